// prism_abi.cu -- the C ABI of libfatiando_b200.so (include/fatiando_b200.h).
//
// Host-side plumbing only: model handles that own the SoA prism arrays in HBM
// ("uploaded once"), staging of host point/result buffers, decomposition of a
// requested field mask into the instantiated component sets, the prism-split
// heuristic for small N, and CUDA-event timing.  All arithmetic lives in the
// kernel translation units.  No entry point has a CPU compute path: if CUDA
// is unusable they fail with FB200_ECUDA.
#include <cuda_runtime.h>
#include <thrust/execution_policy.h>
#include <thrust/sort.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fatiando_b200.h"
#include "abi_internal.h"
#include "prism_fastmath.cuh"
#include "prism_launch.h"
#include "prism_masks.h"
#include "prism_forward.cuh"
#include "prism_mesh.cuh"
#include "prism_mesh_jacobian.cuh"
#include "prism_grid_reuse.cuh"
#include "sphere.cuh"

using namespace fb200;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

// for the other translation units of the library (abi_internal.h)
int fb200::set_error(int code, const std::string &msg) { return fail(code, msg); }

#define CK(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess)                                                     \
            return fail(FB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---------------------------------------------------------------------------
// model handle
// ---------------------------------------------------------------------------
struct fb200_model {
    int device = 0;
    int sm_count = 148;
    bool spheres = false;        // rows hold spheres (sphere.cuh), not prisms
    bool poly_edges = false;     // rows hold polygon edges (polyprism.cuh)
    int64_t m = 0;
    double *d_b[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *d_dens = nullptr;
    double *d_mag[3] = {nullptr, nullptr, nullptr};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // grow-only scratch
    double *d_pts = nullptr;  size_t pts_cap = 0;     // 3*n staged points
    double *d_out = nullptr;  size_t out_cap = 0;     // staged results
    double *d_part = nullptr; size_t part_cap = 0;    // prism-split partials
    // optional node-sharing representation of a regular mesh (prism_mesh.cuh)
    bool has_mesh = false;
    int nxp = 0, nyp = 0, nzp = 0;
    int64_t nnodes = 0;
    double *d_nodes = nullptr;                          // xn | yn | zn
    double *d_wdens = nullptr;                          // node weights of the density
    double *d_wmag[3] = {nullptr, nullptr, nullptr};    // ... of mx, my, mz
    double shift[3] = {0.0, 0.0, 0.0};
    // optional surface form of a relief (prism_face.cuh): top faces + reference-plane nodes
    bool has_surface = false;
    int64_t nfaces = 0, nsnodes = 0;
    double *d_face[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // x1 x2 y1 y2 zf zo w
    double *d_snode[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // xn shx yn shy zn zn w
    int *d_flag = nullptr;
    int last_fallback_points = 0;                      // points the last call re-evaluated per cell
    cudaStream_t copy_stream = nullptr;                // D2H of Jacobian row blocks (overlaps the kernels)
    cudaEvent_t ev_built[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    double *d_iwork = nullptr; size_t iwork_cap = 0;   // per-point flags + index list (ints)
    double *d_sub = nullptr;   size_t sub_cap = 0;     // gathered points and their results (per-point fallback)
    double *d_adj = nullptr; size_t adj_cap = 0;       // node-sharing adjoint: point rows, node list, V
    double *d_grid = nullptr; size_t grid_cap = 0;     // commensurate-grid Jacobian: node and cell tables
    // cell walls whose coordinate the reference computes differently from the two sides
    // (x2 of one cell != x1 of the next, one ulp apart): see fb200_model_set_hazard_planes
    double *d_hazard[3] = {nullptr, nullptr, nullptr};
    int64_t n_hazard[3] = {0, 0, 0};
    double last_ms = 0.0;
    int last_launches = 0;
    std::mutex lock;   // one call at a time per handle
};

// Device memory comes from the stream-ordered allocator (cudaMallocAsync): its
// pool keeps freed blocks cached (release threshold raised in resources_for),
// so a one-shot call -- create model, evaluate, destroy -- pays microseconds,
// not a synchronising cudaMalloc/cudaFree pair per array.
static int grow(double **buf, size_t *cap, size_t need, cudaStream_t s)
{
    if (need <= *cap) return FB200_OK;
    if (*buf) cudaFreeAsync(*buf, s);
    *buf = nullptr;
    *cap = 0;
    CK(cudaMallocAsync((void **)buf, need * sizeof(double), s));
    *cap = need;
    return FB200_OK;
}

// Streams and timing events are pooled per device: creating them costs more
// than evaluating the cookbook-sized models.
struct StreamSet {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
static std::mutex g_pool_lock;
static std::vector<StreamSet> g_pool[64];
static bool g_pool_configured[64] = {false};

static int acquire_streams(int device, StreamSet *out)
{
    {
        std::lock_guard<std::mutex> hold(g_pool_lock);
        if (!g_pool_configured[device]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                uint64_t keep = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            g_pool_configured[device] = true;
        }
        if (!g_pool[device].empty()) {
            *out = g_pool[device].back();
            g_pool[device].pop_back();
            return FB200_OK;
        }
    }
    CK(cudaStreamCreateWithFlags(&out->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&out->ev0));
    CK(cudaEventCreate(&out->ev1));
    return FB200_OK;
}

static void release_streams(int device, const StreamSet &set)
{
    if (!set.stream) return;
    std::lock_guard<std::mutex> hold(g_pool_lock);
    g_pool[device].push_back(set);
}

// ---------------------------------------------------------------------------
// small kernels owned by this file
// ---------------------------------------------------------------------------
// out[out_row[c]*ld_out + l] = scale[c] * sum_s part[(s*nc + c)*ld + l], s ascending:
// the fixed-order second stage of the prism split (reproducible run to run).
struct ReduceArgs {
    const double *part;
    double *out;
    int64_t n, ld, ld_out;
    int nc, nsplit;
    double scale[F_COUNT];
    int out_row[F_COUNT];
};

__global__ void reduce_partials_kernel(const ReduceArgs a)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (l >= a.n) return;
    double s = 0.0;
    for (int k = 0; k < a.nsplit; ++k) s += a.part[((int64_t)k * a.nc + c) * a.ld + l];
    a.out[(int64_t)a.out_row[c] * a.ld_out + l] = s * a.scale[c];
}

// -0.0 -> +0.0 in the model bounds: the fast path reads sign BITS of coordinate
// differences where the reference tests `< 0` (false for -0.0); with canonical
// bounds a difference can only be -0.0 if it is a genuine negative underflow.
__global__ void canonical_zero_kernel(double *p, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] + 0.0;
}

// *flag = 1 (and pflag[l] = 1, when given) for a coordinate c[l] within a few ulp of one of the
// (sorted) planes
__global__ void near_plane_kernel(const double *c, int64_t n, const double *planes, int64_t np, int *flag,
                                  int *pflag = nullptr)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const double v = c[l];
    int64_t lo = 0, hi = np;            // first plane >= v
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (planes[mid] < v) lo = mid + 1; else hi = mid;
    }
    bool near = false;
    for (int64_t k = lo - 1; k <= lo; ++k) {
        if (k < 0 || k >= np) continue;
        const double p = planes[k];
        near = near || fabs(v - p) <= 0x1p-49 * fmax(fabs(v), fabs(p));     // 8 ulp
    }
    if (near) {
        *flag = 1;
        if (pflag) pflag[l] = 1;
    }
}

// ---- per-point fallback of the merged forms ---------------------------------------
// idx[0 .. *count) = the flagged points, ascending (one block-wide ordered pass per 1024 points
// would be overkill: the list is short, an atomic append followed by a sort on the host side is
// not needed either -- every flagged point is evaluated on its own, the order is irrelevant)
__global__ void compact_flags_kernel(const int *pflag, int64_t n, int *idx, int *count)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n && pflag[l]) idx[atomicAdd(count, 1)] = (int)l;
}

__global__ void gather_points_kernel(const double *xp, const double *yp, const double *zp, const int *idx,
                                     int64_t k, int64_t k_pad, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const int l = idx[i];
    out[i] = xp[l]; out[k_pad + i] = yp[l]; out[2 * k_pad + i] = zp[l];
}

__global__ void gather_values_kernel(const double *v, const int *idx, int64_t k, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) out[i] = v[idx[i]];
}

__global__ void add_kernel(double *dst, const double *src, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

// out[c*ld_out + idx[i]] = sub[c*ld_sub + i]
__global__ void scatter_fields_kernel(const double *sub, int64_t ld_sub, const int *idx, int64_t k,
                                      double *out, int64_t ld_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    out[(int64_t)blockIdx.y * ld_out + idx[i]] = sub[(int64_t)blockIdx.y * ld_sub + i];
}

// rows of a Jacobian: out(idx[i] - row0, q) = sub(i, q), either layout; rows outside
// [row0, row0 + nrows) are skipped (host-streamed row blocks)
__global__ void scatter_rows_kernel(const double *sub, int64_t ld_sub, int transposed_sub, const int *idx,
                                    int64_t k, int64_t m, int64_t row0, int64_t nrows, double *out,
                                    int64_t ld_out, int transposed)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (q >= m || i >= k) return;
    const int64_t l = (int64_t)idx[i] - row0;
    if (l < 0 || l >= nrows) return;
    const double v = transposed_sub ? sub[q * ld_sub + i] : sub[i * ld_sub + q];
    if (transposed) out[q * ld_out + l] = v; else out[l * ld_out + q] = v;
}

// ---- node-sharing adjoint helpers (fb200_adjoint) ------------------------------
// rows[r][l]: xpt, shx, ypt, shy, zpt, shz, data*u0 [, data*u1, data*u2]; padding = zero weight
__global__ void adjoint_rows_kernel(const double *xp, const double *yp, const double *zp,
                                    const double *data, int64_t n, int64_t n_pad, int nw, double u0,
                                    double u1, double u2, double shx, double shy, double shz,
                                    double *rows, const int *skip)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_pad) return;
    const bool in = l < n;
    const int64_t s = in ? l : n - 1;
    // points handed to the per-cell kernels (hazard walls) carry no weight here
    const double d = (in && !(skip && skip[s])) ? data[s] : 0.0;
    rows[l] = xp[s];             rows[n_pad + l] = shx;
    rows[2 * n_pad + l] = yp[s]; rows[3 * n_pad + l] = shy;
    rows[4 * n_pad + l] = zp[s]; rows[5 * n_pad + l] = shz;
    rows[6 * n_pad + l] = d * u0;
    if (nw == 3) { rows[7 * n_pad + l] = d * u1; rows[8 * n_pad + l] = d * u2; }
}

// node id = (b*nxp + a)*nzp + c  ->  its coordinates
__global__ void expand_nodes_kernel(const double *xn, const double *yn, const double *zn, int nxp,
                                    int nyp, int nzp, int64_t nn_pad, double *out)
{
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nn = (int64_t)nxp * nyp * nzp;
    if (id >= nn) return;
    const int c = (int)(id % nzp);
    const int64_t col = id / nzp;
    const int a = (int)(col % nxp), b = (int)(col / nxp);
    out[id] = xn[a]; out[nn_pad + id] = yn[b]; out[2 * nn_pad + id] = zn[c];
}

// J^T d of cell q = signed sum of V over its eight corners; corner (x2, y2, z2) carries +,
// every lower bound flips the sign (_prism.pyx:104)
__global__ void adjoint_cells_kernel(const double *V, int nx, int ny, int nz, double scale, double *out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t m = (int64_t)nx * ny * nz;
    if (q >= m) return;
    const int i = (int)(q % nx);
    const int64_t rest = q / nx;
    const int j = (int)(rest % ny), k = (int)(rest / ny);
    const int nxp = nx + 1, nzp = nz + 1;
    auto v = [&](int a, int b, int c) { return V[((int64_t)b * nxp + a) * nzp + c]; };
    const double top = (v(i + 1, j + 1, k + 1) - v(i, j + 1, k + 1)) - (v(i + 1, j, k + 1) - v(i, j, k + 1));
    const double bot = (v(i + 1, j + 1, k) - v(i, j + 1, k)) - (v(i + 1, j, k) - v(i, j, k));
    out[q] = (top - bot) * scale;
}

__global__ void fill_kernel(double *p, int64_t n, double v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// 8 independent DFMA chains per thread, everything in registers.  Operand mix
// register, register, uniform: the fastest of the mixes probed on B200
// (tools/fp64_pipe_probe.cu: 99 % of 148 SM x 64 lanes x clock; register,
// uniform, uniform only reaches 91 %), so this is the true pipe ceiling.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 1e-3 + k; y[k] = a + 1e-12 * (threadIdx.x + k); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = fma(x[k], y[k], b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void debug_math_kernel(int op, const double *a, const double *b, int64_t n, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = 0.0;
    switch (op) {
    case 0: r = fm::log_ratio(a[i], b[i]); break;
    case 1: r = fm::atan_rhp(a[i], b[i]); break;
    case 2: r = fm::div_fast(a[i], b[i]); break;
    case 3: r = fm::root(a[i]); break;
    case 4: r = sqrt(a[i]); break;
    default: break;
    }
    out[i] = r;
}

// ---------------------------------------------------------------------------
// housekeeping
// ---------------------------------------------------------------------------
extern "C" int fb200_abi_version(void) { return FB200_ABI_VERSION; }

extern "C" const char *fb200_last_error(void) { return g_err.c_str(); }

extern "C" int fb200_device_count(int *count)
{
    if (!count) return fail(FB200_EINVAL, "count is NULL");
    *count = 0;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(FB200_ECUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = c;
    if (c == 0) return fail(FB200_ECUDA, "no CUDA device visible");
    return FB200_OK;
}

extern "C" int fb200_device_info(int device, char *name, int name_len, int *sm_count,
                                 int64_t *total_mem_bytes)
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (name && name_len > 0) {
        std::strncpy(name, prop.name, (size_t)name_len - 1);
        name[name_len - 1] = '\0';
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem_bytes) *total_mem_bytes = (int64_t)prop.totalGlobalMem;
    return FB200_OK;
}

// ---------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------
static int upload(double **dst, const double *src, int64_t m, int64_t m_pad, cudaStream_t s)
{
    CK(cudaMallocAsync((void **)dst, (size_t)m_pad * sizeof(double), s));
    CK(cudaMemsetAsync(*dst, 0, (size_t)m_pad * sizeof(double), s));
    if (m > 0) CK(cudaMemcpyAsync(*dst, src, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, s));
    return FB200_OK;
}

extern "C" int fb200_model_destroy(fb200_model *model)
{
    if (!model) return FB200_OK;
    DeviceGuard g(model->device);
    cudaStream_t st = model->stream;
    for (double *p : model->d_b) if (p) cudaFreeAsync(p, st);
    if (model->d_dens) cudaFreeAsync(model->d_dens, st);
    for (double *p : model->d_mag) if (p) cudaFreeAsync(p, st);
    if (model->d_pts) cudaFreeAsync(model->d_pts, st);
    if (model->d_out) cudaFreeAsync(model->d_out, st);
    if (model->d_part) cudaFreeAsync(model->d_part, st);
    if (model->d_nodes) cudaFreeAsync(model->d_nodes, st);
    if (model->d_wdens) cudaFreeAsync(model->d_wdens, st);
    for (double *p : model->d_wmag) if (p) cudaFreeAsync(p, st);
    for (double *p : model->d_face) if (p) cudaFreeAsync(p, st);
    for (int r = 0; r < 7; ++r)
        if (model->d_snode[r] && r != 5) cudaFreeAsync(model->d_snode[r], st);   // row 5 aliases row 4
    if (model->d_flag) cudaFreeAsync(model->d_flag, st);
    if (model->d_iwork) cudaFreeAsync(model->d_iwork, st);
    if (model->d_sub) cudaFreeAsync(model->d_sub, st);
    if (model->d_adj) cudaFreeAsync(model->d_adj, st);
    if (model->d_grid) cudaFreeAsync(model->d_grid, st);
    for (double *p : model->d_hazard) if (p) cudaFreeAsync(p, st);
    if (model->copy_stream) {
        cudaStreamSynchronize(model->copy_stream);
        cudaStreamDestroy(model->copy_stream);
        for (int i = 0; i < 2; ++i) {
            if (model->ev_built[i]) cudaEventDestroy(model->ev_built[i]);
            if (model->ev_copy[i]) cudaEventDestroy(model->ev_copy[i]);
        }
    }
    StreamSet set;
    set.stream = model->stream; set.ev0 = model->ev0; set.ev1 = model->ev1;
    release_streams(model->device, set);
    delete model;
    return FB200_OK;
}

extern "C" int fb200_model_create(int device, int64_t m, const double *x1, const double *x2,
                                  const double *y1, const double *y2, const double *z1,
                                  const double *z2, const double *density, const double *mx,
                                  const double *my, const double *mz, fb200_model **out)
{
    if (!out) return fail(FB200_EINVAL, "out is NULL");
    *out = nullptr;
    if (m < 0) return fail(FB200_EINVAL, "negative prism count");
    if (m > 0 && (!x1 || !x2 || !y1 || !y2 || !z1 || !z2))
        return fail(FB200_EINVAL, "prism bound array is NULL");
    const int nmag = (mx != nullptr) + (my != nullptr) + (mz != nullptr);
    if (nmag != 0 && nmag != 3)
        return fail(FB200_EINVAL, "mx, my, mz must be all NULL or all given");
    int count = 0;
    int rc = fb200_device_count(&count);
    if (rc != FB200_OK) return rc;
    if (device < 0 || device >= count) return fail(FB200_EINVAL, "device index out of range");
    DeviceGuard g(device);
    if (!g.ok) return fail(FB200_ECUDA, "cudaSetDevice failed");

    fb200_model *mod = new fb200_model();
    mod->device = device;
    mod->m = m;
    cudaDeviceGetAttribute(&mod->sm_count, cudaDevAttrMultiProcessorCount, device);
    const int64_t m_pad = std::max<int64_t>(TILE, (m + TILE - 1) / TILE * TILE);
    const double *src[6] = {x1, x2, y1, y2, z1, z2};
#define CKM(expr)                              \
    do {                                       \
        int rc_ = (expr);                      \
        if (rc_ != FB200_OK) {                 \
            std::string keep = g_err;          \
            fb200_model_destroy(mod);          \
            g_err = keep;                      \
            return rc_;                        \
        }                                      \
    } while (0)
    auto cuda_rc = [](cudaError_t e, const char *what) {
        return e == cudaSuccess ? FB200_OK : fail(FB200_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    };
    {
        StreamSet set;
        CKM(acquire_streams(device, &set));
        mod->stream = set.stream; mod->ev0 = set.ev0; mod->ev1 = set.ev1;
    }
    for (int r = 0; r < 6; ++r) {
        CKM(upload(&mod->d_b[r], src[r], m, m_pad, mod->stream));
        canonical_zero_kernel<<<(unsigned)((m_pad + 255) / 256), 256, 0, mod->stream>>>(mod->d_b[r], m_pad);
    }
    if (density) CKM(upload(&mod->d_dens, density, m, m_pad, mod->stream));
    if (nmag == 3) {
        CKM(upload(&mod->d_mag[0], mx, m, m_pad, mod->stream));
        CKM(upload(&mod->d_mag[1], my, m, m_pad, mod->stream));
        CKM(upload(&mod->d_mag[2], mz, m, m_pad, mod->stream));
    }
    CKM(cuda_rc(cudaStreamSynchronize(mod->stream), "cudaStreamSynchronize"));
#undef CKM
    *out = mod;
    return FB200_OK;
}

extern "C" int fb200_model_create_spheres(int device, int64_t m, const double *xc, const double *yc,
                                          const double *zc, const double *radius,
                                          const double *density, const double *mx,
                                          const double *my, const double *mz, fb200_model **out)
{
    if (m < 0 || (m > 0 && (!xc || !yc || !zc || !radius)))
        return fail(FB200_EINVAL, "fb200_model_create_spheres: NULL array or negative size");
    // volume = 4*pi*radius**3/3 with the reference's association (sphere.py:348)
    std::vector<double> vol((size_t)m), zero((size_t)m, 0.0);
    for (int64_t q = 0; q < m; ++q) vol[(size_t)q] = 4 * 3.141592653589793 * std::pow(radius[q], 3.0) / 3;
    int rc = fb200_model_create(device, m, xc, vol.data(), yc, zero.data(), zc, zero.data(), density,
                                mx, my, mz, out);
    if (rc != FB200_OK) return rc;
    (*out)->spheres = true;
    return FB200_OK;
}

extern "C" int fb200_model_create_polyprisms(int device, int64_t nprisms, const int64_t *first_vertex,
                                             const double *vx, const double *vy, const double *z1,
                                             const double *z2, const double *density,
                                             const double *mx, const double *my, const double *mz,
                                             fb200_model **out)
{
    if (nprisms < 0 || (nprisms > 0 && (!first_vertex || !vx || !vy || !z1 || !z2)))
        return fail(FB200_EINVAL, "fb200_model_create_polyprisms: NULL array or negative size");
    const int nmag = (mx != nullptr) + (my != nullptr) + (mz != nullptr);
    if (nmag != 0 && nmag != 3) return fail(FB200_EINVAL, "mx, my, mz must be all NULL or all given");
    // one row per polygon edge: vertex k -> vertex (k+1) % nverts (polyprism.py:264-269)
    std::vector<double> row[6], dens, mag[3];
    for (int64_t q = 0; q < nprisms; ++q) {
        const int64_t v0 = first_vertex[q], nv = first_vertex[q + 1] - v0;
        if (nv < 1) return fail(FB200_EINVAL, "polygon without vertices");
        for (int64_t k = 0; k < nv; ++k) {
            const int64_t a = v0 + k, b = v0 + (k + 1) % nv;
            row[0].push_back(vx[a]); row[1].push_back(vx[b]);
            row[2].push_back(vy[a]); row[3].push_back(vy[b]);
            row[4].push_back(z1[q]); row[5].push_back(z2[q]);
            if (density) dens.push_back(density[q]);
            if (nmag == 3) { mag[0].push_back(mx[q]); mag[1].push_back(my[q]); mag[2].push_back(mz[q]); }
        }
    }
    const int64_t m = (int64_t)row[0].size();
    int rc = fb200_model_create(device, m, row[0].data(), row[1].data(), row[2].data(), row[3].data(),
                                row[4].data(), row[5].data(), density ? dens.data() : nullptr,
                                nmag == 3 ? mag[0].data() : nullptr, nmag == 3 ? mag[1].data() : nullptr,
                                nmag == 3 ? mag[2].data() : nullptr, out);
    if (rc != FB200_OK) return rc;
    (*out)->poly_edges = true;
    return FB200_OK;
}

extern "C" int fb200_model_attach_mesh(fb200_model *mod, int nx, int ny, int nz, const double *xn,
                                       const double *yn, const double *zn,
                                       const double *cell_size, const double *w_density,
                                       const double *w_mx, const double *w_my, const double *w_mz)
{
    if (!mod) return fail(FB200_EINVAL, "model is NULL");
    if (nx < 1 || ny < 1 || nz < 1) return fail(FB200_EINVAL, "mesh shape must be positive");
    if (!xn || !yn || !zn || !cell_size) return fail(FB200_EINVAL, "NULL mesh array");
    const int nmag = (w_mx != nullptr) + (w_my != nullptr) + (w_mz != nullptr);
    if (nmag != 0 && nmag != 3) return fail(FB200_EINVAL, "w_mx, w_my, w_mz must be all NULL or all given");
    // without node weights the mesh geometry alone is attached: enough for the Jacobian
    std::lock_guard<std::mutex> hold(mod->lock);
    DeviceGuard g(mod->device);
    if (!g.ok) return fail(FB200_ECUDA, "cudaSetDevice failed");
    if (mod->has_mesh) return fail(FB200_EINVAL, "a mesh is already attached to this model");
    mod->nxp = nx + 1; mod->nyp = ny + 1; mod->nzp = nz + 1;
    mod->nnodes = (int64_t)mod->nxp * mod->nyp * mod->nzp;
    const int64_t n_pad = (mod->nnodes + TILE - 1) / TILE * TILE;
    const int64_t ncoord = (int64_t)mod->nxp + mod->nyp + mod->nzp;
    CK(cudaMallocAsync((void **)&mod->d_nodes, (size_t)ncoord * 8, mod->stream));
    CK(cudaMemcpyAsync(mod->d_nodes, xn, (size_t)mod->nxp * 8, cudaMemcpyHostToDevice, mod->stream));
    CK(cudaMemcpyAsync(mod->d_nodes + mod->nxp, yn, (size_t)mod->nyp * 8, cudaMemcpyHostToDevice, mod->stream));
    CK(cudaMemcpyAsync(mod->d_nodes + mod->nxp + mod->nyp, zn, (size_t)mod->nzp * 8,
                       cudaMemcpyHostToDevice, mod->stream));
    canonical_zero_kernel<<<(unsigned)((ncoord + 255) / 256), 256, 0, mod->stream>>>(mod->d_nodes, ncoord);
    int rc = FB200_OK;
    if (w_density) rc = upload(&mod->d_wdens, w_density, mod->nnodes, n_pad, mod->stream);
    const double *wm[3] = {w_mx, w_my, w_mz};
    for (int r = 0; r < 3 && rc == FB200_OK && nmag == 3; ++r)
        rc = upload(&mod->d_wmag[r], wm[r], mod->nnodes, n_pad, mod->stream);
    if (rc != FB200_OK) return rc;
    for (int r = 0; r < 3; ++r) mod->shift[r] = 0.00001 * cell_size[r];   // _prism.pyx:328-329
    CK(cudaStreamSynchronize(mod->stream));
    mod->has_mesh = true;
    return FB200_OK;
}

extern "C" int fb200_model_attach_surface(fb200_model *mod, int64_t nfaces, const double *x1,
                                          const double *x2, const double *y1, const double *y2,
                                          const double *z_face, const double *z_other,
                                          const double *w_face, int64_t nnodes, const double *xn,
                                          const double *yn, const double *zn, const double *w_node,
                                          const double *cell_size)
{
    if (!mod) return fail(FB200_EINVAL, "model is NULL");
    if (nfaces < 1 || nnodes < 0) return fail(FB200_EINVAL, "surface needs at least one face");
    if (!x1 || !x2 || !y1 || !y2 || !z_face || !z_other || !w_face || !cell_size)
        return fail(FB200_EINVAL, "NULL face array");
    if (nnodes > 0 && (!xn || !yn || !zn || !w_node)) return fail(FB200_EINVAL, "NULL node array");
    std::lock_guard<std::mutex> hold(mod->lock);
    DeviceGuard g(mod->device);
    if (!g.ok) return fail(FB200_ECUDA, "cudaSetDevice failed");
    if (mod->has_surface) return fail(FB200_EINVAL, "a surface is already attached to this model");
    const int64_t f_pad = (nfaces + TILE - 1) / TILE * TILE;
    const double *face_src[7] = {x1, x2, y1, y2, z_face, z_other, w_face};
    for (int r = 0; r < 7; ++r) {
        int rc = upload(&mod->d_face[r], face_src[r], nfaces, f_pad, mod->stream);
        if (rc != FB200_OK) return rc;
        if (r < 6) canonical_zero_kernel<<<(unsigned)((nfaces + 255) / 256), 256, 0, mod->stream>>>(mod->d_face[r], nfaces);
    }
    if (nnodes > 0) {
        const int64_t n_pad = (nnodes + TILE - 1) / TILE * TILE;
        std::vector<double> shx((size_t)nnodes, 0.00001 * cell_size[0]);   // _prism.pyx:328-329
        std::vector<double> shy((size_t)nnodes, 0.00001 * cell_size[1]);
        const double *node_src[7] = {xn, shx.data(), yn, shy.data(), zn, nullptr, w_node};
        for (int r = 0; r < 7; ++r) {
            if (r == 5) { mod->d_snode[5] = mod->d_snode[4]; continue; }
            int rc = upload(&mod->d_snode[r], node_src[r], nnodes, n_pad, mod->stream);
            if (rc != FB200_OK) return rc;
            if (r == 0 || r == 2 || r == 4)
                canonical_zero_kernel<<<(unsigned)((nnodes + 255) / 256), 256, 0, mod->stream>>>(mod->d_snode[r], nnodes);
        }
        CK(cudaStreamSynchronize(mod->stream));      // shx / shy are locals
    }
    if (!mod->d_flag) CK(cudaMallocAsync((void **)&mod->d_flag, sizeof(int), mod->stream));
    CK(cudaMemsetAsync(mod->d_flag, 0, sizeof(int), mod->stream));
    CK(cudaStreamSynchronize(mod->stream));
    mod->nfaces = nfaces;
    mod->nsnodes = nnodes;
    mod->has_surface = true;
    return FB200_OK;
}

extern "C" int fb200_model_set_hazard_planes(fb200_model *mod, int64_t nx, const double *xw,
                                             int64_t ny, const double *yw, int64_t nz,
                                             const double *zw)
{
    if (!mod) return fail(FB200_EINVAL, "model is NULL");
    const int64_t cnt[3] = {nx, ny, nz};
    const double *src[3] = {xw, yw, zw};
    for (int a = 0; a < 3; ++a)
        if (cnt[a] < 0 || (cnt[a] > 0 && !src[a])) return fail(FB200_EINVAL, "bad hazard plane array");
    std::lock_guard<std::mutex> hold(mod->lock);
    DeviceGuard g(mod->device);
    if (!g.ok) return fail(FB200_ECUDA, "cudaSetDevice failed");
    for (int a = 0; a < 3; ++a) {
        if (mod->d_hazard[a]) { cudaFreeAsync(mod->d_hazard[a], mod->stream); mod->d_hazard[a] = nullptr; }
        mod->n_hazard[a] = 0;
        if (cnt[a] == 0) continue;
        std::vector<double> sorted(src[a], src[a] + cnt[a]);
        std::sort(sorted.begin(), sorted.end());
        CK(cudaMallocAsync((void **)&mod->d_hazard[a], (size_t)cnt[a] * 8, mod->stream));
        CK(cudaMemcpyAsync(mod->d_hazard[a], sorted.data(), (size_t)cnt[a] * 8, cudaMemcpyHostToDevice, mod->stream));
        CK(cudaStreamSynchronize(mod->stream));
        mod->n_hazard[a] = cnt[a];
    }
    if (!mod->d_flag) {
        CK(cudaMallocAsync((void **)&mod->d_flag, sizeof(int), mod->stream));
        CK(cudaMemsetAsync(mod->d_flag, 0, sizeof(int), mod->stream));
        CK(cudaStreamSynchronize(mod->stream));
    }
    return FB200_OK;
}

extern "C" int64_t fb200_model_size(const fb200_model *model) { return model ? model->m : -1; }
extern "C" int fb200_model_device(const fb200_model *model) { return model ? model->device : -1; }

// Page-locked host memory for destinations that are filled again and again (a host-resident
// Jacobian crosses PCIe at DMA speed into it; pageable memory goes through bounce buffers).
extern "C" int fb200_host_alloc(size_t bytes, void **ptr)
{
    if (!ptr) return fail(FB200_EINVAL, "ptr is NULL");
    *ptr = nullptr;
    CK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return FB200_OK;
}

extern "C" int fb200_host_free(void *ptr)
{
    if (ptr) CK(cudaFreeHost(ptr));
    return FB200_OK;
}

extern "C" int fb200_last_fallback_points(const fb200_model *model, int *points)
{
    if (!model || !points) return fail(FB200_EINVAL, "NULL argument");
    *points = model->last_fallback_points;
    return FB200_OK;
}

extern "C" int fb200_last_kernel_ms(const fb200_model *model, double *ms, int *launches)
{
    if (!model) return fail(FB200_EINVAL, "model is NULL");
    if (ms) *ms = model->last_ms;
    if (launches) *launches = model->last_launches;
    return FB200_OK;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
static const uint32_t kGravitySets[] = {
#define X(M) (M),
    FB200_GRAVITY_SETS(X)
#undef X
};
// spheres: gz and the tensor only (sphere.cuh)
static const uint32_t kSphereGravitySets[] = {0x008, 0x010, 0x020, 0x040, 0x080, 0x100, 0x200, 0x3F0, 0x3F8};
static const uint32_t kMagneticSets[] = {
#define X(M) (M),
    FB200_MAGNETIC_SETS(X)
#undef X
};

// Greedy cover of `want` by instantiated sets: the largest subset first.
template <size_t N>
static std::vector<uint32_t> cover(uint32_t want, const uint32_t (&sets)[N])
{
    std::vector<uint32_t> res;
    while (want) {
        uint32_t best = 0;
        for (uint32_t s : sets)
            if ((s & ~want) == 0 && popc(s) > popc(best)) best = s;
        res.push_back(best);
        want &= ~best;
    }
    return res;
}

static void view_of(const fb200_model *mod, bool magnetic, ModelView *v)
{
    for (int r = 0; r < 6; ++r) v->b[r] = mod->d_b[r];
    if (magnetic) {
        for (int r = 0; r < 3; ++r) v->p[r] = mod->d_mag[r];
    } else {
        v->p[0] = mod->d_dens;
        v->p[1] = v->p[2] = nullptr;
    }
    v->m = mod->m;
}

// Thread blocks per SM the prism / node split aims for.  The kernels keep 3-4 blocks resident
// per SM, so 96 blocks per SM are ~24-32 waves: the last, partly filled wave costs 1-2 % of the
// launch instead of the 4-6 % it cost at 20 blocks per SM (5.3 waves on C2), for 47 instead of
// 10 partial sums per point.  Measured (tools/split_sweep.py, profiles/r02_split_sweep.txt):
// C2 18.48 -> 17.52 ms, C3 254.6 -> 246.1 ms, C5 gz+tensor 353.0 -> 348.9 ms; 128 is no better.
// FB200_SPLIT_TARGET overrides it (tuning).
static int64_t split_blocks_per_sm()
{
    const char *e = std::getenv("FB200_SPLIT_TARGET");
    const long v = e ? std::atol(e) : 0;
    return v > 0 ? v : 96;
}

static int reduce_split(fb200_model *mod, int nc, int64_t nsplit, int64_t n, int64_t ld_p,
                        const double *scale, const int *out_row, double *d_out, int64_t ld_out)
{
    ReduceArgs r;
    std::memset(&r, 0, sizeof(r));
    r.part = mod->d_part; r.out = d_out; r.n = n; r.ld = ld_p; r.ld_out = ld_out;
    r.nc = nc; r.nsplit = (int)nsplit;
    for (int k = 0; k < nc; ++k) { r.scale[k] = scale[k]; r.out_row[k] = out_row[k]; }
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)nc);
    reduce_partials_kernel<<<grid, 256, 0, mod->stream>>>(r);
    CK(cudaGetLastError());
    return FB200_OK;
}

// Node-sharing evaluation of one component set (regular mesh attached, no override).
static int run_set_mesh(fb200_model *mod, uint32_t set, uint32_t request, bool allow_split,
                        const double *d_xp, const double *d_yp, const double *d_zp, int64_t n,
                        const double *fvec, const double *scale, double *d_out, int64_t ld_out)
{
    const bool magnetic = (set & MAGNETIC_MASK) != 0;
    MeshArgs a;
    std::memset(&a, 0, sizeof(a));
    a.mesh.xn = mod->d_nodes;
    a.mesh.yn = mod->d_nodes + mod->nxp;
    a.mesh.zn = mod->d_nodes + mod->nxp + mod->nyp;
    a.mesh.nxp = mod->nxp; a.mesh.nyp = mod->nyp; a.mesh.nzp = mod->nzp;
    a.mesh.nnodes = mod->nnodes;
    if (magnetic) for (int r = 0; r < 3; ++r) a.mesh.w[r] = mod->d_wmag[r];
    else a.mesh.w[0] = mod->d_wdens;
    for (int r = 0; r < 3; ++r) { a.mesh.sh[r] = mod->shift[r]; a.fvec[r] = fvec ? fvec[r] : 0.0; }
    a.xp = d_xp; a.yp = d_yp; a.zp = d_zp; a.n = n;
    const int nc = popc(set);
    int c = 0;
    for (int f = 0; f < F_COUNT; ++f) {
        if (!(set & bit(f))) continue;
        const int row = popc(request & (bit(f) - 1));
        a.out_row[c] = row;
        a.scale[c] = scale ? scale[row] : 1.0;
        ++c;
    }
    const int64_t blocks_x = (n + FWD_THREADS - 1) / FWD_THREADS;
    const int64_t tiles = (mod->nnodes + TILE - 1) / TILE;
    int64_t nsplit = 1;
    if (allow_split) {
        const int64_t target = (int64_t)mod->sm_count * split_blocks_per_sm();
        nsplit = std::min<int64_t>(std::max<int64_t>(1, (target + blocks_x - 1) / blocks_x), tiles);
        nsplit = std::min<int64_t>(nsplit, 65535);
    }
    const int64_t chunk = ((tiles + nsplit - 1) / nsplit) * TILE;
    nsplit = (mod->nnodes + chunk - 1) / chunk;
    a.chunk = chunk;
    a.nsplit = (int)nsplit;
    if (nsplit == 1) {
        a.out = d_out;
        a.ld = ld_out;
        CK(launch_mesh_forward(set, a, mod->stream));
        mod->last_launches += 1;
        return FB200_OK;
    }
    const int64_t ld_p = (n + 31) / 32 * 32;
    int rc = grow(&mod->d_part, &mod->part_cap, (size_t)nsplit * nc * ld_p, mod->stream);
    if (rc != FB200_OK) return rc;
    a.out = mod->d_part;
    a.ld = ld_p;
    CK(launch_mesh_forward(set, a, mod->stream));
    rc = reduce_split(mod, nc, nsplit, n, ld_p, a.scale, a.out_row, d_out, ld_out);
    if (rc != FB200_OK) return rc;
    mod->last_launches += 2;
    return FB200_OK;
}

// Surface form of a relief (gravity sets): top faces and reference-plane nodes write
// partial sums side by side, one fixed-order reduction adds them.
static int run_set_surface(fb200_model *mod, uint32_t set, uint32_t request, bool allow_split,
                           const double *d_xp, const double *d_yp, const double *d_zp, int64_t n,
                           const double *scale, double *d_out, int64_t ld_out, int *d_pflag)
{
    ForwardArgs a;
    std::memset(&a, 0, sizeof(a));
    a.xp = d_xp; a.yp = d_yp; a.zp = d_zp; a.n = n;
    a.flag = d_pflag;
    const int nc = popc(set);
    int c = 0;
    for (int f = 0; f < F_COUNT; ++f) {
        if (!(set & bit(f))) continue;
        const int row = popc(request & (bit(f) - 1));
        a.out_row[c] = row;
        a.scale[c] = scale ? scale[row] : 1.0;
        ++c;
    }
    const int64_t blocks_x = (n + FWD_THREADS - 1) / FWD_THREADS;
    auto split_of = [&](int64_t m, int64_t *chunk) {
        const int64_t tiles = (m + TILE - 1) / TILE;
        int64_t ns = 1;
        if (allow_split) {
            const int64_t target = (int64_t)mod->sm_count * split_blocks_per_sm();
            ns = std::min<int64_t>(std::max<int64_t>(1, (target + blocks_x - 1) / blocks_x), tiles);
            ns = std::min<int64_t>(ns, 32000);
        }
        *chunk = ((tiles + ns - 1) / ns) * TILE;
        return (m + *chunk - 1) / *chunk;
    };
    int64_t chunk_f = TILE, chunk_n = TILE;
    const int64_t ns_f = split_of(mod->nfaces, &chunk_f);
    const int64_t ns_n = mod->nsnodes > 0 ? split_of(mod->nsnodes, &chunk_n) : 0;
    const int64_t ns = ns_f + ns_n;
    const int64_t ld_p = (n + 31) / 32 * 32;
    int rc = grow(&mod->d_part, &mod->part_cap, (size_t)ns * nc * ld_p, mod->stream);
    if (rc != FB200_OK) return rc;
    a.ld = ld_p;
    a.nsplit = (int)std::max<int64_t>(ns, 2);          // always the partial-sum output mode
    for (int r = 0; r < 6; ++r) a.model.b[r] = mod->d_face[r];
    a.model.p[0] = mod->d_face[6];
    a.model.m = mod->nfaces;
    a.chunk = chunk_f;
    a.out = mod->d_part;
    CK(launch_forward_face(set, a, (int)ns_f, mod->stream));
    mod->last_launches += 1;
    if (ns_n > 0) {
        for (int r = 0; r < 6; ++r) a.model.b[r] = mod->d_snode[r];
        a.model.p[0] = mod->d_snode[6];
        a.model.m = mod->nsnodes;
        a.chunk = chunk_n;
        a.out = mod->d_part + (size_t)ns_f * nc * ld_p;
        CK(launch_forward_node(set, a, (int)ns_n, mod->stream));
        mod->last_launches += 1;
    }
    rc = reduce_split(mod, nc, ns, n, ld_p, a.scale, a.out_row, d_out, ld_out);
    if (rc != FB200_OK) return rc;
    mod->last_launches += 1;
    return FB200_OK;
}

// One instantiated component set over all points; handles the prism split.
static int run_set(fb200_model *mod, uint32_t set, uint32_t request, bool exact, bool allow_split,
                   const double *d_xp, const double *d_yp, const double *d_zp, int64_t n,
                   const double *override_p, const double *fvec, const double *scale,
                   const double *init, double *d_out, int64_t ld_out)
{
    const bool magnetic = (set & MAGNETIC_MASK) != 0;
    ForwardArgs a;
    std::memset(&a, 0, sizeof(a));
    view_of(mod, magnetic, &a.model);
    a.xp = d_xp; a.yp = d_yp; a.zp = d_zp; a.n = n;
    a.use_override = override_p != nullptr;
    for (int r = 0; r < 3; ++r) a.override_p[r] = override_p ? override_p[magnetic ? r : 0] : 0.0;
    for (int r = 0; r < 3; ++r) a.fvec[r] = fvec ? fvec[r] : 0.0;
    a.init = init;
    const int nc = popc(set);
    int c = 0;
    for (int f = 0; f < F_COUNT; ++f) {
        if (!(set & bit(f))) continue;
        const int row = popc(request & (bit(f) - 1));   // position among the requested fields
        a.out_row[c] = row;
        a.scale[c] = scale ? scale[row] : 1.0;
        ++c;
    }
    // prism split: keep every SM busy when there are few points
    const int64_t blocks_x = (n + FWD_THREADS - 1) / FWD_THREADS;
    const int64_t tiles = (mod->m + TILE - 1) / TILE;
    int64_t nsplit = 1;
    if (allow_split && init == nullptr) {
        const int64_t target = (int64_t)mod->sm_count * split_blocks_per_sm();
        nsplit = std::min<int64_t>(std::max<int64_t>(1, (target + blocks_x - 1) / blocks_x), tiles);
        nsplit = std::min<int64_t>(nsplit, 65535);
    }
    int64_t chunk = ((tiles + nsplit - 1) / nsplit) * TILE;
    nsplit = (mod->m + chunk - 1) / chunk;
    a.chunk = chunk;
    a.nsplit = (int)nsplit;
    auto launch = [&](const ForwardArgs &args) -> cudaError_t {
        if (mod->spheres) return launch_forward_sphere(set, args, mod->stream);
        if (mod->poly_edges) return launch_forward_polyprism(set, args, mod->stream);
        if (exact) return launch_forward_exact(set, args, mod->stream);
        return magnetic ? launch_forward_fast_mag(set, args, mod->stream)
                        : launch_forward_fast_grav(set, args, mod->stream);
    };
    if (nsplit == 1) {
        a.out = d_out;
        a.ld = ld_out;
        CK(launch(a));
        mod->last_launches += 1;
        return FB200_OK;
    }
    const int64_t ld_p = (n + 31) / 32 * 32;
    int rc = grow(&mod->d_part, &mod->part_cap, (size_t)nsplit * nc * ld_p, mod->stream);
    if (rc != FB200_OK) return rc;
    a.out = mod->d_part;
    a.ld = ld_p;
    CK(launch(a));
    rc = reduce_split(mod, nc, nsplit, n, ld_p, a.scale, a.out_row, d_out, ld_out);
    if (rc != FB200_OK) return rc;
    mod->last_launches += 2;
    return FB200_OK;
}

static int forward_impl(fb200_model *mod, const double *xp, const double *yp, const double *zp,
                        int64_t n, uint32_t field_mask, const double *dens_override,
                        const double *pmag_override, const double *fvec, const double *scale,
                        const double *init, double *out, int64_t ld_out, uint32_t flags)
{
    if (!mod) return fail(FB200_EINVAL, "model is NULL");
    if (n < 0) return fail(FB200_EINVAL, "negative point count");
    if (field_mask == 0 || (field_mask & ~(GRAVITY_MASK | MAGNETIC_MASK)))
        return fail(FB200_EINVAL, "invalid field mask");
    if (n > 0 && (!xp || !yp || !zp || !out)) return fail(FB200_EINVAL, "NULL point or output array");
    if (ld_out < n) return fail(FB200_EINVAL, "ld_out smaller than n");
    if (flags & ~(FB200_DEVICE_POINTERS | FB200_REFERENCE_ORDER | FB200_NO_PRISM_SPLIT | FB200_PER_CELL))
        return fail(FB200_EUNSUP, "unsupported flag for fb200_forward");
    const uint32_t gmask = field_mask & GRAVITY_MASK, mmask = field_mask & MAGNETIC_MASK;
    if (gmask && !mod->d_dens && !dens_override && !(mod->has_mesh && mod->d_wdens))
        return fail(FB200_EPROP, "gravity field requested but the model holds no density");
    if (mmask && !mod->d_mag[0] && !pmag_override && !(mod->has_mesh && mod->d_wmag[0]))
        return fail(FB200_EPROP, "magnetic field requested but the model holds no magnetization");
    if ((mmask & bit(F_TF)) && !fvec)
        return fail(FB200_EINVAL, "tf requested without the inducing-field direction");
    const bool sibling = mod->spheres || mod->poly_edges;
    if (sibling && (field_mask & ~SPHERE_FIELDS))
        return fail(FB200_EUNSUP, "spheres and polygonal prisms offer gz, gxx..gzz, tf, bx, by, bz only");
    if (sibling && init) return fail(FB200_EUNSUP, "no accumulate entry points for spheres / polygonal prisms");
    const bool on_device = (flags & FB200_DEVICE_POINTERS) != 0;
    const bool exact = (flags & FB200_REFERENCE_ORDER) != 0;
    const bool split = !exact && !(flags & FB200_NO_PRISM_SPLIT);
    const int nreq = popc(field_mask);
    if (n == 0) return FB200_OK;

    std::lock_guard<std::mutex> hold(mod->lock);
    DeviceGuard g(mod->device);
    if (!g.ok) return fail(FB200_ECUDA, "cudaSetDevice failed");
    mod->last_launches = 0;
    mod->last_ms = 0.0;
    mod->last_fallback_points = 0;

    const double *d_xp = xp, *d_yp = yp, *d_zp = zp, *d_init = init;
    double *d_out = out;
    int64_t ld = ld_out;
    if (!on_device) {
        const int64_t np = (n + 31) / 32 * 32;
        int rc = grow(&mod->d_pts, &mod->pts_cap, (size_t)np * 3, mod->stream);
        if (rc != FB200_OK) return rc;
        rc = grow(&mod->d_out, &mod->out_cap, (size_t)np * nreq, mod->stream);
        if (rc != FB200_OK) return rc;
        CK(cudaMemcpyAsync(mod->d_pts, xp, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
        CK(cudaMemcpyAsync(mod->d_pts + np, yp, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
        CK(cudaMemcpyAsync(mod->d_pts + 2 * np, zp, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
        d_xp = mod->d_pts; d_yp = mod->d_pts + np; d_zp = mod->d_pts + 2 * np;
        d_out = mod->d_out;
        ld = np;
        if (init) {   // running sum starts from the caller's res (single field)
            CK(cudaMemcpyAsync(d_out, init, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
            d_init = d_out;
        }
    }
    if (mod->m == 0) {   // empty model: zeros (times scale), or init unchanged
        if (!init) {
            for (int c = 0; c < nreq; ++c) {
                fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, mod->stream>>>(d_out + (int64_t)c * ld, n, 0.0);
                CK(cudaGetLastError());
            }
        }
    } else {
        CK(cudaEventRecord(mod->ev0, mod->stream));
        // a regular mesh with node weights attached: node-sharing kernel, unless the
        // caller overrides the property, wants the reference order or the per-cell path
        bool nodes_ok = mod->has_mesh && !exact && !init && !(flags & FB200_PER_CELL) &&
                        mod->nzp <= MESH_ZMAX;
        // a relief with its surface form attached: same conditions, gravity only
        bool surface = mod->has_surface && gmask && !exact && !init && !(flags & FB200_PER_CELL) &&
                       !dens_override;
        // Where the reference's two cells disagree on a shared wall by an ulp, a point IN that
        // wall gets direction-dependent terms (atan2 of +-0 denominators) that do not cancel in
        // the reference; the merged forms cannot reproduce that, the per-cell kernels do.  Such
        // points -- and those where a merged reference-plane node meets the thickness-dependent
        // shift of gxz / gyz (prism_face.cuh: EvalNode) -- are flagged one by one, evaluated
        // by the per-cell kernels afterwards and scattered over the merged forms' output.
        const bool hazards = mod->n_hazard[0] || mod->n_hazard[1] || mod->n_hazard[2];
        int *d_pflag = nullptr, *d_idx = nullptr;
        const int64_t n_pad = (n + 63) / 64 * 64;
        if ((nodes_ok && hazards) || surface) {
            int rc = grow(&mod->d_iwork, &mod->iwork_cap, (size_t)n_pad + 8, mod->stream);   // 2 n_pad ints
            if (rc != FB200_OK) return rc;
            d_pflag = reinterpret_cast<int *>(mod->d_iwork);
            d_idx = d_pflag + n_pad;
            CK(cudaMemsetAsync(d_pflag, 0, (size_t)n * sizeof(int), mod->stream));
            CK(cudaMemsetAsync(mod->d_flag, 0, sizeof(int), mod->stream));
            if (hazards) {
                const double *coord[3] = {d_xp, d_yp, d_zp};
                for (int ax = 0; ax < 3; ++ax) {
                    if (!mod->n_hazard[ax]) continue;
                    near_plane_kernel<<<(unsigned)((n + 255) / 256), 256, 0, mod->stream>>>(
                        coord[ax], n, mod->d_hazard[ax], mod->n_hazard[ax], mod->d_flag, d_pflag);
                    CK(cudaGetLastError());
                }
            }
        }
        if (gmask) {
            const bool use_nodes = nodes_ok && mod->d_wdens && !dens_override;
            for (uint32_t set : sibling ? cover(gmask, kSphereGravitySets) : cover(gmask, kGravitySets)) {
                int rc = surface
                    ? run_set_surface(mod, set, field_mask, split, d_xp, d_yp, d_zp, n, scale, d_out, ld, d_pflag)
                    : use_nodes
                    ? run_set_mesh(mod, set, field_mask, split, d_xp, d_yp, d_zp, n, fvec, scale, d_out, ld)
                    : run_set(mod, set, field_mask, exact, split, d_xp, d_yp, d_zp, n,
                              dens_override, fvec, scale, d_init, d_out, ld);
                if (rc != FB200_OK) return rc;
            }
        }
        if (mmask) {
            const bool use_nodes = nodes_ok && mod->d_wmag[0] && !pmag_override;
            for (uint32_t set : cover(mmask, kMagneticSets)) {
                int rc = use_nodes
                    ? run_set_mesh(mod, set, field_mask, split, d_xp, d_yp, d_zp, n, fvec, scale, d_out, ld)
                    : run_set(mod, set, field_mask, exact, split, d_xp, d_yp, d_zp, n,
                              pmag_override, fvec, scale, d_init, d_out, ld);
                if (rc != FB200_OK) return rc;
            }
        }
        if (d_pflag) {
            // the flagged points, cell by cell: exactly what a per-cell call on those points
            // alone returns (the prism split, hence the last bits, depends on the point count)
            int *d_count = mod->d_flag;
            CK(cudaMemsetAsync(d_count, 0, sizeof(int), mod->stream));
            compact_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, mod->stream>>>(d_pflag, n, d_idx, d_count);
            CK(cudaGetLastError());
            int k = 0;
            CK(cudaMemcpyAsync(&k, d_count, sizeof(int), cudaMemcpyDeviceToHost, mod->stream));
            CK(cudaStreamSynchronize(mod->stream));
            mod->last_fallback_points = k;
            if (k > 0) {
                const int64_t k_pad = ((int64_t)k + 31) / 32 * 32;
                int rc = grow(&mod->d_sub, &mod->sub_cap, (size_t)k_pad * (3 + nreq), mod->stream);
                if (rc != FB200_OK) return rc;
                double *sx = mod->d_sub, *sout = mod->d_sub + 3 * k_pad;
                gather_points_kernel<<<(unsigned)((k + 255) / 256), 256, 0, mod->stream>>>(
                    d_xp, d_yp, d_zp, d_idx, k, k_pad, sx);
                CK(cudaGetLastError());
                const bool merged_g = surface || (nodes_ok && mod->d_wdens && !dens_override);
                const bool merged_m = nodes_ok && mod->d_wmag[0] && !pmag_override;
                uint32_t redo = 0;
                if (gmask && merged_g) {
                    redo |= gmask;
                    for (uint32_t set : cover(gmask, kGravitySets)) {
                        rc = run_set(mod, set, field_mask, false, split, sx, sx + k_pad, sx + 2 * k_pad, k,
                                     dens_override, fvec, scale, nullptr, sout, k_pad);
                        if (rc != FB200_OK) return rc;
                    }
                }
                if (mmask && merged_m) {
                    redo |= mmask;
                    for (uint32_t set : cover(mmask, kMagneticSets)) {
                        rc = run_set(mod, set, field_mask, false, split, sx, sx + k_pad, sx + 2 * k_pad, k,
                                     pmag_override, fvec, scale, nullptr, sout, k_pad);
                        if (rc != FB200_OK) return rc;
                    }
                }
                for (int f = 0; f < F_COUNT; ++f) {
                    if (!(redo & bit(f))) continue;
                    const int row = popc(field_mask & (bit(f) - 1));
                    scatter_fields_kernel<<<dim3((unsigned)((k + 255) / 256), 1), 256, 0, mod->stream>>>(
                        sout + (int64_t)row * k_pad, k_pad, d_idx, k, d_out + (int64_t)row * ld, ld);
                    CK(cudaGetLastError());
                }
                mod->last_launches += 2;
            }
        }
        CK(cudaEventRecord(mod->ev1, mod->stream));
    }
    if (!on_device) {
        CK(cudaMemcpy2DAsync(out, (size_t)ld_out * 8, d_out, (size_t)ld * 8, (size_t)n * 8,
                             (size_t)nreq, cudaMemcpyDeviceToHost, mod->stream));
    }
    CK(cudaStreamSynchronize(mod->stream));
    if (mod->m > 0) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, mod->ev0, mod->ev1));
        mod->last_ms = ms;
    }
    return FB200_OK;
}

extern "C" int fb200_forward(fb200_model *model, const double *xp, const double *yp,
                             const double *zp, int64_t n, uint32_t field_mask,
                             const double *dens_override, const double *pmag_override,
                             const double *fvec, const double *scale, double *out,
                             int64_t ld_out, uint32_t flags)
{
    return forward_impl(model, xp, yp, zp, n, field_mask, dens_override, pmag_override, fvec,
                        scale, nullptr, out, ld_out, flags);
}

// ---------------------------------------------------------------------------
// Jacobian
// ---------------------------------------------------------------------------
// second stream + events of the host-destination pipeline, created on first use
static int ensure_copy_stream(fb200_model *mod)
{
    if (mod->copy_stream) return FB200_OK;
    CK(cudaStreamCreateWithFlags(&mod->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&mod->ev_built[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&mod->ev_copy[i], cudaEventDisableTiming));
    }
    return FB200_OK;
}

extern "C" int fb200_jacobian(fb200_model *mod, const double *xp, const double *yp,
                              const double *zp, int64_t n, int field, const double *unit_prop,
                              const double *fvec, double scale, double *out, int64_t ld,
                              uint32_t flags)
{
    if (!mod) return fail(FB200_EINVAL, "model is NULL");
    if (n < 0) return fail(FB200_EINVAL, "negative point count");
    if (field < 0 || field >= F_COUNT) return fail(FB200_EINVAL, "invalid field id");
    if (!unit_prop) return fail(FB200_EINVAL, "unit_prop is NULL");
    if (field == F_TF && !fvec) return fail(FB200_EINVAL, "tf requested without the inducing-field direction");
    if (flags & ~(FB200_DEVICE_POINTERS | FB200_REFERENCE_ORDER | FB200_TRANSPOSED | FB200_PER_CELL))
        return fail(FB200_EUNSUP, "unsupported flag for fb200_jacobian");
    if (mod->spheres && !((SPHERE_FIELDS >> field) & 1u))
        return fail(FB200_EUNSUP, "spheres offer gz, gxx..gzz, tf, bx, by, bz only (sphere.py)");
    if (mod->poly_edges)
        return fail(FB200_EUNSUP, "fb200_jacobian is not offered for polygonal prisms (rows are edges)");
    const bool on_device = (flags & FB200_DEVICE_POINTERS) != 0;
    const bool exact = (flags & FB200_REFERENCE_ORDER) != 0;
    const bool transposed = (flags & FB200_TRANSPOSED) != 0;
    const int64_t m = mod->m;
    if (n == 0 || m == 0) return FB200_OK;
    if (!xp || !yp || !zp || !out) return fail(FB200_EINVAL, "NULL point or output array");
    if (ld < (transposed ? n : m)) return fail(FB200_EINVAL, "ld smaller than the row length");
    const bool magnetic = field >= F_TF;

    std::lock_guard<std::mutex> hold(mod->lock);
    DeviceGuard g(mod->device);
    if (!g.ok) return fail(FB200_ECUDA, "cudaSetDevice failed");
    mod->last_launches = 0;
    mod->last_ms = 0.0;

    JacobianArgs a;
    std::memset(&a, 0, sizeof(a));
    view_of(mod, magnetic, &a.model);
    for (int r = 0; r < 3; ++r) a.unit_p[r] = magnetic ? unit_prop[r] : (r == 0 ? unit_prop[0] : 0.0);
    for (int r = 0; r < 3; ++r) a.fvec[r] = fvec ? fvec[r] : 0.0;
    a.scale = scale;
    // A regular mesh whose prisms are ALL its cells, in mesh order: node-sharing kernel
    // (prism_mesh_jacobian.cuh), unless the reference order or the per-cell path is asked
    // for, the node slabs do not fit in shared memory, or a point lies in a hazard wall.
    bool by_nodes = mod->has_mesh && !exact && !(flags & FB200_PER_CELL) &&
                    m == (int64_t)(mod->nxp - 1) * (mod->nyp - 1) * (mod->nzp - 1);
    // points per block: 4 (two warps per point) measured best at C4 -- 55 KB of slabs per block,
    // three resident blocks per SM overlap their node and cell phases; fewer when the slabs are large
    int mj_points = 4;
    if (by_nodes) {
        while (mj_points > 1 && mesh_jacobian_smem(mod->nxp, mod->nzp, mj_points) > (size_t)200 * 1024)
            mj_points /= 2;
        if (mesh_jacobian_smem(mod->nxp, mod->nzp, mj_points) > (size_t)200 * 1024) by_nodes = false;
    }
    auto launch = [&](const JacobianArgs &args) -> cudaError_t {
        if (by_nodes) {
            MeshJacArgs ma;
            std::memset(&ma, 0, sizeof(ma));
            ma.xn = mod->d_nodes; ma.yn = mod->d_nodes + mod->nxp; ma.zn = mod->d_nodes + mod->nxp + mod->nyp;
            ma.nxp = mod->nxp; ma.nyp = mod->nyp; ma.nzp = mod->nzp;
            for (int r = 0; r < 3; ++r) {
                ma.sh[r] = mod->shift[r]; ma.unit_p[r] = args.unit_p[r]; ma.fvec[r] = args.fvec[r];
            }
            ma.xp = args.xp; ma.yp = args.yp; ma.zp = args.zp; ma.n = args.n;
            ma.scale = args.scale; ma.out = args.out; ma.ld = args.ld;
            ma.transposed = transposed ? 1 : 0;
            ma.points = mj_points;
            return launch_mesh_jacobian(field, ma, mod->stream);
        }
        if (mod->spheres) return launch_jacobian_sphere(field, transposed, args, mod->stream);
        return exact ? launch_jacobian_exact(field, transposed, args, mod->stream)
                     : launch_jacobian_fast(field, transposed, args, mod->stream);
    };
    // Points lying in a hazard wall (see forward_impl) get their rows from the per-cell
    // kernels: flagged one by one, the whole call runs by node sharing, then those rows are
    // recomputed cell by cell and written over.
    int *d_idx = nullptr;
    int n_flagged = 0;
    auto flag_hazards = [&](const double *dx, const double *dy, const double *dz) -> int {
        if (!(mod->n_hazard[0] || mod->n_hazard[1] || mod->n_hazard[2])) return FB200_OK;
        const int64_t n_pad = (n + 63) / 64 * 64;
        int rc = grow(&mod->d_iwork, &mod->iwork_cap, (size_t)n_pad + 8, mod->stream);
        if (rc != FB200_OK) return rc;
        int *d_pflag = reinterpret_cast<int *>(mod->d_iwork);
        d_idx = d_pflag + n_pad;
        CK(cudaMemsetAsync(d_pflag, 0, (size_t)n * sizeof(int), mod->stream));
        CK(cudaMemsetAsync(mod->d_flag, 0, sizeof(int), mod->stream));
        const double *coord[3] = {dx, dy, dz};
        for (int ax = 0; ax < 3; ++ax) {
            if (!mod->n_hazard[ax]) continue;
            near_plane_kernel<<<(unsigned)((n + 255) / 256), 256, 0, mod->stream>>>(
                coord[ax], n, mod->d_hazard[ax], mod->n_hazard[ax], mod->d_flag, d_pflag);
            CK(cudaGetLastError());
        }
        CK(cudaMemsetAsync(mod->d_flag, 0, sizeof(int), mod->stream));
        compact_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, mod->stream>>>(d_pflag, n, d_idx, mod->d_flag);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&n_flagged, mod->d_flag, sizeof(int), cudaMemcpyDeviceToHost, mod->stream));
        CK(cudaStreamSynchronize(mod->stream));
        mod->last_fallback_points = n_flagged;
        return FB200_OK;
    };
    // per-cell rows of the flagged points, in blocks of <= 64 MiB, written over rows
    // [row0, row0 + nrows) of a device matrix `dst` (leading dimension ldd)
    auto redo_flagged = [&](const double *dx, const double *dy, const double *dz, double *dst, int64_t ldd,
                            int64_t row0, int64_t nrows) -> int {
        if (n_flagged == 0) return FB200_OK;
        const int64_t per = std::max<int64_t>(1, std::min<int64_t>(n_flagged, ((int64_t)(64ll << 20) / 8) / std::max<int64_t>(m, 1)));
        const int64_t per_pad = (per + 31) / 32 * 32;
        int rc = grow(&mod->d_sub, &mod->sub_cap, (size_t)per_pad * 3 + (size_t)per_pad * m, mod->stream);
        if (rc != FB200_OK) return rc;
        double *sx = mod->d_sub, *sj = mod->d_sub + 3 * per_pad;
        const bool keep = by_nodes;
        by_nodes = false;
        for (int64_t i0 = 0; i0 < n_flagged; i0 += per) {
            const int64_t k = std::min<int64_t>(per, n_flagged - i0);
            gather_points_kernel<<<(unsigned)((k + 255) / 256), 256, 0, mod->stream>>>(
                dx, dy, dz, d_idx + i0, k, per_pad, sx);
            CK(cudaGetLastError());
            JacobianArgs s2 = a;
            s2.xp = sx; s2.yp = sx + per_pad; s2.zp = sx + 2 * per_pad; s2.n = k;
            s2.out = sj; s2.ld = transposed ? k : m;
            CK(launch(s2));
            scatter_rows_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)k), 256, 0, mod->stream>>>(
                sj, s2.ld, transposed ? 1 : 0, d_idx + i0, k, m, row0, nrows, dst, ldd, transposed ? 1 : 0);
            CK(cudaGetLastError());
            mod->last_launches += 2;
        }
        by_nodes = keep;
        return FB200_OK;
    };
    mod->last_fallback_points = 0;
    // points per launch: grid.y limit, and (host output) a bounded staging block
    const int64_t max_rows_grid = (int64_t)65535 * JAC_TILE;
    if (on_device) {
        if (by_nodes) {
            int rc = flag_hazards(xp, yp, zp);
            if (rc != FB200_OK) return rc;
        }
        CK(cudaEventRecord(mod->ev0, mod->stream));
        for (int64_t r0 = 0; r0 < n; r0 += max_rows_grid) {
            const int64_t rb = std::min(max_rows_grid, n - r0);
            a.xp = xp + r0; a.yp = yp + r0; a.zp = zp + r0; a.n = rb;
            a.out = transposed ? out + r0 : out + r0 * ld;
            a.ld = ld;
            CK(launch(a));
            mod->last_launches += 1;
        }
        if (by_nodes) {
            int rc = redo_flagged(xp, yp, zp, out, ld, 0, n);
            if (rc != FB200_OK) return rc;
        }
        CK(cudaEventRecord(mod->ev1, mod->stream));
        CK(cudaStreamSynchronize(mod->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, mod->ev0, mod->ev1));
        mod->last_ms = ms;
        return FB200_OK;
    }
    // host output: the matrix is built in row blocks on the device and streamed back through
    // two staging blocks -- the kernel of block k+1 runs while block k crosses PCIe
    const int64_t np = (n + 31) / 32 * 32;
    int rc = grow(&mod->d_pts, &mod->pts_cap, (size_t)np * 3, mod->stream);
    if (rc != FB200_OK) return rc;
    CK(cudaMemcpyAsync(mod->d_pts, xp, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
    CK(cudaMemcpyAsync(mod->d_pts + np, yp, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
    CK(cudaMemcpyAsync(mod->d_pts + 2 * np, zp, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
    if (by_nodes) {
        rc = flag_hazards(mod->d_pts, mod->d_pts + np, mod->d_pts + 2 * np);
        if (rc != FB200_OK) return rc;
    }
    // row blocks of <= 128 MiB, two of them in flight
    int64_t rows = std::max<int64_t>(JAC_TILE, ((int64_t)(128ll << 20) / 8 / m) / JAC_TILE * JAC_TILE);
    rows = std::min(rows, std::min(max_rows_grid, (n + JAC_TILE - 1) / JAC_TILE * JAC_TILE));
    rc = grow(&mod->d_out, &mod->out_cap, (size_t)2 * rows * m, mod->stream);
    if (rc != FB200_OK) return rc;
    rc = ensure_copy_stream(mod);
    if (rc != FB200_OK) return rc;
    // pin the destination for the duration of the call when it is large: pageable memory
    // crosses PCIe through the driver's bounce buffers at a few GB/s
    bool pinned_here = false;
    const size_t out_bytes = (size_t)(transposed ? m : n) * (size_t)ld * 8;
    cudaPointerAttributes attr;
    const bool already = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    static const bool no_register = std::getenv("FB200_NO_HOST_REGISTER") != nullptr;
    if (!already && !no_register && out_bytes >= ((size_t)32 << 20))
        pinned_here = cudaHostRegister(out, out_bytes, cudaHostRegisterDefault) == cudaSuccess;
    if (!pinned_here) (void)cudaGetLastError();
    CK(cudaEventRecord(mod->ev0, mod->stream));
    int64_t blk = 0;
    int status = FB200_OK;
    for (int64_t r0 = 0; r0 < n && status == FB200_OK; r0 += rows, ++blk) {
        const int64_t rb = std::min(rows, n - r0);
        double *stage = mod->d_out + (size_t)(blk & 1) * rows * m;
        // the staging block is free once its previous copy (two blocks ago) has finished
        if (blk >= 2 && cudaStreamWaitEvent(mod->stream, mod->ev_copy[blk & 1], 0) != cudaSuccess) { status = FB200_ECUDA; break; }
        a.xp = mod->d_pts + r0; a.yp = mod->d_pts + np + r0; a.zp = mod->d_pts + 2 * np + r0;
        a.n = rb;
        a.out = stage;
        a.ld = transposed ? rb : m;
        if (launch(a) != cudaSuccess) { status = FB200_ECUDA; break; }
        mod->last_launches += 1;
        if (by_nodes && n_flagged) {
            status = redo_flagged(mod->d_pts, mod->d_pts + np, mod->d_pts + 2 * np, stage, a.ld, r0, rb);
            if (status != FB200_OK) break;
        }
        if (cudaEventRecord(mod->ev_built[blk & 1], mod->stream) != cudaSuccess) { status = FB200_ECUDA; break; }
        if (cudaStreamWaitEvent(mod->copy_stream, mod->ev_built[blk & 1], 0) != cudaSuccess) { status = FB200_ECUDA; break; }
        cudaError_t ce;
        if (!transposed)
            ce = cudaMemcpy2DAsync(out + r0 * ld, (size_t)ld * 8, stage, (size_t)m * 8,
                                   (size_t)m * 8, (size_t)rb, cudaMemcpyDeviceToHost, mod->copy_stream);
        else
            ce = cudaMemcpy2DAsync(out + r0, (size_t)ld * 8, stage, (size_t)rb * 8,
                                   (size_t)rb * 8, (size_t)m, cudaMemcpyDeviceToHost, mod->copy_stream);
        if (ce != cudaSuccess || cudaEventRecord(mod->ev_copy[blk & 1], mod->copy_stream) != cudaSuccess) { status = FB200_ECUDA; break; }
    }
    cudaError_t e1 = cudaEventRecord(mod->ev1, mod->stream);
    cudaError_t e2 = cudaStreamSynchronize(mod->copy_stream);
    cudaError_t e3 = cudaStreamSynchronize(mod->stream);
    if (pinned_here) cudaHostUnregister(out);
    if (status != FB200_OK) return fail(status, "Jacobian row-block pipeline failed");
    CK(e1); CK(e2); CK(e3);
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, mod->ev0, mod->ev1));
    mod->last_ms = ms;       // device time of the kernels (the copies overlap them on the copy stream)
    return FB200_OK;
}

// ---------------------------------------------------------------------------
// Jacobian of a regular mesh on a commensurate level grid (prism_grid_reuse.cuh)
// ---------------------------------------------------------------------------
extern "C" int fb200_jacobian_grid(fb200_model *mod, int64_t ni, int64_t nj, int kx, int ky, const double *xu,
                                   const double *yv, const double *zc, int field, const double *unit_prop,
                                   const double *fvec, double scale, double *out, int64_t ld, uint32_t flags)
{
    if (!mod) return fail(FB200_EINVAL, "model is NULL");
    if (field < 0 || field >= F_COUNT) return fail(FB200_EINVAL, "invalid field id");
    if (ni <= 0 || nj <= 0 || kx <= 0 || ky <= 0 || !xu || !yv || !zc || !out || !unit_prop)
        return fail(FB200_EINVAL, "bad argument to fb200_jacobian_grid");
    if (flags & ~FB200_DEVICE_POINTERS) return fail(FB200_EUNSUP, "unsupported flag for fb200_jacobian_grid");
    if (!mod->has_mesh) return fail(FB200_EUNSUP, "the model has no regular-mesh form attached");
    const bool magnetic = field >= F_TF;
    if (field == F_TF && !fvec) return fail(FB200_EINVAL, "tf needs the inducing-field direction");
    const int nx = mod->nxp - 1, ny = mod->nyp - 1, nz = mod->nzp - 1;
    const int64_t m = (int64_t)nx * ny * nz, n = ni * nj;
    if (mod->m != m) return fail(FB200_EUNSUP, "the model's prisms are not all the cells of its mesh");
    if (ld < m) return fail(FB200_EINVAL, "ld smaller than the number of cells");
    // the TMA path moves whole cell rows of nx doubles: 16-byte granularity
    if (nx % 2 || ld % 2 || (reinterpret_cast<uintptr_t>(out) % 16) || (int64_t)nx * 8 > GR_CHUNK_BYTES)
        return fail(FB200_EUNSUP, "commensurate-grid Jacobian needs an even number of cells along x, an even "
                                  "leading dimension and a 16-byte aligned destination");
    const bool on_device = (flags & FB200_DEVICE_POINTERS) != 0;

    std::lock_guard<std::mutex> hold(mod->lock);
    DeviceGuard g(mod->device);
    if (!g.ok) return fail(FB200_ECUDA, "cudaSetDevice failed");
    mod->last_launches = 0;
    mod->last_ms = 0.0;
    mod->last_fallback_points = 0;

    GridPlan gp;
    gp.nx = nx; gp.ny = ny; gp.nz = nz; gp.kx = kx; gp.ky = ky; gp.ni = ni; gp.nj = nj;
    gp.nu = (int64_t)nx * kx + ni; gp.nv = (int64_t)ny * ky + nj;
    gp.ndu = (int64_t)(nx - 1) * kx + ni; gp.ndv = (int64_t)(ny - 1) * ky + nj;
    gp.ne = ((gp.ndu + kx - 1) / kx + 2 + 1) / 2 * 2;
    const size_t nT = (size_t)(nz + 1) * gp.nv * gp.nu;
    const size_t nD = (size_t)nz * gp.ndv * kx * gp.ne;
    const size_t ncoord = (size_t)gp.nu + gp.nv + (nz + 1) + 8;
    int rc = grow(&mod->d_grid, &mod->grid_cap, nT + 2 * nD + ncoord + 16, mod->stream);
    if (rc != FB200_OK) return rc;
    double *T = mod->d_grid, *D0 = T + (nT + 1) / 2 * 2, *D1 = D0 + (nD + 1) / 2 * 2;
    double *dxu = D1 + (nD + 1) / 2 * 2, *dyv = dxu + gp.nu, *dzc = dyv + gp.nv;
    CK(cudaMemcpyAsync(dxu, xu, (size_t)gp.nu * 8, cudaMemcpyHostToDevice, mod->stream));
    CK(cudaMemcpyAsync(dyv, yv, (size_t)gp.nv * 8, cudaMemcpyHostToDevice, mod->stream));
    CK(cudaMemcpyAsync(dzc, zc, (size_t)(nz + 1) * 8, cudaMemcpyHostToDevice, mod->stream));
    double up[3], fv[3];
    for (int r = 0; r < 3; ++r) {
        up[r] = (magnetic ? unit_prop[r] : (r == 0 ? unit_prop[0] : 0.0)) * scale;   // like mesh_jacobian_kernel
        fv[r] = fvec ? fvec[r] : 0.0;
    }
    CK(cudaEventRecord(mod->ev0, mod->stream));
    CK(launch_grid_tables(field, gp, dxu, dyv, dzc, up, fv, mod->shift, T, D0, D1, mod->stream));
    mod->last_launches += 2;
    if (on_device) {
        const int64_t per = 1 << 20;
        for (int64_t r0 = 0; r0 < n; r0 += per) {
            const int64_t rb = std::min(per, n - r0);
            CK(launch_grid_expand(gp, D0, D1, r0, rb, out + r0 * ld, ld, mod->stream));
            mod->last_launches += 1;
        }
        CK(cudaEventRecord(mod->ev1, mod->stream));
        CK(cudaStreamSynchronize(mod->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, mod->ev0, mod->ev1));
        mod->last_ms = ms;
        return FB200_OK;
    }
    // host destination: row blocks through two staging blocks, copies overlap the expansion
    int64_t rows = std::max<int64_t>(1, ((int64_t)(128ll << 20) / 8) / m);
    rows = std::min(rows, n);
    rc = grow(&mod->d_out, &mod->out_cap, (size_t)2 * rows * m, mod->stream);
    if (rc != FB200_OK) return rc;
    rc = ensure_copy_stream(mod);
    if (rc != FB200_OK) return rc;
    int64_t blk = 0;
    for (int64_t r0 = 0; r0 < n; r0 += rows, ++blk) {
        const int64_t rb = std::min(rows, n - r0);
        double *stage = mod->d_out + (size_t)(blk & 1) * rows * m;
        if (blk >= 2) CK(cudaStreamWaitEvent(mod->stream, mod->ev_copy[blk & 1], 0));
        CK(launch_grid_expand(gp, D0, D1, r0, rb, stage, m, mod->stream));
        mod->last_launches += 1;
        CK(cudaEventRecord(mod->ev_built[blk & 1], mod->stream));
        CK(cudaStreamWaitEvent(mod->copy_stream, mod->ev_built[blk & 1], 0));
        CK(cudaMemcpy2DAsync(out + r0 * ld, (size_t)ld * 8, stage, (size_t)m * 8, (size_t)m * 8, (size_t)rb,
                             cudaMemcpyDeviceToHost, mod->copy_stream));
        CK(cudaEventRecord(mod->ev_copy[blk & 1], mod->copy_stream));
    }
    CK(cudaEventRecord(mod->ev1, mod->stream));
    CK(cudaStreamSynchronize(mod->copy_stream));
    CK(cudaStreamSynchronize(mod->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, mod->ev0, mod->ev1));
    mod->last_ms = ms;
    return FB200_OK;
}

// ---------------------------------------------------------------------------
// adjoint J^T d (imaging.py:116-118), never storing J
// ---------------------------------------------------------------------------
extern "C" int fb200_adjoint(fb200_model *mod, const double *xp, const double *yp,
                             const double *zp, int64_t n, int field, const double *unit_prop,
                             const double *fvec, double scale, const double *data, double *out,
                             uint32_t flags)
{
    if (!mod) return fail(FB200_EINVAL, "model is NULL");
    if (n < 0) return fail(FB200_EINVAL, "negative point count");
    if (field < 0 || field >= F_COUNT) return fail(FB200_EINVAL, "invalid field id");
    if (!unit_prop) return fail(FB200_EINVAL, "unit_prop is NULL");
    if (field == F_TF && !fvec) return fail(FB200_EINVAL, "tf requested without the inducing-field direction");
    if (flags & ~(FB200_DEVICE_POINTERS | FB200_REFERENCE_ORDER | FB200_PER_CELL))
        return fail(FB200_EUNSUP, "unsupported flag for fb200_adjoint");
    if (mod->spheres || mod->poly_edges)
        return fail(FB200_EUNSUP, "fb200_adjoint is not offered for spheres / polygonal prisms");
    const bool on_device = (flags & FB200_DEVICE_POINTERS) != 0;
    const bool exact = (flags & FB200_REFERENCE_ORDER) != 0;
    const int64_t m = mod->m;
    if (m == 0) return FB200_OK;
    if (!out || (n > 0 && (!xp || !yp || !zp || !data))) return fail(FB200_EINVAL, "NULL array");
    const bool magnetic = field >= F_TF;

    std::lock_guard<std::mutex> hold(mod->lock);
    DeviceGuard g(mod->device);
    if (!g.ok) return fail(FB200_ECUDA, "cudaSetDevice failed");
    mod->last_launches = 0;
    mod->last_ms = 0.0;

    const double *d_xp = xp, *d_yp = yp, *d_zp = zp, *d_w = data;
    double *d_out = out;
    const int64_t mp = (m + 31) / 32 * 32;
    if (!on_device) {
        const int64_t np = (n + 31) / 32 * 32;
        int rc = grow(&mod->d_pts, &mod->pts_cap, (size_t)np * 4 + 32, mod->stream);
        if (rc != FB200_OK) return rc;
        rc = grow(&mod->d_out, &mod->out_cap, (size_t)mp, mod->stream);
        if (rc != FB200_OK) return rc;
        if (n > 0) {
            CK(cudaMemcpyAsync(mod->d_pts, xp, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
            CK(cudaMemcpyAsync(mod->d_pts + np, yp, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
            CK(cudaMemcpyAsync(mod->d_pts + 2 * np, zp, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
            CK(cudaMemcpyAsync(mod->d_pts + 3 * np, data, (size_t)n * 8, cudaMemcpyHostToDevice, mod->stream));
        }
        d_xp = mod->d_pts; d_yp = mod->d_pts + np; d_zp = mod->d_pts + 2 * np; d_w = mod->d_pts + 3 * np;
        d_out = mod->d_out;
    }
    // whole regular mesh in mesh order: node sharing (EvalNodeRev); points lying in a hazard wall
    // (fb200_model_set_hazard_planes) are taken out of that sum and added by the per-cell kernels
    bool by_nodes = n > 0 && mod->has_mesh && !exact && !(flags & FB200_PER_CELL) &&
                    m == (int64_t)(mod->nxp - 1) * (mod->nyp - 1) * (mod->nzp - 1);
    int *d_pflag = nullptr, *d_idx = nullptr;
    int n_flagged = 0;
    mod->last_fallback_points = 0;
    if (by_nodes && (mod->n_hazard[0] || mod->n_hazard[1] || mod->n_hazard[2])) {
        const int64_t n_pad = (n + 63) / 64 * 64;
        int rc = grow(&mod->d_iwork, &mod->iwork_cap, (size_t)n_pad + 8, mod->stream);
        if (rc != FB200_OK) return rc;
        d_pflag = reinterpret_cast<int *>(mod->d_iwork);
        d_idx = d_pflag + n_pad;
        CK(cudaMemsetAsync(d_pflag, 0, (size_t)n * sizeof(int), mod->stream));
        CK(cudaMemsetAsync(mod->d_flag, 0, sizeof(int), mod->stream));
        const double *coord[3] = {d_xp, d_yp, d_zp};
        for (int ax = 0; ax < 3; ++ax) {
            if (!mod->n_hazard[ax]) continue;
            near_plane_kernel<<<(unsigned)((n + 255) / 256), 256, 0, mod->stream>>>(
                coord[ax], n, mod->d_hazard[ax], mod->n_hazard[ax], mod->d_flag, d_pflag);
            CK(cudaGetLastError());
        }
        CK(cudaMemsetAsync(mod->d_flag, 0, sizeof(int), mod->stream));
        compact_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, mod->stream>>>(d_pflag, n, d_idx, mod->d_flag);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&n_flagged, mod->d_flag, sizeof(int), cudaMemcpyDeviceToHost, mod->stream));
        CK(cudaStreamSynchronize(mod->stream));
        mod->last_fallback_points = n_flagged;
        // the adjoint SUMS over the flagged points: put the (atomically appended) list in point
        // order, so that the result does not change from run to run
        if (n_flagged > 1) thrust::sort(thrust::cuda::par.on(mod->stream), d_idx, d_idx + n_flagged);
    }
    // J^T d of the points (dx, dy, dz, dw)[0 .. cnt) by the per-cell kernels into dst[0 .. m)
    auto per_cell = [&](const double *dx, const double *dy, const double *dz, const double *dw, int64_t cnt,
                        double *dst) -> int {
        // split the points so that every SM has work even for small models
        const int64_t blocks_x = (m + JAC_THREADS - 1) / JAC_THREADS;
        const int64_t tiles = (cnt + JAC_TILE - 1) / JAC_TILE;
        int64_t nsplit = std::min<int64_t>(
            std::max<int64_t>(1, ((int64_t)mod->sm_count * split_blocks_per_sm() + blocks_x - 1) / blocks_x), tiles);
        nsplit = std::min<int64_t>(nsplit, 65535);
        const int64_t rows = ((tiles + nsplit - 1) / nsplit) * JAC_TILE;
        nsplit = (cnt + rows - 1) / rows;
        int rc = grow(&mod->d_part, &mod->part_cap, (size_t)nsplit * mp, mod->stream);
        if (rc != FB200_OK) return rc;
        AdjointArgs a;
        std::memset(&a, 0, sizeof(a));
        view_of(mod, magnetic, &a.j.model);
        for (int r = 0; r < 3; ++r) a.j.unit_p[r] = magnetic ? unit_prop[r] : (r == 0 ? unit_prop[0] : 0.0);
        for (int r = 0; r < 3; ++r) a.j.fvec[r] = fvec ? fvec[r] : 0.0;
        a.j.xp = dx; a.j.yp = dy; a.j.zp = dz; a.j.n = cnt;
        a.weights = dw;
        a.part = mod->d_part;
        a.ld_p = mp;
        a.rows_per_split = rows;
        CK(exact ? launch_adjoint_exact(field, a, (int)nsplit, mod->stream)
                 : launch_adjoint_fast(field, a, (int)nsplit, mod->stream));
        ReduceArgs r;
        std::memset(&r, 0, sizeof(r));
        r.part = mod->d_part; r.out = dst; r.n = m; r.ld = mp; r.ld_out = mp;
        r.nc = 1; r.nsplit = (int)nsplit; r.scale[0] = scale; r.out_row[0] = 0;
        reduce_partials_kernel<<<dim3((unsigned)((m + 255) / 256), 1), 256, 0, mod->stream>>>(r);
        CK(cudaGetLastError());
        mod->last_launches += 2;
        return FB200_OK;
    };
    if (n == 0) {
        fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, mod->stream>>>(d_out, m, 0.0);
        CK(cudaGetLastError());
    } else if (by_nodes) {
        const int nw = magnetic ? 3 : 1;
        const int64_t n_pad = (n + TILE - 1) / TILE * TILE;
        const int64_t nn = mod->nnodes, nn_pad = (nn + 31) / 32 * 32;
        // scratch: point rows | node coordinates | V
        const size_t need = (size_t)(6 + nw) * n_pad + (size_t)4 * nn_pad;
        int rc = grow(&mod->d_adj, &mod->adj_cap, need, mod->stream);
        if (rc != FB200_OK) return rc;
        double *rows = mod->d_adj, *nodes = rows + (size_t)(6 + nw) * n_pad, *V = nodes + (size_t)3 * nn_pad;
        CK(cudaEventRecord(mod->ev0, mod->stream));
        adjoint_rows_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, mod->stream>>>(
            d_xp, d_yp, d_zp, d_w, n, n_pad, nw, unit_prop[0], magnetic ? unit_prop[1] : 0.0,
            magnetic ? unit_prop[2] : 0.0, mod->shift[0], mod->shift[1], mod->shift[2], rows,
            n_flagged ? d_pflag : nullptr);
        CK(cudaGetLastError());
        expand_nodes_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, mod->stream>>>(
            mod->d_nodes, mod->d_nodes + mod->nxp, mod->d_nodes + mod->nxp + mod->nyp, mod->nxp,
            mod->nyp, mod->nzp, nn_pad, nodes);
        CK(cudaGetLastError());
        ForwardArgs a;
        std::memset(&a, 0, sizeof(a));
        for (int r = 0; r < 6; ++r) a.model.b[r] = rows + (size_t)r * n_pad;
        for (int r = 0; r < nw; ++r) a.model.p[r] = rows + (size_t)(6 + r) * n_pad;
        a.model.m = n;
        a.xp = nodes; a.yp = nodes + nn_pad; a.zp = nodes + 2 * nn_pad; a.n = nn;
        for (int r = 0; r < 3; ++r) a.fvec[r] = fvec ? fvec[r] : 0.0;
        const int64_t blocks_x = (nn + FWD_THREADS - 1) / FWD_THREADS;
        const int64_t tiles = n_pad / TILE;
        int64_t nsplit = std::min<int64_t>(
            std::max<int64_t>(1, ((int64_t)mod->sm_count * split_blocks_per_sm() + blocks_x - 1) / blocks_x), tiles);
        nsplit = std::min<int64_t>(nsplit, 32000);
        const int64_t chunk = ((tiles + nsplit - 1) / nsplit) * TILE;
        nsplit = (n + chunk - 1) / chunk;
        rc = grow(&mod->d_part, &mod->part_cap, (size_t)nsplit * nn_pad, mod->stream);
        if (rc != FB200_OK) return rc;
        a.chunk = chunk;
        a.nsplit = (int)std::max<int64_t>(nsplit, 2);     // partial-sum output mode
        a.out = mod->d_part;
        a.ld = nn_pad;
        CK(launch_forward_node_rev(field, a, (int)nsplit, mod->stream));
        const double one[1] = {1.0};
        const int row0[1] = {0};
        rc = reduce_split(mod, 1, nsplit, nn, nn_pad, one, row0, V, nn_pad);
        if (rc != FB200_OK) return rc;
        adjoint_cells_kernel<<<(unsigned)((m + 255) / 256), 256, 0, mod->stream>>>(
            V, mod->nxp - 1, mod->nyp - 1, mod->nzp - 1, scale, d_out);
        CK(cudaGetLastError());
        mod->last_launches = 5;
        if (n_flagged > 0) {
            // the flagged points, cell by cell, added to the node-sharing sum of the others
            const int64_t k = n_flagged, k_pad = (k + 31) / 32 * 32;
            rc = grow(&mod->d_sub, &mod->sub_cap, (size_t)4 * k_pad + (size_t)mp, mod->stream);
            if (rc != FB200_OK) return rc;
            double *sx = mod->d_sub, *sw = sx + 3 * k_pad, *tmp = sw + k_pad;
            gather_points_kernel<<<(unsigned)((k + 255) / 256), 256, 0, mod->stream>>>(d_xp, d_yp, d_zp, d_idx, k,
                                                                                      k_pad, sx);
            CK(cudaGetLastError());
            gather_values_kernel<<<(unsigned)((k + 255) / 256), 256, 0, mod->stream>>>(d_w, d_idx, k, sw);
            CK(cudaGetLastError());
            rc = per_cell(sx, sx + k_pad, sx + 2 * k_pad, sw, k, tmp);
            if (rc != FB200_OK) return rc;
            add_kernel<<<(unsigned)((m + 255) / 256), 256, 0, mod->stream>>>(d_out, tmp, m);
            CK(cudaGetLastError());
            mod->last_launches += 3;
        }
        CK(cudaEventRecord(mod->ev1, mod->stream));
    } else {
        CK(cudaEventRecord(mod->ev0, mod->stream));
        int rc = per_cell(d_xp, d_yp, d_zp, d_w, n, d_out);
        if (rc != FB200_OK) return rc;
        CK(cudaEventRecord(mod->ev1, mod->stream));
    }
    if (!on_device)
        CK(cudaMemcpyAsync(out, d_out, (size_t)m * 8, cudaMemcpyDeviceToHost, mod->stream));
    CK(cudaStreamSynchronize(mod->stream));
    if (n > 0) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, mod->ev0, mod->ev1));
        mod->last_ms = ms;
    }
    return FB200_OK;
}

// ---------------------------------------------------------------------------
// the reference's own FFI shape: one prism accumulated into res
// ---------------------------------------------------------------------------
static int g_default_device = 0;

extern "C" int fb200_set_default_device(int device)
{
    int count = 0;
    int rc = fb200_device_count(&count);
    if (rc != FB200_OK) return rc;
    if (device < 0 || device >= count) return fail(FB200_EINVAL, "device index out of range");
    g_default_device = device;
    return FB200_OK;
}

static int prism_accumulate(int field, const double *xp, const double *yp, const double *zp,
                            int64_t n, double x1, double x2, double y1, double y2, double z1,
                            double z2, const double *props, const double *fvec, double *res)
{
    if (n < 0) return fail(FB200_EINVAL, "negative point count");
    if (n == 0) return FB200_OK;
    if (!xp || !yp || !zp || !res) return fail(FB200_EINVAL, "NULL array");
    const bool magnetic = field >= F_TF;
    fb200_model *mod = nullptr;
    int rc = fb200_model_create(g_default_device, 1, &x1, &x2, &y1, &y2, &z1, &z2,
                                magnetic ? nullptr : props, magnetic ? props : nullptr,
                                magnetic ? props + 1 : nullptr, magnetic ? props + 2 : nullptr, &mod);
    if (rc != FB200_OK) return rc;
    rc = forward_impl(mod, xp, yp, zp, n, bit(field), nullptr, nullptr, fvec, nullptr, res, res, n,
                      FB200_REFERENCE_ORDER);
    std::string keep = g_err;
    fb200_model_destroy(mod);
    g_err = keep;
    return rc;
}

#define GRAVITY_ENTRY(name, F)                                                                  \
    extern "C" int fb200_prism_##name(const double *xp, const double *yp, const double *zp,     \
                                      int64_t n, double x1, double x2, double y1, double y2,    \
                                      double z1, double z2, double density, double *res)        \
    {                                                                                           \
        return prism_accumulate(F, xp, yp, zp, n, x1, x2, y1, y2, z1, z2, &density, nullptr, res); \
    }
GRAVITY_ENTRY(potential, F_POT)
GRAVITY_ENTRY(gx, F_GX)
GRAVITY_ENTRY(gy, F_GY)
GRAVITY_ENTRY(gz, F_GZ)
GRAVITY_ENTRY(gxx, F_GXX)
GRAVITY_ENTRY(gxy, F_GXY)
GRAVITY_ENTRY(gxz, F_GXZ)
GRAVITY_ENTRY(gyy, F_GYY)
GRAVITY_ENTRY(gyz, F_GYZ)
GRAVITY_ENTRY(gzz, F_GZZ)
#undef GRAVITY_ENTRY

extern "C" int fb200_prism_tf(const double *xp, const double *yp, const double *zp, int64_t n,
                              double x1, double x2, double y1, double y2, double z1, double z2,
                              double mx, double my, double mz, double fx, double fy, double fz,
                              double *res)
{
    const double m[3] = {mx, my, mz}, f[3] = {fx, fy, fz};
    return prism_accumulate(F_TF, xp, yp, zp, n, x1, x2, y1, y2, z1, z2, m, f, res);
}

#define MAGNETIC_ENTRY(name, F)                                                                 \
    extern "C" int fb200_prism_##name(const double *xp, const double *yp, const double *zp,     \
                                      int64_t n, double x1, double x2, double y1, double y2,    \
                                      double z1, double z2, double mx, double my, double mz,    \
                                      double *res)                                              \
    {                                                                                           \
        const double m[3] = {mx, my, mz};                                                       \
        return prism_accumulate(F, xp, yp, zp, n, x1, x2, y1, y2, z1, z2, m, nullptr, res);     \
    }
MAGNETIC_ENTRY(bx, F_BX)
MAGNETIC_ENTRY(by, F_BY)
MAGNETIC_ENTRY(bz, F_BZ)
#undef MAGNETIC_ENTRY

// ---------------------------------------------------------------------------
// measurement helpers
// ---------------------------------------------------------------------------
extern "C" int fb200_fp64_peak(int device, double seconds, double *burst_tflops,
                               double *sustained_tflops)
{
    int count = 0;
    int rc = fb200_device_count(&count);
    if (rc != FB200_OK) return rc;
    if (device < 0 || device >= count) return fail(FB200_EINVAL, "device index out of range");
    DeviceGuard g(device);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 2048;
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double flop = (double)blocks * threads * (double)iters * 16.0 * 8.0 * 2.0;
    for (int w = 0; w < 3; ++w) fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    CK(cudaDeviceSynchronize());
    double best = 0.0, total_ms = 0.0, total_flop = 0.0;
    while (total_ms < seconds * 1e3) {
        CK(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, flop / (ms * 1e-3) / 1e12);
        total_ms += ms;
        total_flop += flop;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (burst_tflops) *burst_tflops = best;
    if (sustained_tflops) *sustained_tflops = total_flop / (total_ms * 1e-3) / 1e12;
    return FB200_OK;
}

extern "C" int fb200_flush_l2(int device)
{
    static double *bufs[64] = {nullptr};
    if (device < 0 || device >= 64) return fail(FB200_EINVAL, "device index out of range");
    DeviceGuard g(device);
    const size_t bytes = (size_t)256 << 20;   // 2x the 126 MB L2
    if (!bufs[device]) CK(cudaMalloc((void **)&bufs[device], bytes));
    CK(cudaMemset(bufs[device], 0, bytes));
    CK(cudaDeviceSynchronize());
    return FB200_OK;
}

// op: 0 log_ratio(a, b), 1 atan_rhp(im = a, re = b), 2 div_fast(a, b), 3 root(a), 4 sqrt(a).
// Host arrays; exists so that tests can pin the device math against mpmath/numpy.
extern "C" int fb200_debug_math(int device, int op, const double *a, const double *b, int64_t n,
                                double *out)
{
    if (n <= 0) return FB200_OK;
    if (!a || !b || !out) return fail(FB200_EINVAL, "NULL array");
    DeviceGuard g(device);
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, (size_t)n * 3 * sizeof(double)));
    CK(cudaMemcpy(d, a, (size_t)n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d + n, b, (size_t)n * 8, cudaMemcpyHostToDevice));
    debug_math_kernel<<<(unsigned)((n + 255) / 256), 256>>>(op, d, d + n, n, d + 2 * n);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, d + 2 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return FB200_OK;
}
