// normal.cu -- consumers of a DEVICE-RESIDENT sensitivity matrix: the Gauss-Newton products of
// fatiando/inversion/misfit.py on a row block of J.
//
//   fb200_gram      G = factor * J^T diag(w) J      (misfit.py:250-256: hessian = 2 J^T W J)
//   fb200_matvec    y = J p                         (misfit.py: predicted data of a linear problem)
//   fb200_matvec_t  g = factor * J^T (w .* r)       (misfit.py:280-294: gradient = -2 J^T W r)
//
// J is (n x m), C-order, row = observation (the layout fb200_jacobian writes, eqlayer.py:108-111);
// on several GPUs every rank owns a row block (partition.jacobian_row_sharded) and G, g are summed
// with ONE all-reduce (partition.gram_row_sharded) -- the one place on this path where a
// collective moves real data (12.8 GB for the 40 000^2 matrix of config C4).
//
// The Gram matrix is a rank-n update: 2*n*m^2/2 FLOP on n*m*8 bytes -- at C4 (n = 31 250 per GPU,
// m = 40 000) 1 250 FLOP per byte, FP64-pipe-bound.  gram_kernel is an fp64 SYRK on the FP64
// tensor path (mma.sync m8n8k4): a 128 x 128 tile of G per block of 256 threads, a 64 x 32 warp
// tile of 8 x 4 mma tiles (64 accumulator doubles per thread), the two 32-row
// panels of J staged through a three-stage cp.async ring (8-byte copies, zero-filled past the
// matrix edges, so any n, m and leading dimension work); only tiles on or above the diagonal are
// computed, mirror_kernel fills the rest.  Per four rows of J a warp issues 32 DMMA (8192
// multiply-adds) per 12 shared-memory loads; panel rows are padded by 8 doubles so that a
// fragment load (4 rows x 8 consecutive doubles) takes the minimum of two wavefronts.
// Round-2 history: the same tiling with DFMA register tiles (8 x 8 per thread) stopped at 72 % of
// the DFMA peak -- three-register DFMAs cap at 90 % (tools/fp64_pipe_probe.cu) and the loads take
// issue slots; tools/dmma_probe.cu shows the fp64 mma path running at the same 37 TFLOP/s with
// one eighth of the instructions.  (tcgen05 has no fp64 operand type; mma.sync is the fp64
// tensor path on sm_100a.)
#include <algorithm>
#include <cstdint>
#include <string>

#include "../../include/fatiando_b200.h"
#include "abi_internal.h"

namespace fb200 {

constexpr int GT = 128;          // tile of G (rows and columns)
#ifndef FB200_GK
#define FB200_GK 32
#endif
#ifndef FB200_GSTAGES
#define FB200_GSTAGES 3
#endif
constexpr int GK = FB200_GK;     // rows of J per stage
constexpr int GTHREADS = 256;     // 8 warps, 8 x 8 accumulators per thread
constexpr int GSTAGES = FB200_GSTAGES;

struct GramArgs {
    const double *J;
    const double *w;             // row weights or nullptr
    int64_t n, m, ld;
    double factor;
    double *G;
    int64_t ldg;
    int tiles;                   // tiles per side
};

__device__ __forceinline__ void cp_async8(unsigned dst, const void *src, bool valid)
{
    const int bytes = valid ? 8 : 0;            // src-size 0: the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src, int bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

// Tile (bi <= bj) of a linear block index.  Blocks that run at the same time should share panels
// of J in L2: the upper triangle is walked in SUPER x SUPER groups of tiles (a group's ~100
// resident blocks read SUPER + SUPER panels instead of one each).
constexpr int SUPER = 12;
__device__ __forceinline__ void tile_of(int64_t id, int tiles, int *bi, int *bj)
{
    const int ns = (tiles + SUPER - 1) / SUPER;          // super-tiles per side
    // super-tile rows: row I holds (ns - I) super-tiles; diagonal ones are triangles
    int I = 0;
    for (;; ++I) {
        const int rows = min(SUPER, tiles - I * SUPER);
        // tiles in super-row I: diagonal super-tile (triangle) + full ones to the right
        const int64_t diag = (int64_t)rows * (rows + 1) / 2;
        const int64_t right = (int64_t)rows * (tiles - (I * SUPER + rows));
        if (id < diag + right) {
            if (id < diag) {                               // inside the diagonal super-tile
                int r = 0;
                while (id >= rows - r) { id -= rows - r; ++r; }
                *bi = I * SUPER + r;
                *bj = I * SUPER + r + (int)id;
            } else {
                id -= diag;
                // full super-tiles to the right, SUPER columns each (the last one may be narrower)
                const int64_t per = (int64_t)rows * SUPER;
                const int Jx = (int)(id / per);
                const int64_t in = id - (int64_t)Jx * per;
                const int c0 = I * SUPER + rows + Jx * SUPER;
                const int cols = min(SUPER, tiles - c0);
                *bi = I * SUPER + (int)(in / cols);
                *bj = c0 + (int)(in % cols);
            }
            return;
        }
        id -= diag + right;
    }
}

// fp64 tensor-core path: mma.sync m8n8k4 (DMMA).  tools/dmma_probe.cu measures 37.0 TFLOP/s for it
// on B200 -- the same rate as the DFMA pipe, but one instruction carries 256 multiply-adds of a
// warp instead of 32, so the issue slots that cap a DFMA SYRK at ~72 % are free.
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

constexpr int GTP = GT + 8;      // padded panel row: 4 consecutive rows fall into 2 bank groups

template <bool WEIGHTED, bool VEC>
__global__ void __launch_bounds__(GTHREADS, 1) gram_kernel(const GramArgs a)
{
    extern __shared__ __align__(16) double smem[];
    // stage s: A panel [GK][GTP], B panel [GK][GTP], weights [GK]
    constexpr int STAGE = 2 * GK * GTP + GK;
    int bi, bj;
    tile_of((int64_t)blockIdx.x, a.tiles, &bi, &bj);
    const int64_t qa = (int64_t)bi * GT, qb = (int64_t)bj * GT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 8 warps = 2 (rows) x 4 (columns); a warp owns 64 x 32 of the tile = 8 x 4 mma tiles
    const int wr = (warp >> 2) * 64, wc = (warp & 3) * 32;
    const int fr = lane >> 2, fk = lane & 3;          // fragment row / column index, k index
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const int ktiles = (int)((a.n + GK - 1) / GK);
    // Loader.  VEC: 16-byte copies, a thread owns the column pair c0 of rows r0, r0 + RSTEP, ..
    // (ld even, J 16-byte aligned); otherwise 8-byte copies, any layout.  Column validity is
    // fixed per thread; row validity only matters in the last stage.
    constexpr int PER = VEC ? (GK * GT / 2) / GTHREADS : (GK * GT) / GTHREADS;     // copies per panel
    constexpr int RSTEP = VEC ? GTHREADS / (GT / 2) : GTHREADS / GT;               // rows between copies
    const int r0 = VEC ? tid / (GT / 2) : tid / GT;
    const int c0 = VEC ? (tid % (GT / 2)) * 2 : tid % GT;
    const int a_bytes = VEC ? (qa + c0 + 1 < a.m ? 16 : (qa + c0 < a.m ? 8 : 0)) : (qa + c0 < a.m ? 8 : 0);
    const int b_bytes = VEC ? (qb + c0 + 1 < a.m ? 16 : (qb + c0 < a.m ? 8 : 0)) : (qb + c0 < a.m ? 8 : 0);
    const double *pa = a.J + (int64_t)r0 * a.ld + (a_bytes ? qa + c0 : 0);
    const double *pb = a.J + (int64_t)r0 * a.ld + (b_bytes ? qb + c0 : 0);
    const int64_t row_step = (int64_t)RSTEP * a.ld, stage_step = (int64_t)GK * a.ld;
    const unsigned d0 = (unsigned)(r0 * GTP + c0) * 8u;
    auto issue = [&](int t) {
        const unsigned dst = sbase + (unsigned)((t % GSTAGES) * STAGE) * 8u + d0;
        const int64_t l0 = (int64_t)t * GK + r0;
        const bool full = (int64_t)(t + 1) * GK <= a.n;
        const double *ra = pa, *rb = pb;
        if (VEC && full) {                      // the common case: no row predicates
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const unsigned off = (unsigned)(i * RSTEP * GTP) * 8u;
                cp_async16(dst + off, ra, a_bytes);
                cp_async16(dst + off + (unsigned)(GK * GTP) * 8u, rb, b_bytes);
                ra += row_step;
                rb += row_step;
            }
        } else
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const bool ok = full || l0 + i * RSTEP < a.n;
            const unsigned off = (unsigned)(i * RSTEP * GTP) * 8u;
            if (VEC) {
                cp_async16(dst + off, ok ? ra : a.J, ok ? a_bytes : 0);
                cp_async16(dst + off + (unsigned)(GK * GTP) * 8u, ok ? rb : a.J, ok ? b_bytes : 0);
            } else {
                cp_async8(dst + off, ok ? ra : a.J, ok && a_bytes);
                cp_async8(dst + off + (unsigned)(GK * GTP) * 8u, ok ? rb : a.J, ok && b_bytes);
            }
            ra += row_step;
            rb += row_step;
        }
        if (WEIGHTED && tid < GK) {
            const int64_t l = (int64_t)t * GK + tid;
            cp_async8(sbase + (unsigned)((t % GSTAGES) * STAGE + 2 * GK * GTP + tid) * 8u,
                      a.w + (l < a.n ? l : 0), l < a.n);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        pa += stage_step;
        pb += stage_step;
    };
    for (int t = 0; t < GSTAGES - 1; ++t) {
        if (t < ktiles) issue(t);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int t = 0; t < ktiles; ++t) {
        asm volatile("cp.async.wait_group %0;" ::"n"(GSTAGES - 2) : "memory");
        __syncthreads();                                   // stage t has landed, stage t-1 is free
        if (t + GSTAGES - 1 < ktiles) issue(t + GSTAGES - 1);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const double *As = smem + (t % GSTAGES) * STAGE;
        const double *Bs = As + GK * GTP;
        const double *Ws = Bs + GK * GTP;
#pragma unroll
        for (int k4 = 0; k4 < GK; k4 += 4) {
            // A fragment of mma tile i: J[l = k4 + fk][qa + wr + 8 i + fr]; B likewise with the columns
            double av[8], bv[4];
            const double *ar = As + (k4 + fk) * GTP + wr + fr;
            const double *br = Bs + (k4 + fk) * GTP + wc + fr;
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = ar[8 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = br[8 * j];
            if (WEIGHTED) {
                const double wk = Ws[k4 + fk];
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] *= wk;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j], av[i], bv[j]);
        }
    }
    // accumulator fragment of mma tile (i, j): row 8 i + fr, columns 8 j + 2 fk + (0, 1)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t row = qa + wr + 8 * i + fr;
        if (row >= a.m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = qb + wc + 8 * j + 2 * fk;
            if (col < a.m) a.G[row * a.ldg + col] = acc[i][j][0] * a.factor;
            if (col + 1 < a.m) a.G[row * a.ldg + col + 1] = acc[i][j][1] * a.factor;
        }
    }
}

// G[j][i] = G[i][j] for i < j, 32 x 32 tiles through shared memory
__global__ void mirror_kernel(double *G, int64_t m, int64_t ldg)
{
    __shared__ double t[32][33];
    const int64_t bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    const int64_t r0 = bi * 32, c0 = bj * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t row = r0 + r, col = c0 + threadIdx.x;
        t[r][threadIdx.x] = (row < m && col < m) ? G[row * ldg + col] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t row = c0 + r, col = r0 + threadIdx.x;       // transposed position
        if (row < m && col < m && col < row) G[row * ldg + col] = t[threadIdx.x][r];
    }
}

// y[l] = sum_q J[l][q] p[q]: one warp per row, fixed-order lane sums then a shuffle tree
__global__ void matvec_kernel(const double *J, int64_t n, int64_t m, int64_t ld, const double *p, double *y)
{
    const int64_t l = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (l >= n) return;
    const int lane = threadIdx.x & 31;
    const double *row = J + l * ld;
    double s = 0.0;
    for (int64_t q = lane; q < m; q += 32) s = fma(row[q], p[q], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[l] = s;
}

// part[c][q] = sum over the rows of chunk c of J[l][q] * (w[l] r[l]); thread = column
__global__ void matvec_t_kernel(const double *J, int64_t n, int64_t m, int64_t ld, const double *r,
                                const double *w, int64_t rows_per_chunk, double *part)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const int64_t l0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t l1 = l0 + rows_per_chunk < n ? l0 + rows_per_chunk : n;
    double s = 0.0;
    for (int64_t l = l0; l < l1; ++l) s = fma(J[l * ld + q], w ? w[l] * r[l] : r[l], s);
    part[(int64_t)blockIdx.y * m + q] = s;
}

__global__ void sum_chunks_kernel(const double *part, int64_t m, int chunks, double factor, double *g)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += part[(int64_t)c * m + q];
    g[q] = s * factor;
}

}  // namespace fb200

using namespace fb200;

extern "C" int fb200_gram(int device, const double *J, int64_t n, int64_t m, int64_t ld, const double *w,
                          double factor, double *G, int64_t ldg, double *ms)
{
    if (n < 0 || m < 0) return set_error(FB200_EINVAL, "negative matrix size");
    if (m > 0 && (!G || ldg < m)) return set_error(FB200_EINVAL, "G is NULL or ldg < m");
    if (n > 0 && m > 0 && (!J || ld < m)) return set_error(FB200_EINVAL, "J is NULL or ld < m");
    if (m == 0) return FB200_OK;
    DeviceScope scope(device);
    if (!scope.ok) return set_error(FB200_ECUDA, "cudaSetDevice failed");
    GramArgs a;
    a.J = J; a.w = w; a.n = n; a.m = m; a.ld = ld; a.factor = factor; a.G = G; a.ldg = ldg;
    a.tiles = (int)((m + GT - 1) / GT);
    const int64_t blocks = (int64_t)a.tiles * (a.tiles + 1) / 2;
    if (blocks > 2147483647ll) return set_error(FB200_EUNSUP, "matrix too large for one launch");
    const size_t smem = (size_t)GSTAGES * (2 * GK * GTP + GK) * sizeof(double);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms) {
        FB_CK(cudaEventCreate(&e0));
        FB_CK(cudaEventCreate(&e1));
        FB_CK(cudaEventRecord(e0, 0));
    }
    // 16-byte copies need every row of J to start on a 16-byte boundary and pairs inside the row
    const bool vec = (ld % 2 == 0) && (reinterpret_cast<uintptr_t>(J) % 16 == 0);
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<(unsigned)blocks, GTHREADS, smem>>>(a);
        return cudaGetLastError();
    };
    if (w) FB_CK(vec ? go(gram_kernel<true, true>) : go(gram_kernel<true, false>));
    else FB_CK(vec ? go(gram_kernel<false, true>) : go(gram_kernel<false, false>));
    FB_CK(cudaGetLastError());
    const unsigned mt = (unsigned)((m + 31) / 32);
    mirror_kernel<<<dim3(mt, mt), dim3(32, 8)>>>(G, m, ldg);
    FB_CK(cudaGetLastError());
    if (ms) {
        FB_CK(cudaEventRecord(e1, 0));
        FB_CK(cudaEventSynchronize(e1));
        float t = 0.f;
        FB_CK(cudaEventElapsedTime(&t, e0, e1));
        *ms = t;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    } else {
        FB_CK(cudaDeviceSynchronize());
    }
    return FB200_OK;
}

extern "C" int fb200_matvec(int device, const double *J, int64_t n, int64_t m, int64_t ld, const double *p,
                            double *y)
{
    if (n < 0 || m < 0 || (n > 0 && (!J || !y || ld < m)) || (m > 0 && !p))
        return set_error(FB200_EINVAL, "bad argument to fb200_matvec");
    if (n == 0) return FB200_OK;
    DeviceScope scope(device);
    if (!scope.ok) return set_error(FB200_ECUDA, "cudaSetDevice failed");
    matvec_kernel<<<(unsigned)((n + 7) / 8), 256>>>(J, n, m, ld, p, y);
    FB_CK(cudaGetLastError());
    FB_CK(cudaDeviceSynchronize());
    return FB200_OK;
}

extern "C" int fb200_matvec_t(int device, const double *J, int64_t n, int64_t m, int64_t ld, const double *r,
                              const double *w, double factor, double *g)
{
    if (n < 0 || m < 0 || (m > 0 && !g) || (n > 0 && m > 0 && (!J || !r || ld < m)))
        return set_error(FB200_EINVAL, "bad argument to fb200_matvec_t");
    if (m == 0) return FB200_OK;
    DeviceScope scope(device);
    if (!scope.ok) return set_error(FB200_ECUDA, "cudaSetDevice failed");
    // fixed chunks of rows, added in chunk order: reproducible
    const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(512, (n + 255) / 256));
    const int64_t rows_per_chunk = (std::max<int64_t>(n, 1) + chunks - 1) / chunks;
    double *part = nullptr;
    FB_CK(cudaMallocAsync((void **)&part, (size_t)chunks * m * sizeof(double), 0));
    matvec_t_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)chunks), 256>>>(J, n, m, ld, r, w,
                                                                                  rows_per_chunk, part);
    FB_CK(cudaGetLastError());
    sum_chunks_kernel<<<(unsigned)((m + 255) / 256), 256>>>(part, m, chunks, factor, g);
    FB_CK(cudaGetLastError());
    FB_CK(cudaFreeAsync(part, 0));
    FB_CK(cudaDeviceSynchronize());
    return FB200_OK;
}

// plain device buffers for callers without a tensor library (the ctypes binding)
extern "C" int fb200_device_alloc(int device, size_t bytes, void **ptr)
{
    if (!ptr) return set_error(FB200_EINVAL, "ptr is NULL");
    *ptr = nullptr;
    DeviceScope scope(device);
    if (!scope.ok) return set_error(FB200_ECUDA, "cudaSetDevice failed");
    FB_CK(cudaMalloc(ptr, bytes ? bytes : 1));
    return FB200_OK;
}

extern "C" int fb200_device_free(int device, void *ptr)
{
    DeviceScope scope(device);
    if (!scope.ok) return set_error(FB200_ECUDA, "cudaSetDevice failed");
    if (ptr) FB_CK(cudaFree(ptr));
    return FB200_OK;
}

// kind: 0 host -> device, 1 device -> host, 2 device -> device (synchronous)
extern "C" int fb200_copy(int device, void *dst, const void *src, size_t bytes, int kind)
{
    if (bytes && (!dst || !src)) return set_error(FB200_EINVAL, "NULL pointer");
    if (kind < 0 || kind > 2) return set_error(FB200_EINVAL, "kind must be 0, 1 or 2");
    DeviceScope scope(device);
    if (!scope.ok) return set_error(FB200_ECUDA, "cudaSetDevice failed");
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost
                                                                            : cudaMemcpyDeviceToDevice;
    FB_CK(cudaMemcpy(dst, src, bytes, k));
    return FB200_OK;
}
