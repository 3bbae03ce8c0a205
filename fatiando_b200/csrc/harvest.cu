// harvest.cu -- device-resident effect engine of the planting inversion
// (reference fatiando/gravmag/harvester.py).
//
// The reference keeps, on the host, one Python object per candidate cell with
// its "effect" -- prism.<field>(x, y, z, [cell], prop) for every data set
// (harvester.py:487-493, 720-724, 958-962) -- and, for EVERY candidate of
// EVERY accretion, forms predicted + effect and reduces it twice
// (_misfitfunc :447-456 and _shapefunc :434-444, called from _grow :412-431).
// Here the effects live in HBM (one row of `ntot` doubles per candidate, ntot =
// all data sets concatenated), and three small kernels do the work:
//
//   harvest_effect_kernel    effect rows of up to 8 new candidates, one launch
//                            per data set (the pair evaluators of the forward
//                            path, one cell at a time from a zero accumulator,
//                            then the unit factor -- what one reference call
//                            returns);
//   harvest_evaluate_kernel  block = (data set, candidate): misfit term
//                            sqrt(sum w r^2)/norm and shape-of-anomaly term
//                            ||alpha*obs - pred||, two passes over L2-resident
//                            rows, fixed-order block reductions (reproducible);
//   harvest_accrete_kernel   predicted += effect with the reference's float32
//                            storage of `predicted` (harvester.py:401:
//                            numpy.zeros(..., dtype='f'); p += e rounds to
//                            float32 after a float64 add).
//
// All of it is latency-bound by design (a few thousand points, a few hundred
// candidates); what it buys is that nothing but 2 doubles per candidate and
// data set crosses PCIe per accretion.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fatiando_b200.h"
#include "abi_internal.h"
#include "harvest.cuh"

using namespace fb200;

namespace {

constexpr int HV_EVAL_THREADS = 256;

// sum over the block in a fixed order (shuffle tree inside warps, then warp 0)
__device__ double block_sum(double v, double *scratch)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                   // scratch may still be read from a previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < HV_EVAL_THREADS / 32) ? scratch[threadIdx.x] : 0.0;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) scratch[0] = t;
    }
    __syncthreads();
    return scratch[0];
}

struct EvaluateArgs {
    const double *obs, *w;
    const float *pred;
    const int64_t *off;                // ndata + 1 offsets into the concatenated arrays
    const double *norm;                // ||observed|| per data set
    double *const *chunks;             // effect store: chunk pointers
    int64_t chunk_rows, ntot;
    const int64_t *rows;               // candidate rows, -1 = no effect (the current estimate)
    int ndata;
    double *misfit, *shape;            // [candidate][data set]
};

__global__ void __launch_bounds__(HV_EVAL_THREADS) harvest_evaluate_kernel(const EvaluateArgs a)
{
    __shared__ double scratch[HV_EVAL_THREADS / 32];
    const int d = blockIdx.x;
    const int64_t cand = blockIdx.y;
    const int64_t row = a.rows[cand];
    const double *e = nullptr;
    if (row >= 0) e = a.chunks[row / a.chunk_rows] + (row % a.chunk_rows) * a.ntot;
    const int64_t l0 = a.off[d], l1 = a.off[d + 1];
    const double norm = a.norm[d];
    // pass 1: weighted residual energy (harvester.py:453-455) and <observed, predicted>
    double s_wrr = 0.0, s_op = 0.0;
    for (int64_t l = l0 + threadIdx.x; l < l1; l += HV_EVAL_THREADS) {
        const double o = a.obs[l];
        const double p = (double)a.pred[l] + (e ? e[l] : 0.0);
        const double r = o - p;
        s_wrr += (a.w[l] * r) * r;
        s_op += o * p;
    }
    s_wrr = block_sum(s_wrr, scratch);
    s_op = block_sum(s_op, scratch);
    // pass 2: ||alpha*observed - predicted|| with alpha = <o, p>/||o||^2 (harvester.py:441-443)
    const double alpha = s_op / (norm * norm);
    double s_sh = 0.0;
    for (int64_t l = l0 + threadIdx.x; l < l1; l += HV_EVAL_THREADS) {
        const double p = (double)a.pred[l] + (e ? e[l] : 0.0);
        const double t = alpha * a.obs[l] - p;
        s_sh += t * t;
    }
    s_sh = block_sum(s_sh, scratch);
    if (threadIdx.x == 0) {
        a.misfit[cand * a.ndata + d] = sqrt(s_wrr) / norm;
        a.shape[cand * a.ndata + d] = sqrt(s_sh);
    }
}

__global__ void harvest_accrete_kernel(float *pred, const double *e, int64_t n)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n) pred[l] = (float)((double)pred[l] + e[l]);
}

}  // namespace

struct fb200_harvest {
    int device = 0;
    int ndata = 0;
    int64_t ntot = 0;
    std::vector<int64_t> off;          // ndata + 1
    std::vector<int> field;
    std::vector<double> fvec;          // 3 per data set
    std::vector<double> scale;
    double *d_xyz = nullptr;           // x | y | z, ntot each
    double *d_obs = nullptr, *d_w = nullptr, *d_norm = nullptr;
    float *d_pred = nullptr;
    int64_t *d_off = nullptr;
    std::vector<double *> chunks;      // effect store
    double **d_chunks = nullptr;
    size_t d_chunks_cap = 0;
    int64_t chunk_rows = 0;
    int64_t *d_rows = nullptr;  size_t rows_cap = 0;
    double *d_res = nullptr;    size_t res_cap = 0;
    // pinned staging of the per-call traffic (candidate rows in, 2 doubles per candidate and
    // data set out): one async copy each way instead of staged pageable copies
    int64_t *h_rows = nullptr;
    double *h_res = nullptr;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    std::mutex lock;
};

static double *row_ptr(const fb200_harvest *h, int64_t row)
{
    return h->chunks[(size_t)(row / h->chunk_rows)] + (row % h->chunk_rows) * h->ntot;
}

// make sure rows [0, row] exist in the effect store
static int reserve_row(fb200_harvest *h, int64_t row)
{
    bool grew = false;
    while ((int64_t)h->chunks.size() * h->chunk_rows <= row) {
        double *p = nullptr;
        FB_CK(cudaMallocAsync((void **)&p, (size_t)h->chunk_rows * h->ntot * sizeof(double), h->stream));
        h->chunks.push_back(p);
        grew = true;
    }
    if (grew) {
        if (h->chunks.size() > h->d_chunks_cap) {
            if (h->d_chunks) cudaFreeAsync(h->d_chunks, h->stream);
            h->d_chunks_cap = std::max<size_t>(64, 2 * h->chunks.size());
            FB_CK(cudaMallocAsync((void **)&h->d_chunks, h->d_chunks_cap * sizeof(double *), h->stream));
        }
        // the pointer table is tiny and changes rarely: synchronous staging is fine
        FB_CK(cudaStreamSynchronize(h->stream));
        FB_CK(cudaMemcpy(h->d_chunks, h->chunks.data(), h->chunks.size() * sizeof(double *),
                         cudaMemcpyHostToDevice));
    }
    return FB200_OK;
}

extern "C" int fb200_harvest_create(int device, int ndata, const int64_t *sizes, const int *fields,
                                    const double *x, const double *y, const double *z,
                                    const double *observed, const double *weights,
                                    const double *fvec, const double *scale, fb200_harvest **out)
{
    if (!out) return set_error(FB200_EINVAL, "out is NULL");
    *out = nullptr;
    if (ndata < 1 || !sizes || !fields || !x || !y || !z || !observed || !weights || !scale)
        return set_error(FB200_EINVAL, "fb200_harvest_create: NULL argument or no data set");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(FB200_ECUDA, "no CUDA device visible (there is no CPU path)");
    }
    if (device < 0 || device >= count) return set_error(FB200_EINVAL, "device index out of range");
    DeviceScope g(device);
    if (!g.ok) return set_error(FB200_ECUDA, "cudaSetDevice failed");
    fb200_harvest *h = new fb200_harvest;
    h->device = device;
    h->ndata = ndata;
    h->off.assign(1, 0);
    for (int d = 0; d < ndata; ++d) {
        if (sizes[d] < 1 || fields[d] < 0 || fields[d] >= F_COUNT) {
            delete h;
            return set_error(FB200_EINVAL, "fb200_harvest_create: empty data set or unknown field");
        }
        if (fields[d] == F_TF && !fvec) {
            delete h;
            return set_error(FB200_EINVAL, "total-field data need the inducing-field direction");
        }
        h->off.push_back(h->off.back() + sizes[d]);
        h->field.push_back(fields[d]);
        for (int k = 0; k < 3; ++k) h->fvec.push_back(fvec ? fvec[3 * d + k] : 0.0);
        h->scale.push_back(scale[d]);
    }
    h->ntot = h->off.back();
    // ~32 MiB of effect rows per chunk
    h->chunk_rows = std::max<int64_t>(8, (int64_t)(32u << 20) / (h->ntot * 8));
    std::vector<double> norm((size_t)ndata);
    for (int d = 0; d < ndata; ++d) {
        // numpy.linalg.norm(data) (harvester.py:653): sqrt of the plain dot product
        double s = 0.0;
        for (int64_t l = h->off[d]; l < h->off[d + 1]; ++l) s += observed[l] * observed[l];
        norm[d] = std::sqrt(s);
    }
    const size_t nb = (size_t)h->ntot * sizeof(double);
#define HCK(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            fb200_harvest_destroy(h);                                                     \
            return set_error(FB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
        }                                                                                 \
    } while (0)
    HCK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    HCK(cudaMallocAsync((void **)&h->d_xyz, 3 * nb, h->stream));
    HCK(cudaMallocAsync((void **)&h->d_obs, nb, h->stream));
    HCK(cudaMallocAsync((void **)&h->d_w, nb, h->stream));
    HCK(cudaMallocAsync((void **)&h->d_norm, (size_t)ndata * sizeof(double), h->stream));
    HCK(cudaMallocAsync((void **)&h->d_off, (size_t)(ndata + 1) * sizeof(int64_t), h->stream));
    HCK(cudaMallocAsync((void **)&h->d_pred, (size_t)h->ntot * sizeof(float), h->stream));
    HCK(cudaMemcpyAsync(h->d_xyz, x, nb, cudaMemcpyHostToDevice, h->stream));
    HCK(cudaMemcpyAsync(h->d_xyz + h->ntot, y, nb, cudaMemcpyHostToDevice, h->stream));
    HCK(cudaMemcpyAsync(h->d_xyz + 2 * h->ntot, z, nb, cudaMemcpyHostToDevice, h->stream));
    HCK(cudaMemcpyAsync(h->d_obs, observed, nb, cudaMemcpyHostToDevice, h->stream));
    HCK(cudaMemcpyAsync(h->d_w, weights, nb, cudaMemcpyHostToDevice, h->stream));
    HCK(cudaMemcpyAsync(h->d_norm, norm.data(), (size_t)ndata * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    HCK(cudaMemcpyAsync(h->d_off, h->off.data(), (size_t)(ndata + 1) * sizeof(int64_t),
                        cudaMemcpyHostToDevice, h->stream));
    HCK(cudaMemsetAsync(h->d_pred, 0, (size_t)h->ntot * sizeof(float), h->stream));
    HCK(cudaStreamSynchronize(h->stream));
#undef HCK
    *out = h;
    return FB200_OK;
}

extern "C" int fb200_harvest_destroy(fb200_harvest *h)
{
    if (!h) return FB200_OK;
    DeviceScope g(h->device);
    cudaStream_t s = h->stream;
    if (s) cudaStreamSynchronize(s);
    for (double *p : h->chunks) cudaFree(p);
    cudaFree(h->d_chunks); cudaFree(h->d_xyz); cudaFree(h->d_obs); cudaFree(h->d_w);
    cudaFree(h->d_norm); cudaFree(h->d_off); cudaFree(h->d_pred); cudaFree(h->d_rows); cudaFree(h->d_res);
    if (h->h_rows) cudaFreeHost(h->h_rows);
    if (h->h_res) cudaFreeHost(h->h_res);
    if (s) cudaStreamDestroy(s);
    delete h;
    return FB200_OK;
}

extern "C" int fb200_harvest_effects(fb200_harvest *h, int ncells, const double *bounds,
                                     const double *density, const double *magnetization,
                                     const int *mag_kind, const int64_t *rows, uint32_t flags)
{
    if (!h) return set_error(FB200_EINVAL, "harvest handle is NULL");
    if (ncells < 0 || (ncells > 0 && (!bounds || !rows)))
        return set_error(FB200_EINVAL, "fb200_harvest_effects: NULL argument");
    if (flags & ~FB200_REFERENCE_ORDER) return set_error(FB200_EUNSUP, "unsupported flag for fb200_harvest_effects");
    std::lock_guard<std::mutex> hold(h->lock);
    DeviceScope g(h->device);
    if (!g.ok) return set_error(FB200_ECUDA, "cudaSetDevice failed");
    for (int c = 0; c < ncells; ++c) {
        if (rows[c] < 0) return set_error(FB200_EINVAL, "negative effect row");
        int rc = reserve_row(h, rows[c]);
        if (rc != FB200_OK) return rc;
    }
    for (int c0 = 0; c0 < ncells; c0 += HV_MAX_CELLS) {
        const int nc = std::min(HV_MAX_CELLS, ncells - c0);
        for (int d = 0; d < h->ndata; ++d) {
            EffectArgs a;
            std::memset(&a, 0, sizeof(a));
            a.xp = h->d_xyz + h->off[d];
            a.yp = h->d_xyz + h->ntot + h->off[d];
            a.zp = h->d_xyz + 2 * h->ntot + h->off[d];
            a.n = h->off[d + 1] - h->off[d];
            a.scale = h->scale[d];
            a.ncells = nc;
            const int f = h->field[d];
            const bool magnetic = f >= F_TF;
            for (int k = 0; k < 3; ++k) a.fvec[k] = h->fvec[3 * d + k];
            for (int c = 0; c < nc; ++c) {
                const int q = c0 + c;
                for (int k = 0; k < 6; ++k) a.b[c][k] = bounds[6 * q + k] + 0.0;   // -0.0 -> +0.0
                a.out[c] = row_ptr(h, rows[q]) + h->off[d];
                if (!magnetic) {
                    // harvester.py:690-694: no 'density' in props -> zeros
                    a.has[c] = density && !std::isnan(density[q]);
                    a.p[c][0] = a.has[c] ? density[q] : 0.0;
                } else {
                    const int kind = mag_kind ? mag_kind[q] : 0;
                    a.has[c] = kind != 0 && magnetization != nullptr;
                    if (kind == 1) {
                        // scalar magnetization: along the data set's inducing field
                        // (prism.py:641-645 through harvester.py:958-962)
                        if (f != F_TF)
                            return set_error(FB200_EPROP, "bx/by/bz need a 3-component magnetization");
                        for (int k = 0; k < 3; ++k) a.p[c][k] = magnetization[3 * q] * a.fvec[k];
                    } else if (kind == 2) {
                        for (int k = 0; k < 3; ++k) a.p[c][k] = magnetization[3 * q + k];
                    }
                }
            }
            cudaError_t e = (flags & FB200_REFERENCE_ORDER) ? launch_harvest_effect_exact(f, a, h->stream)
                                                            : launch_harvest_effect_fast(f, a, h->stream);
            if (e != cudaSuccess)
                return set_error(FB200_ECUDA, std::string("harvest_effect_kernel: ") + cudaGetErrorString(e));
            h->launches += 1;
        }
    }
    return FB200_OK;
}

extern "C" int fb200_harvest_accrete(fb200_harvest *h, int64_t row)
{
    if (!h) return set_error(FB200_EINVAL, "harvest handle is NULL");
    std::lock_guard<std::mutex> hold(h->lock);
    if (row < 0 || row >= (int64_t)h->chunks.size() * h->chunk_rows)
        return set_error(FB200_EINVAL, "effect row out of range");
    DeviceScope g(h->device);
    harvest_accrete_kernel<<<(unsigned)((h->ntot + 255) / 256), 256, 0, h->stream>>>(
        h->d_pred, row_ptr(h, row), h->ntot);
    FB_CK(cudaGetLastError());
    h->launches += 1;
    return FB200_OK;
}

extern "C" int fb200_harvest_evaluate(fb200_harvest *h, int64_t ncand, const int64_t *rows,
                                      double *misfit, double *shape)
{
    if (!h) return set_error(FB200_EINVAL, "harvest handle is NULL");
    if (ncand < 0 || (ncand > 0 && (!rows || !misfit || !shape)))
        return set_error(FB200_EINVAL, "fb200_harvest_evaluate: NULL argument");
    if (ncand == 0) return FB200_OK;
    if (ncand > 65535) return set_error(FB200_EINVAL, "more than 65535 candidates in one call");
    std::lock_guard<std::mutex> hold(h->lock);
    const int64_t cap_rows = (int64_t)h->chunks.size() * h->chunk_rows;
    for (int64_t c = 0; c < ncand; ++c)
        if (rows[c] >= cap_rows) return set_error(FB200_EINVAL, "effect row out of range");
    DeviceScope g(h->device);
    if ((size_t)ncand > h->rows_cap) {
        if (h->d_rows) cudaFreeAsync(h->d_rows, h->stream);
        if (h->h_rows) { cudaStreamSynchronize(h->stream); cudaFreeHost(h->h_rows); h->h_rows = nullptr; }
        h->rows_cap = std::max<size_t>(256, 2 * (size_t)ncand);
        FB_CK(cudaMallocAsync((void **)&h->d_rows, h->rows_cap * sizeof(int64_t), h->stream));
        FB_CK(cudaMallocHost((void **)&h->h_rows, h->rows_cap * sizeof(int64_t)));
    }
    const size_t nres = 2 * (size_t)ncand * h->ndata;
    if (nres > h->res_cap) {
        if (h->d_res) cudaFreeAsync(h->d_res, h->stream);
        if (h->h_res) { cudaStreamSynchronize(h->stream); cudaFreeHost(h->h_res); h->h_res = nullptr; }
        h->res_cap = std::max<size_t>(1024, 2 * nres);
        FB_CK(cudaMallocAsync((void **)&h->d_res, h->res_cap * sizeof(double), h->stream));
        FB_CK(cudaMallocHost((void **)&h->h_res, h->res_cap * sizeof(double)));
    }
    std::memcpy(h->h_rows, rows, (size_t)ncand * sizeof(int64_t));
    FB_CK(cudaMemcpyAsync(h->d_rows, h->h_rows, (size_t)ncand * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    EvaluateArgs a;
    a.obs = h->d_obs; a.w = h->d_w; a.pred = h->d_pred; a.off = h->d_off; a.norm = h->d_norm;
    a.chunks = h->d_chunks; a.chunk_rows = h->chunk_rows; a.ntot = h->ntot; a.rows = h->d_rows;
    a.ndata = h->ndata; a.misfit = h->d_res; a.shape = h->d_res + (size_t)ncand * h->ndata;
    dim3 grid((unsigned)h->ndata, (unsigned)ncand);
    harvest_evaluate_kernel<<<grid, HV_EVAL_THREADS, 0, h->stream>>>(a);
    FB_CK(cudaGetLastError());
    h->launches += 1;
    const size_t half = (size_t)ncand * h->ndata;
    FB_CK(cudaMemcpyAsync(h->h_res, h->d_res, 2 * half * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FB_CK(cudaStreamSynchronize(h->stream));
    std::memcpy(misfit, h->h_res, half * sizeof(double));
    std::memcpy(shape, h->h_res + half, half * sizeof(double));
    return FB200_OK;
}

extern "C" int fb200_harvest_predicted(fb200_harvest *h, float *out)
{
    if (!h || !out) return set_error(FB200_EINVAL, "fb200_harvest_predicted: NULL argument");
    std::lock_guard<std::mutex> hold(h->lock);
    DeviceScope g(h->device);
    FB_CK(cudaMemcpyAsync(out, h->d_pred, (size_t)h->ntot * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    FB_CK(cudaStreamSynchronize(h->stream));
    return FB200_OK;
}

extern "C" int fb200_harvest_effect(fb200_harvest *h, int64_t row, double *out)
{
    if (!h || !out) return set_error(FB200_EINVAL, "fb200_harvest_effect: NULL argument");
    std::lock_guard<std::mutex> hold(h->lock);
    if (row < 0 || row >= (int64_t)h->chunks.size() * h->chunk_rows)
        return set_error(FB200_EINVAL, "effect row out of range");
    DeviceScope g(h->device);
    FB_CK(cudaMemcpyAsync(out, row_ptr(h, row), (size_t)h->ntot * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FB_CK(cudaStreamSynchronize(h->stream));
    return FB200_OK;
}

extern "C" int64_t fb200_harvest_launches(const fb200_harvest *h) { return h ? h->launches : -1; }
