// hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the device math headers of fatiando_b200/csrc (prism_exact.cuh,
// prism_fast.cuh) for the HOST, so that the numerics of the fused fast path
// (collapsed log/atan sums, custom log_ratio / atan) can be compared with the
// oracle in a container that has no GPU.  Only tests/ builds and loads this;
// fatiando_b200 never does -- the product has no CPU path.
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (tests/hostsim/Makefile)
#include <cstddef>
#include <cstdint>
#include "../../fatiando_b200/csrc/prism_masks.h"
#include "../../fatiando_b200/csrc/prism_fast.cuh"
#include "../../fatiando_b200/csrc/prism_mesh.cuh"
#include "../../fatiando_b200/csrc/prism_face.cuh"
#include "../../fatiando_b200/csrc/polyprism.cuh"
#include "../../fatiando_b200/csrc/sphere.cuh"

using namespace fb200;

template <uint32_t MASK, class Eval>
static void run(const double *xp, const double *yp, const double *zp, size_t n, size_t m,
                const double *const *b, const double *const *p, const double *f, double *out)
{
    constexpr int NC = popc(MASK);
    const double fv[3] = {f[0], f[1], f[2]};
    for (size_t l = 0; l < n; ++l) {
        double acc[NC];
        for (int c = 0; c < NC; ++c) acc[c] = 0.0;
        for (size_t q = 0; q < m; ++q)
            Eval::template pair<MASK>(acc, xp[l], yp[l], zp[l], b[0][q], b[1][q], b[2][q], b[3][q],
                                      b[4][q], b[5][q], p[0][q], p[1] ? p[1][q] : 0.0,
                                      p[2] ? p[2][q] : 0.0, fv);
        for (int c = 0; c < NC; ++c) out[(size_t)c * n + l] = acc[c];
    }
}

extern "C" int hostsim_forward(uint32_t mask, int exact, const double *xp, const double *yp,
                               const double *zp, size_t n, size_t m, const double *const *bounds,
                               const double *const *props, const double *fvec, double *out)
{
#define CASE(M)                                                                         \
    if (mask == (M)) {                                                                  \
        if (exact) run<(M), EvalExact>(xp, yp, zp, n, m, bounds, props, fvec, out);     \
        else run<(M), EvalFast>(xp, yp, zp, n, m, bounds, props, fvec, out);            \
        return 0;                                                                       \
    }
    FB200_GRAVITY_SETS(CASE)
    FB200_MAGNETIC_SETS(CASE)
#undef CASE
    return -1;
}

// node-sharing evaluation of a regular mesh, same walk as mesh_forward_kernel
template <uint32_t MASK>
static void run_mesh(const double *xp, const double *yp, const double *zp, size_t n, int nxp, int nyp,
                     int nzp, const double *xn, const double *yn, const double *zn,
                     const double *const *w, const double *shv, const double *f, double *out)
{
    constexpr int NC = popc(MASK);
    constexpr bool MAG = (MASK & MAGNETIC_MASK) != 0;
    const double fv[3] = {f[0], f[1], f[2]}, sh[3] = {shv[0], shv[1], shv[2]};
    for (size_t l = 0; l < n; ++l) {
        double acc[NC];
        for (int c = 0; c < NC; ++c) acc[c] = 0.0;
        size_t g = 0;
        for (int b = 0; b < nyp; ++b)
            for (int a = 0; a < nxp; ++a) {
                const double X = xn[a] - xp[l], Y = yn[b] - yp[l];
                const double sxy = X * X + Y * Y;
                for (int c = 0; c < nzp; ++c, ++g) {
                    const double w0 = w[0][g], w1 = MAG ? w[1][g] : 0.0, w2 = MAG ? w[2][g] : 0.0;
                    if (w0 == 0.0 && w1 == 0.0 && w2 == 0.0) continue;
                    node_eval<MASK>(acc, X, Y, zn[c] - zp[l], sxy, column_bits(X, Y), w0, w1, w2, fv, sh,
                                    fm::tab_ref());
                }
            }
        for (int c = 0; c < NC; ++c) out[(size_t)c * n + l] = acc[c];
    }
}

extern "C" int hostsim_mesh_forward(uint32_t mask, const double *xp, const double *yp,
                                    const double *zp, size_t n, int nxp, int nyp, int nzp,
                                    const double *xn, const double *yn, const double *zn,
                                    const double *const *w, const double *sh, const double *fvec,
                                    double *out)
{
#define CASE(M)                                                                          \
    if (mask == (M)) {                                                                   \
        run_mesh<(M)>(xp, yp, zp, n, nxp, nyp, nzp, xn, yn, zn, w, sh, fvec, out);       \
        return 0;                                                                        \
    }
    FB200_GRAVITY_SETS(CASE)
    FB200_MAGNETIC_SETS(CASE)
#undef CASE
    return -1;
}

// surface form of a relief: top faces (EvalFace) + merged reference-plane nodes (EvalNode)
template <uint32_t MASK>
static int run_surface(const double *xp, const double *yp, const double *zp, size_t n, size_t nf,
                       const double *const *face, size_t nn, const double *const *node,
                       const double *cell, double *out)
{
    constexpr int NC = popc(MASK);
    const double f[3] = {0.0, 0.0, 0.0};
    int flag = 0;
    for (size_t l = 0; l < n; ++l) {
        double acc[NC], accn[NC];
        for (int c = 0; c < NC; ++c) acc[c] = accn[c] = 0.0;
        for (size_t q = 0; q < nf; ++q)
            EvalFace::pair<MASK>(acc, xp[l], yp[l], zp[l], face[0][q], face[1][q], face[2][q],
                                 face[3][q], face[4][q], face[5][q], face[6][q], 0.0, 0.0, f);
        for (size_t q = 0; q < nn; ++q)
            EvalNode::pair_flag<MASK>(accn, xp[l], yp[l], zp[l], node[0][q], 0.00001 * cell[0],
                                      node[1][q], 0.00001 * cell[1], node[2][q], 0.0, node[3][q],
                                      0.0, 0.0, f, &flag);
        for (int c = 0; c < NC; ++c) out[(size_t)c * n + l] = acc[c] + accn[c];
    }
    return flag;
}

// returns the flag (1: some point needs the per-cell path), or -1 for an unknown mask
extern "C" int hostsim_surface_forward(uint32_t mask, const double *xp, const double *yp,
                                       const double *zp, size_t n, size_t nf,
                                       const double *const *face, size_t nn,
                                       const double *const *node, const double *cell, double *out)
{
#define CASE(M) if (mask == (M)) return run_surface<(M)>(xp, yp, zp, n, nf, face, nn, node, cell, out);
    FB200_GRAVITY_SETS(CASE)
#undef CASE
    return -1;
}

// sibling families: rows through an evaluator policy (EvalPolyEdge, EvalSphere)
template <uint32_t MASK, class Eval>
static void run_rows(const double *xp, const double *yp, const double *zp, size_t n, size_t m,
                     const double *const *b, const double *const *p, const double *f, double *out)
{
    constexpr int NC = popc(MASK);
    const double fv[3] = {f[0], f[1], f[2]};
    for (size_t l = 0; l < n; ++l) {
        double acc[NC];
        for (int c = 0; c < NC; ++c) acc[c] = 0.0;
        for (size_t q = 0; q < m; ++q)
            Eval::template pair<MASK>(acc, xp[l], yp[l], zp[l], b[0][q], b[1][q], b[2][q], b[3][q],
                                      b[4][q], b[5][q], p[0][q], p[1] ? p[1][q] : 0.0,
                                      p[2] ? p[2][q] : 0.0, fv);
        for (int c = 0; c < NC; ++c) out[(size_t)c * n + l] = acc[c];
    }
}

// kind 0: polygon edges, 1: spheres
extern "C" int hostsim_sibling_forward(int kind, uint32_t mask, const double *xp, const double *yp,
                                       const double *zp, size_t n, size_t m, const double *const *b,
                                       const double *const *p, const double *fvec, double *out)
{
#define CASE(M)                                                                              \
    if (mask == (M)) {                                                                       \
        if (kind == 0) run_rows<(M), EvalPolyEdge>(xp, yp, zp, n, m, b, p, fvec, out);       \
        else run_rows<(M), EvalSphere>(xp, yp, zp, n, m, b, p, fvec, out);                   \
        return 0;                                                                            \
    }
    CASE(0x008) CASE(0x010) CASE(0x020) CASE(0x040) CASE(0x080) CASE(0x100) CASE(0x200)
    CASE(0x3F0) CASE(0x3F8)
    FB200_MAGNETIC_SETS(CASE)
#undef CASE
    return -1;
}

extern "C" void hostsim_log_ratio(const double *num, const double *den, size_t n, double *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = fm::log_ratio(num[i], den[i]);
}
extern "C" void hostsim_atan_rhp(const double *im, const double *re, size_t n, double *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = fm::atan_rhp(im[i], re[i]);
}
extern "C" void hostsim_div(const double *a, const double *b, size_t n, double *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = fm::div_fast(a[i], b[i]);
}
extern "C" void hostsim_log_tab(const double *a, const double *unused, size_t n, double *out)
{
    (void)unused;
    for (size_t i = 0; i < n; ++i) out[i] = fm::log_tab(a[i]);
}
extern "C" void hostsim_atan_tab(const double *im, const double *re, size_t n, double *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = fm::atan_ratio_tab(im[i], re[i]);
}
