/*
 * fatiando_b200.h -- C ABI of the B200-native right-rectangular-prism
 * potential-field forward model (libfatiando_b200.so).
 *
 * This is the drop-in boundary for ONE path of fatiando/fatiando: the native
 * module `fatiando.gravmag._prism` (reference: fatiando/gravmag/_prism.pyx,
 * imported by fatiando/gravmag/prism.py:92-95) and the per-cell sensitivity
 * build that callers layer on top of it (harvester.py:720-724,958-962,
 * imaging.py:116-117, eqlayer.py:100-111).  Plain C types only: pointers,
 * sizes, doubles.  Every entry point returns FB200_OK (0) or a negative
 * FB200_E* code and never aborts; fb200_last_error() gives the message of the
 * calling thread's last failure.
 *
 * Conventions (identical to the reference): x north, y east, z down, SI in;
 * all arrays are IEEE binary64; a prism is the six bounds x1,x2,y1,y2,z1,z2.
 * There is no CPU fallback anywhere behind this header: without a usable
 * CUDA device every compute entry point fails with FB200_ECUDA.
 */
#ifndef FATIANDO_B200_H
#define FATIANDO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_ABI_VERSION 1

/* ---- return codes ------------------------------------------------------ */
#define FB200_OK        0
#define FB200_EINVAL   -1   /* bad argument (null pointer, negative size, ...) */
#define FB200_ECUDA    -2   /* CUDA runtime/driver error, no device, OOM      */
#define FB200_EPROP    -3   /* field needs a property the model does not hold */
#define FB200_EUNSUP   -4   /* unsupported combination of flags               */

/* ---- field ids; bit (1u << id) in a field mask -------------------------
 * Order = the 14 accumulators of _prism.pyx: potential :454, gx :196,
 * gy :223, gz :250, gxx :277, gxy :304, gxz :336, gyy :368, gyz :395,
 * gzz :427, tf :72, bx :106, by :136, bz :166.                              */
enum fb200_field {
    FB200_POTENTIAL = 0, FB200_GX = 1, FB200_GY = 2, FB200_GZ = 3,
    FB200_GXX = 4, FB200_GXY = 5, FB200_GXZ = 6, FB200_GYY = 7,
    FB200_GYZ = 8, FB200_GZZ = 9, FB200_TF = 10, FB200_BX = 11,
    FB200_BY = 12, FB200_BZ = 13, FB200_NFIELDS = 14
};
#define FB200_GRAVITY_MASK  0x03FFu
#define FB200_MAGNETIC_MASK 0x3C00u

/* ---- flags ------------------------------------------------------------- */
/* xp/yp/zp/out (and init) are device pointers on the model's device        */
#define FB200_DEVICE_POINTERS  (1u << 0)
/* evaluate in the reference's exact operation order (corner by corner,
 * running accumulator, libdevice log/atan2) instead of the fused fast path  */
#define FB200_REFERENCE_ORDER  (1u << 1)
/* jacobian only: write the transpose, out[cell * ld + point]
 * (the layout imaging.py:116-117 builds)                                    */
#define FB200_TRANSPOSED       (1u << 2)
/* forward only: never split the prism list over thread blocks, so that every
 * point sums the prisms strictly in list order whatever the number of points
 * (results are then bit-identical between a full run and any observation
 * shard of it; small problems lose occupancy)                               */
#define FB200_NO_PRISM_SPLIT   (1u << 3)
/* forward and jacobian: evaluate cell by cell even when a regular mesh or a
 * relief surface is attached (fb200_model_attach_mesh / _attach_surface)    */
#define FB200_PER_CELL         (1u << 4)

typedef struct fb200_model fb200_model;

/* ---- housekeeping ------------------------------------------------------ */
int         fb200_abi_version(void);
const char *fb200_last_error(void);
/* number of CUDA devices visible to this process (0 and FB200_ECUDA if none) */
int         fb200_device_count(int *count);
/* name / SM count / memory of a device, for logs                            */
int         fb200_device_info(int device, char *name, int name_len,
                              int *sm_count, int64_t *total_mem_bytes);

/* ---- the mesh flattener's upload: SoA prism data, copied to HBM once ----
 * Replaces the per-prism attribute reads of prism.py:138-140 (bounds) and
 * :134-137 / :650-655 (properties).  x1..z2 are host arrays of m doubles.
 * density may be NULL (model usable for magnetic fields only, or with a
 * density override); mx,my,mz are all NULL or all non-NULL.  Prisms are kept
 * in the order given; that order is the summation order.                     */
int fb200_model_create(int device, int64_t m,
                       const double *x1, const double *x2,
                       const double *y1, const double *y2,
                       const double *z1, const double *z2,
                       const double *density,
                       const double *mx, const double *my, const double *mz,
                       fb200_model **out);
/* Sibling model family: spheres (fatiando/gravmag/sphere.py:45-746).  Sphere q has
 * centre (xc, yc, zc), `radius`, and density and/or magnetization like a prism.
 * The handle works with fb200_forward and fb200_jacobian for the fields
 * sphere.py offers -- FB200_GZ, FB200_GXX..FB200_GZZ, FB200_TF, FB200_BX/BY/BZ
 * (anything else: FB200_EUNSUP) -- with the same overrides, unit factors,
 * layouts and flags; FB200_REFERENCE_ORDER is accepted and changes nothing
 * (there is one evaluator, in the reference's operation order).              */
int fb200_model_create_spheres(int device, int64_t m,
                               const double *xc, const double *yc, const double *zc,
                               const double *radius, const double *density,
                               const double *mx, const double *my, const double *mz,
                               fb200_model **out);
/* Sibling model family: polygonal prisms (fatiando/gravmag/polyprism.py:19-1076,
 * Plouff 1976).  Prism q has the polygon vertices
 * (vx, vy)[first_vertex[q] .. first_vertex[q+1]) (clockwise, like the reference
 * asks), top z1[q], bottom z2[q] and density and/or magnetization; first_vertex has
 * nprisms + 1 entries.  The handle works with fb200_forward for the fields
 * polyprism.py offers (FB200_GZ, FB200_GXX..FB200_GZZ, FB200_TF, FB200_BX/BY/BZ),
 * with pmag_override for tf (polyprism.py:19); every polygon edge is one staged
 * row.  Jacobian and adjoint are not offered (FB200_EUNSUP).                    */
int fb200_model_create_polyprisms(int device, int64_t nprisms, const int64_t *first_vertex,
                                  const double *vx, const double *vy,
                                  const double *z1, const double *z2, const double *density,
                                  const double *mx, const double *my, const double *mz,
                                  fb200_model **out);
/* Regular-mesh acceleration (optional).  Declares that the model's prisms are
 * the unmasked cells of a regular nx*ny*nz PrismMesh (mesh.py:495-665) and hands
 * over its node-sharing form: node coordinates xn[nx+1], yn[ny+1], zn[nz+1]
 * (the reference's own bound expressions, mesh.py:629-634), the cell size, and
 * per-node weights = the 2x2x2 signed sum of the property of the <= 8 cells
 * around each node (index (b*(nx+1) + a)*(nz+1) + c, z fastest; masked cells
 * count as zero).  fb200_forward then evaluates the per-corner kernel once per
 * NODE instead of eight times per cell -- same arithmetic per evaluation,
 * ~2.5-3.5x less of it -- unless a property override, FB200_REFERENCE_ORDER or
 * FB200_PER_CELL is given.  The node weights may all be NULL (geometry only).
 * fb200_jacobian uses the geometry when the model's prisms are ALL cells of the
 * mesh in mesh order: every (point, node) kernel value is computed once and
 * each entry formed as the signed sum of its eight node values; fb200_adjoint
 * accumulates per NODE and differences the node sums into cells.              */
int fb200_model_attach_mesh(fb200_model *model, int nx, int ny, int nz,
                            const double *xn, const double *yn, const double *zn,
                            const double *cell_size /* 3 */,
                            const double *w_density /* nodes, or NULL */,
                            const double *w_mx, const double *w_my,
                            const double *w_mz /* nodes each, or all NULL */);
/* Relief acceleration (optional).  Declares that the model's prisms are the
 * cells of a PrismRelief (mesh.py:391-492) on a tight regular grid and hands
 * over its surface form: cell c = rho_c * (G_c(ref) - G_c(z_c)), G = the
 * 4-corner alternating sum over one horizontal face.  Faces: one per cell,
 * x1,x2,y1,y2 (the reference's own bounds), z_face = z_c, z_other = ref,
 * w_face = -rho_c (rho_c = the stored density, negated where z_c > ref --
 * undoing addprop's sign rule, mesh.py:484-488).  Nodes: the reference-plane
 * corners merged over the <= 4 cells that share them, weight = signed 2x2 sum
 * of rho; only non-zero ones need to be listed (for a uniform density, the
 * rim of the grid).  cell_size = (dx, dy).  fb200_forward then evaluates
 * gravity fields from this form (about 1.4x fewer FP64 instructions) unless
 * dens_override, FB200_REFERENCE_ORDER or FB200_PER_CELL is given; a point
 * that would need the thickness-dependent 1e-5 shift of gxz/gyz at a merged
 * node (it lies in the reference plane on a grid line) makes the call fall
 * back to the per-cell kernels by itself.                                    */
int fb200_model_attach_surface(fb200_model *model, int64_t nfaces,
                               const double *x1, const double *x2,
                               const double *y1, const double *y2,
                               const double *z_face, const double *z_other,
                               const double *w_face, int64_t nnodes,
                               const double *xn, const double *yn, const double *zn,
                               const double *w_node, const double *cell_size /* 2 */);
/* Walls the merged forms above must not be trusted on.  The reference computes a
 * wall shared by two cells once from each side (x2 = x1 + dx of one cell,
 * x1 = b0 + dx*(i+1) of the next, mesh.py:629-634; xc +- dx/2 of neighbouring
 * relief nodes, :447-450); where the two roundings differ, a point lying IN the
 * wall gets atan2 terms with a zero denominator from one cell and a one-ulp
 * denominator of either sign from the other -- they need not cancel, and the
 * reference's answer there depends on that rounding.  Pass the coordinates of
 * such walls (both variants); a forward call with a point within 8 ulp of one
 * of them is evaluated by the per-cell kernels, which see the reference's own
 * bounds.  Meshes and reliefs with exactly representable walls need none.    */
int fb200_model_set_hazard_planes(fb200_model *model, int64_t nx, const double *xw,
                                  int64_t ny, const double *yw, int64_t nz, const double *zw);
int     fb200_model_destroy(fb200_model *model);
int64_t fb200_model_size(const fb200_model *model);
int     fb200_model_device(const fb200_model *model);

/* ---- forward fields: the loop of prism.py:131-141 over a whole model ----
 * For every point l and every field f in field_mask (ascending id order):
 *   out[c * ld_out + l] = scale[c] * sum over prisms of field f
 * c = 0,1,... numbering the requested fields.  One pass evaluates all
 * requested components and shares the per-corner r / log / atan2 work.
 *   dens_override : NULL, or pointer to one double used instead of every
 *                   prism's density (prism.py:134-137, `dens=`).
 *   pmag_override : NULL, or 3 doubles used instead of every prism's
 *                   magnetization (prism.py:641-645, `pmag=`).
 *   fvec          : 3 direction cosines of the inducing field; required when
 *                   FB200_TF is requested (prism.py:640), ignored otherwise.
 *   scale         : one factor per requested field (e.g. G*SI2MGAL,
 *                   prism.py:286); NULL means 1.0 for all.
 * Host pointers by default; with FB200_DEVICE_POINTERS all of xp,yp,zp,out
 * are device pointers and nothing is copied.                                 */
int fb200_forward(fb200_model *model,
                  const double *xp, const double *yp, const double *zp,
                  int64_t n, uint32_t field_mask,
                  const double *dens_override, const double *pmag_override,
                  const double *fvec, const double *scale,
                  double *out, int64_t ld_out, uint32_t flags);

/* ---- sensitivity (Jacobian) matrix --------------------------------------
 * out[l * ld + q] = scale * field(point l, prism q alone, unit property):
 * column q is what prism.<field>(x, y, z, [cell_q], dens=1) returns
 * (eqlayer.py:108-111 layout, harvester.py:720-724 / imaging.py:116-117
 * call pattern).  unit_prop: 1 double (gravity) or 3 (magnetic; for
 * FB200_TF the unit magnetization vector, harvester.py:958-962).  fvec as in
 * fb200_forward.  With FB200_TRANSPOSED, out[q * ld + l].
 * With host `out` the matrix is streamed back in row blocks; with
 * FB200_DEVICE_POINTERS it stays resident in HBM (row-sharded by the caller
 * across GPUs by passing each GPU its slice of points).                      */
int fb200_jacobian(fb200_model *model,
                   const double *xp, const double *yp, const double *zp,
                   int64_t n, int field, const double *unit_prop,
                   const double *fvec, double scale,
                   double *out, int64_t ld, uint32_t flags);

/* ---- adjoint of the sensitivity: out[q] = scale * sum_l data[l] * J[l][q] --
 * J^T d without storing J: what imaging.py:116-118 computes layer by layer as
 * dot(array([prism.gz(x, y, z, [p], dens=1) for p in layer]), gz), and the
 * gradient product of misfit.py:280-294.  Arguments as fb200_jacobian; data
 * has n entries, out has m.  Partial sums over fixed contiguous ranges of the
 * points are added in range order (reproducible run to run).                */
int fb200_adjoint(fb200_model *model,
                  const double *xp, const double *yp, const double *zp,
                  int64_t n, int field, const double *unit_prop,
                  const double *fvec, double scale,
                  const double *data, double *out, uint32_t flags);

/* ---- the reference's own FFI shape: one prism accumulated into res ------
 * `_prism.<field>(xp, yp, zp, x1, x2, y1, y2, z1, z2, <props>, res)` adds ONE
 * prism's effect into caller-owned res[0..n) and never zeroes it
 * (_prism.pyx:72-479).  Host pointers.  Evaluated in reference order with
 * res[l] as the starting value of the running sum.                           */
/* device used by the fb200_prism_* entry points below (default 0) */
int fb200_set_default_device(int device);
int fb200_prism_potential(const double *xp, const double *yp, const double *zp, int64_t n,
                          double x1, double x2, double y1, double y2, double z1, double z2,
                          double density, double *res);          /* _prism.pyx:454 */
int fb200_prism_gx(const double *xp, const double *yp, const double *zp, int64_t n,
                   double x1, double x2, double y1, double y2, double z1, double z2,
                   double density, double *res);                 /* _prism.pyx:196 */
int fb200_prism_gy(const double *xp, const double *yp, const double *zp, int64_t n,
                   double x1, double x2, double y1, double y2, double z1, double z2,
                   double density, double *res);                 /* _prism.pyx:223 */
int fb200_prism_gz(const double *xp, const double *yp, const double *zp, int64_t n,
                   double x1, double x2, double y1, double y2, double z1, double z2,
                   double density, double *res);                 /* _prism.pyx:250 */
int fb200_prism_gxx(const double *xp, const double *yp, const double *zp, int64_t n,
                    double x1, double x2, double y1, double y2, double z1, double z2,
                    double density, double *res);                /* _prism.pyx:277 */
int fb200_prism_gxy(const double *xp, const double *yp, const double *zp, int64_t n,
                    double x1, double x2, double y1, double y2, double z1, double z2,
                    double density, double *res);                /* _prism.pyx:304 */
int fb200_prism_gxz(const double *xp, const double *yp, const double *zp, int64_t n,
                    double x1, double x2, double y1, double y2, double z1, double z2,
                    double density, double *res);                /* _prism.pyx:336 */
int fb200_prism_gyy(const double *xp, const double *yp, const double *zp, int64_t n,
                    double x1, double x2, double y1, double y2, double z1, double z2,
                    double density, double *res);                /* _prism.pyx:368 */
int fb200_prism_gyz(const double *xp, const double *yp, const double *zp, int64_t n,
                    double x1, double x2, double y1, double y2, double z1, double z2,
                    double density, double *res);                /* _prism.pyx:395 */
int fb200_prism_gzz(const double *xp, const double *yp, const double *zp, int64_t n,
                    double x1, double x2, double y1, double y2, double z1, double z2,
                    double density, double *res);                /* _prism.pyx:427 */
int fb200_prism_tf(const double *xp, const double *yp, const double *zp, int64_t n,
                   double x1, double x2, double y1, double y2, double z1, double z2,
                   double mx, double my, double mz, double fx, double fy, double fz,
                   double *res);                                 /* _prism.pyx:72  */
int fb200_prism_bx(const double *xp, const double *yp, const double *zp, int64_t n,
                   double x1, double x2, double y1, double y2, double z1, double z2,
                   double mx, double my, double mz, double *res); /* _prism.pyx:106 */
int fb200_prism_by(const double *xp, const double *yp, const double *zp, int64_t n,
                   double x1, double x2, double y1, double y2, double z1, double z2,
                   double mx, double my, double mz, double *res); /* _prism.pyx:136 */
int fb200_prism_bz(const double *xp, const double *yp, const double *zp, int64_t n,
                   double x1, double x2, double y1, double y2, double z1, double z2,
                   double mx, double my, double mz, double *res); /* _prism.pyx:166 */

/* ---- planting inversion: device-resident effect engine -------------------
 * What fatiando/gravmag/harvester.py keeps in Python objects and numpy loops:
 * the "effect" of every candidate cell on every data set
 * (Data.effect, harvester.py:690-694 and :958-962, called from _calc_effect
 * :487-493) and, per accretion, the data-misfit and shape-of-anomaly terms of
 * predicted + effect for every candidate (_misfitfunc :447-456, _shapefunc
 * :434-444, looped by _grow :412-431).  Data sets are concatenated: sizes[d]
 * points each, `fields[d]` the fb200_field they hold, fvec = 3 direction
 * cosines per data set (only read for FB200_TF), scale[d] the unit factor.
 * `weights` has one entry per datum.  The engine owns `predicted` (float32,
 * like harvester.py:401) and an effect store addressed by caller-chosen ROW
 * numbers (a free-list on the host side keeps them dense).                   */
typedef struct fb200_harvest fb200_harvest;
int fb200_harvest_create(int device, int ndata, const int64_t *sizes, const int *fields,
                         const double *x, const double *y, const double *z,
                         const double *observed, const double *weights,
                         const double *fvec /* 3*ndata or NULL */, const double *scale,
                         fb200_harvest **out);
int fb200_harvest_destroy(fb200_harvest *h);
/* effect rows of ncells cells: bounds = 6 doubles per cell (x1,x2,y1,y2,z1,z2);
 * density[c] (NaN = the cell's props have no 'density' -> zero effect on
 * gravity data, harvester.py:691-692); magnetization = 3 doubles per cell with
 * mag_kind[c] 0 = none, 1 = scalar in [3c] (taken along the data set's
 * inducing field, prism.py:641-645), 2 = vector.  flags: 0 or
 * FB200_REFERENCE_ORDER.                                                     */
int fb200_harvest_effects(fb200_harvest *h, int ncells, const double *bounds,
                          const double *density, const double *magnetization,
                          const int *mag_kind, const int64_t *rows, uint32_t flags);
/* predicted += effect[row], rounded to float32 (harvester.py:373-374)        */
int fb200_harvest_accrete(fb200_harvest *h, int64_t row);
/* for each candidate row (or -1 = the current estimate alone) and data set:
 * misfit[c*ndata+d] = sqrt(sum w r r)/||obs||, shape[c*ndata+d] =
 * ||alpha obs - pred||, pred = predicted + effect[row]                       */
int fb200_harvest_evaluate(fb200_harvest *h, int64_t ncand, const int64_t *rows,
                           double *misfit, double *shape);
int fb200_harvest_predicted(fb200_harvest *h, float *out /* all data sets */);
int fb200_harvest_effect(fb200_harvest *h, int64_t row, double *out /* all data sets */);
int64_t fb200_harvest_launches(const fb200_harvest *h);

/* ---- Jacobian of a regular mesh on a commensurate level grid (opt-in) ------
 * The observation points are the regular grid of gridder/point_generation.py:89-96 at ONE
 * level, x slowest (row l = i*nj + j), with a spacing that divides the cell size of the
 * model's regular mesh (a cell = kx x ky grid steps, mesher/mesh.py:626-634), and coordinates
 * such that node - point is the same double for every (node a, point i) with the same
 * u = a*kx - i + (ni - 1): xu[u] (nx*kx + ni values), likewise yv[v] (ny*ky + nj), and
 * zc[c] = zn[c] - zp (nz + 1).  The caller verifies that property (fatiando_b200/prism.py)
 * and passes the three difference tables; the per-corner kernel is then evaluated once per
 * distinct (u, v, c) and the matrix is written by pure data movement (TMA), every entry
 * bit-identical to fb200_jacobian's node-sharing result.  C-order (ni*nj, ncells) only;
 * out on the host, or on the device with FB200_DEVICE_POINTERS.  FB200_EUNSUP when the mesh
 * has an odd number of cells along x or ld is odd (callers fall back to fb200_jacobian).  */
int fb200_jacobian_grid(fb200_model *model, int64_t ni, int64_t nj, int kx, int ky, const double *xu,
                        const double *yv, const double *zc, int field, const double *unit_prop,
                        const double *fvec, double scale, double *out, int64_t ld, uint32_t flags);

/* ---- consumers of a device-resident sensitivity matrix --------------------
 * The Gauss-Newton products of fatiando/inversion/misfit.py on a ROW BLOCK of J that stays in
 * HBM (fb200_jacobian with FB200_DEVICE_POINTERS; rows = this GPU's observations).  All
 * pointers are device pointers on `device`; J is (n x m) C-order with leading dimension ld.
 *
 * fb200_gram:     G = factor * J^T diag(w) J, (m x m) C-order, full symmetric matrix.
 *                 Replaces `safe_dot(jacobian.T, self.weights*jacobian)` and the factor
 *                 2*regul_param of Misfit.hessian (misfit.py:250-256).  w may be NULL.
 *                 *ms (optional) receives the device time of the kernels.
 * fb200_matvec:   y = J p (the predicted data of a linear problem, misfit.py `predicted`).
 * fb200_matvec_t: g = factor * J^T (w .* r); Misfit.gradient (misfit.py:280-294) with
 *                 factor = -2*regul_param.  w may be NULL.
 * Row blocks of several GPUs are combined by one all-reduce of G / g
 * (fatiando_b200/partition.py: gram_row_sharded).                                  */
int fb200_gram(int device, const double *J, int64_t n, int64_t m, int64_t ld, const double *w,
               double factor, double *G, int64_t ldg, double *ms);
int fb200_matvec(int device, const double *J, int64_t n, int64_t m, int64_t ld, const double *p,
                 double *y);
int fb200_matvec_t(int device, const double *J, int64_t n, int64_t m, int64_t ld, const double *r,
                   const double *w, double factor, double *g);
/* Plain device buffers and synchronous copies for bindings without a tensor library
 * (kind: 0 host->device, 1 device->host, 2 device->device).                        */
int fb200_device_alloc(int device, size_t bytes, void **ptr);
int fb200_device_free(int device, void *ptr);
int fb200_copy(int device, void *dst, const void *src, size_t bytes, int kind);

/* ---- measurement helpers (used by bench.py; not part of the data path) --
 * Device time in milliseconds of the kernels launched by the most recent
 * fb200_forward / fb200_jacobian call on this model (CUDA events on the
 * model's stream), and how many kernels that call launched.                  */
int fb200_last_kernel_ms(const fb200_model *model, double *ms, int *launches);
/* How many observation points of the most recent fb200_forward / fb200_jacobian call were
 * re-evaluated by the per-cell kernels because they lie in a wall that the reference rounds
 * differently from its two sides (fb200_model_set_hazard_planes; mesher/mesh.py:447-456,
 * 629-634) or meet the thickness-dependent shift of a merged relief node.          */
int fb200_last_fallback_points(const fb200_model *model, int *points);

/* Page-locked host memory (cudaHostAlloc) for result arrays that are written repeatedly, e.g.
 * the host-resident sensitivity matrix of an inversion loop (eqlayer.py:100-111 rebuilds it per
 * call): fb200_jacobian streams into it at PCIe DMA speed.  Not needed for correctness.       */
int fb200_host_alloc(size_t bytes, void **ptr);
int fb200_host_free(void *ptr);
/* Register-resident DFMA-chain microbenchmark: the FP64 roofline denominator
 * that MEASURED_PEAKS.json does not carry.  Runs for about `seconds`, returns
 * the best (burst) and the whole-run (sustained) rate in TFLOP/s (DFMA = 2). */
int fb200_fp64_peak(int device, double seconds, double *burst_tflops,
                    double *sustained_tflops);
/* Device-wide L2 flush: overwrite a buffer larger than L2 (bench hygiene).   */
int fb200_flush_l2(int device);
/* Evaluate one device math primitive over host arrays, so tests can pin the
 * kernel's building blocks on the GPU: op 0 log(a/b), 1 atan(a/b) for b >= 0,
 * 2 a/b, 3 branch-free sqrt(a), 4 library sqrt(a).                            */
int fb200_debug_math(int device, int op, const double *a, const double *b,
                     int64_t n, double *out);

#ifdef __cplusplus
}
#endif
#endif /* FATIANDO_B200_H */
