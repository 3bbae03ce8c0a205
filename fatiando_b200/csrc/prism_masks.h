// prism_masks.h -- the component sets the forward kernels are instantiated for.
//
// Any requested field mask is covered greedily by these sets (largest subset
// first; the single-field sets guarantee termination), so every request runs
// on a kernel that evaluates exactly the primitives it needs.  Bits are
// (1u << fb200_field), see include/fatiando_b200.h.
#pragma once

//            potential gx gy gz gxx gxy gxz gyy gyz gzz
#define FB200_GRAVITY_SETS(X)                                                   \
    X(0x001) /* potential */ X(0x002) /* gx  */ X(0x004) /* gy  */              \
    X(0x008) /* gz  */       X(0x010) /* gxx */ X(0x020) /* gxy */              \
    X(0x040) /* gxz */       X(0x080) /* gyy */ X(0x100) /* gyz */              \
    X(0x200) /* gzz */                                                          \
    X(0x00E) /* gx gy gz */                                                     \
    X(0x3F0) /* the six tensor components */                                    \
    X(0x3F8) /* gz + tensor */                                                  \
    X(0x3FE) /* gx gy gz + tensor */                                            \
    X(0x3FF) /* all ten gravity fields */

//            tf bx by bz
#define FB200_MAGNETIC_SETS(X)                                                  \
    X(0x0400) /* tf */ X(0x0800) /* bx */ X(0x1000) /* by */ X(0x2000) /* bz */ \
    X(0x3800) /* bx by bz */                                                    \
    X(0x3C00) /* tf bx by bz */
