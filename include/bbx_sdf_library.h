/*
 * bbx_sdf_library.h -- SDF colliders that exist both as host functions (so they
 * can be handed to the reference's UseSDF(world, fn), reference ballbox.h:1688)
 * and as device code (selected with UseSDFDevice, ballbox_ext.h).
 *
 * The reference evaluates the user's SDF through a host function pointer
 * (ballbox.h:2083, :2325-2327, :2430, :2461); a GPU kernel cannot.  The scene
 * author therefore supplies the SDF in this header, which compiles for both
 * sides.  BBX_SDF_WAVY_DET uses bbx_det_sin, a polynomial sine built from
 * + - * and floorf only: with FMA contraction off (gcc -ffp-contract=off,
 * nvcc -fmad=false) it returns the same bits on the CPU and on the GPU, which
 * libm sinf / CUDA sinf do not (they differ by <= 1-2 ulp, amplified by the
 * eps = 1e-3 central difference of SDF_EstimateNormal to ~2e-4 in the normal).
 */
#ifndef BBX_SDF_LIBRARY_H
#define BBX_SDF_LIBRARY_H

#include <math.h>

#if defined(__CUDACC__)
#define BBX_HD __host__ __device__ __forceinline__
#else
#define BBX_HD static inline
#endif

#define BBX_SDF_ID_NONE       0
#define BBX_SDF_ID_PLANE_Y    1
#define BBX_SDF_ID_WAVY_DET   2
#define BBX_SDF_ID_WAVY_LIBM  3
#define BBX_SDF_ID_SPHERE_IN  4
#define BBX_SDF_ID_BAKED      5   /* a host function sampled once on a regular grid (device only; see ballbox_ext.h) */

/* sine with one-step range reduction and an odd degree-11 polynomial; every
 * operation is a single IEEE binary32 op in a fixed order. */
BBX_HD float bbx_det_sin(float x)
{
    float k  = floorf(x * 0.15915494f + 0.5f);
    float r  = x - k * 6.2831855f;
    if (r > 1.5707964f)        r =  3.1415927f - r;
    else if (r < -1.5707964f)  r = -3.1415927f - r;
    float r2 = r * r;
    float p  = -2.5052108e-08f;
    p = p * r2;  p = p + 2.7557319e-06f;
    p = p * r2;  p = p - 1.9841270e-04f;
    p = p * r2;  p = p + 8.3333333e-03f;
    p = p * r2;  p = p - 1.6666667e-01f;
    p = p * r2;
    p = p * r;
    return r + p;
}

BBX_HD float bbx_sdf_eval(int id, const float* prm, float x, float y, float z)
{
    switch (id) {
    case BBX_SDF_ID_PLANE_Y:   return y - prm[0];
    case BBX_SDF_ID_WAVY_DET:  { float w = bbx_det_sin(x) * bbx_det_sin(z); return y - w; }
    case BBX_SDF_ID_WAVY_LIBM: { float w = sinf(x) * sinf(z); return y - w; }
    case BBX_SDF_ID_SPHERE_IN: { float l = sqrtf(x * x + y * y + z * z); return prm[0] - l; }
    default:                   return 1.0e30f;
    }
}

#endif /* BBX_SDF_LIBRARY_H */
