// Shared by the 2-D and 3-D non-local-means kernels (nlm2d.cu, nlm3d.cu).
#pragma once
#include "hipr_common.cuh"

namespace hipr {

__device__ __forceinline__ int reflect_index(int i, int n) {
    // np.pad(mode='reflect') for a pad smaller than n: -k -> k, n - 1 + k -> n - 1 - k
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return min(max(i, 0), n - 1);   // beyond one reflection: only for tile cells no written pixel uses
}

// e^x for x in [-6, 0] to ~4e-11 relative: round(x log2 e) by the 1.5 * 2^52 trick, Cody-Waite reduction
// to |r| <= ln2 / 2, then e^r = 1 + r (1 + r (1/2 + r (1/6 + r q(r)))) with the tail q = 1/4! + r/5! + ...
// + r^8/12! evaluated in float32 (it enters multiplied by r^4 <= 1.5e-2, so its 6e-8 becomes < 4e-11), and
// the exponent added to the high word.  8 FP64 + 8 FP32 operations, branch-free, so the seven pixels of a
// thread interleave; CUDA's exp() cost 65 instructions behind a branch here, and the FP64 pipe is what
// bounds this kernel.  (Weights accurate to 1e-10 keep the denoised image, and the line normalisation
// that amplifies it ~1e4 times, far inside the 1e-5 gate.)
__device__ __forceinline__ double exp_small_neg(double x) {
    const double magic = 6755399441055744.0;
    const double t = fma(x, 1.4426950408889634, magic);
    const int n = __double2loint(t);
    const double nf = t - magic;
    double r = fma(nf, -6.93147180369123816490e-01, x);
    r = fma(nf, -1.90821492927058770002e-10, r);
    const float rf = (float)r;
    float q = 2.08767569878680989792e-09f;            // 1 / 12!
    q = fmaf(q, rf, 2.50521083854417187751e-08f);     // 1 / 11!
    q = fmaf(q, rf, 2.75573192239858906526e-07f);
    q = fmaf(q, rf, 2.75573192239858906526e-06f);
    q = fmaf(q, rf, 2.48015873015873015873e-05f);
    q = fmaf(q, rf, 1.98412698412698412698e-04f);
    q = fmaf(q, rf, 1.38888888888888888889e-03f);
    q = fmaf(q, rf, 8.33333333333333333333e-03f);
    q = fmaf(q, rf, 4.16666666666666666667e-02f);     // 1 / 4!
    double p = fma((double)q, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

}  // namespace hipr
