// K6: non-local-means denoise of the (H, W) channel-sum image -- the step between the channel sum and
// the stencil in every 2-D caller: skimage.restoration.denoise_nl_means(image, h = 0.02),
// syn/..._measurement.py:108 (bio/..._analysis.py:350, 592, 668, 725, 989), fast mode, patch_size 7,
// patch_distance 11, sigma 0.
//
// scikit-image builds one integral image of squared differences per patch shift and accumulates each
// pair of pixels symmetrically (the test-side restatement follows it loop for loop).
// Mathematically that is, for every pixel p:
//     out[p] = sum_t w(p, t) v[p + t] / sum_t w(p, t),      t in [-d, d]^2,
//     w(p, t) = exp(-dist) if dist <= 5 else 0,  twice that for t = 0,
//     dist    = max(sum_{u in W(p)} (v[u] - v[u + t])^2, 0) / (h^2 s^2),
//     W(p)    = rows / cols p - offset + 1 .. p + offset   (a 2*offset square: the implementation's quirk),
// on the reflect-padded image.  This kernel evaluates that directly: a CTA owns a 32 x 32 output tile,
// holds the (32 + 2d + 2 offset - 1)^2 neighbourhood in shared memory and walks the (2d + 1)^2 shifts.
// Per shift, phase 1 forms the horizontal 6-sums of squared differences (lanes along rows, eight
// columns per thread so neighbouring sums share their terms: 26 LDS for 8 sums), phase 2 adds six of
// them vertically for four pixels per thread, applies exp and accumulates weight and weighted value.
// The sum buffers are double-buffered: one __syncthreads per shift.  Compute-bound (~150 instructions
// per thread per shift, 529 shifts): this is 500x the arithmetic of the stencil.
// Everything up to the distance is float64 (the B200's FP64 pipe runs at half the FP32 rate): the hard
// cutoff at dist = 5 makes the estimator discontinuous, and float32 distances would land on the other
// side of it for ~1e-6 of the (pixel, shift) pairs -- thousands per image, each worth up to ~1e-4.
#include "hipr_common.cuh"

namespace hipr {

constexpr int NL_T = 32;              // output tile side
constexpr int NL_OFF = 3;             // patch_size 7
constexpr int NL_N = 2 * NL_OFF;      // window side (6)
constexpr int NL_HR = NL_T + NL_N - 1;  // rows of horizontal sums per tile (37)
constexpr int NL_HS = NL_T + 1;       // row stride of the sum buffers (odd: lanes along rows hit 32 banks)
constexpr int NL_MAX_D = 15;

__device__ __forceinline__ int reflect_index(int i, int n) {
    // np.pad(mode='reflect') for a pad smaller than n: -k -> k, n - 1 + k -> n - 1 - k
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return min(max(i, 0), n - 1);   // beyond one reflection: only for tile cells no written pixel uses
}

template <typename T>
__global__ void __launch_bounds__(256)
nlm2d_kernel(const T *__restrict__ img, int H, int W, int d, double inv_h2s2, T *__restrict__ out) {
    extern __shared__ __align__(16) double nl_smem[];
    const int TS = NL_T + 2 * d + NL_N - 1;      // tile side (59 for d = 11)
    const int TP = TS | 1;                       // odd row stride
    double *tile = nl_smem;                      // [TS][TP]
    double *hb = nl_smem + TS * TP;              // [2][NL_HR][NL_HS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.y * NL_T, c0 = blockIdx.x * NL_T;
    // tile origin: first window row (r0 - offset + 1) minus d
    const int tr0 = r0 - NL_OFF + 1 - d, tc0 = c0 - NL_OFF + 1 - d;
    for (int i = tid; i < TS * TS; i += 256) {
        const int ly = i / TS, lx = i - ly * TS;
        const int gy = reflect_index(tr0 + ly, H), gx = reflect_index(tc0 + lx, W);
        tile[ly * TP + lx] = (double)img[(int64_t)gy * W + gx];
    }
    __syncthreads();
    // phase-1 work item: rows lane + 32 * (warp >> 2), eight columns starting at 8 * (warp & 3)
    const int h_row = lane + 32 * (warp >> 2);
    const int h_col = 8 * (warp & 3);
    const bool h_on = h_row < NL_HR;
    // phase-2 pixels: column lane, rows 4 * warp .. + 3
    const int p_row = 4 * warp;
    double acc_w[4] = {0.0, 0.0, 0.0, 0.0}, acc_v[4] = {0.0, 0.0, 0.0, 0.0};
    int buf = 0;
    auto phase1 = [&](int tr, int tc, double *hbuf) {
        if (!h_on) return;
        const double *a = tile + (h_row + d) * TP + h_col + d;
        const double *b = a + tr * TP + tc;
        double D[8 + NL_N - 1];
#pragma unroll
        for (int k = 0; k < 8 + NL_N - 1; ++k) {
            const double df = a[k] - b[k];
            D[k] = df * df;
        }
        double s = D[0];
#pragma unroll
        for (int k = 1; k < NL_N; ++k) s += D[k];
        hbuf[h_row * NL_HS + h_col] = s;
#pragma unroll
        for (int m = 1; m < 8; ++m) {
            s += D[m + NL_N - 1] - D[m - 1];          // sliding window (float64: ~1e-16 of the largest term)
            hbuf[h_row * NL_HS + h_col + m] = s;
        }
    };
    const int nshift = (2 * d + 1) * (2 * d + 1);
    phase1(-d, -d, hb);
    __syncthreads();
    for (int sidx = 0; sidx < nshift; ++sidx) {
        const int tr = sidx / (2 * d + 1) - d, tc = sidx % (2 * d + 1) - d;
        // phase 1 of the next shift fills the other buffer while this shift's sums are consumed
        if (sidx + 1 < nshift) {
            const int ntr = (sidx + 1) / (2 * d + 1) - d, ntc = (sidx + 1) % (2 * d + 1) - d;
            phase1(ntr, ntc, hb + (buf ^ 1) * NL_HR * NL_HS);
        }
        const double *hcur = hb + buf * NL_HR * NL_HS + p_row * NL_HS + lane;
        double hs[4 + NL_N - 1];
#pragma unroll
        for (int k = 0; k < 4 + NL_N - 1; ++k) hs[k] = hcur[k * NL_HS];
        const double self = (tr == 0 && tc == 0) ? 2.0 : 1.0;
        // the shifted pixel p + t: tile coordinates (p - tile origin) = (row + offset - 1 + d, ...)
        const double *vs = tile + (p_row + NL_OFF - 1 + d + tr) * TP + lane + NL_OFF - 1 + d + tc;
        double box = hs[0];
#pragma unroll
        for (int k = 1; k < NL_N; ++k) box += hs[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q > 0) box += hs[q + NL_N - 1] - hs[q - 1];
            const double dist = fmax(box, 0.0) * inv_h2s2;
            if (dist <= 5.0) {
                // float64 exp: the denoised image feeds the line normalisation (centre - min) / (max - min), whose
                // ranges on a denoised image are ~1e-3 of the values -- a float32 weight (1e-7) would surface as
                // 1e-4 in the score
                const double w = self * exp(-dist);
                acc_w[q] += w;
                acc_v[q] = fma(w, vs[q * TP], acc_v[q]);
            }
        }
        buf ^= 1;
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + p_row + q, c = c0 + lane;
        if (r < H && c < W) out[(int64_t)r * W + c] = (T)(acc_v[q] / acc_w[q]);
    }
}

}  // namespace hipr

using namespace hipr;

extern "C" int hipr_denoise_nl_means_2d(const void *image_dev, int H, int W, int dtype, int patch_size,
                                        int patch_distance, double h, void *out_dev, void *stream) {
    if (!image_dev || !out_dev || H < 1 || W < 1 || !(h > 0.0)) return HIPR_E_ARG;
    if (dtype != HIPR_F32 && dtype != HIPR_F64) return HIPR_E_DTYPE;
    if (patch_size != 7 && patch_size != 6) return HIPR_E_UNSUPPORTED;   // skimage makes an even size odd
    if (patch_distance < 0 || patch_distance > NL_MAX_D) return HIPR_E_UNSUPPORTED;
    const int pad = NL_OFF + patch_distance + 1;
    if (H <= pad || W <= pad) return HIPR_E_PATCH;   // single reflection only (np.pad would wrap again)
    const int d = patch_distance;
    const int TS = NL_T + 2 * d + NL_N - 1, TP = TS | 1;
    const size_t smem = ((size_t)TS * TP + 2 * NL_HR * NL_HS) * sizeof(double);
    const double inv = 1.0 / (h * h * 49.0);
    static bool attr = false;
    if (!attr) {
        HIPR_CUDA(cudaFuncSetAttribute(nlm2d_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        HIPR_CUDA(cudaFuncSetAttribute(nlm2d_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    dim3 grid((unsigned)((W + NL_T - 1) / NL_T), (unsigned)((H + NL_T - 1) / NL_T));
    if (grid.y > 65535) return HIPR_E_RANGE;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HIPR_F32)
        nlm2d_kernel<float><<<grid, 256, smem, st>>>((const float *)image_dev, H, W, d, inv, (float *)out_dev);
    else
        nlm2d_kernel<double><<<grid, 256, smem, st>>>((const double *)image_dev, H, W, d, inv, (double *)out_dev);
    return after_launch();
}
