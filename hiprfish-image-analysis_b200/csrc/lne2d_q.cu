// K3q: the 2-D stencil + epilogue on a 31-bit FIXED-POINT copy of the sum image.
//
// Why: the score is (centre - min) / (max - min) along each line, and on real data the line
// range is ~1/10 of the value, so any rounding of the sum image is amplified ~10-20x.  A
// float32 sum image (6e-8 relative) lands at ~6e-6 in the score -- outside the 1e-5 parity gate
// on some pixels -- and a float64 stencil costs 3-4x the instructions (no DMNMX on sm_100).
// Instead the tile loader maps the float64 channel sums affinely onto integers,
//     q = (S - min) * (0x7E000000 / (max - min)) + 0x00800000,
// with the global min / max that K1 produced.  F1/F2 scores are invariant under that map; F3 and
// ME2 carry their 1e-8 epsilon into q units (eps_q = 1e-8 * max * K).  The integers are stored as
// the bit patterns of positive normal floats, whose ordering is the integer ordering, so the
// 11-sample min / max run on FMNMX3 exactly as a float kernel would -- but every min, max and
// difference is EXACT; the only roundings left are two int->float conversions and one divide
// per line (~1e-7).  Quantisation step: (max - min) * 4.7e-10.
//
// With the pinned (11, 9) table the sample offsets are compile-time immediates of the LDS
// instructions (baked_tables.cuh); any other 11x9 table takes the parameter-bank variant.
#include <type_traits>
#include "hipr_common.cuh"
#include "lne_math.cuh"
#include "baked_tables.cuh"

namespace hipr {

constexpr int Q_TW = 32, Q_TH = 32, Q_P = 11, Q_R = 9, Q_HALF = 5;
constexpr int Q_SW = Q_TW + Q_P - 1;  // 42
constexpr int Q_SH = Q_TH + Q_P - 1;  // 42
constexpr uint32_t Q_BIAS = 0x00800000u;          // smallest positive normal float
constexpr double Q_SPAN = (double)0x7E000000u;    // BIAS + SPAN = 0x7E800000 < +inf

struct TableQ {
    int off[Q_R * Q_P];  // dy * Q_SW + dx
};

constexpr int q_baked_off(int t, int li) {
    return kBaked2D[(t * Q_P + li) * 2] * Q_SW + kBaked2D[(t * Q_P + li) * 2 + 1];
}
// T and LI are compile-time, so the baked offset is a constant expression (an LDS immediate)
template <bool BAKED, int T, int LI>
__device__ __forceinline__ int q_off(const TableQ &tab) {
    if constexpr (BAKED) {
        constexpr int off = q_baked_off(T, LI);
        return off;
    } else {
        return tab.off[T * Q_P + LI];
    }
}
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// (1 - qcv) without the cancellation: 1 - (uq-lq)/(uq+lq+e) = (2 lq + e)/(uq+lq+e)
template <int FLAVOUR>
__device__ __forceinline__ float q_reduce(float (&r)[Q_R]) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < Q_R; ++i) sum += r[i];
    const float mean = sum * (1.0f / Q_R);
    sort_network<float, Q_R>(r);
    const float lq = r[2], uq = r[6];  // np.percentile(.., 25 / 75) of 9 values: order statistics 2 and 6
    float factor;
    if (FLAVOUR == HIPR_FLAVOUR_F1) {
        factor = (uq > 0.f) ? __fdiv_rn(2.f * lq + 1e-8f, uq + lq + 1e-8f) : 1.f;
    } else if (FLAVOUR == HIPR_FLAVOUR_F2 || FLAVOUR == HIPR_FLAVOUR_ME2) {
        const float s = uq + lq;
        factor = (s == 0.f) ? 1.f : __fdiv_rn(2.f * lq, s);   // nan_to_num(0/0) = 0 -> factor 1
    } else {
        factor = __fdiv_rn(2.f * lq + 1e-8f, uq + lq + 1e-8f);
    }
    return mean * factor;
}

template <typename SrcT, int FLAVOUR, bool BAKED>
__global__ void __launch_bounds__(256)
lne2d_q_kernel(const SrcT *__restrict__ img, int Hs, int Ws, int64_t ld, int src_off, int y_begin, int H, int W,
               const __grid_constant__ TableQ tab, const unsigned long long *__restrict__ range,
               float *__restrict__ out) {
    __shared__ float tile[Q_SH * Q_SW];
    __shared__ double red[16];
    const int x0 = blockIdx.x * Q_TW, y0 = y_begin + blockIdx.y * Q_TH;
    // one pass over the tile: each thread keeps its (at most 7) samples in registers, so the
    // tile-local range costs no second read
    constexpr int PER = (Q_SH * Q_SW + 255) / 256;
    double vals[PER];
    double vmax = -__longlong_as_double(0x7ff0000000000000ll), vmin = -vmax;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + 256 * k;
        vals[k] = 0.0;
        if (i < Q_SH * Q_SW) {
            const int ly = i / Q_SW, lx = i - ly * Q_SW;
            const int sy = min(max(y0 + ly - Q_HALF + src_off, 0), Hs - 1);
            const int sx = min(max(x0 + lx - Q_HALF + src_off, 0), Ws - 1);
            double v = (double)img[(int64_t)sy * ld + sx];
            if (v != v) v = 0.0;  // nan_to_num; NaN is not representable in fixed point
            vals[k] = v;
            vmax = fmax(vmax, v);
            vmin = fmin(vmin, v);
        }
    }
    if (range != nullptr) {
        vmax = double_of_key(range[0]);
        vmin = double_of_key(range[1]);
    } else {
        // LOCAL range: F1/F2 are invariant to any affine map, so the tile's own min/max serve (and
        // give a finer grid); the stencil then does not depend on a global reduction
        vmax = warp_max(vmax);
        vmin = -warp_max(-vmin);
        if ((threadIdx.x & 31) == 0) {
            red[(threadIdx.x >> 5) * 2] = vmax;
            red[(threadIdx.x >> 5) * 2 + 1] = vmin;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            vmax = fmax(vmax, red[2 * j]);
            vmin = fmin(vmin, red[2 * j + 1]);
        }
    }
    const double K = (vmax > vmin) ? Q_SPAN / (vmax - vmin) : 0.0;
    const float eps_q = (K > 0.0) ? (float)(1e-8 * fabs(vmax) * K) : 1.0f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + 256 * k;
        if (i < Q_SH * Q_SW) {
            const double qd = fmin(fmax((vals[k] - vmin) * K, 0.0), Q_SPAN);
            tile[i] = __uint_as_float(__double2uint_rn(qd) + Q_BIAS);
        }
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll 1
    for (int k = 0; k < Q_TH / 8; ++k) {
        const int py = ty + 8 * k;
        const int x = x0 + tx, y = y0 + py;
        if (x >= W || y >= H) continue;
        const float *base = tile + py * Q_SW + tx;
        float r[Q_R];
        static_for<0, Q_R>([&](auto tc) {
            constexpr int t = decltype(tc)::value;
            float mn = base[q_off<BAKED, t, 0>(tab)], mx = mn;
            static_for<1, Q_P>([&](auto lc) {
                constexpr int li = decltype(lc)::value;
                const float s = base[q_off<BAKED, t, li>(tab)];
                mn = fminf(mn, s);
                mx = fmaxf(mx, s);
            });
            const float c = base[q_off<BAKED, t, Q_HALF>(tab)];
            const float dq = __uint2float_rn(__float_as_uint(c) - __float_as_uint(mn));
            const float rq = __uint2float_rn(__float_as_uint(mx) - __float_as_uint(mn));
            if (FLAVOUR == HIPR_FLAVOUR_F1 || FLAVOUR == HIPR_FLAVOUR_F2)
                r[t] = __fdividef(dq, rq);                       // 0/0 -> NaN on a flat line
            else if (FLAVOUR == HIPR_FLAVOUR_F3)
                r[t] = __fdividef(dq, rq + eps_q);
            else
                r[t] = __fdividef(dq, fmaxf(rq, eps_q));
        });
        out[(int64_t)y * W + x] = q_reduce<FLAVOUR>(r);
    }
}

template <typename SrcT, bool BAKED>
static int lne2d_q_launch(const SrcT *img, int Hs, int Ws, int64_t ld, int src_off, int y_begin, int H, int W,
                          const TableQ &tab, int flavour, const unsigned long long *range, float *out,
                          cudaStream_t st) {
    // rows [y_begin, H) of the output are computed
    dim3 grid((W + Q_TW - 1) / Q_TW, (H - y_begin + Q_TH - 1) / Q_TH);
    switch (flavour) {
        case HIPR_FLAVOUR_F1:
            lne2d_q_kernel<SrcT, HIPR_FLAVOUR_F1, BAKED><<<grid, 256, 0, st>>>(img, Hs, Ws, ld, src_off, y_begin, H, W, tab, range, out);
            break;
        case HIPR_FLAVOUR_F2:
            lne2d_q_kernel<SrcT, HIPR_FLAVOUR_F2, BAKED><<<grid, 256, 0, st>>>(img, Hs, Ws, ld, src_off, y_begin, H, W, tab, range, out);
            break;
        case HIPR_FLAVOUR_F3:
            lne2d_q_kernel<SrcT, HIPR_FLAVOUR_F3, BAKED><<<grid, 256, 0, st>>>(img, Hs, Ws, ld, src_off, y_begin, H, W, tab, range, out);
            break;
        default:
            return HIPR_E_FLAVOUR;
    }
    return after_launch();
}

}  // namespace hipr

namespace hipr {
// rows [y_begin, y_end) of the score of an UNPADDED (padded == 0) or padded image; range_dev NULL =
// tile-local quantisation (F1 / F2 only)
int lne2d_q_rows(const void *image_dev, int Hs, int Ws, int64_t ld, int padded, int dtype, const int32_t *table_host,
                 int flavour, const uint64_t *range_dev, float *out_dev, int y_begin, int y_end, cudaStream_t st) {
    const int H = padded ? Hs - (Q_P - 1) : Hs, W = padded ? Ws - (Q_P - 1) : Ws;
    const int src_off = padded ? Q_HALF : 0;
    if (H < 1 || W < 1) return HIPR_E_PATCH;
    if (y_begin < 0 || y_end > H || y_begin >= y_end) return HIPR_E_ARG;
    if (!range_dev && flavour != HIPR_FLAVOUR_F1 && flavour != HIPR_FLAVOUR_F2) return HIPR_E_FLAVOUR;
    TableQ tab;
    bool baked = true;
    for (int i = 0; i < Q_R * Q_P; ++i) {
        const int dy = table_host[2 * i], dx = table_host[2 * i + 1];
        if (dy < 0 || dy >= Q_P || dx < 0 || dx >= Q_P) return HIPR_E_TABLE;
        tab.off[i] = dy * Q_SW + dx;
        baked = baked && dy == kBaked2D[2 * i] && dx == kBaked2D[2 * i + 1];
    }
    const unsigned long long *rg = reinterpret_cast<const unsigned long long *>(range_dev);
    if (dtype == HIPR_F64) {
        if (baked) return lne2d_q_launch<double, true>((const double *)image_dev, Hs, Ws, ld, src_off, y_begin, y_end, W, tab, flavour, rg, out_dev, st);
        return lne2d_q_launch<double, false>((const double *)image_dev, Hs, Ws, ld, src_off, y_begin, y_end, W, tab, flavour, rg, out_dev, st);
    }
    if (baked) return lne2d_q_launch<float, true>((const float *)image_dev, Hs, Ws, ld, src_off, y_begin, y_end, W, tab, flavour, rg, out_dev, st);
    return lne2d_q_launch<float, false>((const float *)image_dev, Hs, Ws, ld, src_off, y_begin, y_end, W, tab, flavour, rg, out_dev, st);
}
}  // namespace hipr

using namespace hipr;

extern "C" int hipr_lne2d_q(const void *image_dev, int Hs, int Ws, int64_t ld, int padded, int dtype, int patch_size,
                            int n_dirs, const int32_t *table_host, int flavour, const uint64_t *range_dev,
                            float *out_dev, void *stream) {
    if (!image_dev || !out_dev || !table_host || Hs < 1 || Ws < 1 || ld < Ws) return HIPR_E_ARG;
    if (dtype != HIPR_F32 && dtype != HIPR_F64) return HIPR_E_DTYPE;
    if (patch_size != Q_P || n_dirs != Q_R) return HIPR_E_TABLE;   // callers use hipr_lne2d otherwise
    const int H = padded ? Hs - (Q_P - 1) : Hs;
    if (H < 1) return HIPR_E_PATCH;
    return lne2d_q_rows(image_dev, Hs, Ws, ld, padded, dtype, table_host, flavour, range_dev, out_dev, 0, H,
                        (cudaStream_t)stream);
}
