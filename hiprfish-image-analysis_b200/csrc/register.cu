// K0: registration paste + channel stack + flat-field divide + channel sum, in one pass.
//
// Replaces syn/..._measurement.py:86-105 (the same block in bio/..._analysis.py:330-348 and, without
// shifts, eco/..._measurement.py:147-148):
//     image_registered[i][r, c, :] = image_stack[i][r - shift_row_i, c - shift_col_i, :]   (0 outside)
//     image_channel = np.dstack(image_registered) / calibration_image
//     image_registered_sum = np.sum(image_channel, axis=2)
// E excitation stacks (H, W, c_e) float32 -> the registered (and flat-fielded) cube (H, W, C) float32,
// C = sum c_e, which the scripts save and the per-cell reduction reads later, plus its float64 channel
// sums and their max / min keys.  HBM-bound: 4C B/px read (+4C for a flat field), 4C written.
//
// A CTA takes 64 consecutive pixels of one row.  Pass A gathers the shifted excitation segments (and
// the flat-field values) into shared memory with coalesced loads: per excitation the 64 source pixels
// are one contiguous run of the stack.  Pass B sums
// each pixel out of shared memory in float64 (four threads per pixel; with a flat field the same
// reciprocal + exact-product quotient as K1, hipr_common.cuh).  Pass C streams the tile to the cube.
//
// Tried and removed (round 1): an interior-tile kernel that stages each excitation's run with one 1-D bulk
// async copy (16-byte aligned start up to 3 floats early, next tile's copies in flight) and reads the staged
// runs in place.  Correct, but slower on B200 (1.3 ms, then 2.1 ms with per-excitation unrolled passes, against
// 0.88 ms): with the loads free, the summing and the re-interleaving of five runs of different pixel strides
// into the (pixel, channel) order still cost ~75 issued instructions per element (ncu: 72 % issue slots).
#include <cstdlib>
#include <type_traits>
#include "hipr_common.cuh"

namespace hipr {

constexpr int RG_MAX_E = 8;
constexpr int RG_PX = 64;        // pixels per tile
constexpr int RG_THREADS = 256;  // 4 threads per pixel in pass B
constexpr int RG_MAX_C = 192;    // 2 tiles * 64 px * 192 ch * 4 B = 96 KB dynamic shared memory

struct RegGeom {
    const float *stack[RG_MAX_E];
    int chans[RG_MAX_E];
    int off[RG_MAX_E + 1];     // channel offset of excitation e in the cube
    int srow[RG_MAX_E];
    int scol[RG_MAX_E];
    int E, H, W, C;
};

// Copies the contiguous run src[0, n) into the tile: element i goes to pixel i / ce, channel i % ce of the
// excitation's slot (d0 = destination of this thread's first element; the destination advances by
// dpx * C + dk per step with a carry when the channel index k wraps).  WIDTH loads are issued before the
// first store.
template <int WIDTH>
__device__ __forceinline__ void gather_run(const float *__restrict__ src, int n, float *__restrict__ tp, int d0, int C,
                                           int ce, int dpx, int dk, int k, int tid) {
    const int step = dpx * C + dk, carry = C - ce;
    for (int i = tid; i < n; i += WIDTH * RG_THREADS) {
        float v[WIDTH];
        int dst[WIDTH];
#pragma unroll
        for (int u = 0; u < WIDTH; ++u) {
            const bool in = i + u * RG_THREADS < n;
            v[u] = in ? ldg_stream(src + i + u * RG_THREADS) : 0.f;
            dst[u] = in ? d0 : -1;
            d0 += step;
            k += dk;
            if (k >= ce) { k -= ce; d0 += carry; }
        }
#pragma unroll
        for (int u = 0; u < WIDTH; ++u)
            if (dst[u] >= 0) tp[dst[u]] = v[u];
    }
}


// One tile by the general route: any channel counts, tiles clipped by a shift or by the image edge.
// tp / tq: [RG_PX][C] shared tiles (registered values, flat-field divisors).  Ends with a CTA barrier.
template <bool CALIB>
__device__ __noinline__ void register_tile_general(const RegGeom &g, const float *__restrict__ calib,
                                                   float *__restrict__ cube_out, double *__restrict__ sum_out, int r, int x0,
                                                   float *__restrict__ tp, float *__restrict__ tq, double &vmax, double &vmin) {
    const int tid = threadIdx.x;
    const int C = g.C;
    {
        const int npx = min(RG_PX, g.W - x0);
        const int nel = npx * C;
        const int64_t obase = ((int64_t)r * g.W + x0) * C;
        // ---- pass A: gather.  For one excitation the 64 source pixels are ONE contiguous run of
        // npx * c_e floats (clipped where the shift leaves the frame): coalesced loads, up to eight in flight
        // per thread before the first store, scattered into the excitation's channel slot of each pixel.
        for (int e = 0; e < g.E; ++e) {
            const int ce = g.chans[e], off = g.off[e];
            const int sr = r - g.srow[e];
            // tile pixels [pa, pb) have their source column inside the frame
            int pa = max(g.scol[e] - x0, 0), pb = min(g.W + g.scol[e] - x0, npx);
            const bool row_ok = sr >= 0 && sr < g.H && pb > pa;
            if (!row_ok) { pa = 0; pb = 0; }
            // zero fill outside [pa, pb): the reference pastes into np.zeros
            for (int i = tid; i < (npx - (pb - pa)) * ce; i += RG_THREADS) {
                int px = i / ce;
                const int k = i - px * ce;
                if (px >= pa) px += pb - pa;
                tp[px * C + off + k] = 0.f;
            }
            if (row_ok) {
                const float *src = g.stack[e] + ((int64_t)sr * g.W + (x0 + pa - g.scol[e])) * ce;
                const int n = (pb - pa) * ce;
                const int dpx = RG_THREADS / ce, dk = RG_THREADS - dpx * ce;
                int px = tid / ce, k = tid - px * ce;
                // batch width = what this excitation needs (c_e = 6 -> 2 loads per thread, 32 -> 8)
                const int need = (n + RG_THREADS - 1) / RG_THREADS;
                if (need <= 2) gather_run<2>(src, n, tp, (pa + px) * C + off + k, C, ce, dpx, dk, k, tid);
                else if (need <= 4) gather_run<4>(src, n, tp, (pa + px) * C + off + k, C, ce, dpx, dk, k, tid);
                else if (need <= 6) gather_run<6>(src, n, tp, (pa + px) * C + off + k, C, ce, dpx, dk, k, tid);
                else gather_run<8>(src, n, tp, (pa + px) * C + off + k, C, ce, dpx, dk, k, tid);
            }
        }
        const bool vec = ((obase | nel) & 3) == 0;   // the tile starts and ends on 16 bytes
        if (CALIB) {
            if (vec && (((uintptr_t)calib) & 15u) == 0) {
                const float4 *src4 = reinterpret_cast<const float4 *>(calib + obase);
                float4 *dst4 = reinterpret_cast<float4 *>(tq);
                for (int i = tid; i < (nel >> 2); i += RG_THREADS) dst4[i] = ldg_stream4(src4 + i);
            } else {
                for (int i = tid; i < nel; i += RG_THREADS) tq[i] = ldg_stream(calib + obase + i);
            }
        }
        __syncthreads();
        // ---- pass B: per-pixel channel sums, four threads per pixel (quarter of the channels each)
        {
            const int px = tid >> 2, part = tid & 3;
            const int cper = (C + 3) >> 2;
            const int c0 = part * cper;
            const int cn = max(0, min(cper, C - c0));
            double s = 0.0;
            if (px < npx) {
                const float *p = tp + (size_t)px * C + c0;
                s = CALIB ? sum_channels_div(p, tq + (size_t)px * C + c0, cn) : sum_channels<false>(p, cn);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (px < npx && part == 0) {
                sum_out[(int64_t)r * g.W + x0 + px] = s;
                vmax = fmax(vmax, s);
                vmin = fmin(vmin, s);
            }
        }
        // ---- pass C: the tile of the registered (flat-fielded) cube, coalesced (128-bit when aligned)
        if (cube_out != nullptr) {
            if (vec && (((uintptr_t)cube_out) & 15u) == 0) {
                float4 *dst4 = reinterpret_cast<float4 *>(cube_out + obase);
                const float4 *p4 = reinterpret_cast<const float4 *>(tp);
                const float4 *q4 = reinterpret_cast<const float4 *>(tq);
                for (int i = tid; i < (nel >> 2); i += RG_THREADS) {
                    float4 v = p4[i];
                    if (CALIB) {
                        const float4 w = q4[i];
                        v.x = __fdiv_rn(v.x, w.x);
                        v.y = __fdiv_rn(v.y, w.y);
                        v.z = __fdiv_rn(v.z, w.z);
                        v.w = __fdiv_rn(v.w, w.w);
                    }
                    __stcs(dst4 + i, v);
                }
            } else {
                for (int i = tid; i < nel; i += RG_THREADS) {
                    const float v = tp[i];
                    cube_out[obase + i] = CALIB ? __fdiv_rn(v, tq[i]) : v;
                }
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void rg_publish_range(double vmax, double vmin, unsigned long long *__restrict__ maxkey) {
    if (maxkey != nullptr) {
        vmax = warp_max(vmax);
        vmin = -warp_max(-vmin);
        if ((threadIdx.x & 31) == 0) {
            atomicMax(maxkey, (unsigned long long)key_of_double(vmax));
            atomicMin(maxkey + 1, (unsigned long long)key_of_double(vmin));
        }
    }
}

// any channel layout: every tile by the general route
template <bool CALIB>
__global__ void __launch_bounds__(RG_THREADS)
register_kernel(const __grid_constant__ RegGeom g, const float *__restrict__ calib, float *__restrict__ cube_out,
                double *__restrict__ sum_out, unsigned long long *__restrict__ maxkey) {
    extern __shared__ __align__(16) float rg_smem[];
    float *tp = rg_smem;                           // [RG_PX][C] registered values
    float *tq = rg_smem + (size_t)RG_PX * g.C;     // [RG_PX][C] flat-field divisors (CALIB)
    const int tiles_per_row = (g.W + RG_PX - 1) / RG_PX;
    const int64_t ntiles = (int64_t)g.H * tiles_per_row;
    double vmax = -__longlong_as_double(0x7ff0000000000000ll);
    double vmin = __longlong_as_double(0x7ff0000000000000ll);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        register_tile_general<CALIB>(g, calib, cube_out, sum_out, (int)(tile / tiles_per_row),
                                     (int)(tile % tiles_per_row) * RG_PX, tp, tq, vmax, vmin);
    rg_publish_range(vmax, vmin, maxkey);
}


// ---- the reference's own channel layout, compiled in ------------------------------------------------------------
// The scripts stack the five excitations' 32 + 23 + 20 + 14 + 6 channels (syn/..._measurement.py:86-103).  With the
// channel counts known at compile time an interior tile is straight-line code: the 64-pixel run of excitation e is
// RG_PX * c_e contiguous floats, thread t takes run elements t, t + 256, ... (25 per thread), and the destination of
// run element i -- pixel i / c_e, channel i % c_e, i.e. word i + (i / c_e) * (C - c_e) of the excitation's slot -- is
// a per-thread constant.  The 25 loads of tile k + 1 are issued into registers before tile k is summed and streamed
// out, and stored to the shared tile afterwards (software pipeline; 4-byte asynchronous copies, LDGSTS, into a
// second tile were measured 2x slower: 1.11 ms).  The few tiles clipped by a shift or by the image edge take
// the general route inside the same launch; tiles are dealt round-robin with a stride coprime to the tiles per row
// (contiguous ranges per CTA were measured 2x slower: hundreds of separate DRAM streams).
// Measured history (2048^2 x 95, B200): general kernel 0.88 ms, 894 M warp instructions (div / mod by run-time
// channel counts, clipping and zero fill per excitation: ~1,400 instructions per thread and tile) -> this layout
// with register staging 0.56 ms + 0.21 ms for a second launch over the clipped tiles -> this kernel.
template <int... CE>
struct ChanLayout {
    static constexpr int E = sizeof...(CE);
    static constexpr int ce(int e) {
        constexpr int v[E] = {CE...};
        return v[e];
    }
    static constexpr int off(int e) {
        int o = 0;
        for (int i = 0; i < e; ++i) o += ce(i);
        return o;
    }
    static constexpr int C = off(E);
    static constexpr int slots(int e) { return (RG_PX * ce(e) + RG_THREADS - 1) / RG_THREADS; }
    static constexpr int slot0(int e) {
        int o = 0;
        for (int i = 0; i < e; ++i) o += slots(i);
        return o;
    }
};
using RefLayout = ChanLayout<32, 23, 20, 14, 6>;

template <int A, int B, typename F>
__device__ __forceinline__ void rg_static_for(F &&f) {
    if constexpr (A < B) {
        f(std::integral_constant<int, A>{});
        rg_static_for<A + 1, B>(f);
    }
}


// Correctly rounded float32 quotient (what numpy's float32 / float32 gives) without __fdiv_rn's ~20 instructions:
// reciprocal estimate + one Newton step, quotient, exact remainder, one correction (Markstein).  Exact when both
// magnitudes lie in [2^-60, 2^60] (no intermediate can overflow, underflow or lose bits); anything else -- zeros of
// the paste's fill, denormals, infinities -- takes __fdiv_rn.
__device__ __forceinline__ float div_rn_fast(float v, float w) {
    const unsigned ev = (__float_as_uint(v) >> 23) & 0xffu, ew = (__float_as_uint(w) >> 23) & 0xffu;
    if (ev - 67u < 121u && ew - 67u < 121u) {
        float y = rcp_approx(w);
        y = fmaf(fmaf(-w, y, 1.0f), y, y);
        const float q = v * y;
        return fmaf(fmaf(-q, w, v), y, q);
    }
    return __fdiv_rn(v, w);
}

__device__ __forceinline__ bool rg_tile_interior(const RegGeom &g, int r, int x0) {
    bool interior = x0 + RG_PX <= g.W;
    for (int e = 0; e < g.E; ++e) {
        const int sr = r - g.srow[e], c0 = x0 - g.scol[e];
        interior = interior && sr >= 0 && sr < g.H && c0 >= 0 && c0 + RG_PX <= g.W;
    }
    return interior;
}

template <bool CALIB, typename L>
__global__ void __launch_bounds__(RG_THREADS, CALIB ? 2 : 3)
register_fixed_kernel(const __grid_constant__ RegGeom g, const float *__restrict__ calib, float *__restrict__ cube_out,
                      double *__restrict__ sum_out, unsigned long long *__restrict__ maxkey) {
    extern __shared__ __align__(16) float rg_smem[];
    constexpr int C = L::C, NEL = RG_PX * C, NV4 = NEL / 4;
    const int tid = threadIdx.x;
    const int tiles_per_row = (g.W + RG_PX - 1) / RG_PX;
    const int64_t ntiles = (int64_t)g.H * tiles_per_row;
    const int64_t t0 = blockIdx.x, t1 = ntiles, tstep = gridDim.x;   // strided: the CTAs sweep the image together (DRAM locality)
    // the flat field travels as 16-byte copies: its tiles start on 16 bytes when the base does (NEL * 4 = 24,320 B)
    const bool calib_vec = !CALIB || ((((uintptr_t)calib) & 15u) == 0 && ((g.W * C) & 3) == 0);
    double vmax = -__longlong_as_double(0x7ff0000000000000ll);
    double vmin = __longlong_as_double(0x7ff0000000000000ll);

    // fast route: every full-width tile; tiles cut by a shift or by the image edge predicate each element on its
    // source pixel being inside the frame (the reference pastes into np.zeros).  General route: a ragged last tile
    // of a row (W % 64 != 0), an unaligned flat field.
    auto fast = [&](int x0) { return calib_vec && x0 + RG_PX <= g.W; };   // CTA-uniform
    constexpr int NSLOT = L::slot0(L::E);
    float v[NSLOT];
    // CALIB: the flat-field tile (one contiguous 24,320-byte piece) arrives by a bulk asynchronous copy (TMA engine)
    // into one of two buffers, completion on an mbarrier: no registers, one issuing thread
    float *tp = rg_smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(rg_smem + 3 * NEL);
    uint32_t phase_bits = 0u;                 // bit b: parity of the next completion of buffer b
    uint64_t pol = 0;
    if (CALIB) {
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            fence_barrier_init();
            pol = policy_evict_first();
        }
        __syncthreads();
    }
    // every load of a tile, issued back to back into registers (they land while the previous tile is summed and stored)
    auto load_tile = [&](int64_t tile, int it) {
        const int r = (int)(tile / tiles_per_row), x0 = (int)(tile % tiles_per_row) * RG_PX;
        if (!fast(x0)) return;
        if (CALIB && tid == 0) {
            fence_proxy_async();     // the buffer's earlier generic-proxy accesses are ordered before the async write
            mbar_expect_tx(&bars[it & 1], NEL * 4);
            bulk_g2s(rg_smem + (1 + (it & 1)) * NEL, calib + ((int64_t)r * g.W + x0) * C, NEL * 4, &bars[it & 1], pol);
        }
        auto body = [&](auto clip_c) {
            constexpr bool CLIP = decltype(clip_c)::value;
            rg_static_for<0, L::E>([&](auto ec) {
                constexpr int e = decltype(ec)::value, ce = L::ce(e), n = RG_PX * ce;
                const int sr = r - g.srow[e], c0 = x0 - g.scol[e];
                const bool row_ok = !CLIP || (sr >= 0 && sr < g.H);
                const float *src = g.stack[e] + ((int64_t)sr * g.W + c0) * ce + tid;
                rg_static_for<0, L::slots(e)>([&](auto uc) {
                    constexpr int u = decltype(uc)::value;
                    const int i = u * RG_THREADS + tid;
                    bool ok = ((u + 1) * RG_THREADS <= n) || i < n;
                    if (CLIP) ok = ok && row_ok && (unsigned)(c0 + i / ce) < (unsigned)g.W;
                    v[L::slot0(e) + u] = ok ? ldg_stream(src + u * RG_THREADS) : 0.f;
                });
            });
        };
        if (rg_tile_interior(g, r, x0)) body(std::false_type{});
        else body(std::true_type{});
    };
    auto store_tile = [&](float *tp) {
        rg_static_for<0, L::E>([&](auto ec) {
            constexpr int e = decltype(ec)::value, ce = L::ce(e), n = RG_PX * ce;
            rg_static_for<0, L::slots(e)>([&](auto uc) {
                constexpr int u = decltype(uc)::value;
                const int i = u * RG_THREADS + tid;
                if ((u + 1) * RG_THREADS <= n || i < n) tp[i + (i / ce) * (C - ce) + L::off(e)] = v[L::slot0(e) + u];
            });
        });
    };

    if (t0 < t1) load_tile(t0, 0);
    int it = 0;
    for (int64_t tile = t0; tile < t1; tile += tstep, ++it) {
        const int r = (int)(tile / tiles_per_row), x0 = (int)(tile % tiles_per_row) * RG_PX;
        const bool is_fast = fast(x0);
        float *tq = rg_smem + (1 + (it & 1)) * NEL;
        if (is_fast) store_tile(tp);
        __syncthreads();
        if (tile + tstep < t1) load_tile(tile + tstep, it + 1);
        if (CALIB && is_fast) {
            mbar_wait(&bars[it & 1], (phase_bits >> (it & 1)) & 1u);
            phase_bits ^= 1u << (it & 1);
        }
        if (!is_fast) {
            register_tile_general<CALIB>(g, calib, cube_out, sum_out, r, x0, tp, tq, vmax, vmin);
            continue;
        }
        const int64_t obase = ((int64_t)r * g.W + x0) * C;
        // ---- per-pixel channel sums: four threads per pixel, (C + 3) / 4 channels each (the last part fewer)
        {
            constexpr int CPER = (C + 3) / 4, LASTN = C - 3 * CPER;
            const int px = tid >> 2, part = tid & 3;
            const float *p = tp + px * C + part * CPER;
            double s;
            const float *q = tq + px * C + part * CPER;
            if (part < 3) s = CALIB ? sum_channels_div(p, q, CPER) : sum_channels<false>(p, CPER);
            else s = CALIB ? sum_channels_div(p, q, LASTN) : sum_channels<false>(p, LASTN);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (part == 0) {
                sum_out[(int64_t)r * g.W + x0 + px] = s;
                vmax = fmax(vmax, s);
                vmin = fmin(vmin, s);
            }
        }
        // ---- the tile of the registered (flat-fielded) cube
        if (cube_out != nullptr) {
            if ((((uintptr_t)(cube_out + obase)) & 15u) == 0) {
                float4 *dst4 = reinterpret_cast<float4 *>(cube_out + obase);
                const float4 *p4 = reinterpret_cast<const float4 *>(tp);
                const float4 *q4 = reinterpret_cast<const float4 *>(tq);
#pragma unroll
                for (int k = 0; k < (NV4 + RG_THREADS - 1) / RG_THREADS; ++k) {
                    const int i = k * RG_THREADS + tid;
                    if (i < NV4) {
                        float4 a = p4[i];
                        if (CALIB) {
                            const float4 w = q4[i];
                            a.x = div_rn_fast(a.x, w.x);
                            a.y = div_rn_fast(a.y, w.y);
                            a.z = div_rn_fast(a.z, w.z);
                            a.w = div_rn_fast(a.w, w.w);
                        }
                        __stcs(dst4 + i, a);
                    }
                }
            } else {
                for (int i = tid; i < NEL; i += RG_THREADS) {
                    const float a = tp[i];
                    cube_out[obase + i] = CALIB ? __fdiv_rn(a, tq[i]) : a;
                }
            }
        }
        __syncthreads();      // the tile is rewritten in the next trip
    }
    rg_publish_range(vmax, vmin, maxkey);
}

}  // namespace hipr

using namespace hipr;

extern "C" int hipr_register_stacks(const float *const *stacks_dev, const int32_t *chans, const int32_t *shift_row,
                                    const int32_t *shift_col, int n_stacks, int H, int W, const float *calib_dev,
                                    float *cube_dev, double *sum_dev, uint64_t *maxkey_dev, void *stream) {
    if (!stacks_dev || !chans || !shift_row || !shift_col || !sum_dev || n_stacks <= 0 || H <= 0 || W <= 0)
        return HIPR_E_ARG;
    if (n_stacks > RG_MAX_E) return HIPR_E_RANGE;
    RegGeom g;
    memset(&g, 0, sizeof(g));
    g.E = n_stacks;
    g.H = H;
    g.W = W;
    int C = 0;
    for (int e = 0; e < n_stacks; ++e) {
        if (!stacks_dev[e] || chans[e] <= 0) return HIPR_E_ARG;
        if (((uintptr_t)stacks_dev[e]) & 3u) return HIPR_E_ALIGN;
        g.stack[e] = stacks_dev[e];
        g.chans[e] = chans[e];
        g.off[e] = C;
        g.srow[e] = shift_row[e];
        g.scol[e] = shift_col[e];
        C += chans[e];
    }
    g.off[n_stacks] = C;
    g.C = C;
    if (C > RG_MAX_C) return HIPR_E_RANGE;
    // the reference's paste raises (shape mismatch) when a shift exceeds the frame
    for (int e = 0; e < n_stacks; ++e)
        if (shift_row[e] > H || shift_row[e] < -H || shift_col[e] > W || shift_col[e] < -W) return HIPR_E_RANGE;
    cudaStream_t st = (cudaStream_t)stream;
    if (maxkey_dev) {
        HIPR_CUDA(cudaMemsetAsync(maxkey_dev, 0x00, sizeof(uint64_t), st));
        HIPR_CUDA(cudaMemsetAsync(maxkey_dev + 1, 0xff, sizeof(uint64_t), st));
    }
    const size_t smem = (size_t)(calib_dev ? 2 : 1) * RG_PX * C * sizeof(float);
    const int64_t ntiles = (int64_t)H * ((W + RG_PX - 1) / RG_PX);
    // (register_kernel<true> and <false> have the same function type: one flag per kernel, not per generic lambda)
    static std::atomic<uint64_t> attr_gen[2];
    // the reference's channel layouts -- 95 = 32 + 23 + 20 + 14 + 6 (five lasers) and 63 = 23 + 20 + 14 + 6 (without the
    // 405 nm excitation, syn/..._classify_spectra.py:30-33): the pipelined straight-line kernel (clipped tiles stay on
    // its fast path, a ragged last tile takes the general route inside it)
    auto try_layout = [&](auto layout_tag, int *handled) -> int {
        using L = decltype(layout_tag);
        *handled = 0;
        if (n_stacks != L::E || getenv("HIPR_REGISTER_GENERIC") != nullptr) return HIPR_OK;
        for (int e = 0; e < n_stacks; ++e)
            if (chans[e] != L::ce(e)) return HIPR_OK;
        *handled = 1;
        static std::atomic<uint64_t> attr_fix[2];          // (per layout: the lambda is instantiated once per tag type)
        const size_t smem2 = calib_dev ? (size_t)3 * RG_PX * L::C * 4 + 16 : smem;   // flat field: two divisor tiles in rotation + 2 mbarriers
        auto launch_fixed = [&](auto kern) -> int {
            if (first_use_on_device(attr_fix[calib_dev ? 1 : 0]))
                HIPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * RG_PX * L::C * 4)));
            int per_sm = 1;
            HIPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RG_THREADS, smem2));
            if (per_sm < 1) per_sm = 1;
            int64_t grid = (int64_t)sm_count() * per_sm;
            if (grid > ntiles) grid = ntiles;
            // a stride coprime to the tiles per row spreads the edge-column tiles (predicated loads) over all CTAs
            // instead of the few whose stride class hits columns 0 and W - 64
            auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
            while (grid > 1 && gcd(grid, (W + RG_PX - 1) / RG_PX) != 1) --grid;
            kern<<<(unsigned)grid, RG_THREADS, smem2, st>>>(g, calib_dev, cube_dev, sum_dev,
                                                           reinterpret_cast<unsigned long long *>(maxkey_dev));
            return after_launch();
        };
        return calib_dev ? launch_fixed(register_fixed_kernel<true, L>) : launch_fixed(register_fixed_kernel<false, L>);
    };
    int handled = 0;
    int rc = try_layout(RefLayout{}, &handled);
    if (handled) return rc;
    rc = try_layout(ChanLayout<23, 20, 14, 6>{}, &handled);
    if (handled) return rc;
    auto launch = [&](auto kern) -> int {
        if (first_use_on_device(attr_gen[calib_dev ? 1 : 0]))
            HIPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * RG_PX * RG_MAX_C * 4));
        int per_sm = 1;
        HIPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RG_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        int64_t grid = (int64_t)sm_count() * per_sm;
        if (grid > ntiles) grid = ntiles;
        kern<<<(unsigned)grid, RG_THREADS, smem, st>>>(g, calib_dev, cube_dev, sum_dev,
                                                      reinterpret_cast<unsigned long long *>(maxkey_dev));
        return after_launch();
    };
    return calib_dev ? launch(register_kernel<true>) : launch(register_kernel<false>);
}
