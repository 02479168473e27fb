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

template <bool CALIB>
__global__ void __launch_bounds__(RG_THREADS)
register_kernel(RegGeom g, const float *__restrict__ calib, float *__restrict__ cube_out,
                double *__restrict__ sum_out, unsigned long long *__restrict__ maxkey) {
    extern __shared__ __align__(16) float rg_smem[];
    float *tp = rg_smem;                           // [RG_PX][C] registered values
    float *tq = rg_smem + (size_t)RG_PX * g.C;     // [RG_PX][C] flat-field divisors (CALIB)
    const int tid = threadIdx.x;
    const int C = g.C;
    const int tiles_per_row = (g.W + RG_PX - 1) / RG_PX;
    const int64_t ntiles = (int64_t)g.H * tiles_per_row;
    double vmax = -__longlong_as_double(0x7ff0000000000000ll);
    double vmin = __longlong_as_double(0x7ff0000000000000ll);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r = (int)(tile / tiles_per_row);
        const int x0 = (int)(tile % tiles_per_row) * RG_PX;
        const int npx = min(RG_PX, g.W - x0);
        const int nel = npx * C;
        const int64_t obase = ((int64_t)r * g.W + x0) * C;
        // ---- pass A: gather.  For one excitation the 64 source pixels are ONE contiguous run of
        // npx * c_e floats (clipped where the shift leaves the frame): coalesced loads, up to eight in flight
        // per thread before the first store, scattered into the excitation's channel slot of each pixel.
        // (A per-element formulation with one uniform loop over all excitations was measured 1.7x slower:
        // the index arithmetic per element, not the loads, is what this pass costs.)
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
                // batch width = what this excitation needs (c_e = 6 -> 2 loads per thread, 32 -> 8): predicated-off
                // slots of a fixed 8-wide batch cost as many instructions as live ones (ncu: the gather is
                // issue-bound, 74 % of issue slots)
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
    if (maxkey != nullptr) {
        vmax = warp_max(vmax);
        vmin = -warp_max(-vmin);
        if ((tid & 31) == 0) {
            atomicMax(maxkey, (unsigned long long)key_of_double(vmax));
            atomicMin(maxkey + 1, (unsigned long long)key_of_double(vmin));
        }
    }
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
