// K1: channel-sum prologue.  S[p] = sum_c cube[p, c]  (+ global max of S).
//
// Replaces np.sum(image_channel, axis=2) / np.max, syn/..._measurement.py:105-106.
// HBM-bound: 4*C bytes read per pixel, 4 (or 8) written.  The (npix, C) cube is a flat
// stream; pixel rows are 4*C = 380 bytes (C = 95), so no tensor map with a channel-sized box is
// legal (inner box bytes must be a multiple of 16).  Instead a persistent CTA per SM runs a
// ring of 1-D bulk async copies (cp.async.bulk -> UBLKCP, the TMA engine) of whole pixel
// chunks (128 px * 380 B = 48,640 B, a multiple of 16) into shared memory, completion signalled
// through mbarriers.  Consumer threads then sum one pixel each straight out of shared
// memory: thread t reads words t*C + c, and with C odd the 32 lanes of a warp hit 32 distinct
// banks.  Sums are accumulated in float64 (four independent chains) so the float32 result is
// the correctly rounded sum; the stencil that follows amplifies any error by S/range.
#include <cstdlib>
#include "hipr_common.cuh"

namespace hipr {

constexpr int CS_GROUP_THREADS = 128;  // one consumer thread per pixel of a chunk
constexpr int CS_MAX_GROUPS = 4;       // consumer groups; chunk `it` goes to group it % groups
constexpr int CS_MAX_THREADS = CS_MAX_GROUPS * CS_GROUP_THREADS + 32;  // + one producer warp
constexpr int CS_MAX_STAGES = 8;
constexpr int CS_SMEM_BUDGET = 227 * 1024 - 1024;
// Bytes kept in flight per SM.  Measured on B200 (scratch sweep): 3 x 48,640 B is the
// optimum for C = 95 (6.78 TB/s); 4 stages (195 KB) lose 4 %, 2 stages 9 %.
constexpr int CS_INFLIGHT_TARGET = 150 * 1024;
// INVARIANT: stages % groups == 0, so that a given stage is always consumed by the same group.
// Otherwise a group can reach a stage's full-barrier one phase early (bulk copies may land out
// of order), where a parity wait on a phase that has not started passes immediately.

// CALIB: a second stream, the flat-field divisor (same shape as the cube), lands behind the cube
// chunk in every stage; consumers add p[c] / q[c] (syn/..._measurement.py:104-105; sum_channels_div).
// With twice the bytes and ~3x the arithmetic per pixel, two threads share a pixel (half the channels
// each, 64-pixel chunks): lanes 2k / 2k+1 read words k*C + c and k*C + 48 + c, which for odd C = 95
// still fall on 32 distinct banks.
// InT: float, or raw detector counts (uint16_t / uint8_t) converted on the fly as python-bioformats'
// rescale does, float32(count) / float32(scale) per sample (raw_value); half / a quarter of the bytes.
template <typename OutT, bool CALIB, typename InT>
__global__ void __launch_bounds__(CS_MAX_THREADS, 2)   // 2: caps the kernel at 60 registers (56 used, no spills) so that 4 stencil CTAs of the previous field of view fit beside its one CTA per SM
chansum_bulk_kernel(const InT *__restrict__ cube, const float *__restrict__ calib, int64_t nchunks, int C, int cpx,
                    int stages, int groups, float scale, OutT *__restrict__ out,
                    unsigned long long *__restrict__ maxkey) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t chunk_bytes = (uint32_t)cpx * (uint32_t)C * (uint32_t)sizeof(InT);
    const uint32_t stage_bytes = CALIB ? 2u * chunk_bytes : chunk_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + CS_MAX_STAGES;
    unsigned char *ring = smem_raw + 128;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CS_GROUP_THREADS / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == groups * (CS_GROUP_THREADS / 32)) {
        // ---- producer: one elected lane keeps `stages` bulk copies in flight
        if ((tid & 31) == 0) {
            const uint64_t pol = policy_evict_first();   // the cube is read once
            int s = 0;
            uint32_t round = 0;
            for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                mbar_expect_tx(&full[s], stage_bytes);
                bulk_g2s(ring + (size_t)s * stage_bytes, cube + chunk * (int64_t)cpx * C, chunk_bytes, &full[s], pol);
                if (CALIB)
                    bulk_g2s(ring + (size_t)s * stage_bytes + chunk_bytes, calib + chunk * (int64_t)cpx * C,
                             chunk_bytes, &full[s], pol);
                if (++s == stages) { s = 0; ++round; }
            }
        }
        return;
    }

    // ---- consumers
    constexpr int TPP = CALIB ? 2 : 1;   // threads per pixel
    const int g = warp / (CS_GROUP_THREADS / 32);
    const int t = (tid - g * CS_GROUP_THREADS) / TPP;
    const int part = (tid - g * CS_GROUP_THREADS) % TPP;
    const int cper = (C + TPP - 1) / TPP;
    const int c0 = part * cper;
    const int cn = (C - c0 < cper) ? (C - c0) : cper;
    double vmax = -__longlong_as_double(0x7ff0000000000000ll);  // -inf
    double vmin = __longlong_as_double(0x7ff0000000000000ll);
    int64_t it = 0;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
        if ((int)(it % groups) != g) continue;
        const int s = (int)(it % stages);
        const uint32_t parity = (uint32_t)((it / stages) & 1);
        mbar_wait(&full[s], parity);
        if (t < cpx) {
            const InT *px = reinterpret_cast<const InT *>(ring + (size_t)s * stage_bytes) + (size_t)t * C + c0;
            double sum;
            if constexpr (CALIB)
                sum = sum_channels_div(px, px + (chunk_bytes >> 2), cn);
            else if constexpr (sizeof(InT) == 4)
                sum = sum_channels<false>(px, cn);
            else
                sum = sum_channels_raw(px, cn, scale);
            if (TPP == 2) sum += __shfl_xor_sync(0xffffffffu, sum, 1);   // cpx % 32 == 0: whole warps take this branch
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
            if (part == 0) {
                out[chunk * cpx + t] = (OutT)sum;
                vmax = fmax(vmax, (double)(OutT)sum);
                vmin = fmin(vmin, (double)(OutT)sum);
            }
        } else {
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        }
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

// Generic path (tails, unaligned bases): one warp per pixel, lanes stride
// the channels (coalesced 128-byte requests), float64 butterfly reduction.
template <typename OutT, typename InT>
__global__ void __launch_bounds__(256)
chansum_warp_kernel(const InT *__restrict__ cube, const float *__restrict__ calib, int64_t p0,
                    int64_t p1, int C, float scale, OutT *__restrict__ out,
                    unsigned long long *__restrict__ maxkey) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double vmax = -__longlong_as_double(0x7ff0000000000000ll);
    double vmin = __longlong_as_double(0x7ff0000000000000ll);
    for (int64_t p = p0 + warp0; p < p1; p += nwarps) {
        const InT *px = cube + p * C;
        double a = 0.0;
        if constexpr (sizeof(InT) != 4) {
            for (int c = lane; c < C; c += 32) a += (double)raw_value(px[c], scale);
        } else if (calib == nullptr) {
            for (int c = lane; c < C; c += 32) a += (double)ldg_stream(px + c);
        } else {
            const float *cl = calib + p * C;
            // float64 quotient of the float32 inputs, as numpy computes image/calibration
            for (int c = lane; c < C; c += 32) a += (double)ldg_stream(px + c) / (double)ldg_stream(cl + c);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            out[p] = (OutT)a;
            vmax = fmax(vmax, (double)(OutT)a);
            vmin = fmin(vmin, (double)(OutT)a);
        }
    }
    if (maxkey != nullptr) {
        vmax = warp_max(vmax);
        vmin = -warp_max(-vmin);
        if (lane == 0) {
            atomicMax(maxkey, (unsigned long long)key_of_double(vmax));
            atomicMin(maxkey + 1, (unsigned long long)key_of_double(vmin));
        }
    }
}

template <typename T>
__global__ void normalize_kernel(T *__restrict__ s, int64_t n, const unsigned long long *__restrict__ maxkey) {
    const T m = (T)double_of_key(*maxkey);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        s[i] = s[i] / m;
}

__global__ void normalize_cast_kernel(const double *__restrict__ s, int64_t n,
                                      const unsigned long long *__restrict__ maxkey, float *__restrict__ out) {
    const double m = maxkey ? double_of_key(*maxkey) : 1.0;   // NULL: plain float64 -> float32 cast
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)(s[i] / m);
}

// global min / max of an image as order-preserving keys (for the fixed-point stencil)
template <typename T>
__global__ void __launch_bounds__(256)
image_range_kernel(const T *__restrict__ a, int64_t n, unsigned long long *__restrict__ range) {
    double vmax = -__longlong_as_double(0x7ff0000000000000ll);
    double vmin = __longlong_as_double(0x7ff0000000000000ll);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)a[i];
        vmax = fmax(vmax, v);
        vmin = fmin(vmin, v);
    }
    vmax = warp_max(vmax);
    vmin = -warp_max(-vmin);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(range, (unsigned long long)key_of_double(vmax));
        atomicMin(range + 1, (unsigned long long)key_of_double(vmin));
    }
}

__global__ void maxkey_decode_kernel(const unsigned long long *__restrict__ k, double *__restrict__ out) {
    *out = double_of_key(*k);
}
__global__ void range_decode_kernel(const unsigned long long *__restrict__ k, double *__restrict__ out) {
    out[0] = double_of_key(k[0]);
    out[1] = double_of_key(k[1]);
}
__global__ void range_encode_kernel(const double *__restrict__ v, unsigned long long *__restrict__ k) {
    k[0] = key_of_double(v[0]);
    k[1] = key_of_double(v[1]);
}

template <typename OutT, typename InT>
static int chansum_launch(const InT *cube, const float *calib, int64_t npix, int C, float scale, OutT *out,
                          unsigned long long *maxkey, cudaStream_t st) {
    int64_t done = 0;
    const bool aligned = (((uintptr_t)cube) & 15u) == 0 && (calib == nullptr || (((uintptr_t)calib) & 15u) == 0);
    if (aligned) {
        const int per_px = (calib ? 2 : 1) * C * (int)sizeof(InT);   // bytes per pixel per stage
        int cpx = 0;
        for (int cand = calib ? 64 : 128; cand >= 32; cand -= 32) {   // calib: two threads per pixel
            if ((int64_t)cand * per_px * 2 <= CS_INFLIGHT_TARGET) { cpx = cand; break; }
        }
        if (cpx > 0 && npix >= cpx) {
            const int64_t stage_bytes = (int64_t)cpx * per_px;
            // the flat-field variant has ~3x the arithmetic per byte: a fourth stage + consumer group hides its
            // latency (measured 2048^2 x 95: 0.496 ms with 3 stages, 0.466 ms = 6.84 TB/s with 4)
            int64_t inflight = calib ? 200 * 1024 : CS_INFLIGHT_TARGET;
            if (const char *ev = getenv("HIPR_CS_INFLIGHT_KB")) inflight = (int64_t)atoi(ev) * 1024;   // tuning sweeps only
            int stages = (int)(inflight / stage_bytes);
            if (stages > CS_MAX_STAGES) stages = CS_MAX_STAGES;
            if (stages < 2) stages = 2;
            int groups = 1;
            for (int gcand = CS_MAX_GROUPS; gcand >= 1; --gcand)
                if (stages % gcand == 0) { groups = gcand; break; }
            const int64_t nchunks = npix / cpx;
            const size_t smem = 128 + (size_t)stages * stage_bytes;
            int64_t grid = sm_count();
            if (grid > nchunks) grid = nchunks;
            const unsigned threads = groups * CS_GROUP_THREADS + 32;
            // (one static flag per template instantiation of this function)
            if constexpr (sizeof(InT) == 4) {
                if (calib) {
                    static std::atomic<uint64_t> attr_c{0};
                    auto kern = chansum_bulk_kernel<OutT, true, float>;
                    if (first_use_on_device(attr_c)) {
                        HIPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_BUDGET));
                        HIPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    }
                    kern<<<(unsigned)grid, threads, smem, st>>>(cube, calib, nchunks, C, cpx, stages, groups, scale, out,
                                                               maxkey);
                }
            }
            if (!calib) {
                static std::atomic<uint64_t> attr_n{0};
                auto kern = chansum_bulk_kernel<OutT, false, InT>;
                if (first_use_on_device(attr_n)) {
                    HIPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_BUDGET));
                    // The SM keeps the shared-memory carve-out of the kernel that configured it while that kernel is
                    // resident.  This kernel needs 146 KB, which selects the 164 KB configuration and leaves 17 KB for
                    // the stencil CTAs of the previous field of view that run beside it; asking for the largest
                    // carve-out (it uses no L1: its data arrives by bulk copies) leaves them 81 KB.
                    HIPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                }
                kern<<<(unsigned)grid, threads, smem, st>>>(cube, nullptr, nchunks, C, cpx, stages, groups, scale, out,
                                                           maxkey);
            }
            int e = after_launch();
            if (e) return e;
            done = nchunks * cpx;
        }
    }
    if (done < npix) {
        const int64_t rest = npix - done;
        int64_t blocks = (rest + 7) / 8;  // 8 warps per block, one pixel per warp per trip
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        chansum_warp_kernel<OutT, InT><<<(unsigned)blocks, 256, 0, st>>>(cube, calib, done, npix, C, scale, out, maxkey);
        int e = after_launch();
        if (e) return e;
    }
    return HIPR_OK;
}

// used by the host-buffer pipeline: one band of rows, max key accumulated (not reset)
int chansum_band(const void *cube, int sample_bytes, float scale, int64_t npix, int C, double *out,
                 unsigned long long *maxkey, cudaStream_t st) {
    if (sample_bytes == 4) return chansum_launch<double, float>((const float *)cube, nullptr, npix, C, 1.f, out, maxkey, st);
    if (sample_bytes == 2)
        return chansum_launch<double, uint16_t>((const uint16_t *)cube, nullptr, npix, C, scale, out, maxkey, st);
    return chansum_launch<double, uint8_t>((const uint8_t *)cube, nullptr, npix, C, scale, out, maxkey, st);
}

}  // namespace hipr

using namespace hipr;

extern "C" int hipr_chansum(const float *cube_dev, const float *calib_dev, int64_t npix, int C,
                            void *sum_dev, int sum_dtype, uint64_t *maxkey_dev /* [2]: max, min */, void *stream) {
    if (!cube_dev || !sum_dev || npix <= 0 || C <= 0) return HIPR_E_ARG;
    if (sum_dtype != HIPR_F32 && sum_dtype != HIPR_F64) return HIPR_E_DTYPE;
    if ((((uintptr_t)cube_dev) & 3u) || (calib_dev && (((uintptr_t)calib_dev) & 3u))) return HIPR_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    if (maxkey_dev) {
        HIPR_CUDA(cudaMemsetAsync(maxkey_dev, 0x00, sizeof(uint64_t), st));      // max key: below all
        HIPR_CUDA(cudaMemsetAsync(maxkey_dev + 1, 0xff, sizeof(uint64_t), st));  // min key: above all
    }
    unsigned long long *mk = reinterpret_cast<unsigned long long *>(maxkey_dev);
    if (sum_dtype == HIPR_F32)
        return chansum_launch<float, float>(cube_dev, calib_dev, npix, C, 1.f, (float *)sum_dev, mk, st);
    return chansum_launch<double, float>(cube_dev, calib_dev, npix, C, 1.f, (double *)sum_dev, mk, st);
}

extern "C" int hipr_chansum_raw(const void *cube_dev, int sample_bytes, double scale, int64_t npix, int C,
                                double *sum_dev, uint64_t *maxkey_dev, void *stream) {
    if (!cube_dev || !sum_dev || npix <= 0 || C <= 0 || !(scale > 0.0)) return HIPR_E_ARG;
    if (sample_bytes != 1 && sample_bytes != 2) return HIPR_E_DTYPE;
    if (sample_bytes == 2 && (((uintptr_t)cube_dev) & 1u)) return HIPR_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    if (maxkey_dev) {
        HIPR_CUDA(cudaMemsetAsync(maxkey_dev, 0x00, sizeof(uint64_t), st));
        HIPR_CUDA(cudaMemsetAsync(maxkey_dev + 1, 0xff, sizeof(uint64_t), st));
    }
    return chansum_band(cube_dev, sample_bytes, (float)scale, npix, C, sum_dev,
                        reinterpret_cast<unsigned long long *>(maxkey_dev), st);
}

extern "C" int hipr_normalize(void *sum_dev, int sum_dtype, int64_t npix, const uint64_t *maxkey_dev,
                              void *stream) {
    if (!sum_dev || !maxkey_dev || npix <= 0) return HIPR_E_ARG;
    if (sum_dtype != HIPR_F32 && sum_dtype != HIPR_F64) return HIPR_E_DTYPE;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (npix + 1023) / 1024;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const unsigned long long *mk = reinterpret_cast<const unsigned long long *>(maxkey_dev);
    if (sum_dtype == HIPR_F32)
        normalize_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float *)sum_dev, npix, mk);
    else
        normalize_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((double *)sum_dev, npix, mk);
    return after_launch();
}

extern "C" int hipr_maxkey_decode(const uint64_t *maxkey_dev, double *max_dev, void *stream) {
    if (!maxkey_dev || !max_dev) return HIPR_E_ARG;
    maxkey_decode_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long *>(maxkey_dev), max_dev);
    return after_launch();
}

extern "C" int hipr_normalize_cast(const double *sum_dev, int64_t npix, const uint64_t *maxkey_dev, float *out_dev,
                                   void *stream) {
    if (!sum_dev || !out_dev || npix <= 0) return HIPR_E_ARG;
    int64_t blocks = (npix + 1023) / 1024;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    normalize_cast_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        sum_dev, npix, reinterpret_cast<const unsigned long long *>(maxkey_dev), out_dev);
    return after_launch();
}

extern "C" int hipr_image_range(const void *image_dev, int dtype, int64_t n, uint64_t *range_dev, void *stream) {
    if (!image_dev || !range_dev || n <= 0) return HIPR_E_ARG;
    if (dtype != HIPR_F32 && dtype != HIPR_F64) return HIPR_E_DTYPE;
    cudaStream_t st = (cudaStream_t)stream;
    HIPR_CUDA(cudaMemsetAsync(range_dev, 0x00, sizeof(uint64_t), st));
    HIPR_CUDA(cudaMemsetAsync(range_dev + 1, 0xff, sizeof(uint64_t), st));
    int64_t blocks = (n + 2047) / 2048;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    unsigned long long *rg = reinterpret_cast<unsigned long long *>(range_dev);
    if (dtype == HIPR_F32)
        image_range_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float *)image_dev, n, rg);
    else
        image_range_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((const double *)image_dev, n, rg);
    return after_launch();
}

extern "C" int hipr_range_decode(const uint64_t *range_dev, double *maxmin_dev, void *stream) {
    if (!range_dev || !maxmin_dev) return HIPR_E_ARG;
    range_decode_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long *>(range_dev),
                                                          maxmin_dev);
    return after_launch();
}

extern "C" int hipr_range_encode(const double *maxmin_dev, uint64_t *range_dev, void *stream) {
    if (!range_dev || !maxmin_dev) return HIPR_E_ARG;
    range_encode_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(maxmin_dev,
                                                          reinterpret_cast<unsigned long long *>(range_dev));
    return after_launch();
}
