// K5: per-cell mean spectra over a label image.
//
// Replaces the C x regionprops loop, syn/..._measurement.py:167-172:
//     avgint[cell, k] = mean(img[..., k][seg == label(cell)]),  rows in ascending label order,
//     avgint_norm     = avgint / max(avgint, axis=1)
// HBM-bound on the cube (4*C B/px) -- except that background pixels (label <= 0) contribute
// nothing, so their 380-byte channel vectors are never fetched.
//
// accumulate: a warp owns 32 consecutive pixels; lane = channel (c = lane + 32*j).  The warp
// walks its foreground pixels (ballot), adding channel vectors in float32 registers while the
// label is unchanged (labels come from a watershed: long runs), and flushes a run with one float64
// red.global.add per channel plus one integer add for the pixel count.  Loads for up to four
// pixels are issued before any is consumed; the next group's labels are prefetched.  Counts are integer atomics: bit-exact.
// The kernel lives on memory-level parallelism: ~40 resident warps per SM x 12 loads each.  A
// tile-based variant with on-chip label slots (8x fewer atomics) was built and measured 2.5x
// SLOWER: its 120 registers and serial row walk cut the loads in flight; the atomics were never
// the limit (fire-and-forget REDs).  The grid is exactly one wave of resident CTAs.
// finalize: one CTA compacts the labels present (ascending); a second kernel, one warp per
// present label, forms the means and their row-max normalisation.
#include <cstdlib>
#include "hipr_common.cuh"

namespace hipr {

#ifndef HIPR_CA_BATCH
#define HIPR_CA_BATCH 6   // measured 2048^2 FOV: 4 -> 0.111 ms, 6 -> 0.106 ms, 8 -> 0.112 ms (74 registers)
#endif
constexpr int CA_BATCH = HIPR_CA_BATCH;   // foreground pixels whose loads are issued before any is consumed

template <typename LabelT, int CK, bool PREFETCH>
__global__ void __launch_bounds__(256)
cell_accumulate_kernel(const float *__restrict__ cube, const LabelT *__restrict__ labels, int64_t npix, int C,
                       int c_base, int64_t max_label, double *__restrict__ sums, int *__restrict__ counts,
                       int *__restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t ngroups = (npix + 31) >> 5;
    bool chan_ok[CK];
#pragma unroll
    for (int j = 0; j < CK; ++j) chan_ok[j] = (c_base + lane + 32 * j) < C;

    // the labels of the NEXT group are requested before the current group is processed, so a
    // background-only group costs no exposed memory latency
    long long lab_next = 0;
    if (PREFETCH && warp0 < ngroups && (warp0 << 5) + lane < npix) lab_next = (long long)labels[(warp0 << 5) + lane];
    for (int64_t grp = warp0; grp < ngroups; grp += nwarps) {
        long long lab;
        if (PREFETCH) {
            lab = lab_next;
            const int64_t gn = grp + nwarps;
            lab_next = 0;
            if (gn < ngroups && (gn << 5) + lane < npix) lab_next = (long long)labels[(gn << 5) + lane];
        } else {
            lab = 0;
            if ((grp << 5) + lane < npix) lab = (long long)labels[(grp << 5) + lane];
        }
        if (lab > max_label) {
            if (overflow) atomicAdd(overflow, 1);
            lab = 0;
        }
        unsigned fg = __ballot_sync(0xffffffffu, lab > 0);
        if (fg == 0) continue;
        long long cur = -1;
        int run = 0;
        float acc[CK];
#pragma unroll
        for (int j = 0; j < CK; ++j) acc[j] = 0.f;
        while (fg) {
            int q[CA_BATCH];
            long long ql[CA_BATCH];
            float v[CA_BATCH][CK];
            int n = 0;
#pragma unroll
            for (int u = 0; u < CA_BATCH; ++u) {
                q[u] = -1;
                if (fg) {
                    q[u] = __ffs(fg) - 1;
                    fg &= fg - 1;
                    n = u + 1;
                }
            }
#pragma unroll
            for (int u = 0; u < CA_BATCH; ++u) {
                if (q[u] >= 0) {
                    ql[u] = __shfl_sync(0xffffffffu, lab, q[u]);
                    const float *px = cube + ((grp << 5) + q[u]) * (int64_t)C + c_base + lane;
#pragma unroll
                    for (int j = 0; j < CK; ++j) v[u][j] = chan_ok[j] ? ldg_stream(px + 32 * j) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < CA_BATCH; ++u) {
                if (u < n) {
                    if (ql[u] != cur) {
                        if (run > 0) {
#pragma unroll
                            for (int j = 0; j < CK; ++j)
                                if (chan_ok[j]) atomicAdd(&sums[cur * C + c_base + lane + 32 * j], (double)acc[j]);
                            if (lane == 0 && c_base == 0) atomicAdd(&counts[cur], run);
                        }
                        cur = ql[u];
                        run = 0;
#pragma unroll
                        for (int j = 0; j < CK; ++j) acc[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < CK; ++j) acc[j] += v[u][j];
                    ++run;
                }
            }
        }
        if (run > 0) {
#pragma unroll
            for (int j = 0; j < CK; ++j)
                if (chan_ok[j]) atomicAdd(&sums[cur * C + c_base + lane + 32 * j], (double)acc[j]);
            if (lane == 0 && c_base == 0) atomicAdd(&counts[cur], run);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Bulk-copy variant (opt-in, HIPR_CELL_BULK=1; measured SLOWER than the load variant above on B200: 0.173 vs
// 0.112 ms on the 25 %-foreground synthetic FOV, 0.370 vs 0.275 ms = 4.3 vs 5.8 TB/s on a 98 %-foreground label
// image -- with two 12 KB stages per warp only 8 warps fit an SM, and their per-pixel shared-memory walk is
// latency-bound; kept as the starting point for a 16-pixel-group / more-warps version): the same run-length
// accumulation, but the channel vectors come
// through the TMA engine instead of per-lane loads.  Foreground pixels of a 32-pixel group are contiguous bytes of
// the cube (a pixel is 4 C bytes), so ONE 1-D bulk copy (cp.async.bulk, SASS UBLKCP) fetches the span from the
// group's first to its last foreground pixel -- rounded out to 16-byte boundaries -- into a shared-memory stage that
// belongs to the warp; background-only groups are never fetched.  Every warp runs its own two-stage pipeline (own
// mbarriers, no CTA-wide synchronisation): it issues the copies of groups k + 1 and k + 2 before it consumes group
// k, lane = channel straight out of shared memory (consecutive lanes, consecutive words: no bank conflicts).  Bytes in
// flight cost no registers: 8 warps x 2 stages x 12 KB = 195 KB per SM against ~55 KB for the load variant, which
// is what an HBM-bound gather needs (the load variant reached 3.6 TB/s on the bytes it fetched).
// ---------------------------------------------------------------------------------------------------------
constexpr int CBK_STAGES = 2;

template <typename LabelT, int CK>
__global__ void __launch_bounds__(256, 1)
cell_accumulate_bulk_kernel(const float *__restrict__ cube, const LabelT *__restrict__ labels, int64_t npix, int C,
                            int64_t max_label, int stage_bytes, double *__restrict__ sums, int *__restrict__ counts,
                            int *__restrict__ overflow) {
    extern __shared__ __align__(128) unsigned char cbk_smem[];
    __shared__ __align__(8) uint64_t bars[8 * CBK_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned char *my_stage = cbk_smem + (size_t)warp * CBK_STAGES * stage_bytes;
    uint64_t *my_bar = bars + warp * CBK_STAGES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < CBK_STAGES; ++s) mbar_init(&my_bar[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t policy = policy_evict_first();
    const int64_t warp0 = (int64_t)blockIdx.x * nw + warp;
    const int64_t nwarps = (int64_t)gridDim.x * nw;
    const int64_t ngroups = (npix + 31) >> 5;
    const int64_t px_bytes = (int64_t)C * 4;
    const unsigned char *cube_b = reinterpret_cast<const unsigned char *>(cube);
    bool chan_ok[CK];
#pragma unroll
    for (int j = 0; j < CK; ++j) chan_ok[j] = (lane + 32 * j) < C;

    long long lab_s[CBK_STAGES];
    unsigned fg_s[CBK_STAGES];
    int off_s[CBK_STAGES];          // byte offset of the group's FIRST foreground pixel inside the stage
    unsigned par_s[CBK_STAGES];
#pragma unroll
    for (int s = 0; s < CBK_STAGES; ++s) { lab_s[s] = 0; fg_s[s] = 0; off_s[s] = 0; par_s[s] = 0; }

    auto issue = [&](int64_t grp, int s) {
        long long lab = 0;
        if (grp < ngroups && (grp << 5) + lane < npix) lab = (long long)labels[(grp << 5) + lane];
        if (lab > max_label) {
            if (overflow) atomicAdd(overflow, 1);
            lab = 0;
        }
        const unsigned fg = __ballot_sync(0xffffffffu, lab > 0);
        lab_s[s] = lab;
        fg_s[s] = fg;
        if (fg == 0) return;
        const int first = __ffs(fg) - 1, last = 31 - __clz(fg);
        const int64_t b0 = ((grp << 5) + first) * px_bytes, b1 = ((grp << 5) + last + 1) * px_bytes;
        const int64_t a0 = b0 & ~(int64_t)15, a1 = (b1 + 15) & ~(int64_t)15;
        off_s[s] = (int)(b0 - a0);
        if (lane == 0) {
            mbar_expect_tx(&my_bar[s], (uint32_t)(a1 - a0));
            bulk_g2s(my_stage + (size_t)s * stage_bytes, cube_b + a0, (uint32_t)(a1 - a0), &my_bar[s], policy);
        }
    };
    auto consume = [&](int s) {
        unsigned fg = fg_s[s];
        if (fg == 0) return;
        mbar_wait(&my_bar[s], par_s[s]);
        par_s[s] ^= 1u;
        const int first = __ffs(fg) - 1;
        const float *base = reinterpret_cast<const float *>(my_stage + (size_t)s * stage_bytes + off_s[s]) + lane;
        const long long lab = lab_s[s];
        long long cur = -1;
        int run = 0;
        float acc[CK];
#pragma unroll
        for (int j = 0; j < CK; ++j) acc[j] = 0.f;
        while (fg) {
            const int q = __ffs(fg) - 1;
            fg &= fg - 1;
            const long long ql = __shfl_sync(0xffffffffu, lab, q);
            const float *px = base + (q - first) * C;
            float v[CK];
#pragma unroll
            for (int j = 0; j < CK; ++j) v[j] = chan_ok[j] ? px[32 * j] : 0.f;
            if (ql != cur) {
                if (run > 0) {
#pragma unroll
                    for (int j = 0; j < CK; ++j)
                        if (chan_ok[j]) atomicAdd(&sums[cur * C + lane + 32 * j], (double)acc[j]);
                    if (lane == 0) atomicAdd(&counts[cur], run);
                }
                cur = ql;
                run = 0;
#pragma unroll
                for (int j = 0; j < CK; ++j) acc[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < CK; ++j) acc[j] += v[j];
            ++run;
        }
        if (run > 0) {
#pragma unroll
            for (int j = 0; j < CK; ++j)
                if (chan_ok[j]) atomicAdd(&sums[cur * C + lane + 32 * j], (double)acc[j]);
            if (lane == 0) atomicAdd(&counts[cur], run);
        }
        __syncwarp();          // every lane has read the stage before lane 0 refills it
    };

    // prologue: the first CBK_STAGES groups of this warp are in flight before anything is consumed
#pragma unroll
    for (int s = 0; s < CBK_STAGES; ++s) issue(warp0 + s * nwarps, s);
    for (int64_t grp = warp0; grp < ngroups; grp += CBK_STAGES * nwarps) {
#pragma unroll
        for (int s = 0; s < CBK_STAGES; ++s) {
            if (grp + s * nwarps >= ngroups) break;
            consume(s);
            issue(grp + (s + CBK_STAGES) * nwarps, s);
        }
    }
}

template <typename LabelT>
__global__ void __launch_bounds__(256)
label_max_kernel(const LabelT *__restrict__ labels, int64_t npix, unsigned long long *__restrict__ out) {
    long long m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
        const long long v = (long long)labels[i];
        m = v > m ? v : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, (unsigned long long)m);
}

// Single CTA: exclusive scan of (count > 0) over labels 1..max_label -> the present labels in
// ascending order (regionprops order) and their pixel counts.
__global__ void __launch_bounds__(1024)
cell_compact_kernel(const int *__restrict__ counts, int64_t max_label, int *__restrict__ n_cells,
                    long long *__restrict__ labels_out, long long *__restrict__ area_out) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 1; base <= max_label; base += 1024) {
        const int64_t lab = base + tid;
        const int cnt = (lab <= max_label) ? counts[lab] : 0;
        const int flag = cnt > 0;
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;  // inclusive totals
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - flag;
        if (flag) {
            labels_out[before] = lab;
            area_out[before] = cnt;
        }
        __syncthreads();
        if (tid == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) *n_cells = carry;
}

// One warp per present label: mean spectrum and its row-max normalisation.
__global__ void __launch_bounds__(256)
cell_rows_kernel(const double *__restrict__ sums, int C, const int *__restrict__ n_cells,
                 const long long *__restrict__ labels_out, const long long *__restrict__ area_out,
                 double *__restrict__ avg_out, double *__restrict__ norm_out) {
    const int lane = threadIdx.x & 31;
    const int n = *n_cells;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += nwarps) {
        const long long lab = labels_out[row];
        const double npx = (double)area_out[row];
        double mx = -__longlong_as_double(0x7ff0000000000000ll);
        bool anynan = false;
        for (int c = lane; c < C; c += 32) {
            const double a = sums[lab * C + c] / npx;
            avg_out[row * C + c] = a;
            anynan |= (a != a);
            mx = fmax(mx, a);
        }
        mx = warp_max(mx);
        // np.max propagates NaN
        if (__any_sync(0xffffffffu, anynan)) mx = __longlong_as_double(0x7ff8000000000000ll);
        for (int c = lane; c < C; c += 32) {
            const double a = sums[lab * C + c] / npx;
            norm_out[row * C + c] = a / mx;
        }
    }
}

template <typename LabelT>
static int accumulate_launch(const float *cube, const LabelT *labels, int64_t npix, int64_t row_len, int C,
                             int64_t max_label, double *sums, int *counts, int *overflow, cudaStream_t st) {
    (void)row_len;   // reserved: the run-length kernel treats the label image as a flat array
    const int64_t ngroups = (npix + 31) / 32;
    // bulk-copy variant: needs spans that can be rounded out to 16 bytes inside the cube
    static const bool bulk_on = [] { const char *e = getenv("HIPR_CELL_BULK"); return e && e[0] == '1'; }();
    if (bulk_on && C <= 128 && (((uintptr_t)cube) & 15) == 0 && (npix * (int64_t)C) % 4 == 0) {
        const int stage_bytes = (32 * C * 4 + 16 + 127) / 128 * 128;
        int warps = 8;
        while (warps > 1 && (size_t)warps * CBK_STAGES * stage_bytes > 200 * 1024) --warps;
        const size_t smem = (size_t)warps * CBK_STAGES * stage_bytes;
        int64_t blocks = sm_count();
        if (blocks > (ngroups + warps - 1) / warps) blocks = (ngroups + warps - 1) / warps;
        const int ck = (C + 31) / 32;
#define HIPR_CELL_BULK_LAUNCH(CKV)                                                                                  \
    do {                                                                                                            \
        auto kern = cell_accumulate_bulk_kernel<LabelT, CKV>;                                                       \
        static std::atomic<uint64_t> attr_done{0};                                                                  \
        if (first_use_on_device(attr_done)) HIPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
        kern<<<(unsigned)blocks, warps * 32, smem, st>>>(cube, labels, npix, C, max_label, stage_bytes, sums, counts, overflow); \
    } while (0)
        switch (ck) {
            case 1: HIPR_CELL_BULK_LAUNCH(1); break;
            case 2: HIPR_CELL_BULK_LAUNCH(2); break;
            case 3: HIPR_CELL_BULK_LAUNCH(3); break;
            default: HIPR_CELL_BULK_LAUNCH(4); break;
        }
#undef HIPR_CELL_BULK_LAUNCH
        return after_launch();
    }
    for (int c_base = 0; c_base < C; c_base += 128) {
        const int rem = C - c_base;
        const int ck = rem >= 97 ? 4 : (rem + 31) / 32;
        int *ovf = c_base == 0 ? overflow : nullptr;
#define HIPR_CELL_LAUNCH(CKV)                                                                                    \
    do {                                                                                                         \
        /* label prefetch: 0.094 vs 0.111 ms per 2048^2 FOV; forcing 6 CTAs/SM (40 regs, spills): 0.116 */      \
        auto kern = cell_accumulate_kernel<LabelT, CKV, true>;                                                   \
        static int per_sm = 0;                                                                                   \
        if (per_sm == 0) {                                                                                       \
            HIPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));                     \
            if (per_sm < 1) per_sm = 1;                                                                          \
        }                                                                                                        \
        int64_t blocks = (int64_t)sm_count() * per_sm; /* exactly one wave: no tail of a partial wave */         \
        if (blocks > (ngroups + 7) / 8) blocks = (ngroups + 7) / 8;                                              \
        kern<<<(unsigned)blocks, 256, 0, st>>>(cube, labels, npix, C, c_base, max_label, sums, counts, ovf);     \
    } while (0)
        switch (ck) {
            case 1: HIPR_CELL_LAUNCH(1); break;
            case 2: HIPR_CELL_LAUNCH(2); break;
            case 3: HIPR_CELL_LAUNCH(3); break;
            default: HIPR_CELL_LAUNCH(4); break;
        }
#undef HIPR_CELL_LAUNCH
        int e = after_launch();
        if (e) return e;
    }
    return HIPR_OK;
}

int cell_compact_launch(const int *counts, int64_t max_label, int *n_cells, long long *labels_out, long long *area_out,
                        cudaStream_t st) {
    cell_compact_kernel<<<1, 1024, 0, st>>>(counts, max_label, n_cells, labels_out, area_out);
    return after_launch();
}

}  // namespace hipr

using namespace hipr;

extern "C" int hipr_label_max(const void *labels_dev, int label_bytes, int64_t npix, int64_t *max_dev, void *stream) {
    if (!labels_dev || !max_dev || npix <= 0) return HIPR_E_ARG;
    if (label_bytes != 4 && label_bytes != 8) return HIPR_E_DTYPE;
    cudaStream_t st = (cudaStream_t)stream;
    HIPR_CUDA(cudaMemsetAsync(max_dev, 0, sizeof(int64_t), st));
    int64_t blocks = (npix + 2047) / 2048;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (label_bytes == 4)
        label_max_kernel<int><<<(unsigned)blocks, 256, 0, st>>>((const int *)labels_dev, npix, (unsigned long long *)max_dev);
    else
        label_max_kernel<long long><<<(unsigned)blocks, 256, 0, st>>>((const long long *)labels_dev, npix, (unsigned long long *)max_dev);
    return after_launch();
}

extern "C" int hipr_cell_spectra_accumulate(const float *cube_dev, const void *labels_dev, int label_bytes,
                                            int64_t npix, int64_t row_len, int C, int64_t max_label,
                                            double *sums_dev, int32_t *counts_dev, int32_t *overflow_dev,
                                            void *stream) {
    if (!cube_dev || !labels_dev || !sums_dev || !counts_dev || npix <= 0 || C <= 0 || max_label < 0) return HIPR_E_ARG;
    if (label_bytes != 4 && label_bytes != 8) return HIPR_E_DTYPE;
    if (npix > 0x7fffffffLL) return HIPR_E_RANGE;  // int32 pixel counts
    cudaStream_t st = (cudaStream_t)stream;
    if (label_bytes == 4)
        return accumulate_launch<int>(cube_dev, (const int *)labels_dev, npix, row_len, C, max_label, sums_dev, counts_dev, overflow_dev, st);
    return accumulate_launch<long long>(cube_dev, (const long long *)labels_dev, npix, row_len, C, max_label, sums_dev, counts_dev, overflow_dev, st);
}

extern "C" int hipr_cell_spectra_reset(double *sums_dev, int32_t *counts_dev, int64_t max_label, int C, void *stream) {
    if (!sums_dev || !counts_dev || max_label < 0 || C <= 0) return HIPR_E_ARG;
    HIPR_CUDA(cudaMemsetAsync(sums_dev, 0, (size_t)(max_label + 1) * C * sizeof(double), (cudaStream_t)stream));
    HIPR_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)(max_label + 1) * sizeof(int32_t), (cudaStream_t)stream));
    return HIPR_OK;
}

extern "C" int hipr_cell_spectra_finalize(const double *sums_dev, const int32_t *counts_dev, int64_t max_label, int C,
                                          int32_t *n_cells_dev, int64_t *labels_out, int64_t *area_out,
                                          double *avgint_out, double *avgint_norm_out, void *stream) {
    if (!sums_dev || !counts_dev || !n_cells_dev || !labels_out || !area_out || !avgint_out || !avgint_norm_out ||
        C <= 0 || max_label < 0)
        return HIPR_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int e = cell_compact_launch(counts_dev, max_label, n_cells_dev, (long long *)labels_out, (long long *)area_out, st);
    if (e) return e;
    int64_t blocks = (max_label + 7) / 8;   // upper bound on the rows: one warp each
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cell_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(sums_dev, C, n_cells_dev, (const long long *)labels_out,
                                                       (const long long *)area_out, avgint_out, avgint_norm_out);
    return after_launch();
}
