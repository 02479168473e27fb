// K5: per-cell mean spectra over a label image.
//
// Replaces the C x regionprops loop, syn/..._measurement.py:167-172:
//     avgint[cell, k] = mean(img[..., k][seg == label(cell)]),  rows in ascending label order,
//     avgint_norm     = avgint / max(avgint, axis=1)
// HBM-bound on the cube (4*C B/px) -- except that background pixels (label <= 0) contribute
// nothing, so their 380-byte channel vectors are never fetched.
//
// accumulate: a warp owns 32 consecutive pixels; lane = channel (c = lane + 32*j).  The warp
// walks its foreground pixels, adding channel vectors in float32 registers while the label is
// unchanged (labels come from a watershed: long runs), and flushes a run with one float64
// red.global.add per channel plus one integer add for the pixel count.  Loads for up to four
// pixels are issued before any is consumed.  Counts are integer atomics: bit-exact.
// finalize: one CTA compacts the labels present (ascending) and forms the means.
#include "hipr_common.cuh"

namespace hipr {

template <typename LabelT, int CK>
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

    for (int64_t grp = warp0; grp < ngroups; grp += nwarps) {
        const int64_t p = (grp << 5) + lane;
        long long lab = 0;
        if (p < npix) lab = (long long)labels[p];
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
            // up to four foreground pixels per trip: issue all loads, then consume
            int q[4];
            long long ql[4];
            float v[4][CK];
            int n = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                q[u] = -1;
                if (fg) {
                    q[u] = __ffs(fg) - 1;
                    fg &= fg - 1;
                    n = u + 1;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (q[u] >= 0) {
                    ql[u] = __shfl_sync(0xffffffffu, lab, q[u]);
                    const float *px = cube + ((grp << 5) + q[u]) * (int64_t)C + c_base + lane;
#pragma unroll
                    for (int j = 0; j < CK; ++j) v[u][j] = chan_ok[j] ? ldg_stream(px + 32 * j) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
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

// Single CTA: exclusive scan of (count > 0) over labels 1..max_label, then one warp per cell row.
__global__ void __launch_bounds__(1024)
cell_finalize_kernel(const double *__restrict__ sums, const int *__restrict__ counts, int64_t max_label, int C,
                     int *__restrict__ n_cells, long long *__restrict__ labels_out, long long *__restrict__ area_out,
                     double *__restrict__ avg_out, double *__restrict__ norm_out) {
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
    const int n = carry;
    if (tid == 0) *n_cells = n;
    // rows: one warp per cell
    for (int row = warp; row < n; row += 32) {
        const long long lab = labels_out[row];
        const double inv_n = (double)area_out[row];
        double mx = -__longlong_as_double(0x7ff0000000000000ll);
        bool anynan = false;
        for (int c = lane; c < C; c += 32) {
            const double a = sums[lab * C + c] / inv_n;
            avg_out[(int64_t)row * C + c] = a;
            anynan |= (a != a);
            mx = fmax(mx, a);
        }
        mx = warp_max(mx);
        // np.max propagates NaN
        if (__any_sync(0xffffffffu, anynan)) mx = __longlong_as_double(0x7ff8000000000000ll);
        for (int c = lane; c < C; c += 32)
            norm_out[(int64_t)row * C + c] = avg_out[(int64_t)row * C + c] / mx;
    }
}

template <typename LabelT>
static int accumulate_launch(const float *cube, const LabelT *labels, int64_t npix, int C, int64_t max_label,
                             double *sums, int *counts, int *overflow, cudaStream_t st) {
    const int64_t ngroups = (npix + 31) / 32;
    int64_t blocks = (ngroups + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    for (int c_base = 0; c_base < C; c_base += 128) {
        const int rem = C - c_base;
        const int ck = rem >= 97 ? 4 : (rem + 31) / 32;
        int *ovf = c_base == 0 ? overflow : nullptr;
        switch (ck) {
            case 1: cell_accumulate_kernel<LabelT, 1><<<(unsigned)blocks, 256, 0, st>>>(cube, labels, npix, C, c_base, max_label, sums, counts, ovf); break;
            case 2: cell_accumulate_kernel<LabelT, 2><<<(unsigned)blocks, 256, 0, st>>>(cube, labels, npix, C, c_base, max_label, sums, counts, ovf); break;
            case 3: cell_accumulate_kernel<LabelT, 3><<<(unsigned)blocks, 256, 0, st>>>(cube, labels, npix, C, c_base, max_label, sums, counts, ovf); break;
            default: cell_accumulate_kernel<LabelT, 4><<<(unsigned)blocks, 256, 0, st>>>(cube, labels, npix, C, c_base, max_label, sums, counts, ovf); break;
        }
        int e = after_launch();
        if (e) return e;
    }
    return HIPR_OK;
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
                                            int64_t npix, int C, int64_t max_label, double *sums_dev,
                                            int32_t *counts_dev, int32_t *overflow_dev, void *stream) {
    if (!cube_dev || !labels_dev || !sums_dev || !counts_dev || npix <= 0 || C <= 0 || max_label < 0) return HIPR_E_ARG;
    if (label_bytes != 4 && label_bytes != 8) return HIPR_E_DTYPE;
    if (npix > 0x7fffffffLL) return HIPR_E_RANGE;  // int32 pixel counts
    cudaStream_t st = (cudaStream_t)stream;
    if (label_bytes == 4)
        return accumulate_launch<int>(cube_dev, (const int *)labels_dev, npix, C, max_label, sums_dev, counts_dev, overflow_dev, st);
    return accumulate_launch<long long>(cube_dev, (const long long *)labels_dev, npix, C, max_label, sums_dev, counts_dev, overflow_dev, st);
}

extern "C" int hipr_cell_spectra_finalize(const double *sums_dev, const int32_t *counts_dev, int64_t max_label, int C,
                                          int32_t *n_cells_dev, int64_t *labels_out, int64_t *area_out,
                                          double *avgint_out, double *avgint_norm_out, void *stream) {
    if (!sums_dev || !counts_dev || !n_cells_dev || !labels_out || !area_out || !avgint_out || !avgint_norm_out ||
        C <= 0 || max_label < 0)
        return HIPR_E_ARG;
    cell_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(sums_dev, counts_dev, max_label, C, n_cells_dev,
                                                              (long long *)labels_out, (long long *)area_out,
                                                              avgint_out, avgint_norm_out);
    return after_launch();
}
