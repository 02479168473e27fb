// ABI housekeeping and the host-buffer entry points: what a numpy caller binds.  The cube is
// streamed host -> device in row bands on a copy stream while the channel sum (or the per-cell
// accumulation) of the previous band runs on a compute stream, so the end-to-end time is the
// PCIe time of the cube plus a short tail.
#include <mutex>
#include <string.h>
#include <cstdlib>
#include <thread>
#include <vector>
#include "hipr_common.cuh"

namespace hipr {

std::atomic<int64_t> g_launches{0};

int chansum_band(const void *cube, int sample_bytes, float scale, int64_t npix, int C, double *out,
                 unsigned long long *maxkey, cudaStream_t st);

constexpr int NBUF = 3;
struct Workspace {
    std::mutex mu;
    cudaStream_t copy = nullptr, comp = nullptr;
    cudaEvent_t copied[NBUF] = {}, freed[NBUF] = {}, done = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;   // device-side timing of the last host call
    float last_ms = -1.f;
    void *band[NBUF] = {};
    size_t band_bytes = 0;
    void *aux[6] = {};
    size_t aux_bytes[6] = {};
    // pageable callers: page-locked staging ring the caller's array is copied through by a few host threads
    void *stage[NBUF] = {};
    size_t stage_bytes = 0;
    cudaEvent_t staged_out[NBUF] = {};   // the H2D copy out of stage[i] has completed
    // the cell table of the last hipr_cell_spectra_host call stays in aux[5] for hipr_cell_spectra_host_fetch
    int64_t last_cells = -1, last_max_label = 0;
    int last_C = 0;
    bool ready = false;
};
static Workspace g_ws;

static int ws_init(Workspace &w) {
    if (w.ready) return HIPR_OK;
    HIPR_CUDA(cudaStreamCreateWithFlags(&w.copy, cudaStreamNonBlocking));
    HIPR_CUDA(cudaStreamCreateWithFlags(&w.comp, cudaStreamNonBlocking));
    for (int i = 0; i < NBUF; ++i) {
        HIPR_CUDA(cudaEventCreateWithFlags(&w.copied[i], cudaEventDisableTiming));
        HIPR_CUDA(cudaEventCreateWithFlags(&w.freed[i], cudaEventDisableTiming));
    }
    HIPR_CUDA(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
    HIPR_CUDA(cudaEventCreate(&w.t0));
    HIPR_CUDA(cudaEventCreate(&w.t1));
    w.ready = true;
    return HIPR_OK;
}
static int ws_bands(Workspace &w, size_t bytes) {
    if (bytes <= w.band_bytes) return HIPR_OK;
    for (int i = 0; i < NBUF; ++i) {
        if (w.band[i]) cudaFree(w.band[i]);
        w.band[i] = nullptr;
    }
    w.band_bytes = 0;
    for (int i = 0; i < NBUF; ++i) HIPR_CUDA(cudaMalloc(&w.band[i], bytes));
    w.band_bytes = bytes;
    return HIPR_OK;
}
static int ws_stage(Workspace &w, size_t bytes) {
    if (bytes <= w.stage_bytes) return HIPR_OK;
    for (int i = 0; i < NBUF; ++i) {
        if (w.stage[i]) cudaFreeHost(w.stage[i]);
        w.stage[i] = nullptr;
        if (!w.staged_out[i]) HIPR_CUDA(cudaEventCreateWithFlags(&w.staged_out[i], cudaEventDisableTiming));
    }
    w.stage_bytes = 0;
    for (int i = 0; i < NBUF; ++i) HIPR_CUDA(cudaHostAlloc(&w.stage[i], bytes, cudaHostAllocDefault));
    w.stage_bytes = bytes;
    return HIPR_OK;
}

// true when `p` is ordinary pageable host memory (numpy's allocator): a direct cudaMemcpyAsync from it is staged by
// the driver at ~11 GB/s (measured: 143 ms for a 1.59 GB cube); page-locked memory goes at the PCIe rate
static bool is_pageable(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// dst <- src with a few threads (one slice each)
static void parallel_copy(void *dst, const void *src, size_t bytes, int nthreads) {
    if (nthreads <= 1 || bytes < (4u << 20)) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t slice = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    for (int t = 1; t < nthreads; ++t) {
        const size_t o = (size_t)t * slice;
        if (o >= bytes) break;
        const size_t n = (bytes - o < slice) ? bytes - o : slice;
        th.emplace_back([=] { memcpy((char *)dst + o, (const char *)src + o, n); });
    }
    memcpy(dst, src, slice < bytes ? slice : bytes);
    for (auto &t : th) t.join();
}

static int ws_aux(Workspace &w, int slot, size_t bytes) {
    if (bytes <= w.aux_bytes[slot]) return HIPR_OK;
    if (w.aux[slot]) cudaFree(w.aux[slot]);
    w.aux[slot] = nullptr;
    w.aux_bytes[slot] = 0;
    HIPR_CUDA(cudaMalloc(&w.aux[slot], bytes));
    w.aux_bytes[slot] = bytes;
    return HIPR_OK;
}

// rows per band so that a band is ~32 MiB and a whole number of rows
static int band_rows(int64_t row_bytes, int64_t nrows) {
    int64_t r = (32ll << 20) / (row_bytes > 0 ? row_bytes : 1);
    if (r < 1) r = 1;
    if (r > nrows) r = nrows;
    return (int)r;
}

}  // namespace hipr

using namespace hipr;

extern "C" int hipr_abi_version(void) { return HIPR_ABI_VERSION; }

extern "C" int64_t hipr_launch_count(void) { return g_launches.load(); }

extern "C" int hipr_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return HIPR_E_NODEVICE;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return HIPR_E_NODEVICE;
    return n;
}

extern "C" const char *hipr_error_string(int code) {
    switch (code) {
        case HIPR_OK: return "ok";
        case HIPR_E_ARG: return "invalid argument (null pointer or non-positive size)";
        case HIPR_E_DTYPE: return "unsupported dtype code";
        case HIPR_E_PATCH: return "patch_size must be odd, 3..31, and no larger than the image";
        case HIPR_E_TABLE: return "line table entry outside the patch, or table too large";
        case HIPR_E_FLAVOUR: return "unknown epilogue flavour for this entry point";
        case HIPR_E_ALIGN: return "pointer not aligned to its element type";
        case HIPR_E_RANGE: return "size exceeds the index range or the caller's capacity";
        case HIPR_E_NODEVICE: return "no CUDA device";
        case HIPR_E_UNSUPPORTED: return "request outside this entry point's fast path";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

extern "C" int hipr_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return HIPR_E_ARG;
    HIPR_CUDA(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
    return HIPR_OK;
}
extern "C" int hipr_host_free(void *ptr) {
    if (!ptr) return HIPR_OK;
    HIPR_CUDA(cudaFreeHost(ptr));
    return HIPR_OK;
}
extern "C" int hipr_host_release_workspace(void) {
    Workspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    cudaDeviceSynchronize();
    for (int i = 0; i < NBUF; ++i) {
        if (w.band[i]) cudaFree(w.band[i]);
        w.band[i] = nullptr;
    }
    w.band_bytes = 0;
    for (int i = 0; i < NBUF; ++i) {
        if (w.stage[i]) cudaFreeHost(w.stage[i]);
        w.stage[i] = nullptr;
    }
    w.stage_bytes = 0;
    for (int i = 0; i < 6; ++i) {
        if (w.aux[i]) cudaFree(w.aux[i]);
        w.aux[i] = nullptr;
        w.aux_bytes[i] = 0;
    }
    return HIPR_OK;
}

// cube_host: (H, W, C) samples of `sample_bytes` (4 = float32; 2 / 1 = raw uint16 / uint8 counts, value =
// float32(count) / float32(scale))
static int neighbor2d_host_impl(const void *cube_host, int sample_bytes, float scale, int H, int W, int C,
                                int patch_size, int n_dirs, const int32_t *table_host, int flavour, float *score_host,
                                float *sum_host, double denoise_h = 0.0) {
    if (!cube_host || !score_host || H < 1 || W < 1 || C < 1) return HIPR_E_ARG;
    Workspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    int e = ws_init(w);
    if (e) return e;
    const int64_t row_bytes = (int64_t)W * C * sample_bytes;
    const int rows = band_rows(row_bytes, H);
    if ((e = ws_bands(w, (size_t)rows * row_bytes))) return e;
    const size_t img_bytes = (size_t)H * W * 4;
    if ((e = ws_aux(w, 0, 4 * img_bytes))) return e;  // float64 sum image (+ float64 score, general parameters)
    if ((e = ws_aux(w, 1, img_bytes))) return e;      // score, then the float32 normalised sum
    if ((e = ws_aux(w, 2, 64))) return e;             // max / min keys
    double *sum_dev = (double *)w.aux[0];
    float *score_dev = (float *)w.aux[1];
    unsigned long long *key = (unsigned long long *)w.aux[2];
    HIPR_CUDA(cudaEventRecord(w.t0, w.copy));            // device clock starts before the first H2D
    HIPR_CUDA(cudaStreamWaitEvent(w.comp, w.t0, 0));
    HIPR_CUDA(cudaMemsetAsync(key, 0x00, 8, w.comp));
    HIPR_CUDA(cudaMemsetAsync(key + 1, 0xff, 8, w.comp));
    const bool pageable = is_pageable(cube_host);
    int copy_threads = (int)std::thread::hardware_concurrency();
    copy_threads = copy_threads > 8 ? 8 : copy_threads;     // measured with 8: 143 -> 48 ms per 1.59 GB cube
    if (const char *ev = getenv("HIPR_HOST_COPY_THREADS")) copy_threads = atoi(ev);
    copy_threads = copy_threads > 16 ? 16 : (copy_threads < 1 ? 1 : copy_threads);
    if (pageable && (e = ws_stage(w, (size_t)rows * row_bytes))) return e;
    int b = 0;
    for (int r0 = 0; r0 < H; r0 += rows, ++b) {
        const int nr = (H - r0 < rows) ? H - r0 : rows;
        const int slot = b % NBUF;
        const void *src = (const char *)cube_host + (int64_t)r0 * row_bytes;
        if (pageable) {
            // host threads copy band b into the page-locked ring while band b - 1 crosses PCIe and band b - 2 is summed
            if (b >= NBUF) HIPR_CUDA(cudaEventSynchronize(w.staged_out[slot]));
            parallel_copy(w.stage[slot], src, (size_t)nr * row_bytes, copy_threads);
            src = w.stage[slot];
        }
        if (b >= NBUF) HIPR_CUDA(cudaStreamWaitEvent(w.copy, w.freed[slot], 0));
        HIPR_CUDA(cudaMemcpyAsync(w.band[slot], src, (size_t)nr * row_bytes, cudaMemcpyHostToDevice, w.copy));
        if (pageable) HIPR_CUDA(cudaEventRecord(w.staged_out[slot], w.copy));
        HIPR_CUDA(cudaEventRecord(w.copied[slot], w.copy));
        HIPR_CUDA(cudaStreamWaitEvent(w.comp, w.copied[slot], 0));
        if ((e = chansum_band(w.band[slot], sample_bytes, scale, (int64_t)nr * W, C, sum_dev + (int64_t)r0 * W, key,
                              w.comp)))
            return e;
        HIPR_CUDA(cudaEventRecord(w.freed[slot], w.comp));
    }
    if (denoise_h > 0.0) {
        // syn/..._measurement.py:106-124 in full: /max -> NL-means -> float64 stencil (the denoised image is too smooth
        // for the fixed-point grid).  aux[0] holds the sums and, behind them, the denoised image; aux[3]
        // the float64 score.
        double *den = sum_dev + (size_t)H * W;
        if ((e = ws_aux(w, 3, 2 * img_bytes))) return e;
        double *score64 = (double *)w.aux[3];
        if ((e = hipr_normalize(sum_dev, HIPR_F64, (int64_t)H * W, (const uint64_t *)key, w.comp))) return e;
        if ((e = hipr_denoise_nl_means_2d(sum_dev, H, W, HIPR_F64, 7, 11, denoise_h, den, w.comp))) return e;
        if ((e = hipr_lne2d(den, H, W, W, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour, nullptr, score64, w.comp)))
            return e;
        if ((e = hipr_normalize_cast(score64, (int64_t)H * W, nullptr, score_dev, w.comp))) return e;
        HIPR_CUDA(cudaMemcpyAsync(score_host, score_dev, img_bytes, cudaMemcpyDeviceToHost, w.comp));
        if (sum_host) {
            if ((e = hipr_normalize_cast(den, (int64_t)H * W, nullptr, score_dev, w.comp))) return e;
            HIPR_CUDA(cudaMemcpyAsync(sum_host, score_dev, img_bytes, cudaMemcpyDeviceToHost, w.comp));
        }
        HIPR_CUDA(cudaEventRecord(w.t1, w.comp));
        HIPR_CUDA(cudaStreamSynchronize(w.comp));
        HIPR_CUDA(cudaEventElapsedTime(&w.last_ms, w.t0, w.t1));
        return HIPR_OK;
    }
    // fixed-point stencil for the (11, 9) table every pipeline uses; float64 kernel otherwise
    // (F1 / F2: every tile quantised with its own range, the finest grid and the same as the device-resident
    // pipeline, bit for bit; F3's epsilon needs the global range)
    const bool tile_local = (flavour == HIPR_FLAVOUR_F1 || flavour == HIPR_FLAVOUR_F2);
    e = hipr_lne2d_q(sum_dev, H, W, W, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour,
                     tile_local ? nullptr : (const uint64_t *)key, score_dev, w.comp);
    if (e == HIPR_E_TABLE && !(patch_size == 11 && n_dirs == 9)) {
        // general parameters: float64 stencil into the upper half of aux[0], then cast
        double *score64 = sum_dev + (size_t)H * W;
        if ((e = hipr_lne2d(sum_dev, H, W, W, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour,
                            (const uint64_t *)key, score64, w.comp)))
            return e;
        e = hipr_normalize_cast(score64, (int64_t)H * W, nullptr, score_dev, w.comp);
    }
    if (e) return e;
    HIPR_CUDA(cudaMemcpyAsync(score_host, score_dev, img_bytes, cudaMemcpyDeviceToHost, w.comp));
    if (sum_host) {
        if ((e = hipr_normalize_cast(sum_dev, (int64_t)H * W, (const uint64_t *)key, score_dev, w.comp))) return e;
        HIPR_CUDA(cudaMemcpyAsync(sum_host, score_dev, img_bytes, cudaMemcpyDeviceToHost, w.comp));
    }
    HIPR_CUDA(cudaEventRecord(w.t1, w.comp));            // ... and stops after the last D2H
    HIPR_CUDA(cudaStreamSynchronize(w.comp));
    HIPR_CUDA(cudaEventElapsedTime(&w.last_ms, w.t0, w.t1));
    return HIPR_OK;
}

extern "C" int hipr_neighbor2d_host(const float *cube_host, int H, int W, int C, int patch_size, int n_dirs,
                                    const int32_t *table_host, int flavour, float *score_host, float *sum_host) {
    return neighbor2d_host_impl(cube_host, 4, 1.f, H, W, C, patch_size, n_dirs, table_host, flavour, score_host, sum_host);
}

// 3-D: cube_host (X, Y, Z, C) float32 -> score_host (X, Y, Z) float32: channel sum -> /max -> edge pad -> 72 x 11 line
// profiles -> epilogue, bio/..._analysis.py:807-817 (ME2), :900-917 (F2), :1102-1125 (F3), the cube streamed in bands
// of whole x-planes under the channel sum.
extern "C" int hipr_neighbor3d_host(const float *cube_host, int X, int Y, int Z, int C, int patch_size, int n_dirs,
                                    const int32_t *table_host, int flavour, float *score_host) {
    if (!cube_host || !score_host || !table_host || X < 1 || Y < 1 || Z < 1 || C < 1) return HIPR_E_ARG;
    Workspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    int e = ws_init(w);
    if (e) return e;
    const int64_t plane_px = (int64_t)Y * Z, plane_bytes = plane_px * C * 4, nvox = (int64_t)X * plane_px;
    const int planes = band_rows(plane_bytes, X);
    if ((e = ws_bands(w, (size_t)planes * plane_bytes))) return e;
    if ((e = ws_aux(w, 0, (size_t)nvox * 8))) return e;     // float64 sum volume
    if ((e = ws_aux(w, 1, (size_t)nvox * 4))) return e;     // score
    if ((e = ws_aux(w, 2, 64))) return e;
    double *sum_dev = (double *)w.aux[0];
    float *score_dev = (float *)w.aux[1];
    unsigned long long *key = (unsigned long long *)w.aux[2];
    HIPR_CUDA(cudaEventRecord(w.t0, w.copy));
    HIPR_CUDA(cudaStreamWaitEvent(w.comp, w.t0, 0));
    HIPR_CUDA(cudaMemsetAsync(key, 0x00, 8, w.comp));
    HIPR_CUDA(cudaMemsetAsync(key + 1, 0xff, 8, w.comp));
    const bool pageable = is_pageable(cube_host);
    int copy_threads = (int)std::thread::hardware_concurrency();
    copy_threads = copy_threads > 8 ? 8 : (copy_threads < 1 ? 1 : copy_threads);
    if (pageable && (e = ws_stage(w, (size_t)planes * plane_bytes))) return e;
    int b = 0;
    for (int x0 = 0; x0 < X; x0 += planes, ++b) {
        const int nx = (X - x0 < planes) ? X - x0 : planes;
        const int slot = b % NBUF;
        const void *src = (const char *)cube_host + (int64_t)x0 * plane_bytes;
        if (pageable) {
            if (b >= NBUF) HIPR_CUDA(cudaEventSynchronize(w.staged_out[slot]));
            parallel_copy(w.stage[slot], src, (size_t)nx * plane_bytes, copy_threads);
            src = w.stage[slot];
        }
        if (b >= NBUF) HIPR_CUDA(cudaStreamWaitEvent(w.copy, w.freed[slot], 0));
        HIPR_CUDA(cudaMemcpyAsync(w.band[slot], src, (size_t)nx * plane_bytes, cudaMemcpyHostToDevice, w.copy));
        if (pageable) HIPR_CUDA(cudaEventRecord(w.staged_out[slot], w.copy));
        HIPR_CUDA(cudaEventRecord(w.copied[slot], w.copy));
        HIPR_CUDA(cudaStreamWaitEvent(w.comp, w.copied[slot], 0));
        if ((e = chansum_band(w.band[slot], 4, 1.f, (int64_t)nx * plane_px, C, sum_dev + (int64_t)x0 * plane_px, key, w.comp)))
            return e;
        HIPR_CUDA(cudaEventRecord(w.freed[slot], w.comp));
    }
    // fixed-point stencil for the reference's (11, 9, 9) table; the float64 stencil for any other table
    e = hipr_lne3d_q(sum_dev, X, Y, Z, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour, 0, (const uint64_t *)key,
                     score_dev, w.comp);
    if (e == HIPR_E_UNSUPPORTED) {
        if ((e = ws_aux(w, 3, (size_t)nvox * 8))) return e;
        double *score64 = (double *)w.aux[3];
        if ((e = hipr_lne3d(sum_dev, X, Y, Z, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour, (const uint64_t *)key,
                            score64, w.comp)))
            return e;
        e = hipr_normalize_cast(score64, nvox, nullptr, score_dev, w.comp);
    }
    if (e) return e;
    HIPR_CUDA(cudaMemcpyAsync(score_host, score_dev, (size_t)nvox * 4, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaEventRecord(w.t1, w.comp));
    HIPR_CUDA(cudaStreamSynchronize(w.comp));
    HIPR_CUDA(cudaEventElapsedTime(&w.last_ms, w.t0, w.t1));
    return HIPR_OK;
}

extern "C" int hipr_neighbor2d_host_denoise(const float *cube_host, int H, int W, int C, int patch_size, int n_dirs,
                                            const int32_t *table_host, int flavour, double denoise_h, float *score_host,
                                            float *sum_host) {
    if (!(denoise_h > 0.0)) return HIPR_E_ARG;
    return neighbor2d_host_impl(cube_host, 4, 1.f, H, W, C, patch_size, n_dirs, table_host, flavour, score_host, sum_host,
                                denoise_h);
}

extern "C" int hipr_neighbor2d_host_raw(const void *cube_host, int sample_bytes, double scale, int H, int W, int C,
                                        int patch_size, int n_dirs, const int32_t *table_host, int flavour,
                                        float *score_host, float *sum_host) {
    if (sample_bytes != 1 && sample_bytes != 2) return HIPR_E_DTYPE;
    if (!(scale > 0.0)) return HIPR_E_ARG;
    if (sample_bytes == 2 && (((uintptr_t)cube_host) & 1u)) return HIPR_E_ALIGN;
    return neighbor2d_host_impl(cube_host, sample_bytes, (float)scale, H, W, C, patch_size, n_dirs, table_host, flavour,
                                score_host, sum_host);
}

extern "C" double hipr_host_last_elapsed_ms(void) { return (double)g_ws.last_ms; }

extern "C" int hipr_cell_spectra_host(const float *cube_host, const void *labels_host, int label_bytes, int64_t npix,
                                      int64_t row_len, int C, int64_t capacity, int64_t *n_cells, int64_t *labels_out,
                                      int64_t *area_out, double *avgint_out, double *avgint_norm_out) {
    if (!cube_host || !labels_host || !n_cells || !labels_out || !area_out || !avgint_out || !avgint_norm_out ||
        npix < 1 || C < 1 || capacity < 0)
        return HIPR_E_ARG;
    if (label_bytes != 4 && label_bytes != 8) return HIPR_E_DTYPE;
    Workspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    int e = ws_init(w);
    if (e) return e;
    // labels first (small), max label back to the host to size the accumulators
    if ((e = ws_aux(w, 3, (size_t)npix * label_bytes))) return e;
    if ((e = ws_aux(w, 2, 64))) return e;
    void *labels_dev = w.aux[3];
    int64_t *scalar_dev = (int64_t *)w.aux[2];
    HIPR_CUDA(cudaMemcpyAsync(labels_dev, labels_host, (size_t)npix * label_bytes, cudaMemcpyHostToDevice, w.comp));
    if ((e = hipr_label_max(labels_dev, label_bytes, npix, scalar_dev, w.comp))) return e;
    int64_t max_label = 0;
    HIPR_CUDA(cudaMemcpyAsync(&max_label, scalar_dev, 8, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaStreamSynchronize(w.comp));
    *n_cells = 0;
    if (max_label <= 0) return HIPR_OK;
    const size_t sums_bytes = (size_t)(max_label + 1) * C * 8;
    const size_t cnt_bytes = (size_t)(max_label + 1) * 4;
    if ((e = ws_aux(w, 4, sums_bytes + cnt_bytes + 16))) return e;
    double *sums = (double *)w.aux[4];
    int32_t *counts = (int32_t *)((char *)w.aux[4] + sums_bytes);
    HIPR_CUDA(cudaMemsetAsync(w.aux[4], 0, sums_bytes + cnt_bytes + 16, w.comp));
    const int64_t px_bytes = (int64_t)C * 4;
    int64_t band_px = (32ll << 20) / px_bytes;
    const int64_t W = (row_len > 0 && npix % row_len == 0) ? row_len : npix;
    if (W < npix) {
        band_px = (band_px / W) / 32 * 32 * W;   // whole 32-row tiles per band
        if (band_px < W) band_px = W;
    } else {
        band_px &= ~31ll;
        if (band_px < 32) band_px = 32;
    }
    if ((e = ws_bands(w, (size_t)band_px * px_bytes))) return e;
    const bool pageable = is_pageable(cube_host);
    int copy_threads = (int)std::thread::hardware_concurrency();
    copy_threads = copy_threads > 8 ? 8 : (copy_threads < 1 ? 1 : copy_threads);
    if (pageable && (e = ws_stage(w, (size_t)band_px * px_bytes))) return e;
    int b = 0;
    for (int64_t p0 = 0; p0 < npix; p0 += band_px, ++b) {
        const int64_t np = (npix - p0 < band_px) ? npix - p0 : band_px;
        const int slot = b % NBUF;
        const void *src = cube_host + p0 * C;
        if (pageable) {
            if (b >= NBUF) HIPR_CUDA(cudaEventSynchronize(w.staged_out[slot]));
            parallel_copy(w.stage[slot], src, (size_t)np * px_bytes, copy_threads);
            src = w.stage[slot];
        }
        if (b >= NBUF) HIPR_CUDA(cudaStreamWaitEvent(w.copy, w.freed[slot], 0));
        HIPR_CUDA(cudaMemcpyAsync(w.band[slot], src, (size_t)np * px_bytes, cudaMemcpyHostToDevice, w.copy));
        if (pageable) HIPR_CUDA(cudaEventRecord(w.staged_out[slot], w.copy));
        HIPR_CUDA(cudaEventRecord(w.copied[slot], w.copy));
        HIPR_CUDA(cudaStreamWaitEvent(w.comp, w.copied[slot], 0));
        if ((e = hipr_cell_spectra_accumulate((const float *)w.band[slot], (const char *)labels_dev + p0 * label_bytes,
                                              label_bytes, np, (W < npix) ? W : 0, C, max_label, sums, counts, nullptr,
                                              w.comp)))
            return e;
        HIPR_CUDA(cudaEventRecord(w.freed[slot], w.comp));
    }
    // finalize into device scratch sized by max_label, then copy the valid rows back
    const size_t row_bytes = (size_t)C * 8;
    const size_t fin_bytes = 16 + (size_t)max_label * (16 + 2 * row_bytes);
    if ((e = ws_aux(w, 5, fin_bytes))) return e;
    char *fin = (char *)w.aux[5];
    int32_t *n_dev = (int32_t *)fin;
    int64_t *lab_dev = (int64_t *)(fin + 16);
    int64_t *area_dev = lab_dev + max_label;
    double *avg_dev = (double *)(area_dev + max_label);
    double *norm_dev = avg_dev + (size_t)max_label * C;
    if ((e = hipr_cell_spectra_finalize(sums, counts, max_label, C, n_dev, lab_dev, area_dev, avg_dev, norm_dev,
                                        w.comp)))
        return e;
    int32_t n = 0;
    HIPR_CUDA(cudaMemcpyAsync(&n, n_dev, 4, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaStreamSynchronize(w.comp));
    *n_cells = n;
    w.last_cells = n;
    w.last_max_label = max_label;
    w.last_C = C;
    if (n > capacity) return HIPR_E_RANGE;     // the table stays on the device: hipr_cell_spectra_host_fetch
    if (n == 0) return HIPR_OK;
    HIPR_CUDA(cudaMemcpyAsync(labels_out, lab_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(area_out, area_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(avgint_out, avg_dev, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(avgint_norm_out, norm_dev, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaStreamSynchronize(w.comp));
    return HIPR_OK;
}

extern "C" int hipr_cell_spectra_host_fetch(int64_t capacity, int64_t *labels_out, int64_t *area_out, double *avgint_out,
                                            double *avgint_norm_out) {
    if (!labels_out || !area_out || !avgint_out || !avgint_norm_out) return HIPR_E_ARG;
    Workspace &w = g_ws;
    std::lock_guard<std::mutex> lock(w.mu);
    if (w.last_cells < 0 || !w.aux[5]) return HIPR_E_ARG;      // no table to fetch
    const int64_t n = w.last_cells, max_label = w.last_max_label;
    if (n > capacity) return HIPR_E_RANGE;
    if (n == 0) return HIPR_OK;
    const size_t row_bytes = (size_t)w.last_C * 8;
    char *fin = (char *)w.aux[5];
    int64_t *lab_dev = (int64_t *)(fin + 16);
    int64_t *area_dev = lab_dev + max_label;
    double *avg_dev = (double *)(area_dev + max_label);
    double *norm_dev = avg_dev + (size_t)max_label * w.last_C;
    HIPR_CUDA(cudaMemcpyAsync(labels_out, lab_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(area_out, area_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(avgint_out, avg_dev, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(avgint_norm_out, norm_dev, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaStreamSynchronize(w.comp));
    return HIPR_OK;
}
