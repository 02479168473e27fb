// ABI housekeeping and the host-buffer entry points: what a numpy caller binds.  The cube is
// streamed host -> device in row bands on a copy stream while the channel sum (or the per-cell
// accumulation) of the previous band runs on a compute stream, so the end-to-end time is the
// PCIe time of the cube plus a short tail.
#include <mutex>
#include <string.h>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cstdio>
#include "hipr_common.cuh"

// in this file a failing CUDA call also reports where, when HIPR_DEBUG is set (the host pipelines enqueue dozens of
// calls per entry point; the ABI's return value only carries the code)
#undef HIPR_CUDA
#define HIPR_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            if (getenv("HIPR_DEBUG")) fprintf(stderr, "hipr: %s:%d: %s -> %s\n", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (int)_e;                                                                         \
        }                                                                                           \
    } while (0)

namespace hipr {

std::atomic<int64_t> g_launches{0};

int chansum_band(const void *cube, int sample_bytes, float scale, int64_t npix, int C, double *out,
                 unsigned long long *maxkey, cudaStream_t st);
template <typename T>
int gather_launch(const T *src, int64_t stride_a, int64_t stride_b, int inner, int64_t nrows, int rowlen, int K,
                  const int *offs, T *out, cudaStream_t st);
int check_table(const int32_t *table, int n_dirs, int P, int ndim);

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
    // batch entry point (hipr_neighbor2d_host_batch): a third stream for denoise + stencil + score read-back of FOV i
    // while FOV i + 1 is uploaded and summed, and two sets of per-FOV buffers
    cudaStream_t post = nullptr;
    cudaEvent_t summed[2] = {}, post_done[2] = {};
    void *baux[2][4] = {};
    size_t baux_bytes[2][4] = {};
    // the cell table of the last hipr_cell_spectra_host call stays in aux[5] for hipr_cell_spectra_host_fetch
    int64_t last_cells = -1, last_max_label = 0;
    int last_C = 0;
    bool ready = false;
};
// One workspace per device: its streams, events and buffers belong to the device that was current when they were
// created, so a process that drives several GPUs (or switches devices between calls) never shares them.
constexpr int kMaxDevices = 32;
static Workspace g_ws_dev[kMaxDevices];
static Workspace *ws_current() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    return &g_ws_dev[dev];
}
static thread_local float t_last_ms = -1.f;      // hipr_host_last_elapsed_ms: the calling thread's last host call

// An error return after the first enqueue must not leave copies from the caller's buffers (or into them) in flight:
// the caller may free them.  Unless dismissed, the guard drains both streams on scope exit.
struct DrainOnError {
    Workspace &w;
    bool armed = true;
    explicit DrainOnError(Workspace &ws) : w(ws) {}
    void dismiss() { armed = false; }
    ~DrainOnError() {
        if (!armed) return;
        if (w.copy) cudaStreamSynchronize(w.copy);
        if (w.comp) cudaStreamSynchronize(w.comp);
        if (w.post) cudaStreamSynchronize(w.post);
    }
};

// host threads that copy a pageable array into the page-locked staging ring (HIPR_HOST_COPY_THREADS overrides)
static int host_copy_threads() {
    int n = (int)std::thread::hardware_concurrency();
    n = n > 8 ? 8 : n;                                      // measured with 8: 143 -> 48 ms per 1.59 GB cube
    if (const char *ev = getenv("HIPR_HOST_COPY_THREADS")) n = atoi(ev);
    return n > 16 ? 16 : (n < 1 ? 1 : n);
}

static int ws_init(Workspace &w) {
    if (w.ready) return HIPR_OK;
    HIPR_CUDA(cudaStreamCreateWithFlags(&w.copy, cudaStreamNonBlocking));
    // the compute stream (channel sums of the bands as they arrive) has the highest priority: the hardware dispatches a
    // later kernel's CTAs only when the earlier one has none left to dispatch, unless its stream's priority is higher --
    // without it a band's channel sum waits behind the whole denoise grid of the previous FOV on the `post` stream, the
    // band ring fills and the upload stalls (batch entry point: 33.6 ms per FOV instead of the PCIe time)
    int prio_lo = 0, prio_hi = 0;
    HIPR_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    HIPR_CUDA(cudaStreamCreateWithPriority(&w.comp, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < NBUF; ++i) {
        HIPR_CUDA(cudaEventCreateWithFlags(&w.copied[i], cudaEventDisableTiming));
        HIPR_CUDA(cudaEventCreateWithFlags(&w.freed[i], cudaEventDisableTiming));
    }
    HIPR_CUDA(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
    HIPR_CUDA(cudaStreamCreateWithFlags(&w.post, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        HIPR_CUDA(cudaEventCreateWithFlags(&w.summed[i], cudaEventDisableTiming));
        HIPR_CUDA(cudaEventCreateWithFlags(&w.post_done[i], cudaEventDisableTiming));
    }
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

static int ws_baux(Workspace &w, int set, int slot, size_t bytes) {
    if (bytes <= w.baux_bytes[set][slot]) return HIPR_OK;
    if (w.baux[set][slot]) cudaFree(w.baux[set][slot]);
    w.baux[set][slot] = nullptr;
    w.baux_bytes[set][slot] = 0;
    HIPR_CUDA(cudaMalloc(&w.baux[set][slot], bytes));
    w.baux_bytes[set][slot] = bytes;
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
static void fov_cache_release_current_device();
extern "C" int hipr_host_release_workspace(void) {
    fov_cache_release_current_device();
    Workspace *wp = ws_current();
    if (!wp) return HIPR_E_NODEVICE;
    Workspace &w = *wp;
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
    for (int k = 0; k < 2; ++k)
        for (int i = 0; i < 4; ++i) {
            if (w.baux[k][i]) cudaFree(w.baux[k][i]);
            w.baux[k][i] = nullptr;
            w.baux_bytes[k][i] = 0;
        }
    return HIPR_OK;
}

// cube_host: (H, W, C) samples of `sample_bytes` (4 = float32; 2 / 1 = raw uint16 / uint8 counts, value =
// float32(count) / float32(scale))
static int neighbor2d_host_impl(const void *cube_host, int sample_bytes, float scale, int H, int W, int C,
                                int patch_size, int n_dirs, const int32_t *table_host, int flavour, float *score_host,
                                float *sum_host, double denoise_h = 0.0) {
    if (!cube_host || !score_host || H < 1 || W < 1 || C < 1) return HIPR_E_ARG;
    Workspace *wp = ws_current();
    if (!wp) return HIPR_E_NODEVICE;
    Workspace &w = *wp;
    std::lock_guard<std::mutex> lock(w.mu);
    int e = ws_init(w);
    if (e) return e;
    DrainOnError guard(w);
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
    const int copy_threads = host_copy_threads();
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
        t_last_ms = w.last_ms;
        guard.dismiss();
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
    t_last_ms = w.last_ms;
    guard.dismiss();
    return HIPR_OK;
}

extern "C" int hipr_neighbor2d_host(const float *cube_host, int H, int W, int C, int patch_size, int n_dirs,
                                    const int32_t *table_host, int flavour, float *score_host, float *sum_host) {
    return neighbor2d_host_impl(cube_host, 4, 1.f, H, W, C, patch_size, n_dirs, table_host, flavour, score_host, sum_host);
}

// Results that are far larger than their inputs (the literal gathers: 792 B per pixel, 576 B per voxel) leave the
// device in row bands: `produce(r0, nr, band_dev, stream)` fills a band of `nr` rows on the compute stream while the
// previous bands cross PCIe on the copy stream -- straight into out_host when it is page-locked, otherwise through
// the page-locked staging ring and a few host threads (a direct device -> pageable copy is staged by the driver at
// a few GB/s; measured 1.58 s for a 3.3 GB result against 0.10 s this way and 0.06 s into page-locked memory).
template <typename Produce>
static int banded_to_host(Workspace &w, int64_t nrows, int64_t row_bytes, void *out_host, Produce produce) {
    int e;
    const int rows = band_rows(row_bytes, nrows);
    if ((e = ws_bands(w, (size_t)rows * row_bytes))) return e;
    const bool pageable = is_pageable(out_host);
    const int copy_threads = host_copy_threads();
    if (pageable && (e = ws_stage(w, (size_t)rows * row_bytes))) return e;
    struct Pending { int slot; int64_t off; size_t bytes; };
    std::vector<Pending> pend;          // pageable: bands whose staged copy still has to reach out_host
    size_t drained = 0;
    auto drain = [&](size_t upto) -> int {
        for (; drained < upto; ++drained) {
            const Pending &q = pend[drained];
            HIPR_CUDA(cudaEventSynchronize(w.staged_out[q.slot]));
            parallel_copy((char *)out_host + q.off, w.stage[q.slot], q.bytes, copy_threads);
        }
        return HIPR_OK;
    };
    int b = 0;
    for (int64_t r0 = 0; r0 < nrows; r0 += rows, ++b) {
        const int nr = (int)((nrows - r0 < rows) ? nrows - r0 : rows);
        const int slot = b % NBUF;
        const size_t bytes = (size_t)nr * row_bytes;
        if (b >= NBUF) HIPR_CUDA(cudaStreamWaitEvent(w.comp, w.freed[slot], 0));   // its previous band has left the device
        if ((e = produce(r0, nr, w.band[slot], w.comp))) return e;
        HIPR_CUDA(cudaEventRecord(w.copied[slot], w.comp));
        HIPR_CUDA(cudaStreamWaitEvent(w.copy, w.copied[slot], 0));
        if (pageable) {
            // stage[slot] is free once the host has copied band b - NBUF out of it
            if (b >= NBUF && (e = drain((size_t)(b - NBUF + 1)))) return e;
            HIPR_CUDA(cudaMemcpyAsync(w.stage[slot], w.band[slot], bytes, cudaMemcpyDeviceToHost, w.copy));
            HIPR_CUDA(cudaEventRecord(w.staged_out[slot], w.copy));
            pend.push_back(Pending{slot, r0 * row_bytes, bytes});
        } else {
            HIPR_CUDA(cudaMemcpyAsync((char *)out_host + r0 * row_bytes, w.band[slot], bytes, cudaMemcpyDeviceToHost, w.copy));
        }
        HIPR_CUDA(cudaEventRecord(w.freed[slot], w.copy));
    }
    if (pageable && (e = drain(pend.size()))) return e;
    HIPR_CUDA(cudaEventRecord(w.t1, w.copy));
    HIPR_CUDA(cudaStreamSynchronize(w.copy));
    HIPR_CUDA(cudaStreamSynchronize(w.comp));
    HIPR_CUDA(cudaEventElapsedTime(&w.last_ms, w.t0, w.t1));
    t_last_ms = w.last_ms;
    return HIPR_OK;
}

// The strict drop-in of line_profile_2d_v2 (eco/neighbor2d.pyx:8-64) from and to HOST arrays: padded float64 image in,
// the literal (H, W, n_dirs, P) float64 gather out -- 792 B per pixel at (11, 9), 3.3 GB for a 2048^2 image, so the
// call is the PCIe time of the OUTPUT.  The image is uploaded once; the gather runs in row bands of ~32 MiB of output.
extern "C" int hipr_line_profile_2d_host(const double *image_padded_host, int Hp, int Wp, int patch_size, int n_dirs,
                                         const int32_t *table_host, double *out_host) {
    if (!image_padded_host || !out_host) return HIPR_E_ARG;
    int e = check_table(table_host, n_dirs, patch_size, 2);
    if (e) return e;
    const int P = patch_size, H = Hp - (P - 1), W = Wp - (P - 1), K = n_dirs * P;
    if (H < 1 || W < 1) return HIPR_E_PATCH;
    int lin[HIPR_MAX_TABLE];
    for (int i = 0; i < K; ++i) {
        const int dy = table_host[2 * i], dx = table_host[2 * i + 1];
        if (dy < 0 || dy >= P || dx < 0 || dx >= P) return HIPR_E_TABLE;
        lin[i] = dy * Wp + dx;
    }
    Workspace *wp = ws_current();
    if (!wp) return HIPR_E_NODEVICE;
    Workspace &w = *wp;
    std::lock_guard<std::mutex> lock(w.mu);
    if ((e = ws_init(w))) return e;
    DrainOnError guard(w);
    if ((e = ws_aux(w, 0, (size_t)Hp * Wp * sizeof(double)))) return e;
    double *img_dev = (double *)w.aux[0];
    HIPR_CUDA(cudaEventRecord(w.t0, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(img_dev, image_padded_host, (size_t)Hp * Wp * sizeof(double), cudaMemcpyHostToDevice, w.comp));
    e = banded_to_host(w, H, (int64_t)W * K * sizeof(double), out_host,
                       [&](int64_t r0, int nr, void *band, cudaStream_t st) {
                           return gather_launch<double>(img_dev + r0 * Wp, Wp, 0, 1, nr, W, K, lin, (double *)band, st);
                       });
    if (e) return e;
    guard.dismiss();
    return HIPR_OK;
}

// line_profile_v2 (bio/neighbor.pyx:115-181) from and to host arrays: (X, Y, Z, n_dirs, P) float64, 6,336 B per voxel at
// (11, 9, 9) -- the reference's callers chunk the volume to 100^2 / 200^2 columns for that reason
// (bio/..._analysis.py:900-904, 1105-1112); a 100 x 100 x 64 chunk is a 4 GB result.  Bands of whole x-planes.
extern "C" int hipr_line_profile_3d_host(const double *volume_padded_host, int Xp, int Yp, int Zp, int patch_size, int n_dirs,
                                         const int32_t *table_host, double *out_host) {
    if (!volume_padded_host || !out_host || patch_size < 1) return HIPR_E_ARG;
    const int P = patch_size, X = Xp - (P - 1), Y = Yp - (P - 1), Z = Zp - (P - 1);
    if (X < 1 || Y < 1 || Z < 1) return HIPR_E_PATCH;
    Workspace *wp = ws_current();
    if (!wp) return HIPR_E_NODEVICE;
    Workspace &w = *wp;
    std::lock_guard<std::mutex> lock(w.mu);
    int e = ws_init(w);
    if (e) return e;
    DrainOnError guard(w);
    const size_t vol_bytes = (size_t)Xp * Yp * Zp * sizeof(double);
    if ((e = ws_aux(w, 0, vol_bytes))) return e;
    double *vol_dev = (double *)w.aux[0];
    HIPR_CUDA(cudaEventRecord(w.t0, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(vol_dev, volume_padded_host, vol_bytes, cudaMemcpyHostToDevice, w.comp));
    e = banded_to_host(w, X, (int64_t)Y * Z * n_dirs * P * sizeof(double), out_host,
                       [&](int64_t x0, int nx, void *band, cudaStream_t st) {
                           return hipr_line_profile_3d(vol_dev + x0 * (int64_t)Yp * Zp, nx + P - 1, Yp, Zp, HIPR_F64, P, n_dirs,
                                                       table_host, band, st);
                       });
    if (e) return e;
    guard.dismiss();
    return HIPR_OK;
}

// The same for line_profile_memory_efficient_v2 (bio/neighbor.pyx:186-263), the 3-D stencil the z-stack pipelines
// call (bio/..._analysis.py:456, :812): padded float64 volume (Xp, Yp, Zp) in, the (X, Y, Z, n_dirs) float64
// per-direction values out (576 B per voxel at 72 directions), in bands of whole x-planes.
extern "C" int hipr_lne3d_dirs_host(const double *volume_padded_host, int Xp, int Yp, int Zp, int patch_size, int n_dirs,
                                    const int32_t *table_host, double *out_host) {
    if (!volume_padded_host || !out_host || patch_size < 1) return HIPR_E_ARG;
    const int P = patch_size, X = Xp - (P - 1), Y = Yp - (P - 1), Z = Zp - (P - 1);
    if (X < 1 || Y < 1 || Z < 1) return HIPR_E_PATCH;
    Workspace *wp = ws_current();
    if (!wp) return HIPR_E_NODEVICE;
    Workspace &w = *wp;
    std::lock_guard<std::mutex> lock(w.mu);
    int e = ws_init(w);
    if (e) return e;
    DrainOnError guard(w);
    const size_t vol_bytes = (size_t)Xp * Yp * Zp * sizeof(double);
    if ((e = ws_aux(w, 0, vol_bytes))) return e;
    double *vol_dev = (double *)w.aux[0];
    HIPR_CUDA(cudaEventRecord(w.t0, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(vol_dev, volume_padded_host, vol_bytes, cudaMemcpyHostToDevice, w.comp));
    e = banded_to_host(w, X, (int64_t)Y * Z * n_dirs * sizeof(double), out_host,
                       [&](int64_t x0, int nx, void *band, cudaStream_t st) {
                           // output planes [x0, x0 + nx) read padded planes [x0, x0 + nx + P - 1)
                           return hipr_lne3d_dirs(vol_dev + x0 * (int64_t)Yp * Zp, nx + P - 1, Yp, Zp, 1, HIPR_F64, P, n_dirs,
                                                  table_host, nullptr, band, st);
                       });
    if (e) return e;
    guard.dismiss();
    return HIPR_OK;
}

// 3-D: cube_host (X, Y, Z, C) float32 -> score_host (X, Y, Z) float32: channel sum -> /max -> edge pad -> 72 x 11 line
// profiles -> epilogue, bio/..._analysis.py:807-817 (ME2), :900-917 (F2), :1102-1125 (F3), the cube streamed in bands
// of whole x-planes under the channel sum.
static int neighbor3d_host_impl(const float *cube_host, int X, int Y, int Z, int C, int patch_size, int n_dirs,
                                const int32_t *table_host, int flavour, float *score_host, double denoise_h,
                                int denoise_distance) {
    if (!cube_host || !score_host || !table_host || X < 1 || Y < 1 || Z < 1 || C < 1) return HIPR_E_ARG;
    Workspace *wp = ws_current();
    if (!wp) return HIPR_E_NODEVICE;
    Workspace &w = *wp;
    std::lock_guard<std::mutex> lock(w.mu);
    int e = ws_init(w);
    if (e) return e;
    DrainOnError guard(w);
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
    const int copy_threads = host_copy_threads();
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
    if (denoise_h > 0.0) {
        // bio/..._analysis.py:452-462 in full: /max -> 3-D NL-means -> float64 stencil (the denoised volume is too smooth
        // for the fixed-point grid, as in 2-D) -> float32 score.  aux[3]: denoised volume, then the float64 score;
        // aux[4]: the reflect-padded copy the denoise works on
        const int64_t wbytes = hipr_denoise_nl_means_3d_workspace(X, Y, Z, denoise_distance);
        if (wbytes < 0) return HIPR_E_UNSUPPORTED;
        if ((e = ws_aux(w, 3, (size_t)nvox * 8))) return e;
        if ((e = ws_aux(w, 4, (size_t)wbytes))) return e;
        double *den = (double *)w.aux[3];
        if ((e = hipr_normalize(sum_dev, HIPR_F64, nvox, (const uint64_t *)key, w.comp))) return e;
        if ((e = hipr_denoise_nl_means_3d(sum_dev, X, Y, Z, HIPR_F64, 7, denoise_distance, denoise_h, den, w.aux[4], wbytes, w.comp)))
            return e;
        if ((e = hipr_lne3d(den, X, Y, Z, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour, nullptr, sum_dev, w.comp))) return e;
        if ((e = hipr_normalize_cast(sum_dev, nvox, nullptr, score_dev, w.comp))) return e;
        HIPR_CUDA(cudaMemcpyAsync(score_host, score_dev, (size_t)nvox * 4, cudaMemcpyDeviceToHost, w.comp));
        HIPR_CUDA(cudaEventRecord(w.t1, w.comp));
        HIPR_CUDA(cudaStreamSynchronize(w.comp));
        HIPR_CUDA(cudaEventElapsedTime(&w.last_ms, w.t0, w.t1));
        t_last_ms = w.last_ms;
        guard.dismiss();
        return HIPR_OK;
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
    t_last_ms = w.last_ms;
    guard.dismiss();
    return HIPR_OK;
}

extern "C" int hipr_neighbor3d_host(const float *cube_host, int X, int Y, int Z, int C, int patch_size, int n_dirs,
                                    const int32_t *table_host, int flavour, float *score_host) {
    return neighbor3d_host_impl(cube_host, X, Y, Z, C, patch_size, n_dirs, table_host, flavour, score_host, 0.0, 11);
}

extern "C" int hipr_neighbor3d_host_denoise(const float *cube_host, int X, int Y, int Z, int C, int patch_size, int n_dirs,
                                            const int32_t *table_host, int flavour, double denoise_h, int denoise_distance,
                                            float *score_host) {
    if (!(denoise_h > 0.0)) return HIPR_E_ARG;
    return neighbor3d_host_impl(cube_host, X, Y, Z, C, patch_size, n_dirs, table_host, flavour, score_host, denoise_h,
                                denoise_distance);
}

extern "C" int hipr_neighbor2d_host_denoise(const float *cube_host, int H, int W, int C, int patch_size, int n_dirs,
                                            const int32_t *table_host, int flavour, double denoise_h, float *score_host,
                                            float *sum_host) {
    if (!(denoise_h > 0.0)) return HIPR_E_ARG;
    return neighbor2d_host_impl(cube_host, 4, 1.f, H, W, C, patch_size, n_dirs, table_host, flavour, score_host, sum_host,
                                denoise_h);
}

// A batch of fields of view from host memory (config 3: hundreds of FOVs per run; the scripts process them one after
// another, syn/..._measurement.py:161-173 inside the Snakefile's per-sample loop).  FOV i + 1 is uploaded band by band
// and summed on the copy / compute streams while FOV i's normalisation, NL-means denoise (denoise_h > 0: lines 106-124
// in full), stencil and score read-back run on a third stream from a second set of buffers: per FOV the call costs the
// PCIe time of the cube, where one hipr_neighbor2d_host_denoise call per FOV costs that plus ~6 ms of denoise + stencil.
// Results are bit-identical to the single-FOV entry points.
extern "C" int hipr_neighbor2d_host_batch(const float *const *cubes_host, int n_fov, int H, int W, int C, int patch_size,
                                          int n_dirs, const int32_t *table_host, int flavour, double denoise_h,
                                          float *const *scores_host) {
    if (!cubes_host || !scores_host || n_fov < 1 || H < 1 || W < 1 || C < 1 || denoise_h < 0.0) return HIPR_E_ARG;
    for (int i = 0; i < n_fov; ++i)
        if (!cubes_host[i] || !scores_host[i]) return HIPR_E_ARG;
    Workspace *wp = ws_current();
    if (!wp) return HIPR_E_NODEVICE;
    Workspace &w = *wp;
    std::lock_guard<std::mutex> lock(w.mu);
    int e = ws_init(w);
    if (e) return e;
    DrainOnError guard(w);
    const int64_t row_bytes = (int64_t)W * C * 4;
    const int rows = band_rows(row_bytes, H);
    if ((e = ws_bands(w, (size_t)rows * row_bytes))) return e;
    const size_t img_bytes = (size_t)H * W * 4;
    for (int k = 0; k < 2; ++k) {
        if ((e = ws_baux(w, k, 0, 4 * img_bytes))) return e;   // float64 sums + float64 denoised image
        if ((e = ws_baux(w, k, 1, img_bytes))) return e;       // float32 score
        if ((e = ws_baux(w, k, 2, 64))) return e;              // range keys
        if ((e = ws_baux(w, k, 3, 2 * img_bytes))) return e;   // float64 score (denoise / general parameters)
    }
    const int copy_threads = host_copy_threads();
    bool any_pageable = false;
    for (int i = 0; i < n_fov; ++i) any_pageable = any_pageable || is_pageable(cubes_host[i]);
    if (any_pageable && (e = ws_stage(w, (size_t)rows * row_bytes))) return e;
    HIPR_CUDA(cudaEventRecord(w.t0, w.copy));
    HIPR_CUDA(cudaStreamWaitEvent(w.comp, w.t0, 0));
    HIPR_CUDA(cudaStreamWaitEvent(w.post, w.t0, 0));
    int b = 0;                                                 // bands since the start of the call (ring slots carry over)
    for (int i = 0; i < n_fov; ++i) {
        const int k = i & 1;
        double *sum_dev = (double *)w.baux[k][0];
        float *score_dev = (float *)w.baux[k][1];
        unsigned long long *key = (unsigned long long *)w.baux[k][2];
        double *score64 = (double *)w.baux[k][3];
        if (i >= 2) HIPR_CUDA(cudaStreamWaitEvent(w.comp, w.post_done[k], 0));   // FOV i - 2 has left this set
        HIPR_CUDA(cudaMemsetAsync(key, 0x00, 8, w.comp));
        HIPR_CUDA(cudaMemsetAsync(key + 1, 0xff, 8, w.comp));
        const bool pageable = is_pageable(cubes_host[i]);
        for (int r0 = 0; r0 < H; r0 += rows, ++b) {
            const int nr = (H - r0 < rows) ? H - r0 : rows;
            const int slot = b % NBUF;
            const void *src = (const char *)cubes_host[i] + (int64_t)r0 * row_bytes;
            if (pageable) {
                if (b >= NBUF) HIPR_CUDA(cudaEventSynchronize(w.staged_out[slot]));
                parallel_copy(w.stage[slot], src, (size_t)nr * row_bytes, copy_threads);
                src = w.stage[slot];
            }
            if (b >= NBUF) HIPR_CUDA(cudaStreamWaitEvent(w.copy, w.freed[slot], 0));
            HIPR_CUDA(cudaMemcpyAsync(w.band[slot], src, (size_t)nr * row_bytes, cudaMemcpyHostToDevice, w.copy));
            if (any_pageable) HIPR_CUDA(cudaEventRecord(w.staged_out[slot], w.copy));   // (the ring's events exist only then)
            HIPR_CUDA(cudaEventRecord(w.copied[slot], w.copy));
            HIPR_CUDA(cudaStreamWaitEvent(w.comp, w.copied[slot], 0));
            if ((e = chansum_band(w.band[slot], 4, 1.f, (int64_t)nr * W, C, sum_dev + (int64_t)r0 * W, key, w.comp))) return e;
            HIPR_CUDA(cudaEventRecord(w.freed[slot], w.comp));
        }
        HIPR_CUDA(cudaEventRecord(w.summed[k], w.comp));
        HIPR_CUDA(cudaStreamWaitEvent(w.post, w.summed[k], 0));
        if (denoise_h > 0.0) {
            double *den = sum_dev + (size_t)H * W;
            if ((e = hipr_normalize(sum_dev, HIPR_F64, (int64_t)H * W, (const uint64_t *)key, w.post))) return e;
            if ((e = hipr_denoise_nl_means_2d(sum_dev, H, W, HIPR_F64, 7, 11, denoise_h, den, w.post))) return e;
            if ((e = hipr_lne2d(den, H, W, W, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour, nullptr, score64, w.post)))
                return e;
            if ((e = hipr_normalize_cast(score64, (int64_t)H * W, nullptr, score_dev, w.post))) return e;
        } else {
            const bool tile_local = (flavour == HIPR_FLAVOUR_F1 || flavour == HIPR_FLAVOUR_F2);
            e = hipr_lne2d_q(sum_dev, H, W, W, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour,
                             tile_local ? nullptr : (const uint64_t *)key, score_dev, w.post);
            if (e == HIPR_E_TABLE && !(patch_size == 11 && n_dirs == 9)) {
                if ((e = hipr_lne2d(sum_dev, H, W, W, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour,
                                    (const uint64_t *)key, score64, w.post)))
                    return e;
                e = hipr_normalize_cast(score64, (int64_t)H * W, nullptr, score_dev, w.post);
            }
            if (e) return e;
        }
        HIPR_CUDA(cudaMemcpyAsync(scores_host[i], score_dev, img_bytes, cudaMemcpyDeviceToHost, w.post));
        HIPR_CUDA(cudaEventRecord(w.post_done[k], w.post));
    }
    HIPR_CUDA(cudaEventRecord(w.t1, w.post));
    HIPR_CUDA(cudaStreamSynchronize(w.post));
    HIPR_CUDA(cudaStreamSynchronize(w.comp));
    HIPR_CUDA(cudaStreamSynchronize(w.copy));
    HIPR_CUDA(cudaEventElapsedTime(&w.last_ms, w.t0, w.t1));
    t_last_ms = w.last_ms;
    guard.dismiss();
    return HIPR_OK;
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

extern "C" double hipr_host_last_elapsed_ms(void) { return (double)t_last_ms; }

extern "C" int hipr_cell_spectra_host(const float *cube_host, const void *labels_host, int label_bytes, int64_t npix,
                                      int64_t row_len, int C, int64_t capacity, int64_t *n_cells, int64_t *labels_out,
                                      int64_t *area_out, double *avgint_out, double *avgint_norm_out) {
    if (!cube_host || !labels_host || !n_cells || !labels_out || !area_out || !avgint_out || !avgint_norm_out ||
        npix < 1 || C < 1 || capacity < 0)
        return HIPR_E_ARG;
    if (label_bytes != 4 && label_bytes != 8) return HIPR_E_DTYPE;
    Workspace *wp = ws_current();
    if (!wp) return HIPR_E_NODEVICE;
    Workspace &w = *wp;
    std::lock_guard<std::mutex> lock(w.mu);
    int e = ws_init(w);
    if (e) return e;
    DrainOnError guard(w);
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
    if (max_label <= 0) {
        guard.dismiss();
        return HIPR_OK;
    }
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
    const int copy_threads = host_copy_threads();
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
    if (n > capacity || n == 0) {
        guard.dismiss();                       // everything enqueued so far has completed
        return n ? HIPR_E_RANGE : HIPR_OK;     // n > capacity: the table stays on the device (hipr_cell_spectra_host_fetch)
    }
    HIPR_CUDA(cudaMemcpyAsync(labels_out, lab_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(area_out, area_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(avgint_out, avg_dev, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaMemcpyAsync(avgint_norm_out, norm_dev, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, w.comp));
    HIPR_CUDA(cudaStreamSynchronize(w.comp));
    guard.dismiss();
    return HIPR_OK;
}

extern "C" int hipr_cell_spectra_host_fetch(int64_t capacity, int64_t *labels_out, int64_t *area_out, double *avgint_out,
                                            double *avgint_norm_out) {
    if (!labels_out || !area_out || !avgint_out || !avgint_norm_out) return HIPR_E_ARG;
    Workspace *wp = ws_current();
    if (!wp) return HIPR_E_NODEVICE;
    Workspace &w = *wp;
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


// ---------------------------------------------------------------------------------------------------------------
// Field-of-view handle: ONE upload of the cube for the score map and the per-cell spectra.
// The scripts hold `image_registered` across both steps (syn/..._measurement.py:161-173: generate_2d_segmentation
// returns it, the regionprops loop reads it again after the watershed); with the two host entry points above the
// cube crosses PCIe twice (29 + 32 ms per 2048^2 FOV).  hipr_fov_upload streams it to the device once, in row bands
// under the channel sum, and keeps cube, sums and range resident; hipr_fov_score and hipr_fov_cell_spectra then cost
// a stencil / a label upload + one pass over the resident cube.
// ---------------------------------------------------------------------------------------------------------------
namespace hipr {
struct Fov {
    int device = 0, H = 0, W = 0, C = 0;
    float *cube = nullptr;
    double *sum = nullptr;
    float *score = nullptr;            // (H, W) float32 scratch: score, then the normalised sum
    unsigned long long *keys = nullptr;
    void *labels = nullptr;            // grown on demand
    size_t labels_bytes = 0;
    void *cells = nullptr;             // accumulators + compacted table
    size_t cells_bytes = 0;
    cudaStream_t copy = nullptr, comp = nullptr;
    cudaEvent_t copied = nullptr, t0 = nullptr, t1 = nullptr;
    float last_ms = -1.f;
    std::mutex mu;
};
struct DeviceScope {                   // run on the handle's device, restore the caller's afterwards
    int prev = -1;
    bool ok = true;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
static void fov_free(Fov *f) {
    if (!f) return;
    if (f->comp) cudaStreamSynchronize(f->comp);
    if (f->copy) cudaStreamSynchronize(f->copy);
    cudaFree(f->cube);
    cudaFree(f->sum);
    cudaFree(f->score);
    cudaFree(f->keys);
    cudaFree(f->labels);
    cudaFree(f->cells);
    if (f->copied) cudaEventDestroy(f->copied);
    if (f->t0) cudaEventDestroy(f->t0);
    if (f->t1) cudaEventDestroy(f->t1);
    if (f->copy) cudaStreamDestroy(f->copy);
    if (f->comp) cudaStreamDestroy(f->comp);
    delete f;
}
}  // namespace hipr

// A released handle's buffers are kept (one per device) for the next upload of the same shape: allocating and freeing
// 1.6 GB of device memory costs more than the upload itself (measured 128 ms per FOV without this, 31 ms with).
static Fov *g_fov_cache[kMaxDevices] = {};
static std::mutex g_fov_cache_mu;

static void fov_cache_release_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return;
    Fov *c = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_fov_cache_mu);
        c = g_fov_cache[dev];
        g_fov_cache[dev] = nullptr;
    }
    fov_free(c);
}

extern "C" int hipr_fov_upload(const float *cube_host, int H, int W, int C, void **handle_out) {
    if (!cube_host || !handle_out || H < 1 || W < 1 || C < 1) return HIPR_E_ARG;
    *handle_out = nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return HIPR_E_NODEVICE;
    Fov *f = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_fov_cache_mu);
        Fov *c = g_fov_cache[dev];
        if (c) {
            g_fov_cache[dev] = nullptr;
            if (c->H == H && c->W == W && c->C == C) f = c;
            else fov_free(c);
        }
    }
    const bool reused = (f != nullptr);
    if (!f) f = new Fov;
    f->device = dev;
    f->H = H; f->W = W; f->C = C;
    const size_t npix = (size_t)H * W;
    int e = HIPR_OK;
    auto fail = [&](int code) {
        fov_free(f);
        return code;
    };
#define HIPR_FOV(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail((int)_e); } while (0)
    if (!reused) {
        HIPR_FOV(cudaMalloc((void **)&f->cube, npix * C * sizeof(float)));
        HIPR_FOV(cudaMalloc((void **)&f->sum, npix * sizeof(double)));
        HIPR_FOV(cudaMalloc((void **)&f->score, npix * sizeof(float)));
        HIPR_FOV(cudaMalloc((void **)&f->keys, 64));
        HIPR_FOV(cudaStreamCreateWithFlags(&f->copy, cudaStreamNonBlocking));
        HIPR_FOV(cudaStreamCreateWithFlags(&f->comp, cudaStreamNonBlocking));
        HIPR_FOV(cudaEventCreateWithFlags(&f->copied, cudaEventDisableTiming));
        HIPR_FOV(cudaEventCreate(&f->t0));
        HIPR_FOV(cudaEventCreate(&f->t1));
    }
    HIPR_FOV(cudaEventRecord(f->t0, f->copy));
    HIPR_FOV(cudaStreamWaitEvent(f->comp, f->t0, 0));
    HIPR_FOV(cudaMemsetAsync(f->keys, 0x00, 8, f->comp));
    HIPR_FOV(cudaMemsetAsync(f->keys + 1, 0xff, 8, f->comp));
    // bands of ~32 MiB straight into the resident cube; the channel sum of band b runs under the copy of band b + 1.
    // Page-locked memory (hipr_host_alloc) goes at the PCIe rate; pageable memory is staged by the driver.
    const int64_t row_bytes = (int64_t)W * C * 4;
    const int rows = band_rows(row_bytes, H);
    for (int r0 = 0; r0 < H; r0 += rows) {
        const int nr = (H - r0 < rows) ? H - r0 : rows;
        float *dst = f->cube + (size_t)r0 * W * C;
        HIPR_FOV(cudaMemcpyAsync(dst, (const char *)cube_host + (int64_t)r0 * row_bytes, (size_t)nr * row_bytes,
                                 cudaMemcpyHostToDevice, f->copy));
        HIPR_FOV(cudaEventRecord(f->copied, f->copy));
        HIPR_FOV(cudaStreamWaitEvent(f->comp, f->copied, 0));
        if ((e = chansum_band(dst, 4, 1.f, (int64_t)nr * W, C, f->sum + (size_t)r0 * W, f->keys, f->comp))) return fail(e);
    }
    HIPR_FOV(cudaEventRecord(f->t1, f->comp));
    HIPR_FOV(cudaStreamSynchronize(f->comp));      // the caller may reuse cube_host
    HIPR_FOV(cudaEventElapsedTime(&f->last_ms, f->t0, f->t1));
#undef HIPR_FOV
    t_last_ms = f->last_ms;
    *handle_out = f;
    return HIPR_OK;
}

extern "C" int hipr_fov_release(void *handle) {
    Fov *f = reinterpret_cast<Fov *>(handle);
    if (!f) return HIPR_OK;
    DeviceScope scope(f->device);
    cudaStreamSynchronize(f->comp);
    cudaStreamSynchronize(f->copy);
    Fov *old = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_fov_cache_mu);
        old = g_fov_cache[f->device];
        g_fov_cache[f->device] = f;          // keep the newest; hipr_host_release_workspace frees it
    }
    fov_free(old);
    return HIPR_OK;
}

// device pointers of a handle's resident arrays (for callers that continue on the device): cube (H, W, C) float32,
// channel sums (H, W) float64, range keys (2 uint64)
extern "C" int hipr_fov_device_arrays(void *handle, const float **cube_dev, const double **sum_dev, const uint64_t **range_dev) {
    Fov *f = reinterpret_cast<Fov *>(handle);
    if (!f) return HIPR_E_ARG;
    if (cube_dev) *cube_dev = f->cube;
    if (sum_dev) *sum_dev = f->sum;
    if (range_dev) *range_dev = reinterpret_cast<const uint64_t *>(f->keys);
    return HIPR_OK;
}

extern "C" int hipr_fov_score(void *handle, int patch_size, int n_dirs, const int32_t *table_host, int flavour,
                              float *score_host, float *sum_host) {
    Fov *f = reinterpret_cast<Fov *>(handle);
    if (!f || !score_host || !table_host) return HIPR_E_ARG;
    std::lock_guard<std::mutex> lock(f->mu);
    DeviceScope scope(f->device);
    if (!scope.ok) return HIPR_E_NODEVICE;
    const int H = f->H, W = f->W;
    const size_t img_bytes = (size_t)H * W * 4;
    struct Drain {
        cudaStream_t s;
        bool armed = true;
        ~Drain() { if (armed) cudaStreamSynchronize(s); }
    } drain{f->comp};
    HIPR_CUDA(cudaEventRecord(f->t0, f->comp));
    const bool tile_local = (flavour == HIPR_FLAVOUR_F1 || flavour == HIPR_FLAVOUR_F2);
    int e = hipr_lne2d_q(f->sum, H, W, W, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour,
                         tile_local ? nullptr : reinterpret_cast<const uint64_t *>(f->keys), f->score, f->comp);
    if (e == HIPR_E_TABLE && !(patch_size == 11 && n_dirs == 9)) {
        // general parameters: float64 stencil into scratch, then cast
        double *score64 = nullptr;
        HIPR_CUDA(cudaMallocAsync((void **)&score64, (size_t)H * W * 8, f->comp));
        e = hipr_lne2d(f->sum, H, W, W, 0, HIPR_F64, patch_size, n_dirs, table_host, flavour,
                       reinterpret_cast<const uint64_t *>(f->keys), score64, f->comp);
        if (!e) e = hipr_normalize_cast(score64, (int64_t)H * W, nullptr, f->score, f->comp);
        cudaFreeAsync(score64, f->comp);
    }
    if (e) return e;
    HIPR_CUDA(cudaMemcpyAsync(score_host, f->score, img_bytes, cudaMemcpyDeviceToHost, f->comp));
    if (sum_host) {
        if ((e = hipr_normalize_cast(f->sum, (int64_t)H * W, reinterpret_cast<const uint64_t *>(f->keys), f->score, f->comp)))
            return e;
        HIPR_CUDA(cudaMemcpyAsync(sum_host, f->score, img_bytes, cudaMemcpyDeviceToHost, f->comp));
    }
    HIPR_CUDA(cudaEventRecord(f->t1, f->comp));
    HIPR_CUDA(cudaStreamSynchronize(f->comp));
    drain.armed = false;
    HIPR_CUDA(cudaEventElapsedTime(&f->last_ms, f->t0, f->t1));
    t_last_ms = f->last_ms;
    return HIPR_OK;
}

extern "C" int hipr_fov_cell_spectra(void *handle, const void *labels_host, int label_bytes, int64_t capacity,
                                     int64_t *n_cells, int64_t *labels_out, int64_t *area_out, double *avgint_out,
                                     double *avgint_norm_out) {
    Fov *f = reinterpret_cast<Fov *>(handle);
    if (!f || !labels_host || !n_cells || !labels_out || !area_out || !avgint_out || !avgint_norm_out || capacity < 0)
        return HIPR_E_ARG;
    if (label_bytes != 4 && label_bytes != 8) return HIPR_E_DTYPE;
    std::lock_guard<std::mutex> lock(f->mu);
    DeviceScope scope(f->device);
    if (!scope.ok) return HIPR_E_NODEVICE;
    const int64_t npix = (int64_t)f->H * f->W;
    const int C = f->C;
    struct Drain {
        cudaStream_t s;
        bool armed = true;
        ~Drain() { if (armed) cudaStreamSynchronize(s); }
    } drain{f->comp};
    if ((size_t)npix * label_bytes > f->labels_bytes) {
        cudaFree(f->labels);
        f->labels = nullptr;
        f->labels_bytes = 0;
        HIPR_CUDA(cudaMalloc(&f->labels, (size_t)npix * label_bytes));
        f->labels_bytes = (size_t)npix * label_bytes;
    }
    HIPR_CUDA(cudaEventRecord(f->t0, f->comp));
    HIPR_CUDA(cudaMemcpyAsync(f->labels, labels_host, (size_t)npix * label_bytes, cudaMemcpyHostToDevice, f->comp));
    int64_t *scalar_dev = reinterpret_cast<int64_t *>(f->keys + 4);      // the handle's 64-byte scratch, past the keys
    int e = hipr_label_max(f->labels, label_bytes, npix, scalar_dev, f->comp);
    if (e) return e;
    int64_t max_label = 0;
    HIPR_CUDA(cudaMemcpyAsync(&max_label, scalar_dev, 8, cudaMemcpyDeviceToHost, f->comp));
    HIPR_CUDA(cudaStreamSynchronize(f->comp));
    *n_cells = 0;
    if (max_label <= 0) {
        drain.armed = false;
        return HIPR_OK;
    }
    const size_t row_bytes = (size_t)C * 8;
    const size_t sums_bytes = (size_t)(max_label + 1) * row_bytes, cnt_bytes = ((size_t)(max_label + 1) * 4 + 15) & ~(size_t)15;
    const size_t fin_off = sums_bytes + cnt_bytes + 16;
    const size_t need = fin_off + 16 + (size_t)max_label * (16 + 2 * row_bytes);
    if (need > f->cells_bytes) {
        cudaFree(f->cells);
        f->cells = nullptr;
        f->cells_bytes = 0;
        HIPR_CUDA(cudaMalloc(&f->cells, need));
        f->cells_bytes = need;
    }
    double *sums = reinterpret_cast<double *>(f->cells);
    int32_t *counts = reinterpret_cast<int32_t *>((char *)f->cells + sums_bytes);
    char *fin = (char *)f->cells + fin_off;
    int32_t *n_dev = reinterpret_cast<int32_t *>(fin);
    int64_t *lab_dev = reinterpret_cast<int64_t *>(fin + 16);
    int64_t *area_dev = lab_dev + max_label;
    double *avg_dev = reinterpret_cast<double *>(area_dev + max_label);
    double *norm_dev = avg_dev + (size_t)max_label * C;
    HIPR_CUDA(cudaMemsetAsync(f->cells, 0, fin_off, f->comp));
    if ((e = hipr_cell_spectra_accumulate(f->cube, f->labels, label_bytes, npix, f->W, C, max_label, sums, counts, nullptr,
                                          f->comp)))
        return e;
    if ((e = hipr_cell_spectra_finalize(sums, counts, max_label, C, n_dev, lab_dev, area_dev, avg_dev, norm_dev, f->comp)))
        return e;
    int32_t n = 0;
    HIPR_CUDA(cudaMemcpyAsync(&n, n_dev, 4, cudaMemcpyDeviceToHost, f->comp));
    HIPR_CUDA(cudaStreamSynchronize(f->comp));
    *n_cells = n;
    if (n > capacity || n == 0) {
        drain.armed = false;
        return n ? HIPR_E_RANGE : HIPR_OK;    // HIPR_E_RANGE: call again with capacity >= *n_cells (the cube is resident)
    }
    HIPR_CUDA(cudaMemcpyAsync(labels_out, lab_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, f->comp));
    HIPR_CUDA(cudaMemcpyAsync(area_out, area_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, f->comp));
    HIPR_CUDA(cudaMemcpyAsync(avgint_out, avg_dev, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, f->comp));
    HIPR_CUDA(cudaMemcpyAsync(avgint_norm_out, norm_dev, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, f->comp));
    HIPR_CUDA(cudaEventRecord(f->t1, f->comp));
    HIPR_CUDA(cudaStreamSynchronize(f->comp));
    drain.armed = false;
    HIPR_CUDA(cudaEventElapsedTime(&f->last_ms, f->t0, f->t1));
    t_last_ms = f->last_ms;
    return HIPR_OK;
}
