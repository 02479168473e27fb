// Split-mosaic exchange over NVLink peer memory (BASELINE config 5), without a collective library call.
//
// One stitched mosaic is cut into row slabs, one per GPU.  The stencil runs on the 1-channel sum image, so a
// slab needs 5 rows of that image from the slab above and below, plus the global max / min of the sums.
// Every rank owns ONE peer-mapped buffer (cudaMalloc + cudaIpc handle, opened by all the other ranks):
//
//   ext[parity]  (rows_max + 10, W) float64   rows [0, 5) top halo, [5, 5 + rows) own sums, then bottom halo
//   keys[parity] (world, 2) uint64            max / min keys of every rank's slab
//   flags        (world) uint64               flags[r] = last epoch whose data from rank r has landed here
//
// The channel-sum kernel writes the slab's sums straight into ext[parity] rows [5, 5 + rows).  Then
//   push: one CTA copies the slab's first / last 5 rows into the neighbours' halo rows and its keys into every
//         rank's key table with plain stores through the peer mappings (NVLink), fences system-wide and
//         releases flags[rank] = epoch on every peer;
//   wait: one CTA spins (bounded) until flags[r] >= epoch for every r, then reduces the key table to the global
//         range the stencil consumes.
// Two parities: a rank can be at most one exchange ahead of its neighbours (its next push follows its own
// wait, which needs the neighbours' current flags, which they release only after their previous stencil -- same
// stream -- has read the other parity).
#include <mutex>
#include "hipr_common.cuh"

namespace hipr {

int chansum_band(const void *cube, int sample_bytes, float scale, int64_t npix, int C, double *out,
                 unsigned long long *maxkey, cudaStream_t st);
int lne2d_q_rows(const void *image_dev, int Hs, int Ws, int64_t ld, int padded, int dtype, const int32_t *table_host,
                 int flavour, const uint64_t *range_dev, float *out_dev, int y_begin, int y_end, cudaStream_t st);

constexpr int MP_HALO = 5;
constexpr int MP_MAX_BANDS = 16;
constexpr int MP_MAX_DEVICES = 16;

struct MosaicSide {
    cudaStream_t s = nullptr;
    cudaEvent_t start = nullptr, pushed = nullptr, done = nullptr, k1[MP_MAX_BANDS] = {};
    bool ready = false;
};
static MosaicSide g_mside[MP_MAX_DEVICES];
static std::mutex g_mside_mu;

static int mosaic_side(MosaicSide **out) {
    int dev = 0;
    HIPR_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MP_MAX_DEVICES) return HIPR_E_RANGE;
    MosaicSide &sd = g_mside[dev];
    if (!sd.ready) {
        HIPR_CUDA(cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking));
        HIPR_CUDA(cudaEventCreateWithFlags(&sd.start, cudaEventDisableTiming));
        HIPR_CUDA(cudaEventCreateWithFlags(&sd.pushed, cudaEventDisableTiming));
        HIPR_CUDA(cudaEventCreateWithFlags(&sd.done, cudaEventDisableTiming));
        for (int i = 0; i < MP_MAX_BANDS; ++i) HIPR_CUDA(cudaEventCreateWithFlags(&sd.k1[i], cudaEventDisableTiming));
        sd.ready = true;
    }
    *out = &sd;
    return HIPR_OK;
}
constexpr int MP_MAX_WORLD = 64;

// how long a wait kernel spins for its peers before it gives up (SM clocks; hipr_mosaic_p2p_set_timeout_ms)
static std::atomic<long long> g_wait_timeout_clocks{20ll * 1000 * 1000 * 1000};   // ~10 s at 1.97 GHz

struct MosaicLayout {
    int64_t ext_elems;     // doubles per parity
    int64_t keys_off;      // byte offsets from the base
    int64_t flags_off;
    int64_t bytes;
};
__host__ __device__ inline MosaicLayout mosaic_layout(int rows_max, int W, int world) {
    MosaicLayout l;
    l.ext_elems = (int64_t)(rows_max + 2 * MP_HALO) * W;
    l.keys_off = 2 * l.ext_elems * 8;
    l.flags_off = l.keys_off + (int64_t)2 * world * 2 * 8;
    l.bytes = l.flags_off + (int64_t)world * 8;
    l.bytes = (l.bytes + 255) / 256 * 256;
    return l;
}

struct PeerBases {
    unsigned char *base[MP_MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024)
mosaic_push_kernel(PeerBases peers, int rank, int world, int rows, int rows_up, int rows_max, int W, int parity,
                   const unsigned long long *__restrict__ keys_local, unsigned long long epoch) {
    const MosaicLayout l = mosaic_layout(rows_max, W, world);
    const double *own = reinterpret_cast<const double *>(peers.base[rank]) + parity * l.ext_elems;
    const int64_t n = (int64_t)MP_HALO * W;
    if (rank > 0) {          // my first rows -> bottom halo of the slab above
        double *dst = reinterpret_cast<double *>(peers.base[rank - 1]) + parity * l.ext_elems + (int64_t)(MP_HALO + rows_up) * W;
        const double *src = own + (int64_t)MP_HALO * W;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
    if (rank + 1 < world) {  // my last rows -> top halo of the slab below
        double *dst = reinterpret_cast<double *>(peers.base[rank + 1]) + parity * l.ext_elems;
        const double *src = own + (int64_t)rows * W;          // ext row MP_HALO + rows - MP_HALO
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
    if ((int)threadIdx.x < world) {
        unsigned long long *k = reinterpret_cast<unsigned long long *>(peers.base[threadIdx.x] + l.keys_off) +
                                ((int64_t)parity * world + rank) * 2;
        k[0] = keys_local[0];
        k[1] = keys_local[1];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world)
        st_release_sys(reinterpret_cast<unsigned long long *>(peers.base[threadIdx.x] + l.flags_off) + rank, epoch);
}

__global__ void __launch_bounds__(MP_MAX_WORLD)
mosaic_wait_kernel(unsigned char *base, int world, int rows_max, int W, int parity, unsigned long long epoch,
                   long long timeout_clocks, unsigned long long *__restrict__ range_out, int *__restrict__ error) {
    const MosaicLayout l = mosaic_layout(rows_max, W, world);
    const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(base + l.flags_off);
    __shared__ int failed;
    if (threadIdx.x == 0) failed = 0;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
            if (clock64() - t0 > timeout_clocks) {      // a peer died: report, do not hang the GPU
                atomicExch(&failed, 1);
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long *k = reinterpret_cast<const unsigned long long *>(base + l.keys_off) + (int64_t)parity * world * 2;
        unsigned long long kmax = 0ull, kmin = ~0ull;
        for (int r = 0; r < world; ++r) {
            kmax = k[2 * r] > kmax ? k[2 * r] : kmax;
            kmin = k[2 * r + 1] < kmin ? k[2 * r + 1] : kmin;
        }
        range_out[0] = kmax;
        range_out[1] = kmin;
        if (failed) *error = 1;
    }
}

// A wait kernel that timed out set *error: the stencil that followed read stale halo rows (and, for F3, a stale
// range), so the score of this call is poisoned with NaN.  The failure is then visible in the data itself, without a
// host synchronisation, even if the caller never polls the flag.  One CTA reads the flag; nothing else happens when
// it is clear.
__global__ void __launch_bounds__(256)
mosaic_guard_kernel(const int *__restrict__ error, float *__restrict__ score, int64_t n) {
    if (*error == 0) return;
    const float nan = __int_as_float(0x7fc00000);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) score[i] = nan;
}

}  // namespace hipr

using namespace hipr;

extern "C" int hipr_mosaic_p2p_set_timeout_ms(double ms) {
    if (!(ms > 0.0) || ms > 3.6e6) return HIPR_E_ARG;
    int dev = 0, khz = 0;
    HIPR_CUDA(cudaGetDevice(&dev));
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess || khz <= 0) khz = 1965000;
    g_wait_timeout_clocks.store((long long)(ms * (double)khz));
    return HIPR_OK;
}

extern "C" int hipr_mosaic_p2p_guard(const int32_t *error_dev, float *score_dev, int64_t n, void *stream) {
    if (!error_dev || !score_dev || n < 1) return HIPR_E_ARG;
    mosaic_guard_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(error_dev, score_dev, n);
    return after_launch();
}

extern "C" int64_t hipr_mosaic_p2p_bytes(int rows_max, int W, int world) {
    if (rows_max < MP_HALO || W < 1 || world < 1 || world > MP_MAX_WORLD) return HIPR_E_ARG;
    return mosaic_layout(rows_max, W, world).bytes;
}

extern "C" int hipr_p2p_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return HIPR_E_ARG;
    HIPR_CUDA(cudaMalloc(ptr, (size_t)bytes));
    HIPR_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    return HIPR_OK;
}
extern "C" int hipr_p2p_free(void *ptr) {
    if (ptr) HIPR_CUDA(cudaFree(ptr));
    return HIPR_OK;
}
extern "C" int hipr_p2p_get_handle(void *ptr, void *handle64) {
    if (!ptr || !handle64) return HIPR_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    HIPR_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
    return HIPR_OK;
}
extern "C" int hipr_p2p_open_handle(const void *handle64, void **peer_ptr) {
    if (!handle64 || !peer_ptr) return HIPR_E_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    HIPR_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return HIPR_OK;
}
extern "C" int hipr_p2p_close_handle(void *peer_ptr) {
    if (peer_ptr) HIPR_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return HIPR_OK;
}

extern "C" int hipr_mosaic_p2p_rows_ptr(void *base, int rows_max, int W, int world, int parity, int with_top_halo,
                                        double **rows_ptr) {
    if (!base || !rows_ptr || (parity != 0 && parity != 1)) return HIPR_E_ARG;
    const MosaicLayout l = mosaic_layout(rows_max, W, world);
    *rows_ptr = reinterpret_cast<double *>(base) + parity * l.ext_elems + (with_top_halo ? 0 : (int64_t)MP_HALO * W);
    return HIPR_OK;
}

extern "C" int hipr_mosaic_p2p_exchange(void *const *bases_host, int rank, int world, int rows, int rows_up,
                                        int rows_max, int W, int parity, const uint64_t *keys_local_dev, uint64_t epoch,
                                        uint64_t *range_out_dev, int32_t *error_dev, void *stream) {
    if (!bases_host || !keys_local_dev || !range_out_dev || !error_dev || world < 1 || world > MP_MAX_WORLD || rank < 0 ||
        rank >= world || rows < MP_HALO || rows > rows_max || W < 1 || (parity != 0 && parity != 1))
        return HIPR_E_ARG;
    PeerBases pb;
    memset(&pb, 0, sizeof(pb));
    for (int r = 0; r < world; ++r) {
        if (!bases_host[r]) return HIPR_E_ARG;
        pb.base[r] = reinterpret_cast<unsigned char *>(bases_host[r]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    mosaic_push_kernel<<<1, 1024, 0, st>>>(pb, rank, world, rows, rows_up, rows_max, W, parity,
                                           reinterpret_cast<const unsigned long long *>(keys_local_dev),
                                           (unsigned long long)epoch);
    int e = after_launch();
    if (e) return e;
    const long long timeout = g_wait_timeout_clocks.load();
    mosaic_wait_kernel<<<1, MP_MAX_WORLD, 0, st>>>(pb.base[rank], world, rows_max, W, parity, (unsigned long long)epoch,
                                                   timeout, reinterpret_cast<unsigned long long *>(range_out_dev), error_dev);
    return after_launch();
}


// One slab of a split mosaic, cube -> score, with the exchange AND the stencil hidden under the channel sum:
// the slab is cut into row bands; the first and the last band are summed first and their edge rows pushed to
// the neighbours; then band b + 1 is summed (HBM-bound, caller's stream) while the stencil of band b (SM-bound,
// tile-local quantisation) runs on a side stream; the two edge bands' stencils follow the wait kernel.  F3
// needs the global range for its epsilon and runs unbanded (sum, exchange, stencil in order).
extern "C" int hipr_mosaic_p2p_score(const float *cube_slab_dev, int C, void *const *bases_host, int rank, int world,
                                     int rows, int rows_up, int rows_max, int W, int parity, uint64_t epoch,
                                     const int32_t *table_host, int flavour, int bands, uint64_t *keys_local_dev,
                                     uint64_t *range_dev, int32_t *error_dev, float *score_dev, void *stream) {
    if (!cube_slab_dev || !bases_host || !table_host || !keys_local_dev || !range_dev || !error_dev || !score_dev ||
        C < 1 || world < 1 || world > MP_MAX_WORLD || rank < 0 || rank >= world || rows < MP_HALO || rows > rows_max ||
        W < 1 || (parity != 0 && parity != 1))
        return HIPR_E_ARG;
    if (flavour != HIPR_FLAVOUR_F1 && flavour != HIPR_FLAVOUR_F2 && flavour != HIPR_FLAVOUR_F3) return HIPR_E_FLAVOUR;
    cudaStream_t st = (cudaStream_t)stream;
    const MosaicLayout l = mosaic_layout(rows_max, W, world);
    double *ext = reinterpret_cast<double *>(bases_host[rank]) + parity * l.ext_elems;
    double *own = ext + (int64_t)MP_HALO * W;
    const int n_top = rank > 0 ? MP_HALO : 0, n_bottom = rank + 1 < world ? MP_HALO : 0;
    const double *img = n_top ? ext : own;                 // the extended image the stencil reads
    const int Hs = rows + n_top + n_bottom;
    float *out_ext = score_dev - (int64_t)n_top * W;       // extended-image row n_top is score row 0
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(keys_local_dev);
    HIPR_CUDA(cudaMemsetAsync(keys_local_dev, 0x00, 8, st));
    HIPR_CUDA(cudaMemsetAsync(keys_local_dev + 1, 0xff, 8, st));
    if (bands > MP_MAX_BANDS) bands = MP_MAX_BANDS;
    int band_rows = bands > 1 ? (((rows + bands - 1) / bands) + 31) / 32 * 32 : rows;
    if (flavour == HIPR_FLAVOUR_F3 || band_rows < 64) band_rows = rows;
    const int nb = (rows + band_rows - 1) / band_rows;
    int e;
    if (nb < 3) {
        if ((e = chansum_band(cube_slab_dev, 4, 1.f, (int64_t)rows * W, C, own, keys, st))) return e;
        if ((e = hipr_mosaic_p2p_exchange(bases_host, rank, world, rows, rows_up, rows_max, W, parity, keys_local_dev, epoch,
                                          range_dev, error_dev, stream)))
            return e;
        if ((e = lne2d_q_rows(img, Hs, W, W, 0, HIPR_F64, table_host, flavour, range_dev, out_ext, n_top, n_top + rows, st)))
            return e;
        return hipr_mosaic_p2p_guard(error_dev, score_dev, (int64_t)rows * W, stream);
    }
    std::lock_guard<std::mutex> lock(g_mside_mu);
    MosaicSide *sd = nullptr;
    if ((e = mosaic_side(&sd))) return e;
    auto band = [&](int b, int &r0, int &r1) {
        r0 = b * band_rows;
        r1 = (r0 + band_rows < rows) ? r0 + band_rows : rows;
    };
    auto k1 = [&](int b) -> int {
        int r0, r1;
        band(b, r0, r1);
        return chansum_band(cube_slab_dev + (int64_t)r0 * W * C, 4, 1.f, (int64_t)(r1 - r0) * W, C, own + (int64_t)r0 * W, keys, st);
    };
    auto stencil = [&](int b) -> int {
        int r0, r1;
        band(b, r0, r1);
        return lne2d_q_rows(img, Hs, W, W, 0, HIPR_F64, table_host, flavour, nullptr, out_ext, n_top + r0, n_top + r1, sd->s);
    };
    HIPR_CUDA(cudaEventRecord(sd->start, st));
    HIPR_CUDA(cudaStreamWaitEvent(sd->s, sd->start, 0));
    if ((e = k1(0))) return e;
    if ((e = k1(nb - 1))) return e;
    // push the edge rows (the keys are partial here and unused: F1 / F2 quantise tile by tile)
    PeerBases pb;
    memset(&pb, 0, sizeof(pb));
    for (int r = 0; r < world; ++r) {
        if (!bases_host[r]) return HIPR_E_ARG;
        pb.base[r] = reinterpret_cast<unsigned char *>(bases_host[r]);
    }
    mosaic_push_kernel<<<1, 1024, 0, st>>>(pb, rank, world, rows, rows_up, rows_max, W, parity, keys, (unsigned long long)epoch);
    if ((e = after_launch())) return e;
    HIPR_CUDA(cudaEventRecord(sd->pushed, st));
    for (int b = 1; b <= nb - 2; ++b) {
        if ((e = k1(b))) return e;
        HIPR_CUDA(cudaEventRecord(sd->k1[b], st));
        HIPR_CUDA(cudaStreamWaitEvent(sd->s, sd->k1[b], 0));
        if (b >= 2 && (e = stencil(b - 1))) return e;      // bands b - 2, b - 1, b exist
    }
    if (nb >= 3 && (e = stencil(nb - 2))) return e;         // its lower neighbour, the last band, was summed first
    // the two edge bands need the neighbours' rows
    HIPR_CUDA(cudaStreamWaitEvent(sd->s, sd->pushed, 0));
    mosaic_wait_kernel<<<1, MP_MAX_WORLD, 0, sd->s>>>(pb.base[rank], world, rows_max, W, parity, (unsigned long long)epoch,
                                                      g_wait_timeout_clocks.load(),
                                                      reinterpret_cast<unsigned long long *>(range_dev), error_dev);
    if ((e = after_launch())) return e;
    if ((e = stencil(0))) return e;
    if ((e = stencil(nb - 1))) return e;
    HIPR_CUDA(cudaEventRecord(sd->done, sd->s));
    HIPR_CUDA(cudaStreamWaitEvent(st, sd->done, 0));
    return hipr_mosaic_p2p_guard(error_dev, score_dev, (int64_t)rows * W, stream);
}
