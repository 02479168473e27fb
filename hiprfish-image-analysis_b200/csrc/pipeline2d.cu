// hipr_neighbor2d: the cube -> score pipeline as ONE call that overlaps its two kernels.
//
// K1 (channel sum) is HBM-bound with one persistent CTA per SM; K3q (fixed-point stencil) is
// shared-memory/ALU-bound and small (7 KB, 256 threads per CTA), so its CTAs fit beside K1's on
// the same SMs.  The image is cut into row bands: K1 runs band after band on the caller's stream
// while the stencil of the previous band runs on a side stream -- only the last band's stencil is
// exposed.  The stencil can start before the whole sum image exists because F1/F2 are invariant
// to any affine map of the image: each tile is quantised with its own min/max (lne2d_q.cu), no
// global max is needed.  F3 needs the global max for its epsilon, so it runs the two kernels back
// to back.  The caller's stream sees one ordered operation (fork/join with events).
#include <mutex>
#include "hipr_common.cuh"

namespace hipr {

int chansum_band(const void *cube, int sample_bytes, float scale, int64_t npix, int C, double *out,
                 unsigned long long *maxkey, cudaStream_t st);
int lne2d_q_rows(const void *image_dev, int Hs, int Ws, int64_t ld, int padded, int dtype, const int32_t *table_host,
                 int flavour, const uint64_t *range_dev, float *out_dev, int y_begin, int y_end, cudaStream_t st);

constexpr int PL_MAX_BANDS = 16;
constexpr int PL_MAX_DEVICES = 16;
struct SideStream {
    cudaStream_t s = nullptr;
    cudaEvent_t start = nullptr, done = nullptr, k1[PL_MAX_BANDS] = {};
    bool ready = false;
};
static SideStream g_side[PL_MAX_DEVICES];
static std::mutex g_side_mu;

static int side_for_current_device(SideStream **out) {
    int dev = 0;
    HIPR_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= PL_MAX_DEVICES) return HIPR_E_RANGE;
    SideStream &sd = g_side[dev];
    if (!sd.ready) {
        HIPR_CUDA(cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking));
        HIPR_CUDA(cudaEventCreateWithFlags(&sd.start, cudaEventDisableTiming));
        HIPR_CUDA(cudaEventCreateWithFlags(&sd.done, cudaEventDisableTiming));
        for (int i = 0; i < PL_MAX_BANDS; ++i) HIPR_CUDA(cudaEventCreateWithFlags(&sd.k1[i], cudaEventDisableTiming));
        sd.ready = true;
    }
    *out = &sd;
    return HIPR_OK;
}

}  // namespace hipr

using namespace hipr;

extern "C" int hipr_neighbor2d(const float *cube_dev, int H, int W, int C, int patch_size, int n_dirs,
                               const int32_t *table_host, int flavour, float *score_dev, double *sum_dev,
                               uint64_t *range_dev, int bands, void *stream) {
    if (!cube_dev || !score_dev || !sum_dev || !range_dev || !table_host || H < 1 || W < 1 || C < 1) return HIPR_E_ARG;
    if (patch_size != 11 || n_dirs != 9) return HIPR_E_UNSUPPORTED;
    if (flavour != HIPR_FLAVOUR_F1 && flavour != HIPR_FLAVOUR_F2 && flavour != HIPR_FLAVOUR_F3) return HIPR_E_FLAVOUR;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *rg = reinterpret_cast<unsigned long long *>(range_dev);
    HIPR_CUDA(cudaMemsetAsync(range_dev, 0x00, 8, st));
    HIPR_CUDA(cudaMemsetAsync(range_dev + 1, 0xff, 8, st));
    const bool local = (flavour != HIPR_FLAVOUR_F3);
    if (bands <= 0) bands = 1;   // measured: banding does not pay; overlap FOVs over two streams instead
    if (bands > PL_MAX_BANDS) bands = PL_MAX_BANDS;
    // bands of a multiple of 32 rows (keeps every band 16-byte aligned and tile-aligned)
    int band_rows = (((H + bands - 1) / bands) + 31) / 32 * 32;
    if (!local || H < 128) band_rows = H;
    const int nb = (H + band_rows - 1) / band_rows;
    int e;
    if (nb == 1) {
        if ((e = chansum_band(cube_dev, 4, 1.f, (int64_t)H * W, C, sum_dev, rg, st))) return e;
        return lne2d_q_rows(sum_dev, H, W, W, 0, HIPR_F64, table_host, flavour, local ? nullptr : range_dev, score_dev, 0,
                            H, st);
    }
    std::lock_guard<std::mutex> lock(g_side_mu);
    SideStream *sd = nullptr;
    if ((e = side_for_current_device(&sd))) return e;
    HIPR_CUDA(cudaEventRecord(sd->start, st));
    HIPR_CUDA(cudaStreamWaitEvent(sd->s, sd->start, 0));
    for (int b = 0; b < nb; ++b) {
        const int r0 = b * band_rows, r1 = (r0 + band_rows < H) ? r0 + band_rows : H;
        if ((e = chansum_band(cube_dev + (int64_t)r0 * W * C, 4, 1.f, (int64_t)(r1 - r0) * W, C, sum_dev + (int64_t)r0 * W, rg, st)))
            return e;
        HIPR_CUDA(cudaEventRecord(sd->k1[b], st));
        HIPR_CUDA(cudaStreamWaitEvent(sd->s, sd->k1[b], 0));
        if (b >= 1) {
            // the stencil of band b-1 reads 5 rows into band b, which now exists
            if ((e = lne2d_q_rows(sum_dev, H, W, W, 0, HIPR_F64, table_host, flavour, nullptr, score_dev,
                                  (b - 1) * band_rows, r0, sd->s)))
                return e;
        }
    }
    if ((e = lne2d_q_rows(sum_dev, H, W, W, 0, HIPR_F64, table_host, flavour, nullptr, score_dev, (nb - 1) * band_rows, H,
                          sd->s)))
        return e;
    HIPR_CUDA(cudaEventRecord(sd->done, sd->s));
    HIPR_CUDA(cudaStreamWaitEvent(st, sd->done, 0));
    return HIPR_OK;
}
