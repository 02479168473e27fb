// K6-3D: non-local-means denoise of the (X, Y, Z) channel-sum volume -- the step between the channel sum and the
// 72-direction stencil in the z-stack caller: skimage.restoration.denoise_nl_means(volume, h = 0.03),
// bio/..._analysis.py:454, fast mode, patch_size 7, patch_distance 11, sigma 0 (scikit-image >= 0.15 semantics: a
// 3-D array is a volume; INTEGRATION.md section 4 for the <= 0.14 ambiguity).
//
// scikit-image builds one 3-D integral image of squared differences per patch shift and accumulates each pair of
// voxels symmetrically.  Voxel by voxel that is (the test side's denoise_nl_means_3d_direct, equal to its loop-for-loop
// restatement to 1e-15):
//     out[p] = sum_t w(p, t) v[p + t] / sum_t w(p, t),      t in [-d, d]^3,  twice the weight for t = 0,
//     w(p, t) = exp(-dist) if dist <= 5 else 0,
//     dist    = max(sum_{u in W(p)} (v[u] - v[u + t])^2, 0) / (h^2 s^3),   W(p) = p - 2 .. p + 3 per axis (6^3),
// on the reflect-padded volume.  (2d + 1)^3 = 12,167 shifts x 216 window voxels: a volume costs ~12,167 / 529 x 6 =
// 140 times a 2-D image of the same voxel count, which is why the separable running sums below matter.
//
// A pad kernel first writes the reflect-padded volume (pad offset + d + 1 = 15, as skimage; dimensions rounded up to
// whole tiles) so that every shifted read is one linear offset.  A CTA (16 warps) owns an 8 x 11 x 27 output tile: its
// unshifted 13 x 16 x 32 window region sits in shared memory (z along the 32 lanes: 27 + 5 = one warp width).  Per shift:
//   phase 1  warp j, lane k: for region row j, the 13 squared differences along x (shifted samples straight from
//            global memory / L1: coalesced rows of 32 doubles) and their eight sliding 6-sums -> Sx[x][j][k];
//   phase 2  warp = (plane x, half of the rows), lane k: eleven Sx[x][.][k], six sliding 6-sums along y in registers,
//            then the 6-sum along z across lanes (shuffles), exp, accumulate weight and weighted value for six voxels
//            per thread (first version: 8 warps, two rows / eleven voxels per thread, 154 registers: 102 ms per
//            128 x 132 x 54 volume; 16 warps: 86 ms; 32 warps with the planes of phase 1 split in halves: 100 ms -- the
//            redundant differences cost more than the extra warps hide).
// Sx is double-buffered (phase 1 of shift i + 1 precedes phase 2 of shift i): one __syncthreads per shift.
// Everything is float64, for the reason given in nlm2d.cu (the hard cutoff).
#include "hipr_common.cuh"
#include "nlm_common.cuh"

namespace hipr {

constexpr int N3_TX = 8, N3_TY = 11, N3_TZ = 27;
constexpr int N3_OFF = 3, N3_N = 2 * N3_OFF;                 // patch 7: window side 6
constexpr int N3_RX = N3_TX + N3_N - 1;                      // 13 region planes
constexpr int N3_RY = N3_TY + N3_N - 1;                      // 16 region rows
constexpr int N3_RZ = N3_TZ + N3_N - 1;                      // 32 region columns = lanes
constexpr int N3_THREADS = 512;                              // 16 warps: one region row each in phase 1, (plane, half of the rows) in phase 2
constexpr int N3_NQ = 6;                                     // voxels per thread in phase 2 (rows 0..5 / 6..10)
static_assert(N3_RZ == 32 && N3_RY == 16 && N3_TX == 8 && N3_THREADS == 32 * N3_RY && 2 * N3_NQ >= N3_TY, "the thread mapping below assumes these");

// reflect-padded copy (np.pad(mode='reflect') by `pad`), dimensions (Xp, Yp, Zp) >= (X, Y, Z) + 2 pad; cells beyond
// the padded volume (tile round-up) are zero and never reach a written voxel
template <typename T>
__global__ void nlm3d_pad_kernel(const T *__restrict__ vol, int X, int Y, int Z, int pad, int Xp, int Yp, int Zp,
                                 double *__restrict__ out) {
    const int64_t n = (int64_t)Xp * Yp * Zp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int zp = (int)(i % Zp);
        const int64_t r = i / Zp;
        const int yp = (int)(r % Yp), xp = (int)(r / Yp);
        double v = 0.0;
        if (xp < X + 2 * pad && yp < Y + 2 * pad && zp < Z + 2 * pad)
            v = (double)vol[((int64_t)reflect_index(xp - pad, X) * Y + reflect_index(yp - pad, Y)) * Z + reflect_index(zp - pad, Z)];
        out[i] = v;
    }
}

template <typename T>
__global__ void __launch_bounds__(N3_THREADS, 1)
nlm3d_kernel(const double *__restrict__ vp, int Yp, int Zp, int pad, int X, int Y, int Z, int d, double inv_h2s3,
             T *__restrict__ out) {
    extern __shared__ __align__(16) double n3_smem[];
    double *A = n3_smem;                                       // [13][16][32]
    double *Sx = n3_smem + N3_RX * N3_RY * N3_RZ;              // [2][8][16][32]
    constexpr int SXN = N3_TX * N3_RY * N3_RZ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nzb = (Z + N3_TZ - 1) / N3_TZ, nyb = (Y + N3_TY - 1) / N3_TY;
    const int z0 = (blockIdx.x % nzb) * N3_TZ, y0 = ((blockIdx.x / nzb) % nyb) * N3_TY, x0 = (blockIdx.x / (nzb * nyb)) * N3_TX;
    const int plane = Yp * Zp;                 // 32-bit strides and shift offsets (the launcher checks 12 * plane < 2^31): one IMAD.WIDE per address
    // region origin in the padded volume: tile origin - offset + 1
    const double *rbase = vp + ((int64_t)(x0 + pad - N3_OFF + 1) * Yp + (y0 + pad - N3_OFF + 1)) * Zp + (z0 + pad - N3_OFF + 1);
    for (int i = tid; i < N3_RX * N3_RY * N3_RZ; i += N3_THREADS) {
        const int k = i & 31, j = (i >> 5) & 15, ii = i >> 9;
        A[i] = rbase[(int64_t)ii * plane + (int64_t)j * Zp + k];
    }
    __syncthreads();
    // phase 1: row j = warp of the region, column k = lane
    const double *a1 = A + warp * N3_RZ + lane;
    const double *b1 = rbase + (int64_t)warp * Zp + lane;
    auto phase1 = [&](int soff, int which) {
        double *dst = Sx + which * SXN + warp * N3_RZ + lane;
        const double *b = b1 + soff;
        double D[N3_RX];
#pragma unroll
        for (int i = 0; i < N3_RX; ++i) {
            const double df = a1[i * N3_RY * N3_RZ] - __ldg(b + i * plane);
            D[i] = df * df;
        }
        double s = D[0];
#pragma unroll
        for (int i = 1; i < N3_N; ++i) s += D[i];
        dst[0] = s;
#pragma unroll
        for (int x = 1; x < N3_TX; ++x) {
            s += D[x + N3_N - 1] - D[x - 1];               // sliding window along x
            dst[x * N3_RY * N3_RZ] = s;
        }
    };
    // phase 2: plane px = warp & 7 of the tile, rows yb .. yb + 5 (yb = 0 or 6; row 11 does not exist), column lane
    const int px = warp & 7, yb = (warp >> 3) * N3_NQ;
    double acc_w[N3_NQ], acc_v[N3_NQ];
#pragma unroll
    for (int q = 0; q < N3_NQ; ++q) acc_w[q] = acc_v[q] = 0.0;
    // centre of voxel (x0 + px, y0 + yb, z0 + lane) in the padded volume
    const double *vc = vp + ((int64_t)(x0 + px + pad) * Yp + (y0 + yb + pad)) * Zp + (z0 + pad) + lane;
    const bool lane_on = lane < N3_TZ;
    const int side = 2 * d + 1;
    const int nshift = side * side * side;
    // shifts in (tx, ty, tz) raster order, tz fastest; the linear offset of the shifted samples advances incrementally
    // (an integer division by the run-time side length per shift cost ~80 instructions per thread and shift)
    int soff = -d * plane - d * Zp - d;
    int cy = 0, cz = 0;                                          // ty + d, tz + d of `soff`
    auto advance = [&](int &off, int &ky, int &kz) {
        ++off;
        if (++kz == side) {
            kz = 0;
            off += Zp - side;
            if (++ky == side) {
                ky = 0;
                off += plane - side * Zp;
            }
        }
    };
    phase1(soff, 0);
    __syncthreads();
    int buf = 0;
    for (int s = 0; s < nshift; ++s) {
        int soff_next = soff;
        int ny = cy, nz = cz;
        advance(soff_next, ny, nz);
        if (s + 1 < nshift) phase1(soff_next, buf ^ 1);
        const double *hcur = Sx + buf * SXN + px * N3_RY * N3_RZ + lane;
        double hs[N3_NQ + N3_N - 1];
#pragma unroll
        for (int j = 0; j < N3_NQ + N3_N - 1; ++j) hs[j] = hcur[min(yb + j, N3_RY - 1) * N3_RZ];   // row 16 only feeds the unused row 11
        double sy = hs[0];
#pragma unroll
        for (int j = 1; j < N3_N; ++j) sy += hs[j];
#pragma unroll
        for (int q = 0; q < N3_NQ; ++q) {
            if (q > 0) sy += hs[q + N3_N - 1] - hs[q - 1];     // sliding window along y
            // 6-sum along z: lanes l .. l + 5 (lanes >= 27 produce unused values)
            // 6 = 4 + 2: pairs, quads, then quad(l) + pair(l + 4) -- three shuffles instead of five
            const double s2 = sy + __shfl_down_sync(0xffffffffu, sy, 1);
            const double s4 = s2 + __shfl_down_sync(0xffffffffu, s2, 2);
            const double box = s4 + __shfl_down_sync(0xffffffffu, s2, 4);
            const double dist = fabs(box) * inv_h2s3;
            const int hi = __double2hiint(dist);
            const bool inside = (hi < 0x40140000) || (hi == 0x40140000 && __double2loint(dist) == 0);   // dist <= 5.0
            const bool on = lane_on && yb + q < N3_TY;
            double w = exp_small_neg(-dist);
            w = (inside && on) ? w : 0.0;
            const double v = on ? __ldg(vc + (soff + q * Zp)) : 0.0;
            acc_w[q] += w;
            acc_v[q] = fma(w, v, acc_v[q]);
        }
        buf ^= 1;
        soff = soff_next;
        cy = ny;
        cz = nz;
        __syncthreads();
    }
    const int x = x0 + px, z = z0 + lane;
    if (lane_on && x < X && z < Z) {
#pragma unroll
        for (int q = 0; q < N3_NQ; ++q) {
            const int y = y0 + yb + q;
            if (yb + q >= N3_TY || y >= Y) break;
            // the zero shift counts twice
            const double w = acc_w[q] + 1.0, v = acc_v[q] + vc[(int64_t)q * Zp];
            out[((int64_t)x * Y + y) * Z + z] = (T)(v / w);
        }
    }
}

}  // namespace hipr

using namespace hipr;

extern "C" int hipr_denoise_nl_means_3d(const void *volume_dev, int X, int Y, int Z, int dtype, int patch_size,
                                        int patch_distance, double h, void *out_dev, void *workspace_dev,
                                        int64_t workspace_bytes, void *stream) {
    if (X < 1 || Y < 1 || Z < 1 || !(h > 0.0)) return HIPR_E_ARG;
    if (dtype != HIPR_F32 && dtype != HIPR_F64) return HIPR_E_DTYPE;
    if (patch_size != 7 && patch_size != 6) return HIPR_E_UNSUPPORTED;   // skimage makes an even size odd
    if (patch_distance < 0 || patch_distance > 15) return HIPR_E_UNSUPPORTED;
    const int d = patch_distance, pad = N3_OFF + d + 1;
    if (X <= pad || Y <= pad || Z <= pad) return HIPR_E_PATCH;           // single reflection only
    const int nxb = (X + N3_TX - 1) / N3_TX, nyb = (Y + N3_TY - 1) / N3_TY, nzb = (Z + N3_TZ - 1) / N3_TZ;
    const int Xp = nxb * N3_TX + 2 * pad, Yp = nyb * N3_TY + 2 * pad, Zp = nzb * N3_TZ + 2 * pad;
    const int64_t need = (int64_t)Xp * Yp * Zp * (int64_t)sizeof(double);
    if (!volume_dev || !out_dev || !workspace_dev) return HIPR_E_ARG;
    if (workspace_bytes < need) return HIPR_E_RANGE;
    if ((int64_t)nxb * nyb * nzb > 0x7fffffffLL || 12ll * Yp * Zp > 0x7fffffffLL) return HIPR_E_RANGE;
    cudaStream_t st = (cudaStream_t)stream;
    double *vp = (double *)workspace_dev;
    const int64_t np_ = (int64_t)Xp * Yp * Zp;
    int64_t pblocks = (np_ + 255) / 256;
    if (pblocks > (int64_t)sm_count() * 16) pblocks = (int64_t)sm_count() * 16;
    if (dtype == HIPR_F32)
        nlm3d_pad_kernel<float><<<(unsigned)pblocks, 256, 0, st>>>((const float *)volume_dev, X, Y, Z, pad, Xp, Yp, Zp, vp);
    else
        nlm3d_pad_kernel<double><<<(unsigned)pblocks, 256, 0, st>>>((const double *)volume_dev, X, Y, Z, pad, Xp, Yp, Zp, vp);
    int e = after_launch();
    if (e) return e;
    const size_t smem = ((size_t)N3_RX * N3_RY * N3_RZ + 2 * (size_t)N3_TX * N3_RY * N3_RZ) * sizeof(double);
    static std::atomic<uint64_t> attr{0};
    if (first_use_on_device(attr)) {
        HIPR_CUDA(cudaFuncSetAttribute(nlm3d_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HIPR_CUDA(cudaFuncSetAttribute(nlm3d_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const double inv = 1.0 / (h * h * 343.0);
    const unsigned grid = (unsigned)((int64_t)nxb * nyb * nzb);
    if (dtype == HIPR_F32)
        nlm3d_kernel<float><<<grid, N3_THREADS, smem, st>>>(vp, Yp, Zp, pad, X, Y, Z, d, inv, (float *)out_dev);
    else
        nlm3d_kernel<double><<<grid, N3_THREADS, smem, st>>>(vp, Yp, Zp, pad, X, Y, Z, d, inv, (double *)out_dev);
    return after_launch();
}

// bytes of the workspace hipr_denoise_nl_means_3d needs for an (X, Y, Z) volume (the reflect-padded float64 copy)
extern "C" int64_t hipr_denoise_nl_means_3d_workspace(int X, int Y, int Z, int patch_distance) {
    if (X < 1 || Y < 1 || Z < 1 || patch_distance < 0 || patch_distance > 15) return -1;
    const int pad = N3_OFF + patch_distance + 1;
    const int64_t Xp = (int64_t)((X + N3_TX - 1) / N3_TX) * N3_TX + 2 * pad, Yp = (int64_t)((Y + N3_TY - 1) / N3_TY) * N3_TY + 2 * pad,
                  Zp = (int64_t)((Z + N3_TZ - 1) / N3_TZ) * N3_TZ + 2 * pad;
    return Xp * Yp * Zp * (int64_t)sizeof(double);
}
