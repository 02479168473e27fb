// 2-D stencils.
//   K2  line_profile_2d  : literal gather, eco/neighbor2d.pyx:56-63          (write-bound)
//   K3  lne2d            : gather + epilogue fused, eco/neighbor2d.pyx:56-63 +
//                          syn/..._measurement.py:111-124 (F1) / bio F2, F3  (smem/ALU-bound)
// The line table arrives from the host (hash-pinned data, never recomputed on the device) and
// travels in the kernel parameter bank, so each sample costs one address add with a constant
// operand plus one shared-memory load.
#include "hipr_common.cuh"
#include "lne_math.cuh"

namespace hipr {

struct Table2D {
    int off[HIPR_MAX_TABLE];  // linear offsets, meaning depends on the kernel
};

// ---------------------------------------------------------------------------------------
// K3 fast path: P = 11, 9 directions.  32x32 output tile per 256-thread CTA, 42x42 input tile
// in shared memory (edge clamp = np.pad(mode='edge') when the source is unpadded), thread
// (tx, ty) scores pixels (tx, ty + 8k).  Lanes of a warp read consecutive words: no conflicts.
// ---------------------------------------------------------------------------------------
constexpr int L2_TW = 32, L2_TH = 32, L2_P = 11, L2_R = 9, L2_HALF = 5;
constexpr int L2_SW = L2_TW + L2_P - 1;  // 42
constexpr int L2_SH = L2_TH + L2_P - 1;  // 42

template <typename T, int FLAVOUR>
__global__ void __launch_bounds__(256)
lne2d_p11r9_kernel(const T *__restrict__ img, int Hs, int Ws, int64_t ld, int src_off, int H, int W,
                   const __grid_constant__ Table2D tab,  // off = dy * L2_SW + dx (patch coords)
                   const unsigned long long *__restrict__ maxkey, T *__restrict__ out) {
    __shared__ T tile[L2_SH * L2_SW];
    const int x0 = blockIdx.x * L2_TW, y0 = blockIdx.y * L2_TH;
    const bool scale = (maxkey != nullptr);
    const T vmax = scale ? (T)double_of_key(*maxkey) : (T)1;
    for (int i = threadIdx.x; i < L2_SH * L2_SW; i += 256) {
        const int ly = i / L2_SW, lx = i - ly * L2_SW;
        // output pixel (y, x) has its patch origin at source (y + src_off - HALF, x + src_off - HALF)
        int sy = y0 + ly - L2_HALF + src_off, sx = x0 + lx - L2_HALF + src_off;
        sy = min(max(sy, 0), Hs - 1);
        sx = min(max(sx, 0), Ws - 1);
        T v = img[(int64_t)sy * ld + sx];
        if (scale) v = Num<T>::div(v, vmax);
        if (FLAVOUR == HIPR_FLAVOUR_F1 || FLAVOUR == HIPR_FLAVOUR_F2) v = nan_to_num<T>(v);
        tile[i] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll 1
    for (int k = 0; k < L2_TH / 8; ++k) {
        const int py = ty + 8 * k;
        const int x = x0 + tx, y = y0 + py;
        if (x >= W || y >= H) continue;
        const T *base = tile + py * L2_SW + tx;
        T r[L2_R];
#pragma unroll
        for (int t = 0; t < L2_R; ++t) {
            T mn = base[tab.off[t * L2_P]], mx = mn, centre = mn;
            bool bad = (mn != mn);
#pragma unroll
            for (int li = 1; li < L2_P; ++li) {
                const T s = base[tab.off[t * L2_P + li]];
                mn = Num<T>::mn(mn, s);
                mx = Num<T>::mx(mx, s);
                if (li == L2_HALF) centre = s;
                if (FLAVOUR != HIPR_FLAVOUR_F1 && FLAVOUR != HIPR_FLAVOUR_F2) bad |= (s != s);
            }
            r[t] = line_rel<T, FLAVOUR>(centre, mn, mx, bad);
        }
        out[(int64_t)y * W + x] = reduce_dirs<T, L2_R, FLAVOUR>(r);
    }
}

// ---------------------------------------------------------------------------------------
// K3 generic path: any odd patch size <= 31 and any direction count <= 128.  One thread per
// pixel straight from global memory (the sum image is L2-resident); correctness path.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
lne2d_generic_kernel(const T *__restrict__ img, int Hs, int Ws, int64_t ld, int src_off, int H, int W,
                     int P, int R, const __grid_constant__ Table2D tab,  // off[(t*P+li)*2 + {0,1}] = dy, dx
                     int flavour, const unsigned long long *__restrict__ maxkey, T *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)H * W) return;
    const int y = (int)(idx / W), x = (int)(idx - (int64_t)y * W);
    const int half = (P - 1) / 2;
    const bool scale = (maxkey != nullptr);
    const T vmax = scale ? (T)double_of_key(*maxkey) : (T)1;
    const bool n2n = (flavour == HIPR_FLAVOUR_F1 || flavour == HIPR_FLAVOUR_F2);
    T r[HIPR_MAX_DIRS];
    for (int t = 0; t < R; ++t) {
        T mn = (T)0, mx = (T)0, centre = (T)0;
        bool bad = false;
        for (int li = 0; li < P; ++li) {
            int sy = y + tab.off[(t * P + li) * 2] - half + src_off;
            int sx = x + tab.off[(t * P + li) * 2 + 1] - half + src_off;
            sy = min(max(sy, 0), Hs - 1);
            sx = min(max(sx, 0), Ws - 1);
            T s = img[(int64_t)sy * ld + sx];
            if (scale) s = Num<T>::div(s, vmax);
            if (n2n) s = nan_to_num<T>(s);
            bad |= (s != s);
            if (li == 0) { mn = s; mx = s; }
            else { mn = Num<T>::mn(mn, s); mx = Num<T>::mx(mx, s); }
            if (li == half) centre = s;
        }
        switch (flavour) {
            case HIPR_FLAVOUR_F1: r[t] = line_rel<T, HIPR_FLAVOUR_F1>(centre, mn, mx, bad); break;
            case HIPR_FLAVOUR_F2: r[t] = line_rel<T, HIPR_FLAVOUR_F2>(centre, mn, mx, bad); break;
            case HIPR_FLAVOUR_F3: r[t] = line_rel<T, HIPR_FLAVOUR_F3>(centre, mn, mx, bad); break;
            default: r[t] = line_rel<T, HIPR_FLAVOUR_ME2>(centre, mn, mx, bad); break;
        }
    }
    out[idx] = reduce_dirs_runtime<T>(r, R, flavour);
}

// ---------------------------------------------------------------------------------------
// K2 literal gather (2-D and 3-D share it): out[pix, k] = src[lin(pix) + off[k]], k < K.
// A CTA owns a run of consecutive pixels of one image row, so its output is one contiguous
// span written with 16-byte stores; sources come through L1/L2 (the padded image is small).
// ---------------------------------------------------------------------------------------
constexpr int GA_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(GA_THREADS)
gather_kernel(const T *__restrict__ src, int64_t stride_a, int64_t stride_b, int inner,
              int64_t nrows /*output rows = a * inner + b*/, int rowlen /*output pixels per row*/,
              int chunk /*pixels per CTA*/, int K, const __grid_constant__ Table2D offs /*K linear offsets*/,
              T *__restrict__ out) {
    // the table travels in the kernel parameter bank (<= 3840 B): no device copy to keep coherent with the launch
    // stream or the current device; staged in shared memory because it is indexed by a runtime k
    extern __shared__ int s_off[];
    for (int i = threadIdx.x; i < K; i += GA_THREADS) s_off[i] = offs.off[i];
    __syncthreads();
    const int chunks_per_row = (rowlen + chunk - 1) / chunk;
    constexpr int VEC = 16 / sizeof(T);
    for (int64_t w = blockIdx.x; w < nrows * chunks_per_row; w += gridDim.x) {
        const int64_t row = w / chunks_per_row;
        const int c0 = (int)(w - row * chunks_per_row) * chunk;
        const int npx = min(chunk, rowlen - c0);
        const int64_t ra = row / inner, rb = row - ra * inner;
        const T *sbase = src + ra * stride_a + rb * stride_b + c0;
        T *obase = out + (row * rowlen + c0) * (int64_t)K;
        const int total = npx * K;
        // peel to a 16-byte boundary of the output, then whole vectors, then the tail
        int head = (int)(((16 - ((uintptr_t)obase & 15)) & 15) / sizeof(T));
        if (head > total) head = total;
        for (int e = threadIdx.x; e < head; e += GA_THREADS) {
            const int p = e / K, k = e - p * K;
            obase[e] = sbase[p + s_off[k]];
        }
        const int nvec = (total - head) / VEC;
        for (int v = threadIdx.x; v < nvec; v += GA_THREADS) {
            const int e0 = head + v * VEC;
            T vals[VEC];
            int p = e0 / K, k = e0 - p * K;
#pragma unroll
            for (int u = 0; u < VEC; ++u) {
                vals[u] = sbase[p + s_off[k]];
                if (++k == K) { k = 0; ++p; }
            }
            *reinterpret_cast<int4 *>(obase + e0) = *reinterpret_cast<const int4 *>(vals);
        }
        for (int e = head + nvec * VEC + threadIdx.x; e < total; e += GA_THREADS) {
            const int p = e / K, k = e - p * K;
            obase[e] = sbase[p + s_off[k]];
        }
    }
}

// 2-D form (K <= 128) at the write roofline: a CTA takes 64 consecutive pixels of a row and builds their whole
// [64][K] output block in shared memory first -- warp w gathers sample k = w, w + 8, ... for 32 consecutive pixels per
// instruction, so a load touches one or two lines (in the kernel above the 32 lanes of a load spread over the K
// samples of two pixels, ~11 image rows: L1-tag bound, 3.8 TB/s) and the store lands at word p * K + k (K odd: no bank
// conflicts) -- then streams the block out, contiguous, with 16-byte stores.
// KT > 0: K known at compile time (99 = the (11, 9) table every 2-D pipeline uses): a thread's sample offsets -- the
// same for every block -- live in registers and its loads are issued in two batches of up to 13 before the stores
// (the run-time-K form keeps 4 in flight and re-reads the offsets from shared memory: 0.69 ms against the 0.44 ms a
// plain fill of the same 3.3 GB takes).
// PX: pixels per block.  64 for the 2-D tables (K <= 128); 8 for the pinned 3-D table (KT = 792: 8 voxels x 792 samples
// are the same 50 KB block, a warp's load covers 4 samples x 8 consecutive voxels).
template <typename T, int KT, int PX>
__global__ void __launch_bounds__(GA_THREADS)
gather_tile_kernel(const T *__restrict__ src, int64_t stride_a, int64_t stride_b, int inner, int64_t nrows, int rowlen,
                   int K_rt, const __grid_constant__ Table2D offs, T *__restrict__ out) {
    constexpr int GT_PX = PX;
    // shared-memory row stride: K, except for the 3-D table, where 796 (a multiple of 4, so rows stay 16-byte aligned)
    // takes the 8 voxels of a store instruction off the same banks (792 = 24 * 33: 8-voxel columns collide four ways)
    constexpr int KP_CT = (KT == 792) ? 796 : KT;
    extern __shared__ __align__(16) unsigned char gt_smem[];
    const int K = KT > 0 ? KT : K_rt;
    const int KP = KT > 0 ? KP_CT : K_rt;
    T *tile = reinterpret_cast<T *>(gt_smem);
    int *s_off = reinterpret_cast<int *>(gt_smem + (size_t)GT_PX * KP * sizeof(T));
    constexpr int KSTEP = GA_THREADS / GT_PX;                       // 4 samples per pass of the CTA
    constexpr int NK = KT > 0 ? (KT + KSTEP - 1) / KSTEP : 1;       // samples per thread
    const int p = threadIdx.x & (GT_PX - 1), k0 = threadIdx.x / GT_PX;
    int my_off[NK];
    if (KT > 0) {
#pragma unroll
        for (int u = 0; u < NK; ++u) my_off[u] = (k0 + u * KSTEP < KT) ? offs.off[k0 + u * KSTEP] : 0;
    } else {
        for (int i = threadIdx.x; i < K; i += GA_THREADS) s_off[i] = offs.off[i];
        __syncthreads();
    }
    const int chunks_per_row = (rowlen + GT_PX - 1) / GT_PX;
    constexpr int VEC = 16 / sizeof(T);
    for (int64_t w = blockIdx.x; w < nrows * chunks_per_row; w += gridDim.x) {
        const int64_t row = w / chunks_per_row;
        const int c0 = (int)(w - row * chunks_per_row) * GT_PX;
        const int npx = min(GT_PX, rowlen - c0);
        const int64_t ra = row / inner, rb = row - ra * inner;
        const T *sbase = src + ra * stride_a + rb * stride_b + c0;
        if (p < npx) {
            T *tp = tile + p * KP + k0;
            const T *sp = sbase + p;
            if constexpr (KT > 0) {
                constexpr int H1 = (NK + 1) / 2;
                T v[H1];
#pragma unroll
                for (int u = 0; u < H1; ++u) v[u] = sp[my_off[u]];
#pragma unroll
                for (int u = 0; u < H1; ++u) tp[u * KSTEP] = v[u];
#pragma unroll
                for (int u = H1; u < NK; ++u) v[u - H1] = sp[my_off[u]];   // the last pass is partial: offset 0 is a valid read
#pragma unroll
                for (int u = H1; u < NK; ++u)
                    if (u < NK - 1 || k0 + u * KSTEP < KT) tp[u * KSTEP] = v[u - H1];
            } else {
                int k = k0;
                for (; k + 3 * KSTEP < K; k += 4 * KSTEP) {
                    const T v0 = sp[s_off[k]], v1 = sp[s_off[k + KSTEP]], v2 = sp[s_off[k + 2 * KSTEP]], v3 = sp[s_off[k + 3 * KSTEP]];
                    tp[k - k0] = v0;
                    tp[k - k0 + KSTEP] = v1;
                    tp[k - k0 + 2 * KSTEP] = v2;
                    tp[k - k0 + 3 * KSTEP] = v3;
                }
                for (; k < K; k += KSTEP) tp[k - k0] = sp[s_off[k]];
            }
        }
        __syncthreads();
        T *obase = out + (row * rowlen + c0) * (int64_t)K;
        const int total = npx * K;
        const int nvec = total / VEC;                      // obase is 16-byte aligned (checked by the launcher)
        const int4 *t4 = reinterpret_cast<const int4 *>(tile);
        int4 *o4 = reinterpret_cast<int4 *>(obase);
        if constexpr (KT > 0 && KP_CT != KT) {
            constexpr int VR = KT / VEC, VRP = KP_CT / VEC;          // vectors per pixel row: output, shared memory
            for (int v = threadIdx.x; v < nvec; v += GA_THREADS) {
                const int pr = v / VR, c = v - pr * VR;
                __stcs(o4 + v, t4[pr * VRP + c]);
            }
        } else {
            for (int v = threadIdx.x; v < nvec; v += GA_THREADS) __stcs(o4 + v, t4[v]);
            for (int e = nvec * VEC + threadIdx.x; e < total; e += GA_THREADS) obase[e] = tile[e];
        }
        __syncthreads();
    }
}

int check_table(const int32_t *table, int n_dirs, int P, int ndim) {
    if (!table) return HIPR_E_ARG;
    if (P < 3 || (P & 1) == 0 || P > HIPR_MAX_PATCH) return HIPR_E_PATCH;
    if (n_dirs < 1 || n_dirs > HIPR_MAX_DIRS || n_dirs * P > HIPR_MAX_TABLE) return HIPR_E_TABLE;
    return HIPR_OK;
    (void)ndim;
}

template <typename T>
static int lne2d_dispatch(const T *img, int Hs, int Ws, int64_t ld, int padded, int P, int R,
                          const int32_t *table, int flavour, const unsigned long long *maxkey, T *out,
                          cudaStream_t st) {
    const int H = padded ? Hs - (P - 1) : Hs;
    const int W = padded ? Ws - (P - 1) : Ws;
    const int src_off = padded ? (P - 1) / 2 : 0;
    if (H < 1 || W < 1) return HIPR_E_PATCH;
    for (int i = 0; i < R * P * 2; ++i)
        if (table[i] < 0 || table[i] >= P) return HIPR_E_TABLE;
    Table2D tab;
    if (P == L2_P && R == L2_R) {
        for (int i = 0; i < R * P; ++i) tab.off[i] = table[2 * i] * L2_SW + table[2 * i + 1];
        dim3 grid((W + L2_TW - 1) / L2_TW, (H + L2_TH - 1) / L2_TH);
        switch (flavour) {
            case HIPR_FLAVOUR_F1:
                lne2d_p11r9_kernel<T, HIPR_FLAVOUR_F1><<<grid, 256, 0, st>>>(img, Hs, Ws, ld, src_off, H, W, tab, maxkey, out);
                break;
            case HIPR_FLAVOUR_F2:
                lne2d_p11r9_kernel<T, HIPR_FLAVOUR_F2><<<grid, 256, 0, st>>>(img, Hs, Ws, ld, src_off, H, W, tab, maxkey, out);
                break;
            case HIPR_FLAVOUR_F3:
                lne2d_p11r9_kernel<T, HIPR_FLAVOUR_F3><<<grid, 256, 0, st>>>(img, Hs, Ws, ld, src_off, H, W, tab, maxkey, out);
                break;
            default:
                return HIPR_E_FLAVOUR;
        }
        return after_launch();
    }
    if (flavour != HIPR_FLAVOUR_F1 && flavour != HIPR_FLAVOUR_F2 && flavour != HIPR_FLAVOUR_F3)
        return HIPR_E_FLAVOUR;
    if (R * P * 2 > HIPR_MAX_TABLE) return HIPR_E_TABLE;
    for (int i = 0; i < R * P * 2; ++i) tab.off[i] = table[i];
    const int64_t n = (int64_t)H * W;
    lne2d_generic_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(img, Hs, Ws, ld, src_off, H, W, P, R,
                                                                        tab, flavour, maxkey, out);
    return after_launch();
}

template <typename T>
int gather_launch(const T *src, int64_t stride_a, int64_t stride_b, int inner, int64_t nrows, int rowlen,
                  int K, const int *lin, T *out, cudaStream_t st) {
    if (K < 1 || K > HIPR_MAX_TABLE) return HIPR_E_TABLE;
    Table2D offs;
    memcpy(offs.off, lin, K * sizeof(int));
    // the shared-memory block kernel, when every block of the output starts on 16 bytes: 64-pixel blocks for the 2-D
    // tables, 8-voxel blocks for the pinned 3-D table
    const int px_blk = (K <= 128) ? 64 : 8;
    // (3-D in float64 stays on the element-order kernel: measured 1.25 ms against 1.31 ms here, the 64-bit stores of 8-voxel
    // blocks conflict four ways; float32: 0.93 -> 0.64 ms per 96 x 128 x 64 volume)
    if ((K <= 128 || K == 792) && (((uintptr_t)out) & 15u) == 0 && ((int64_t)rowlen * K * sizeof(T)) % 16 == 0 &&
        ((int64_t)px_blk * K * sizeof(T)) % 16 == 0) {
        const int kp = (K == 792) ? 796 : K;
        const size_t smem = (size_t)px_blk * kp * sizeof(T) + (size_t)K * sizeof(int);
        static std::atomic<uint64_t> attr_gt{0};
        if (first_use_on_device(attr_gt)) {
            HIPR_CUDA(cudaFuncSetAttribute(gather_tile_kernel<T, 0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 128 * 8 + 128 * 4));
            HIPR_CUDA(cudaFuncSetAttribute(gather_tile_kernel<T, 99, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 128 * 8 + 128 * 4));
            HIPR_CUDA(cudaFuncSetAttribute(gather_tile_kernel<T, 792, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 796 * 8 + 792 * 4));
        }
        auto kern = (K == 99) ? gather_tile_kernel<T, 99, 64> : (K == 792) ? gather_tile_kernel<T, 792, 8> : gather_tile_kernel<T, 0, 64>;
        int per_sm = 1;
        HIPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GA_THREADS, smem));
        const int64_t work = nrows * ((rowlen + px_blk - 1) / px_blk);
        int64_t grid = (int64_t)sm_count() * (per_sm < 1 ? 1 : per_sm);
        if (grid > work) grid = work;
        kern<<<(unsigned)grid, GA_THREADS, smem, st>>>(src, stride_a, stride_b, inner, nrows, rowlen, K, offs, out);
        return after_launch();
    }
    int chunk = 128;
    if (K > 256) chunk = 16;   // 3-D: 792 values per voxel
    const int64_t work = nrows * ((rowlen + chunk - 1) / chunk);
    int64_t grid = (int64_t)sm_count() * 8;
    if (grid > work) grid = work;
    gather_kernel<T><<<(unsigned)grid, GA_THREADS, K * sizeof(int), st>>>(src, stride_a, stride_b, inner, nrows, rowlen,
                                                                         chunk, K, offs, out);
    return after_launch();
}
template int gather_launch<float>(const float *, int64_t, int64_t, int, int64_t, int, int, const int *, float *,
                                  cudaStream_t);
template int gather_launch<double>(const double *, int64_t, int64_t, int, int64_t, int, int, const int *, double *,
                                   cudaStream_t);

}  // namespace hipr

using namespace hipr;

extern "C" int hipr_line_profile_2d(const void *image_padded_dev, int Hp, int Wp, int dtype, int patch_size,
                                    int n_dirs, const int32_t *table_host, void *out_dev, void *stream) {
    if (!image_padded_dev || !out_dev) return HIPR_E_ARG;
    if (dtype != HIPR_F32 && dtype != HIPR_F64) return HIPR_E_DTYPE;
    int e = check_table(table_host, n_dirs, patch_size, 2);
    if (e) return e;
    const int P = patch_size, H = Hp - (P - 1), W = Wp - (P - 1);
    if (H < 1 || W < 1) return HIPR_E_PATCH;
    int lin[HIPR_MAX_TABLE];
    for (int i = 0; i < n_dirs * P; ++i) {
        const int dy = table_host[2 * i], dx = table_host[2 * i + 1];
        if (dy < 0 || dy >= P || dx < 0 || dx >= P) return HIPR_E_TABLE;
        lin[i] = dy * Wp + dx;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HIPR_F32)
        return gather_launch<float>((const float *)image_padded_dev, Wp, 0, 1, H, W, n_dirs * P, lin,
                                    (float *)out_dev, st);
    return gather_launch<double>((const double *)image_padded_dev, Wp, 0, 1, H, W, n_dirs * P, lin,
                                 (double *)out_dev, st);
}

extern "C" int hipr_lne2d(const void *image_dev, int Hs, int Ws, int64_t ld, int padded, int dtype, int patch_size,
                          int n_dirs, const int32_t *table_host, int flavour, const uint64_t *maxkey_dev,
                          void *out_dev, void *stream) {
    if (!image_dev || !out_dev || Hs < 1 || Ws < 1 || ld < Ws) return HIPR_E_ARG;
    if (dtype != HIPR_F32 && dtype != HIPR_F64) return HIPR_E_DTYPE;
    int e = check_table(table_host, n_dirs, patch_size, 2);
    if (e) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long *mk = reinterpret_cast<const unsigned long long *>(maxkey_dev);
    if (dtype == HIPR_F32)
        return lne2d_dispatch<float>((const float *)image_dev, Hs, Ws, ld, padded, patch_size, n_dirs, table_host,
                                     flavour, mk, (float *)out_dev, st);
    return lne2d_dispatch<double>((const double *)image_dev, Hs, Ws, ld, padded, patch_size, n_dirs, table_host,
                                  flavour, mk, (double *)out_dev, st);
}
