// Per-pixel arithmetic shared by the 2-D and 3-D fused stencils: the numpy epilogues of the
// measurement scripts, restated per pixel.
//   F1  syn/..._measurement.py:111-124      F2  bio/..._analysis.py:905-917
//   F3  bio/..._analysis.py:1114-1125       ME2 bio/neighbor.pyx:256-262 + bio/...:812-817
//   V3  bio/neighbor.pyx:335-348
#pragma once
#include "hipr_common.cuh"
#include "sortnet_gen.cuh"

namespace hipr {

template <typename T> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float big() { return 3.402823466e+38f; }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double mn(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double mx(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double big() { return 1.7976931348623157e+308; }
    static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000ll); }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
};

// np.nan_to_num: NaN -> 0, +-inf -> +-largest finite (applied to the samples by F1 and F2).
template <typename T>
__device__ __forceinline__ T nan_to_num(T v) {
    if (v != v) return (T)0;
    return Num<T>::mx(Num<T>::mn(v, Num<T>::big()), -Num<T>::big());
}

// Relative centre value of one line: (centre - min) / range, per flavour.
// mn/mx are the NaN-ignoring min/max of the line; `bad` says a sample was NaN (F3/ME2/V3 only;
// F1/F2 have already mapped NaN to 0), in which case numpy's min/max propagate the NaN.
template <typename T, int FLAVOUR>
__device__ __forceinline__ T line_rel(T centre, T mn, T mx, bool bad) {
    const T range = mx - mn;
    T r;
    if (FLAVOUR == HIPR_FLAVOUR_F1 || FLAVOUR == HIPR_FLAVOUR_F2) {
        r = Num<T>::div(centre - mn, range);  // 0/0 -> NaN on a flat line, as numpy
    } else if (FLAVOUR == HIPR_FLAVOUR_F3) {
        r = Num<T>::div(centre - mn, range + (T)1e-8);
        if (bad) r = Num<T>::nan();
    } else {  // ME2, V3: lp_range = max(lp_range, 1e-8)
        r = Num<T>::div(centre - mn, Num<T>::mx(range, (T)1e-8));
        if (bad) r = Num<T>::nan();
    }
    return r;
}

// Knuth's merge-exchange (Batcher) network, valid for any N.  Fully unrolled so v[] stays in
// registers; only the order statistics read afterwards stay live, the rest is dead code.
template <typename T, int N>
__device__ __forceinline__ void sort_network(T (&v)[N]) {
    if constexpr (N == 72) {
        // nvcc does not unroll the nested merge-exchange loops at this size (v[] would fall to
        // local memory), so the comparator list is generated (gen_sortnet.py) and already pruned
        // to the comparators that reach the quartile order statistics.
#define HIPR_CE(a, b) { const T lo = Num<T>::mn(v[a], v[b]); v[b] = Num<T>::mx(v[a], v[b]); v[a] = lo; }
        HIPR_SORTNET_72(HIPR_CE)
#undef HIPR_CE
    } else {
#pragma unroll
    for (int p = 1; p < N; p <<= 1) {
#pragma unroll
        for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
            for (int j = k % p; j <= N - 1 - k; j += 2 * k) {
#pragma unroll
                for (int i = 0; i <= (k - 1 < N - j - k - 1 ? k - 1 : N - j - k - 1); ++i) {
                    if ((i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        const T a = v[i + j], b = v[i + j + k];
                        v[i + j] = Num<T>::mn(a, b);
                        v[i + j + k] = Num<T>::mx(a, b);
                    }
                }
            }
        }
    }
    }
}

// numpy's percentile 'linear' method at q = num/4 over N sorted values: virtual index
// num*(N-1)/4, lerp as numpy/lib/_function_base_impl.py:_lerp (b - (b-a)*(1-t) when t >= 0.5).
template <typename T, int N, int NUM>
__device__ __forceinline__ T quartile_sorted(const T (&v)[N]) {
    constexpr int k = (NUM * (N - 1)) / 4;
    constexpr int rem = (NUM * (N - 1)) % 4;
    if (rem == 0) return v[k];
    constexpr int k1 = (k + 1 < N) ? k + 1 : k;
    const T a = v[k], b = v[k1];
    const T t = (T)rem * (T)0.25;
    const T d = b - a;
    return (rem >= 2) ? (b - d * ((T)1 - t)) : (a + d * t);
}

// The same lerp from the two order statistics around the virtual index (N = 72: ranks 17 / 18 and 53 / 54).
template <typename T, int NUM>
__device__ __forceinline__ T quartile_pair(T a, T b) {
    constexpr int rem = (NUM * (72 - 1)) % 4;
    const T t = (T)rem * (T)0.25;
    const T d = b - a;
    return (rem >= 2) ? (b - d * ((T)1 - t)) : (a + d * t);
}

// Reduction over the N directions -> one score.
template <typename T, int N, int FLAVOUR>
__device__ __forceinline__ T reduce_dirs(T (&r)[N]) {
    T sum = (T)0;
#pragma unroll
    for (int i = 0; i < N; ++i) sum += r[i];
    const T mean = sum / (T)N;
    sort_network<T, N>(r);
    const T lq = quartile_sorted<T, N, 1>(r);
    const T uq = quartile_sorted<T, N, 3>(r);
    if (FLAVOUR == HIPR_FLAVOUR_F1) {
        T qcv = (T)0;
        if (uq > (T)0) qcv = Num<T>::div(uq - lq, uq + lq + (T)1e-8);
        return mean * ((T)1 - qcv);
    } else if (FLAVOUR == HIPR_FLAVOUR_F2 || FLAVOUR == HIPR_FLAVOUR_ME2) {
        T qcv = Num<T>::div(uq - lq, uq + lq);
        qcv = nan_to_num<T>(qcv);
        return mean * ((T)1 - qcv);
    } else if (FLAVOUR == HIPR_FLAVOUR_F3) {
        const T qcv = Num<T>::div(uq - lq, uq + lq + (T)1e-8);
        return mean * ((T)1 - qcv);
    } else {  // V3: pixel_uq = p25, pixel_lq = p75 (names swapped in the reference)
        return mean * Num<T>::div(lq - uq, lq + uq + (T)1e-8);
    }
}

// Same reduction for a run-time direction count (generic, slow path): order statistics by
// rank counting, O(n^2), values in local memory.
template <typename T>
__device__ inline T kth_smallest(const T *r, int n, int k) {
    for (int i = 0; i < n; ++i) {
        int less = 0, eq = 0;
        for (int j = 0; j < n; ++j) {
            less += (r[j] < r[i]);
            eq += (r[j] == r[i]);
        }
        if (less <= k && k < less + eq) return r[i];
    }
    return r[0];
}
template <typename T>
__device__ inline T quartile_runtime(const T *r, int n, int num) {
    const int k = (num * (n - 1)) / 4;
    const int rem = (num * (n - 1)) % 4;
    const T a = kth_smallest(r, n, k);
    if (rem == 0) return a;
    const T b = kth_smallest(r, n, k + 1 < n ? k + 1 : k);
    const T t = (T)rem * (T)0.25;
    const T d = b - a;
    return (rem >= 2) ? (b - d * ((T)1 - t)) : (a + d * t);
}
template <typename T>
__device__ inline T reduce_dirs_runtime(const T *r, int n, int flavour) {
    T sum = (T)0;
    bool anynan = false;
    for (int i = 0; i < n; ++i) {
        sum += r[i];
        anynan |= (r[i] != r[i]);
    }
    const T mean = sum / (T)n;
    if (anynan) return Num<T>::nan();
    const T lq = quartile_runtime(r, n, 1);
    const T uq = quartile_runtime(r, n, 3);
    if (flavour == HIPR_FLAVOUR_F1) {
        T qcv = (T)0;
        if (uq > (T)0) qcv = Num<T>::div(uq - lq, uq + lq + (T)1e-8);
        return mean * ((T)1 - qcv);
    } else if (flavour == HIPR_FLAVOUR_F2 || flavour == HIPR_FLAVOUR_ME2) {
        T qcv = nan_to_num<T>(Num<T>::div(uq - lq, uq + lq));
        return mean * ((T)1 - qcv);
    } else if (flavour == HIPR_FLAVOUR_F3) {
        return mean * ((T)1 - Num<T>::div(uq - lq, uq + lq + (T)1e-8));
    }
    return mean * Num<T>::div(lq - uq, lq + uq + (T)1e-8);
}

}  // namespace hipr
