// K1 variants: CPX pixels per stage, STAGES, GROUPS consumer groups of CPX threads, HINT, MODE (0 = full sum, 1 = touch only)
#include <cstdio>
#include "hipr_common.cuh"
namespace hipr { std::atomic<int64_t> g_launches{0}; }
using namespace hipr;
template <int CPX, int STAGES, int GROUPS, bool HINT, int MODE>
__global__ void __launch_bounds__(GROUPS * CPX + 32, 1)
k1(const float* __restrict__ cube, int64_t nchunks, int C, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_bytes = CPX * C * 4u;
    uint64_t* full = (uint64_t*)smem_raw; uint64_t* empty = full + 8;
    float* ring = (float*)(smem_raw + 128);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CPX / 32); } fence_barrier_init(); }
    __syncthreads();
    if (warp == GROUPS * (CPX / 32)) {
        if ((tid & 31) == 0) {
            const uint64_t pol = policy_evict_first();
            int s = 0; uint32_t round = 0;
            for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                mbar_expect_tx(&full[s], stage_bytes);
                if (HINT) bulk_g2s((unsigned char*)ring + (size_t)s * stage_bytes, cube + chunk * (int64_t)CPX * C, stage_bytes, &full[s], pol);
                else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32((unsigned char*)ring + (size_t)s * stage_bytes)), "l"(cube + chunk * (int64_t)CPX * C), "r"(stage_bytes), "r"(smem_u32(&full[s])) : "memory");
                if (++s == STAGES) { s = 0; ++round; }
            }
        }
        return;
    }
    const int g = warp / (CPX / 32), t = tid - g * CPX;
    int64_t it = 0;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
        if ((int)(it % GROUPS) != g) continue;
        const int s = (int)(it % STAGES);
        mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
        const float* px = ring + (size_t)s * (stage_bytes >> 2) + (size_t)t * C;
        double sum;
        if (MODE == 0) sum = sum_channels<false>(px, C);
        else if (MODE == 2) sum = sum_channels<true>(px, C);
        else sum = (double)px[0];
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        out[chunk * CPX + t] = sum;
    }
}
template <typename F> float timeit(F f, int n = 20) {
    for (int i = 0; i < 3; ++i) f();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); for (int i = 0; i < n; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / n;
}
template <int CPX, int STAGES, int GROUPS, bool HINT, int MODE>
void run(const float* d, double* out, const char* name) {
    const int C = 95; const int64_t npix = 2048 * 2048; const int64_t nchunks = npix / CPX;
    size_t smem = 128 + (size_t)STAGES * CPX * C * 4;
    printf("[%s] ", name); fflush(stdout);
    auto k = k1<CPX, STAGES, GROUPS, HINT, MODE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    float ms = 1e9;
    for (int rep = 0; rep < 1; ++rep) { float m = timeit([&] { k<<<148, GROUPS * CPX + 32, smem>>>(d, nchunks, C, out); }); ms = m < ms ? m : ms; }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-44s %.4f ms  %.0f GB/s  %s\n", name, ms, npix * 380.0 / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e)); fflush(stdout);
}
int main() {
    size_t bytes = (size_t)2048 * 2048 * 380;
    float* d; cudaMalloc(&d, bytes); cudaMemset(d, 0, bytes);
    double* out; cudaMalloc(&out, 2048 * 2048 * 8);
    run<128, 3, 2, false, 0>(d, out, "128 x3 g2");
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
