// micro-benchmark: pure-read bandwidth, LDG.128 vs 1-D bulk copy (TMA) with different depths
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void ldg_read(const float4* __restrict__ p, size_t n, float* out) {
    float acc = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p + i));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + i + stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(p + i + 2 * stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(p + i + 3 * stride));
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
    }
    for (; i < n; i += stride) { float4 a = p[i]; acc += a.x + a.y + a.z + a.w; }
    if (acc == 123.456f) *out = acc;
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// bulk: each CTA loops over chunks; STAGES in flight; consumers just touch one word per chunk
template <int STAGES>
__global__ void bulk_read(const char* __restrict__ p, size_t nchunks, uint32_t chunk_bytes, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = (uint64_t*)sm;
    unsigned char* buf = sm + 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    float acc = 0.f;
    size_t it = 0;
    size_t first = blockIdx.x;
    // prime
    size_t issued = 0;
    for (size_t c = first; c < nchunks && issued < STAGES; c += gridDim.x, ++issued) {
        int s = issued % STAGES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + (size_t)s * chunk_bytes)), "l"(p + c * chunk_bytes), "r"(chunk_bytes), "r"(s32(&full[s])) : "memory");
    }
    size_t next = first + issued * gridDim.x;
    for (size_t c = first; c < nchunks; c += gridDim.x, ++it) {
        int s = it % STAGES;
        uint32_t parity = (it / STAGES) & 1;
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(&full[s])), "r"(parity) : "memory");
        acc += *(volatile float*)(buf + (size_t)s * chunk_bytes);
        if (next < nchunks) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + (size_t)s * chunk_bytes)), "l"(p + next * chunk_bytes), "r"(chunk_bytes), "r"(s32(&full[s])) : "memory");
            next += gridDim.x;
        }
    }
    if (acc == 123.456f) *out = acc;
}
template <typename F> float timeit(F f, int n = 20) {
    for (int i = 0; i < 3; ++i) f();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); for (int i = 0; i < n; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / n;
}
int main() {
    size_t bytes = (size_t)2048 * 2048 * 380;   // the cube
    char* d; cudaMalloc(&d, bytes); cudaMemset(d, 0, bytes);
    float* out; cudaMalloc(&out, 4);
    for (int bpsm : {2, 4, 8, 16}) {
        float ms = timeit([&] { ldg_read<<<148 * bpsm, 512>>>((const float4*)d, bytes / 16, out); });
        printf("LDG.128  %2d CTAs/SM x512: %.4f ms  %.0f GB/s\n", bpsm, ms, bytes / ms / 1e6);
    }
    uint32_t chunk = 48640;
    size_t nchunks = bytes / chunk;
    cudaFuncSetAttribute(bulk_read<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(bulk_read<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(bulk_read<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    { float ms = timeit([&] { bulk_read<4><<<148, 64, 128 + 4 * chunk>>>(d, nchunks, chunk, out); }); printf("bulk 48640 x4 stages, 1 CTA/SM: %.4f ms %.0f GB/s\n", ms, bytes / ms / 1e6); }
    { float ms = timeit([&] { bulk_read<2><<<148 * 2, 64, 128 + 2 * chunk>>>(d, nchunks, chunk, out); }); printf("bulk 48640 x2 stages, 2 CTA/SM: %.4f ms %.0f GB/s\n", ms, bytes / ms / 1e6); }
    chunk = 24320; nchunks = bytes / chunk;
    { float ms = timeit([&] { bulk_read<8><<<148, 64, 128 + 8 * chunk>>>(d, nchunks, chunk, out); }); printf("bulk 24320 x8 stages, 1 CTA/SM: %.4f ms %.0f GB/s\n", ms, bytes / ms / 1e6); }
    { float ms = timeit([&] { bulk_read<4><<<148 * 2, 64, 128 + 4 * chunk>>>(d, nchunks, chunk, out); }); printf("bulk 24320 x4 stages, 2 CTA/SM: %.4f ms %.0f GB/s\n", ms, bytes / ms / 1e6); }
    chunk = 12160; nchunks = bytes / chunk;
    { float ms = timeit([&] { bulk_read<8><<<148 * 2, 64, 128 + 8 * chunk>>>(d, nchunks, chunk, out); }); printf("bulk 12160 x8 stages, 2 CTA/SM: %.4f ms %.0f GB/s\n", ms, bytes / ms / 1e6); }
    cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
    return 0;
}
