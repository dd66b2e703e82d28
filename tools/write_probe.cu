// Write-only bandwidth of one B200 by store mechanism (development probe; built into tools/_dev/write_probe):
// which of them can carry the mostly-zero dX / dL tiles of the categorical step and the fp32 configurations of the k-best kernel.
//   memset            cudaMemsetAsync
//   stg128 / stg256   persistent grid, st.global.L1::no_allocate.v4 / .v8 (256-bit, sm_100+), a warp writes 512 B / 1 KB per instruction
//   bulk<bytes,nbuf>  one warp = nbuf shared-memory buffers of `bytes`, cp.async.bulk shared -> global (TMA without a tensor map),
//                     the warp re-touches a buffer (one 128-bit store per lane) before it stores it again, as the step kernel does
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/write_probe.cu -o tools/_dev/write_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void stg128_kernel(float4* p, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p + i), "f"(0.f) : "memory");
}
__global__ void stg256_kernel(float4* p, size_t n32)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + 2 * i), "f"(0.f) : "memory");
}
template <int NBUF>
__global__ void bulk_kernel(char* p, size_t bytes_total, int chunk)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned char* mine = smem + (size_t)warp * NBUF * chunk;
    for (int i = lane * 16; i < NBUF * chunk; i += 512) *reinterpret_cast<float4*>(mine + i) = make_float4(0, 0, 0, 0);
    const size_t nchunks = bytes_total / chunk;
    int b = 0;
    for (size_t t = (size_t)warp * gridDim.x + blockIdx.x; t < nchunks; t += (size_t)gridDim.x * nwarps) {
        // at most NBUF - 1 older stores may still be reading their buffers
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
        __syncwarp();
        unsigned char* buf = mine + (size_t)b * chunk;
        *reinterpret_cast<float4*>(buf + lane * 16) = make_float4((float)t, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + t * chunk),
                         "r"((uint32_t)__cvta_generic_to_shared(buf)), "r"((uint32_t)chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        b = (b + 1) % NBUF;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <class F> static float time_ms(F f, int reps = 20)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main()
{
    const size_t bytes = (size_t)1 << 30;
    char* p = nullptr;
    CK(cudaMalloc(&p, bytes));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto rep = [&](const char* name, float ms) { printf("%-28s %.4f ms  %.0f GB/s\n", name, ms, bytes / ms / 1e6); fflush(stdout); };
    rep("memset", time_ms([&] { cudaMemsetAsync(p, 0, bytes); }));
    for (int cta : {1, 2, 4, 8}) {
        char nm[64];
        snprintf(nm, 64, "stg128 %dx256 per SM", cta);
        rep(nm, time_ms([&] { stg128_kernel<<<sms * cta, 256>>>((float4*)p, bytes / 16); }));
        snprintf(nm, 64, "stg256 %dx256 per SM", cta);
        rep(nm, time_ms([&] { stg256_kernel<<<sms * cta, 256>>>((float4*)p, bytes / 32); }));
    }
    CK(cudaFuncSetAttribute(bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int chunk : {16384, 8192, 4096, 2048}) {
        for (int warps : {4, 8, 12}) {
            char nm[64];
            if ((size_t)warps * chunk <= 196608) {
                snprintf(nm, 64, "bulk %5d B x1, %2d warps", chunk, warps);
                rep(nm, time_ms([&] { bulk_kernel<1><<<sms, warps * 32, (size_t)warps * chunk>>>(p, bytes, chunk); }));
            }
            if ((size_t)warps * chunk * 2 <= 196608) {
                snprintf(nm, 64, "bulk %5d B x2, %2d warps", chunk, warps);
                rep(nm, time_ms([&] { bulk_kernel<2><<<sms, warps * 32, (size_t)warps * chunk * 2>>>(p, bytes, chunk); }));
            }
            if ((size_t)warps * chunk * 4 <= 196608) {
                snprintf(nm, 64, "bulk %5d B x4, %2d warps", chunk, warps);
                rep(nm, time_ms([&] { bulk_kernel<4><<<sms, warps * 32, (size_t)warps * chunk * 4>>>(p, bytes, chunk); }));
            }
        }
    }
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return 0;
}
