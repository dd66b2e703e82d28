#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, const double* in, int n)
{
    double s = in[0];
    double t[8];
    for (int i = 0; i < 8; ++i) t[i] = in[1 + i];
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) s = __dsub_rn(s, t[u]);
    }
    long long c1 = clock64();
    // smem-fed chain
    __shared__ double sh[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = in[i & 7];
    __syncthreads();
    double s2 = in[0];
    long long c2 = clock64();
    for (int r = 0; r < n / 4; ++r) {
        const double2* p = reinterpret_cast<const double2*>(sh);
#pragma unroll 16
        for (int u = 0; u < 16; ++u) { double2 v = p[(u + r) & 127]; s2 = __dsub_rn(s2, v.x); s2 = __dsub_rn(s2, v.y); }
    }
    long long c3 = clock64();
    // float chain for comparison
    float f = (float)in[0];
    long long c4 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) f = __fsub_rn(f, (float)t[u]);
    }
    long long c5 = clock64();
    // two interleaved double chains
    double a = in[0], b = in[1];
    long long c6 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { a = __dsub_rn(a, t[u]); b = __dsub_rn(b, t[u]); }
    }
    long long c7 = clock64();
    if (threadIdx.x == 0) {
        out[0] = s + s2 + f + a + b;
        cyc[0] = c1 - c0; cyc[1] = c3 - c2; cyc[2] = c5 - c4; cyc[3] = c7 - c6;
    }
}
int main()
{
    double h[256]; for (int i = 0; i < 256; ++i) h[i] = 1e-3 * (i + 1);
    double *in, *out; long long* cyc;
    cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, 8); cudaMalloc(&cyc, 64);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int n = 1024;
    for (int threads : {1, 32, 128}) {
        k<<<1, threads>>>(out, cyc, in, n);
        cudaDeviceSynchronize();
        long long c[4]; cudaMemcpy(c, cyc, 32, cudaMemcpyDeviceToHost);
        printf("threads %d: DADD reg chain %.2f cyc/op; smem-fed chain %.2f cyc/op; FADD chain %.2f; two DADD chains %.2f cyc per pair\n", threads,
               (double)c[0] / (8.0 * n), (double)c[1] / (32.0 * (n / 4)), (double)c[2] / (8.0 * n), (double)c[3] / (8.0 * n));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
