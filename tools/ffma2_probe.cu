// Microbenchmark: issue throughput of FFMA vs FFMA2 (packed f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b)
{
    float2 acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {  // scalar: 2 FFMA per pair
                    acc[i].x = fmaf(acc[i].x, a, b);
                    acc[i].y = fmaf(acc[i].y, a, b);
                } else {          // packed: 1 FFMA2 per pair
                    acc[i] = __ffma2_rn(acc[i], A, B);
                }
            }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
int main()
{
    float *d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
            else k<1><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma = 148.0 * 8 * 256 * (double)iters * 4 * 8 * 2;  // scalar FMAs
            if (rep == 2) printf("%s: %.3f ms  %.2f TFMA/s\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9);
        }
    }
    return 0;
}
