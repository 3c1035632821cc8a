// Measures the FP32 FMA roofline of the device with scalar FFMA and packed FFMA2 (fma.rn.f32x2) chains.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_peak fp32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template<int PACKED> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b)
{
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (PACKED) v[i] = __ffma2_rn(v[i], A, B);
            else { v[i].x = __fmaf_rn(v[i].x, a, b); v[i].y = __fmaf_rn(v[i].y, a, b); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nblk = p.multiProcessorCount * 8, iters = 20000;
    float* out; cudaMalloc(&out, nblk * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int packed = 0; packed < 2; packed++)
        for (int rep = 0; rep < 3; rep++)
        {
            cudaEventRecord(e0);
            if (packed) k<1><<<nblk, 256>>>(out, iters, 0.999f, 0.001f); else k<0><<<nblk, 256>>>(out, iters, 0.999f, 0.001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flop = (double)nblk * 256 * iters * 16 * 2;
            printf("%s rep %d: %.3f ms  %.2f TFLOP/s (SMs %d, clockRate %d kHz)\n", packed ? "FFMA2" : "FFMA ", rep, ms, flop / ms * 1e-9, p.multiProcessorCount, p.clockRate);
        }
    return 0;
}
