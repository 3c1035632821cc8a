// Device microbenchmarks that size the force-kernel design: FP32 issue rates (scalar FFMA vs packed FFMA2),
// MUFU rate, SHFL rate, and L2 reduction (RED) throughput for the j-force scatter patterns.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template<int MODE> __global__ void __launch_bounds__(256) k_pipe(float* out, int iters, float a, float b)
{
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = make_float2(threadIdx.x * 0.001f + i + 1.0f, i * 0.5f + 1.0f);
    float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (MODE == 0) { v[i].x = __fmaf_rn(v[i].x, a, b); v[i].y = __fmaf_rn(v[i].y, a, b); }
            if (MODE == 1) v[i] = __ffma2_rn(v[i], A, B);
            if (MODE == 2) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i].x)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i].y)); }
            if (MODE == 3) { v[i].x = __shfl_xor_sync(0xffffffffu, v[i].x, 1); v[i].y = __shfl_xor_sync(0xffffffffu, v[i].y, 2); }
            if (MODE == 4) { v[i] = __ffma2_rn(v[i], A, B); v[i].x = fmaxf(v[i].x, b); }               // FFMA2 + 1 ALU op
            if (MODE == 5) { v[i] = __ffma2_rn(v[i], A, B); v[i].x = fmaxf(v[i].x, b); v[i].y = fminf(v[i].y, a + 5.f); } // FFMA2 + 2 ALU
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// RED patterns: one "tile" per warp iteration scatters the force of 8 j-atoms (24 floats) to a pseudo-random cluster.
template<int MODE> __global__ void __launch_bounds__(128) k_red(float4* f, int nclusters, int iters)
{
    const int lane = threadIdx.x & 31;
    unsigned  h = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 2654435761u + 12345u;
    for (int it = 0; it < iters; it++)
    {
        h            = h * 1664525u + 1013904223u;
        const int cj = (h >> 8) % nclusters;
        float*    fa = reinterpret_cast<float*>(f + (size_t)cj * 8);
        if (MODE == 0) // current kernel: 16 lanes x/y scalar + 8 lanes z scalar
        {
            const int il = lane & 7, jq = lane >> 3;
            float*    p = fa + 4 * (jq + ((il & 1) ? 4 : 0));
            if (il < 4) atomicAdd(p + ((il & 2) ? 1 : 0), 1.0f);
            if (il < 2) atomicAdd(p + 2, 1.0f);
        }
        if (MODE == 1) // 8 lanes, one float4 each
        {
            if (lane < 8) atomicAdd(f + (size_t)cj * 8 + lane, make_float4(1.f, 1.f, 1.f, 0.f));
        }
        if (MODE == 2) // 8 lanes, float2 + scalar
        {
            if (lane < 8)
            {
                atomicAdd(reinterpret_cast<float2*>(fa + 4 * lane), make_float2(1.f, 1.f));
                atomicAdd(fa + 4 * lane + 2, 1.0f);
            }
        }
        if (MODE == 3) // 24 lanes, scalar, float3-packed layout (12 B per atom)
        {
            if (lane < 24) atomicAdd(reinterpret_cast<float*>(f) + (size_t)cj * 24 + lane, 1.0f);
        }
        if (MODE == 4) // 6 lanes x float4 over a 96-byte float3-packed block
        {
            if (lane < 6) atomicAdd(reinterpret_cast<float4*>(reinterpret_cast<float*>(f) + (size_t)cj * 24) + lane, make_float4(1.f, 1.f, 1.f, 1.f));
        }
    }
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, clockRate %d kHz\n", p.name, nsm, p.clockRate);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    {
        int    nblk = nsm * 8, iters = 20000;
        float* out;
        cudaMalloc(&out, nblk * 256 * sizeof(float));
        const char* names[] = { "FFMA scalar", "FFMA2 packed", "MUFU.RSQ", "SHFL.BFLY", "FFMA2 + 1 ALU", "FFMA2 + 2 ALU" };
        for (int mode = 0; mode < 6; mode++)
            for (int rep = 0; rep < 2; rep++)
            {
                cudaEventRecord(e0);
                switch (mode)
                {
                    case 0: k_pipe<0><<<nblk, 256>>>(out, iters, 0.999f, 0.001f); break;
                    case 1: k_pipe<1><<<nblk, 256>>>(out, iters, 0.999f, 0.001f); break;
                    case 2: k_pipe<2><<<nblk, 256>>>(out, iters / 4, 0.999f, 0.001f); break;
                    case 3: k_pipe<3><<<nblk, 256>>>(out, iters / 4, 0.999f, 0.001f); break;
                    case 4: k_pipe<4><<<nblk, 256>>>(out, iters, 0.999f, 0.001f); break;
                    case 5: k_pipe<5><<<nblk, 256>>>(out, iters, 0.999f, 0.001f); break;
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                int    it  = (mode == 2 || mode == 3) ? iters / 4 : iters;
                double ops = (double)nblk * 256 * it * 16; // lane-ops (2 per packed instruction)
                double per_sm_clk = ops / (ms * 1e-3) / nsm / (p.clockRate * 1e3);
                if (rep) printf("%-16s %.3f ms  %.1f lane-ops/clk/SM (at nominal clock)  %.2f T lane-ops/s\n", names[mode], ms, per_sm_clk, ops / ms * 1e-9);
            }
        cudaFree(out);
    }
    {
        const char* names[] = { "16+8 scalar RED (float4 slots)", "8 x RED.v4", "8 x (RED.v2 + RED)", "24 scalar RED (float3 packed)", "6 x RED.v4 (float3 packed)" };
        for (int ncl = 3168; ncl <= 3168 * 64; ncl *= 8)
        {
            float4* f;
            cudaMalloc(&f, (size_t)ncl * 8 * sizeof(float4));
            cudaMemset(f, 0, (size_t)ncl * 8 * sizeof(float4));
            for (int mode = 0; mode < 5; mode++)
                for (int rep = 0; rep < 2; rep++)
                {
                    int nblk = nsm * 16, iters = 2000;
                    cudaEventRecord(e0);
                    switch (mode)
                    {
                        case 0: k_red<0><<<nblk, 128>>>(f, ncl, iters); break;
                        case 1: k_red<1><<<nblk, 128>>>(f, ncl, iters); break;
                        case 2: k_red<2><<<nblk, 128>>>(f, ncl, iters); break;
                        case 3: k_red<3><<<nblk, 128>>>(f, ncl, iters); break;
                        case 4: k_red<4><<<nblk, 128>>>(f, ncl, iters); break;
                    }
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    double tiles = (double)nblk * 4 * iters;
                    if (rep) printf("clusters %7d  %-32s %.3f ms  %.1f G tiles/s  (%.0f G float-adds/s)\n", ncl, names[mode], ms, tiles / ms * 1e-6, tiles * 24 / ms * 1e-6);
                }
            cudaFree(f);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
