/* pairmath_bench: the arithmetic ceiling of the force kernel's plain step, measured without any memory traffic.
 * Includes force.cu itself and runs pair_fscal<Ewald, geometric, F> exactly as k_force's plain step does (2 i-pairs per lane and
 * step, i / j accumulators, the half-warp exchange), with j-atoms synthesised in registers, at a chosen number of resident warps
 * per SM.  Prints cycles per step per SM sub-partition: compare with the kernel's measured figure (kernel_sweep.py).
 * build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude -Igmxapi_b200/csrc profiles/tools/pairmath_bench.cu -o scratch/pairmath_bench */
#include "../../gmxapi_b200/csrc/force.cu"
int nb_fail(b200nb_context*, int code, const std::string&) { return code; } /* the one host symbol force.cu needs from b200nb.cu */

namespace
{
template<int EEL, int MODE>
__global__ void __launch_bounds__(32, 16) k_bench(int nsteps, const float* __restrict__ kconst, NbParamsDev P, float* out)
{
    const int lane = threadIdx.x & 31;
    KConst K;
    {
        const float4 k0 = __ldg(reinterpret_cast<const float4*>(kconst)), k1 = __ldg(reinterpret_cast<const float4*>(kconst) + 1),
                     k2 = __ldg(reinterpret_cast<const float4*>(kconst) + 2);
        K.rc2 = k0.x, K.beta2 = k0.z, K.fd4 = k0.w, K.fd3 = k1.x, K.fn6 = k1.y, K.fn5 = k1.z, K.fd2 = k1.w, K.fd1 = k2.x, K.fd0 = k2.y;
    }
    IData I[2];
    for (int p = 0; p < 2; p++)
    {
        I[p].x = make_float2(0.1f * lane + p, 0.2f * lane + p);
        I[p].y = make_float2(0.3f + p, 0.15f * lane);
        I[p].z = make_float2(0.05f * lane, 0.4f + p);
        I[p].q = make_float2(0.4f, -0.8f);
        I[p].c6n = make_float2(-0.1f, -0.12f);
        I[p].c12 = make_float2(0.01f, 0.012f);
        I[p].t0 = I[p].t1 = 0;
        I[p].g0 = I[p].g1 = dup(0.f);
    }
    float2 fix[2] = { dup(0.f), dup(0.f) }, fiy[2] = { dup(0.f), dup(0.f) }, fiz[2] = { dup(0.f), dup(0.f) };
    float  acc = 0.f, ev = 0.f, ec = 0.f;
    JAtom  J;
    J.xq = make_float4(0.3f + 0.01f * lane, 0.2f, 0.1f, 0.5f);
    J.lj = make_float2(0.3f, 0.1f);
    J.slot = lane;
    for (int s = 0; s < nsteps; s++)
    {
        J.xq.x += 0.001f; /* a new j-atom every step, no memory */
        J.xq.y -= 0.0007f;
        float2 fjx = dup(0.f), fjy = dup(0.f), fjz = dup(0.f);
#pragma unroll
        for (int p = 0; p < 2; p++)
        {
            float2       dx, dy, dz;
            const float2 fs = pair_fscal<EEL, true, false, false, false>(I[p], J, P, K, nullptr, nullptr, 1.f, 1.f, true, true, dx, dy, dz, ev, ec);
            fix[p] = fma2(fs, dx, fix[p]);
            fiy[p] = fma2(fs, dy, fiy[p]);
            fiz[p] = fma2(fs, dz, fiz[p]);
            fjx    = fma2(fs, dx, fjx);
            fjy    = fma2(fs, dy, fjy);
            fjz    = fma2(fs, dz, fjz);
        }
        if (MODE == 1)
        {
            /* the exchange of reduce_store_j without the red */
            const bool  upper = lane >= 16;
            const float sx = -fjx.x - fjx.y, sy = -fjy.x - fjy.y, sz = -fjz.x - fjz.y;
            const float rcv = __shfl_xor_sync(0xffffffffu, upper ? sx : sz, 16);
            const float rcy = __shfl_xor_sync(0xffffffffu, sy, 16);
            acc += (upper ? sz : sx) + rcv + (upper ? 0.0f : sy + rcy);
        }
        else
            acc += fjx.x + fjx.y + fjy.x + fjy.y + fjz.x + fjz.y;
    }
    out[blockIdx.x * 32 + lane] = acc + fix[0].x + fix[1].y + fiy[0].x + fiy[1].y + fiz[0].y + fiz[1].x;
}
} // namespace

int main()
{
    float  kc[12] = { 0.81f, 3.47f, 12.04f, 0.0011193462567257629232f / 3.47f, 0.014866955030185295499f / 3.47f, -1.7357322914161492954e-8f,
                     1.4703624142580877519e-6f, 0.11583842382862377919f / 3.47f, 0.50736591960530292870f / 3.47f, 1.0f / 3.47f, 0, 0 };
    float *d_k, *d_o;
    cudaMalloc(&d_k, sizeof(kc));
    cudaMemcpy(d_k, kc, sizeof(kc), cudaMemcpyHostToDevice);
    cudaMalloc(&d_o, 148 * 64 * 32 * 4);
    NbParamsDev P{};
    P.two_k_rf = 0.6f;
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    const double ghz = pr.clockRate * 1e-6;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int nsteps = 20000;
    for (int mode = 0; mode < 2; mode++)
        for (int eel = 1; eel >= 0; eel--)
            for (int w : { 4, 8, 12, 16 })
            {
                for (int rep = 0; rep < 2; rep++)
                {
                    cudaEventRecord(e0);
                    if (eel == 1 && mode == 0) k_bench<1, 0><<<148 * w, 32>>>(nsteps, d_k, P, d_o);
                    if (eel == 1 && mode == 1) k_bench<1, 1><<<148 * w, 32>>>(nsteps, d_k, P, d_o);
                    if (eel == 0 && mode == 0) k_bench<0, 0><<<148 * w, 32>>>(nsteps, d_k, P, d_o);
                    if (eel == 0 && mode == 1) k_bench<0, 1><<<148 * w, 32>>>(nsteps, d_k, P, d_o);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                /* w warps per SM = w/4 per sub-partition; each does nsteps steps */
                const double cyc = ms * 1e-3 * ghz * 1e9 / (nsteps * (w / 4.0));
                printf("%s %s  %2d warps/SM: %.3f ms, %.1f cycles per step per SMSP (FMA-pipe floor: %d)\n", eel ? "ewald" : "rf   ",
                       mode ? "with exchange" : "math only    ", w, ms, cyc, eel ? 144 : 92);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
