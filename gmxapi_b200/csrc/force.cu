/* b200nb force kernels: Lennard-Jones + {reaction-field / plain cut-off, Ewald real space (analytical)}
 * cluster-pair force and energy with shift-force reduction, hand-written for sm_100a.
 *
 * Replaces (paths relative to /root/reference/src/gromacs):
 *   nbnxm/cuda/nbnxm_cuda_kernel.cuh:150-650 (K1) with its reductions nbnxm_cuda_kernel_utils.cuh:477-700;
 *   arithmetic parity target = the CPU SIMD kernels nbnxm/kernels_simd_2xmm/kernel_inner.h:226-880 and
 *   kernel_outer.h:395-452 (self terms), simd/simd_math.h:1609-1722 (Ewald correction polynomials).
 *
 * Design (not a port of the reference CUDA kernel):
 *  - one warp per list entry = one 8-atom i-cluster + shift vs a run of 8-atom j-clusters;
 *    the i-atom lives in registers for the whole entry: no shared memory, no per-tile i reloads;
 *  - lane = il + 8*jq evaluates the two atom pairs (il, jq) and (il, jq+4) of a tile at once, held as
 *    float2 register pairs, so the arithmetic maps onto Blackwell's packed FP32 pipe instructions
 *    (fma.rn.f32x2 / mul / add -> SASS FFMA2/FMUL2/FADD2) which need half the issue slots per flop;
 *  - the pair-interleaved coordinate layout (b200nb_internal.h) makes every j operand ONE 8-byte,
 *    broadcast-friendly read-only load straight into an aligned register pair;
 *  - j-forces: transposed butterfly over the 8 lanes sharing a j atom (7 shuffles per tile instead of
 *    18), then scalar red.global; i-forces: registers, 6 shuffles + one vector atomic per entry;
 *  - r^2 is evaluated with the reference's operand roles and operation order so the in-range pair set is
 *    bit-identical (see nb_rsq in b200nb_internal.h).
 */
#include <cstdio>

#include "b200nb_internal.h"

#ifndef B200NB_USE_F32X2
#define B200NB_USE_F32X2 1
#endif

namespace
{

__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
#if B200NB_USE_F32X2
    return __fmul2_rn(a, b);
#else
    return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
#endif
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
#if B200NB_USE_F32X2
    return __fadd2_rn(a, b);
#else
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
#endif
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
#if B200NB_USE_F32X2
    return __ffma2_rn(a, b, c);
#else
    return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
#endif
}
__device__ __forceinline__ float2 dup(float a)
{
    return make_float2(a, a);
}
__device__ __forceinline__ float2 ldg2(const float* p)
{
    return __ldg(reinterpret_cast<const float2*>(p));
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct JData
{
    float2 x, y, z, q, c6, c12;
    int    cj;
};

template<bool GEOM>
__device__ __forceinline__ void load_j(JData& d, int cj, int jq, const float* __restrict__ xq, const float* __restrict__ lj,
                                       const int* __restrict__ atype, const float2* __restrict__ nbfp, int tioff)
{
    const float* jb = xq + (size_t)cj * NB_XQ_STRIDE + 2 * jq;
    d.cj            = cj;
    d.x             = ldg2(jb);
    d.y             = ldg2(jb + 8);
    d.z             = ldg2(jb + 16);
    d.q             = ldg2(jb + 24);
    if (GEOM)
    {
        const float* lb = lj + (size_t)cj * NB_LJ_STRIDE + 2 * jq;
        d.c6            = ldg2(lb);
        d.c12           = ldg2(lb + 8);
    }
    else
    {
        const int2   t  = __ldg(reinterpret_cast<const int2*>(atype + (size_t)cj * 8 + 2 * jq));
        const float2 pa = __ldg(nbfp + tioff + t.x), pb = __ldg(nbfp + tioff + t.y);
        d.c6            = make_float2(pa.x, pb.x);
        d.c12           = make_float2(pa.y, pb.y);
    }
}

/* simd/simd_math.h:1609-1650 pmeForceCorrection, two arguments at once */
__device__ __forceinline__ float2 pme_force_corr2(float2 z2)
{
    const float2 FN6 = dup(-1.7357322914161492954e-8f), FN5 = dup(1.4703624142580877519e-6f),
                 FN4 = dup(-0.000053401640219807709149f), FN3 = dup(0.0010054721316683106153f),
                 FN2 = dup(-0.019278317264888380590f), FN1 = dup(0.069670166153766424023f),
                 FN0 = dup(-0.75225204789749321333f);
    const float2 FD4 = dup(0.0011193462567257629232f), FD3 = dup(0.014866955030185295499f),
                 FD2 = dup(0.11583842382862377919f), FD1 = dup(0.50736591960530292870f), FD0 = dup(1.0f);
    const float2 z4 = mul2(z2, z2);
    float2       d0 = fma2(FD4, z4, FD2), d1 = fma2(FD3, z4, FD1);
    d0              = fma2(d0, z4, FD0);
    d0              = fma2(d1, z2, d0);
    float2 n0 = fma2(FN6, z4, FN4), n1 = fma2(FN5, z4, FN3);
    n0 = fma2(n0, z4, FN2);
    n1 = fma2(n1, z4, FN1);
    n0 = fma2(n0, z4, FN0);
    n0 = fma2(n1, z2, n0);
    const float2 r = make_float2(rcp_approx(d0.x), rcp_approx(d0.y));
    return mul2(n0, r);
}

/* simd/simd_math.h:1687-1722 pmePotentialCorrection */
__device__ __forceinline__ float2 pme_pot_corr2(float2 z2)
{
    const float2 VN6 = dup(1.9296833005951166339e-8f), VN5 = dup(-1.4213390571557850962e-6f),
                 VN4 = dup(0.000041603292906656984871f), VN3 = dup(-0.00013134036773265025626f),
                 VN2 = dup(0.038657983986041781264f), VN1 = dup(0.11285044772717598220f),
                 VN0 = dup(1.1283802385263030286f);
    const float2 VD3 = dup(0.0066752224023576045451f), VD2 = dup(0.078647795836373922256f),
                 VD1 = dup(0.43336185284710920150f), VD0 = dup(1.0f);
    const float2 z4 = mul2(z2, z2);
    float2       d1 = fma2(VD3, z4, VD1), d0 = fma2(VD2, z4, VD0);
    d0              = fma2(d1, z2, d0);
    float2 n0 = fma2(VN6, z4, VN4), n1 = fma2(VN5, z4, VN3);
    n0 = fma2(n0, z4, VN2);
    n1 = fma2(n1, z4, VN1);
    n0 = fma2(n0, z4, VN0);
    n0 = fma2(n1, z2, n0);
    const float2 r = make_float2(rcp_approx(d0.x), rcp_approx(d0.y));
    return mul2(n0, r);
}

struct IData
{
    float2 x, y, z, q, c6, c12; /* duplicated i-atom values */
    float  qraw;
    int    il, jq, ci, shift;
};

/* One tile: lane computes pairs (il, jq) [.x] and (il, jq+4) [.y].  Returns the force on the i-atom in
 * (tx,ty,tz) (both pairs, packed) and accumulates energies. */
template<int EEL, bool VF, bool MASKED>
__device__ __forceinline__ void tile_pairs(const IData& I, const JData& J, const NbParamsDev& P, uint64_t mask, bool diag, float2& tx,
                                           float2& ty, float2& tz, float& evdw, float& ecoul)
{
    const float2 m1 = dup(-1.0f);
    const float2 dx = fma2(J.x, m1, I.x); /* xi - xj, exact product */
    const float2 dy = fma2(J.y, m1, I.y);
    const float2 dz = fma2(J.z, m1, I.z);
    const float2 r2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
    bool wa = r2.x < P.rc2, wb = r2.y < P.rc2;
    float2 inter = dup(1.0f);
    if (MASKED)
    {
        inter.x = (float)((mask >> (I.jq * 8 + I.il)) & 1ull);
        inter.y = (float)((mask >> ((I.jq + 4) * 8 + I.il)) & 1ull);
        if (diag)
        {
            /* self tile: only j > i (nbnxm/pairlist.cpp:880-904, kernel_gpu_ref.cpp:223-226) */
            wa = wa && (I.jq > I.il);
            wb = wb && (I.jq + 4 > I.il);
        }
    }
    const float2 r2c  = make_float2(fmaxf(r2.x, NB_MIN_RSQ), fmaxf(r2.y, NB_MIN_RSQ));
    const float2 rinv = make_float2(rsqrt_approx(r2c.x), rsqrt_approx(r2c.y));
    const float2 rinvsq = mul2(rinv, rinv);
    float2 rinv_ex = rinv;
    if (MASKED) rinv_ex = mul2(rinv, inter);

    /* Lennard-Jones: F*r = 12 C12 r^-12 - 6 C6 r^-6 with the 6/12 folded into the parameters */
    const float2 c6 = mul2(I.c6, J.c6), c12 = mul2(I.c12, J.c12);
    float2       rinv6 = mul2(mul2(rinvsq, rinvsq), rinvsq);
    if (MASKED) rinv6 = mul2(rinv6, inter);
    const float2 frlj6 = mul2(c6, rinv6);
    const float2 frlj12 = mul2(mul2(c12, rinv6), rinv6);
    const float2 frlj   = fma2(frlj6, m1, frlj12);

    const float2 qq = mul2(I.q, J.q);
    float2       frcoul;
    float2       vcoul = dup(0.0f);
    if (EEL == 1)
    {
        const float2 brsq   = mul2(dup(P.beta2), r2c);
        const float2 ewcorr = mul2(dup(P.beta), pme_force_corr2(brsq));
        frcoul              = mul2(qq, fma2(ewcorr, brsq, rinv_ex));
        if (VF)
        {
            float2 vsub = mul2(dup(P.beta), pme_pot_corr2(brsq));
            vsub        = fma2(dup(P.sh_ewald), inter, vsub);
            vcoul       = mul2(qq, fma2(vsub, m1, rinv_ex));
        }
    }
    else
    {
        frcoul = mul2(qq, fma2(r2c, dup(-P.two_k_rf), rinv_ex));
        if (VF) vcoul = mul2(qq, add2(rinv_ex, fma2(r2c, dup(P.k_rf), dup(-P.c_rf))));
    }
    float2 fscal = mul2(rinvsq, add2(frcoul, frlj));
    fscal.x      = wa ? fscal.x : 0.0f;
    fscal.y      = wb ? fscal.y : 0.0f;
    tx           = mul2(fscal, dx);
    ty           = mul2(fscal, dy);
    tz           = mul2(fscal, dz);
    if (VF)
    {
        /* kernels_simd_2xmm/kernel_inner.h:612-626: V = (FrLJ12 + c12*cpot12)/12 - (FrLJ6 + c6*cpot6)/6 */
        float2 v6  = mul2(dup(1.0f / 6.0f), fma2(c6, dup(P.disp_cpot), frlj6));
        float2 v12 = mul2(dup(1.0f / 12.0f), fma2(c12, dup(P.rep_cpot), frlj12));
        float2 vlj = fma2(v6, m1, v12);
        if (MASKED) vlj = mul2(vlj, inter);
        evdw += (wa ? vlj.x : 0.0f) + (wb ? vlj.y : 0.0f);
        ecoul += (wa ? vcoul.x : 0.0f) + (wb ? vcoul.y : 0.0f);
    }
}

/* Transposed butterfly over the 8 lanes (lane bits 0-2) that share the same two j atoms: 7 shuffles leave
 * every lane with one fully reduced x-or-y component and the z component of one of the two atoms. */
__device__ __forceinline__ void reduce_store_j(const float2 tx, const float2 ty, const float2 tz, int il, int jq, int cj,
                                               float4* __restrict__ f)
{
    const unsigned full = 0xffffffffu;
    const bool     b0 = il & 1, b1 = il & 2;
    float kx = b0 ? tx.y : tx.x, ky = b0 ? ty.y : ty.x, kz = b0 ? tz.y : tz.x;
    float sx = b0 ? tx.x : tx.y, sy = b0 ? ty.x : ty.y, sz = b0 ? tz.x : tz.y;
    kx += __shfl_xor_sync(full, sx, 1);
    ky += __shfl_xor_sync(full, sy, 1);
    kz += __shfl_xor_sync(full, sz, 1);
    float v = b1 ? ky : kx, s = b1 ? kx : ky;
    v += __shfl_xor_sync(full, s, 2);
    kz += __shfl_xor_sync(full, kz, 2);
    v += __shfl_xor_sync(full, v, 4);
    kz += __shfl_xor_sync(full, kz, 4);
    float* fa = reinterpret_cast<float*>(f + ((size_t)cj * 8 + jq + (b0 ? 4 : 0)));
    if (il < 4) atomicAdd(fa + (b1 ? 1 : 0), -v);
    if (il < 2) atomicAdd(fa + 2, -kz);
}

template<int EEL, bool GEOM, bool VF>
__global__ void __launch_bounds__(128)
k_force(const Entry* __restrict__ entries, long long nentries, const int* __restrict__ tcj, const uint64_t* __restrict__ tmask,
        const float* __restrict__ xq, const float* __restrict__ lj, const int* __restrict__ atype, const float2* __restrict__ nbfp,
        const float* __restrict__ shift_vec, float4* __restrict__ f, float* __restrict__ fshift, double* __restrict__ energy,
        const NbParamsDev P, const int intra)
{
    const long long e = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= nentries) return;
    const int4 ev = __ldg(reinterpret_cast<const int4*>(entries) + e);
    const int  start = ev.z, end = ev.w;
    if (start >= end) return;
    const int lane = threadIdx.x & 31;
    IData     I;
    I.il    = lane & 7;
    I.jq    = lane >> 3;
    I.ci    = ev.x;
    I.shift = ev.y & 255;
    const int nmask = ev.y >> 8;
    {
        const float* xb = xq + (size_t)I.ci * NB_XQ_STRIDE + nb_pairpos(I.il);
        /* the reference adds the shift to the i-atom before the subtraction: kernel_outer.h:482-489 */
        I.x    = dup(__fadd_rn(__ldg(xb), __ldg(shift_vec + 3 * I.shift)));
        I.y    = dup(__fadd_rn(__ldg(xb + 8), __ldg(shift_vec + 3 * I.shift + 1)));
        I.z    = dup(__fadd_rn(__ldg(xb + 16), __ldg(shift_vec + 3 * I.shift + 2)));
        I.qraw = __ldg(xb + 24);
        I.q    = dup(P.epsfac * I.qraw);
    }
    int tioff = 0;
    if (GEOM)
    {
        const float* lb = lj + (size_t)I.ci * NB_LJ_STRIDE + nb_pairpos(I.il);
        I.c6            = dup(__ldg(lb));
        I.c12           = dup(__ldg(lb + 8));
    }
    else
    {
        tioff = __ldg(atype + (size_t)I.ci * 8 + nb_pairpos(I.il)) * P.ntypes;
        I.c6 = I.c12 = dup(1.0f);
    }
    float2 fix = dup(0.f), fiy = dup(0.f), fiz = dup(0.f);
    float  evdw = 0.f, ecoul = 0.f;

    /* software pipeline: j data one tile ahead, j-cluster index two tiles ahead */
    JData cur, nxt;
    load_j<GEOM>(cur, __ldg(tcj + start), I.jq, xq, lj, atype, nbfp, tioff);
    int cj_next = (start + 1 < end) ? __ldg(tcj + start + 1) : cur.cj;
    for (int t = start; t < end; t++)
    {
        const int cj_next2 = (t + 2 < end) ? __ldg(tcj + t + 2) : cj_next;
        load_j<GEOM>(nxt, cj_next, I.jq, xq, lj, atype, nbfp, tioff);
        float2 tx, ty, tz;
        if (t - start < nmask)
        {
            const uint64_t mask = __ldg(reinterpret_cast<const unsigned long long*>(tmask) + t);
            const bool     diag = intra && I.shift == B200NB_CENTRAL && cur.cj == I.ci;
            tile_pairs<EEL, VF, true>(I, cur, P, mask, diag, tx, ty, tz, evdw, ecoul);
            if (VF && diag && I.jq == 0)
            {
                /* self term, once per atom: kernel_outer.h:408-452 */
                ecoul -= P.epsfac * I.qraw * I.qraw * P.self_sub;
            }
        }
        else
        {
            tile_pairs<EEL, VF, false>(I, cur, P, ~0ull, false, tx, ty, tz, evdw, ecoul);
        }
        fix = add2(fix, tx);
        fiy = add2(fiy, ty);
        fiz = add2(fiz, tz);
        reduce_store_j(tx, ty, tz, I.il, I.jq, cur.cj, f);
        cur     = nxt;
        cj_next = cj_next2;
    }
    /* i-force: sum the two packed halves, then over the 4 jq groups (lane bits 3-4) */
    const unsigned full = 0xffffffffu;
    float          fx = fix.x + fix.y, fy = fiy.x + fiy.y, fz = fiz.x + fiz.y;
    fx += __shfl_xor_sync(full, fx, 8);
    fy += __shfl_xor_sync(full, fy, 8);
    fz += __shfl_xor_sync(full, fz, 8);
    fx += __shfl_xor_sync(full, fx, 16);
    fy += __shfl_xor_sync(full, fy, 16);
    fz += __shfl_xor_sync(full, fz, 16);
    if (I.jq == 0) atomicAdd(f + ((size_t)I.ci * 8 + I.il), make_float4(fx, fy, fz, 0.f));
    if (VF)
    {
        /* shift force = sum of the i-forces of this entry (kernel_outer.h:620-640; the CUDA kernel skips
         * the central shift, nbnxm_cuda_kernel.cuh:624-628) */
        if (I.shift != B200NB_CENTRAL)
        {
            fx += __shfl_xor_sync(full, fx, 1);
            fy += __shfl_xor_sync(full, fy, 1);
            fz += __shfl_xor_sync(full, fz, 1);
            fx += __shfl_xor_sync(full, fx, 2);
            fy += __shfl_xor_sync(full, fy, 2);
            fz += __shfl_xor_sync(full, fz, 2);
            fx += __shfl_xor_sync(full, fx, 4);
            fy += __shfl_xor_sync(full, fy, 4);
            fz += __shfl_xor_sync(full, fz, 4);
            if (lane == 0)
            {
                atomicAdd(fshift + 3 * I.shift, fx);
                atomicAdd(fshift + 3 * I.shift + 1, fy);
                atomicAdd(fshift + 3 * I.shift + 2, fz);
            }
        }
        for (int o = 16; o > 0; o >>= 1)
        {
            evdw += __shfl_xor_sync(full, evdw, o);
            ecoul += __shfl_xor_sync(full, ecoul, o);
        }
        if (lane == 0)
        {
            atomicAdd(energy, (double)evdw);
            atomicAdd(energy + 1, (double)ecoul);
        }
    }
}

template<int EEL, bool GEOM, bool VF>
int launch(b200nb_context* h, const PairList& L, int intra)
{
    const unsigned nblk = (unsigned)((L.nentries + 3) / 4);
    k_force<EEL, GEOM, VF><<<nblk, 128, 0, h->stream>>>(L.entries, L.nentries, L.cj, L.mask, h->d_xq, h->d_lj, h->d_atype,
                                                        reinterpret_cast<const float2*>(h->d_nbfp), h->d_shift_vec, h->d_f, h->d_fshift,
                                                        h->d_energy, h->dp, intra);
    h->nlaunches++;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return nb_fail(h, B200NB_ERR_CUDA, std::string("force kernel launch: ") + cudaGetErrorString(err));
    return 0;
}

} // namespace

int nb_launch_force_kernel(b200nb_context* h, int loc, int flags)
{
    const PairList& L = h->inner[loc];
    if (L.nentries == 0) return 0;
    const bool vf    = (flags & (B200NB_FLAG_ENERGY | B200NB_FLAG_VIRIAL)) != 0;
    const bool ewald = h->dp.eeltype == B200NB_EEL_EWALD;
    const int  intra = (loc == 0);
    if (ewald)
    {
        if (h->comb_geom) return vf ? launch<1, true, true>(h, L, intra) : launch<1, true, false>(h, L, intra);
        return vf ? launch<1, false, true>(h, L, intra) : launch<1, false, false>(h, L, intra);
    }
    if (h->comb_geom) return vf ? launch<0, true, true>(h, L, intra) : launch<0, true, false>(h, L, intra);
    return vf ? launch<0, false, true>(h, L, intra) : launch<0, false, false>(h, L, intra);
}
