/* b200nb force kernels: Lennard-Jones + {reaction-field / plain cut-off, Ewald real space (analytical)}
 * cluster-pair force and energy with shift-force reduction, hand-written for sm_100a.
 *
 * Replaces (paths relative to /root/reference/src/gromacs):
 *   nbnxm/cuda/nbnxm_cuda_kernel.cuh:150-650 (K1) with its reductions nbnxm_cuda_kernel_utils.cuh:477-700;
 *   arithmetic parity target = the CPU SIMD kernels nbnxm/kernels_simd_2xmm/kernel_inner.h:226-880 and
 *   kernel_outer.h:395-452 (self terms), simd/simd_math.h:1609-1722 (Ewald correction polynomials).
 *
 * Design (not a port of the reference CUDA kernel), sized with profiles/tools/microbench.cu and the diagnostic builds of
 * profiles/r1/v_sweep_pair_loop_diagnostics.txt: on sm_100 the FP32 pipe retires 128 lane-FMAs/clk/SM whether issued as
 * scalar FFMA or packed FFMA2; a packed instruction costs its sub-partition two issue cycles and every other instruction
 * one, and that -- instruction issue with the FP32 pipe as its largest client -- is what bounds this kernel, not latency.
 * ALU-pipe instructions (FSEL, FMNMX, LOP3, IADD3) run at half rate on their own pipe, SHFL at ~0.44 warp-instructions/
 * clk/SMSP, MUFU at 16 lanes/clk/SM.  So the kernel is written to spend its issue slots on FP32 math:
 *  - one warp per list entry = one 8-atom i-cluster + shift against a run of PACKED tiles of 8 j-atom slots each
 *    (PackedList, b200nb_internal.h: j-atoms with no pair inside the list radius were dropped when the list was packed);
 *  - lane = jl + 8*ih holds the j-atom in slot jl of the tile and the TWO i-atoms (2*ih, 2*ih+1) as packed float2 registers
 *    for the whole entry: all pair arithmetic is packed (fma.rn.f32x2 -> FFMA2/FMUL2/FADD2), the j operands enter as the
 *    scalar-broadcast operand form of those instructions, so a tile costs two 16-byte shared-memory loads (xyzq;
 *    LJ pair + the j-atom's slot index; the 4 lanes sharing a j-atom hit the same address) and no register shuffling;
 *  - j-forces: in-lane add of the two pairs, 2-stage reduce-scatter (3 shuffles) over the 4 lanes sharing the j-atom,
 *    then one scalar red.global.add.f32 per lane: 32 lanes cover the float4 force slots of the tile's 8 j-atoms;
 *  - i-forces stay in registers; per entry one transposed butterfly over the 8 j-lanes and one 16-byte red per atom;
 *  - tiles that carry exclusion masks are sorted to the front of an entry (b200nb.cu k_search) and run through a
 *    separate code path; the unmasked path has no mask logic and no r^2 clamp;
 *  - LJ is evaluated as (c12*r^-6 - c6)*r^-6, the Ewald correction polynomials keep their coefficients as
 *    instruction immediates (folding beta^3 into them would cost seven registers); out-of-range lanes are discarded by
 *    select, so garbage there cannot poison a sum;
 *  - r^2 is evaluated with the reference's operand roles and operation order so the in-range pair set is
 *    bit-identical (see nb_rsq in b200nb_internal.h);
 *  - the j-atom data go through a two-buffer ring of NB_CHUNK tiles in shared memory, filled with cp.async (LDGSTS): each
 *    lane reads one j-slot index (coalesced) and gathers that atom's 16-byte xyzq and 8-byte LJ pair while the previous
 *    chunk is being computed; the pair loop never touches L2, and a warp needs 4 KB whatever the length of its entry.
 *    (The first version loaded j data with LDG one tile ahead and spent its time in long-scoreboard stalls; the second
 *    staged a whole entry at once, which tied the entry length to the shared memory per warp.)  The i-cluster lives in
 *    registers, which beats shared memory.
 *    TMA (cp.async.bulk) was considered and not used: the stream is a gather of 16-byte atoms by slot index, so a bulk copy
 *    would move one atom per instruction issued by one elected lane plus mbarrier traffic, while LDGSTS moves 512 B per
 *    warp instruction with per-lane addresses and needs only wait_group + syncwarp;
 *  - LJ force switch / potential switch / VdW cut-off below the Coulomb cut-off are a second set of instantiations (GEN):
 *    the plain kernels, which every BASELINE configuration uses, pay nothing for them.
 */
#include <cstdio>

#include "b200nb_internal.h"

namespace
{

__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 dup(float a) { return make_float2(a, a); }

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
#ifdef B200NB_DIAG_NO_MUFU /* diagnostic build only (wrong results): an FMA in place of the special-function unit */
    r = __fmaf_rn(x, 0.001f, 1.0f);
#else
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
#endif
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
#ifdef B200NB_DIAG_NO_MUFU
    r = __fmaf_rn(x, -0.5f, 1.5f);
#else
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
#endif
    return r;
}

__device__ __forceinline__ float exp_approx(float x) /* e^x through MUFU.EX2; relative error ~1e-7 for the |x| <= 12 used here */
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}

struct JAtom
{
    float4 xq;
    float2 lj; /* GEOM: sqrt(6 C6), sqrt(12 C12); table: atom type in lj.x (as int bits) */
    int    slot; /* grid slot of the j-atom: index into xq / f */
};

/* The j-atom data of a whole entry is staged in shared memory with cp.async (LDGSTS) before the pair loop, so the
 * loop itself never waits on L2.  Per warp and packed tile 256 B: [8 x float4 xyzq][8 x {lj.x, lj.y, slot, -}]; the 8
 * lanes of a quarter-warp read 128 contiguous bytes (no bank conflicts), the 4 quarter-warps the same ones (broadcast). */
#define NB_TILE_SMEM 256
__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void* gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem_addr, const void* gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gptr) : "memory");
}

/* s_lane = this warp's staging area + jl*16 */
__device__ __forceinline__ void load_j(JAtom& J, const unsigned char* s_lane, int t)
{
    J.xq           = *reinterpret_cast<const float4*>(s_lane + t * NB_TILE_SMEM);
    const float4 a = *reinterpret_cast<const float4*>(s_lane + t * NB_TILE_SMEM + 128);
    J.lj           = make_float2(a.x, a.y);
    J.slot         = __float_as_int(a.z);
}

/* simd/simd_math.h:1609-1650 pmeForceCorrection: denominator and numerator */
/* all five coefficients arrive divided by beta (KConst), so 1/den is already beta / denominator: saves the multiplication by
 * beta per pair at the cost of three more loop-invariant registers */
__device__ __forceinline__ float2 pme_force_den(float2 z2, float2 z4, float fd4, float fd3, float fd2, float fd1, float fd0)
{
    const float2 FD4 = dup(fd4), FD3 = dup(fd3), FD2 = dup(fd2), FD1 = dup(fd1), FD0 = dup(fd0);
    float2 d0 = fma2(FD4, z4, FD2), d1 = fma2(FD3, z4, FD1);
    d0        = fma2(d0, z4, FD0);
    return fma2(d1, z2, d0);
}
__device__ __forceinline__ float2 pme_force_num(float2 z2, float2 z4, float fn6, float fn5)
{
    const float2 FN6 = dup(fn6), FN5 = dup(fn5),
                 FN4 = dup(-0.000053401640219807709149f), FN3 = dup(0.0010054721316683106153f),
                 FN2 = dup(-0.019278317264888380590f), FN1 = dup(0.069670166153766424023f),
                 FN0 = dup(-0.75225204789749321333f);
    float2 n0 = fma2(FN6, z4, FN4), n1 = fma2(FN5, z4, FN3);
    n0        = fma2(n0, z4, FN2);
    n1        = fma2(n1, z4, FN1);
    n0        = fma2(n0, z4, FN0);
    return fma2(n1, z2, n0);
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

/* Loop-invariant scalars. They are read from a small global array (b200nb_context::d_kconst) rather than from kernel
 * parameters or literals: a loaded value must stay in its register, whereas ptxas rematerialises constant-bank values and
 * immediates inside the pair loop (7 issue slots per tile in the first version). */
struct KConst
{
    float rc2, beta2, fd4, fd3, fd2, fd1, fd0, fn6, fn5; /* fd*: pmeForceCorrection denominator coefficients / beta */
};
struct IData
{
    float2 x, y, z;  /* the two i-atoms of this lane, shift already added */
    float2 q;        /* epsfac * q_i */
    float2 c6n, c12; /* GEOM: -sqrt(6 C6_i), sqrt(12 C12_i) */
    int    t0, t1;   /* table path: type_i * ntypes */
    float2 g0, g1;   /* LJ-PME (GEN kernels): nbfp_comb of the two i-atoms' types */
};

/* One tile: lane computes pairs (i0, jl) [.x] and (i1, jl) [.y].  Returns the force ON THE i-ATOMS in (tx,ty,tz). */
template<int EEL, bool GEOM, bool VF, bool MASKED, bool GEN>
__device__ __forceinline__ void tile_pairs(const IData& I, const JAtom& J, const NbParamsDev& P, const KConst& K, const float2* __restrict__ nbfp,
                                           const float2* __restrict__ nbfp_comb, float inter0, float inter1, bool ok0, bool ok1, float2& tx, float2& ty, float2& tz,
                                           float& evdw, float& ecoul)
{
    const float2 m1 = dup(-1.0f);
    const float2 dx = fma2(dup(J.xq.x), m1, I.x); /* xi - xj: the product is exact */
    const float2 dy = fma2(dup(J.xq.y), m1, I.y);
    const float2 dz = fma2(dup(J.xq.z), m1, I.z);
    float2       r2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
    bool         wa = r2.x < K.rc2, wb = r2.y < K.rc2;
    float2       inter;
    if (MASKED)
    {
        wa    = wa && ok0;
        wb    = wb && ok1;
        inter = make_float2(inter0, inter1);
        r2    = make_float2(fmaxf(r2.x, NB_MIN_RSQ), fmaxf(r2.y, NB_MIN_RSQ));
    }
    const float2 rinv   = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
    const float2 rinvsq = mul2(rinv, rinv);
    float2       rinv_ex = rinv;
    if (MASKED) rinv_ex = mul2(rinv, inter);

    float2 c6n, c12;
    if (GEOM)
    {
        c6n = mul2(I.c6n, dup(J.lj.x));
        c12 = mul2(I.c12, dup(J.lj.y));
    }
    else
    {
        const int    tj = __float_as_int(J.lj.x);
        const float2 pa = __ldg(nbfp + I.t0 + tj), pb = __ldg(nbfp + I.t1 + tj);
        c6n             = make_float2(-pa.x, -pb.x);
        c12             = make_float2(pa.y, pb.y);
    }
    float2 rinv6 = mul2(mul2(rinvsq, rinvsq), rinvsq);
    if (MASKED) rinv6 = mul2(rinv6, inter);
    float2 fsum; /* F*r summed over LJ and Coulomb */
    float2 frlj6, frlj12;
    float2 vlj = dup(0.0f);
    if (GEN)
    {
        /* LJ with a force or potential switch and / or a VdW cut-off shorter than the Coulomb one: the arithmetic of
         * kernels_reference/kernel_ref_inner.h:152-262 (cuda/nbnxm_cuda_kernel_utils.cuh calculate_force_switch_F[_E],
         * calculate_potential_switch_F[_E]; VDW_CUTOFF_CHECK nbnxm_cuda_kernel.cuh:540-548).  c6n = -6 C6. */
        frlj6  = mul2(c6n, rinv6);
        frlj12 = mul2(mul2(c12, rinv6), rinv6);
        fsum   = add2(frlj12, frlj6);
        const bool need_v = VF || P.vdw_modifier == B200NB_VDW_POTSWITCH;
        if (need_v)
        {
            const float2 v6  = mul2(dup(1.0f / 6.0f), fma2(c6n, dup(P.disp_cpot), frlj6));
            const float2 v12 = mul2(dup(1.0f / 12.0f), fma2(c12, dup(P.rep_cpot), frlj12));
            vlj              = add2(v12, v6);
        }
        if (P.vdw_modifier != B200NB_VDW_POTSHIFT)
        {
            const float2 r   = mul2(r2, rinv);
            float2       rsw = add2(r, dup(-P.rvdw_switch));
            rsw              = make_float2(fmaxf(rsw.x, 0.0f), fmaxf(rsw.y, 0.0f));
            const float2 rsw2 = mul2(rsw, rsw);
            if (P.vdw_modifier == B200NB_VDW_FORCESWITCH)
            {
                const float2 a = fma2(c6n, fma2(dup(P.disp_c3), rsw, dup(P.disp_c2)), mul2(c12, fma2(dup(P.rep_c3), rsw, dup(P.rep_c2))));
                fsum           = fma2(a, mul2(rsw2, r), fsum);
                if (VF)
                {
                    const float2 b = fma2(c6n, fma2(dup(-0.25f * P.disp_c3), rsw, dup(-P.disp_c2 * (1.0f / 3.0f))),
                                          mul2(c12, fma2(dup(-0.25f * P.rep_c3), rsw, dup(-P.rep_c2 * (1.0f / 3.0f)))));
                    vlj            = fma2(b, mul2(rsw2, rsw), vlj);
                }
                if (MASKED) vlj = mul2(vlj, inter);
            }
            else
            {
                if (MASKED) vlj = mul2(vlj, inter);
                const float2 sw  = fma2(fma2(fma2(dup(P.sw_c5), rsw, dup(P.sw_c4)), rsw, dup(P.sw_c3)), mul2(rsw2, rsw), dup(1.0f));
                const float2 dsw = mul2(fma2(fma2(dup(5.0f * P.sw_c5), rsw, dup(4.0f * P.sw_c4)), rsw, dup(3.0f * P.sw_c3)), rsw2);
                fsum             = fma2(mul2(dsw, vlj), mul2(r, dup(-1.0f)), mul2(sw, fsum));
                vlj              = mul2(sw, vlj);
            }
        }
        else if (MASKED)
        {
            vlj = mul2(vlj, inter);
        }
        if (!GEOM && P.ljpme)
        {
            /* LJ-PME: subtract the grid part of the dispersion (kernel_ref_inner.h:207-250; cuda calculate_lj_ewald_comb_geom_F[_E],
             * calculate_lj_ewald_comb_LB_F_E).  r^-6 here is NOT masked by the exclusions: excluded pairs inside the cut-off
             * keep this correction, exactly like the Coulomb exclusion correction. */
            const float2 gj = __ldg(nbfp_comb + __float_as_int(J.lj.x));
            float2       c6grid;
            if (P.ljpme == 1)
            {
                c6grid = make_float2(I.g0.x * gj.x, I.g1.x * gj.x);
            }
            else
            {
                const float2 sg = make_float2(I.g0.x + gj.x, I.g1.x + gj.x);
                const float2 s2 = mul2(sg, sg);
                c6grid          = mul2(make_float2(I.g0.y * gj.y, I.g1.y * gj.y), mul2(mul2(s2, s2), s2));
            }
            const float2 rinv6nm = mul2(mul2(rinvsq, rinvsq), rinvsq);
            const float2 cr2     = mul2(dup(P.lje_coeff2), r2);
            const float2 nex     = make_float2(-exp_approx(-cr2.x), -exp_approx(-cr2.y)); /* -exp(-cr2) */
            const float2 poly    = fma2(fma2(dup(0.5f), cr2, dup(1.0f)), cr2, dup(1.0f));
            fsum                 = fma2(c6grid, fma2(nex, fma2(rinv6nm, poly, dup(P.lje_coeff6_6)), rinv6nm), fsum);
            if (VF)
            {
                float2 sh = dup(P.sh_lj_ewald);
                if (MASKED) sh = mul2(sh, inter);
                vlj = fma2(mul2(c6grid, dup(1.0f / 6.0f)), fma2(rinv6nm, fma2(nex, poly, dup(1.0f)), sh), vlj);
            }
        }
        /* VdW cut-off shorter than the Coulomb cut-off (PME load balancing grows rcoulomb, rvdw stays) */
        const bool va = r2.x < P.rvdw2, vb = r2.y < P.rvdw2;
        fsum.x = va ? fsum.x : 0.0f;
        fsum.y = vb ? fsum.y : 0.0f;
        vlj.x  = va ? vlj.x : 0.0f;
        vlj.y  = vb ? vlj.y : 0.0f;
    }
    else if (VF)
    {
        frlj6  = mul2(c6n, rinv6);               /* -6 C6 r^-6 */
        frlj12 = mul2(mul2(c12, rinv6), rinv6);  /* 12 C12 r^-12 */
        fsum   = add2(frlj12, frlj6);
    }
    else
    {
        fsum = mul2(fma2(c12, rinv6, c6n), rinv6);
    }
    const float2 qq = mul2(I.q, dup(J.xq.w));
    float2       vcoul = dup(0.0f);
    if (EEL == 1)
    {
        const float2 z2 = mul2(dup(K.beta2), r2);
        const float2 z4 = mul2(z2, z2);
        const float2 den = pme_force_den(z2, z4, K.fd4, K.fd3, K.fd2, K.fd1, K.fd0);
        const float2 num = pme_force_num(z2, z4, K.fn6, K.fn5);
        const float2 t   = mul2(num, make_float2(rcp_approx(den.x), rcp_approx(den.y))); /* beta * pmecorrF(z2) */
        fsum             = fma2(qq, fma2(t, z2, rinv_ex), fsum);
        if (VF)
        {
            float2 vsub = mul2(dup(P.beta), pme_pot_corr2(z2));
            if (MASKED) vsub = fma2(dup(P.sh_ewald), inter, vsub);
            else vsub = add2(vsub, dup(P.sh_ewald));
            vcoul = mul2(qq, fma2(vsub, m1, rinv_ex));
        }
    }
    else
    {
        fsum = fma2(qq, fma2(r2, dup(-P.two_k_rf), rinv_ex), fsum);
        if (VF) vcoul = mul2(qq, add2(rinv_ex, fma2(r2, dup(P.k_rf), dup(-P.c_rf))));
    }
    float2 fscal = mul2(rinvsq, fsum);
    fscal.x      = wa ? fscal.x : 0.0f;
    fscal.y      = wb ? fscal.y : 0.0f;
    tx           = mul2(fscal, dx);
    ty           = mul2(fscal, dy);
    tz           = mul2(fscal, dz);
    if (VF)
    {
        if (!GEN)
        {
            /* kernels_simd_2xmm/kernel_inner.h:612-626: V = (FrLJ12 + c12*cpot12)/12 - (FrLJ6 + c6*cpot6)/6 */
            float2 v6  = mul2(dup(1.0f / 6.0f), fma2(c6n, dup(P.disp_cpot), frlj6)); /* = -(FrLJ6 + c6 cpot6)/6 */
            float2 v12 = mul2(dup(1.0f / 12.0f), fma2(c12, dup(P.rep_cpot), frlj12));
            vlj        = add2(v12, v6);
            if (MASKED) vlj = mul2(vlj, inter);
        }
        evdw += (wa ? vlj.x : 0.0f) + (wb ? vlj.y : 0.0f);
        ecoul += (wa ? vcoul.x : 0.0f) + (wb ? vcoul.y : 0.0f);
    }
}

#ifndef NB_CHUNK
#define NB_CHUNK 8 /* packed tiles per ring buffer (multiple of 4: one staging round covers 32 j-atoms) */
#endif
/* NT plain (unmasked, force-only) tiles evaluated in lock step: every step of the arithmetic is written for all NT tiles
 * before the next step, so the instruction stream carries NT independent dependency chains.  The kernel uses NT = 1: one
 * tile in flight at 32 warps/SM beat two at 20 (profiles/r1/g_sweep_one_entry_per_warp.txt); the generality is kept because
 * the statement order below (not the tile count) is what makes ptxas emit the packed sequence this file is tuned around. */
template<int EEL, bool GEOM, int NT>
__device__ __forceinline__ void tile_pairs_multi(const IData& I, const JAtom (&J)[NT], const NbParamsDev& P, const KConst& K,
                                                 const float2* __restrict__ nbfp, float2& fix, float2& fiy, float2& fiz, float (&sx)[NT],
                                                 float (&sy)[NT], float (&sz)[NT])
{
    const float2 m1 = dup(-1.0f);
    float2       dx[NT], dy[NT], dz[NT], r2[NT], rinv[NT], rinvsq[NT], rinv6[NT], c6n[NT], c12[NT], fsum[NT], qq[NT];
#pragma unroll
    for (int u = 0; u < NT; u++) dx[u] = fma2(dup(J[u].xq.x), m1, I.x);
#pragma unroll
    for (int u = 0; u < NT; u++) dy[u] = fma2(dup(J[u].xq.y), m1, I.y);
#pragma unroll
    for (int u = 0; u < NT; u++) dz[u] = fma2(dup(J[u].xq.z), m1, I.z);
#pragma unroll
    for (int u = 0; u < NT; u++) r2[u] = mul2(dy[u], dy[u]);
#pragma unroll
    for (int u = 0; u < NT; u++) r2[u] = fma2(dx[u], dx[u], r2[u]);
#pragma unroll
    for (int u = 0; u < NT; u++) r2[u] = fma2(dz[u], dz[u], r2[u]);
#pragma unroll
    for (int u = 0; u < NT; u++) rinv[u] = make_float2(rsqrt_approx(r2[u].x), rsqrt_approx(r2[u].y));
#pragma unroll
    for (int u = 0; u < NT; u++)
    {
        if (GEOM)
        {
            c6n[u] = mul2(I.c6n, dup(J[u].lj.x));
            c12[u] = mul2(I.c12, dup(J[u].lj.y));
        }
        else
        {
            const int    tj = __float_as_int(J[u].lj.x);
            const float2 pa = __ldg(nbfp + I.t0 + tj), pb = __ldg(nbfp + I.t1 + tj);
            c6n[u]          = make_float2(-pa.x, -pb.x);
            c12[u]          = make_float2(pa.y, pb.y);
        }
        qq[u] = mul2(I.q, dup(J[u].xq.w));
    }
    float2 z2[NT], z4[NT], den[NT], num[NT];
    if (EEL == 1)
    {
#pragma unroll
        for (int u = 0; u < NT; u++) z2[u] = mul2(dup(K.beta2), r2[u]);
#pragma unroll
        for (int u = 0; u < NT; u++) z4[u] = mul2(z2[u], z2[u]);
#pragma unroll
        for (int u = 0; u < NT; u++) den[u] = pme_force_den(z2[u], z4[u], K.fd4, K.fd3, K.fd2, K.fd1, K.fd0);
#pragma unroll
        for (int u = 0; u < NT; u++) den[u] = make_float2(rcp_approx(den[u].x), rcp_approx(den[u].y));
#pragma unroll
        for (int u = 0; u < NT; u++) num[u] = pme_force_num(z2[u], z4[u], K.fn6, K.fn5);
    }
#pragma unroll
    for (int u = 0; u < NT; u++) rinvsq[u] = mul2(rinv[u], rinv[u]);
#pragma unroll
    for (int u = 0; u < NT; u++) rinv6[u] = mul2(mul2(rinvsq[u], rinvsq[u]), rinvsq[u]);
#pragma unroll
    for (int u = 0; u < NT; u++) fsum[u] = mul2(fma2(c12[u], rinv6[u], c6n[u]), rinv6[u]);
#pragma unroll
    for (int u = 0; u < NT; u++)
    {
        if (EEL == 1) fsum[u] = fma2(qq[u], fma2(mul2(num[u], den[u]), z2[u], rinv[u]), fsum[u]);
        else fsum[u] = fma2(qq[u], fma2(r2[u], dup(-P.two_k_rf), rinv[u]), fsum[u]);
    }
#pragma unroll
    for (int u = 0; u < NT; u++)
    {
        float2 fscal = mul2(rinvsq[u], fsum[u]);
        fscal.x      = (r2[u].x < K.rc2) ? fscal.x : 0.0f;
        fscal.y      = (r2[u].y < K.rc2) ? fscal.y : 0.0f;
        /* force on the two i-atoms: one packed FMA per component into the entry's accumulators; force on the j-atom: minus
         * the sum over the lane's two pairs, one scalar FMUL + FFMA per component (4 FMA-pipe cycles per component instead
         * of the 5 of packed multiply, packed add, scalar add) */
        fix   = fma2(fscal, dx[u], fix);
        fiy   = fma2(fscal, dy[u], fiy);
        fiz   = fma2(fscal, dz[u], fiz);
        sx[u] = __fmaf_rn(-fscal.y, dx[u].y, -(fscal.x * dx[u].x));
        sy[u] = __fmaf_rn(-fscal.y, dy[u].y, -(fscal.x * dy[u].x));
        sz[u] = __fmaf_rn(-fscal.y, dz[u].y, -(fscal.x * dz[u].x));
    }
}

/* j-force: both pairs of the lane act on the same j-atom: sum them (negated: the force on j is minus the force on i), then
 * a 2-stage reduce-scatter over the 4 lanes (bits 4 and 3) that share the j-atom leaves lane class (b4,b3) = (0,0) with the
 * x total, (0,1) y, (1,0) z and (1,1) a second copy of z; every lane then issues ONE scalar red.global.add.f32 to float
 * slot 2*b4+b3 of f[cj*8+jl] -- the copy lands in the float4's pad slot, which nothing reads.  3 shuffles, 4 selects, no
 * divergent branch (a predicated 16-byte red by 8 lanes compiled to BSSY/BRA/BSYNC plus a convergence check before the
 * next shuffle: 7 more instructions per tile). */
struct LaneClass
{
    bool  b4, b3, b3or4, b3only;
    char* f_lane; /* f + component slot (2*b4+b3) * 4 bytes */
};
__device__ __forceinline__ void reduce_store_j(const float2 tx, const float2 ty, const float2 tz, const LaneClass& C, int jslot)
{
    const unsigned full = 0xffffffffu;
    const float sx = -tx.x - tx.y, sy = -ty.x - ty.y, sz = -tz.x - tz.y;
    float k0 = C.b4 ? sz : sx;
    k0 += __shfl_xor_sync(full, C.b4 ? sx : sz, 16);   /* b4=0: x of both halves, b4=1: z of both halves */
    const float k1 = sy + __shfl_xor_sync(full, sy, 16); /* y of both halves (used by the b4=0 lanes) */
    float v = C.b3only ? k1 : k0;
    v += __shfl_xor_sync(full, C.b3or4 ? k0 : k1, 8);
    atomicAdd(reinterpret_cast<float*>(C.f_lane + (size_t)(unsigned)jslot * 16u), v);
}

template<int NT>
__device__ __forceinline__ void reduce_store_j_multi(const float (&sx)[NT], const float (&sy)[NT], const float (&sz)[NT], const LaneClass& C,
                                                     const int (&jslot)[NT])
{
    const unsigned full = 0xffffffffu;
    float          k0[NT], k1[NT], v[NT];
#pragma unroll
    for (int u = 0; u < NT; u++) k0[u] = __shfl_xor_sync(full, C.b4 ? sx[u] : sz[u], 16);
#pragma unroll
    for (int u = 0; u < NT; u++) k1[u] = __shfl_xor_sync(full, sy[u], 16);
#pragma unroll
    for (int u = 0; u < NT; u++) k0[u] += C.b4 ? sz[u] : sx[u], k1[u] += sy[u];
#pragma unroll
    for (int u = 0; u < NT; u++) v[u] = __shfl_xor_sync(full, C.b3or4 ? k0[u] : k1[u], 8);
#pragma unroll
    for (int u = 0; u < NT; u++) v[u] += C.b3only ? k1[u] : k0[u];
#pragma unroll
    for (int u = 0; u < NT; u++)
    {
#ifdef B200NB_DIAG_NO_RED /* diagnostic build only: drops the j-force scatter (wrong results) to measure its cost */
        if (v[u] == 12345.678f)
#endif
            atomicAdd(reinterpret_cast<float*>(C.f_lane + (size_t)(unsigned)jslot[u] * 16u), v[u]);
    }
}

#ifndef B200NB_FORCE_WARPS
#define B200NB_FORCE_WARPS 1 /* warps (= list entries) per CTA: small CTAs refill an SM's warp slots at entry granularity */
#endif
#ifndef B200NB_FORCE_MIN_BLOCKS
#define B200NB_FORCE_MIN_BLOCKS (32 / B200NB_FORCE_WARPS) /* 32 resident warps per SM = 64 registers per thread */
#endif
template<int EEL, bool GEOM, bool VF, bool GEN>
__global__ void __launch_bounds__(32 * B200NB_FORCE_WARPS, B200NB_FORCE_MIN_BLOCKS)
k_force(const Entry* __restrict__ entries, long long nentries, const int* __restrict__ pja, const uint64_t* __restrict__ tmask,
        const float4* __restrict__ xq, const float2* __restrict__ lj, const int* __restrict__ atype, const float2* __restrict__ nbfp,
        const float* __restrict__ shift_vec, float4* __restrict__ f, float* __restrict__ fshift, double* __restrict__ energy,
        const __grid_constant__ NbParamsDev P, const int intra, const int maxt, const float* __restrict__ kconst,
        const float2* __restrict__ nbfp_comb)
{
    const long long e = (long long)blockIdx.x * B200NB_FORCE_WARPS + (threadIdx.x >> 5);
    if (e >= nentries) return;
    const int lane = threadIdx.x & 31;
    /* Entry e owns the packed tiles [e*maxt, (e+1)*maxt) (maxt = the list's pitch).  Its j data go through a two-buffer ring of
     * NB_CHUNK tiles in shared memory: chunk c+1 is gathered with cp.async while chunk c is being computed, so the shared
     * memory per warp is constant (2 x NB_CHUNK x 256 B) whatever the entry's length, and entries can hold a whole
     * (i-cluster, shift) list.  The kernel is issue-bound (profiles/r1/v_*): what an entry costs is its instruction count,
     * not the latency of these loads -- long entries pay the prologue / epilogue instructions less often. */
    const int* const ja = pja + (size_t)e * maxt * 8;
    const int4       ev = __ldg(reinterpret_cast<const int4*>(entries) + e);
    KConst K;
    {
        const float4 k0 = __ldg(reinterpret_cast<const float4*>(kconst)), k1 = __ldg(reinterpret_cast<const float4*>(kconst) + 1),
                     k2 = __ldg(reinterpret_cast<const float4*>(kconst) + 2);
        K.rc2 = k0.x, K.beta2 = k0.z, K.fd4 = k0.w, K.fd3 = k1.x, K.fn6 = k1.y, K.fn5 = k1.z, K.fd2 = k1.w, K.fd1 = k2.x, K.fd0 = k2.y;
    }
    /* Programmatic dependent launch: everything above reads only the list; from here on the kernel touches xq and f, so wait
     * for the completion of the preceding kernel of the stream (k_step_begin) -- a no-op for a normally serialised launch. */
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int  start = ev.z, end = ev.w;
    const bool self  = VF && NB_ENTRY_SELF(ev.y);
    if (start >= end && !self) return;
    const int jl = lane & 7, ih = lane >> 3;
    const int ci = ev.x, shift = NB_ENTRY_SHIFT(ev.y), nmask = NB_ENTRY_NMASK(ev.y);
    const int ntile = end - start;
    const unsigned full = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CHUNK_BYTES = NB_CHUNK * NB_TILE_SMEM;
    unsigned char* const sj = smem_raw + (threadIdx.x >> 5) * (2 * CHUNK_BYTES);
    const unsigned       s0 = (unsigned)__cvta_generic_to_shared(sj) + (lane >> 3) * NB_TILE_SMEM + (lane & 7) * 16;
    auto stage = [&](int c) {
        /* each lane gathers one j-atom per round: 16 B xyzq, 8 B LJ pair (or 4 B type), and keeps the slot index beside them */
        const unsigned sb = s0 + (c & 1) * CHUNK_BYTES;
#pragma unroll
        for (int r = 0; r < NB_CHUNK / 4; r++)
        {
            const int a = c * (NB_CHUNK * 8) + 32 * r + lane;
            if (a < ntile * 8)
            {
                const int      slot = __ldg(ja + a);
                const unsigned dx   = sb + r * 4 * NB_TILE_SMEM;
                cp_async16(dx, xq + slot);
                if (GEOM) cp_async8(dx + 128, lj + slot);
                else cp_async4(dx + 128, atype + slot);
                asm volatile("st.shared.s32 [%0], %1;" ::"r"(dx + 136), "r"(slot) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0);
    LaneClass C;
    C.b4     = (lane & 16) != 0;
    C.b3     = (lane & 8) != 0;
    C.b3or4  = C.b3 || C.b4;
    C.b3only = C.b3 && !C.b4;
    C.f_lane = reinterpret_cast<char*>(f) + 4 * (2 * (int)C.b4 + (int)C.b3);
    IData I;
    {
        const float4 a = __ldg(xq + (size_t)ci * 8 + 2 * ih), b = __ldg(xq + (size_t)ci * 8 + 2 * ih + 1);
        const float  sx = __ldg(shift_vec + 3 * shift), sy = __ldg(shift_vec + 3 * shift + 1), sz = __ldg(shift_vec + 3 * shift + 2);
        /* the reference adds the shift to the i-atom before the subtraction: kernel_outer.h:482-489 */
        I.x = make_float2(__fadd_rn(a.x, sx), __fadd_rn(b.x, sx));
        I.y = make_float2(__fadd_rn(a.y, sy), __fadd_rn(b.y, sy));
        I.z = make_float2(__fadd_rn(a.z, sz), __fadd_rn(b.z, sz));
        I.q = make_float2(P.epsfac * a.w, P.epsfac * b.w);
        if (GEOM)
        {
            const float4 l = __ldg(reinterpret_cast<const float4*>(lj + (size_t)ci * 8 + 2 * ih));
            I.c6n          = make_float2(-l.x, -l.z);
            I.c12          = make_float2(l.y, l.w);
            I.t0 = I.t1 = 0;
        }
        else
        {
            const int2 t = __ldg(reinterpret_cast<const int2*>(atype + (size_t)ci * 8 + 2 * ih));
            I.t0         = t.x * P.ntypes;
            I.t1         = t.y * P.ntypes;
            I.c6n = I.c12 = dup(0.0f);
            if (GEN && P.ljpme)
            {
                I.g0 = __ldg(nbfp_comb + t.x);
                I.g1 = __ldg(nbfp_comb + t.y);
            }
        }
    }
    float2 fix = dup(0.f), fiy = dup(0.f), fiz = dup(0.f);
    float  evdw = 0.f, ecoul = 0.f;
    if (self && jl < 2)
    {
        /* Coulomb self term, once per i-atom: kernel_outer.h:408-452 (fillers carry q = 0) */
        const float qi = (jl == 0 ? I.q.x : I.q.y);
        ecoul -= qi * qi * P.self_q2;
        if (GEN && !GEOM && P.ljpme)
        {
            /* LJ Ewald self interaction, kernel_ref_outer.h:316-321: 0.5 * (6 C6_ii) / 6 * coeff^6 / 6 */
            const int ti = (jl == 0 ? I.t0 : I.t1);
            evdw += 0.5f * __ldg(nbfp + ti + ti / P.ntypes).x * (1.0f / 6.0f) * P.lje_coeff6_6;
        }
    }
    const uint2* const emask = reinterpret_cast<const uint2*>(tmask) + (size_t)e * maxt;
    const int          nchunk = (ntile + NB_CHUNK - 1) / NB_CHUNK;
    int                t      = 0;
    for (int c = 0; c < nchunk; c++)
    {
        if (c + 1 < nchunk) stage(c + 1);
        else asm volatile("cp.async.commit_group;" ::: "memory"); /* keeps "all but the newest group" = chunk c landed */
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        /* load_j addresses tile t at s_lane + t*256: bias the base so that tile c*NB_CHUNK falls on this chunk's buffer */
        const unsigned char* const s_lane = sj + (c & 1) * CHUNK_BYTES + jl * 16 - c * CHUNK_BYTES;
        const int                  tend   = min(ntile, (c + 1) * NB_CHUNK);

        /* ---- tiles with exclusion masks (sorted to the front of the entry) ---- */
        for (const int tm = min(nmask, tend); t < tm; t++)
        {
            JAtom J;
            load_j(J, s_lane, t);
            const uint2 m = __ldg(emask + t);
            /* mask word w, bit `lane`: pair (i-atom 2*ih + w, j-slot jl) interacts */
            const float in0 = (float)((m.x >> lane) & 1u), in1 = (float)((m.y >> lane) & 1u);
            bool        ok0 = true, ok1 = true;
            if (intra && shift == B200NB_CENTRAL && (J.slot >> 3) == ci)
            {
                /* j-atom of the i-cluster itself: only j > i (nbnxm/pairlist.cpp:880-904, kernel_gpu_ref.cpp:223-226) */
                ok0 = (J.slot & 7) > 2 * ih;
                ok1 = (J.slot & 7) > 2 * ih + 1;
            }
            float2 tx, ty, tz;
            tile_pairs<EEL, GEOM, VF, true, GEN>(I, J, P, K, nbfp, nbfp_comb, in0, in1, ok0, ok1, tx, ty, tz, evdw, ecoul);
            fix = add2(fix, tx);
            fiy = add2(fiy, ty);
            fiz = add2(fiz, tz);
            reduce_store_j(tx, ty, tz, C, J.slot);
        }

        /* ---- plain tiles ---- */
        if (!VF && !GEN)
        {
            for (; t < tend; t++)
            {
                JAtom J[1];
                int   js[1];
                float sx[1], sy[1], sz[1];
                load_j(J[0], s_lane, t);
                js[0] = J[0].slot;
                tile_pairs_multi<EEL, GEOM, 1>(I, J, P, K, nbfp, fix, fiy, fiz, sx, sy, sz);
#ifndef B200NB_DIAG_NO_JFORCE /* diagnostic build only: no j-forces at all (wrong results), isolates the pair arithmetic */
                reduce_store_j_multi<1>(sx, sy, sz, C, js);
#endif
            }
        }
        for (; t < tend; t++)
        {
            JAtom J;
            load_j(J, s_lane, t);
            float2 tx, ty, tz;
            tile_pairs<EEL, GEOM, VF, false, GEN>(I, J, P, K, nbfp, nbfp_comb, 1.f, 1.f, true, true, tx, ty, tz, evdw, ecoul);
            fix = add2(fix, tx);
            fiy = add2(fiy, ty);
            fiz = add2(fiz, tz);
            reduce_store_j(tx, ty, tz, C, J.slot);
        }
        __syncwarp(); /* every lane is done with this buffer before the next iteration's gather overwrites it */
    }

    /* ---- i-forces: reduce over the 8 j-lanes (bits 0-2). Stage 1 is transposed: even lanes keep atom i0, odd lanes i1. */
    const bool     odd  = lane & 1;
    float          kx = odd ? fix.y : fix.x, ky = odd ? fiy.y : fiy.x, kz = odd ? fiz.y : fiz.x;
    const float    sx = odd ? fix.x : fix.y, sy = odd ? fiy.x : fiy.y, sz = odd ? fiz.x : fiz.y;
    kx += __shfl_xor_sync(full, sx, 1);
    ky += __shfl_xor_sync(full, sy, 1);
    kz += __shfl_xor_sync(full, sz, 1);
    kx += __shfl_xor_sync(full, kx, 2);
    ky += __shfl_xor_sync(full, ky, 2);
    kz += __shfl_xor_sync(full, kz, 2);
    kx += __shfl_xor_sync(full, kx, 4);
    ky += __shfl_xor_sync(full, ky, 4);
    kz += __shfl_xor_sync(full, kz, 4);
    /* lanes with jl in {0,1} hold the total force on i-atom 2*ih + jl */
    if (jl < 2 && ntile > 0) atomicAdd(f + ((size_t)ci * 8 + 2 * ih + jl), make_float4(kx, ky, kz, 0.f));
    if (VF)
    {
        /* shift force = sum of the i-forces of this entry (kernel_outer.h:620-640; the CUDA kernel skips the central
         * shift, nbnxm_cuda_kernel.cuh:624-628) */
        if (shift != B200NB_CENTRAL)
        {
            kx += __shfl_xor_sync(full, kx, 1);
            ky += __shfl_xor_sync(full, ky, 1);
            kz += __shfl_xor_sync(full, kz, 1);
            kx += __shfl_xor_sync(full, kx, 8);
            ky += __shfl_xor_sync(full, ky, 8);
            kz += __shfl_xor_sync(full, kz, 8);
            kx += __shfl_xor_sync(full, kx, 16);
            ky += __shfl_xor_sync(full, ky, 16);
            kz += __shfl_xor_sync(full, kz, 16);
            if (lane == 0)
            {
                float* fs = fshift + (int)(e & (NB_OUT_COPIES - 1)) * NB_FSHIFT_PITCH + 3 * shift;
                atomicAdd(fs, kx);
                atomicAdd(fs + 1, ky);
                atomicAdd(fs + 2, kz);
            }
        }
        for (int o = 16; o > 0; o >>= 1)
        {
            evdw += __shfl_xor_sync(full, evdw, o);
            ecoul += __shfl_xor_sync(full, ecoul, o);
        }
        if (lane == 0)
        {
            double* en = energy + 2 * (int)(e & (NB_OUT_COPIES - 1));
            atomicAdd(en, (double)evdw);
            atomicAdd(en + 1, (double)ecoul);
        }
    }
}

template<int EEL, bool GEOM, bool VF, bool GEN>
int launch(b200nb_context* h, const PackedList& L, int intra)
{
    /* One entry per single-warp CTA; CTAs start in index order, i.e. largest entries first.  A persistent variant (resident
     * warps walking the sorted list in boustrophedon order, equal tiles per warp) was measured and is 10-18 % slower: warps
     * that start equal entries together also wait for their prologue loads together
     * (profiles/r1/y_sweep_persistent_boustrophedon.txt). */
    const unsigned nblk = (unsigned)((L.nentries + B200NB_FORCE_WARPS - 1) / B200NB_FORCE_WARPS);
    const int      maxt = L.pitch;
    const size_t smem = (size_t)B200NB_FORCE_WARPS * 2 * NB_CHUNK * NB_TILE_SMEM; /* the two ring buffers of each warp */
    cudaLaunchConfig_t cfg{};
    cfg.gridDim          = dim3(nblk);
    cfg.blockDim         = dim3(32 * B200NB_FORCE_WARPS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1; /* overlap our list loads with the tail of the preceding kernel */
    cfg.attrs    = at;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_force<EEL, GEOM, VF, GEN>, (const Entry*)L.entries, (long long)L.nentries, (const int*)L.ja, (const uint64_t*)L.mask,
                       reinterpret_cast<const float4*>(h->d_xq), reinterpret_cast<const float2*>(h->d_lj), (const int*)h->d_atype,
                       reinterpret_cast<const float2*>(h->d_nbfp), (const float*)h->d_shift_vec, h->d_f, h->d_fshift, h->d_energy, h->dp,
                       intra, maxt, (const float*)h->d_kconst, reinterpret_cast<const float2*>(h->d_nbfp_comb));
    h->nlaunches++;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return nb_fail(h, B200NB_ERR_CUDA, std::string("force kernel launch: ") + cudaGetErrorString(err));
    return 0;
}

} // namespace

int nb_launch_force_kernel(b200nb_context* h, int loc, int flags)
{
    const PackedList& L = h->packed[loc];
    if (L.nentries == 0) return 0;
    const bool vf    = (flags & (B200NB_FLAG_ENERGY | B200NB_FLAG_VIRIAL)) != 0;
    const bool ewald = h->dp.eeltype == B200NB_EEL_EWALD;
    const int  intra = (loc == 0);
    /* the plain kernels cover LJ cut-off + potential shift with rvdw == rcoulomb (every BASELINE configuration); the
     * general ones add the force / potential switch and the twin-range check (cuda/nbnxm_cuda.cu:165-282 kernel table) */
    const bool gen = h->dp.vdw_modifier != B200NB_VDW_POTSHIFT || h->dp.rvdw2 < h->dp.rc2 || h->dp.ljpme != 0;
    /* LJ-PME reads its per-type grid parameters through the atom types, which only the type-table kernels stage */
    const bool geom = h->comb_geom && h->dp.ljpme == 0;
#define NB_PICK(E, G)                                                                                           \
    (gen ? (vf ? launch<E, G, true, true>(h, L, intra) : launch<E, G, false, true>(h, L, intra))               \
         : (vf ? launch<E, G, true, false>(h, L, intra) : launch<E, G, false, false>(h, L, intra)))
    if (ewald) return geom ? NB_PICK(1, true) : NB_PICK(1, false);
    return geom ? NB_PICK(0, true) : NB_PICK(0, false);
#undef NB_PICK
}
