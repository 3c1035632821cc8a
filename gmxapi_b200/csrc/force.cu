/* b200nb force kernels: Lennard-Jones + {reaction-field / plain cut-off, Ewald real space (analytical or tabulated)}
 * cluster-pair force and energy with shift-force reduction, hand-written for sm_100a.
 *
 * Replaces (paths relative to /root/reference/src/gromacs):
 *   nbnxm/cuda/nbnxm_cuda_kernel.cuh:150-650 (K1) with its reductions nbnxm_cuda_kernel_utils.cuh:477-700;
 *   arithmetic parity target = the CPU SIMD kernels nbnxm/kernels_simd_2xmm/kernel_inner.h:226-880 and
 *   kernel_outer.h:395-452 (self terms), simd/simd_math.h:1609-1722 (Ewald correction polynomials).
 *
 * Design (not a port of the reference CUDA kernel), sized with profiles/tools/microbench.cu, the diagnostic builds of
 * profiles/r1/v_sweep_pair_loop_diagnostics.txt and the ncu source pages under profiles/: on sm_100 the FP32 pipe retires
 * 128 lane-FMAs/clk/SM whether issued as scalar FFMA or packed FFMA2; a packed instruction keeps the pipe busy for two
 * cycles, everything else competes for the remaining issue slots.
 *  - the unit of the list is a HALF-ENTRY: four i-atoms (one half of an 8-atom i-cluster) + shift against a run of PACKED
 *    j-atoms (PackedList, b200nb_internal.h) -- the j-atoms that have at least one of their FOUR pairs inside the list radius.
 *    Packing per i-quad instead of per i-cluster raises the share of in-range lanes from 54 % to 65 %: the one lever that
 *    removes FP32 work instead of overhead;
 *  - one warp runs TWO half-entries side by side, lanes 0-15 the first, lanes 16-31 the second; they are neighbours in the
 *    size-sorted list, so they have the same number of steps (a step = 16 j-atoms per half-entry) up to the few warps that
 *    straddle a size boundary; a half that runs out first continues on far-away dummy atoms;
 *  - lane = jl + 16*ih owns ONE j-atom per step (jl of its half-entry's list) and FOUR i-atoms as two packed float2 pairs
 *    held in registers for the whole entry.  All pair arithmetic is packed (fma.rn.f32x2 -> FFMA2/FMUL2), the j operands
 *    enter in the scalar-broadcast operand form;
 *  - the j-atom's data (16-byte xyzq, 8-byte LJ pair) are gathered straight into the lane's registers ONE STEP AHEAD of their use,
 *    the slot index two steps ahead: two loads + one index load per step, no shared memory.  What makes this work is where the
 *    loads sit: at the top of the step, pinned there by the (deferred) j-force red of the previous step (see load_j).  The
 *    round-2 kernels before this one staged the stream through a 4-stage per-lane cp.async ring in shared memory (three steps
 *    of lead): ncu showed the L1TEX data pipe at 86 % with it (cp.async fills, record reads, the slot hand-over) and the
 *    register form is 2-5 % faster at every size (profiles/r2/s_sweep_register_gather.txt); prefetch instructions (CCTL) for a
 *    longer lead cost 10-15 % and were dropped (u_, w_sweep_*.txt);
 *  - j-forces: accumulated in the lane over its 4 pairs with the same packed FMAs that feed the i accumulators and
 *    reduced by the lane itself with one 16-byte red (see red_j): the two halves of a warp hold different j-atoms, so there is
 *    no exchange (the cluster-granular layout needed 2 shuffles + 5 selects / negations per step to share a j-atom between halves);
 *  - i-forces stay in registers (12 floats per lane); per half-entry one transposed butterfly over its 16 lanes
 *    (15 shuffles) and one 16-byte red per i-atom;
 *  - steps that carry exclusion masks are sorted to the front of a half-entry (k_pack) and run through a separate code
 *    path; the unmasked path has no mask logic and no r^2 clamp;
 *  - LJ is evaluated as (c12*r^-6 - c6)*r^-6, the Ewald correction polynomials keep their coefficients as
 *    instruction immediates; out-of-range lanes are discarded by select, so garbage there cannot poison a sum;
 *  - r^2 is evaluated with the reference's operand roles and operation order so the in-range pair set is
 *    bit-identical (see nb_rsq in b200nb_internal.h);
 *  - TMA (cp.async.bulk / tile::gather4) is not used: the j stream is a gather of single 16-byte atoms by slot index
 *    whose consumer is ONE lane; a per-lane load delivers it into that lane's registers with no barrier, no shared memory
 *    and no elected-thread issue; gather4 moves rows of a 2-D tensor into shared memory for a whole CTA and needs an mbarrier
 *    round trip per stage plus a shared-memory read per lane;
 *  - LJ force switch / potential switch / VdW cut-off below the Coulomb cut-off / LJ-PME are a second set of
 *    instantiations (GEN): the plain kernels, which every BASELINE configuration uses, pay nothing for them.
 * Measured and dropped (profiles/r2/k_*): a persistent form with cross-entry pipelining (k_force_p) --
 * it removed the per-entry latency and was 25-30 % slower by its extra bookkeeping instructions.
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>

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
    float2 lj;   /* GEOM: sqrt(6 C6), sqrt(12 C12); table: atom type in lj.x (as int bits) */
    int    slot; /* grid slot of the j-atom: index into xq / f */
};

#define NB_JSTEP 16 /* j-atoms per step of a half-entry: lane jl + 16*ih holds j-atom jl of the step of half-entry ih of the warp */
/* The j-atom gather into registers.  Plain (weak) global loads in volatile asm with a memory clobber, NOT the read-only path:
 * the loads of step s + 1 are issued at the TOP of step s, in front of the red that flushes the j-forces of step s - 1, and a
 * weak load may not be moved across a later reduction to an address ptxas cannot tell apart -- which pins the loads a whole
 * step ahead of their first use.  (An ld.global.nc placed there was sunk by ptxas to the end of the step, right in front of
 * its consumer: 41 % of the stall samples of profiles/r2/a_ncu_source_k_force_water192k.txt.) */
template<bool GEOM>
__device__ __forceinline__ void load_j(JAtom& J, int slot, const float4* __restrict__ xq, const float2* __restrict__ lj, const int* __restrict__ atype)
{
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(J.xq.x), "=f"(J.xq.y), "=f"(J.xq.z), "=f"(J.xq.w) : "l"(xq + slot) : "memory");
#ifdef B200NB_DIAG_NO_LJ /* diagnostic build only (wrong results): what the second gather instruction of a step costs */
    if (GEOM) J.lj = make_float2(0.05f, 0.003f);
    else
#endif
    if (GEOM) asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(J.lj.x), "=f"(J.lj.y) : "l"(lj + slot) : "memory");
    else asm volatile("ld.global.f32 %0, [%1];" : "=f"(J.lj.x) : "l"(atype + slot) : "memory");
    J.slot = slot;
}
__device__ __forceinline__ int load_index(const int* p)
{
    int v;
#ifdef B200NB_INDEX_EVICT_FIRST /* experiment: the slot-index stream (read once) marked evict-first in L2, so that it does not push out xq / f */
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
#else
    asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#endif
    return v;
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
    float2 x, y, z;  /* two i-atoms of this lane, shift already added */
    float2 q;        /* epsfac * q_i */
    float2 c6n, c12; /* GEOM: -sqrt(6 C6_i), sqrt(12 C12_i) */
    int    t0, t1;   /* table path: type_i * ntypes */
    float2 g0, g1;   /* LJ-PME (GEN kernels): nbfp_comb of the two i-atoms' types */
};

/* Two pairs: (i0, j) in .x and (i1, j) in .y.  Returns F/r (zero outside the cut-off or for a switched-off pair) and the
 * distance vector xi - xj; adds the pair energies to evdw / ecoul when VF. */
template<int EEL, bool GEOM, bool VF, bool MASKED, bool GEN>
__device__ __forceinline__ float2 pair_fscal(const IData& I, const JAtom& J, const NbParamsDev& P, const KConst& K,
                                             const float2* __restrict__ nbfp, const float2* __restrict__ nbfp_comb, float inter0, float inter1,
                                             bool ok0, bool ok1, float2& dx, float2& dy, float2& dz, float& evdw, float& ecoul)
{
    const float2 m1 = dup(-1.0f);
    dx              = fma2(dup(J.xq.x), m1, I.x); /* xi - xj: the product is exact */
    dy              = fma2(dup(J.xq.y), m1, I.y);
    dz              = fma2(dup(J.xq.z), m1, I.z);
    float2 r2       = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
    bool   wa = r2.x < K.rc2, wb = r2.y < K.rc2;
    float2 inter;
    if (MASKED)
    {
        wa    = wa && ok0;
        wb    = wb && ok1;
        inter = make_float2(inter0, inter1);
        r2    = make_float2(fmaxf(r2.x, NB_MIN_RSQ), fmaxf(r2.y, NB_MIN_RSQ));
    }
    const float2 rinv    = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
    float2       c6n, c12;
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
    const float2 qq = mul2(I.q, dup(J.xq.w));
    /* Ewald correction polynomials first: their reciprocal (MUFU) then overlaps the LJ arithmetic */
    float2 z2 = dup(0.0f), ewt = dup(0.0f);
    if (EEL == 1)
    {
        z2               = mul2(dup(K.beta2), r2);
        const float2 z4  = mul2(z2, z2);
        const float2 den = pme_force_den(z2, z4, K.fd4, K.fd3, K.fd2, K.fd1, K.fd0);
        const float2 rd  = make_float2(rcp_approx(den.x), rcp_approx(den.y));
        const float2 num = pme_force_num(z2, z4, K.fn6, K.fn5);
        ewt              = mul2(num, rd); /* beta * pmecorrF(z2) */
    }
    if (EEL == 2)
    {
        /* tabulated Ewald correction (the reference's EL_EWALD_TAB kernels: interpolate_coulomb_force_r,
         * cuda/nbnxm_cuda_kernel_utils.cuh:380-391; kernel_gpu_ref.cpp:262-272): F(r) interpolated linearly between the points of
         * interaction_const_t::coulombEwaldTables->tableF; the table holds {F[i], F[i+1] - F[i]} so one 8-byte load serves a pair.
         * ewt becomes -F_table(r) * r, to be added to int_bit / r like the analytical term. */
        if (VF) z2 = mul2(dup(K.beta2), r2);
        const float2 r  = mul2(r2, rinv);
        float2       rt = mul2(r, dup(P.tab_scale));
        rt              = make_float2(fminf(rt.x, P.tab_max), fminf(rt.y, P.tab_max)); /* lanes beyond the cut-off stay inside the table */
        const int    i0 = (int)rt.x, i1 = (int)rt.y;
        const float2 t0 = __ldg(P.ewald_tab + i0), t1 = __ldg(P.ewald_tab + i1);
        const float2 ft = make_float2(__fmaf_rn(rt.x - (float)i0, t0.y, t0.x), __fmaf_rn(rt.y - (float)i1, t1.y, t1.x));
        ewt             = mul2(mul2(ft, r), m1);
    }
    const float2 rinvsq  = mul2(rinv, rinv);
    float2       rinv_ex = rinv;
    if (MASKED) rinv_ex = mul2(rinv, inter);
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
    float2 vcoul = dup(0.0f);
    if (EEL >= 1)
    {
        if (EEL == 1) fsum = fma2(qq, fma2(ewt, z2, rinv_ex), fsum);
        else fsum = fma2(qq, add2(ewt, rinv_ex), fsum);
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
    return fscal;
}

/* j-force of a step: the lane's packed accumulators hold, per component, the sums over its even (.x) and odd (.y) i-atoms
 * of F/r * d (the force ON THE i-ATOMS); the force on the lane's j-atom is minus their sum.  Every lane has its own j-atom, so
 * it reduces all three components itself with ONE 16-byte red into the j-atom's float4 force slot.  Measured
 * (profiles/r2/q_sweep_half_entries.txt, zl_sweep_jred_split_register_gather.txt): red.v2 {x, y} + red {z} saves the fourth, useless
 * add in L2 but doubles the requests on the SM -> L2 reduction path, which saturates at ~240 G lane-level reductions/s: 487 us
 * instead of 403 us at 1 M atoms, for Ewald and reaction field alike (B200NB_JRED_SPLIT builds it). */
__device__ __forceinline__ void red_j(const float sx, const float sy, const float sz, float4* __restrict__ f, const int jslot)
{
    float* const fp = reinterpret_cast<float*>(f + jslot);
#ifdef B200NB_DIAG_NO_RED /* diagnostic build only: drops the j-force scatter (wrong results) to measure its cost */
    if (sx == 12345.678f)
#endif
    {
#ifndef B200NB_JRED_SPLIT
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(fp), "f"(sx), "f"(sy), "f"(sz), "f"(0.0f) : "memory");
#else
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(fp), "f"(sx), "f"(sy) : "memory");
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(fp + 2), "f"(sz) : "memory");
#endif
    }
}
#ifndef B200NB_FORCE_WARPS
#define B200NB_FORCE_WARPS 1 /* warps (= list entries) per CTA: small CTAs refill an SM's warp slots at entry granularity */
#endif
#ifndef B200NB_MB_PLAIN
#define B200NB_MB_PLAIN 20
#endif
/* resident warps per SM the register allocation aims at: 20 (<= 96 registers) for the force-only kernels, 16 for the energy /
 * modifier kernels, which need up to 128 registers */
#define B200NB_FORCE_MIN_BLOCKS(VF, GEN) (((VF) || (GEN) ? 16 : B200NB_MB_PLAIN) / B200NB_FORCE_WARPS)
template<int EEL, bool GEOM, bool VF, bool GEN>
__global__ void __launch_bounds__(32 * B200NB_FORCE_WARPS, B200NB_FORCE_MIN_BLOCKS(VF, GEN))
k_force(const Entry* __restrict__ entries, long long nwarps, const int* __restrict__ pja, const uint64_t* __restrict__ tmask,
        const float4* __restrict__ xq, const float2* __restrict__ lj, const int* __restrict__ atype, const float2* __restrict__ nbfp,
        const float* __restrict__ shift_vec, float4* __restrict__ f, float* __restrict__ fshift, double* __restrict__ energy,
        const __grid_constant__ NbParamsDev P, const int intra, const float* __restrict__ kconst,
        const float2* __restrict__ nbfp_comb)
{
    const long long w = (long long)blockIdx.x * B200NB_FORCE_WARPS + (threadIdx.x >> 5);
    if (w >= nwarps) return;
    const int      lane = threadIdx.x & 31, jl = lane & (NB_JSTEP - 1), ih = lane >> 4;
    const unsigned full = 0xffffffffu;
    /* Warp w runs the half-entries 2w (lanes 0-15) and 2w + 1 (lanes 16-31) of the size-sorted list.  Both headers are loaded
     * by every lane (two broadcast loads from warp-uniform addresses), so that the step counts bounding the loops are provably
     * warp-uniform for the compiler (uniform datapath, no divergence bookkeeping around the loop bodies); everything else that
     * derives from a header is a per-HALF value picked by ih. */
    const int4 evA = __ldg(reinterpret_cast<const int4*>(entries) + 2 * w), evB = __ldg(reinterpret_cast<const int4*>(entries) + 2 * w + 1);
    const int4 ev  = ih ? evB : evA;
    KConst K;
    {
        const float4 k0 = __ldg(reinterpret_cast<const float4*>(kconst)), k1 = __ldg(reinterpret_cast<const float4*>(kconst) + 1),
                     k2 = __ldg(reinterpret_cast<const float4*>(kconst) + 2);
        K.rc2 = k0.x, K.beta2 = k0.z, K.fd4 = k0.w, K.fd3 = k1.x, K.fn6 = k1.y, K.fn5 = k1.z, K.fd2 = k1.w, K.fd1 = k2.x, K.fd0 = k2.y;
    }
    /* the half-entry owns the steps [ev.z, ev.w) of the j-slot array (16 slots per step) and of the mask array (64 bits per step) */
    const int nstep_my = ev.w - ev.z;
    const int nstep    = max(evA.w - evA.z, evB.w - evB.z);
    /* this lane's slot of step k is ja[16 k].  A half-entry that is shorter than the one sharing its warp runs the difference on
     * far-away dummy atoms: k_pad_partner (b200nb.cu) has written their slots behind its own steps, so the loops below need no
     * per-half bounds -- only the mask fetch and the outputs look at nstep_my */
    const int* const ja = pja + (size_t)ev.z * NB_JSTEP + jl;
    /* the j slots of the first two steps: list data only, so they may be fetched before the dependency wait below */
    int slot_n = __ldg(ja), slot_nn = __ldg(ja + NB_JSTEP); /* always valid: a row ends in NB_PACK_TAIL steps of dummy atoms */
    /* Programmatic dependent launch: everything above reads only the list; from here on the kernel touches xq and f, so wait
     * for the completion of the preceding kernel of the stream (k_step_begin) -- a no-op for a normally serialised launch. */
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool self = VF && NB_ENTRY_SELF(ev.y);
    if (nstep == 0 && !(VF && (NB_ENTRY_SELF(evA.y) || NB_ENTRY_SELF(evB.y)))) return;
    const int  ci = ev.x, shift = NB_ENTRY_SHIFT(ev.y), half = NB_ENTRY_HALF(ev.y);
    const int  nmask_my = min(NB_ENTRY_NMASK(ev.y), nstep_my);
    const int  nmask    = max(min(NB_ENTRY_NMASK(evA.y), evA.w - evA.z), min(NB_ENTRY_NMASK(evB.y), evB.w - evB.z));
    JAtom Jn; /* the j-atom of the NEXT step, in flight */
    Jn.xq = make_float4(0.f, 0.f, 0.f, 0.f), Jn.lj = dup(0.f), Jn.slot = 0;
    load_j<GEOM>(Jn, slot_n, xq, lj, atype);
    IData I[2];
    {
        const float sx = __ldg(shift_vec + 3 * shift), sy = __ldg(shift_vec + 3 * shift + 1), sz = __ldg(shift_vec + 3 * shift + 2);
#pragma unroll
        for (int p = 0; p < 2; p++)
        {
            const size_t ia = (size_t)ci * 8 + 4 * half + 2 * p;
            const float4 a = __ldg(xq + ia), b = __ldg(xq + ia + 1);
            /* the reference adds the shift to the i-atom before the subtraction: kernel_outer.h:482-489 */
            I[p].x = make_float2(__fadd_rn(a.x, sx), __fadd_rn(b.x, sx));
            I[p].y = make_float2(__fadd_rn(a.y, sy), __fadd_rn(b.y, sy));
            I[p].z = make_float2(__fadd_rn(a.z, sz), __fadd_rn(b.z, sz));
            I[p].q = make_float2(P.epsfac * a.w, P.epsfac * b.w);
            I[p].g0 = I[p].g1 = dup(0.0f);
            if (GEOM)
            {
                const float4 l = __ldg(reinterpret_cast<const float4*>(lj + ia));
                I[p].c6n       = make_float2(-l.x, -l.z);
                I[p].c12       = make_float2(l.y, l.w);
                I[p].t0 = I[p].t1 = 0;
            }
            else
            {
                const int2 t = __ldg(reinterpret_cast<const int2*>(atype + ia));
                I[p].t0      = t.x * P.ntypes;
                I[p].t1      = t.y * P.ntypes;
                I[p].c6n = I[p].c12 = dup(0.0f);
                if (GEN && P.ljpme)
                {
                    I[p].g0 = __ldg(nbfp_comb + t.x);
                    I[p].g1 = __ldg(nbfp_comb + t.y);
                }
            }
        }
    }
#define NB_LOAD_I(IP, p) const IData& IP = I[p];
    float2 fix[2] = { dup(0.f), dup(0.f) }, fiy[2] = { dup(0.f), dup(0.f) }, fiz[2] = { dup(0.f), dup(0.f) };
    float  evdw = 0.f, ecoul = 0.f;
    if (self && jl < 4)
    {
        /* Coulomb self term, once per i-atom (lane jl of the half-entry takes i-atom 4*half + jl): kernel_outer.h:408-452
         * (fillers carry q = 0) */
        const float2 qp = (jl & 2) ? I[1].q : I[0].q;
        const float  qi = (jl & 1) ? qp.y : qp.x;
        ecoul -= qi * qi * P.self_q2;
        if (GEN && !GEOM && P.ljpme)
        {
            /* LJ Ewald self interaction, kernel_ref_outer.h:316-321: 0.5 * (6 C6_ii) / 6 * coeff^6 / 6 */
            const int ta = (jl & 2) ? I[1].t0 : I[0].t0, tb = (jl & 2) ? I[1].t1 : I[0].t1;
            const int ti = (jl & 1) ? tb : ta;
            evdw += 0.5f * __ldg(nbfp + ti + ti / P.ntypes).x * (1.0f / 6.0f) * P.lje_coeff6_6;
        }
    }
    /* masks of the leading steps: 64 bits per step, bit 16*k + jl = pair (i-atom 4*half + k, j-atom jl) interacts */
    const uint2* const emask = reinterpret_cast<const uint2*>(tmask) + ev.z;
    int                s     = 0;
    /* Top of step s: the registers Jn hold (or are about to receive) the j-atom of step s, slot_nn the slot of step s + 1.
     * next_j() hands Jn over as the current j-atom, issues the gather of step s + 1 and the fetch of the slot of step s + 2, and
     * then flushes the j-force of step s - 1 (pending in pjx .. pslot): that red is what keeps the loads above it. */
    const int* jp = ja + 2 * NB_JSTEP; /* slot index of step s + 2 */
    float      pjx = 0.f, pjy = 0.f, pjz = 0.f;
    int        pslot = slot_n; /* the first flush adds zeros to the first j-atom's slot */
    auto next_j = [&](JAtom& J) {
        J = Jn;
        /* unconditional: past the last step these fetch the dummy atoms that end every row (NB_PACK_TAIL), so that the loop has
         * no predicates and the registers of Jn are simply renamed from step to step */
        load_j<GEOM>(Jn, slot_nn, xq, lj, atype);
        slot_nn = load_index(jp);
        jp += NB_JSTEP;
        red_j(pjx, pjy, pjz, f, pslot);
    };
#define NB_STORE_J(fjx, fjy, fjz, slot) pjx = -(fjx).x - (fjx).y, pjy = -(fjy).x - (fjy).y, pjz = -(fjz).x - (fjz).y, pslot = (slot)

    /* ---- steps with exclusion masks (sorted to the front of a half-entry; one or two per half-entry) ---- */
    for (; s < nmask; s++)
    {
        JAtom J;
        next_j(J);
        uint2 m = make_uint2(~0u, ~0u);
        if (s < nmask_my) m = __ldg(emask + s);
        /* j-atom of the i-cluster itself: only j > i (nbnxm/pairlist.cpp:880-904, kernel_gpu_ref.cpp:223-226) */
        const bool diag = intra && shift == B200NB_CENTRAL && (J.slot >> 3) == ci;
        const int  jin  = J.slot & 7;
        float2     fjx = dup(0.f), fjy = dup(0.f), fjz = dup(0.f);
#pragma unroll
        for (int p = 0; p < 2; p++)
        {
            const unsigned wd  = p ? m.y : m.x;
            const float    in0 = (float)((wd >> jl) & 1u), in1 = (float)((wd >> (16 + jl)) & 1u);
            const bool     ok0 = !diag || jin > 4 * half + 2 * p, ok1 = !diag || jin > 4 * half + 2 * p + 1;
            float2         dx, dy, dz;
            NB_LOAD_I(Ip, p)
            const float2   fs = pair_fscal<EEL, GEOM, VF, true, GEN>(Ip, J, P, K, nbfp, nbfp_comb, in0, in1, ok0, ok1, dx, dy, dz, evdw, ecoul);
            fix[p] = fma2(fs, dx, fix[p]);
            fiy[p] = fma2(fs, dy, fiy[p]);
            fiz[p] = fma2(fs, dz, fiz[p]);
            fjx    = fma2(fs, dx, fjx);
            fjy    = fma2(fs, dy, fjy);
            fjz    = fma2(fs, dz, fjz);
        }
        NB_STORE_J(fjx, fjy, fjz, J.slot);
    }
    /* ---- plain steps ---- */
#pragma unroll 2
    for (; s < nstep; s++)
    {
        JAtom J;
        next_j(J);
        float2 fjx = dup(0.f), fjy = dup(0.f), fjz = dup(0.f);
#pragma unroll
        for (int p = 0; p < 2; p++)
        {
            float2       dx, dy, dz;
            NB_LOAD_I(Ip, p)
            const float2 fs = pair_fscal<EEL, GEOM, VF, false, GEN>(Ip, J, P, K, nbfp, nbfp_comb, 1.f, 1.f, true, true, dx, dy, dz, evdw, ecoul);
            fix[p] = fma2(fs, dx, fix[p]);
            fiy[p] = fma2(fs, dy, fiy[p]);
            fiz[p] = fma2(fs, dz, fiz[p]);
#ifndef B200NB_DIAG_NO_JFORCE /* diagnostic build only: no j-forces at all (wrong results), isolates the pair arithmetic */
            fjx = fma2(fs, dx, fjx);
            fjy = fma2(fs, dy, fjy);
            fjz = fma2(fs, dz, fjz);
#endif
        }
#ifndef B200NB_DIAG_NO_JFORCE
        NB_STORE_J(fjx, fjy, fjz, J.slot);
#endif
    }
    if (nstep > 0) red_j(pjx, pjy, pjz, f, pslot); /* the j-force of the last step */
#undef NB_STORE_J
    /* ---- i-forces: 12 floats per lane (2 pairs x 2 atoms x 3) summed over the 16 j-lanes of this half (lane bits 0-3),
     * transposed so that every exchange halves what is still carried: bit 3 picks the pair, bit 2 the atom of the pair ---- */
    const bool b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
    float      r0, r1, r2, r3, r4, r5;
#define NB_STAGE_A(out, v0, v1)                                          \
    {                                                                    \
        const float keep_ = b3 ? (v1) : (v0), send_ = b3 ? (v0) : (v1);  \
        out               = keep_ + __shfl_xor_sync(full, send_, 8);     \
    }
    NB_STAGE_A(r0, fix[0].x, fix[1].x)
    NB_STAGE_A(r1, fix[0].y, fix[1].y)
    NB_STAGE_A(r2, fiy[0].x, fiy[1].x)
    NB_STAGE_A(r3, fiy[0].y, fiy[1].y)
    NB_STAGE_A(r4, fiz[0].x, fiz[1].x)
    NB_STAGE_A(r5, fiz[0].y, fiz[1].y)
#undef NB_STAGE_A
    float kx = (b2 ? r1 : r0) + __shfl_xor_sync(full, b2 ? r0 : r1, 4);
    float ky = (b2 ? r3 : r2) + __shfl_xor_sync(full, b2 ? r2 : r3, 4);
    float kz = (b2 ? r5 : r4) + __shfl_xor_sync(full, b2 ? r4 : r5, 4);
    kx += __shfl_xor_sync(full, kx, 2);
    ky += __shfl_xor_sync(full, ky, 2);
    kz += __shfl_xor_sync(full, kz, 2);
    kx += __shfl_xor_sync(full, kx, 1);
    ky += __shfl_xor_sync(full, ky, 1);
    kz += __shfl_xor_sync(full, kz, 1);
    /* the lanes with bits 0-1 clear hold the total force on i-atom 4*half + 2*b3 + b2 of their half-entry */
    if ((lane & 3) == 0 && nstep_my > 0) atomicAdd(f + ((size_t)ci * 8 + 4 * half + 2 * (int)b3 + (int)b2), make_float4(kx, ky, kz, 0.f));
    if (VF)
    {
        /* shift force = sum of the i-forces of a half-entry (kernel_outer.h:620-640; the CUDA kernel skips the central
         * shift, nbnxm_cuda_kernel.cuh:624-628): summed over the 16 lanes of the half, the two halves may carry different shifts */
        kx += __shfl_xor_sync(full, kx, 4);
        ky += __shfl_xor_sync(full, ky, 4);
        kz += __shfl_xor_sync(full, kz, 4);
        kx += __shfl_xor_sync(full, kx, 8);
        ky += __shfl_xor_sync(full, ky, 8);
        kz += __shfl_xor_sync(full, kz, 8);
        if (jl == 0 && shift != B200NB_CENTRAL && nstep_my > 0)
        {
            float* fs = fshift + (int)(w & (NB_OUT_COPIES - 1)) * NB_FSHIFT_PITCH + 3 * shift;
            atomicAdd(fs, kx);
            atomicAdd(fs + 1, ky);
            atomicAdd(fs + 2, kz);
        }
        for (int o = 16; o > 0; o >>= 1)
        {
            evdw += __shfl_xor_sync(full, evdw, o);
            ecoul += __shfl_xor_sync(full, ecoul, o);
        }
        if (lane == 0)
        {
            double* en = energy + 2 * (int)(w & (NB_OUT_COPIES - 1));
            atomicAdd(en, (double)evdw);
            atomicAdd(en + 1, (double)ecoul);
        }
    }
}

template<int EEL, bool GEOM, bool VF, bool GEN>
int launch(b200nb_context* h, const PackedList& L, int intra)
{
    cudaLaunchConfig_t cfg{};
    cfg.blockDim         = dim3(32 * B200NB_FORCE_WARPS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream           = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1; /* overlap our list loads with the tail of the preceding kernel */
    cfg.attrs    = at;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    /* two half-entries per single-warp CTA; CTAs start in index order, i.e. largest half-entries first */
    const long long nwarps = L.nentries / 2;
    cfg.gridDim            = dim3((unsigned)((nwarps + B200NB_FORCE_WARPS - 1) / B200NB_FORCE_WARPS));
    cudaLaunchKernelEx(&cfg, k_force<EEL, GEOM, VF, GEN>, (const Entry*)L.entries, nwarps, (const int*)L.ja, (const uint64_t*)L.mask,
                       reinterpret_cast<const float4*>(h->d_xq), reinterpret_cast<const float2*>(h->d_lj), (const int*)h->d_atype,
                       reinterpret_cast<const float2*>(h->d_nbfp), (const float*)h->d_shift_vec, h->d_f, h->d_fshift, h->d_energy, h->dp,
                       intra, (const float*)h->d_kconst, reinterpret_cast<const float2*>(h->d_nbfp_comb));
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
    /* tabulated Ewald correction (b200nb_set_ewald_table; the reference's EL_EWALD_TAB choice, nbnxm_gpu_data_mgmt.cpp:118-154):
     * plain kernels only -- the reference's CUDA backend has no tabulated kernel with LJ-PME either way round that matters here */
#define NB_PICK_PLAIN(E, G) (vf ? launch<E, G, true, false>(h, L, intra) : launch<E, G, false, false>(h, L, intra))
    if (ewald && h->dp.ewald_tab != nullptr && !gen) return geom ? NB_PICK_PLAIN(2, true) : NB_PICK_PLAIN(2, false);
#undef NB_PICK_PLAIN
    if (ewald) return geom ? NB_PICK(1, true) : NB_PICK(1, false);
    return geom ? NB_PICK(0, true) : NB_PICK(0, false);
#undef NB_PICK
}
