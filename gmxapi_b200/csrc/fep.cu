/* b200nb: the perturbed-pair (free-energy) kernel on the device.
 *
 * Replaces CPU code of the reference: gmxlib/nonbonded/nb_free_energy.cpp:203-860 (nb_free_energy_kernel, the scalar instantiation:
 * the reference has no SIMD or GPU form of it and runs it on the host beside the GPU nonbonded kernels, mdlib/sim_util.cpp:1658
 * do_force -> nonbonded_verlet_t::dispatchFreeEnergyKernel, nbnxm/kerneldispatch.cpp:458-567), for the flavours built so far:
 * reaction-field / plain cut-off or Ewald electrostatics (the long-range part subtracted unsoftened, :693-737, evaluated directly
 * instead of from the reference's spline table), cut-off LJ with potential shift, soft-core with r-power 6 (lambda power 1 or 2) or
 * none, LJ potential shift or potential switch, rvdw <= rcoulomb, LJ-PME (potential shift; the grid part subtracted unsoftened, :725-770, evaluated
 * directly as well).  The LJ force switch (which the reference's kernel does not have either) is refused.
 * The pair list comes from the caller in t_nblist form (mdtypes/nblist.h:117-137; b200nb_fep_upload_list) -- what
 * nbnxm/pairlist.cpp:1699-1872 make_fep_list produces: every pair within the list radius with a perturbed atom, excluded pairs
 * flagged, perturbed atoms listed with themselves -- or is built on the device from the gridded coordinates (b200nb_fep_build_list).
 * As in the reference the perturbed atoms carry zero charge and the filler LJ type in the normal atom data
 * (nbnxn_atomdata_mask_fep, nbnxm/atomdata.cpp), so the cluster-pair kernels compute nothing for them; this kernel adds the
 * forces of the perturbed pairs into the same grid-order force buffer and the same shift-force replicas.
 *
 * One warp per i-entry (one i-atom + shift, a handful to a few dozen j-atoms): lanes take the j-atoms, the i-force / energies /
 * dV/dlambda are reduced over the warp.  Not a throughput kernel: the perturbed pairs are ~1e-3 of the pairs of a typical system.
 */
#include <cstdio>
#include <vector>

#include "b200nb_internal.h"

namespace
{

struct FepDev
{
    float rc, rc2, epsfac, k_rf, c_rf, disp_cpot, rep_cpot;
    float rvdw; /* <= rc = rcoulomb (shorter where PME load balancing grew rcoulomb); the list cut-off is the larger one, :300-301 */
    float LFC[2], LFV[2];
    float lfac_coul[2], dlfac_coul[2], lfac_vdw[2], dlfac_vdw[2];
    float alpha_coul, alpha_vdw, sigma6_def, sigma6_min;
    int   soft_core, sc_differ, ntypes;
    int   ewald;
    float beta, sh_ewald;
    int   pot_switch; /* LJ potential switch from rvdw_switch to rc (interaction_const_t::vdw_switch) */
    float rvdw_switch, sw_c3, sw_c4, sw_c5;
    int   ljpme; /* 0 none, 1 geometric, 2 Lorentz-Berthelot grid combination rule (vdwtype = evdwPME) */
    float lje_coeff2, sh_lj_ewald; /* ewaldcoeff_lj^2, interaction_const_t::sh_lj_ewald */
};

/* the grid C6 of a type pair, times six like nbfp's C6 (fr->ljpme_c6grid, mdlib/forcerec.cpp:157-195), from the per-type
 * nbfp_comb entries the cluster-pair kernels use (b200nb_set_vdw).  One corner differs from make_ljpme_c6grid: for a type with
 * C6 > 0 and C12 = 0 the Lorentz-Berthelot rule has no sigma, the reference's table falls back to sqrt(C6_i C6_j) there while its
 * cluster kernels' nbfp_comb (nbnxm/atomdata.cpp:303-321) -- and so this function -- give such a type no grid C6 at all */
__device__ __forceinline__ float grid_c6(int rule, float2 gi, float2 gj)
{
    if (rule == 1) return gi.x * gj.x;
    const float sg = gi.x + gj.x, s2 = sg * sg;
    return gi.y * gj.y * (s2 * s2 * s2);
}

/* r^(1/6) of 1 / (alpha sigma^6 + r^6) and its inverse: pthRoot, nb_free_energy.cpp:81-87 */
__device__ __forceinline__ void sixth_root(float rpinv, float& rinv_eff, float& r_eff)
{
    r_eff    = rsqrtf(cbrtf(rpinv));
    rinv_eff = 1.0f / r_eff;
}

__global__ void __launch_bounds__(128)
k_fep(int nri, const int* __restrict__ iinr, const int* __restrict__ shift, const int* __restrict__ jindex, const int* __restrict__ jjnr,
      const signed char* __restrict__ excl_fep, const float4* __restrict__ xq, const int* __restrict__ slot_of_atom, const float* __restrict__ shift_vec,
      const int* __restrict__ typeA, const int* __restrict__ typeB, const float* __restrict__ qA, const float* __restrict__ qB,
      const float2* __restrict__ nbfp, const float2* __restrict__ nbfp_comb, const __grid_constant__ FepDev P, float4* __restrict__ f, float* __restrict__ fshift, double* __restrict__ out4)
{
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= nri) return;
    const int    ii = iinr[n], is = shift[n];
    const int    si = slot_of_atom[ii];
    const float4 xi = xq[si];
    const float  ix = shift_vec[3 * is] + xi.x, iy = shift_vec[3 * is + 1] + xi.y, iz = shift_vec[3 * is + 2] + xi.z;
    const float  iqA = P.epsfac * qA[ii], iqB = P.epsfac * qB[ii];
    const int    tiA = typeA[ii] * P.ntypes, tiB = typeB[ii] * P.ntypes;
    float2       giA = make_float2(0.f, 0.f), giB = giA;
    if (P.ljpme) giA = nbfp_comb[typeA[ii]], giB = nbfp_comb[typeB[ii]];
    const float  DLF[2] = { -1.f, 1.f };
    float        vctot = 0, vvtot = 0, fix = 0, fiy = 0, fiz = 0, dvdl_coul = 0, dvdl_vdw = 0;
    int          nwithin = 0;
    for (int k = jindex[n] + lane; k < jindex[n + 1]; k += 32)
    {
        const int    jnr = jjnr[k], sj = slot_of_atom[jnr];
        const float4 xj  = xq[sj];
        const float  dx = ix - xj.x, dy = iy - xj.y, dz = iz - xj.z;
        const float  rsq = dx * dx + dy * dy + dz * dz;
        const bool   included = excl_fep[k] != 0;
        if (rsq >= P.rc2 && included) continue;
        nwithin++;
        float rinv = 0.f, r = 0.f, rp, rpm2;
        if (rsq > 0)
        {
            rinv = rsqrtf(rsq);
            r    = rsq * rinv;
        }
        if (P.soft_core)
        {
            rpm2 = rsq * rsq;
            rp   = rpm2 * rsq;
        }
        else
        {
            rpm2 = rinv * rinv;
            rp   = 1.f;
        }
        float       Fscal = 0.f;
        const float qq[2] = { iqA * qA[jnr], iqB * qB[jnr] };
        float       c6grid[2] = { 0.f, 0.f }; /* nbfp_grid[tj[i]], :609, :764 */
        if (P.ljpme)
        {
            c6grid[0] = grid_c6(P.ljpme, giA, nbfp_comb[typeA[jnr]]);
            c6grid[1] = grid_c6(P.ljpme, giB, nbfp_comb[typeB[jnr]]);
        }
        if (included)
        {
            const float2 pa = nbfp[tiA + typeA[jnr]], pb = nbfp[tiB + typeB[jnr]];
            const float  c6[2] = { pa.x, pb.x }, c12[2] = { pa.y, pb.y };
            float        sigma6[2] = { 0.f, 0.f }, alpha_vdw_eff = 0.f, alpha_coul_eff = 0.f;
            if (P.soft_core)
            {
#pragma unroll
                for (int i = 0; i < 2; i++)
                {
                    if (c6[i] > 0 && c12[i] > 0) sigma6[i] = fmaxf(0.5f * c12[i] / c6[i], P.sigma6_min);
                    else sigma6[i] = P.sigma6_def;
                }
                if (!(c12[0] > 0 && c12[1] > 0)) /* soft-core only where an end state has no repulsion: :498-509 */
                {
                    alpha_vdw_eff  = P.alpha_vdw;
                    alpha_coul_eff = P.alpha_coul;
                }
            }
#pragma unroll
            for (int i = 0; i < 2; i++)
            {
                float FscalC = 0.f, FscalV = 0.f, Vcoul = 0.f, Vvdw = 0.f;
                if (qq[i] != 0 || c6[i] != 0 || c12[i] != 0)
                {
                    float rinvC, rinvV, rC, rV, rpinvC, rpinvV;
                    if (P.soft_core)
                    {
                        rpinvC = 1.0f / (alpha_coul_eff * P.lfac_coul[i] * sigma6[i] + rp);
                        sixth_root(rpinvC, rinvC, rC);
                        if (P.sc_differ)
                        {
                            rpinvV = 1.0f / (alpha_vdw_eff * P.lfac_vdw[i] * sigma6[i] + rp);
                            sixth_root(rpinvV, rinvV, rV);
                        }
                        else
                        {
                            rpinvV = rpinvC, rinvV = rinvC, rV = rC;
                        }
                    }
                    else
                    {
                        rpinvC = rpinvV = 1.f;
                        rinvC = rinvV = rinv;
                        rC = rV = r;
                    }
                    if (qq[i] != 0 && (P.ewald ? r < P.rc : rC < P.rc)) /* :565-581 */
                    {
                        if (P.ewald) /* plain (soft-cored) 1/r; the long-range part is subtracted below, unsoftened */
                        {
                            Vcoul  = qq[i] * (rinvC - P.sh_ewald);
                            FscalC = qq[i] * rinvC;
                        }
                        else
                        {
                            Vcoul  = qq[i] * (rinvC + P.k_rf * rC * rC - P.c_rf);
                            FscalC = qq[i] * (rinvC - 2.0f * P.k_rf * rC * rC);
                        }
                    }
                    if ((c6[i] != 0 || c12[i] != 0) && (P.ljpme ? r < P.rvdw : rV < P.rvdw)) /* :586-611 */
                    {
                        float rinv6;
                        if (P.soft_core) rinv6 = rpinvV;
                        else
                        {
                            rinv6 = rinvV * rinvV;
                            rinv6 = rinv6 * rinv6 * rinv6;
                        }
                        const float v6 = c6[i] * rinv6, v12 = c12[i] * rinv6 * rinv6;
                        Vvdw   = (v12 + c12[i] * P.rep_cpot) * (1.0f / 12.0f) - (v6 + c6[i] * P.disp_cpot) * (1.0f / 6.0f);
                        FscalV = v12 - v6;
                        if (P.ljpme) Vvdw += c6grid[i] * P.sh_lj_ewald * (1.0f / 6.0f); /* the grid potential at the cut-off */
                        if (P.pot_switch) /* :613-625, on the (soft-cored) distance */
                        {
                            const float d = fmaxf(rV - P.rvdw_switch, 0.f), d2 = d * d;
                            const float sw  = 1.0f + d2 * d * (P.sw_c3 + d * (P.sw_c4 + d * P.sw_c5));
                            const float dsw = d2 * (3.0f * P.sw_c3 + d * (4.0f * P.sw_c4 + d * 5.0f * P.sw_c5));
                            FscalV = FscalV * sw - rV * Vvdw * dsw;
                            Vvdw *= sw;
                        }
                    }
                    FscalC *= rpinvC;
                    FscalV *= rpinvV;
                }
                vctot += P.LFC[i] * Vcoul;
                vvtot += P.LFV[i] * Vvdw;
                Fscal += P.LFC[i] * FscalC * rpm2;
                Fscal += P.LFV[i] * FscalV * rpm2;
                dvdl_coul += Vcoul * DLF[i];
                dvdl_vdw += Vvdw * DLF[i];
                if (P.soft_core)
                {
                    dvdl_coul += P.LFC[i] * alpha_coul_eff * P.dlfac_coul[i] * FscalC * sigma6[i];
                    dvdl_vdw += P.LFV[i] * alpha_vdw_eff * P.dlfac_vdw[i] * FscalV * sigma6[i];
                }
            }
        }
        else if (!P.ewald) /* excluded pair: its reaction-field correction, no soft-core; an atom listed with itself counts half: :669-691 */
        {
            const float FF = -2.0f * P.k_rf;
            float       VV = P.k_rf * rsq - P.c_rf;
            if (ii == jnr) VV *= 0.5f;
#pragma unroll
            for (int i = 0; i < 2; i++)
            {
                vctot += P.LFC[i] * qq[i] * VV;
                Fscal += P.LFC[i] * qq[i] * FF;
                dvdl_coul += DLF[i] * qq[i] * VV;
            }
        }
        if (P.ewald && (r < P.rc || !included))
        {
            /* :693-737: the reciprocal-space part of the pair, erf(beta r) / r, subtracted unsoftened (also for excluded pairs and for
             * a perturbed atom with itself, half).  The reference interpolates it from its cubic-spline table; evaluated directly
             * here (the parity tests against the reference kernel measure the difference: forces 1-3e-6) */
            float v_lr, f_lr;
            if (rsq > 0)
            {
                const float br = P.beta * r;
                v_lr           = erff(br) * rinv;
                f_lr           = (v_lr - 1.1283791670955126f * P.beta * expf(-br * br)) * rinv * rinv; /* -(dv/dr) / r */
            }
            else
            {
                v_lr = 1.1283791670955126f * P.beta;
                f_lr = 0.f;
            }
            if (ii == jnr) v_lr *= 0.5f;
#pragma unroll
            for (int i = 0; i < 2; i++)
            {
                vctot -= P.LFC[i] * qq[i] * v_lr;
                Fscal -= P.LFC[i] * qq[i] * f_lr;
                dvdl_coul -= (DLF[i] * qq[i]) * v_lr;
            }
        }
        if (P.ljpme && r < P.rvdw)
        {
            /* :725-770: the grid (reciprocal-space) part of the dispersion, g(x) / r^6 with g = 1 - exp(-x) (1 + x + x^2 / 2),
             * x = (ewaldcoeff_lj r)^2 (tables/forcetable.cpp v_lj_ewald_lr), taken off unsoftened -- for excluded pairs and a
             * perturbed atom with itself (half) too.  The reference interpolates it and its derivative from the cubic-spline table
             * vdwEwaldTables (it avoids the closed form because that cancels for small r, :738-741); evaluated directly here, by
             * the series g = exp(-x) x^3 / 6 (1 + x/4 + x^2/20 + ...) below x = 1, which has no 1 / r^6 in it at all */
            const float b2 = P.lje_coeff2, x = b2 * rsq, ex = expf(-x), b6 = b2 * b2 * b2;
            float       v_lr, f_lr; /* v, -(dv/dr) / r */
            if (rsq > 0)
            {
                const float rinv2 = rinv * rinv;
                if (x < 1.0f)
                {
                    const float ser = x * (1.0f / 4 + x * (1.0f / 20 + x * (1.0f / 120 + x * (1.0f / 840 + x * (1.0f / 6720
                                      + x * (1.0f / 60480 + x * (1.0f / 604800 + x * (1.0f / 6652800 + x * (1.0f / 79833600)))))))));
                    v_lr = b6 * ex * (1.0f + ser) * (1.0f / 6.0f);
                    f_lr = b6 * ex * ser * rinv2;
                }
                else
                {
                    const float g = 1.0f - ex * (1.0f + x + 0.5f * x * x), rinv6 = rinv2 * rinv2 * rinv2;
                    v_lr = g * rinv6;
                    f_lr = (6.0f * g * rinv6 - b6 * ex) * rinv2;
                }
            }
            else
            {
                v_lr = b6 * (1.0f / 6.0f);
                f_lr = 0.f;
            }
            const float FF = f_lr * (1.0f / 6.0f);
            float       VV = v_lr * (1.0f / 6.0f);
            if (ii == jnr) VV *= 0.5f;
#pragma unroll
            for (int i = 0; i < 2; i++)
            {
                vvtot += P.LFV[i] * c6grid[i] * VV;
                Fscal += P.LFV[i] * c6grid[i] * FF;
                dvdl_vdw += (DLF[i] * c6grid[i]) * VV;
            }
        }
        const float tx = Fscal * dx, ty = Fscal * dy, tz = Fscal * dz;
        fix += tx, fiy += ty, fiz += tz;
        atomicAdd(f + sj, make_float4(-tx, -ty, -tz, 0.f));
    }
    const unsigned full = 0xffffffffu;
    for (int o = 16; o > 0; o >>= 1)
    {
        fix += __shfl_xor_sync(full, fix, o);
        fiy += __shfl_xor_sync(full, fiy, o);
        fiz += __shfl_xor_sync(full, fiz, o);
        vctot += __shfl_xor_sync(full, vctot, o);
        vvtot += __shfl_xor_sync(full, vvtot, o);
        dvdl_coul += __shfl_xor_sync(full, dvdl_coul, o);
        dvdl_vdw += __shfl_xor_sync(full, dvdl_vdw, o);
        nwithin += __shfl_xor_sync(full, nwithin, o);
    }
    if (lane == 0)
    {
        if (nwithin > 0)
        {
            atomicAdd(f + si, make_float4(fix, fiy, fiz, 0.f));
            if (is != B200NB_CENTRAL) /* the central shift carries no virial; the cluster-pair kernels skip it as well */
            {
                float* fs = fshift + (n & (NB_OUT_COPIES - 1)) * NB_FSHIFT_PITCH + 3 * is;
                atomicAdd(fs, fix), atomicAdd(fs + 1, fiy), atomicAdd(fs + 2, fiz);
            }
            atomicAdd(out4, (double)vctot);
            atomicAdd(out4 + 1, (double)vvtot);
        }
        atomicAdd(out4 + 2, (double)dvdl_coul);
        atomicAdd(out4 + 3, (double)dvdl_vdw);
    }
}

/* ---- the perturbed pair list, built on the device ----
 * What nbnxm/pairlist.cpp:1699-1872 make_fep_list hands to the free-energy kernel: every atom pair within the list radius with at
 * least one perturbed atom (excluded pairs stay in the list, flagged 0), every perturbed atom with itself (flag 0), grouped into
 * i-entries of one i-atom and one shift.  The reference cuts these pairs out of its cluster-pair list while building it; here the
 * cluster-pair path never sees them (perturbed atoms are masked in its atom data), so the list is a search of its own over the
 * same grid: one warp per perturbed atom walks the cluster bounding boxes, lanes 8 x 4 test the atoms of four clusters at a time.
 * A pair of two perturbed atoms is listed from the one with the lower index.  Two passes (count, fill) around one scan; the order
 * of the j-atoms within an entry is the grid order, so the list and the sums over it are reproducible. */
struct FepListArgs
{
    float box[3], inv_box[3];
    float off[3]; /* triclinic cell: box[YY][XX], box[ZZ][XX], box[ZZ][YY] */
    int   triclinic;
    int   pbc[3];
    float rlist2;
    int   nclusters;
};

/* d reduced to the image with |d_k| <= box_kk / 2, z then y then x (pbcutil/pbc_aiuc_cuda.cuh:60-125); t = the shift of the first
 * point's image (-sh).  Within the list radius (at most half the smallest diagonal element: build_pairlist's max_cutoff2 check)
 * that image is the only one in range. */
__device__ __forceinline__ void fep_reduce(const FepListArgs& A, float* d, int* t)
{
    float sh = A.pbc[2] ? rintf(d[2] * A.inv_box[2]) : 0.f;
    d[0] -= sh * A.off[1], d[1] -= sh * A.off[2], d[2] -= sh * A.box[2];
    t[2] = -(int)sh;
    sh   = A.pbc[1] ? rintf(d[1] * A.inv_box[1]) : 0.f;
    d[0] -= sh * A.off[0], d[1] -= sh * A.box[1];
    t[1] = -(int)sh;
    sh   = A.pbc[0] ? rintf(d[0] * A.inv_box[0]) : 0.f;
    d[0] -= sh * A.box[0];
    t[0] = -(int)sh;
}

template<bool FILL>
__global__ void __launch_bounds__(128)
k_fep_list(int npert, const int* __restrict__ pert, const unsigned char* __restrict__ is_pert, const float4* __restrict__ xq,
           const int* __restrict__ slot_of_atom, const int* __restrict__ atom_index, const float* __restrict__ bb,
           const int* __restrict__ excl_off, const int* __restrict__ excl_idx, const float* __restrict__ shift_vec, const FepListArgs A,
           int* __restrict__ cnt, const int* __restrict__ off, const int* __restrict__ eidx, int* __restrict__ iinr, int* __restrict__ shift,
           int* __restrict__ jindex, int* __restrict__ jjnr, signed char* __restrict__ excl_fep)
{
    __shared__ int s_cnt[4][B200NB_SHIFTS + 3];
    __shared__ int s_q[4][32];
    const int      w = threadIdx.x >> 5, lane = threadIdx.x & 31, ip = blockIdx.x * 4 + w;
    if (ip >= npert) return; /* warp-uniform; only warp-level synchronisation below */
    const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
    const int      p  = pert[ip];
    const float4   xp = xq[slot_of_atom[p]];
    const float    xpv[3] = { xp.x, xp.y, xp.z };
    int            pe0 = 0, pe1 = 0;
    if (excl_off) pe0 = excl_off[p], pe1 = excl_off[p + 1];
    for (int k = lane; k < B200NB_SHIFTS; k += 32) s_cnt[w][k] = 0;
    __syncwarp();
    for (int c0 = 0; c0 < A.nclusters; c0 += 32)
    {
        const int c   = c0 + lane;
        bool      hit = false;
        if (c < A.nclusters)
        {
            const float* b = bb + (size_t)c * 6;
            if (b[0] <= b[3]) /* not an all-filler cluster */
            {
                float dc[3], half[3], d2 = 0.f;
                int   tc[3];
                bool  ambiguous = false;
#pragma unroll
                for (int d = 0; d < 3; d++)
                {
                    half[d] = 0.5f * (b[3 + d] - b[d]) + 1e-4f; /* margin: the atom test below decides */
                    dc[d]   = xpv[d] - 0.5f * (b[3 + d] + b[d]);
                }
                fep_reduce(A, dc, tc);
#pragma unroll
                for (int d = 0; d < 3; d++)
                {
                    const float t = fabsf(dc[d]) - half[d];
                    if (t > 0.f) d2 += t * t;
                    /* a box that reaches across the half-cell plane: its atoms can reduce with another lattice vector than its
                     * centre.  Rectangular cells: that image is only nearer, the test stays conservative.  Triclinic: the other
                     * image moves the lower components too -- take the cluster */
                    ambiguous = ambiguous || (A.pbc[d] && fabsf(dc[d]) + half[d] > 0.5f * A.box[d]);
                }
                if (A.triclinic && ambiguous) d2 = 0.f;
                hit = d2 < A.rlist2;
            }
        }
        const unsigned m = __ballot_sync(full, hit);
        if (hit) s_q[w][__popc(m & lt)] = c;
        const int nq = __popc(m);
        __syncwarp();
        for (int q0 = 0; q0 < nq; q0 += 4)
        {
            const int qi = q0 + (lane >> 3);
            bool      ok = false;
            int       j = -1, s = B200NB_CENTRAL;
            if (qi < nq)
            {
                const int sj = s_q[w][qi] * 8 + (lane & 7);
                j            = atom_index[sj];
                if (j >= 0 && !(j != p && is_pert[j] && j < p))
                {
                    const float4 xj    = xq[sj];
                    float        dv[3] = { xp.x - xj.x, xp.y - xj.y, xp.z - xj.z };
                    int          t[3];
                    fep_reduce(A, dv, t);
                    if (t[0] >= -2 && t[0] <= 2 && t[1] >= -1 && t[1] <= 1 && t[2] >= -1 && t[2] <= 1)
                    {
                        s = 5 * (3 * (t[2] + 1) + (t[1] + 1)) + t[0] + 2; /* pbcutil/ishift.h:50 */
                        const float r2 = nb_rsq(xp.x + shift_vec[3 * s], xp.y + shift_vec[3 * s + 1], xp.z + shift_vec[3 * s + 2], xj.x, xj.y, xj.z);
                        ok = r2 < A.rlist2 && (j != p || s == B200NB_CENTRAL);
                    }
                }
            }
            const unsigned act = __ballot_sync(full, ok);
            if (ok)
            {
                const unsigned same = __match_any_sync(act, s);
                const int      rank = __popc(same & lt), base = s_cnt[w][s];
                __syncwarp(act);
                if (rank == 0) s_cnt[w][s] = base + __popc(same);
                if (FILL)
                {
                    /* excluded: within the central image and on either atom's exclusion list (or the atom itself) */
                    signed char flag = 1;
                    if (j == p) flag = 0;
                    else if (s == B200NB_CENTRAL && excl_off)
                    {
                        for (int e = pe0; e < pe1; e++)
                            if (excl_idx[e] == j) flag = 0;
                        for (int e = excl_off[j]; e < excl_off[j + 1]; e++)
                            if (excl_idx[e] == p) flag = 0;
                    }
                    const int k = off[ip * B200NB_SHIFTS + s] + base + rank;
                    jjnr[k]     = j;
                    excl_fep[k] = flag;
                }
            }
            __syncwarp();
        }
        __syncwarp();
    }
    for (int k = lane; k < B200NB_SHIFTS; k += 32)
    {
        const int n = s_cnt[w][k];
        if (!FILL) cnt[ip * B200NB_SHIFTS + k] = n;
        else if (n > 0)
        {
            const int e = eidx[ip * B200NB_SHIFTS + k];
            iinr[e] = p, shift[e] = k, jindex[e] = off[ip * B200NB_SHIFTS + k];
        }
    }
}

/* exclusive scans of the counts and of (count > 0) over m slots, one block; totals[0] = pairs, totals[1] = non-empty entries */
__global__ void __launch_bounds__(1024) k_fep_scan(int m, const int* __restrict__ cnt, int* __restrict__ off, int* __restrict__ eidx, int* __restrict__ totals)
{
    __shared__ int s_a[32], s_b[32], s_carry[2];
    const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry[0] = s_carry[1] = 0;
    __syncthreads();
    for (int k0 = 0; k0 < m; k0 += 1024)
    {
        const int k = k0 + threadIdx.x;
        const int a = k < m ? cnt[k] : 0, b = a > 0;
        int       ia = a, ib = b;
        for (int o = 1; o < 32; o <<= 1)
        {
            const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) ia += ta, ib += tb;
        }
        if (lane == 31) s_a[w] = ia, s_b[w] = ib;
        __syncthreads();
        if (w == 0)
        {
            int va = s_a[lane], vb = s_b[lane];
            for (int o = 1; o < 32; o <<= 1)
            {
                const int ta = __shfl_up_sync(0xffffffffu, va, o), tb = __shfl_up_sync(0xffffffffu, vb, o);
                if (lane >= o) va += ta, vb += tb;
            }
            s_a[lane] = va, s_b[lane] = vb;
        }
        __syncthreads();
        const int ca = s_carry[0] + (w ? s_a[w - 1] : 0), cb = s_carry[1] + (w ? s_b[w - 1] : 0);
        if (k < m) off[k] = ca + ia - a, eidx[k] = cb + ib - b;
        __syncthreads();
        if (threadIdx.x == 0) s_carry[0] += s_a[31], s_carry[1] += s_b[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[0] = s_carry[0], totals[1] = s_carry[1];
}

template<typename T>
int upload(b200nb_context* h, T** dst, const T* src, size_t n)
{
    cudaFree(*dst);
    *dst = nullptr;
    NB_CUDA(h, cudaMalloc((void**)dst, sizeof(T) * std::max<size_t>(n, 1)));
    if (n) NB_CUDA(h, cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
    return 0;
}

} // namespace

extern "C" int b200nb_fep_set_atoms(b200nb_t* h, const int* typeA_host, const int* typeB_host, const float* qA_host, const float* qB_host)
{
    if (!h || !typeA_host || !typeB_host || !qA_host || !qB_host) return nb_fail(h, B200NB_ERR_ARG, "fep_set_atoms: bad argument");
    if (h->natoms < 1) return nb_fail(h, B200NB_ERR_STATE, "fep_set_atoms: set_atoms first");
    for (int a = 0; a < h->natoms; a++)
        if (typeA_host[a] < 0 || typeA_host[a] >= h->hp.ntypes || typeB_host[a] < 0 || typeB_host[a] >= h->hp.ntypes)
            return nb_fail(h, B200NB_ERR_ARG, "fep_set_atoms: atom type out of range");
    cudaSetDevice(h->device);
    FepState& F = h->fep;
    if (upload(h, &F.d_typeA, typeA_host, h->natoms) || upload(h, &F.d_typeB, typeB_host, h->natoms) || upload(h, &F.d_qA, qA_host, h->natoms)
        || upload(h, &F.d_qB, qB_host, h->natoms))
        return B200NB_ERR_CUDA;
    F.natoms = h->natoms;
    /* the perturbed atoms: charge or type differs between the end states (t_mdatoms::bPerturbed without the mass) */
    std::vector<int>           pert;
    std::vector<unsigned char> is_pert(h->natoms, 0);
    for (int a = 0; a < h->natoms; a++)
        if (typeA_host[a] != typeB_host[a] || qA_host[a] != qB_host[a]) pert.push_back(a), is_pert[a] = 1;
    if (upload(h, &F.d_pert, pert.data(), pert.size()) || upload(h, &F.d_is_pert, is_pert.data(), is_pert.size())) return B200NB_ERR_CUDA;
    F.npert = (int)pert.size();
    h->generation++;
    if (!F.d_out)
    {
        NB_CUDA(h, cudaMalloc((void**)&F.d_out, sizeof(double) * 4));
        NB_CUDA(h, cudaMemset(F.d_out, 0, sizeof(double) * 4));
    }
    return 0;
}

extern "C" int b200nb_fep_upload_list(b200nb_t* h, int nri, const int* iinr, const int* shift, const int* jindex, const int* jjnr,
                                      const signed char* excl_fep)
{
    if (!h || nri < 0 || (nri && (!iinr || !shift || !jindex || !jjnr || !excl_fep))) return nb_fail(h, B200NB_ERR_ARG, "fep_upload_list: bad argument");
    FepState& F = h->fep;
    if (F.natoms < 1) return nb_fail(h, B200NB_ERR_STATE, "fep_upload_list: fep_set_atoms first");
    const int nrj = nri ? jindex[nri] : 0;
    for (int n = 0; n < nri; n++)
        if (iinr[n] < 0 || iinr[n] >= F.natoms || shift[n] < 0 || shift[n] >= B200NB_SHIFTS || jindex[n + 1] < jindex[n])
            return nb_fail(h, B200NB_ERR_ARG, "fep_upload_list: bad i-entry");
    for (int k = 0; k < nrj; k++)
        if (jjnr[k] < 0 || jjnr[k] >= F.natoms) return nb_fail(h, B200NB_ERR_ARG, "fep_upload_list: j-atom out of range");
    cudaSetDevice(h->device);
    if (upload(h, &F.d_iinr, iinr, nri) || upload(h, &F.d_shift, shift, nri) || upload(h, &F.d_jindex, jindex, (size_t)nri + (nri ? 1 : 0))
        || upload(h, &F.d_jjnr, jjnr, nrj) || upload(h, &F.d_excl, excl_fep, nrj))
        return B200NB_ERR_CUDA;
    F.nri = nri, F.nrj = nrj;
    F.cap_nri = F.cap_nrj = 0; /* exact-size arrays: a later b200nb_fep_build_list allocates its own */
    h->generation++; /* captured steps that include the free-energy kernel carry the list pointers by value */
    return 0;
}

extern "C" int b200nb_fep_build_list(b200nb_t* h, int* nri_out, int* nrj_out)
{
    if (!h) return B200NB_ERR_ARG;
    FepState& F = h->fep;
    if (F.natoms != h->natoms || F.natoms < 1) return nb_fail(h, B200NB_ERR_STATE, "fep_build_list: fep_set_atoms for the current atoms first");
    if (!h->grid[0].valid || !h->have_params) return nb_fail(h, B200NB_ERR_STATE, "fep_build_list: set_params and put_on_grid first");
    if (h->dd.window) return nb_fail(h, B200NB_ERR_ARG, "fep_build_list: not built for decomposed runs");
    cudaSetDevice(h->device);
    FepListArgs A{};
    for (int d = 0; d < 3; d++)
    {
        A.box[d] = h->box[d], A.inv_box[d] = h->box[d] > 0.f ? 1.0f / h->box[d] : 0.f, A.pbc[d] = h->pbc[d] && h->box[d] > 0.f;
        A.off[d] = h->box_off[d];
        if (A.pbc[d] && h->box[d] < 2.0f * h->hp.rlist_outer)
            return nb_fail(h, B200NB_ERR_ARG, "fep_build_list: a periodic dimension narrower than twice the list radius");
    }
    A.triclinic = A.off[0] != 0.f || A.off[1] != 0.f || A.off[2] != 0.f;
    /* (every diagonal element >= 2 rlist, checked above, is what makes the reduced image the only one in range: |d| < rlist
     * bounds every component by half its diagonal element) */
    A.rlist2    = h->dp.rlist_outer2;
    A.nclusters = h->npad / 8;
    F.nri       = 0;
    if (nri_out) *nri_out = 0;
    if (nrj_out) *nrj_out = 0;
    if (F.npert == 0) return 0;
    const size_t m = (size_t)F.npert * B200NB_SHIFTS;
    if (3 * m + 2 > F.cap_scratch)
    {
        cudaFree(F.d_scratch);
        F.d_scratch = nullptr, F.cap_scratch = 0;
        NB_CUDA(h, cudaMalloc((void**)&F.d_scratch, sizeof(int) * (3 * m + 2)));
        F.cap_scratch = 3 * m + 2;
    }
    int *d_cnt = F.d_scratch, *d_off = d_cnt + m, *d_eidx = d_off + m, *d_tot = d_eidx + m;
    const unsigned nblk = (unsigned)((F.npert + 3) / 4);
    const float4*  xq   = reinterpret_cast<const float4*>(h->d_xq);
    k_fep_list<false><<<nblk, 128, 0, h->stream>>>(F.npert, F.d_pert, F.d_is_pert, xq, h->d_slot_of_atom, h->d_atom_index, h->d_bb, h->d_excl_off, h->d_excl_idx,
                                                   h->d_shift_vec, A, d_cnt, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    k_fep_scan<<<1, 1024, 0, h->stream>>>((int)m, d_cnt, d_off, d_eidx, d_tot);
    int tot[2] = { 0, 0 };
    NB_CUDA(h, cudaMemcpyAsync(tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaStreamSynchronize(h->stream)); /* the one host round trip of a rebuild: the list's sizes */
    const int nrj = tot[0], nri = tot[1];
    if ((size_t)nri + 1 > F.cap_nri || (size_t)nrj > F.cap_nrj || !F.d_iinr)
    {
        /* the list outgrew its arrays (or was uploaded with exact sizes): new ones with 20 % head-room */
        cudaFree(F.d_iinr), cudaFree(F.d_shift), cudaFree(F.d_jindex), cudaFree(F.d_jjnr), cudaFree(F.d_excl);
        F.d_iinr = F.d_shift = F.d_jindex = F.d_jjnr = nullptr, F.d_excl = nullptr, F.cap_nri = F.cap_nrj = 0;
        const size_t ci = (size_t)(1.2 * nri) + 64, cj = (size_t)(1.2 * nrj) + 256;
        if (cudaMalloc((void**)&F.d_iinr, sizeof(int) * ci) != cudaSuccess || cudaMalloc((void**)&F.d_shift, sizeof(int) * ci) != cudaSuccess
            || cudaMalloc((void**)&F.d_jindex, sizeof(int) * ci) != cudaSuccess || cudaMalloc((void**)&F.d_jjnr, sizeof(int) * cj) != cudaSuccess
            || cudaMalloc((void**)&F.d_excl, cj) != cudaSuccess)
            return nb_fail(h, B200NB_ERR_CUDA, "fep_build_list: out of device memory");
        F.cap_nri = ci, F.cap_nrj = cj;
    }
    k_fep_list<true><<<nblk, 128, 0, h->stream>>>(F.npert, F.d_pert, F.d_is_pert, xq, h->d_slot_of_atom, h->d_atom_index, h->d_bb, h->d_excl_off, h->d_excl_idx,
                                                  h->d_shift_vec, A, nullptr, d_off, d_eidx, F.d_iinr, F.d_shift, F.d_jindex, F.d_jjnr, F.d_excl);
    NB_CUDA(h, cudaMemcpyAsync(F.d_jindex + nri, &d_tot[0], sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    NB_CUDA(h, cudaGetLastError());
    h->nlaunches += 3;
    F.nri = nri, F.nrj = nrj;
    h->generation++;
    if (nri_out) *nri_out = nri;
    if (nrj_out) *nrj_out = nrj;
    return 0;
}

extern "C" int b200nb_fep_get_list(b200nb_t* h, int* iinr_host, int* shift_host, int* jindex_host, int* jjnr_host, signed char* excl_fep_host)
{
    if (!h || !iinr_host || !shift_host || !jindex_host || !jjnr_host || !excl_fep_host) return nb_fail(h, B200NB_ERR_ARG, "fep_get_list: bad argument");
    FepState& F = h->fep;
    cudaSetDevice(h->device);
    if (F.nri == 0)
    {
        jindex_host[0] = 0;
        return 0;
    }
    /* on the context's stream, behind the fill pass of b200nb_fep_build_list (which returns without waiting for it) */
    NB_CUDA(h, cudaMemcpyAsync(iinr_host, F.d_iinr, sizeof(int) * F.nri, cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaMemcpyAsync(shift_host, F.d_shift, sizeof(int) * F.nri, cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaMemcpyAsync(jindex_host, F.d_jindex, sizeof(int) * (F.nri + 1), cudaMemcpyDeviceToHost, h->stream));
    if (F.nrj > 0)
    {
        NB_CUDA(h, cudaMemcpyAsync(jjnr_host, F.d_jjnr, sizeof(int) * F.nrj, cudaMemcpyDeviceToHost, h->stream));
        NB_CUDA(h, cudaMemcpyAsync(excl_fep_host, F.d_excl, F.nrj, cudaMemcpyDeviceToHost, h->stream));
    }
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200nb_fep_launch(b200nb_t* h, const b200nb_fep_params_t* p)
{
    if (!h || !p) return nb_fail(h, B200NB_ERR_ARG, "fep_launch: bad argument");
    FepState& F = h->fep;
    if (F.natoms != h->natoms || !F.d_out) return nb_fail(h, B200NB_ERR_STATE, "fep_launch: fep_set_atoms for the current atoms first");
    if (!h->grid[0].valid) return nb_fail(h, B200NB_ERR_STATE, "fep_launch: put_on_grid first");
    if (h->dp.vdw_modifier == B200NB_VDW_FORCESWITCH)
        return nb_fail(h, B200NB_ERR_ARG, "fep_launch: the LJ force switch is not built for perturbed pairs (the reference's free-energy kernel does not have it either)");
    if (h->dp.ljpme && !h->d_nbfp_comb) return nb_fail(h, B200NB_ERR_STATE, "fep_launch: LJ-PME without its per-type grid parameters (b200nb_set_vdw)");
    if (p->sc_power != 1 && p->sc_power != 2) return nb_fail(h, B200NB_ERR_ARG, "fep_launch: sc_power must be 1 or 2");
    if (F.nri == 0) return 0;
    cudaSetDevice(h->device);
    FepDev D{};
    D.rc = h->hp.rc, D.rc2 = h->dp.rc2, D.epsfac = h->dp.epsfac, D.k_rf = h->dp.k_rf, D.c_rf = h->dp.c_rf, D.disp_cpot = h->dp.disp_cpot, D.rep_cpot = h->dp.rep_cpot;
    D.ntypes = h->dp.ntypes;
    D.rvdw   = h->dp.rvdw2 < h->dp.rc2 ? sqrtf(h->dp.rvdw2) : h->hp.rc; /* b200nb_set_vdw keeps rvdw <= rc */
    D.ewald = h->dp.eeltype == B200NB_EEL_EWALD, D.beta = h->dp.beta, D.sh_ewald = h->dp.sh_ewald;
    D.pot_switch = h->dp.vdw_modifier == B200NB_VDW_POTSWITCH, D.rvdw_switch = h->dp.rvdw_switch;
    D.sw_c3 = h->dp.sw_c3, D.sw_c4 = h->dp.sw_c4, D.sw_c5 = h->dp.sw_c5;
    D.ljpme = h->dp.ljpme, D.lje_coeff2 = h->dp.lje_coeff2, D.sh_lj_ewald = h->dp.sh_lj_ewald; /* set_vdw admits LJ-PME with the potential shift only */
    /* interaction_const_t::SoftCoreParameters (mdtypes/interaction_const.cpp:47-56) */
    D.alpha_vdw  = p->sc_alpha;
    D.alpha_coul = p->sc_coul ? p->sc_alpha : 0.f;
    const float s2 = p->sc_sigma * p->sc_sigma, m2 = p->sc_sigma_min * p->sc_sigma_min;
    D.sigma6_def = s2 * s2 * s2;
    D.sigma6_min = p->sc_coul ? m2 * m2 * m2 : 0.f;
    D.soft_core  = !(D.alpha_coul == 0.f && D.alpha_vdw == 0.f);
    D.sc_differ  = D.soft_core && !(p->lambda_coul == p->lambda_vdw && D.alpha_coul == D.alpha_vdw);
    D.LFC[0] = 1.f - p->lambda_coul, D.LFC[1] = p->lambda_coul, D.LFV[0] = 1.f - p->lambda_vdw, D.LFV[1] = p->lambda_vdw;
    const float DLF[2] = { -1.f, 1.f }, lp = (float)p->sc_power;
    for (int i = 0; i < 2; i++) /* nb_free_energy.cpp:363-370 */
    {
        D.lfac_coul[i]  = p->sc_power == 2 ? (1 - D.LFC[i]) * (1 - D.LFC[i]) : (1 - D.LFC[i]);
        D.dlfac_coul[i] = DLF[i] * lp / 6.0f * (p->sc_power == 2 ? (1 - D.LFC[i]) : 1);
        D.lfac_vdw[i]   = p->sc_power == 2 ? (1 - D.LFV[i]) * (1 - D.LFV[i]) : (1 - D.LFV[i]);
        D.dlfac_vdw[i]  = DLF[i] * lp / 6.0f * (p->sc_power == 2 ? (1 - D.LFV[i]) : 1);
    }
    k_fep<<<(unsigned)((F.nri + 3) / 4), 128, 0, h->stream>>>(F.nri, F.d_iinr, F.d_shift, F.d_jindex, F.d_jjnr, F.d_excl, reinterpret_cast<const float4*>(h->d_xq),
                                                            h->d_slot_of_atom, h->d_shift_vec, F.d_typeA, F.d_typeB, F.d_qA, F.d_qB,
                                                            reinterpret_cast<const float2*>(h->d_nbfp), reinterpret_cast<const float2*>(h->d_nbfp_comb), D,
                                                            h->d_f, h->d_fshift, F.d_out);
    h->nlaunches++;
    NB_CUDA(h, cudaGetLastError());
    return 0;
}

extern "C" int b200nb_fep_get_outputs(b200nb_t* h, double out4_host[4])
{
    if (!h || !out4_host) return nb_fail(h, B200NB_ERR_ARG, "fep_get_outputs: bad argument");
    FepState& F = h->fep;
    if (!F.d_out) return nb_fail(h, B200NB_ERR_STATE, "fep_get_outputs: fep_set_atoms first");
    cudaSetDevice(h->device);
    NB_CUDA(h, cudaMemcpyAsync(out4_host, F.d_out, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaMemsetAsync(F.d_out, 0, sizeof(double) * 4, h->stream)); /* read and reset: the sums of the launches since the last read */
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

/* The free-energy kernel as part of b200nb_step / b200nb_compute (inside the captured step graph, after the force kernel);
 * p == NULL takes it out again.  The list is the one current at capture: build or upload it before the next step. */
extern "C" int b200nb_fep_in_step(b200nb_t* h, const b200nb_fep_params_t* p)
{
    if (!h) return B200NB_ERR_ARG;
    h->fep.in_step = p != nullptr;
    if (p) h->fep.step_params = *p;
    h->generation++;
    return 0;
}

int nb_fep_enqueue_in_step(b200nb_context* h)
{
    if (!h->fep.in_step) return 0;
    return b200nb_fep_launch(h, &h->fep.step_params);
}

void nb_fep_free(b200nb_context* h)
{
    FepState& F = h->fep;
    cudaFree(F.d_typeA), cudaFree(F.d_typeB), cudaFree(F.d_qA), cudaFree(F.d_qB), cudaFree(F.d_iinr), cudaFree(F.d_shift), cudaFree(F.d_jindex);
    cudaFree(F.d_jjnr), cudaFree(F.d_excl), cudaFree(F.d_out), cudaFree(F.d_pert), cudaFree(F.d_is_pert), cudaFree(F.d_scratch);
    F = FepState{};
}
