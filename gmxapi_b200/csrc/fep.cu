/* b200nb: the perturbed-pair (free-energy) kernel on the device.
 *
 * Replaces CPU code of the reference: gmxlib/nonbonded/nb_free_energy.cpp:203-860 (nb_free_energy_kernel, the scalar instantiation:
 * the reference has no SIMD or GPU form of it and runs it on the host beside the GPU nonbonded kernels, mdlib/sim_util.cpp
 * do_nb_verlet -> nonbonded_verlet_t::dispatchFreeEnergyKernel, nbnxm/kerneldispatch.cpp:486-588), for the flavours built so far:
 * reaction-field / plain cut-off or Ewald electrostatics (the long-range part subtracted unsoftened, :693-737, evaluated directly
 * instead of from the reference's spline table), cut-off LJ with potential shift, soft-core with r-power 6 (lambda power 1 or 2) or
 * none.  LJ-PME and the LJ potential switch are refused.
 * The pair list comes from the caller in t_nblist form (mdtypes/nblist.h:117-137; b200nb_fep_upload_list) -- what
 * nbnxm/pairlist.cpp:1699-1872 make_fep_list produces: every pair within the list radius with a perturbed atom, excluded pairs
 * flagged, perturbed atoms listed with themselves; building it on the device from the cluster-pair search is the next step.
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
    float LFC[2], LFV[2];
    float lfac_coul[2], dlfac_coul[2], lfac_vdw[2], dlfac_vdw[2];
    float alpha_coul, alpha_vdw, sigma6_def, sigma6_min;
    int   soft_core, sc_differ, ntypes;
    int   ewald;
    float beta, sh_ewald;
};

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
      const float2* __restrict__ nbfp, const __grid_constant__ FepDev P, float4* __restrict__ f, float* __restrict__ fshift, double* __restrict__ out4)
{
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= nri) return;
    const int    ii = iinr[n], is = shift[n];
    const int    si = slot_of_atom[ii];
    const float4 xi = xq[si];
    const float  ix = shift_vec[3 * is] + xi.x, iy = shift_vec[3 * is + 1] + xi.y, iz = shift_vec[3 * is + 2] + xi.z;
    const float  iqA = P.epsfac * qA[ii], iqB = P.epsfac * qB[ii];
    const int    tiA = typeA[ii] * P.ntypes, tiB = typeB[ii] * P.ntypes;
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
                    if ((c6[i] != 0 || c12[i] != 0) && rV < P.rc) /* :588-607 */
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
    F.nri = nri;
    return 0;
}

extern "C" int b200nb_fep_launch(b200nb_t* h, const b200nb_fep_params_t* p)
{
    if (!h || !p) return nb_fail(h, B200NB_ERR_ARG, "fep_launch: bad argument");
    FepState& F = h->fep;
    if (F.natoms != h->natoms || !F.d_out) return nb_fail(h, B200NB_ERR_STATE, "fep_launch: fep_set_atoms for the current atoms first");
    if (!h->grid[0].valid) return nb_fail(h, B200NB_ERR_STATE, "fep_launch: put_on_grid first");
    if (h->dp.vdw_modifier != B200NB_VDW_POTSHIFT || h->dp.rvdw2 < h->dp.rc2 || h->dp.ljpme != 0)
        return nb_fail(h, B200NB_ERR_ARG, "fep_launch: only cut-off LJ with potential shift and rvdw = rcoulomb is built for perturbed pairs");
    if (p->sc_power != 1 && p->sc_power != 2) return nb_fail(h, B200NB_ERR_ARG, "fep_launch: sc_power must be 1 or 2");
    if (F.nri == 0) return 0;
    cudaSetDevice(h->device);
    FepDev D{};
    D.rc = h->hp.rc, D.rc2 = h->dp.rc2, D.epsfac = h->dp.epsfac, D.k_rf = h->dp.k_rf, D.c_rf = h->dp.c_rf, D.disp_cpot = h->dp.disp_cpot, D.rep_cpot = h->dp.rep_cpot;
    D.ntypes = h->dp.ntypes;
    D.ewald = h->dp.eeltype == B200NB_EEL_EWALD, D.beta = h->dp.beta, D.sh_ewald = h->dp.sh_ewald;
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
                                                            reinterpret_cast<const float2*>(h->d_nbfp), D, h->d_f, h->d_fshift, F.d_out);
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

void nb_fep_free(b200nb_context* h)
{
    FepState& F = h->fep;
    cudaFree(F.d_typeA), cudaFree(F.d_typeB), cudaFree(F.d_qA), cudaFree(F.d_qB), cudaFree(F.d_iinr), cudaFree(F.d_shift), cudaFree(F.d_jindex);
    cudaFree(F.d_jjnr), cudaFree(F.d_excl), cudaFree(F.d_out);
    F = FepState{};
}
