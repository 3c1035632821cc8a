/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * C-ABI harness around the UNMODIFIED reference nbnxm CPU path, compiled from the
 * sources where they lie under /root/reference by oracle/build_ref.sh into
 * oracle/_ref/libgmxref_nbnxm.so.  It assembles a nonbonded_verlet_t by hand the way
 * the reference's own benchmark does (src/gromacs/nbnxm/benchmark/bench_setup.cpp:170-232
 * setupNbnxmForBenchInstance; api/nblib/gmxsetup.cpp:175-208 setupNbnxmInstance) and
 * exposes: forces / shift forces / energies from the plain-C 4x4, SIMD 4xN/2xNN or
 * GPU-emulation 8x8x8 kernels, the grid atom order, the in-range atom-pair set of the
 * reference list, and wall-clock timings of grid / search / kernel.
 *
 * Used by: tests/ (validation of oracle/nbnxm_oracle.c), bench.py cpu_baseline and
 * `bench.py --impl reference`.
 */
#include "gmxpre.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "gromacs/ewald/ewald_utils.h"
#include "gromacs/gmxlib/nrnb.h"
#include "gromacs/gpu_utils/hostallocator.h"
#include "gromacs/math/units.h"
#include "gromacs/math/vec.h"
#include "gromacs/mdlib/gmx_omp_nthreads.h"
#include "gromacs/mdtypes/enerdata.h"
#include "gromacs/mdtypes/forcerec.h"
#include "gromacs/nbnxm/benchmark/bench_coords.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/nbnxm/atomdata.h"
#include "gromacs/nbnxm/gridset.h"
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/nbnxm_simd.h"
#include "gromacs/nbnxm/pairlist.h"
#include "gromacs/nbnxm/pairlistset.h"
#include "gromacs/nbnxm/pairlistsets.h"
#include "gromacs/nbnxm/pairsearch.h"
#include "gromacs/pbcutil/ishift.h"
#include "gromacs/pbcutil/pbc.h"
#include "gromacs/simd/simd.h"
#include "gromacs/simd/vector_operations.h"
#include "gromacs/tables/forcetable.h"
#include "gromacs/utility/listoflists.h"
#include "gromacs/utility/logger.h"

#include "ref_harness.h"

namespace
{

double nowSeconds()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/* r^2 exactly as the reference SIMD kernels evaluate it: gmx::norm2 on SimdReal
 * (simd/vector_operations.h:106-115), compiled here with the same flags as the kernels,
 * i-atom shifted first (kernels_simd_2xmm/kernel_outer.h:482-489), dx = xi - xj. */
float simdRsq(float xi, float yi, float zi, float xj, float yj, float zj)
{
#if GMX_SIMD_HAVE_REAL
    using gmx::SimdReal;
    SimdReal                                   dx = SimdReal(xi) - SimdReal(xj);
    SimdReal                                   dy = SimdReal(yi) - SimdReal(yj);
    SimdReal                                   dz = SimdReal(zi) - SimdReal(zj);
    SimdReal                                   r2 = gmx::norm2(dx, dy, dz);
    alignas(GMX_SIMD_ALIGNMENT) float          tmp[GMX_SIMD_REAL_WIDTH];
    gmx::store(tmp, r2);
    return tmp[0];
#else
    float dx = xi - xj, dy = yi - yj, dz = zi - zj;
    return dx * dx + dy * dy + dz * dz;
#endif
}

struct Instance
{
    std::unique_ptr<nonbonded_verlet_t> nbv;
    interaction_const_t                 ic;
    std::vector<gmx::RVec>              x;
    std::vector<int>                    atomInfo;
    gmx::ListOfLists<int>               excls;
    matrix                              box;
    rvec                                shiftVec[SHIFTS];
    t_forcerec*                         fr = nullptr; // zero-filled, only shift_vec / bBHAM are read
    int                                 natoms   = 0;
    int                                 nthreads = 1;
    Nbnxm::KernelType                   kernelType;
    double                              tGrid = 0, tSearch = 0;
};

Nbnxm::KernelType kernelTypeFromInt(int k)
{
    switch (k)
    {
        case GMXREF_KERNEL_PLAINC_4X4: return Nbnxm::KernelType::Cpu4x4_PlainC;
        case GMXREF_KERNEL_SIMD_4XN: return Nbnxm::KernelType::Cpu4xN_Simd_4xN;
        case GMXREF_KERNEL_SIMD_2XNN: return Nbnxm::KernelType::Cpu4xN_Simd_2xNN;
        case GMXREF_KERNEL_GPUREF_8X8X8: return Nbnxm::KernelType::Cpu8x8x8_PlainC;
        default: return Nbnxm::KernelType::NotSet;
    }
}

} // namespace

extern "C" {

int gmxref_simd_width(void)
{
#if GMX_SIMD_HAVE_REAL
    return GMX_SIMD_REAL_WIDTH;
#else
    return 0;
#endif
}

int gmxref_default_simd_kernel(void)
{
#if defined GMX_NBNXN_SIMD_2XNN && !defined GMX_NBNXN_SIMD_4XN
    return GMXREF_KERNEL_SIMD_2XNN;
#elif defined GMX_NBNXN_SIMD_4XN
    return GMXREF_KERNEL_SIMD_4XN;
#else
    return GMXREF_KERNEL_PLAINC_4X4;
#endif
}

float gmxref_ewald_coeff(float rc, float rtol)
{
    return calc_ewaldcoeff_q(rc, rtol);
}

float gmxref_simd_rsq(float xi, float yi, float zi, float xj, float yj, float zj)
{
    return simdRsq(xi, yi, zi, xj, yj, zj);
}

void* gmxref_create(const gmxref_system* s, const gmxref_params* p)
{
    Nbnxm::KernelType kernelType = kernelTypeFromInt(p->kernel);
    if (kernelType == Nbnxm::KernelType::NotSet)
    {
        return nullptr;
    }
#ifndef GMX_NBNXN_SIMD_4XN
    if (kernelType == Nbnxm::KernelType::Cpu4xN_Simd_4xN)
    {
        return nullptr;
    }
#endif
#ifndef GMX_NBNXN_SIMD_2XNN
    if (kernelType == Nbnxm::KernelType::Cpu4xN_Simd_2xNN)
    {
        return nullptr;
    }
#endif
    auto* inst       = new Instance;
    inst->natoms     = s->natoms;
    inst->nthreads   = p->nthreads > 0 ? p->nthreads : 1;
    inst->kernelType = kernelType;
    const int nth    = inst->nthreads;
    gmx_omp_nthreads_set(emntPairsearch, nth);
    gmx_omp_nthreads_set(emntNonbonded, nth);
    gmx_omp_nthreads_set(emntDefault, nth);

    /* interaction constants (bench_setup.cpp:134-166, api/nblib/gmxsetup.cpp:226-284) */
    interaction_const_t& ic = inst->ic;
    ic.vdwtype               = evdwCUT;
    ic.vdw_modifier          = eintmodPOTSHIFT;
    ic.rvdw                  = p->rc;
    ic.coulomb_modifier      = eintmodPOTSHIFT;
    ic.rcoulomb              = p->rc;
    ic.dispersion_shift.cpot = p->disp_cpot;
    ic.repulsion_shift.cpot  = p->rep_cpot;
    /* the derived constants come from the caller exactly as init_interaction_const (mdlib/forcerec.cpp:850-874) stores
     * them; that function is static in forcerec.cpp, so the test side restates its three formulas */
    if (p->rvdw > 0) ic.rvdw = p->rvdw;
    if (p->vdw_modifier == 1)
    {
        ic.vdw_modifier        = eintmodFORCESWITCH;
        ic.rvdw_switch         = p->rvdw_switch;
        ic.dispersion_shift.c2 = p->disp_c2;
        ic.dispersion_shift.c3 = p->disp_c3;
        ic.repulsion_shift.c2  = p->rep_c2;
        ic.repulsion_shift.c3  = p->rep_c3;
    }
    else if (p->vdw_modifier == 2)
    {
        ic.vdw_modifier  = eintmodPOTSWITCH;
        ic.rvdw_switch   = p->rvdw_switch;
        ic.vdw_switch.c3 = p->sw_c3;
        ic.vdw_switch.c4 = p->sw_c4;
        ic.vdw_switch.c5 = p->sw_c5;
    }
    if (p->ljpme)
    {
        ic.vdwtype        = evdwPME;
        ic.ljpme_comb_rule = (p->ljpme == 1) ? eljpmeGEOM : eljpmeLB;
        ic.ewaldcoeff_lj  = p->ewaldcoeff_lj;
        ic.sh_lj_ewald    = p->sh_lj_ewald;
    }
    ic.epsilon_r             = 1;
    ic.epsfac                = p->epsfac;
    ic.k_rf                  = p->k_rf;
    ic.c_rf                  = p->c_rf;
    ic.sh_ewald              = p->sh_ewald;
    switch (p->eeltype)
    {
        case GMXREF_EEL_CUT: ic.eeltype = eelCUT; break;
        case GMXREF_EEL_RF: ic.eeltype = eelRF; break;
        default: ic.eeltype = eelPME; break;
    }
    Nbnxm::KernelSetup kernelSetup;
    kernelSetup.kernelType = kernelType;
    const bool simdKernel  = (kernelType == Nbnxm::KernelType::Cpu4xN_Simd_4xN
                             || kernelType == Nbnxm::KernelType::Cpu4xN_Simd_2xNN);
    kernelSetup.ewaldExclusionType = (simdKernel && p->eeltype == GMXREF_EEL_EWALD_ANA)
                                             ? Nbnxm::EwaldExclusionType::Analytical
                                             : Nbnxm::EwaldExclusionType::Table;
    if (ic.eeltype == eelPME)
    {
        ic.ewaldcoeff_q       = p->ewaldcoeff;
        ic.coulombEwaldTables = std::make_unique<EwaldCorrectionTables>();
        /* mdlib/forcerec.cpp:724-763 init_ewald_f_table, Coulomb only */
        const real tableScale = ewald_spline3_table_scale(ic, true, false);
        const int  tableSize  = static_cast<int>(ic.rcoulomb * tableScale) + 2;
        *ic.coulombEwaldTables =
                generateEwaldCorrectionTables(tableSize, tableScale, ic.ewaldcoeff_q, v_q_ewald_lr);
    }

    clear_mat(inst->box);
    inst->box[XX][XX] = s->box[0];
    inst->box[YY][YY] = s->box[1];
    inst->box[ZZ][ZZ] = s->box[2];
    inst->box[YY][XX] = p->box_offdiag[0];
    inst->box[ZZ][XX] = p->box_offdiag[1];
    inst->box[ZZ][YY] = p->box_offdiag[2];
    calc_shifts(inst->box, inst->shiftVec);
    inst->fr            = static_cast<t_forcerec*>(std::calloc(1, sizeof(t_forcerec)));
    inst->fr->shift_vec = inst->shiftVec;

    inst->x.resize(s->natoms);
    inst->atomInfo.assign(s->natoms, 0);
    std::vector<int>  types(s->type, s->type + s->natoms);
    std::vector<real> charges(s->q, s->q + s->natoms);
    for (int a = 0; a < s->natoms; a++)
    {
        inst->x[a] = { s->x[3 * a], s->x[3 * a + 1], s->x[3 * a + 2] };
        /* bench_system.cpp:176-189: all atoms flagged VdW + Q unless the caller asks for
         * the exact per-atom flags (half-LJ optimisation). */
        bool hasVdw = true, hasQ = true;
        if (p->exact_atom_flags)
        {
            const int t = s->type[a];
            hasVdw      = false;
            for (int t2 = 0; t2 < s->ntypes; t2++)
            {
                if (s->nbfp[(t * s->ntypes + t2) * 2] != 0 || s->nbfp[(t * s->ntypes + t2) * 2 + 1] != 0)
                {
                    hasVdw = true;
                }
            }
            hasQ = (s->q[a] != 0);
        }
        if (hasVdw)
        {
            SET_CGINFO_HAS_VDW(inst->atomInfo[a]);
        }
        if (hasQ)
        {
            SET_CGINFO_HAS_Q(inst->atomInfo[a]);
        }
        if (p->perturbed && p->perturbed[a])
        {
            SET_CGINFO_FEP(inst->atomInfo[a]);
        }
        const int n0 = s->excl_off[a], n1 = s->excl_off[a + 1];
        inst->excls.pushBackListOfSize(n1 - n0);
        gmx::ArrayRef<int> e = inst->excls.back();
        for (int k = n0; k < n1; k++)
        {
            e[k - n0] = s->excl_idx[k];
        }
    }
    if (p->put_in_box)
    {
        put_atoms_in_box(PbcType::Xyz, inst->box, inst->x);
    }

    const auto pin = gmx::PinningPolicy::CannotBePinned;
    const bool     haveFep = p->perturbed != nullptr;
    PairlistParams pairlistParams(kernelType, haveFep, p->rlist, false);
    if (p->rlist_inner > 0 && p->rlist_inner < p->rlist)
    {
        /* dynamic pruning set up as pairlist_tuning.cpp:506-572 would */
        pairlistParams.useDynamicPruning      = true;
        pairlistParams.rlistInner             = p->rlist_inner;
        pairlistParams.nstlistPrune           = p->nstlist_prune > 0 ? p->nstlist_prune : 4;
        pairlistParams.numRollingPruningParts = 1;
        pairlistParams.lifetime               = 100;
    }
    auto pairlistSets = std::make_unique<PairlistSets>(pairlistParams, false, p->min_ilist_count);
    auto pairSearch   = std::make_unique<PairSearch>(PbcType::Xyz, false, nullptr, nullptr,
                                                   pairlistParams.pairlistType, haveFep, nth, pin);
    auto atomData     = std::make_unique<nbnxn_atomdata_t>(pin);
    inst->nbv = std::make_unique<nonbonded_verlet_t>(std::move(pairlistSets), std::move(pairSearch),
                                                     std::move(atomData), kernelSetup, nullptr, nullptr);
    std::vector<real> nbfp(s->nbfp, s->nbfp + 2 * s->ntypes * s->ntypes);
    nbnxn_atomdata_init(gmx::MDLogger(), inst->nbv->nbat.get(), kernelType, p->comb_rule, s->ntypes,
                        nbfp, 1, nth);

    const rvec lower   = { 0, 0, 0 };
    const rvec upper   = { inst->box[XX][XX], inst->box[YY][YY], inst->box[ZZ][ZZ] };
    const real density = s->natoms / det(inst->box);
    t_nrnb     nrnb;
    double     t0 = nowSeconds();
    nbnxn_put_on_grid(inst->nbv.get(), inst->box, 0, lower, upper, nullptr, { 0, s->natoms }, density,
                      inst->atomInfo, inst->x, 0, nullptr);
    double t1 = nowSeconds();
    inst->nbv->constructPairlist(gmx::InteractionLocality::Local, inst->excls, 0, &nrnb);
    double t2     = nowSeconds();
    inst->tGrid   = t1 - t0;
    inst->tSearch = t2 - t1;
    inst->nbv->setAtomProperties(types, charges, inst->atomInfo);
    return inst;
}

void gmxref_destroy(void* h)
{
    auto* inst = static_cast<Instance*>(h);
    if (inst)
    {
        std::free(inst->fr);
        delete inst;
    }
}

/* Re-grid + re-search with timing (for the CPU baseline of the search stage). */
int gmxref_regrid_research(void* h, double* tGrid, double* tSearch)
{
    auto*      inst    = static_cast<Instance*>(h);
    const rvec lower   = { 0, 0, 0 };
    const rvec upper   = { inst->box[XX][XX], inst->box[YY][YY], inst->box[ZZ][ZZ] };
    const real density = inst->natoms / det(inst->box);
    t_nrnb     nrnb;
    double     t0 = nowSeconds();
    nbnxn_put_on_grid(inst->nbv.get(), inst->box, 0, lower, upper, nullptr, { 0, inst->natoms },
                      density, inst->atomInfo, inst->x, 0, nullptr);
    double t1 = nowSeconds();
    inst->nbv->constructPairlist(gmx::InteractionLocality::Local, inst->excls, 0, &nrnb);
    double t2 = nowSeconds();
    *tGrid    = t1 - t0;
    *tSearch  = t2 - t1;
    return 0;
}

void gmxref_setup_times(void* h, double* tGrid, double* tSearch)
{
    auto* inst = static_cast<Instance*>(h);
    *tGrid     = inst->tGrid;
    *tSearch   = inst->tSearch;
}

/* Full force evaluation as GmxForceCalculator::compute does (api/nblib/gmxcalculator.cpp:70-83):
 * convertCoordinates -> dispatchNonbondedKernel(clearF) -> atomdata_add_nbat_f_to_f.
 * x may be NULL (use the coordinates given at create). f[3N] is overwritten, fshift[135]
 * overwritten (zeros unless want_virial), energies[2] = {Vvdw, Vcoul}. */
int gmxref_compute(void* h, const float* x, int want_energy, int want_virial, float* f, float* fshift,
                   float* energies)
{
    auto* inst = static_cast<Instance*>(h);
    if (x)
    {
        for (int a = 0; a < inst->natoms; a++)
        {
            inst->x[a] = { x[3 * a], x[3 * a + 1], x[3 * a + 2] };
        }
    }
    inst->nbv->convertCoordinates(gmx::AtomLocality::Local, false, inst->x);
    gmx::StepWorkload stepWork;
    stepWork.computeForces = true;
    stepWork.computeEnergy = want_energy != 0;
    stepWork.computeVirial = want_virial != 0 || want_energy != 0;
    gmx_enerdata_t enerd(1, 0);
    t_nrnb         nrnb = { 0 };
    inst->nbv->dispatchNonbondedKernel(gmx::InteractionLocality::Local, inst->ic, stepWork,
                                       enbvClearFYes, *inst->fr, &enerd, &nrnb);
    std::vector<gmx::RVec> force(inst->natoms, { 0, 0, 0 });
    if (inst->kernelType == Nbnxm::KernelType::Cpu8x8x8_PlainC)
    {
        /* nbnxm.cpp:167-171 skips the reduction for GPU-layout lists without a physical GPU;
         * call the same reduction (atomdata.cpp:1425) directly. */
        reduceForces(inst->nbv->nbat.get(), gmx::AtomLocality::All, inst->nbv->pairSearch_->gridSet(),
                     as_rvec_array(force.data()));
    }
    else
    {
        inst->nbv->atomdata_add_nbat_f_to_f(gmx::AtomLocality::All, force);
    }
    for (int a = 0; a < inst->natoms; a++)
    {
        f[3 * a]     = force[a][XX];
        f[3 * a + 1] = force[a][YY];
        f[3 * a + 2] = force[a][ZZ];
    }
    if (fshift)
    {
        std::vector<gmx::RVec> fs(SHIFTS, { 0, 0, 0 });
        if (stepWork.computeVirial)
        {
            nbnxn_atomdata_add_nbat_fshift_to_fshift(*inst->nbv->nbat, fs);
        }
        for (int sIdx = 0; sIdx < SHIFTS; sIdx++)
        {
            for (int d = 0; d < DIM; d++)
            {
                fshift[3 * sIdx + d] = fs[sIdx][d];
            }
        }
    }
    if (energies)
    {
        energies[0] = enerd.grpp.ener[egLJSR][0];
        energies[1] = enerd.grpp.ener[egCOULSR][0];
    }
    return 0;
}

/* Kernel-only timing, the protocol of bench_setup.cpp:303-326: one pre-iteration with
 * force clearing, then niter iterations without clearing. Returns seconds per iteration. */
double gmxref_time_kernel(void* h, int want_energy, int nwarm, int niter)
{
    auto* inst = static_cast<Instance*>(h);
    inst->nbv->convertCoordinates(gmx::AtomLocality::Local, false, inst->x);
    gmx::StepWorkload stepWork;
    stepWork.computeForces = true;
    stepWork.computeEnergy = want_energy != 0;
    stepWork.computeVirial = want_energy != 0;
    gmx_enerdata_t enerd(1, 0);
    t_nrnb         nrnb = { 0 };
    for (int i = 0; i < (nwarm > 0 ? nwarm : 1); i++)
    {
        inst->nbv->dispatchNonbondedKernel(gmx::InteractionLocality::Local, inst->ic, stepWork,
                                           enbvClearFYes, *inst->fr, &enerd, &nrnb);
    }
    double t0 = nowSeconds();
    for (int i = 0; i < niter; i++)
    {
        inst->nbv->dispatchNonbondedKernel(gmx::InteractionLocality::Local, inst->ic, stepWork,
                                           enbvClearFNo, *inst->fr, &enerd, &nrnb);
    }
    return (nowSeconds() - t0) / (niter > 0 ? niter : 1);
}

/* Whole-step timing: x convert + kernel (clearing f) + force un-sort, per iteration. */
double gmxref_time_step(void* h, int want_energy, int nwarm, int niter)
{
    auto*              inst = static_cast<Instance*>(h);
    std::vector<float> f(3 * size_t(inst->natoms));
    std::vector<float> fs(3 * SHIFTS);
    float              e[2];
    for (int i = 0; i < nwarm; i++)
    {
        gmxref_compute(h, nullptr, want_energy, 0, f.data(), fs.data(), e);
    }
    double t0 = nowSeconds();
    for (int i = 0; i < niter; i++)
    {
        gmxref_compute(h, nullptr, want_energy, 0, f.data(), fs.data(), e);
    }
    return (nowSeconds() - t0) / (niter > 0 ? niter : 1);
}

/* Order of atoms on the reference grid: out[k] = original atom index at grid slot k or -1
 * for a filler. Returns the number of slots (<= cap written). */
int gmxref_grid_order(void* h, int* out, int cap)
{
    auto* inst    = static_cast<Instance*>(h);
    auto  indices = inst->nbv->pairSearch_->gridSet().atomIndices();
    int   n       = inst->nbv->pairSearch_->gridSet().numGridAtomsTotal();
    for (int k = 0; k < n && k < cap; k++)
    {
        out[k] = indices[k];
    }
    return n;
}

void gmxref_grid_dims(void* h, int* ncx, int* ncy, float* cellx, float* celly, int* natomsPadded)
{
    auto*       inst = static_cast<Instance*>(h);
    const auto& g    = inst->nbv->pairSearch_->gridSet().grids()[0];
    *ncx             = g.dimensions().numCells[XX];
    *ncy             = g.dimensions().numCells[YY];
    *cellx           = g.dimensions().cellSize[XX];
    *celly           = g.dimensions().cellSize[YY];
    *natomsPadded    = inst->nbv->pairSearch_->gridSet().numGridAtomsTotal();
}

/* List statistics: number of cluster pairs in the list (na_ci x na_cj tiles) and the number of
 * atom pairs the kernel evaluates ("total" pairs of bench_setup.cpp:316). */
void gmxref_list_stats(void* h, long long* nClusterPairs, long long* nAtomPairsComputed, int* na_ci,
                       int* na_cj)
{
    auto*       inst = static_cast<Instance*>(h);
    const auto& set  = inst->nbv->pairlistSets().pairlistSet(gmx::InteractionLocality::Local);
    long long   ncp  = 0;
    if (inst->kernelType == Nbnxm::KernelType::Cpu8x8x8_PlainC)
    {
        const NbnxnPairlistGpu* l = set.gpuList();
        *na_ci                    = l->na_ci;
        *na_cj                    = l->na_cj;
        for (const auto& cj4 : l->cj4)
        {
            for (int jm = 0; jm < c_nbnxnGpuJgroupSize; jm++)
            {
                for (int ic = 0; ic < c_nbnxnGpuNumClusterPerSupercluster; ic++)
                {
                    ncp += (cj4.imei[0].imask >> (jm * c_nbnxnGpuNumClusterPerSupercluster + ic)) & 1;
                }
            }
        }
    }
    else
    {
        for (const auto& l : set.cpuLists())
        {
            *na_ci = l.na_ci;
            *na_cj = l.na_cj;
            for (const auto& ci : l.ci)
            {
                ncp += ci.cj_ind_end - ci.cj_ind_start;
            }
        }
    }
    *nClusterPairs      = ncp;
    *nAtomPairsComputed = ncp * (*na_ci) * (*na_cj);
}

/* The in-range atom-pair set of the reference list: every list entry x interaction mask bit
 * whose r^2 (simdRsq above, i-atom shifted) is < rc^2, excluding filler atoms and
 * pairs where both... (nothing else: masks already contain topology and half-list exclusions).
 * Pairs are returned in ORIGINAL atom indices as (i, j, shift) triples with i the shifted atom.
 * Pass pairs == NULL to only count. Returns the number of pairs (may exceed cap; only cap
 * are written). */
long long gmxref_pair_set(void* h, float rc, int* pairs, long long cap)
{
    auto*       inst    = static_cast<Instance*>(h);
    const auto& gridSet = inst->nbv->pairSearch_->gridSet();
    auto        indices = gridSet.atomIndices();
    const auto& set     = inst->nbv->pairlistSets().pairlistSet(gmx::InteractionLocality::Local);
    const float rc2     = rc * rc;
    long long   n       = 0;

    auto testPair = [&](int gi, int gj, int shift) {
        const int ai = indices[gi], aj = indices[gj];
        if (ai < 0 || aj < 0)
        {
            return;
        }
        const float xi = inst->x[ai][XX] + inst->shiftVec[shift][XX];
        const float yi = inst->x[ai][YY] + inst->shiftVec[shift][YY];
        const float zi = inst->x[ai][ZZ] + inst->shiftVec[shift][ZZ];
        const float r2 = simdRsq(xi, yi, zi, inst->x[aj][XX], inst->x[aj][YY], inst->x[aj][ZZ]);
        if (r2 < rc2)
        {
            if (pairs && n < cap)
            {
                pairs[3 * n]     = ai;
                pairs[3 * n + 1] = aj;
                pairs[3 * n + 2] = shift;
            }
            n++;
        }
    };

    if (inst->kernelType == Nbnxm::KernelType::Cpu8x8x8_PlainC)
    {
        /* layout walk as kernels_reference/kernel_gpu_ref.cpp:127-250 */
        const NbnxnPairlistGpu* l = set.gpuList();
        constexpr int           cs = c_nbnxnGpuClusterSize;
        for (const nbnxn_sci_t& sci : l->sci)
        {
            for (int cj4Ind = sci.cj4_ind_start; cj4Ind < sci.cj4_ind_end; cj4Ind++)
            {
                const nbnxn_cj4_t& cj4 = l->cj4[cj4Ind];
                const auto&        e0  = l->excl[cj4.imei[0].excl_ind];
                const auto&        e1  = l->excl[cj4.imei[1].excl_ind];
                for (int jm = 0; jm < c_nbnxnGpuJgroupSize; jm++)
                {
                    const int cj = cj4.cj[jm];
                    for (int im = 0; im < c_nbnxnGpuNumClusterPerSupercluster; im++)
                    {
                        if (!((cj4.imei[0].imask >> (jm * c_nbnxnGpuNumClusterPerSupercluster + im)) & 1))
                        {
                            continue;
                        }
                        const int ci = sci.sci * c_nbnxnGpuNumClusterPerSupercluster + im;
                        for (int ii = 0; ii < cs; ii++)
                        {
                            for (int jj = 0; jj < cs; jj++)
                            {
                                const auto& ex = (jj < cs / 2) ? e0 : e1;
                                const int   bit = jm * c_nbnxnGpuNumClusterPerSupercluster + im;
                                /* kernel_gpu_ref.cpp:223-226: explicit half-list skip on the diagonal */
                                if ((sci.shift & NBNXN_CI_SHIFT) == CENTRAL && ci == cj
                                    && cj * cs + jj <= ci * cs + ii)
                                {
                                    continue;
                                }
                                if ((ex.pair[(jj & (cs / 2 - 1)) * cs + ii] >> bit) & 1)
                                {
                                    testPair(ci * cs + ii, cj * cs + jj, sci.shift & NBNXN_CI_SHIFT);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    else
    {
        for (const auto& l : set.cpuLists())
        {
            for (const nbnxn_ci_t& ci : l.ci)
            {
                const int shift = ci.shift & NBNXN_CI_SHIFT;
                for (int k = ci.cj_ind_start; k < ci.cj_ind_end; k++)
                {
                    const nbnxn_cj_t& cj = l.cj[k];
                    for (int ii = 0; ii < l.na_ci; ii++)
                    {
                        for (int jj = 0; jj < l.na_cj; jj++)
                        {
                            if ((cj.excl >> (ii * l.na_cj + jj)) & 1)
                            {
                                testPair(ci.ci * l.na_ci + ii, cj.cj * l.na_cj + jj, shift);
                            }
                        }
                    }
                }
            }
        }
    }
    return n;
}

/* The reference-built 8x8x8 list and grid-ordered atom data, as Nbnxm::gpu_init_pairlist / gpu_init_atomdata receive them
 * (nbnxm_gpu_data_mgmt.cpp:251-311): the inputs of b200nb_upload_pairlist / b200nb_set_grid_atoms in the parity tests of the
 * drop-in path.  Only for GMXREF_KERNEL_GPUREF_8X8X8 instances.  Sizes first (all pointers NULL), then the data. */
int gmxref_gpu_list(void* h, int* nsci, int* ncj4, int* nexcl, int* nslots, int* sci, int* cj4, unsigned* excl, float* xq, int* type)
{
    auto* inst = static_cast<Instance*>(h);
    if (inst->kernelType != Nbnxm::KernelType::Cpu8x8x8_PlainC) return 1;
    const NbnxnPairlistGpu* l    = inst->nbv->pairlistSets().pairlistSet(gmx::InteractionLocality::Local).gpuList();
    const nbnxn_atomdata_t& nbat = *inst->nbv->nbat;
    *nsci   = static_cast<int>(l->sci.size());
    *ncj4   = static_cast<int>(l->cj4.size());
    *nexcl  = static_cast<int>(l->excl.size());
    *nslots = nbat.numAtoms();
    static_assert(sizeof(nbnxn_sci_t) == 16 && sizeof(nbnxn_cj4_t) == 32 && sizeof(nbnxn_excl_t) == 128, "list element layout");
    if (sci) std::memcpy(sci, l->sci.data(), l->sci.size() * sizeof(nbnxn_sci_t));
    if (cj4) std::memcpy(cj4, l->cj4.data(), l->cj4.size() * sizeof(nbnxn_cj4_t));
    if (excl) std::memcpy(excl, l->excl.data(), l->excl.size() * sizeof(nbnxn_excl_t));
    if (xq)
    {
        if (nbat.XFormat != nbatXYZQ) return 2;
        std::memcpy(xq, nbat.x().data(), static_cast<size_t>(nbat.numAtoms()) * 4 * sizeof(float));
    }
    if (type) std::memcpy(type, nbat.params().type.data(), static_cast<size_t>(nbat.numAtoms()) * sizeof(int));
    return 0;
}

/* The reference's benchmark water (benchmark/bench_coords.h:47-49: 1000 SPC/E molecules equilibrated at 300 K, 1 bar, cubic box
 * 3.10736 nm), the base tile of BenchmarkSystem (bench_system.cpp:90-151).  out: 3000 x 3 floats; returns the atom count. */
int gmxref_bench_coordinates1000(float* out, int cap_atoms, float* box_edge)
{
    const int n = static_cast<int>(coordinates1000.size());
    *box_edge   = box1000[XX][XX];
    if (out == nullptr || cap_atoms < n) return n;
    for (int i = 0; i < n; i++)
        for (int d = 0; d < DIM; d++) out[3 * i + d] = coordinates1000[i][d];
    return n;
}

/* The Ewald correction force table of this instance (interaction_const_t::coulombEwaldTables: tableF and its scale), what
 * init_ewald_coulomb_force_table uploads for the reference's tabulated GPU kernels (nbnxm_gpu_data_mgmt.cpp:71-83).
 * Returns the number of table points (0: no Ewald tables); writes at most cap. */
int gmxref_ewald_table(void* h, float* table_f, int cap, float* scale)
{
    auto* inst = static_cast<Instance*>(h);
    if (!inst->ic.coulombEwaldTables) return 0;
    const auto& t = *inst->ic.coulombEwaldTables;
    const int   n = static_cast<int>(t.tableF.size());
    *scale        = t.scale;
    if (table_f)
        for (int i = 0; i < n && i < cap; i++) table_f[i] = t.tableF[i];
    return n;
}

/* forces of the last gmxref_compute in GRID order (nbat->out[0].f, 3 floats per slot): what gpu_launch_cpyback delivers */
int gmxref_grid_forces(void* h, float* f, int cap_slots)
{
    auto*                   inst = static_cast<Instance*>(h);
    const nbnxn_atomdata_t& nbat = *inst->nbv->nbat;
    const int               n    = nbat.numAtoms();
    if (n > cap_slots) return -1;
    std::memcpy(f, nbat.out[0].f.data(), static_cast<size_t>(n) * 3 * sizeof(float));
    return n;
}

} // extern "C"


/* ---- the reference's free-energy kernel on a caller-built perturbed pair list ---------------------------------------------------------- */
#include "gromacs/gmxlib/nonbonded/nb_free_energy.h"
#include "gromacs/gmxlib/nonbonded/nonbonded.h"
#include "gromacs/mdtypes/forceoutput.h"
#include "gromacs/mdtypes/inputrec.h"
#include "gromacs/mdtypes/mdatom.h"
#include "gromacs/mdtypes/nblist.h"
#include "gromacs/utility/arrayref.h"

int gmxref_fep_kernel(int natoms, const float* x, const float* shift_vec, int ntypes, const float* nbfp, const int* typeA, const int* typeB,
                      const float* qA, const float* qB, int nri, const int* iinr, const int* shift, const int* jindex, const int* jjnr,
                      const char* excl_fep, const gmxref_fep_params* p, float* f, float* fshift, float* out4)
{
    /* interaction constants as init_interaction_const would store them (mdlib/forcerec.cpp:850-874) */
    interaction_const_t ic;
    ic.eeltype          = (p->ewaldcoeff > 0.0f) ? eelPME : ((p->k_rf != 0.0f) ? eelRF : eelCUT);
    ic.coulomb_modifier = (p->ewaldcoeff > 0.0f) ? eintmodPOTSHIFT : eintmodNONE;
    ic.vdwtype          = evdwCUT;
    ic.vdw_modifier     = p->rvdw_switch > 0.0f ? eintmodPOTSWITCH : eintmodPOTSHIFT;
    ic.rvdw_switch      = p->rvdw_switch;
    ic.rcoulomb = ic.rvdw = p->rc;
    if (p->rvdw > 0.0f) ic.rvdw = p->rvdw;
    ic.epsfac             = p->epsfac;
    ic.k_rf               = p->k_rf;
    ic.c_rf               = p->c_rf;
    ic.dispersion_shift.cpot = p->disp_cpot;
    ic.repulsion_shift.cpot  = p->rep_cpot;
    const bool ljpme = p->ljpme_comb_rule != 0;
    if (ljpme)
    {
        ic.vdwtype       = evdwPME;
        ic.ewaldcoeff_lj = p->ewaldcoeff_lj;
        ic.sh_lj_ewald   = p->sh_lj_ewald;
    }
    if (ic.eeltype == eelPME || ljpme)
    {
        /* mdlib/forcerec.cpp:724-763 init_ewald_f_table: one spacing for both tables, obeying both accuracy requirements */
        if (ic.eeltype == eelPME) ic.ewaldcoeff_q = p->ewaldcoeff, ic.sh_ewald = p->sh_ewald;
        const real tableScale = ewald_spline3_table_scale(ic, ic.eeltype == eelPME, ljpme);
        const int  tableSize  = static_cast<int>(ic.rcoulomb * tableScale) + 2;
        if (ic.eeltype == eelPME)
        {
            ic.coulombEwaldTables  = std::make_unique<EwaldCorrectionTables>();
            *ic.coulombEwaldTables = generateEwaldCorrectionTables(tableSize, tableScale, ic.ewaldcoeff_q, v_q_ewald_lr);
        }
        if (ljpme)
        {
            ic.vdwEwaldTables  = std::make_unique<EwaldCorrectionTables>();
            *ic.vdwEwaldTables = generateEwaldCorrectionTables(tableSize, tableScale, ic.ewaldcoeff_lj, v_lj_ewald_lr);
        }
    }
    t_lambda fepvals{};
    fepvals.sc_alpha     = p->sc_alpha;
    fepvals.sc_power     = p->sc_power;
    fepvals.sc_r_power   = 6.0;
    fepvals.sc_sigma     = p->sc_sigma;
    fepvals.sc_sigma_min = p->sc_sigma_min;
    fepvals.bScCoul      = p->sc_coul != 0;
    ic.softCoreParameters = std::make_unique<interaction_const_t::SoftCoreParameters>(fepvals);

    std::vector<gmx::RVec> shiftVec(SHIFTS);
    for (int s = 0; s < SHIFTS; s++) shiftVec[s] = { shift_vec[3 * s], shift_vec[3 * s + 1], shift_vec[3 * s + 2] };
    t_forcerec* fr = static_cast<t_forcerec*>(std::calloc(1, sizeof(t_forcerec)));
    fr->ic         = &ic;
    fr->shift_vec  = as_rvec_array(shiftVec.data());
    fr->ntype      = ntypes;
    new (&fr->nbfp) std::vector<real>(nbfp, nbfp + 2 * ntypes * ntypes);
    fr->use_simd_kernels = FALSE;
    /* the grid C6 per type pair, at the positions of C6 in nbfp: make_ljpme_c6grid (static in mdlib/forcerec.cpp:157-195), from the
     * diagonal of nbfp (which holds 6 C6 and 12 C12, forcerec.cpp mk_nbfp) */
    std::vector<real> c6grid;
    if (ljpme)
    {
        fr->ljpme_combination_rule = p->ljpme_comb_rule == 2 ? eljpmeLB : eljpmeGEOM;
        c6grid.assign(2 * static_cast<size_t>(ntypes) * ntypes, 0);
        for (int i = 0; i < ntypes; i++)
        {
            for (int j = 0; j < ntypes; j++)
            {
                const real c6i = nbfp[2 * (i * ntypes + i)] / 6.0, c12i = nbfp[2 * (i * ntypes + i) + 1] / 12.0;
                const real c6j = nbfp[2 * (j * ntypes + j)] / 6.0, c12j = nbfp[2 * (j * ntypes + j) + 1] / 12.0;
                real       c6  = std::sqrt(c6i * c6j);
                if (fr->ljpme_combination_rule == eljpmeLB && !gmx_numzero(c6) && !gmx_numzero(c12i) && !gmx_numzero(c12j))
                {
                    const real sigmai = gmx::sixthroot(c12i / c6i), sigmaj = gmx::sixthroot(c12j / c6j);
                    const real epsi = c6i * c6i / c12i, epsj = c6j * c6j / c12j;
                    c6 = std::sqrt(epsi * epsj) * gmx::power6(0.5 * (sigmai + sigmaj));
                }
                c6grid[2 * (ntypes * i + j)] = c6 * 6.0;
            }
        }
        fr->ljpme_c6grid = c6grid.data();
    }

    std::vector<real> cA(qA, qA + natoms), cB(qB, qB + natoms);
    std::vector<int>  tA(typeA, typeA + natoms), tB(typeB, typeB + natoms);
    t_mdatoms md{};
    md.chargeA = cA.data();
    md.chargeB = cB.data();
    md.typeA   = tA.data();
    md.typeB   = tB.data();

    std::vector<int>  vi(iinr, iinr + nri), vs(shift, shift + nri), vg(nri, 0), vj(jindex, jindex + nri + 1), vjj(jjnr, jjnr + jindex[nri]);
    std::vector<char> ve(excl_fep, excl_fep + jindex[nri]);
    t_nblist nl{};
    nl.nri = nl.maxnri = nri;
    nl.nrj = nl.maxnrj = jindex[nri];
    nl.iinr     = vi.data();
    nl.gid      = vg.data();
    nl.shift    = vs.data();
    nl.jindex   = vj.data();
    nl.jjnr     = vjj.data();
    nl.excl_fep = ve.data();

    gmx::PaddedVector<gmx::RVec> xx(natoms), ff(natoms, { 0, 0, 0 });
    for (int a = 0; a < natoms; a++) xx[a] = { x[3 * a], x[3 * a + 1], x[3 * a + 2] };
    std::vector<gmx::RVec>     fsh(SHIFTS, { 0, 0, 0 });
    gmx::ForceWithShiftForces forces(ff.arrayRefWithPadding(), true, fsh);

    real lambda[efptNR] = { 0 }, dvdl[efptNR] = { 0 }, vc = 0, vv = 0;
    lambda[efptCOUL]    = p->lambda_coul;
    lambda[efptVDW]     = p->lambda_vdw;
    nb_kernel_data_t kd{};
    kd.flags          = GMX_NONBONDED_DO_FORCE | GMX_NONBONDED_DO_SHIFTFORCE | GMX_NONBONDED_DO_POTENTIAL | GMX_NONBONDED_DO_SR;
    kd.lambda         = lambda;
    kd.dvdl           = dvdl;
    kd.energygrp_elec = &vc;
    kd.energygrp_vdw  = &vv;
    t_nrnb nrnb{};
    gmx_nb_free_energy_kernel(&nl, as_rvec_array(xx.data()), &forces, fr, &md, &kd, &nrnb);

    for (int a = 0; a < natoms; a++)
        for (int d = 0; d < 3; d++) f[3 * a + d] = ff[a][d];
    for (int s = 0; s < SHIFTS; s++)
        for (int d = 0; d < 3; d++) fshift[3 * s + d] = fsh[s][d];
    out4[0] = vc, out4[1] = vv, out4[2] = dvdl[efptCOUL], out4[3] = dvdl[efptVDW];
    fr->nbfp.~vector();
    std::free(fr);
    return 0;
}

/* ---- listed ("bonded") interactions the reference's GPU bonded module covers (listed_forces/gpubonded.h:84-85 fTypesOnGpu):
 * the reference's CPU functions for the same types (listed_forces/bonded.cpp calculateSimpleBond, pairs.cpp do_pairs), compiled
 * from the tree.  TEST INFRASTRUCTURE: pins oracle/nbnxm_oracle.c orc_bonded. ---- */
#include "gromacs/listed_forces/bonded.h"
#include "gromacs/listed_forces/pairs.h"
#include "gromacs/topology/idef.h"
#include "gromacs/topology/ifunc.h"

extern "C" int gmxref_bonded(int kind, int nbonds, const int* iatoms, int nparams, const float* params6, int natoms, const float* x,
                             const float* q, const float* box9, float epsfac_fudge, int want_virial_energy, float* f, float* fshift,
                             double* energy2)
{
    static const int ftypes[GMXREF_BONDED_KINDS] = { F_BONDS, F_ANGLES, F_UREY_BRADLEY, F_PDIHS, F_RBDIHS, F_IDIHS, F_PIDIHS, F_LJ14 };
    if (kind < 0 || kind >= GMXREF_BONDED_KINDS) return -1;
    const int                ftype = ftypes[kind];
    const int                nral  = interaction_function[ftype].nratoms;
    std::vector<t_iparams>   ip(nparams);
    for (int t = 0; t < nparams; t++)
    {
        const float* p = params6 + 6 * t;
        std::memset(&ip[t], 0, sizeof(t_iparams));
        switch (ftype)
        {
            case F_BONDS:
            case F_ANGLES:
            case F_IDIHS:
                ip[t].harmonic.rA = ip[t].harmonic.rB = p[0];
                ip[t].harmonic.krA = ip[t].harmonic.krB = p[1];
                break;
            case F_UREY_BRADLEY:
                ip[t].u_b.thetaA = ip[t].u_b.thetaB = p[0];
                ip[t].u_b.kthetaA = ip[t].u_b.kthetaB = p[1];
                ip[t].u_b.r13A = ip[t].u_b.r13B = p[2];
                ip[t].u_b.kUBA = ip[t].u_b.kUBB = p[3];
                break;
            case F_PDIHS:
            case F_PIDIHS:
                ip[t].pdihs.phiA = ip[t].pdihs.phiB = p[0];
                ip[t].pdihs.cpA = ip[t].pdihs.cpB = p[1];
                ip[t].pdihs.mult                  = (int)p[2];
                break;
            case F_RBDIHS:
                for (int k = 0; k < 6; k++) ip[t].rbdihs.rbcA[k] = ip[t].rbdihs.rbcB[k] = p[k];
                break;
            case F_LJ14:
                ip[t].lj14.c6A = ip[t].lj14.c6B = p[0];
                ip[t].lj14.c12A = ip[t].lj14.c12B = p[1];
                break;
        }
    }
    matrix box;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) box[i][j] = box9[3 * i + j];
    t_pbc pbc;
    set_pbc(&pbc, PbcType::Xyz, box);
    std::vector<gmx::RVec> xv(natoms);
    for (int a = 0; a < natoms; a++) xv[a] = { x[3 * a], x[3 * a + 1], x[3 * a + 2] };
    std::vector<real> f4((size_t)natoms * 4, 0.0f);
    std::vector<real> fs(SHIFTS * 3, 0.0f);
    std::vector<int>  glob(natoms);
    for (int a = 0; a < natoms; a++) glob[a] = a;
    t_mdatoms         md{};
    std::vector<real> qa(q, q + natoms);
    md.chargeA = qa.data();
    md.chargeB = qa.data();
    energy2[0] = energy2[1] = 0.0;
    real dvdl               = 0;
    if (ftype == F_LJ14)
    {
        /* the plain analytical code path (pairs.cpp:636-676: no tables) -- forces only; the reference computes 1-4 energies and
         * the virial from its spline tables on the CPU and analytically on the GPU (gpubondedkernels.cu:658-718) */
        interaction_const_t ic;
        ic.vdwtype = evdwCUT;
        ic.eeltype = eelCUT;
        ic.epsfac  = epsfac_fudge;
        t_forcerec* fr = static_cast<t_forcerec*>(std::calloc(1, sizeof(t_forcerec)));
        fr->ic               = &ic;
        fr->fudgeQQ          = 1.0f;
        fr->use_simd_kernels = FALSE;
        gmx::StepWorkload stepWork;
        real              lambda[efptNR] = { 0 };
        real              dvdl4[efptNR]  = { 0 };
        do_pairs(F_LJ14, nbonds * (nral + 1), iatoms, ip.data(), as_rvec_array(xv.data()), reinterpret_cast<rvec4*>(f4.data()),
                 reinterpret_cast<rvec*>(fs.data()), &pbc, lambda, dvdl4, &md, fr, false, stepWork, nullptr, glob.data());
        std::free(fr);
    }
    else
    {
        const BondedKernelFlavor flavor = want_virial_energy ? BondedKernelFlavor::ForcesAndVirialAndEnergy : BondedKernelFlavor::ForcesNoSimd;
        energy2[0] = calculateSimpleBond(ftype, nbonds * (nral + 1), iatoms, ip.data(), as_rvec_array(xv.data()), reinterpret_cast<rvec4*>(f4.data()),
                                         reinterpret_cast<rvec*>(fs.data()), &pbc, 0.0f, &dvdl, &md, nullptr, glob.data(), flavor);
    }
    for (int a = 0; a < natoms; a++)
        for (int d = 0; d < 3; d++) f[3 * a + d] = f4[4 * a + d];
    for (int k = 0; k < SHIFTS * 3; k++) fshift[k] = fs[k];
    return 0;
}


extern "C" int gmxref_fep_list(void* h, int* nri, int* nrj, int cap_nri, int cap_nrj, int* iinr, int* shift, int* jindex, int* jjnr, char* excl_fep)
{
    auto*      inst  = static_cast<Instance*>(h);
    const auto lists = inst->nbv->pairlistSets().pairlistSet(gmx::InteractionLocality::Local).fepLists();
    int        ni = 0, nj = 0;
    for (const auto& l : lists) ni += l->nri, nj += l->nrj;
    *nri = ni, *nrj = nj;
    if (cap_nri == 0 && cap_nrj == 0) return 0;
    if (cap_nri < ni || cap_nrj < nj) return -1;
    int ki = 0, kj = 0;
    for (const auto& l : lists)
    {
        for (int n = 0; n < l->nri; n++)
        {
            iinr[ki] = l->iinr[n], shift[ki] = l->shift[n], jindex[ki] = kj;
            for (int k = l->jindex[n]; k < l->jindex[n + 1]; k++) jjnr[kj] = l->jjnr[k], excl_fep[kj] = l->excl_fep[k], kj++;
            ki++;
        }
    }
    jindex[ki] = kj;
    return 0;
}
