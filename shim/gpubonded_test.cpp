/* gmx::GpuBonded on libb200nb (shim/gpubonded_b200.cpp) driven the way do_force() drives it (mdlib/sim_util.cpp:1340-1356,
 * 1421-1426, 1548-1551, 959-967), next to the reference's own CPU functions for the same interactions (listed_forces/bonded.cpp
 * calculateSimpleBond, pairs.cpp do_pairs) in the same process:
 *   nonbonded_verlet_t with KernelType::Gpu8x8x8 and the Nbnxm::gpu_* shim (set up as shim/nbnxm_bench_gpu.cpp does) on the
 *   reference's BenchmarkSystem(1) (3000 atoms of water) | GpuBonded(ffparams, epsfac * fudgeQQ, ...) |
 *   updateInteractionListsAndDeviceBuffers(nbv->getGridIndices(), idef, gpu_get_xq, gpu_get_f, gpu_get_fshift) |
 *   per step: gpu_clear_outputs, gpu_copy_xq_to_gpu, setPbcAndlaunchKernel, dispatchNonbondedKernel, gpu_launch_cpyback,
 *   launchEnergyTransfer, gpu_wait_finish_task, atomdata_add_nbat_f_to_f, waitAccumulateEnergyTerms + clearEnergies.
 * The listed interactions are laid over the water molecules (bonds / angle / Urey-Bradley inside a molecule, torsions and 1-4
 * pairs across consecutive molecules): not a force field, but valid input for both implementations, all eight types, with many
 * interactions across the periodic boundary.
 * The forces of the step with the bonded kernel minus the forces of the step without it are compared with the CPU functions'
 * forces; energies per type and the shift forces likewise.  Prints one JSON object; exit code 0 iff everything agrees.
 * GPUBONDED_TEST_CPU_ONLY=1: the CPU leg alone (the build check on a machine without a GPU). */
#include "gmxpre.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <memory>
#include <vector>

#include <omp.h>

#include "gromacs/gpu_utils/device_stream_manager.h"
#include "gromacs/hardware/device_information.h"
#include "gromacs/gmxlib/nrnb.h"
#include "gromacs/listed_forces/bonded.h"
#include "gromacs/listed_forces/gpubonded.h"
#include "gromacs/listed_forces/pairs.h"
#include "gromacs/math/units.h"
#include "gromacs/mdlib/gmx_omp_nthreads.h"
#include "gromacs/mdtypes/enerdata.h"
#include "gromacs/mdtypes/forcerec.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/mdatom.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/nbnxm/atomdata.h"
#include "gromacs/nbnxm/benchmark/bench_system.h"
#include "gromacs/nbnxm/gpu_data_mgmt.h"
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/nbnxm_gpu.h"
#include "gromacs/nbnxm/pairlistset.h"
#include "gromacs/nbnxm/pairlistsets.h"
#include "gromacs/nbnxm/pairsearch.h"
#include "gromacs/pbcutil/ishift.h"
#include "gromacs/pbcutil/pbc.h"
#include "gromacs/topology/forcefieldparameters.h"
#include "gromacs/topology/idef.h"
#include "gromacs/topology/ifunc.h"
#include "gromacs/utility/logger.h"

namespace
{

constexpr real c_cutoff  = 0.9;
constexpr real c_fudgeQQ = 0.5;

gmx::DeviceStreamManager& streamManager()
{
    static DeviceInformation        deviceInfo{};
    static gmx::DeviceStreamManager manager(deviceInfo, false, gmx::SimulationWorkload(), false);
    return manager;
}

/* reaction field, as shim/nbnxm_bench_gpu.cpp setupInteractionConst */
void setupInteractionConst(interaction_const_t* ic)
{
    ic->vdwtype               = evdwCUT;
    ic->vdw_modifier          = eintmodPOTSHIFT;
    ic->rvdw                  = c_cutoff;
    ic->eeltype               = eelRF;
    ic->coulomb_modifier      = eintmodPOTSHIFT;
    ic->rcoulomb              = c_cutoff;
    ic->epsfac                = ONE_4PI_EPS0;
    ic->k_rf                  = 0.5 * std::pow(ic->rcoulomb, -3);
    ic->c_rf                  = 1 / ic->rcoulomb + ic->k_rf * ic->rcoulomb * ic->rcoulomb;
    ic->dispersion_shift.cpot = -1.0 / gmx::power6(ic->rvdw);
    ic->repulsion_shift.cpot  = -1.0 / gmx::power12(ic->rvdw);
}

std::unique_ptr<nonbonded_verlet_t> setupNbnxmGpu(const gmx::BenchmarkSystem& system, const interaction_const_t& ic)
{
    const auto         pinPolicy = gmx::PinningPolicy::PinnedIfSupported;
    Nbnxm::KernelSetup kernelSetup;
    kernelSetup.kernelType         = Nbnxm::KernelType::Gpu8x8x8;
    kernelSetup.ewaldExclusionType = Nbnxm::EwaldExclusionType::Analytical;
    PairlistParams pairlistParams(kernelSetup.kernelType, false, c_cutoff, false);
    auto           pairlistSets = std::make_unique<PairlistSets>(pairlistParams, false, 0);
    auto pairSearch = std::make_unique<PairSearch>(PbcType::Xyz, false, nullptr, nullptr, pairlistParams.pairlistType, false, 1, pinPolicy);
    auto atomData   = std::make_unique<nbnxn_atomdata_t>(pinPolicy);
    nbnxn_atomdata_init(gmx::MDLogger(), atomData.get(), kernelSetup.kernelType, 0, system.numAtomTypes, system.nonbondedParameters, 1, 1);
    NbnxmGpu* gpuNbv = Nbnxm::gpu_init(streamManager(), &ic, pairlistParams, atomData.get(), false);
    auto nbv = std::make_unique<nonbonded_verlet_t>(std::move(pairlistSets), std::move(pairSearch), std::move(atomData), kernelSetup, gpuNbv, nullptr);
    t_nrnb     nrnb;
    const rvec lowerCorner = { 0, 0, 0 };
    const rvec upperCorner = { system.box[XX][XX], system.box[YY][YY], system.box[ZZ][ZZ] };
    const real atomDensity = system.coordinates.size() / det(system.box);
    nbnxn_put_on_grid(nbv.get(), system.box, 0, lowerCorner, upperCorner, nullptr, { 0, int(system.coordinates.size()) }, atomDensity,
                      system.atomInfoAllVdw, system.coordinates, 0, nullptr);
    nbv->setAtomProperties(system.atomTypes, system.charges, system.atomInfoAllVdw);
    nbnxn_atomdata_copy_shiftvec(false, system.forceRec.shift_vec, nbv->nbat.get());
    Nbnxm::gpu_init_atomdata(nbv->gpu_nbv, nbv->nbat.get());
    nbv->constructPairlist(gmx::InteractionLocality::Local, system.excls, 0, &nrnb);
    return nbv;
}

/* one step of the nonbonded + (optionally) bonded GPU work, in do_force()'s order */
void step(nonbonded_verlet_t* nbv, gmx::GpuBonded* gpuBonded, const gmx::BenchmarkSystem& system, const interaction_const_t& ic,
          const gmx::StepWorkload& stepWork, gmx_enerdata_t* enerd, std::vector<gmx::RVec>* f, std::vector<gmx::RVec>* shiftForces)
{
    t_nrnb nrnb = { 0 };
    nbv->convertCoordinates(gmx::AtomLocality::Local, false, system.coordinates);
    std::fill(f->begin(), f->end(), gmx::RVec{ 0, 0, 0 });
    std::fill(shiftForces->begin(), shiftForces->end(), gmx::RVec{ 0, 0, 0 });
    Nbnxm::gpu_clear_outputs(nbv->gpu_nbv, stepWork.computeVirial);
    Nbnxm::gpu_copy_xq_to_gpu(nbv->gpu_nbv, nbv->nbat.get(), gmx::AtomLocality::Local);
    if (gpuBonded && gpuBonded->haveInteractions()) gpuBonded->setPbcAndlaunchKernel(PbcType::Xyz, system.box, true, stepWork);
    nbv->dispatchNonbondedKernel(gmx::InteractionLocality::Local, ic, stepWork, enbvClearFYes, system.forceRec, enerd, &nrnb);
    Nbnxm::gpu_launch_cpyback(nbv->gpu_nbv, nbv->nbat.get(), stepWork, gmx::AtomLocality::Local);
    if (gpuBonded && gpuBonded->haveInteractions() && stepWork.computeEnergy) gpuBonded->launchEnergyTransfer();
    Nbnxm::gpu_wait_finish_task(nbv->gpu_nbv, stepWork, gmx::AtomLocality::Local, enerd->grpp.ener[egLJSR].data(),
                                enerd->grpp.ener[egCOULSR].data(), *shiftForces, nullptr);
    nbv->atomdata_add_nbat_f_to_f(gmx::AtomLocality::All, *f);
    if (gpuBonded && gpuBonded->haveInteractions() && stepWork.computeEnergy)
    {
        gpuBonded->waitAccumulateEnergyTerms(enerd);
        gpuBonded->clearEnergies();
    }
}

t_iparams harmonic(real r, real k)
{
    t_iparams p;
    std::memset(&p, 0, sizeof(p));
    p.harmonic.rA = p.harmonic.rB = r;
    p.harmonic.krA = p.harmonic.krB = k;
    return p;
}

} // namespace

int main()
{
    gmx_omp_nthreads_set(emntPairsearch, 1);
    gmx_omp_nthreads_set(emntNonbonded, 1);
    gmx_omp_nthreads_set(emntDefault, 1);
    gmx_omp_nthreads_set(emntBonded, 1);
    const bool cpuOnly = std::getenv("GPUBONDED_TEST_CPU_ONLY") != nullptr;

    gmx::BenchmarkSystem system(1);
    const int            n    = int(system.coordinates.size());
    const int            nmol = n / 3;
    interaction_const_t  ic;
    setupInteractionConst(&ic);

    /* parameter sets, one per interaction type of the GPU bonded module */
    gmx_ffparams_t ffparams;
    ffparams.fudgeQQ = c_fudgeQQ;
    auto add = [&](int ftype, const t_iparams& p) {
        ffparams.functype.push_back(ftype);
        ffparams.iparams.push_back(p);
        return int(ffparams.functype.size()) - 1;
    };
    t_iparams p;
    const int tBond = add(F_BONDS, harmonic(0.09, 3.45e5)); /* the water of the benchmark is rigid at 0.1 nm / 109.47 degrees */
    const int tAng  = add(F_ANGLES, harmonic(104.0, 383.0));
    std::memset(&p, 0, sizeof(p));
    p.u_b.thetaA = p.u_b.thetaB = 104.5, p.u_b.kthetaA = p.u_b.kthetaB = 300.0, p.u_b.r13A = p.u_b.r13B = 0.15, p.u_b.kUBA = p.u_b.kUBB = 2.0e4;
    const int tUB = add(F_UREY_BRADLEY, p);
    std::memset(&p, 0, sizeof(p));
    p.pdihs.phiA = p.pdihs.phiB = 30.0, p.pdihs.cpA = p.pdihs.cpB = 5.0, p.pdihs.mult = 3;
    const int tPdih = add(F_PDIHS, p);
    std::memset(&p, 0, sizeof(p));
    const real rb[NR_RBDIHS] = { 9.28, 12.16, -13.12, -3.06, 26.24, -31.5 };
    for (int k = 0; k < NR_RBDIHS; k++) p.rbdihs.rbcA[k] = p.rbdihs.rbcB[k] = rb[k];
    const int tRb   = add(F_RBDIHS, p);
    const int tIdih = add(F_IDIHS, harmonic(10.0, 100.0));
    std::memset(&p, 0, sizeof(p));
    p.pdihs.phiA = p.pdihs.phiB = 180.0, p.pdihs.cpA = p.pdihs.cpB = 4.6, p.pdihs.mult = 2;
    const int tPidih = add(F_PIDIHS, p);
    std::memset(&p, 0, sizeof(p));
    p.lj14.c6A = p.lj14.c6B = 1.3e-3, p.lj14.c12A = p.lj14.c12B = 1.3e-6;
    const int tLj14 = add(F_LJ14, p);

    InteractionDefinitions idef(ffparams);
    for (int m = 0; m < nmol; m++)
    {
        const int o = 3 * m, h1 = o + 1, h2 = o + 2;
        idef.il[F_BONDS].push_back(tBond, std::array<int, 2>{ o, h1 });
        idef.il[F_BONDS].push_back(tBond, std::array<int, 2>{ o, h2 });
        if (m % 2 == 0) idef.il[F_ANGLES].push_back(tAng, std::array<int, 3>{ h1, o, h2 });
        else idef.il[F_UREY_BRADLEY].push_back(tUB, std::array<int, 3>{ h1, o, h2 });
        if (m + 1 < nmol)
        {
            const int o2 = o + 3, h3 = o + 4;
            switch (m % 4)
            {
                case 0: idef.il[F_PDIHS].push_back(tPdih, std::array<int, 4>{ h1, o, o2, h3 }); break;
                case 1: idef.il[F_RBDIHS].push_back(tRb, std::array<int, 4>{ h1, o, o2, h3 }); break;
                case 2: idef.il[F_IDIHS].push_back(tIdih, std::array<int, 4>{ o, h1, h2, o2 }); break;
                default: idef.il[F_PIDIHS].push_back(tPidih, std::array<int, 4>{ o, h1, h2, o2 }); break;
            }
            idef.il[F_LJ14].push_back(tLj14, std::array<int, 2>{ h1, h3 });
        }
    }

    /* ---- the reference's CPU functions ---- */
    t_pbc pbc;
    set_pbc(&pbc, PbcType::Xyz, system.box);
    std::vector<real> f4(size_t(n) * 4, 0.0F);
    std::vector<real> fsCpu(SHIFTS * 3, 0.0F);
    std::vector<int>  glob(n);
    for (int a = 0; a < n; a++) glob[a] = a;
    t_mdatoms md{};
    md.chargeA = const_cast<real*>(system.charges.data());
    md.chargeB = md.chargeA;
    double eCpu[F_NRE] = { 0 };
    real   dvdl        = 0;
    for (int ftype : gmx::fTypesOnGpu)
    {
        const InteractionList& il = idef.il[ftype];
        if (il.empty()) continue;
        if (ftype == F_LJ14)
        {
            /* the analytical code path (no tables): forces only; energies and virial come from tables on the CPU */
            t_forcerec* fr       = static_cast<t_forcerec*>(std::calloc(1, sizeof(t_forcerec)));
            fr->ic               = &ic;
            fr->fudgeQQ          = c_fudgeQQ;
            fr->use_simd_kernels = FALSE;
            gmx::StepWorkload forceOnly;
            real              lambda[efptNR] = { 0 }, dvdl4[efptNR] = { 0 };
            do_pairs(F_LJ14, il.size(), il.iatoms.data(), ffparams.iparams.data(), as_rvec_array(system.coordinates.data()),
                     reinterpret_cast<rvec4*>(f4.data()), reinterpret_cast<rvec*>(fsCpu.data()), &pbc, lambda, dvdl4, &md, fr, false, forceOnly,
                     nullptr, glob.data());
            std::free(fr);
        }
        else
        {
            eCpu[ftype] = calculateSimpleBond(ftype, il.size(), il.iatoms.data(), ffparams.iparams.data(), as_rvec_array(system.coordinates.data()),
                                              reinterpret_cast<rvec4*>(f4.data()), reinterpret_cast<rvec*>(fsCpu.data()), &pbc, 0.0F, &dvdl, &md,
                                              nullptr, glob.data(), BondedKernelFlavor::ForcesAndVirialAndEnergy);
        }
    }
    double fCpuSq = 0;
    for (int a = 0; a < n; a++)
        for (int d = 0; d < 3; d++) fCpuSq += double(f4[4 * a + d]) * f4[4 * a + d];
    if (cpuOnly)
    {
        std::printf("{\"atoms\": %d, \"bonds\": %d, \"cpu_force_sumsq\": %.6e, \"cpu_e_bonds\": %.6e}\n", n, idef.il[F_BONDS].size() / 3, fCpuSq,
                    eCpu[F_BONDS]);
        return (fCpuSq > 0 && std::isfinite(fCpuSq)) ? 0 : 1;
    }

    /* ---- the same through GpuBonded on the nonbonded module's device buffers ---- */
    gmx::StepWorkload stepWork;
    stepWork.computeForces          = true;
    stepWork.computeNonbondedForces = true;
    stepWork.computeVirial          = true;
    stepWork.computeEnergy          = true;
    auto            nbv = setupNbnxmGpu(system, ic);
    gmx::GpuBonded  gpuBonded(ffparams, ic.epsfac * c_fudgeQQ, streamManager().context(),
                             streamManager().stream(gmx::DeviceStreamType::NonBondedLocal), nullptr);
    gpuBonded.updateInteractionListsAndDeviceBuffers(nbv->getGridIndices(), idef, Nbnxm::gpu_get_xq(nbv->gpu_nbv), Nbnxm::gpu_get_f(nbv->gpu_nbv),
                                                     Nbnxm::gpu_get_fshift(nbv->gpu_nbv));
    gmx_enerdata_t         enerdNb(1, 0), enerd(1, 0);
    std::vector<gmx::RVec> fNb(n), fAll(n), fsNb(SHIFTS), fsAll(SHIFTS);
    step(nbv.get(), nullptr, system, ic, stepWork, &enerdNb, &fNb, &fsNb);
    step(nbv.get(), &gpuBonded, system, ic, stepWork, &enerd, &fAll, &fsAll);
    /* force-only flavour of both kernels */
    gmx::StepWorkload forceOnly;
    forceOnly.computeForces          = true;
    forceOnly.computeNonbondedForces = true;
    gmx_enerdata_t         enerdF(1, 0);
    std::vector<gmx::RVec> fNbF(n), fAllF(n), fsDummy(SHIFTS);
    step(nbv.get(), nullptr, system, ic, forceOnly, &enerdF, &fNbF, &fsDummy);
    step(nbv.get(), &gpuBonded, system, ic, forceOnly, &enerdF, &fAllF, &fsDummy);

    /* the lists without the 1-4 pairs (a second update, as at a later search step): the CPU leg's shift forces hold every type
     * but those (their analytical CPU path is force-only), so this is the step the shift forces are compared on */
    idef.il[F_LJ14].clear();
    gpuBonded.updateInteractionListsAndDeviceBuffers(nbv->getGridIndices(), idef, Nbnxm::gpu_get_xq(nbv->gpu_nbv), Nbnxm::gpu_get_f(nbv->gpu_nbv),
                                                     Nbnxm::gpu_get_fshift(nbv->gpu_nbv));
    gmx_enerdata_t         enerdNo14(1, 0);
    std::vector<gmx::RVec> fNo14(n), fsNo14(SHIFTS);
    step(nbv.get(), &gpuBonded, system, ic, stepWork, &enerdNo14, &fNo14, &fsNo14);

    double num = 0, numF = 0, fsNum = 0, fsMax = 0;
    bool   finite = true;
    for (int a = 0; a < n; a++)
        for (int d = 0; d < 3; d++)
        {
            const double ref = f4[4 * a + d];
            const double g   = double(fAll[a][d]) - fNb[a][d], gF = double(fAllF[a][d]) - fNbF[a][d];
            finite           = finite && std::isfinite(g) && std::isfinite(gF);
            num += (g - ref) * (g - ref);
            numF += (gF - ref) * (gF - ref);
        }
    for (int s = 0; s < SHIFTS; s++)
        for (int d = 0; d < 3; d++)
        {
            if (s == CENTRAL) continue; /* the central shift carries no virial; b200nb does not accumulate it */
            fsMax = std::max(fsMax, std::fabs(double(fsCpu[3 * s + d])));
            fsNum = std::max(fsNum, std::fabs(double(fsNo14[s][d]) - fsNb[s][d] - fsCpu[3 * s + d]));
        }
    const double relRms = std::sqrt(num / fCpuSq), relRmsF = std::sqrt(numF / fCpuSq);
    double       eErr   = 0;
    std::printf("{\"atoms\": %d, \"force_rel_rms_gpu_vs_cpu\": %.3e, \"force_only_rel_rms_gpu_vs_cpu\": %.3e, \"energies\": {", n, relRms, relRmsF);
    bool first = true;
    for (int ftype : gmx::fTypesOnGpu)
    {
        if (ftype == F_LJ14) continue;
        const double g = enerd.term[ftype];
        eErr           = std::max(eErr, std::fabs(g - eCpu[ftype]) / std::max(1e-30, std::fabs(eCpu[ftype])));
        std::printf("%s\"%s\": [%.6e, %.6e]", first ? "" : ", ", interaction_function[ftype].name, g, eCpu[ftype]);
        first = false;
    }
    std::printf("}, \"energy_max_rel_err\": %.3e, \"lj14\": %.6e, \"coul14\": %.6e, \"fshift_max_abs_diff\": %.3e, \"fshift_max\": %.3e}\n", eErr,
                double(enerd.grpp.ener[egLJ14][0]), double(enerd.grpp.ener[egCOUL14][0]), fsNum, fsMax);
    const bool ok = finite && relRms < 1e-5 && relRmsF < 1e-5 && eErr < 2e-5 && fsNum <= 2e-5 * fsMax && enerd.grpp.ener[egLJ14][0] != 0
                    && enerdNo14.grpp.ener[egLJ14][0] == 0;
    return ok ? 0 : 1;
}
