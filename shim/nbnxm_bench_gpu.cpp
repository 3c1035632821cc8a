/* The reference's nonbonded benchmark protocol (`gmx nonbonded-benchmark`: nbnxm/benchmark/bench_setup.cpp:170-343) with the GPU
 * backend, in C++, through the UNMODIFIED nbnxm module and the Nbnxm::gpu_* shim (shim/nbnxm_b200.cpp -> libb200nb.so).
 *
 * The reference's own tool passes gpu_nbv = nullptr (bench_setup.cpp:198-199) and its own BenchmarkSystem, so this driver
 * repeats its set-up with KernelType::Gpu8x8x8 and Nbnxm::gpu_init, on the reference's own system generator
 * (gmx::BenchmarkSystem, nbnxm/benchmark/bench_system.cpp, compiled from the reference tree: BenchmarkSystem(32) = 96 000 atoms is
 * BASELINE configs[2]), and runs, per step, the nonbonded part of do_force() (mdlib/sim_util.cpp:1388-2035) for one rank:
 *   nbv->convertCoordinates | gpu_copy_xq_to_gpu | gpu_clear_outputs | nbv->dispatchNonbondedKernel (-> gpu_launch_kernel) |
 *   gpu_launch_cpyback | gpu_wait_finish_task | nbv->atomdata_add_nbat_f_to_f
 * next to the same step on the reference's CPU SIMD kernel (the default 2xMM / 4xM setup of the build) in the same process, and
 * compares the forces.  Grid and pair search run on the CPU in both (as for the reference's own CUDA backend).
 * Prints one JSON object.  usage: nbnxm_bench_gpu [sizeFactor=32] [pme|rf] [iterations=50] [threads=all]
 * Exit code 0 iff the forces agree within 1e-5 relative RMS. */
#include "gmxpre.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <memory>
#include <string>
#include <vector>

#include <omp.h>

#include "gromacs/ewald/ewald_utils.h"
#include "gromacs/gmxlib/nrnb.h"
#include "gromacs/gpu_utils/device_stream_manager.h"
#include "gromacs/hardware/device_information.h"
#include "gromacs/math/units.h"
#include "gromacs/mdlib/forcerec.h"
#include "gromacs/mdlib/gmx_omp_nthreads.h"
#include "gromacs/mdtypes/enerdata.h"
#include "gromacs/mdtypes/forcerec.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/nbnxm/atomdata.h"
#include "gromacs/nbnxm/benchmark/bench_system.h"
#include "gromacs/nbnxm/gpu_data_mgmt.h"
#include "gromacs/nbnxm/gridset.h"
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/nbnxm_gpu.h"
#include "gromacs/nbnxm/nbnxm_simd.h"
#include "gromacs/nbnxm/pairlistset.h"
#include "gromacs/nbnxm/pairlistsets.h"
#include "gromacs/nbnxm/pairsearch.h"
#include "gromacs/pbcutil/ishift.h"
#include "gromacs/pbcutil/pbc.h"
#include "gromacs/utility/logger.h"

namespace
{

constexpr real c_cutoff = 0.9; /* BASELINE configs[1-3]: rc = rlist = 0.9 nm */

double seconds()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/* setupInteractionConst, bench_setup.cpp:137-166 */
void setupInteractionConst(interaction_const_t* ic, bool pme)
{
    ic->vdwtype              = evdwCUT;
    ic->vdw_modifier         = eintmodPOTSHIFT;
    ic->rvdw                 = c_cutoff;
    ic->eeltype              = pme ? eelPME : eelRF;
    ic->coulomb_modifier     = eintmodPOTSHIFT;
    ic->rcoulomb             = c_cutoff;
    ic->epsfac               = ONE_4PI_EPS0;
    ic->k_rf                 = 0.5 * std::pow(ic->rcoulomb, -3);
    ic->c_rf                 = 1 / ic->rcoulomb + ic->k_rf * ic->rcoulomb * ic->rcoulomb;
    ic->dispersion_shift.cpot = -1.0 / gmx::power6(ic->rvdw);
    ic->repulsion_shift.cpot  = -1.0 / gmx::power12(ic->rvdw);
    if (pme)
    {
        ic->ewaldcoeff_q       = calc_ewaldcoeff_q(c_cutoff, 1e-5);
        ic->coulombEwaldTables = std::make_unique<EwaldCorrectionTables>();
        init_interaction_const_tables(nullptr, ic, 0);
    }
}

/* setupNbnxmForBenchInstance, bench_setup.cpp:170-232, with the kernel type and the GPU object as parameters */
std::unique_ptr<nonbonded_verlet_t> setupNbnxm(const gmx::BenchmarkSystem& system, const interaction_const_t& ic, bool useGpu, int numThreads)
{
    const auto pinPolicy = useGpu ? gmx::PinningPolicy::PinnedIfSupported : gmx::PinningPolicy::CannotBePinned;
    Nbnxm::KernelSetup kernelSetup;
    if (useGpu)
    {
        kernelSetup.kernelType = Nbnxm::KernelType::Gpu8x8x8;
    }
    else
    {
#ifdef GMX_NBNXN_SIMD_2XNN
        kernelSetup.kernelType = Nbnxm::KernelType::Cpu4xN_Simd_2xNN;
#else
        kernelSetup.kernelType = Nbnxm::KernelType::Cpu4xN_Simd_4xN;
#endif
    }
    kernelSetup.ewaldExclusionType = Nbnxm::EwaldExclusionType::Analytical;
    PairlistParams pairlistParams(kernelSetup.kernelType, false, c_cutoff, false);
    auto           pairlistSets = std::make_unique<PairlistSets>(pairlistParams, false, 0);
    auto pairSearch = std::make_unique<PairSearch>(PbcType::Xyz, false, nullptr, nullptr, pairlistParams.pairlistType, false, numThreads, pinPolicy);
    auto atomData   = std::make_unique<nbnxn_atomdata_t>(pinPolicy);
    nbnxn_atomdata_init(gmx::MDLogger(), atomData.get(), kernelSetup.kernelType, 0 /* geometric rule */, system.numAtomTypes,
                        system.nonbondedParameters, 1, useGpu ? 1 : numThreads); /* one output buffer with a GPU: nbnxm_setup.cpp:431-433 */
    NbnxmGpu* gpuNbv = nullptr;
    if (useGpu)
    {
        static DeviceInformation        deviceInfo{};
        static gmx::DeviceStreamManager deviceStreamManager(deviceInfo, false, gmx::SimulationWorkload(), false);
        gpuNbv = Nbnxm::gpu_init(deviceStreamManager, &ic, pairlistParams, atomData.get(), false);
    }
    auto nbv = std::make_unique<nonbonded_verlet_t>(std::move(pairlistSets), std::move(pairSearch), std::move(atomData), kernelSetup, gpuNbv, nullptr);

    t_nrnb     nrnb;
    const rvec lowerCorner = { 0, 0, 0 };
    const rvec upperCorner = { system.box[XX][XX], system.box[YY][YY], system.box[ZZ][ZZ] };
    const real atomDensity = system.coordinates.size() / det(system.box);
    nbnxn_put_on_grid(nbv.get(), system.box, 0, lowerCorner, upperCorner, nullptr, { 0, int(system.coordinates.size()) }, atomDensity,
                      system.atomInfoAllVdw, system.coordinates, 0, nullptr);
    nbv->setAtomProperties(system.atomTypes, system.charges, system.atomInfoAllVdw);
    if (useGpu)
    {
        /* the order of do_force() on a search step (mdlib/sim_util.cpp:1190,1327-1366) */
        nbnxn_atomdata_copy_shiftvec(false, system.forceRec.shift_vec, nbv->nbat.get());
        Nbnxm::gpu_init_atomdata(nbv->gpu_nbv, nbv->nbat.get());
    }
    nbv->constructPairlist(gmx::InteractionLocality::Local, system.excls, 0, &nrnb);
    return nbv;
}

/* one step; returns nothing, forces in f */
void step(nonbonded_verlet_t* nbv, const gmx::BenchmarkSystem& system, const interaction_const_t& ic, const gmx::StepWorkload& stepWork,
          gmx_enerdata_t* enerd, t_nrnb* nrnb, std::vector<gmx::RVec>* f, std::vector<gmx::RVec>* shiftForces)
{
    nbv->convertCoordinates(gmx::AtomLocality::Local, false, system.coordinates);
    std::fill(f->begin(), f->end(), gmx::RVec{ 0, 0, 0 });
    if (nbv->useGpu())
    {
        Nbnxm::gpu_clear_outputs(nbv->gpu_nbv, stepWork.computeVirial);
        Nbnxm::gpu_copy_xq_to_gpu(nbv->gpu_nbv, nbv->nbat.get(), gmx::AtomLocality::Local);
    }
    nbv->dispatchNonbondedKernel(gmx::InteractionLocality::Local, ic, stepWork, enbvClearFYes, system.forceRec, enerd, nrnb);
    if (nbv->useGpu())
    {
        Nbnxm::gpu_launch_cpyback(nbv->gpu_nbv, nbv->nbat.get(), stepWork, gmx::AtomLocality::Local);
        Nbnxm::gpu_wait_finish_task(nbv->gpu_nbv, stepWork, gmx::AtomLocality::Local, enerd->grpp.ener[egLJSR].data(),
                                    enerd->grpp.ener[egCOULSR].data(), *shiftForces, nullptr);
    }
    nbv->atomdata_add_nbat_f_to_f(gmx::AtomLocality::All, *f);
}

} // namespace

int main(int argc, char** argv)
{
    const int  sizeFactor = argc > 1 ? atoi(argv[1]) : 32;
    const bool pme        = !(argc > 2 && std::strcmp(argv[2], "rf") == 0);
    const int  iterations = argc > 3 ? atoi(argv[3]) : 50;
    const int  numThreads = argc > 4 ? atoi(argv[4]) : omp_get_max_threads();
    gmx_omp_nthreads_set(emntPairsearch, numThreads);
    gmx_omp_nthreads_set(emntNonbonded, numThreads);
    gmx_omp_nthreads_set(emntDefault, numThreads);
    const bool cpuOnly = std::getenv("NBNXM_BENCH_CPU_ONLY") != nullptr; /* the build check on a machine without a GPU */

    gmx::BenchmarkSystem system(sizeFactor);
    const size_t         n = system.coordinates.size();
    interaction_const_t  ic;
    setupInteractionConst(&ic, pme);
    gmx::StepWorkload stepWork;
    stepWork.computeForces          = true;
    stepWork.computeNonbondedForces = true;
    t_nrnb                 nrnb = { 0 };
    gmx_enerdata_t         enerd(1, 0);
    std::vector<gmx::RVec> shiftForces(SHIFTS, gmx::RVec{ 0, 0, 0 });

    /* analytical count of useful pairs, bench_setup.cpp:273-276 */
    const double density = n / det(system.box);
    const double useful  = n * 0.5 * (density * 4.0 / 3.0 * M_PI * std::pow(double(c_cutoff), 3) + 1);

    double tSearch[2] = { 0, 0 }, tStep[2] = { 0, 0 };
    std::vector<gmx::RVec> f[2] = { std::vector<gmx::RVec>(n), std::vector<gmx::RVec>(n) };
    for (int gpu = 0; gpu < (cpuOnly ? 1 : 2); gpu++)
    {
        double t0  = seconds();
        auto   nbv = setupNbnxm(system, ic, gpu != 0, numThreads);
        tSearch[gpu] = seconds() - t0;
        for (int i = 0; i < 3; i++) step(nbv.get(), system, ic, stepWork, &enerd, &nrnb, &f[gpu], &shiftForces);
        t0 = seconds();
        for (int i = 0; i < iterations; i++) step(nbv.get(), system, ic, stepWork, &enerd, &nrnb, &f[gpu], &shiftForces);
        tStep[gpu] = (seconds() - t0) / iterations;
    }
    double num = 0, den = 0;
    bool   finite = true;
    const int g = cpuOnly ? 0 : 1;
    for (size_t a = 0; a < n; a++)
    {
        for (int d = 0; d < DIM; d++)
        {
            finite = finite && std::isfinite(f[g][a][d]);
            num += double(f[g][a][d] - f[0][a][d]) * double(f[g][a][d] - f[0][a][d]);
            den += double(f[0][a][d]) * double(f[0][a][d]);
        }
    }
    const double rel = std::sqrt(num / den);
    std::printf("{\"system\": \"BenchmarkSystem(%d)\", \"atoms\": %zu, \"interaction\": \"LJ + %s\", \"rc\": %.2f, \"threads\": %d, "
                "\"useful_pairs_analytical\": %.0f, \"cpu_kernel\": \"reference SIMD (%s)\", \"cpu_setup_s\": %.4f, \"cpu_ms_per_step\": %.4f, "
                "\"cpu_pairs_per_s\": %.4g, \"gpu_setup_s\": %.4f, \"gpu_ms_per_step\": %.4f, \"gpu_pairs_per_s\": %.4g, "
                "\"force_rel_rms_gpu_vs_cpu\": %.3g, \"step\": \"convertCoordinates + [xq H2D] + dispatchNonbondedKernel + [f D2H, wait] + "
                "atomdata_add_nbat_f_to_f\"}\n",
                sizeFactor, n, pme ? "Ewald real space (analytical)" : "reaction field", double(c_cutoff), numThreads, useful,
#ifdef GMX_NBNXN_SIMD_2XNN
                "2xMM",
#else
                "4xM",
#endif
                tSearch[0], tStep[0] * 1e3, useful / tStep[0], tSearch[g], tStep[g] * 1e3, useful / tStep[g], rel);
    return (finite && rel < 1e-5) ? 0 : 1;
}
