/* The reference's own nblib force tests -- NBlibTest.ArgonForcesAreCorrect and NBlibTest.SpcMethanolForcesAreCorrect,
 * api/nblib/tests/nbkernelsystem.cpp:69-84,187-202 -- run through the UNMODIFIED public nblib API
 * (SimulationState -> ForceCalculator -> compute) with NBKernelOptions::useGpu = true, i.e. through
 * GmxForceCalculator -> nonbonded_verlet_t -> Nbnxm::gpu_* (shim/nbnxm_b200.cpp) -> libb200nb.so, and once more with the CPU
 * kernel the reference test uses (SimdNo).  The test systems are the reference's (api/nblib/tests/testsystems.cpp, compiled
 * from the reference tree).  Prints one JSON object {"argon": {"gpu": [[fx,fy,fz],...], "cpu": [...]}, "spc_methanol": {...}};
 * tests/test_gpu_shim.py compares it with the reference's XML golden data (tests/golden/*.json) at the reference's tolerance
 * (200 ULP, api/nblib/tests/testhelpers.h:73-77).  Exit code 0 iff every GPU force is finite and within 1e-4 relative of
 * the CPU path (a coarse self-check; the real assertion is the Python test's). */
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <string>
#include <vector>

#include "nblib/forcecalculator.h"
#include "nblib/kerneloptions.h"
#include "nblib/simulationstate.h"
#include "nblib/tests/testsystems.h"

using namespace nblib;

static std::vector<Vec3> forcesOf(SimulationState simState, bool useGpu, CoulombType coulomb)
{
    NBKernelOptions options;
    options.useGpu      = useGpu;
    options.nbnxmSimd   = SimdKernels::SimdNo;
    options.coulombType = coulomb;
    ForceCalculator     forceCalculator(simState, options);
    gmx::ArrayRef<Vec3> forces(simState.forces());
    forceCalculator.compute(simState.coordinates(), forces);
    if (useGpu)
    {
        /* a second evaluation: outputs are cleared and coordinates re-sent every call */
        forceCalculator.compute(simState.coordinates(), forces);
    }
    return std::vector<Vec3>(forces.begin(), forces.end());
}

static void printForces(const char* key, const std::vector<Vec3>& f, bool last)
{
    std::printf("\"%s\": [", key);
    for (size_t i = 0; i < f.size(); i++)
    {
        std::printf("%s[%.9g, %.9g, %.9g]", i ? ", " : "", f[i][0], f[i][1], f[i][2]);
    }
    std::printf("]%s", last ? "" : ", ");
}

static int runSystem(const char* name, SimulationState simState, CoulombType coulomb, bool last)
{
    const std::vector<Vec3> cpu = forcesOf(simState, false, coulomb);
    /* NBLIB_GPU_TEST_CPU_ONLY=1: the CPU leg alone (the build check on a machine without a GPU) */
    const std::vector<Vec3> gpu = std::getenv("NBLIB_GPU_TEST_CPU_ONLY") ? cpu : forcesOf(simState, true, coulomb);
    std::printf("\"%s\": {", name);
    printForces("gpu", gpu, false);
    printForces("cpu", cpu, true);
    std::printf("}%s", last ? "" : ", ");
    double num = 0, den = 0;
    int    bad = gpu.size() != cpu.size();
    for (size_t i = 0; i < gpu.size() && !bad; i++)
    {
        for (int d = 0; d < 3; d++)
        {
            if (!std::isfinite(gpu[i][d]))
            {
                bad = 1;
            }
            num += (gpu[i][d] - cpu[i][d]) * (gpu[i][d] - cpu[i][d]);
            den += cpu[i][d] * cpu[i][d];
        }
    }
    return bad || !(std::sqrt(num / den) < 1e-4);
}

int main()
{
    int bad = 0;
    std::printf("{");
    {
        ArgonSimulationStateBuilder builder;
        bad |= runSystem("argon", builder.setupSimulationState(), CoulombType::Cutoff, false);
    }
    {
        SpcMethanolSimulationStateBuilder builder;
        bad |= runSystem("spc_methanol", builder.setupSimulationState(), CoulombType::Cutoff, false);
    }
    {
        /* beyond the reference's test: the same molecules with reaction-field and Ewald real-space electrostatics */
        SpcMethanolSimulationStateBuilder builder;
        bad |= runSystem("spc_methanol_rf", builder.setupSimulationState(), CoulombType::ReactionField, false);
    }
    {
        SpcMethanolSimulationStateBuilder builder;
        bad |= runSystem("spc_methanol_pme", builder.setupSimulationState(), CoulombType::Pme, true);
    }
    std::printf("}\n");
    return bad;
}
