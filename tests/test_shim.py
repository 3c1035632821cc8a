"""The compiled reference-side binding (shim/): the reference's own nbnxm module and nblib ForceCalculator, built from the
reference tree with the GPU sub-interface enabled, driving libb200nb.so through shim/nbnxm_b200.cpp.

CPU part: the library exports every Nbnxm::gpu_* symbol nbnxm_gpu.h / gpu_data_mgmt.h declare, nothing it needs is unresolved,
and the test program's CPU leg reproduces the reference's golden forces.  GPU part: the reference's nblib force tests
(api/nblib/tests/nbkernelsystem.cpp:69-84,187-202) with NBKernelOptions::useGpu = true at the reference's tolerance."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "shim", "_build", "libgmx_nbnxm_b200.so")
EXE = os.path.join(ROOT, "shim", "_build", "nblib_gpu_test")
BENCH = os.path.join(ROOT, "shim", "_build", "nbnxm_bench_gpu")
BONDED = os.path.join(ROOT, "shim", "_build", "gpubonded_test")
needs_shim = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(EXE)),
                                reason="shim/_build not built (needs the reference tree: shim/build_shim.sh)")

# nbnxm_gpu.h:138-355 and gpu_data_mgmt.h:72-138
GPU_INTERFACE = ["gpu_init", "gpu_init_pairlist", "gpu_init_atomdata", "gpu_pme_loadbal_update_param", "gpu_upload_shiftvec",
                 "gpu_clear_outputs", "gpu_free", "gpu_get_timings", "gpu_reset_timings", "gpu_min_ci_balanced",
                 "gpu_is_kernel_ewald_analytical", "gpu_get_command_stream", "gpu_get_xq", "gpu_get_f", "gpu_get_fshift",
                 "gpu_copy_xq_to_gpu", "gpu_launch_kernel", "gpu_launch_kernel_pruneonly", "gpu_launch_cpyback", "gpu_try_finish_task",
                 "gpu_wait_finish_task", "nbnxn_gpu_init_x_to_nbat_x", "nbnxn_gpu_x_to_nbat_x", "nbnxnInsertNonlocalGpuDependency",
                 "setupGpuShortRangeWork", "haveGpuShortRangeWork", "nbnxn_wait_x_on_device", "getGpuForces"]


def ulps(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)

    def key(v):  # monotonic integer image of a float (sign-magnitude -> two's complement)
        i = v.view(np.int32).astype(np.int64)
        return np.where(i < 0, -(i & 0x7FFFFFFF), i)
    return np.abs(key(a) - key(b))


def golden(name):
    return np.array(json.load(open(os.path.join(ROOT, "tests", "golden", name)))["forces"], np.float32)


def run(cpu_only):
    env = dict(os.environ)
    if cpu_only:
        env["NBLIB_GPU_TEST_CPU_ONLY"] = "1"
    r = subprocess.run([EXE], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    return json.loads(r.stdout)


@needs_shim
def test_shim_exports_the_gpu_interface():
    out = subprocess.run(["nm", "-D", "-C", "--defined-only", LIB], capture_output=True, text=True).stdout
    for fn in GPU_INTERFACE:
        assert "Nbnxm::%s(" % fn in out, fn
    # the reference's nbnxm module in the same library calls them (undefined Nbnxm::gpu_* would mean the CPU-build stubs were
    # compiled in instead of real calls): the library has no unresolved symbols at all (-z defs at link time) and needs libb200nb
    need = subprocess.run(["readelf", "-d", LIB], capture_output=True, text=True).stdout
    assert "libb200nb.so" in need
    # and nothing CUDA in the shim's own object: it is host code over the C ABI
    shim_obj = os.path.join(ROOT, "shim", "_build", "obj", "shim_nbnxm_b200.o")
    und = subprocess.run(["nm", "-u", shim_obj], capture_output=True, text=True).stdout
    assert "b200nb_upload_pairlist" in und and "cuda" not in und.lower()


@needs_shim
def test_shim_build_cpu_leg_reproduces_reference_goldens():
    d = run(cpu_only=True)
    # the reference's own CPU kernel in this build: 184 / 229 ULP from its XML data (SURVEY 8c measured the same 229 on its own build)
    assert ulps(d["argon"]["cpu"], golden("argon12_forces.json")).max() <= 256
    assert ulps(d["spc_methanol"]["cpu"], golden("spc_methanol_forces.json")).max() <= 256


@pytest.mark.gpu
@needs_shim
def test_reference_nblib_force_tests_with_use_gpu():
    """ArgonForcesAreCorrect / SpcMethanolForcesAreCorrect through nblib::ForceCalculator with useGpu = true.  Tolerance: the
    reference's 200 ULP (testhelpers.h:73-77) against its XML data; where the reference's own CPU kernel misses that on this
    build (one SPC-methanol component, 229 ULP), twice its deviation."""
    d = run(cpu_only=False)
    for key, gold in (("argon", "argon12_forces.json"), ("spc_methanol", "spc_methanol_forces.json")):
        ref = golden(gold)
        gpu, cpu = np.array(d[key]["gpu"], np.float32), np.array(d[key]["cpu"], np.float32)
        assert gpu.shape == ref.shape
        bar = np.maximum(200, 2 * ulps(cpu, ref))
        assert np.all(ulps(gpu, ref) <= bar), (key, ulps(gpu, ref).max())
    for key in ("spc_methanol_rf", "spc_methanol_pme"):
        gpu, cpu = np.array(d[key]["gpu"], np.float64), np.array(d[key]["cpu"], np.float64)
        assert np.sqrt(((gpu - cpu) ** 2).sum() / (cpu ** 2).sum()) < 1e-5, key


def run_bench(args, cpu_only):
    env = dict(os.environ, OMP_PROC_BIND="spread", OMP_PLACES="cores")
    if cpu_only:
        env["NBNXM_BENCH_CPU_ONLY"] = "1"
    r = subprocess.run([BENCH] + [str(a) for a in args], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-800:], r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


@needs_shim
def test_benchmark_driver_cpu_leg():
    d = run_bench([1, "rf", 2, 2], cpu_only=True)
    assert d["atoms"] == 3000 and d["cpu_ms_per_step"] > 0


@pytest.mark.gpu
@needs_shim
@pytest.mark.parametrize("size,eel", [(32, "pme"), (8, "rf")])
def test_reference_benchmark_protocol_with_gpu_backend(size, eel):
    """BASELINE configs[2] (BenchmarkSystem(32), 96 000 atoms) and configs[1] through C++: the reference's nonbonded-benchmark
    set-up and step (nbnxm/benchmark/bench_setup.cpp:170-343) with KernelType::Gpu8x8x8 -- unmodified nbnxm module, reference-built
    grid and pair list, Nbnxm::gpu_* shim, B200 kernels -- against the reference's CPU SIMD kernel in the same process."""
    d = run_bench([size, eel, 20], cpu_only=False)
    assert d["atoms"] == 3000 * size
    assert d["force_rel_rms_gpu_vs_cpu"] < 1e-5
    assert d["gpu_ms_per_step"] < d["cpu_ms_per_step"]


def run_bonded(cpu_only):
    env = dict(os.environ)
    if cpu_only:
        env["GPUBONDED_TEST_CPU_ONLY"] = "1"
    r = subprocess.run([BONDED], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-800:], r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


@needs_shim
def test_gpubonded_class_is_defined_and_cpu_leg_runs():
    """gmx::GpuBonded (listed_forces/gpubonded.h:99-172) is defined by the shim library on the C ABI, and the test program's CPU
    leg (the reference's calculateSimpleBond / do_pairs on the listed interactions laid over the benchmark water) runs."""
    out = subprocess.run(["nm", "-D", "-C", "--defined-only", LIB], capture_output=True, text=True).stdout
    for fn in ("GpuBonded(gmx_ffparams_t const&", "updateInteractionListsAndDeviceBuffers(", "setPbcAndlaunchKernel(", "launchKernel(",
               "launchEnergyTransfer(", "waitAccumulateEnergyTerms(", "clearEnergies(", "haveInteractions("):
        assert "gmx::GpuBonded::" + fn in out, fn
    und = subprocess.run(["nm", "-u", os.path.join(ROOT, "shim", "_build", "obj", "shim_gpubonded_b200.o")], capture_output=True, text=True).stdout
    assert "b200nb_bonded_set_list" in und and "b200nb_bonded_launch" in und and "cuda" not in und.lower()
    d = run_bonded(cpu_only=True)
    assert d["atoms"] == 3000 and d["bonds"] == 2000 and d["cpu_force_sumsq"] > 0 and d["cpu_e_bonds"] > 1e4


@pytest.mark.gpu
@needs_shim
def test_reference_gpubonded_interface_on_b200nb():
    """gmx::GpuBonded driven the way do_force() drives it -- updateInteractionListsAndDeviceBuffers with Nbnxm::gpu_get_xq,
    setPbcAndlaunchKernel between gpu_copy_xq_to_gpu and the nonbonded kernel, launchEnergyTransfer, waitAccumulateEnergyTerms --
    through the unmodified nbnxm module and the shim, against the reference's CPU functions in the same process: all eight
    interaction types, forces (energy / virial and force-only flavours), energies per type, shift forces."""
    d = run_bonded(cpu_only=False)
    assert d["force_rel_rms_gpu_vs_cpu"] < 1e-5 and d["force_only_rel_rms_gpu_vs_cpu"] < 1e-5
    assert d["energy_max_rel_err"] < 2e-5 and d["fshift_max_abs_diff"] <= 2e-5 * d["fshift_max"]
    assert d["lj14"] != 0 and d["coul14"] != 0
