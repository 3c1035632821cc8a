#!/bin/bash
# Builds shim/_build/libgmx_nbnxm_b200.so: the reference's nbnxm module with its GPU sub-interface ENABLED (GMX_GPU_CUDA=1,
# so the Nbnxm::gpu_* calls in nbnxm.cpp / pairlist.cpp / kerneldispatch.cpp / prunekerneldispatch.cpp are real external
# calls, not the empty stubs of a CPU build) + nblib's ForceCalculator with the GPU hook of shim/nblib_gmxsetup_gpu.patch + shim/nblib_gmxcalculator_gpu.patch + our
# implementation of that sub-interface (shim/nbnxm_b200.cpp) and of gmx::GpuBonded (shim/gpubonded_b200.cpp) on libb200nb.so;
# shim/_build/nblib_gpu_test, the reference's
# nblib force tests (api/nblib/tests/nbkernelsystem.cpp:69-202) run with NBKernelOptions::useGpu = true; and
# shim/_build/nbnxm_bench_gpu, the reference's nonbonded-benchmark protocol (nbnxm/benchmark/bench_setup.cpp) with the GPU backend.
# Reference sources are compiled where they lie under /root/reference (never copied into the repo; the two nblib files the
# patch touches are patched in shim/_build/src/, which is git-ignored).  No cmake: g++ on the files directly, as
# oracle/build_ref.sh does for the CPU build, whose objects (GPU-independent files) are reused.
# Usage: shim/build_shim.sh [reference-root]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${1:-/root/reference}"
R="$REF/src"
N="$REF/api/nblib"
OUT="$HERE/_build"
OBJ="$OUT/obj"
[ -d "$R/gromacs/nbnxm" ] || { echo "reference sources not found at $REF"; exit 3; }
[ -f "$ROOT/oracle/_ref/obj/nbnxm_grid.o" ] || bash "$ROOT/oracle/build_ref.sh" "$REF"
[ -f "$ROOT/gmxapi_b200/libb200nb.so" ] || make -C "$ROOT/gmxapi_b200/csrc"
mkdir -p "$OBJ" "$OUT/cfg" "$OUT/src"
CXX="${CXX_REF:-/usr/bin/g++}"
CUDA_INC="${CUDA_HOME:-/usr/local/cuda}/include"
# the hand-written feature header of the CPU build with the GPU switches turned on
sed -e 's/^#define GMX_GPU 0/#define GMX_GPU 1/' -e 's/^#define GMX_GPU_CUDA 0/#define GMX_GPU_CUDA 1/' "$ROOT/oracle/refcfg/config.h" > "$OUT/cfg/config.h.new"
cmp -s "$OUT/cfg/config.h.new" "$OUT/cfg/config.h" || mv "$OUT/cfg/config.h.new" "$OUT/cfg/config.h"
cp -u "$ROOT/oracle/refcfg/gmxpre-config.h" "$OUT/cfg/"
# nblib with the GPU hook: patched copies of the two files the hook touches (the diffs are listed in INTEGRATION.md)
for f in gmxsetup gmxcalculator; do
  cp "$N/$f.cpp" "$OUT/src/$f.cpp.tmp"
  patch -s "$OUT/src/$f.cpp.tmp" "$HERE/nblib_${f}_gpu.patch"
  cmp -s "$OUT/src/$f.cpp.tmp" "$OUT/src/$f.cpp" || mv "$OUT/src/$f.cpp.tmp" "$OUT/src/$f.cpp"
  rm -f "$OUT/src/$f.cpp.tmp" "$OUT/src/$f.cpp.tmp.orig"
done
FLAGS="-std=c++17 -O2 -mavx512f -mfma -fopenmp -fexcess-precision=fast -DHAVE_CONFIG_H -DGMX_DOUBLE=0 -fPIC -w \
 -I$OUT/cfg -I$R -I$R/external -I$R/external/thread_mpi/include -I$REF/api -I$REF/api/nblib -I$CUDA_INC -I$ROOT/include"
# 1. reference files whose code depends on the GPU switches (everything else comes from the CPU build's objects)
SRCS=""
for f in nbnxm/atomdata.cpp nbnxm/kerneldispatch.cpp nbnxm/nbnxm.cpp nbnxm/nbnxm_setup.cpp nbnxm/pairlist.cpp nbnxm/prunekerneldispatch.cpp \
         gpu_utils/hostallocator.cpp gpu_utils/device_stream_manager.cpp gpu_utils/device_context.cpp topology/exclusionblocks.cpp \
         mdlib/rf_util.cpp mdtypes/md_enums.cpp; do
  SRCS="$SRCS $R/gromacs/$f"
done
# 2. nblib (api/nblib): unmodified files + the two patched copies
for f in box.cpp forcecalculator.cpp integrator.cpp interactions.cpp molecules.cpp particletype.cpp simulationstate.cpp topology.cpp \
         topologyhelpers.cpp; do
  SRCS="$SRCS $N/$f"
done
SRCS="$SRCS $OUT/src/gmxsetup.cpp $OUT/src/gmxcalculator.cpp"
# 3. ours
SRCS="$SRCS $HERE/nbnxm_b200.cpp $HERE/gpubonded_b200.cpp $HERE/gmx_link_stubs.cpp"
compile_one() {
  src="$1"; obj="$OBJ/$(echo "$src" | sed "s#$R/gromacs/##; s#$N/#nblib_#; s#$OUT/src/#nblib_#; s#$HERE/#shim_#; s#/#_#g; s#\.cpp\$#.o#")"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$OUT/cfg/config.h" -nt "$obj" ] || [ "$ROOT/include/b200nb.h" -nt "$obj" ]; then
    $CXX $FLAGS -c "$src" -o "$obj" || exit 255
  fi
}
export -f compile_one; export R N HERE OBJ OUT CXX FLAGS ROOT
echo $SRCS | tr ' ' '\n' | xargs -P "$(nproc)" -I{} bash -c 'compile_one {}'
# GPU-independent objects of the CPU build (kernels, grid, search helpers, utilities, stand-ins for unbuilt subsystems)
CPUOBJ=$(ls "$ROOT"/oracle/_ref/obj/*.o | grep -v -e ref_harness.o -e nbnxm_atomdata.o -e nbnxm_kerneldispatch.o -e nbnxm_nbnxm.o \
         -e nbnxm_nbnxm_setup.o -e nbnxm_pairlist.o -e nbnxm_prunekerneldispatch.o -e gpu_utils_hostallocator.o)
$CXX -shared -fopenmp -o "$OUT/libgmx_nbnxm_b200.so" $OBJ/*.o $CPUOBJ -Wl,-z,defs -Wl,--allow-multiple-definition -L"$ROOT/gmxapi_b200" -lb200nb \
  -Wl,-rpath,'$ORIGIN/../../gmxapi_b200' -L"${CUDA_HOME:-/usr/local/cuda}/lib64" -lcudart -lm
$CXX $FLAGS -o "$OUT/nblib_gpu_test" "$HERE/nblib_gpu_test.cpp" "$N/tests/testsystems.cpp" -L"$OUT" -lgmx_nbnxm_b200 -Wl,-rpath,'$ORIGIN' -L"$ROOT/gmxapi_b200" -lb200nb \
  -Wl,-rpath,'$ORIGIN/../../gmxapi_b200' -fopenmp
# the reference's benchmark protocol with the GPU backend, on the reference's own BenchmarkSystem (bench_system.cpp from the tree)
$CXX $FLAGS -o "$OUT/nbnxm_bench_gpu" "$HERE/nbnxm_bench_gpu.cpp" "$R/gromacs/nbnxm/benchmark/bench_system.cpp" -L"$OUT" -lgmx_nbnxm_b200 \
  -Wl,-rpath,'$ORIGIN' -L"$ROOT/gmxapi_b200" -lb200nb -Wl,-rpath,'$ORIGIN/../../gmxapi_b200' -fopenmp
# gmx::GpuBonded on libb200nb (gpubonded_b200.cpp, in the library) driven as do_force() drives it, against the reference's CPU functions
$CXX $FLAGS -o "$OUT/gpubonded_test" "$HERE/gpubonded_test.cpp" "$R/gromacs/nbnxm/benchmark/bench_system.cpp" -L"$OUT" -lgmx_nbnxm_b200 \
  -Wl,-rpath,'$ORIGIN' -L"$ROOT/gmxapi_b200" -lb200nb -Wl,-rpath,'$ORIGIN/../../gmxapi_b200' -fopenmp
echo "built $OUT/libgmx_nbnxm_b200.so, $OUT/nblib_gpu_test, $OUT/nbnxm_bench_gpu and $OUT/gpubonded_test"
