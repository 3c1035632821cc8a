#!/bin/bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the reference's nbnxm CPU path (grid, pair search, atomdata, plain-C / SIMD / GPU-emulation
# kernels) directly with g++ from the sources where they lie under /root/reference -- no cmake, no
# reference build system, no copies of reference sources -- plus our C-ABI harness, into
# oracle/_ref/libgmxref_nbnxm.so. The feature macros cmake would generate are stated by hand in
# oracle/refcfg/config.h.  Usage: oracle/build_ref.sh [reference-root]   (default /root/reference)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
R="$REF/src"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
[ -d "$R/gromacs/nbnxm" ] || { echo "reference sources not found at $REF"; exit 3; }
mkdir -p "$OBJ"
CXX="${CXX_REF:-/usr/bin/g++}"
# flags the reference's cmake would pick for gcc + AVX-512 Release (cmake/gmxCFlags.cmake:253-254,390;
# cmake/gmxSimdFlags.cmake:269)
FLAGS="-std=c++17 -O3 -mavx512f -mfma -fopenmp -fexcess-precision=fast -funroll-all-loops \
 -DHAVE_CONFIG_H -DGMX_DOUBLE=0 -fPIC -ffunction-sections -fdata-sections -w \
 -I$HERE/refcfg -I$R -I$R/external -I$R/external/thread_mpi/include"
SRCS=$(ls $R/gromacs/nbnxm/*.cpp $R/gromacs/nbnxm/kernels_reference/*.cpp \
          $R/gromacs/nbnxm/kernels_simd_2xmm/*.cpp $R/gromacs/nbnxm/kernels_simd_4xm/*.cpp | grep -v nbnxm_gpu_data_mgmt)
for f in pbcutil/pbc.cpp mdlib/enerdata_utils.cpp utility/alignedallocator.cpp gpu_utils/hostallocator.cpp \
         math/functions.cpp utility/smalloc.cpp utility/stringutil.cpp tables/forcetable.cpp \
         ewald/ewald_utils.cpp math/utilities.cpp utility/logger.cpp \
         gmxlib/nonbonded/nb_free_energy.cpp mdtypes/interaction_const.cpp \
         listed_forces/bonded.cpp listed_forces/pairs.cpp listed_forces/restcbt.cpp pbcutil/pbc_simd.cpp topology/ifunc.cpp; do
  SRCS="$SRCS $R/gromacs/$f"
done
SRCS="$SRCS $HERE/ref_harness.cpp $HERE/ref_stubs.cpp"
compile_one() {
  src="$1"; obj="$OBJ/$(echo "$src" | sed "s#$R/gromacs/##; s#$HERE/##; s#/#_#g; s#\.cpp\$#.o#")"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/refcfg/config.h" -nt "$obj" ] || [ "$HERE/ref_harness.h" -nt "$obj" ]; then
    $CXX $FLAGS -c "$src" -o "$obj" || exit 1
  fi
}
export -f compile_one; export R HERE OBJ CXX FLAGS
echo $SRCS | tr ' ' '\n' | xargs -P "$(nproc)" -I{} bash -c 'compile_one {}'
$CXX -shared -fopenmp -o "$OUT/libgmxref_nbnxm.so" $OBJ/*.o -Wl,--gc-sections -Wl,-z,defs -lm
echo "built $OUT/libgmxref_nbnxm.so"
