/* TEST INFRASTRUCTURE ONLY.
 * Minimal stand-ins for the handful of GROMACS utility symbols the reference nbnxm CPU path
 * references but which live in subsystems we do not compile (fatal-error reporting, wallcycle
 * counters, thread-count registry, DD / Verlet-buffer tuning entry points that the harness never
 * reaches). Bodies are our own; they either do the obvious thing or abort loudly if reached.
 */
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>

#include "gromacs/math/vectypes.h"
#include "gromacs/utility/arrayref.h"
#include "gromacs/utility/basedefinitions.h"
#include "gromacs/utility/real.h"

struct gmx_wallcycle;
struct gmx_mtop_t;
struct t_inputrec;
struct t_commrec;
struct gmx_domdec_t;
struct gmx_domdec_zones_t;
struct t_nblist;
struct t_forcerec;
struct t_mdatoms;
struct nb_kernel_data_t;
struct t_nrnb;
struct t_graph
{
    int dummy;
};
struct gmx_moltype_t;
struct VerletbufListSetup
{
    int cluster_size_i;
    int cluster_size_j;
};
enum class ListSetupType;
enum class PbcType : int;
namespace gmx
{
class ForceWithShiftForces;
}

[[noreturn]] static void unreachable(const char* what)
{
    std::fprintf(stderr, "gmxref stub reached: %s\n", what);
    std::abort();
}

FILE*    debug        = nullptr;
gmx_bool gmx_debug_at = FALSE;
const char* efpt_names[16] = { "fep", "mass", "coul", "vdw", "bonded", "restraint", "temperature", nullptr };

void gmx_fatal(int /*fatal_errno*/, const char* file, int line, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    std::fprintf(stderr, "gmx_fatal at %s:%d: ", file, line);
    std::vfprintf(stderr, fmt, ap);
    std::fprintf(stderr, "\n");
    va_end(ap);
    std::abort();
}

void gmx_warning(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    std::vfprintf(stderr, fmt, ap);
    std::fprintf(stderr, "\n");
    va_end(ap);
}

void _range_check(int n, int n_min, int n_max, const char* warn_str, const char* var, const char* file, int line)
{
    if (n < n_min || n >= n_max)
    {
        std::fprintf(stderr, "range check failed %s:%d %s=%d [%d,%d) %s\n", file, line, var, n, n_min, n_max,
                     warn_str ? warn_str : "");
        std::abort();
    }
}

namespace gmx
{
namespace internal
{
void assertHandler(const char* condition, const char* msg, const char* func, const char* file, int line)
{
    std::fprintf(stderr, "GMX assertion failed: %s (%s) in %s at %s:%d\n", condition, msg, func, file, line);
    std::abort();
}
} // namespace internal

std::string findLibraryFile(const char* /*filename*/, bool /*bAddCWD*/, bool /*bFatal*/)
{
    unreachable("findLibraryFile");
}
} // namespace gmx

/* thread-count registry (mdlib/gmx_omp_nthreads.h) */
static int g_nthreads[32] = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1 };
int  gmx_omp_nthreads_get(int mod) { return g_nthreads[mod] > 0 ? g_nthreads[mod] : 1; }
void gmx_omp_nthreads_set(int mod, int nthreads) { g_nthreads[mod] = nthreads; }

/* wallcycle: the harness passes wcycle == nullptr everywhere */
void wallcycle_start(gmx_wallcycle*, int) {}
void wallcycle_start_nocount(gmx_wallcycle*, int) {}
double wallcycle_stop(gmx_wallcycle*, int) { return 0; }
void wallcycle_sub_start(gmx_wallcycle*, int) {}
void wallcycle_sub_start_nocount(gmx_wallcycle*, int) {}
void wallcycle_sub_stop(gmx_wallcycle*, int) {}

/* thread_mpi mutex used by smalloc's aligned allocation bookkeeping: single setup thread */
struct tMPI_Thread_mutex_t;
int tMPI_Thread_mutex_lock(tMPI_Thread_mutex_t*) { return 0; }
int tMPI_Thread_mutex_unlock(tMPI_Thread_mutex_t*) { return 0; }

void gmx_sumd(int, double*, const t_commrec*) {}

/* never reached from the harness */
real calcVerletBufferSize(const gmx_mtop_t&, real, const t_inputrec&, int, int, real, const VerletbufListSetup&)
{
    unreachable("calcVerletBufferSize");
}
VerletbufListSetup verletbufGetSafeListSetup(ListSetupType) { unreachable("verletbufGetSafeListSetup"); }
gmx_bool change_dd_cutoff(t_commrec*, const matrix, gmx::ArrayRef<const gmx::RVec>, real)
{
    unreachable("change_dd_cutoff");
}
gmx_domdec_zones_t* domdec_zones(gmx_domdec_t*) { unreachable("domdec_zones"); }
int  nonbondedMtsFactorStub();
namespace gmx
{
int nonbondedMtsFactor(const t_inputrec&) { unreachable("nonbondedMtsFactor"); }
} // namespace gmx
/* gmx_nb_free_energy_kernel: the real one is compiled in (gmxlib/nonbonded/nb_free_energy.cpp, oracle/build_ref.sh) */
bool haveFepPerturbedNBInteractions(const gmx_mtop_t&) { unreachable("haveFepPerturbedNBInteractions"); }
int  inputrec2nboundeddim(const t_inputrec*) { unreachable("inputrec2nboundeddim"); }
void shift_self(const t_graph&, const matrix, rvec*) { unreachable("shift_self"); }
void mk_mshift(FILE*, t_graph*, PbcType, const matrix, const rvec*) { unreachable("mk_mshift"); }
void pr_ivecs(FILE*, int, const char*, const ivec*, int, gmx_bool) {}
void pr_rvecs(FILE*, int, const char*, const rvec*, int) {}
t_graph mk_graph_moltype(const gmx_moltype_t&) { unreachable("mk_graph_moltype"); }

/* listed_forces/bonded.cpp references the restraint functions of subsystems we do not compile; never reached for the types the
 * harness drives (gmxref_bonded) */
union t_iparams;
struct t_pbc;
typedef real rvec4[4]; /* topology/ifunc.h:55 */
struct t_fcdata;
int  glatnr(const int* global_atom_index, int i) { return global_atom_index ? global_atom_index[i] + 1 : i + 1; }
real orires(int, const int*, const t_iparams*, const rvec*, rvec4*, rvec*, const t_pbc*, real, real*, const t_mdatoms*, t_fcdata*, int*)
{
    unreachable("orires");
}
real ta_disres(int, const int*, const t_iparams*, const rvec*, rvec4*, rvec*, const t_pbc*, real, real*, const t_mdatoms*, t_fcdata*, int*)
{
    unreachable("ta_disres");
}

#include <omp.h>
#include <exception>
#include <vector>
#include "gromacs/fileio/xvgr.h"
int gmx_omp_get_thread_num() { return omp_get_thread_num(); }
void _gmx_error(const char* key, const std::string& msg, const char* file, int line)
{
    std::fprintf(stderr, "gmx_error(%s) at %s:%d: %s\n", key, file, line, msg.c_str());
    std::abort();
}
namespace gmx
{
[[noreturn]] void processExceptionAsFatalError(const std::exception& ex)
{
    std::fprintf(stderr, "exception in reference code: %s\n", ex.what());
    std::abort();
}
} // namespace gmx
gmx::MultiDimArray<std::vector<double>, gmx::dynamicExtents2D> readXvgData(const std::string&)
{
    unreachable("readXvgData");
}
