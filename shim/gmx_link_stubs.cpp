/* TEST / BUILD SCAFFOLDING of shim/build_shim.sh, not product code.
 * nblib's GmxForceCalculator owns a t_forcerec and an interaction_const_t, whose out-of-line members live in mdlib/forcerec.cpp
 * -- a translation unit that pulls in most of mdrun.  The shim build links nblib + nbnxm only, so the handful of symbols those
 * two objects need are provided here with the obvious bodies; plus the pinned-memory and stream primitives the reference's
 * gpu_utils would compile with nvcc (pinning.cu, device_stream.cu), restated on the CUDA runtime API. */
#include "gmxpre.h"

#include <cuda_runtime_api.h>

#include "gromacs/ewald/ewald_utils.h"
#include "gromacs/gpu_utils/device_stream.h"
#include "gromacs/gpu_utils/pinning.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/tables/forcetable.h"
#include "gromacs/utility/fatalerror.h"
#include "gromacs/utility/smalloc.h"

/* t_forcerec holds unique_ptrs to a dozen mdrun subsystems that are only forward-declared in its header; nblib leaves all of
 * them null.  Empty stand-in definitions make the (empty) destructor instantiable here without compiling those subsystems. */
class DispersionCorrection
{
};
struct t_fcdata
{
};
class ListedForces
{
};
struct ForceHelperBuffers_placeholder;
namespace gmx
{
class WholeMoleculeTransform
{
};
class PmePpCommGpu
{
};
class GpuForceReduction
{
};
class GpuBonded
{
};
} // namespace gmx
#include "gromacs/mdtypes/forcerec.h"
#include "gromacs/nbnxm/nbnxm.h"

t_forcerec::t_forcerec() = default;
t_forcerec::~t_forcerec()
{
    sfree(shift_vec);
}

/* mdlib/forcerec.cpp:765-785 init_interaction_const_tables: Ewald correction tables for the tabulated CPU kernels */
void init_interaction_const_tables(FILE* /* fp */, interaction_const_t* ic, const real rlist)
{
    if (EEL_PME_EWALD(ic->eeltype))
    {
        const real tableScale = ewald_spline3_table_scale(*ic, true, false);
        const int  tableSize  = static_cast<int>((std::max(ic->rcoulomb, rlist) + 0.0) * tableScale) + 2;
        *ic->coulombEwaldTables = generateEwaldCorrectionTables(tableSize, tableScale, ic->ewaldcoeff_q, v_q_ewald_lr);
    }
}

namespace gmx
{
/* gpu_utils/pinning.cu */
void pinBuffer(void* pointer, std::size_t numBytes) noexcept
{
    if (numBytes)
    {
        cudaHostRegister(pointer, numBytes, cudaHostRegisterDefault);
    }
}
void unpinBuffer(void* pointer) noexcept
{
    cudaHostUnregister(pointer);
}
} // namespace gmx

/* gpu_utils/device_stream.cu */
DeviceStream::DeviceStream(const DeviceContext& /* deviceContext */, DeviceStreamPriority priority, const bool /* useTiming */)
{
    int         highest = 0;
    cudaError_t stat    = cudaSuccess;
    if (priority == DeviceStreamPriority::High && cudaDeviceGetStreamPriorityRange(nullptr, &highest) == cudaSuccess)
    {
        stat = cudaStreamCreateWithPriority(&stream_, cudaStreamDefault, highest);
    }
    else
    {
        stat = cudaStreamCreate(&stream_);
    }
    if (stat != cudaSuccess)
    {
        gmx_fatal(FARGS, "Could not create CUDA stream: %s", cudaGetErrorString(stat));
    }
}
DeviceStream::~DeviceStream()
{
    if (stream_ != nullptr)
    {
        cudaStreamDestroy(stream_);
        stream_ = nullptr;
    }
}
cudaStream_t DeviceStream::stream() const
{
    return stream_;
}
bool DeviceStream::isValid() const
{
    return stream_ != nullptr;
}
void DeviceStream::synchronize() const
{
    cudaStreamSynchronize(stream_);
}

/* api/nblib/util/user.cpp also holds the Maxwell-Boltzmann velocity generator, which drags in gmx::random and the exception
 * formatter; the force path needs only these two helpers of it */
#include <cmath>
#include <string>
#include <vector>

#include "nblib/basicdefinitions.h"
#include "nblib/vector.h"
namespace nblib
{
namespace detail
{
std::string next_token(std::string& s, const std::string& delimiter)
{
    const std::size_t pos   = s.find(delimiter);
    std::string       token = s.substr(0, pos);
    s.erase(0, pos == std::string::npos ? std::string::npos : pos + delimiter.length());
    return token;
}
} // namespace detail
bool checkNumericValues(const std::vector<Vec3>& values)
{
    for (const auto& v : values)
    {
        if (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2]))
        {
            return false;
        }
    }
    return true;
}
} // namespace nblib

/* topology/idef.cpp holds the InteractionDefinitions constructor next to the parameter printers (which drag in the text writer
 * and the pr_* helpers); the listed-forces test (shim/gpubonded_test.cpp) needs only the constructor: it binds the parameter
 * tables of the force field (idef.h:372-400) */
#include "gromacs/topology/forcefieldparameters.h"
#include "gromacs/topology/idef.h"
InteractionDefinitions::InteractionDefinitions(const gmx_ffparams_t& ffparams) :
    iparams(ffparams.iparams), functype(ffparams.functype), cmap_grid(ffparams.cmap_grid)
{
}
