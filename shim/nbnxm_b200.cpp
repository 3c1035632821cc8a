/* The reference's GPU sub-interface of the nbnxm module, implemented on libb200nb.so.
 *
 * Defines every `Nbnxm::gpu_*` symbol that src/gromacs/nbnxm/nbnxm_gpu.h:138-355 and gpu_data_mgmt.h:72-138 declare with
 * GPU_FUNC_QUALIFIER / CUDA_FUNC_QUALIFIER, i.e. exactly what the reference's own nbnxm/cuda/*.cu and
 * nbnxm_gpu_data_mgmt.cpp provide in a GMX_GPU=CUDA build.  Linked in their place, the UNMODIFIED nbnxm module
 * (nbnxm.cpp, pairlist.cpp, kerneldispatch.cpp, prunekerneldispatch.cpp, atomdata.cpp, compiled with GMX_GPU_CUDA=1) drives the
 * B200 kernels through its usual call sites:
 *   nonbonded_verlet_t::constructPairlist  -> gpu_init_pairlist          (pairlist.cpp:4247-4262)
 *   nonbonded_verlet_t::dispatchNonbondedKernel -> gpu_launch_kernel     (kerneldispatch.cpp:420-560, Gpu8x8x8 case)
 *   nonbonded_verlet_t::dispatchPruneKernelGpu  -> gpu_launch_kernel_pruneonly (prunekerneldispatch.cpp:99-113)
 *   do_force / nblib                        -> gpu_init_atomdata, gpu_upload_shiftvec, gpu_copy_xq_to_gpu, gpu_launch_cpyback,
 *                                              gpu_wait_finish_task, gpu_clear_outputs (mdlib/sim_util.cpp:1338-1933)
 * Gridding and pair search stay on the CPU where the reference does them for its CUDA backend; the shim hands their products
 * to the library (b200nb_set_grid_atoms / b200nb_upload_pairlist).  Host code only: no CUDA in this file.
 *
 * Built by shim/build_shim.sh against the headers under /root/reference/src (never copied). */
#include "gmxpre.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <vector>

#include "gromacs/gpu_utils/device_stream_manager.h"
#include "gromacs/gpu_utils/devicebuffer_datatype.h"
#include "gromacs/gpu_utils/gpu_utils.h"
#include "gromacs/math/units.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/locality.h"
#include "gromacs/mdtypes/md_enums.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/nbnxm/atomdata.h"
#include "gromacs/nbnxm/gpu_data_mgmt.h"
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/nbnxm_gpu.h"
#include "gromacs/nbnxm/pairlist.h"
#include "gromacs/nbnxm/pairlistparams.h"
#include "gromacs/nbnxm/pairlistsets.h"
#include "gromacs/pbcutil/ishift.h"
#include "gromacs/timing/gpu_timing.h"
#include "gromacs/utility/arrayref.h"
#include "gromacs/utility/fatalerror.h"

#include "b200nb.h"

static_assert(sizeof(nbnxn_sci_t) == sizeof(b200nb_sci_t) && sizeof(nbnxn_cj4_t) == sizeof(b200nb_cj4_t)
                      && sizeof(nbnxn_excl_t) == sizeof(b200nb_excl_t),
              "the C ABI mirrors NbnxnPairlistGpu's element layout (nbnxm/pairlist.h:174-225)");
static_assert(SHIFTS == B200NB_SHIFTS && CENTRAL == B200NB_CENTRAL, "pbcutil/ishift.h");

/* the reference only ever holds a pointer to this type (nbnxm.h:219: raw NbnxmGpu* gpu_nbv) */
struct NbnxmGpu
{
    b200nb_t* h = nullptr;
    bool      twoLocalities = false;
    int       numSlots = 0, numSlotsLocal = 0;
    /* rolling pruning: the part the next prune-only launch takes, per locality (gpu_plist::rollingPruningPart,
     * nbnxm/gpu_types_common.h:182-220) */
    int  rollingPart[2]     = { 0, 0 };
    int  rollingNumParts[2] = { 0, 0 };
    bool haveFreshList[2]   = { false, false };
    bool haveList[2]        = { false, false };
    bool ewaldTabulated     = false;
    /* nblib builds the pair list BEFORE it sets the atom properties (api/nblib/gmxsetup.cpp:316-320; mdrun does it the other way
     * round, mdlib/sim_util.cpp:1327-1366): a list that arrives before the atom data of its grid is kept until gpu_init_atomdata */
    bool                      haveAtomData = false;
    bool                      pending[2]   = { false, false };
    std::vector<nbnxn_sci_t>  pendingSci[2];
    std::vector<nbnxn_cj4_t>  pendingCj4[2];
    std::vector<nbnxn_excl_t> pendingExcl[2];
    /* staged outputs of the step in flight (NBStagingData, nbnxm/gpu_types_common.h:126-136) */
    bool                      didEnergy = false, didVirial = false;
    gmx_wallclock_gpu_nbnxn_t timings{};
};

namespace
{

void check(NbnxmGpu* nb, int rc, const char* what)
{
    /* the reference reports GPU failures through gmx_fatal (CU_RET_ERR, cuda/nbnxm_cuda.cu), never through return codes */
    if (rc != B200NB_OK)
    {
        gmx_fatal(FARGS, "B200 nonbonded backend, %s: %s", what, nb && nb->h ? b200nb_last_error(nb->h) : "no context");
    }
}

int locIndex(gmx::InteractionLocality iloc)
{
    return iloc == gmx::InteractionLocality::Local ? 0 : 1;
}

/* interaction_const_t + PairlistParams + nbat params -> b200nb_set_params / b200nb_set_vdw: the translation init_nbparam /
 * set_cutoff_parameters do for NBParamGpu (nbnxm_gpu_data_mgmt.cpp:166-245, cuda/nbnxm_cuda_data_mgmt.cu:121-227) */
void setParameters(NbnxmGpu* nb, const interaction_const_t* ic, const PairlistParams& listParams, const nbnxn_atomdata_t::Params& nbatParams)
{
    const int          ntWithFiller = nbatParams.numTypes; /* the last type is the zero-parameter filler (atomdata.cpp:456) */
    const int          nt           = ntWithFiller - 1;
    std::vector<float> nbfp(static_cast<size_t>(nt) * nt * 2);
    for (int i = 0; i < nt; i++)
    {
        for (int j = 0; j < nt; j++)
        {
            nbfp[(i * nt + j) * 2]     = nbatParams.nbfp[(i * ntWithFiller + j) * 2];
            nbfp[(i * nt + j) * 2 + 1] = nbatParams.nbfp[(i * ntWithFiller + j) * 2 + 1];
        }
    }
    b200nb_params_t p{};
    p.ntypes      = nt;
    p.nbfp_host   = nbfp.data();
    p.rc          = ic->rcoulomb;
    p.rlist_outer = listParams.rlistOuter;
    p.rlist_inner = listParams.useDynamicPruning ? listParams.rlistInner : listParams.rlistOuter;
    if (EEL_PME_EWALD(ic->eeltype))
    {
        p.eeltype = B200NB_EEL_EWALD; /* analytical real-space correction: gpu_is_kernel_ewald_analytical() answers true */
    }
    else if (EEL_RF(ic->eeltype))
    {
        p.eeltype = B200NB_EEL_RF;
    }
    else if (ic->eeltype == eelCUT)
    {
        p.eeltype = B200NB_EEL_CUT;
    }
    else
    {
        gmx_fatal(FARGS, "Electrostatics type %s is not supported by the B200 nonbonded backend", eel_names[ic->eeltype]);
    }
    p.epsfac     = ic->epsfac;
    p.k_rf       = ic->k_rf;
    p.c_rf       = ic->c_rf;
    p.ewald_beta = ic->ewaldcoeff_q;
    p.sh_ewald   = ic->sh_ewald;
    p.disp_cpot  = ic->dispersion_shift.cpot;
    p.rep_cpot   = ic->repulsion_shift.cpot;
    /* nbnxn_atomdata_params_init detected the rule already (atomdata.cpp:462-525): ljcrGEOM = 0 in its enum, ours 1 = geometric */
    p.comb_rule           = nbatParams.comb_rule == ljcrGEOM ? 1 : 2;
    p.max_tiles_per_entry = 0;
    check(nb, b200nb_set_params(nb->h, &p), "set_params");
    /* nbnxn_gpu_pick_ewald_kernel_type (nbnxm_gpu_data_mgmt.cpp:118-154): analytical by default on current GPUs, tabulated on request */
    nb->ewaldTabulated = p.eeltype == B200NB_EEL_EWALD && ic->coulombEwaldTables != nullptr && !ic->coulombEwaldTables->tableF.empty()
                         && (getenv("GMX_GPU_NB_TAB_EWALD") != nullptr || getenv("GMX_CUDA_NB_TAB_EWALD") != nullptr);
    if (nb->ewaldTabulated)
    {
        const EwaldCorrectionTables& t = *ic->coulombEwaldTables;
        check(nb, b200nb_set_ewald_table(nb->h, t.tableF.data(), static_cast<int>(t.tableF.size()), t.scale), "set_ewald_table");
    }

    const bool ljPme = ic->vdwtype == evdwPME;
    if (ic->vdwtype != evdwCUT && !ljPme)
    {
        gmx_fatal(FARGS, "VdW type %s is not supported by the B200 nonbonded backend", evdw_names[ic->vdwtype]);
    }
    if (ic->vdw_modifier != eintmodPOTSHIFT && ic->vdw_modifier != eintmodNONE || ic->rvdw < ic->rcoulomb || ljPme)
    {
        b200nb_vdw_t v{};
        switch (ic->vdw_modifier)
        {
            case eintmodNONE:
            case eintmodPOTSHIFT: v.vdw_modifier = B200NB_VDW_POTSHIFT; break;
            case eintmodFORCESWITCH: v.vdw_modifier = B200NB_VDW_FORCESWITCH; break;
            case eintmodPOTSWITCH: v.vdw_modifier = B200NB_VDW_POTSWITCH; break;
            default: gmx_fatal(FARGS, "VdW modifier %s is not supported by the B200 nonbonded backend", eintmod_names[ic->vdw_modifier]);
        }
        v.rvdw        = ic->rvdw;
        v.rvdw_switch = ic->rvdw_switch;
        v.disp_c2     = ic->dispersion_shift.c2;
        v.disp_c3     = ic->dispersion_shift.c3;
        v.rep_c2      = ic->repulsion_shift.c2;
        v.rep_c3      = ic->repulsion_shift.c3;
        v.sw_c3       = ic->vdw_switch.c3;
        v.sw_c4       = ic->vdw_switch.c4;
        v.sw_c5       = ic->vdw_switch.c5;
        if (ljPme)
        {
            v.ljpme_comb_rule = ic->ljpme_comb_rule == eljpmeGEOM ? 1 : 2;
            v.ewaldcoeff_lj   = ic->ewaldcoeff_lj;
            v.sh_lj_ewald     = ic->sh_lj_ewald;
        }
        check(nb, b200nb_set_vdw(nb->h, &v), "set_vdw");
    }
}

void uploadList(NbnxmGpu* nb, int l, const nbnxn_sci_t* sci, size_t nsci, const nbnxn_cj4_t* cj4, size_t ncj4, const nbnxn_excl_t* excl, size_t nexcl)
{
    check(nb,
          b200nb_upload_pairlist(nb->h, l, reinterpret_cast<const b200nb_sci_t*>(sci), static_cast<int>(nsci),
                                 reinterpret_cast<const b200nb_cj4_t*>(cj4), static_cast<int>(ncj4),
                                 reinterpret_cast<const b200nb_excl_t*>(excl), static_cast<int>(nexcl)),
          "upload_pairlist");
    nb->haveList[l]      = true;
    nb->haveFreshList[l] = true; /* the library prunes a fresh list at upload: nothing left for the first prune-only call */
    nb->rollingPart[l]   = 0;
}

void slotRange(const NbnxmGpu* nb, gmx::AtomLocality aloc, int* begin, int* end)
{
    /* getGpuAtomRange, nbnxm/gpu_common_utils.h: local = [0, natoms_local), non-local = the rest */
    switch (aloc)
    {
        case gmx::AtomLocality::Local: *begin = 0, *end = nb->numSlotsLocal; break;
        case gmx::AtomLocality::NonLocal: *begin = nb->numSlotsLocal, *end = nb->numSlots; break;
        default: *begin = 0, *end = nb->numSlots; break;
    }
}

} // namespace

namespace Nbnxm
{

NbnxmGpu* gpu_init(const gmx::DeviceStreamManager& /* deviceStreamManager */,
                   const interaction_const_t*      ic,
                   const PairlistParams&           listParams,
                   const nbnxn_atomdata_t*         nbat,
                   bool                            bLocalAndNonlocal)
{
    auto* nb          = new NbnxmGpu();
    nb->twoLocalities = bLocalAndNonlocal;
    int device        = 0;
    if (const char* env = getenv("B200NB_DEVICE"))
    {
        device = atoi(env);
    }
    /* the library issues its work on its own stream; the reference's DeviceStreamManager streams serve its other GPU modules */
    if (b200nb_create(&nb->h, device) != B200NB_OK)
    {
        delete nb;
        gmx_fatal(FARGS, "Could not create the B200 nonbonded context on device %d (is a B200 visible?)", device);
    }
    setParameters(nb, ic, listParams, nbat->params());
    return nb;
}

void gpu_free(NbnxmGpu* nb)
{
    if (nb == nullptr)
    {
        return;
    }
    b200nb_destroy(nb->h);
    delete nb;
}

void gpu_pme_loadbal_update_param(const nonbonded_verlet_t* nbv, const interaction_const_t* ic)
{
    if (!nbv || !nbv->useGpu())
    {
        return;
    }
    NbnxmGpu* nb = nbv->gpu_nbv;
    setParameters(nb, ic, nbv->pairlistSets().params(), nbv->nbat->params());
    nb->haveList[0] = nb->haveList[1] = false; /* the next search uploads lists for the new cut-off */
}

void gpu_init_atomdata(NbnxmGpu* nb, const nbnxn_atomdata_t* nbat)
{
    GMX_RELEASE_ASSERT(nbat->XFormat == nbatXYZQ, "the GPU path stores x and q interleaved (atomdata.cpp:659-662)");
    nb->numSlots      = nbat->numAtoms();
    nb->numSlotsLocal = nbat->natoms_local;
    check(nb, b200nb_set_grid_atoms(nb->h, nb->numSlots, nbat->x().data(), nbat->params().type.data()), "set_grid_atoms");
    check(nb, b200nb_set_shift_vec(nb->h, reinterpret_cast<const float*>(nbat->shift_vec.data())), "set_shift_vec");
    nb->haveList[0] = nb->haveList[1] = false;
    nb->haveAtomData                  = true;
    for (int l = 0; l < 2; l++)
    {
        if (nb->pending[l])
        {
            uploadList(nb, l, nb->pendingSci[l].data(), nb->pendingSci[l].size(), nb->pendingCj4[l].data(), nb->pendingCj4[l].size(),
                       nb->pendingExcl[l].data(), nb->pendingExcl[l].size());
            nb->pending[l] = false;
            nb->pendingSci[l].clear(), nb->pendingCj4[l].clear(), nb->pendingExcl[l].clear();
        }
    }
}

void gpu_upload_shiftvec(NbnxmGpu* nb, const nbnxn_atomdata_t* nbatom)
{
    check(nb, b200nb_set_shift_vec(nb->h, reinterpret_cast<const float*>(nbatom->shift_vec.data())), "set_shift_vec");
}

void gpu_init_pairlist(NbnxmGpu* nb, const NbnxnPairlistGpu* h_nblist, gmx::InteractionLocality iloc)
{
    const int l = locIndex(iloc);
    if (!nb->haveAtomData)
    {
        nb->pendingSci[l].assign(h_nblist->sci.begin(), h_nblist->sci.end());
        nb->pendingCj4[l].assign(h_nblist->cj4.begin(), h_nblist->cj4.end());
        nb->pendingExcl[l].assign(h_nblist->excl.begin(), h_nblist->excl.end());
        nb->pending[l] = true;
        return;
    }
    uploadList(nb, l, h_nblist->sci.data(), h_nblist->sci.size(), h_nblist->cj4.data(), h_nblist->cj4.size(), h_nblist->excl.data(),
               h_nblist->excl.size());
}

void gpu_copy_xq_to_gpu(NbnxmGpu* nb, const nbnxn_atomdata_t* nbdata, gmx::AtomLocality aloc)
{
    int begin, end;
    slotRange(nb, aloc, &begin, &end);
    check(nb, b200nb_copy_xq_grid(nb->h, nbdata->x().data(), begin, end), "copy_xq_grid");
}

void gpu_launch_kernel(NbnxmGpu* nb, const gmx::StepWorkload& stepWork, gmx::InteractionLocality iloc)
{
    const int l = locIndex(iloc);
    if (!nb->haveList[l])
    {
        return; /* canSkipNonbondedWork, nbnxm/gpu_common.h */
    }
    const int flags = (stepWork.computeEnergy ? B200NB_FLAG_ENERGY : 0) | (stepWork.computeVirial ? B200NB_FLAG_VIRIAL : 0);
    nb->didEnergy   = stepWork.computeEnergy;
    nb->didVirial   = stepWork.computeVirial;
    check(nb, b200nb_launch_force(nb->h, l, flags), "launch_force");
    nb->haveFreshList[l] = false;
}

void gpu_launch_kernel_pruneonly(NbnxmGpu* nb, gmx::InteractionLocality iloc, int numParts)
{
    const int l = locIndex(iloc);
    if (!nb->haveList[l])
    {
        return;
    }
    if (nb->haveFreshList[l])
    {
        /* cuda/nbnxm_cuda.cu:608-622: a fresh list is pruned as a whole -- done at upload */
        nb->haveFreshList[l] = false;
        return;
    }
    if (nb->rollingNumParts[l] != numParts)
    {
        nb->rollingNumParts[l] = numParts;
        nb->rollingPart[l]     = 0;
    }
    check(nb, b200nb_launch_prune(nb->h, l, nb->rollingPart[l], numParts), "launch_prune");
    nb->rollingPart[l] = (nb->rollingPart[l] + 1) % numParts;
}

void gpu_launch_cpyback(NbnxmGpu* nb, nbnxn_atomdata_t* nbatom, const gmx::StepWorkload& /* stepWork */, gmx::AtomLocality aloc)
{
    int begin, end;
    slotRange(nb, aloc, &begin, &end);
    check(nb, b200nb_get_f_grid(nb->h, nbatom->out[0].f.data(), begin, end), "get_f_grid");
}

bool gpu_try_finish_task(NbnxmGpu*                nb,
                         const gmx::StepWorkload& stepWork,
                         gmx::AtomLocality        aloc,
                         real*                    e_lj,
                         real*                    e_el,
                         gmx::ArrayRef<gmx::RVec> shiftForces,
                         GpuTaskCompletion /* completionKind */,
                         gmx_wallcycle* /* wcycle */)
{
    check(nb, b200nb_synchronize(nb->h), "synchronize");
    /* gpu_reduce_staged_outputs (gpu_common.h:249-277): energies and shift forces are reduced once, with the local task */
    if (aloc == gmx::AtomLocality::Local && (stepWork.computeEnergy || stepWork.computeVirial))
    {
        float  fshift[B200NB_SHIFTS * 3] = { 0 };
        double e[2]                      = { 0, 0 };
        check(nb, b200nb_get_outputs(nb->h, stepWork.computeVirial ? fshift : nullptr, stepWork.computeEnergy ? e : nullptr), "get_outputs");
        if (stepWork.computeEnergy)
        {
            *e_lj += static_cast<real>(e[0]);
            *e_el += static_cast<real>(e[1]);
        }
        if (stepWork.computeVirial)
        {
            for (int s = 0; s < SHIFTS; s++)
            {
                shiftForces[s][XX] += fshift[3 * s];
                shiftForces[s][YY] += fshift[3 * s + 1];
                shiftForces[s][ZZ] += fshift[3 * s + 2];
            }
        }
    }
    return true;
}

float gpu_wait_finish_task(NbnxmGpu*                nb,
                           const gmx::StepWorkload& stepWork,
                           gmx::AtomLocality        aloc,
                           real*                    e_lj,
                           real*                    e_el,
                           gmx::ArrayRef<gmx::RVec> shiftForces,
                           gmx_wallcycle*           wcycle)
{
    gpu_try_finish_task(nb, stepWork, aloc, e_lj, e_el, shiftForces, GpuTaskCompletion::Wait, wcycle);
    return 0.0F;
}

void gpu_clear_outputs(NbnxmGpu* nb, bool /* computeVirial */)
{
    check(nb, b200nb_clear_outputs(nb->h), "clear_outputs");
}

gmx_wallclock_gpu_nbnxn_t* gpu_get_timings(NbnxmGpu* nb)
{
    return nb ? &nb->timings : nullptr;
}

void gpu_reset_timings(nonbonded_verlet_t* nbv)
{
    if (nbv && nbv->gpu_nbv)
    {
        nbv->gpu_nbv->timings = gmx_wallclock_gpu_nbnxn_t{};
    }
}

int gpu_min_ci_balanced(NbnxmGpu* /* nb */)
{
    /* the library re-balances on the device (k_pack: entries of <= max_tiles cluster pairs, largest first), so the CPU search
     * need not split its sci entries for us (get_nsubpair_target, pairlist.cpp:2485-2587) */
    return 0;
}

bool gpu_is_kernel_ewald_analytical(const NbnxmGpu* nb)
{
    return !nb->ewaldTabulated;
}

void setupGpuShortRangeWork(NbnxmGpu* /* nb */, const gmx::GpuBonded* /* gpuBonded */, gmx::InteractionLocality /* iLocality */) {}

bool haveGpuShortRangeWork(const NbnxmGpu* nb, gmx::AtomLocality aLocality)
{
    return nb->haveList[aLocality == gmx::AtomLocality::NonLocal ? 1 : 0];
}

/* ---- CUDA_FUNC_QUALIFIER entry points of the GPU buffer-ops / GPU-update path (mdlib/sim_util.cpp:1043-1108): that path hands
 * DeviceBuffers of the reference's other GPU modules to the backend.  With this shim the coordinates arrive through
 * gpu_copy_xq_to_gpu and leave through gpu_launch_cpyback (the path nblib and the default `mdrun -nb gpu` take); the device-buffer
 * path of the library is b200nb_step, reached through the C ABI directly. */
const DeviceStream* gpu_get_command_stream(NbnxmGpu* /* nb */, gmx::InteractionLocality /* iloc */)
{
    return nullptr;
}
void* gpu_get_xq(NbnxmGpu* nb)
{
    /* mdrun passes this pointer (and gpu_get_f / gpu_get_fshift) on to gmx::GpuBonded::updateInteractionListsAndDeviceBuffers
     * (mdlib/sim_util.cpp:1340-1356) and never looks inside: our GpuBonded (shim/gpubonded_b200.cpp) takes it as the b200nb
     * context, whose xq / f / fshift buffers the bonded kernel of the same library works on */
    return nb ? static_cast<void*>(nb->h) : nullptr;
}
DeviceBuffer<gmx::RVec> gpu_get_f(NbnxmGpu* /* nb */)
{
    return DeviceBuffer<gmx::RVec>{};
}
DeviceBuffer<gmx::RVec> gpu_get_fshift(NbnxmGpu* /* nb */)
{
    return DeviceBuffer<gmx::RVec>{};
}
void* getGpuForces(NbnxmGpu* /* nb */)
{
    return nullptr;
}
void nbnxn_gpu_init_x_to_nbat_x(const Nbnxm::GridSet& /* gridSet */, NbnxmGpu* /* gpu_nbv */) {}
void nbnxn_gpu_x_to_nbat_x(const Nbnxm::Grid& /* grid */,
                           bool /* setFillerCoords */,
                           NbnxmGpu* /* gpu_nbv */,
                           DeviceBuffer<gmx::RVec> /* d_x */,
                           GpuEventSynchronizer* /* xReadyOnDevice */,
                           gmx::AtomLocality /* locality */,
                           int /* gridId */,
                           int /* numColumnsMax */)
{
    gmx_fatal(FARGS, "GPU buffer ops (GMX_USE_GPU_BUFFER_OPS) are not routed through the B200 nbnxm shim; use the default coordinate path");
}
void nbnxnInsertNonlocalGpuDependency(const NbnxmGpu* /* nb */, gmx::InteractionLocality /* interactionLocality */)
{
    /* local and non-local work share the library's one stream in this path: already ordered */
}
void nbnxn_wait_x_on_device(NbnxmGpu* nb)
{
    check(nb, b200nb_synchronize(nb->h), "synchronize");
}

} // namespace Nbnxm
