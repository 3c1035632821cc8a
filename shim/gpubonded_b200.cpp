/* gmx::GpuBonded (listed_forces/gpubonded.h:99-172) on libb200nb's C ABI: the class mdrun drives for the listed interactions it
 * offloads (mdlib/sim_util.cpp:1340-1356 updateInteractionListsAndDeviceBuffers at search steps, :1421-1426 / :1518-1523
 * setPbcAndlaunchKernel, :1548-1551 launchEnergyTransfer, :959-967 waitAccumulateEnergyTerms + clearEnergies).
 * Compiled against the reference's headers in place of listed_forces/gpubonded_impl.cu; host code only.
 *
 * The reference hands the nonbonded module's device buffers to this class (xqDevice = Nbnxm::gpu_get_xq(), forces, shift forces).
 * Our Nbnxm::gpu_get_xq (shim/nbnxm_b200.cpp) returns the b200nb context instead: the bonded kernel lives in the same library
 * and works on that context's xq / f / fshift, so the three arguments collapse into one opaque pointer and the call sites in
 * sim_util.cpp stay as they are.  The context behind the shim holds the REFERENCE-built grid, whose atoms are grid slots, so the
 * lists are converted with nbnxnAtomOrder at every update, as the reference does (a context that gridded the atoms itself keeps
 * the lists in atom order and needs no update at search steps: b200nb_bonded_set_list, include/b200nb.h). */
#include "gmxpre.h"

#include <algorithm>
#include <array>
#include <vector>

#include "gromacs/listed_forces/gpubonded.h"
#include "gromacs/mdtypes/enerdata.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/pbcutil/pbc.h"
#include "gromacs/topology/forcefieldparameters.h"
#include "gromacs/topology/idef.h"
#include "gromacs/topology/ifunc.h"
#include "gromacs/utility/fatalerror.h"

#include "b200nb.h"

namespace gmx
{

class GpuBonded::Impl
{
public:
    Impl(const gmx_ffparams_t& ffparams, float electrostaticsScaleFactor) : scale_(electrostaticsScaleFactor)
    {
        /* 6 floats per parameter set: the t_iparams fields the set's function type reads (include/b200nb.h) */
        params6_.assign(6 * static_cast<size_t>(ffparams.numTypes()), 0.0F);
        for (int t = 0; t < ffparams.numTypes(); t++)
        {
            const t_iparams& ip = ffparams.iparams[t];
            float*           p  = params6_.data() + 6 * static_cast<size_t>(t);
            switch (ffparams.functype[t])
            {
                case F_BONDS:
                case F_ANGLES:
                case F_IDIHS: p[0] = ip.harmonic.rA, p[1] = ip.harmonic.krA; break;
                case F_UREY_BRADLEY: p[0] = ip.u_b.thetaA, p[1] = ip.u_b.kthetaA, p[2] = ip.u_b.r13A, p[3] = ip.u_b.kUBA; break;
                case F_PDIHS:
                case F_PIDIHS: p[0] = ip.pdihs.phiA, p[1] = ip.pdihs.cpA, p[2] = static_cast<float>(ip.pdihs.mult); break;
                case F_RBDIHS:
                    for (int k = 0; k < NR_RBDIHS; k++) p[k] = ip.rbdihs.rbcA[k];
                    break;
                case F_LJ14: p[0] = ip.lj14.c6A, p[1] = ip.lj14.c12A; break;
                default: break;
            }
        }
        numTypes_ = ffparams.numTypes();
    }

    void check(int rc, const char* what) const
    {
        if (rc != B200NB_OK) gmx_fatal(FARGS, "b200nb %s failed: %s", what, h_ ? b200nb_last_error(h_) : "no context");
    }

    void update(ArrayRef<const int> nbnxnAtomOrder, const InteractionDefinitions& idef, void* context)
    {
        h_ = static_cast<b200nb_t*>(context);
        if (!h_) gmx_fatal(FARGS, "GpuBonded: the nonbonded module handed over no device context");
        /* The shim's context holds the reference-built grid (b200nb_set_grid_atoms): its atoms are grid slots, so the lists are
         * converted to that order here, at every search step -- convertIlistToNbnxnOrder, gpubonded_impl.cu:144-162.
         * nbnxnAtomOrder: the grid position of each atom (nonbonded_verlet_t::getGridIndices = GridSet::cells) */
        const int numAtoms = static_cast<int>(nbnxnAtomOrder.size());
        haveInteractions_ = false;
        std::vector<int> converted;
        for (int k = 0; k < numFTypesOnGpu; k++)
        {
            const InteractionList& il   = idef.il[fTypesOnGpu[k]];
            const int              nral = interaction_function[fTypesOnGpu[k]].nratoms;
            const int              n    = il.size() / (nral + 1);
            converted.assign(il.iatoms.begin(), il.iatoms.end());
            for (int i = 0; i < n; i++)
                for (int a = 1; a <= nral; a++)
                {
                    int& v = converted[i * (nral + 1) + a];
                    if (v < 0 || v >= numAtoms) gmx_fatal(FARGS, "GpuBonded: atom %d of a listed interaction is not on the grid", v);
                    v = nbnxnAtomOrder[v];
                }
            check(b200nb_bonded_set_list(h_, k, n, converted.data(), numTypes_, params6_.data()), "bonded_set_list");
            haveInteractions_ = haveInteractions_ || n > 0;
        }
    }

    b200nb_t*          h_ = nullptr;
    std::vector<float> params6_;
    int                numTypes_ = 0;
    float              scale_;
    bool               haveInteractions_ = false;
    double             energies_[B200NB_BONDED_KINDS + 1] = {};
};

/* fTypesOnGpu and the kinds of include/b200nb.h enumerate the same types in the same order */
static_assert(numFTypesOnGpu == B200NB_BONDED_KINDS, "interaction types of the GPU bonded module");
static_assert(fTypesOnGpu[B200NB_BONDED_BONDS] == F_BONDS && fTypesOnGpu[B200NB_BONDED_ANGLES] == F_ANGLES
                      && fTypesOnGpu[B200NB_BONDED_UREY_BRADLEY] == F_UREY_BRADLEY && fTypesOnGpu[B200NB_BONDED_PDIHS] == F_PDIHS
                      && fTypesOnGpu[B200NB_BONDED_RBDIHS] == F_RBDIHS && fTypesOnGpu[B200NB_BONDED_IDIHS] == F_IDIHS
                      && fTypesOnGpu[B200NB_BONDED_PIDIHS] == F_PIDIHS && fTypesOnGpu[B200NB_BONDED_LJ14] == F_LJ14,
              "order of the interaction types");

GpuBonded::GpuBonded(const gmx_ffparams_t& ffparams, const float electrostaticsScaleFactor, const DeviceContext& /*deviceContext*/,
                     const DeviceStream& /*deviceStream*/, gmx_wallcycle* /*wcycle*/) :
    impl_(new Impl(ffparams, electrostaticsScaleFactor))
{
}

GpuBonded::~GpuBonded() = default;

void GpuBonded::updateInteractionListsAndDeviceBuffers(ArrayRef<const int> nbnxnAtomOrder, const InteractionDefinitions& idef, void* d_xq,
                                                       DeviceBuffer<RVec> /*d_f*/, DeviceBuffer<RVec> /*d_fShift*/)
{
    impl_->update(nbnxnAtomOrder, idef, d_xq);
}

void GpuBonded::setPbc(PbcType pbcType, const matrix box, bool canMoleculeSpanPbc)
{
    /* gpubonded_impl.cu:312-316: setPbcAiuc(canMoleculeSpanPbc ? numPbcDimensions(pbcType) : 0, box, &pbcAiuc) */
    float box9[9];
    for (int i = 0; i < DIM; i++)
        for (int j = 0; j < DIM; j++) box9[DIM * i + j] = box[i][j];
    impl_->check(b200nb_bonded_set_pbc(impl_->h_, box9, canMoleculeSpanPbc ? numPbcDimensions(pbcType) : 0), "bonded_set_pbc");
}

bool GpuBonded::haveInteractions() const
{
    return impl_->haveInteractions_;
}

void GpuBonded::launchKernel(const gmx::StepWorkload& stepWork)
{
    const int flags = (stepWork.computeEnergy ? B200NB_FLAG_ENERGY : 0) | (stepWork.computeVirial ? B200NB_FLAG_VIRIAL : 0);
    impl_->check(b200nb_bonded_launch(impl_->h_, flags, impl_->scale_), "bonded_launch");
}

void GpuBonded::setPbcAndlaunchKernel(PbcType pbcType, const matrix box, bool canMoleculeSpanPbc, const gmx::StepWorkload& stepWork)
{
    setPbc(pbcType, box, canMoleculeSpanPbc);
    launchKernel(stepWork);
}

void GpuBonded::launchEnergyTransfer()
{
    /* read and reset in one call: clearEnergies() has nothing left to do */
    impl_->check(b200nb_bonded_get_energies(impl_->h_, impl_->energies_), "bonded_get_energies");
}

void GpuBonded::waitAccumulateEnergyTerms(gmx_enerdata_t* enerd)
{
    for (int k = 0; k < numFTypesOnGpu; k++)
        if (fTypesOnGpu[k] != F_LJ14) enerd->term[fTypesOnGpu[k]] += static_cast<real>(impl_->energies_[k]);
    gmx_grppairener_t* grppener = &enerd->grpp;
    GMX_RELEASE_ASSERT(grppener->nener == 1, "No energy group support for bondeds on the GPU");
    grppener->ener[egLJ14][0] += static_cast<real>(impl_->energies_[B200NB_BONDED_LJ14]);
    grppener->ener[egCOUL14][0] += static_cast<real>(impl_->energies_[B200NB_BONDED_KINDS]);
    for (double& e : impl_->energies_) e = 0.0;
}

void GpuBonded::clearEnergies() {}

} // namespace gmx
