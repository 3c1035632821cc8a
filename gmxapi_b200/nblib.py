"""Host-side mirror of the nblib operator interface for the nonbonded path.

Same names, argument meaning and error behaviour as the reference's C++ API so that tests read like
api/nblib/tests/nbkernelsystem.cpp:
  NBKernelOptions   api/nblib/kerneloptions.h:86-112
  SimulationState   api/nblib/simulationstate.h (coordinates, box, topology: types / charges / nonbonded
                    parameters / exclusions)
  ForceCalculator   api/nblib/forcecalculator.h: ForceCalculator(SimulationState, NBKernelOptions),
                    compute(coordinates, forces), updatePairList(coordinates, box)
Setup follows GmxSetupDirector::setupGmxForceCalculator (api/nblib/gmxsetup.cpp:306-325).
`useGpu` is real here -- and mandatory: the reference notes "currently GPUs are not supported"
(kerneloptions.h:88-89); this implementation has no CPU path at all.
"""
import enum
from dataclasses import dataclass

import numpy as np

from . import lib as _lib
from .systems import ONE_4PI_EPS0, ewald_beta, ewald_beta_lj, rf_constants


class CoulombType(enum.Enum):  # api/nblib/kerneloptions.h:72-79
    Pme = 0
    Cutoff = 1
    ReactionField = 2


class VdwModifier(enum.Enum):  # interaction_const_t::vdw_modifier (eintmodPOTSHIFT / FORCESWITCH / POTSWITCH)
    PotentialShift = 0
    ForceSwitch = 1
    PotentialSwitch = 2


class LjPme(enum.Enum):  # vdwtype = evdwPME with ljpme_combination_rule (eljpmeGEOM / eljpmeLB); the real-space part only
    Off = 0
    Geometric = 1
    LorentzBerthelot = 2


@dataclass
class NBKernelOptions:
    useGpu: bool = True
    pairlistCutoff: float = 1.0
    computeVirialAndEnergy: bool = False
    coulombType: CoulombType = CoulombType.Pme
    # extensions beyond the reference's options: dynamic pruning radii (PairlistParams, pairlistparams.h:105-131)
    rlistOuter: float = 0.0
    rlistInner: float = 0.0
    maxTilesPerEntry: int = 0  # list balancing granularity (split_sci_entry, pairlist.cpp:2077-2194); 0 = library default
    epsilonRf: float = 1.0  # interaction_const_t::epsilon_rf as gmxsetup.cpp:251-253 leaves it; 0 = infinity
    # LJ modifier and VdW cut-off (the reference's nblib fixes potential shift and rvdw = rcoulomb, gmxsetup.cpp:231-240;
    # mdrun reaches the others through the .mdp options vdw-modifier, rvdw, rvdw-switch)
    vdwModifier: VdwModifier = VdwModifier.PotentialShift
    vdwCutoff: float = 0.0  # rvdw <= pairlistCutoff (= rcoulomb); 0 = the same
    vdwSwitch: float = 0.0  # rvdw-switch
    ljPme: LjPme = LjPme.Off  # subtract the LJ-PME grid part in real space (the mesh part is not this library's business)
    ljPmeEwaldCoeff: float = 0.0  # ewaldcoeff_lj; 0 = calc_ewaldcoeff_lj(rvdw, 1e-3)
    # NBKernelOptions::useTabulatedEwaldCorr (api/nblib/kerneloptions.h): the tabulated instead of the analytical Ewald correction
    # (the reference's EL_EWALD_TAB GPU kernels).  The table is F(r) = erf(beta r)/r^2 - 2 beta/sqrt(pi) exp(-beta^2 r^2)/r at
    # 2000 points per nm; a caller that has the reference's own table (interaction_const_t::coulombEwaldTables) passes that one to
    # NbnxmGpu.set_ewald_table instead.
    useTabulatedEwaldCorr: bool = False
    # interaction_const_t::sh_ewald = erfc(beta rc) / rc, the Coulomb potential shift mdrun applies with coulomb-modifier =
    # potential-shift (initCoulombEwaldParameters, mdtypes/interaction_const.cpp); the reference's nblib leaves it 0
    ewaldPotentialShift: bool = False
    device: int = 0


class InputException(ValueError):  # nblib::InputException (api/nblib/exception.h)
    pass


def interaction_kwargs(options):
    """setupInteractionConst (api/nblib/gmxsetup.cpp:226-284) -> b200nb_set_params keywords."""
    rc = float(options.pairlistCutoff)
    kw = dict(epsfac=ONE_4PI_EPS0)
    if options.coulombType == CoulombType.Pme:
        beta = float(np.float32(ewald_beta(rc, 1e-5)))
        kw.update(eeltype=_lib.EEL_EWALD, ewald_beta=beta)
        if options.ewaldPotentialShift:
            from math import erfc
            kw.update(sh_ewald=float(np.float32(erfc(beta * rc) / rc)))
    elif options.coulombType == CoulombType.Cutoff:
        k, c = rf_constants(rc, eps_rf=1.0)
        kw.update(eeltype=_lib.EEL_CUT, k_rf=k, c_rf=c)
    elif options.coulombType == CoulombType.ReactionField:
        k, c = rf_constants(rc, eps_rf=options.epsilonRf)
        kw.update(eeltype=_lib.EEL_RF, k_rf=k, c_rf=c)
    else:
        raise InputException("Unsupported electrostatic interaction")
    return kw


def ewald_force_table(beta, rc, scale=2000.0):
    """(tableF, scale): the Ewald correction force -(d/dr)(erf(beta r)/r) at r = i / scale, i = 0 .. rc*scale + 2 (the quantity
    EwaldCorrectionTables::tableF holds, tables/forcetable.cpp; F(0) = 0)."""
    from math import erf
    n = int(rc * scale) + 3
    r = np.arange(n, dtype=np.float64) / scale
    f = np.zeros(n, np.float64)
    rr = r[1:]
    f[1:] = np.array([erf(beta * v) for v in rr]) / rr ** 2 - 2.0 * beta / np.sqrt(np.pi) * np.exp(-(beta * rr) ** 2) / rr
    return f.astype(np.float32), float(scale)


def configure_interactions(nb, nonbonded_parameters, options, rlist_outer=None):
    """b200nb_set_params (+ b200nb_set_vdw) from NBKernelOptions: what setupInteractionConst and gpu_init make of the
    interaction constants (api/nblib/gmxsetup.cpp:226-284, nbnxm/nbnxm_gpu_data_mgmt.cpp:166-245)."""
    rc = float(options.pairlistCutoff)
    kw = interaction_kwargs(options)
    rvdw = float(options.vdwCutoff) or rc
    if rvdw > rc:
        raise InputException("vdwCutoff must not exceed pairlistCutoff")
    vk = _lib.vdw_modifier_constants(options.vdwModifier.value, rvdw, options.vdwSwitch)
    nb.set_params(nonbonded_parameters, rc, rlist_outer=rlist_outer or options.rlistOuter or rc,
                  rlist_inner=options.rlistInner or 0.0, max_tiles_per_entry=options.maxTilesPerEntry,
                  disp_cpot=vk["disp_cpot"], rep_cpot=vk["rep_cpot"], **kw)
    if options.ljPme != LjPme.Off:
        if options.vdwModifier != VdwModifier.PotentialShift:
            raise InputException("LJ-PME goes with the potential-shift modifier only")
        bl = float(np.float32(options.ljPmeEwaldCoeff or ewald_beta_lj(rvdw, 1e-3)))
        nb.set_vdw(options.vdwModifier.value, rvdw, options.vdwSwitch, vk, ljpme=options.ljPme.value, ewaldcoeff_lj=bl,
                   sh_lj_ewald=_lib.lj_ewald_shift(bl, rvdw))
    elif options.vdwModifier != VdwModifier.PotentialShift or rvdw < rc:
        nb.set_vdw(options.vdwModifier.value, rvdw, options.vdwSwitch, vk)
    if options.useTabulatedEwaldCorr and options.coulombType == CoulombType.Pme:
        nb.set_ewald_table(*ewald_force_table(kw["ewald_beta"], rc))


class SimulationState:
    def __init__(self, coordinates, box, types, charges, nonbondedParameters, excl_off=None, excl_idx=None):
        self.coordinates = np.ascontiguousarray(coordinates, dtype=np.float32).reshape(-1, 3)
        box = np.ascontiguousarray(box, dtype=np.float32)
        self.box = box.reshape(3, 3) if box.size == 9 else box.reshape(3)  # edge lengths, or the triclinic box matrix (rows a, b, c)
        self.types = np.ascontiguousarray(types, dtype=np.int32)
        self.charges = np.ascontiguousarray(charges, dtype=np.float32)
        self.nonbondedParameters = np.ascontiguousarray(nonbondedParameters, dtype=np.float32)
        n = self.types.shape[0]
        if self.coordinates.shape[0] != n or self.charges.shape[0] != n:
            raise InputException("coordinates, types and charges must have the same length")
        if not np.all(np.isfinite(self.coordinates)):
            raise InputException("non-finite coordinates")  # simulationstate.cpp checks isRealValued
        if excl_off is None:
            excl_off = np.arange(n + 1, dtype=np.int32)
            excl_idx = np.arange(n, dtype=np.int32)
        self.excl_off = np.ascontiguousarray(excl_off, dtype=np.int32)
        self.excl_idx = np.ascontiguousarray(excl_idx, dtype=np.int32)

    @classmethod
    def from_system(cls, s):
        box = s.box_matrix if np.any(getattr(s, "box_offdiag", np.zeros(3)) != 0) else s.box
        return cls(s.x, box, s.types, s.q, s.nbfp, s.excl_off, s.excl_idx)


class ForceCalculator:
    def __init__(self, state, options):
        if not options.useGpu:
            raise InputException("gmxapi_b200 only provides the GPU nonbonded path (useGpu must be true)")
        self.options = options
        self.state = state
        rc = float(options.pairlistCutoff)
        self.nb = _lib.NbnxmGpu(options.device)
        configure_interactions(self.nb, state.nonbondedParameters, options)
        self.nb.set_atoms(state.types, state.charges, state.excl_off, state.excl_idx)
        self._set_particles_on_grid(state.coordinates, state.box)
        self.nb.build_pairlist()  # constructPairList, gmxsetup.cpp:299-303

    def _set_particles_on_grid(self, coordinates, box):
        # GmxForceCalculator::setParticlesOnGrid, api/nblib/gmxcalculator.cpp:85-103
        box = np.ascontiguousarray(box, dtype=np.float32)
        diag = np.ascontiguousarray(np.diag(box.reshape(3, 3))) if box.size == 9 else box.reshape(3)
        if not np.all(diag > 0):
            raise InputException("box must be positive")
        self.nb.set_box(box)  # 3 edge lengths, or a 3 x 3 triclinic box matrix: the grid covers the brick spanned by its diagonal
        self.nb.put_on_grid(coordinates, np.zeros(3, np.float32), diag)

    def compute(self, coordinates=None, forces=None):
        """Returns forces[n,3] (float32). With computeVirialAndEnergy also keeps .energies / .shiftForces."""
        x = self.state.coordinates if coordinates is None else coordinates
        flags = (_lib.FLAG_ENERGY | _lib.FLAG_VIRIAL) if self.options.computeVirialAndEnergy else 0
        f, fs, elj, eel = self.nb.compute(x, flags, forces)
        self.shiftForces, self.energies = fs, (elj, eel)
        return f

    def updatePairList(self, coordinates, box):
        # ForceCalculator::updatePairList (forcecalculator.cpp:62-67) re-grids; unlike the reference (which
        # leaves the old list in place) we also rebuild the list so it stays valid.
        self._set_particles_on_grid(coordinates, box)
        self.nb.build_pairlist()
