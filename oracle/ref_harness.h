/* TEST INFRASTRUCTURE ONLY. C ABI of oracle/_ref/libgmxref_nbnxm.so (see ref_harness.cpp). */
#ifndef B200NB_ORACLE_REF_HARNESS_H
#define B200NB_ORACLE_REF_HARNESS_H
#ifdef __cplusplus
extern "C" {
#endif

enum { GMXREF_KERNEL_PLAINC_4X4 = 0, GMXREF_KERNEL_SIMD_4XN = 1, GMXREF_KERNEL_SIMD_2XNN = 2, GMXREF_KERNEL_GPUREF_8X8X8 = 3 };
enum { GMXREF_EEL_CUT = 0, GMXREF_EEL_RF = 1, GMXREF_EEL_EWALD_ANA = 2, GMXREF_EEL_EWALD_TAB = 3 };

typedef struct
{
    int          natoms;
    const float* x;        /* 3*natoms */
    float        box[3];   /* rectangular */
    int          ntypes;
    const float* nbfp;     /* ntypes*ntypes*2: 6*C6, 12*C12 (the fr->nbfp convention) */
    const int*   type;     /* natoms */
    const float* q;        /* natoms */
    const int*   excl_off; /* natoms+1, CSR; each atom's list includes itself */
    const int*   excl_idx;
} gmxref_system;

typedef struct
{
    float rc;           /* rvdw = rcoulomb */
    float rlist;        /* outer list radius */
    float rlist_inner;  /* <=0: no dynamic pruning */
    int   nstlist_prune;
    int   eeltype;      /* GMXREF_EEL_* */
    float epsfac, k_rf, c_rf, ewaldcoeff, sh_ewald;
    float disp_cpot, rep_cpot; /* potential-shift constants: -rc^-6, -rc^-12 or 0 */
    int   kernel;       /* GMXREF_KERNEL_* */
    int   comb_rule;    /* enbnxninitcombrule: 0 detect, 1 geom, 2 LB, 3 none */
    int   nthreads;
    int   exact_atom_flags; /* 0: all atoms flagged VdW+Q (bench default) */
    int   put_in_box;
    int   min_ilist_count;  /* GPU list balancing target, 0 = none */
    /* LJ modifier and twin-range cut-off (interaction_const_t::vdw_modifier, rvdw, rvdw_switch) */
    float rvdw;             /* 0: = rc; < rc: VdW cut-off check (Ewald kernels only, kerneldispatch.cpp:175-200) */
    int   vdw_modifier;     /* 0 potential shift (disp_cpot / rep_cpot above), 1 force switch, 2 potential switch */
    float rvdw_switch;
    float disp_c2, disp_c3, rep_c2, rep_c3; /* shift_consts_t::c2, c3 as force_switch_constants makes them (forcerec.cpp:787-801) */
    float sw_c3, sw_c4, sw_c5;              /* switch_consts_t (forcerec.cpp:803-816) */
    /* LJ-PME: vdwtype = evdwPME with ljpme_comb_rule GEOM (1) or LB (2; plain-C kernel only, kerneldispatch.cpp:219-233);
     * the caller also passes comb_rule = 1 / 2 for the atom-data initialisation (nbnxm_setup.cpp:399-410) */
    int   ljpme;
    float ewaldcoeff_lj, sh_lj_ewald;
    /* triclinic cell: the off-diagonal elements box[YY][XX], box[ZZ][XX], box[ZZ][YY] of the lower-triangular box matrix
     * (gmxref_system::box is its diagonal); all zero = rectangular */
    float box_offdiag[3];
    /* free-energy perturbation: per atom 1 = perturbed (atom info bit SET_CGINFO_FEP; the search then builds the perturbed pair
     * lists, nbnxm/pairlist.cpp make_fep_list, and takes those pairs out of the cluster-pair list); NULL = none */
    const unsigned char* perturbed;
} gmxref_params;

int    gmxref_simd_width(void);
int    gmxref_default_simd_kernel(void);
float  gmxref_ewald_coeff(float rc, float rtol);
float  gmxref_simd_rsq(float xi, float yi, float zi, float xj, float yj, float zj);
void*  gmxref_create(const gmxref_system* s, const gmxref_params* p);
void   gmxref_destroy(void* h);
int    gmxref_regrid_research(void* h, double* tGrid, double* tSearch);
void   gmxref_setup_times(void* h, double* tGrid, double* tSearch);
int    gmxref_compute(void* h, const float* x, int want_energy, int want_virial, float* f, float* fshift, float* energies);
double gmxref_time_kernel(void* h, int want_energy, int nwarm, int niter);
double gmxref_time_step(void* h, int want_energy, int nwarm, int niter);
int    gmxref_grid_order(void* h, int* out, int cap);
void   gmxref_grid_dims(void* h, int* ncx, int* ncy, float* cellx, float* celly, int* natomsPadded);
void   gmxref_list_stats(void* h, long long* nClusterPairs, long long* nAtomPairsComputed, int* na_ci, int* na_cj);
long long gmxref_pair_set(void* h, float rc, int* pairs, long long cap);
int    gmxref_gpu_list(void* h, int* nsci, int* ncj4, int* nexcl, int* nslots, int* sci, int* cj4, unsigned* excl, float* xq, int* type);
int    gmxref_grid_forces(void* h, float* f, int cap_slots);
int    gmxref_ewald_table(void* h, float* table_f, int cap, float* scale);
int    gmxref_bench_coordinates1000(float* out, int cap_atoms, float* box_edge);

/* The perturbed pair lists the reference's search built (gmxref_params::perturbed; PairlistSet::fepLists(), one t_nblist per search
 * thread), concatenated: call with cap_nri = cap_nrj = 0 for the sizes, then with arrays (jindex: nri + 1). */
int gmxref_fep_list(void* h, int* nri, int* nrj, int cap_nri, int cap_nrj, int* iinr, int* shift, int* jindex, int* jjnr, char* excl_fep);

/* The reference's perturbed-pair (free-energy) kernel, gmxlib/nonbonded/nb_free_energy.cpp gmx_nb_free_energy_kernel, on a pair list
 * given by the caller in t_nblist form (nri i-entries {iinr, shift, jindex}, jjnr, excl_fep: 1 = the pair interacts, 0 = excluded).
 * Reaction-field / plain cut-off or Ewald electrostatics and cut-off LJ with potential shift (the flavours our FEP kernel covers). */
typedef struct
{
    float rc;                 /* rcoulomb = rvdw */
    float epsfac, k_rf, c_rf;
    float disp_cpot, rep_cpot;
    float lambda_coul, lambda_vdw;
    float sc_alpha;           /* 0: no soft-core */
    int   sc_power;           /* 1 or 2 */
    float sc_sigma, sc_sigma_min;
    int   sc_coul;            /* soft-core also on Coulomb (t_lambda::bScCoul) */
    float ewaldcoeff, sh_ewald; /* ewaldcoeff > 0: Ewald electrostatics (eelPME; the kernel subtracts the tabulated long-range part) */
    float rvdw_switch;          /* > 0: vdw_modifier = eintmodPOTSWITCH from rvdw_switch to rc (the caller passes disp_cpot = rep_cpot = 0) */
    int   ljpme_comb_rule;      /* 0: cut-off LJ; 1 / 2: vdwtype = evdwPME with the geometric / Lorentz-Berthelot grid rule (eljpmeGEOM / eljpmeLB) */
    float ewaldcoeff_lj, sh_lj_ewald; /* interaction_const_t::ewaldcoeff_lj, sh_lj_ewald */
    float rvdw;                 /* > 0: rvdw < rcoulomb = rc (what PME load balancing leaves behind; disp_cpot / rep_cpot / sh_lj_ewald are for rvdw) */
} gmxref_fep_params;
int gmxref_fep_kernel(int natoms, const float* x, const float* shift_vec, int ntypes, const float* nbfp, const int* typeA, const int* typeB,
                      const float* qA, const float* qB, int nri, const int* iinr, const int* shift, const int* jindex, const int* jjnr,
                      const char* excl_fep, const gmxref_fep_params* p, float* f, float* fshift, float* out4);

/* The reference's CPU functions for the listed interactions its GPU bonded module covers (listed_forces/bonded.cpp
 * calculateSimpleBond; pairs.cpp do_pairs, analytical force-only path for LJ14).  kind: 0 bonds, 1 angles, 2 Urey-Bradley, 3 proper
 * dihedrals, 4 Ryckaert-Bellemans, 5 improper (harmonic) dihedrals, 6 periodic improper dihedrals, 7 LJ-14 pairs.  iatoms: per
 * interaction {parameter index, atoms...} (t_ilist); params6: 6 floats per parameter set ({r0, k}; {theta0, k}; {theta0, ktheta,
 * r13, kUB}; {phi0, k, multiplicity}; {C0..C5}; {xi0, k}; as 3; {c6, c12}).  f[natoms*3], fshift[45*3] are overwritten. */
#define GMXREF_BONDED_KINDS 8
int gmxref_bonded(int kind, int nbonds, const int* iatoms, int nparams, const float* params6, int natoms, const float* x,
                  const float* q, const float* box9, float epsfac_fudge, int want_virial_energy, float* f, float* fshift,
                  double* energy2);

#ifdef __cplusplus
}
#endif
#endif
