/* b200nb -- B200-native (sm_100a) short-range nonbonded path, C ABI.
 *
 * Drop-in boundary for the nbnxm hot path of kassonlab/gmxapi (GROMACS 2021): everything below is what a
 * thin C++ shim implementing the reference's GPU sub-interface (src/gromacs/nbnxm/nbnxm_gpu.h:138-355,
 * gpu_data_mgmt.h:72-138) and nblib's GmxForceCalculator (api/nblib/gmxcalculator.cpp:70-103) binds to.
 * Plain C types only; every function returns 0 on success or a negative B200NB_ERR_* code, and
 * b200nb_last_error() gives the message (the reference aborts via gmx_fatal / exceptions instead,
 * e.g. grid.cpp:425 "Lost particles while sorting", cuda/nbnxm_cuda.cu:142-149 grid-size overflow).
 *
 * All paths run on the GPU; there is no CPU fallback.  Pointers named *_host are host memory, *_dev
 * device memory; functions taking `on_device` accept either.
 *
 * Reference paths below are relative to /root/reference/src/gromacs unless they start with api/.
 */
#ifndef B200NB_H
#define B200NB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NB_VERSION 1
#define B200NB_SHIFTS 45  /* pbcutil/ishift.h:40-48 */
#define B200NB_CENTRAL 22 /* pbcutil/ishift.h:47 */
#define B200NB_CLUSTER 8  /* nbnxm/pairlistparams.h:66 c_nbnxnGpuClusterSize */

enum
{
    B200NB_OK            = 0,
    B200NB_ERR_ARG       = -1,
    B200NB_ERR_CUDA      = -2,
    B200NB_ERR_STATE     = -3,
    B200NB_ERR_CAPACITY  = -4,
    B200NB_ERR_LOSTATOMS = -5
};

enum
{
    B200NB_EEL_CUT   = 0, /* eelCUT: evaluated as reaction-field with k_rf = 0 (nbnxm/kerneldispatch.cpp:168-171) */
    B200NB_EEL_RF    = 1, /* eelRF */
    B200NB_EEL_EWALD = 2  /* eelPME/eelEWALD real space, analytical correction (cuda/nbnxm_cuda_kernel.cuh:573-594) */
};

enum
{
    B200NB_VDW_POTSHIFT    = 0, /* eintmodPOTSHIFT / eintmodNONE: plain cut-off, constant potential shift (disp_cpot, rep_cpot) */
    B200NB_VDW_FORCESWITCH = 1, /* eintmodFORCESWITCH: vdwktLJFORCESWITCH (nbnxm/kerneldispatch.cpp:208-216, cuda evdwSwitch kernels) */
    B200NB_VDW_POTSWITCH   = 2  /* eintmodPOTSWITCH */
};

enum
{
    B200NB_FLAG_ENERGY = 1, /* StepWorkload::computeEnergy */
    B200NB_FLAG_VIRIAL = 2  /* StepWorkload::computeVirial: accumulate the 45 shift forces */
};

/* Interaction parameters: mdtypes/interaction_const.h:107-172 -> NBParamGpu (nbnxm/gpu_types_common.h:63-124),
 * filled the way set_cutoff_parameters / init_nbparam do (nbnxm_gpu_data_mgmt.cpp:166-186,
 * cuda/nbnxm_cuda_data_mgmt.cu:121-227) plus PairlistParams (nbnxm/pairlistparams.h:105-131). */
typedef struct
{
    int          ntypes;      /* atom types, without the filler type the library appends (atomdata.cpp:456) */
    const float* nbfp_host;   /* ntypes*ntypes*2 floats {6*C6, 12*C12}: the fr->nbfp convention (atomdata.cpp:498-506) */
    float        rc;          /* rvdw == rcoulomb */
    float        rlist_outer; /* PairlistParams::rlistOuter: search radius */
    float        rlist_inner; /* PairlistParams::rlistInner: dynamic-pruning radius; >= rlist_outer disables pruning */
    int          eeltype;     /* B200NB_EEL_* */
    float        epsfac;      /* interaction_const_t::epsfac */
    float        k_rf, c_rf;
    float        ewald_beta;  /* ewaldcoeff_q */
    float        sh_ewald;
    float        disp_cpot;   /* dispersion_shift.cpot (-rc^-6 with potential shift) */
    float        rep_cpot;    /* repulsion_shift.cpot  (-rc^-12) */
    int          comb_rule;   /* 0 = detect geometric rule with tol 1e-5 (atomdata.cpp:462-525), 1 = force geometric, 2 = type table */
    int          max_tiles_per_entry; /* list balancing granularity (pairlist.cpp:2077-2194 split_sci_entry); 0 = default */
} b200nb_params_t;

/* Van der Waals modifier and VdW cut-off: the part of interaction_const_t (mdtypes/interaction_const.h:107-172) that
 * init_interaction_const derives in mdlib/forcerec.cpp:850-874 (force_switch_constants :787-801, potential_switch_constants
 * :803-816) and NBParamGpu carries as dispersion_shift / repulsion_shift / vdw_switch / rvdw_switch / rvdw_sq
 * (nbnxm/gpu_types_common.h:63-124).  Optional: without it the library evaluates LJ cut-off + potential shift, rvdw = rc. */
typedef struct
{
    int   vdw_modifier; /* B200NB_VDW_* */
    float rvdw;         /* <= rc (= rcoulomb).  rvdw < rc is the twin-range case PME load balancing produces
                           (gpu_pme_loadbal_update_param, nbnxm_gpu_data_mgmt.cpp:188-205); 0 means rc */
    float rvdw_switch;  /* start of the switching range */
    float disp_c2, disp_c3, rep_c2, rep_c3; /* shift_consts_t::c2, c3 of dispersion_shift / repulsion_shift (force switch); the
                                               matching cpot values go in b200nb_params_t::disp_cpot / rep_cpot */
    float sw_c3, sw_c4, sw_c5;              /* switch_consts_t vdw_switch (potential switch) */
    /* LJ-PME (vdwtype = evdwPME): the real-space kernel subtracts the grid part of the dispersion (evdwTypeEWALDGEOM / EWALDLB
     * kernels, cuda/nbnxm_cuda_kernel_utils.cuh calculate_lj_ewald_comb_*; kernels_reference/kernel_ref_inner.h:207-250).  The
     * per-type grid parameters are derived from the diagonal of nbfp as set_lj_parameter_data does (atomdata.cpp:291-322).
     * Needs vdw_modifier = B200NB_VDW_POTSHIFT; the reciprocal-space part is outside this library. */
    int   ljpme_comb_rule; /* 0 = no LJ-PME, 1 = geometric (eljpmeGEOM), 2 = Lorentz-Berthelot (eljpmeLB) */
    float ewaldcoeff_lj;   /* interaction_const_t::ewaldcoeff_lj */
    float sh_lj_ewald;     /* interaction_const_t::sh_lj_ewald (mdlib/forcerec.cpp:709-717) */
} b200nb_vdw_t;

typedef struct b200nb_context b200nb_t;

/* ---- lifetime: Nbnxm::gpu_init / gpu_free (cuda/nbnxm_cuda_data_mgmt.cu:242,395) ------------------------ */
int         b200nb_create(b200nb_t** out, int device);
void        b200nb_destroy(b200nb_t* h);
const char* b200nb_last_error(const b200nb_t* h);
/* The CUDA stream (cudaStream_t) all work of this context is issued on; callers that bring their own
 * device buffers order against it (the reference exposes DeviceStream objects the same way). */
void* b200nb_stream(b200nb_t* h);
/* Use the caller's stream instead (the reference hands its DeviceStreamManager streams to gpu_init the same way,
 * cuda/nbnxm_cuda_data_mgmt.cu:242-291). NULL returns to a private stream. The caller keeps ownership. */
int   b200nb_set_stream(b200nb_t* h, void* cuda_stream);
int   b200nb_synchronize(b200nb_t* h);

/* ---- parameters: init_nbparam / gpu_pme_loadbal_update_param (nbnxm_gpu_data_mgmt.cpp:225) -------------- */
int b200nb_set_params(b200nb_t* h, const b200nb_params_t* p);
/* LJ modifier / VdW cut-off; call after b200nb_set_params (which resets them to potential shift, rvdw = rc).  Only the
 * force kernel changes: the pair list is built for rlist >= rc either way. */
int b200nb_set_vdw(b200nb_t* h, const b200nb_vdw_t* v);

/* Tabulated Ewald force correction (the reference's EL_EWALD_TAB kernels; chosen there by nbnxn_gpu_pick_ewald_kernel_type,
 * nbnxm_gpu_data_mgmt.cpp:118-154; table upload init_ewald_coulomb_force_table :71-83): table_f_host = n points of
 * interaction_const_t::coulombEwaldTables->tableF, scale = its tableScale (points per nm).  Call after b200nb_set_params with
 * Ewald electrostatics; n = 0 returns to the analytical correction (the default).  Applies to the plain LJ kernels. */
int b200nb_set_ewald_table(b200nb_t* h, const float* table_f_host, int n, float scale);

/* ---- atoms: nbnxn_atomdata_set (atomdata.cpp:955-977) + the exclusion ListOfLists handed to
 * constructPairlist (nbnxm.h:263).  Exclusions are CSR over LOCAL atom indices, each atom's list contains
 * itself (bench_system.cpp:192-195).  natoms counts every atom this rank holds (home + halo). */
int b200nb_set_atoms(b200nb_t* h, int natoms, const int* type_host, const float* q_host,
                     const int* excl_off_host, const int* excl_idx_host);

/* ---- box: forcerec shift_vec = calc_shifts(box) (api/nblib/gmxsetup.cpp:286-292, pbcutil/pbc.cpp:1187).
 * pbc_dims[d] = 0 switches periodic images off along d (a dimension decomposed over ranks,
 * pairlist.cpp:3168-3176). */
int b200nb_set_box(b200nb_t* h, const float box[3], const int pbc_dims[3]);
/* the same for a TRICLINIC cell: box9 = the lower-triangular box matrix of the reference, rows a, b, c (`matrix box`,
 * pbcutil/pbc.cpp; limits of check_box apply); atoms in the brick put_atoms_in_box leaves them in.  The shift vectors follow
 * calc_shifts (pbc.cpp:1187-1202), the x-shift range nbnxm/pairlist.cpp:3181-3188, the largest list radius max_cutoff2
 * (pbc.cpp:179-208).  Fully periodic, single domain. */
int b200nb_set_box_triclinic(b200nb_t* h, const float box9[9]);

/* ---- gridding: nbnxn_put_on_grid (nbnxm.cpp:58-75 -> gridset.cpp:135-241 -> grid.cpp:103-1445).
 * Grid 0 = home atoms, grid 1 = halo atoms (nbnxn_put_on_grid_nonlocal, nbnxm.cpp:77-95).  Atoms
 * [atom_begin, atom_end) of x (natoms*3 floats, original order) are binned on 2-D columns between
 * lower/upper, sorted into 8-atom clusters, and their bounding boxes computed, all on the device.
 * density <= 0: computed from the home grid (grid.cpp:138-141). */
int b200nb_put_on_grid(b200nb_t* h, int grid_index, const float lower[3], const float upper[3], int atom_begin,
                       int atom_end, float density, const float* x, int x_on_device);

/* ---- pair search: nonbonded_verlet_t::constructPairlist (pairlist.cpp:3921-4199) + gpu_init_pairlist
 * (nbnxm_gpu_data_mgmt.cpp:251-311): device-built cluster-pair list at rlist_outer with exclusion masks,
 * split into balanced entries, followed by the fresh-list prune to rlist_inner
 * (cuda/nbnxm_cuda.cu:510-517).  Grid 0 x grid 0 is a half list, grid 0 x grid 1 a full list. */
int b200nb_build_pairlist(b200nb_t* h);

/* ---- reference-built grid and list: the drop-in path behind Nbnxm::gpu_* (shim/nbnxm_b200.cpp) ---------------------
 * With mdrun / nblib in front, gridding and pair search stay on the CPU, where the reference does them for its own CUDA
 * backend, and the backend receives their products.  These four calls take exactly what the reference hands to
 * gpu_init_atomdata / gpu_init_pairlist / gpu_copy_xq_to_gpu / gpu_launch_cpyback, so no reference call site changes.
 * In this mode the context has no atom-order view: use b200nb_launch_force / _launch_prune / _clear_outputs / _get_outputs
 * with the calls below; b200nb_step, b200nb_compute, b200nb_set_x and b200nb_get_f need b200nb_put_on_grid instead. */
typedef struct
{
    int sci, shift, cj4_ind_start, cj4_ind_end;
} b200nb_sci_t; /* = nbnxn_sci_t, nbnxm/pairlist.h:174-188 */
typedef struct
{
    int cj[4];
    struct
    {
        unsigned int imask;
        int          excl_ind;
    } imei[2];
} b200nb_cj4_t; /* = nbnxn_cj4_t, nbnxm/pairlist.h:190-205 */
typedef struct
{
    unsigned int pair[32];
} b200nb_excl_t; /* = nbnxn_excl_t, nbnxm/pairlist.h:208-225 */
/* gpu_init_atomdata (nbnxm_gpu_data_mgmt.cpp; cuda/nbnxm_cuda_data_mgmt.cu:297-348): nslots = nbat->numAtoms(), xq_host =
 * nbat->x() (nbatXYZQ: 4 floats per slot), type_host = nbat->params().type (fillers carry the zero-parameter type `ntypes`). */
int b200nb_set_grid_atoms(b200nb_t* h, int nslots, const float* xq_host, const int* type_host);
/* gpu_init_pairlist (nbnxm_gpu_data_mgmt.cpp:251-311): the NbnxnPairlistGpu of one locality (sci, cj4, excl arrays as they
 * are); followed on the device by the fresh-list prune to rlist_inner and the re-packing for the force kernel. */
int b200nb_upload_pairlist(b200nb_t* h, int locality, const b200nb_sci_t* sci, int nsci, const b200nb_cj4_t* cj4, int ncj4,
                           const b200nb_excl_t* excl, int nexcl);
/* gpu_upload_shiftvec (cuda/nbnxm_cuda_data_mgmt.cu:283-295): nbat->shift_vec, 45 x 3 floats, as calc_shifts(box) made them. */
int b200nb_set_shift_vec(b200nb_t* h, const float* shift_vec_host);
/* gpu_copy_xq_to_gpu (cuda/nbnxm_cuda.cu:395-465): slots [slot_begin, slot_end) of nbat->x(); asynchronous on the stream. */
int b200nb_copy_xq_grid(b200nb_t* h, const float* xq_host, int slot_begin, int slot_end);
/* gpu_launch_cpyback, force part (cuda/nbnxm_cuda.cu:720-814): grid-ordered forces of slots [slot_begin, slot_end) into
 * nbat->out[0].f (3 floats per slot, overwritten); asynchronous on the stream: b200nb_synchronize before reading. */
int b200nb_get_f_grid(b200nb_t* h, float* f_host, int slot_begin, int slot_end);

/* ---- per step ------------------------------------------------------------------------------------------- */
/* gpu_copy_xq_to_gpu + nbnxn_gpu_x_to_nbat_x (cuda/nbnxm_cuda.cu:395,829): new coordinates (original
 * order, natoms*3) -> grid-ordered device layout.  atom range [atom_begin, atom_end) lets the caller
 * update home and halo atoms separately. */
int b200nb_set_x(b200nb_t* h, const float* x, int x_on_device, int atom_begin, int atom_end);
/* gpu_clear_outputs (cuda/nbnxm_cuda_data_mgmt.cu:350-377) */
int b200nb_clear_outputs(b200nb_t* h);
/* gpu_launch_kernel (cuda/nbnxm_cuda.cu:484-591). locality: 0 = local (home-home), 1 = non-local (home-halo),
 * -1 = both. */
int b200nb_launch_force(b200nb_t* h, int locality, int flags);
/* gpu_launch_kernel_pruneonly (cuda/nbnxm_cuda.cu:593-718): rolling prune of part `part` of `num_parts`. */
int b200nb_launch_prune(b200nb_t* h, int locality, int part, int num_parts);
/* gpu_launch_cpyback + atomdata_add_nbat_f_to_f (cuda/nbnxm_cuda.cu:720-814, atomdata.cpp:1425-1468,
 * mdlib/gpuforcereduction_impl.cu:70-104): f_out[a] (+)= f_grid[cell[a]] for atoms [atom_begin, atom_end). */
int b200nb_get_f(b200nb_t* h, float* f, int f_on_device, int accumulate, int atom_begin, int atom_end);
/* gpu_wait_finish_task's staged outputs (gpu_common.h:249-277): shift forces [45*3] and {E_lj, E_el};
 * values are ADDED to the caller's buffers like the reference does.  Either pointer may be NULL. */
int b200nb_get_outputs(b200nb_t* h, float* fshift_host, double* energies_host);

/* ---- the nblib call: GmxForceCalculator::compute (api/nblib/gmxcalculator.cpp:70-83):
 * x_host (natoms*3) in, f_host (natoms*3) overwritten; fshift_host[135]/energies_host[2] overwritten when
 * not NULL.  Synchronous.  Pinned host buffers (cudaHostAlloc / cudaHostRegister: what gmx::HostVector gives the
 * reference's GPU path, nbnxm_setup.cpp:417-418) are read and written in place by the kernels over PCIe; pageable
 * buffers are staged through the context's own pinned scratch. */
int b200nb_compute(b200nb_t* h, const float* x_host, int flags, float* f_host, float* fshift_host,
                   double* energies_host);
/* The same step with coordinates and forces resident on the device (the reference's GPU buffer-ops path:
 * nbnxn_gpu_x_to_nbat_x + gpu_launch_kernel + GpuForceReduction, mdlib/sim_util.cpp:1043-1108): three launches
 * (x -> grid layout fused with gpu_clear_outputs; force kernel; grid-ordered f -> atom order), asynchronous on the
 * context's stream.  f_dev is overwritten. */
int b200nb_step(b200nb_t* h, const float* x_dev, int flags, float* f_dev);

/* ---- halo exchange helpers: packSendBufKernel / unpackRecvBufKernel (domdec/gpuhaloexchange_impl.cu:77-131).
 * index_dev: n local atom indices. pack: out[k] = x[index[k]] + shift; unpack: f[index[k]] += in[k]. */
int b200nb_halo_pack_x(b200nb_t* h, const float* x_dev, const int* index_dev, int n, const float shift[3],
                       float* out_dev);
int b200nb_halo_unpack_f(b200nb_t* h, float* f_dev, const int* index_dev, int n, const float* in_dev);

/* ---- domain-decomposed step over peer-memory halo windows ---------------------------------------------------
 * Replaces dd_move_x / dd_move_f (domdec/domdec.cpp:260-460) and GpuHaloExchange (domdec/gpuhaloexchange_impl.cu:133-444)
 * on the per-step path, for a 1-D (x-slab) decomposition (b200nb_dd_set_plan) and for 2-D / 3-D decompositions with up to 16
 * halo LINKS per rank (b200nb_dd_set_links; e.g. the 13 half-shell neighbours of a 2 x 2 x 2 grid).  Each rank creates ONE
 * window in its device memory and hands its CUDA IPC handle to its neighbours; the neighbours' kernels store halo coordinates /
 * halo forces straight into it over NVLink and raise a per-link flag, the owner's kernels wait on the flags (bounded to 10 s).
 * No host synchronisation and no library collective inside a step.  b200nb_halo_pack_x / _unpack_f above remain for callers
 * that move the data themselves (the pair-search step, where the halo composition changes, goes through them). */
/* ipc_handle_out: 64 bytes (cudaIpcMemHandle_t) for peers in other processes; window_dev_out: the device pointer, for peers
 * in this process.  max_halo bounds the halo atoms, max_send the send entries (atom, link) of the plans that may be set later. */
int b200nb_dd_create_window(b200nb_t* h, int max_halo, int max_send, void* ipc_handle_out, void** window_dev_out);
/* Opens a neighbour's window as peer number `peer` (0..15).  Give either the peer's IPC handle or, inside one process, its
 * window pointer; peer_max_halo = the peer's max_halo.  A neighbour reached over several links is opened once.
 * With b200nb_dd_set_plan: peer 0 = the -x neighbour (gets our halo coordinates), peer 1 = the +x neighbour. */
int b200nb_dd_open_peer(b200nb_t* h, int peer, const void* ipc_handle, void* same_process_window, int peer_max_halo);
/* One halo link = one (neighbour offset) connection; link k of every rank belongs to the same offset, so this rank's receive
 * side of link k and its source's send side of link k are the two ends of one connection (they share flag index k). */
typedef struct
{
    /* send side: nsend of our home atoms go to peer `send_peer`, `shift` added (dd_move_x's box shift, domdec.cpp:300-318);
     * they become halo atoms peer_halo_offset .. peer_halo_offset + nsend - 1 of the destination */
    int        send_peer, nsend;
    const int* send_idx_host;
    float      shift[3];
    int        peer_halo_offset;
    /* receive side: nrecv halo atoms arrive from peer `recv_peer` (they follow the halo atoms of the links before this one);
     * the forces we compute on them return to the source's send entries peer_entry_offset .. + nrecv - 1 (= the number of
     * entries the source sends over its links before this one); fshift_index: shift-force slot that also receives those
     * forces when the atoms arrived across a periodic edge (domdec.cpp:426-458), else -1 */
    int recv_peer, nrecv, peer_entry_offset, fshift_index;
} b200nb_dd_link_t;
/* Plan of the current search interval: local atoms [0, nhome) home, [nhome, nhome + nhalo) halo in link order. */
int b200nb_dd_set_links(b200nb_t* h, int nhome, int nhalo, int nlinks, const b200nb_dd_link_t* links);
/* The 1-D form: one link, sending to peer 0, receiving from peer 1.  send_idx_host[nsend]: home atoms sent to the -x neighbour
 * with `shift` added; halo_fshift_index: shift-force slot that also receives the forces computed HERE on the halo atoms when
 * those arrived across the periodic edge (the last slab), or -1. */
int b200nb_dd_set_plan(b200nb_t* h, int nhome, int nhalo, const int* send_idx_host, int nsend, const float shift[3],
                       int halo_fshift_index);
/* One decomposed step, asynchronous on the context's stream, replayed as one CUDA graph of 7 kernels on two branches:
 * home x -> grid + clear | local kernel || push of the halo x into the neighbours' windows | wait, halo x -> grid | non-local
 * kernel | push of the halo forces || wait, add, un-sort.  x_home in, f_home out: nhome*3 floats, device or pinned host memory.  Every neighbour must
 * call it the same number of times. */
int b200nb_dd_step(b200nb_t* h, const float* x_home, float* f_home, int flags);
/* after b200nb_synchronize: B200NB_ERR_STATE if a halo flag timed out in any step since the last call */
int b200nb_dd_status(b200nb_t* h);

/* ---- repartitioning on the device (the coordinate-dependent work of dd_partition_system; gmxapi_b200/csrc/dd_partition.cu) ----
 * Replaces CPU code of the reference: domdec/redistribute.cpp:505-760 (dd_redistribute_cg: which atoms left, where to, compaction,
 * packing per direction), pbcutil/pbc.cpp put_atoms_in_box, domdec/partition.cpp:2058-2440 + domdec/domdec.cpp setup_dd_communication
 * (send lists = home atoms within the cut-off of a face), domdec/localtopology.cpp make_exclusions_zone + ga2la (exclusions of the local
 * atoms in local indices).  All pointers named *_dev are device memory of this context's GPU; coordinates stay on the device, the
 * host sees counts.  Messages between ranks (leavers, halo indices) are the caller's: search-step traffic of a few thousand atoms. */
/* x-slab decomposition: wraps the home atoms into the box (in place) and classifies each: 0 stays, 1 goes to the -x neighbour,
 * 2 to the +x neighbour, 3 moved more than one domain (the caller treats a non-zero count of 3 as fatal, like the reference).
 * bounds_host: nranks + 1 slab boundaries, float32(k * box_x / nranks). */
int b200nb_dd_wrap_classify(b200nb_t* h, float* x_dev, int n, const float box[3], const float* bounds_host, int nranks, int rank, int* code_dev);
/* the same for an N-D grid of domains (grid[3] cells, this rank at coords[3]): code = 9 (ox+1) + 3 (oy+1) + (oz+1) with the offset
 * -1 / 0 / +1 of the atom's new cell from this rank's per dimension (13 = stays; the other cell of a two-cell dimension is +1),
 * 27 = moved more than one domain */
int b200nb_dd_wrap_classify_nd(b200nb_t* h, float* x_dev, int n, const float box[3], const int grid[3], const int coords[3], int* code_dev);
/* code 1 for the home atoms the rank seeing this domain at half-shell offset[3] needs: within rlist of the lower face where
 * offset = +1 (x - lo < rlist), of the upper face where offset = -1 (hi - x <= rlist); else 0 */
int b200nb_dd_select_boundary(b200nb_t* h, const float* x_dev, int n, const float lo[3], const float hi[3], const int offset[3], float rlist,
                              int* code_dev);
/* code 1 for the home atoms within rlist of the slab's lower face (x - lo < rlist in float32: the halo of the -x neighbour), else 0 */
int b200nb_dd_select_lower_face(b200nb_t* h, const float* x_dev, int n, float lo, float rlist, int* code_dev);
/* stable partition of the indices 0 .. n-1 by code (0 .. ncodes-1, ncodes <= 32): idx_dev = the indices with code 0 in ascending
 * order, then those with code 1, ...; counts_host[k] = how many have code k (synchronises the stream for these few ints) */
int b200nb_dd_partition_indices(b200nb_t* h, const int* code_dev, int n, int ncodes, int* idx_dev, int* counts_host);
/* the message for a neighbour: per listed atom 4 ints {global index, x, y, z bit patterns} */
int b200nb_dd_pack_atoms(b200nb_t* h, const int* idx_dev, int m, const int* gid_dev, const float* x_dev, int* out4_dev);
/* the new home set in ascending global index: the stayers (listed by stay_idx_dev, ascending already) merged with the arrived
 * messages (any order); writes nstay + narrived global indices and coordinates */
int b200nb_dd_merge_home(b200nb_t* h, const int* stay_idx_dev, int nstay, const int* gid_dev, const float* x_dev, const int* arrived4_dev,
                         int narrived, int* gid_out_dev, float* x_out_dev);
int b200nb_dd_gather_int(b200nb_t* h, const int* idx_dev, int m, const int* in_dev, int* out_dev);
/* the replicated global topology (types, charges, exclusions as CSR over global atom indices), uploaded once */
int b200nb_dd_set_global_topology(b200nb_t* h, int nglobal, const int* type_host, const float* q_host, const int* excl_off_host,
                                  const int* excl_idx_host);
/* b200nb_set_atoms for home + halo atoms given by their global indices ON THE DEVICE: types and charges gathered, exclusions
 * renumbered to local indices (partners that are not local are dropped: they cannot be in range on this rank) */
int b200nb_dd_set_local_atoms(b200nb_t* h, const int* local_gid_dev, int nlocal);

/* ---- introspection used by the parity tests and the bench ------------------------------------------------ */
/* ---- perturbed (free-energy) pairs (gmxapi_b200/csrc/fep.cu) ----
 * Replaces CPU code of the reference: the free-energy kernel gmxlib/nonbonded/nb_free_energy.cpp:203-860, which the reference runs
 * on the host beside its GPU kernels (nonbonded_verlet_t::dispatchFreeEnergyKernel, nbnxm/kerneldispatch.cpp:458-567).  Built so
 * far: Ewald (real-space part: plain soft-cored 1/r - sh_ewald, minus the unsoftened erf(beta r)/r, :693-737) / reaction-field /
 * plain cut-off electrostatics, cut-off LJ with potential shift or potential switch (:613-625, on the soft-cored distance),
 * LJ-PME (b200nb_set_vdw's ljpme_comb_rule, both grid rules: cut-off on the plain distance, the grid potential at the cut-off and
 * the grid part of the dispersion taken off unsoftened, :586-611 and :725-770, evaluated directly instead of from the reference's
 * spline table), rvdw <= rcoulomb, soft-core (r-power 6, lambda power 1 or 2) or none; the force switch (not in the reference's
 * kernel either) returns B200NB_ERR_ARG.  As in the reference (nbnxn_atomdata_mask_fep) the caller
 * gives the perturbed atoms zero charge and a type without LJ in b200nb_set_atoms, so the cluster-pair kernels skip them, and
 * hands over the perturbed pair list in t_nblist form (mdtypes/nblist.h:117-137; what nbnxm/pairlist.cpp:1699-1872 make_fep_list
 * builds: every pair within the list radius with a perturbed atom, excl_fep = 0 for excluded pairs and for a perturbed atom listed
 * with itself).  b200nb_fep_launch goes between b200nb_clear_outputs / b200nb_launch_force and b200nb_get_f: it adds into the same
 * forces and shift forces. */
typedef struct
{
    float lambda_coul, lambda_vdw; /* nb_kernel_data_t::lambda[efptCOUL], [efptVDW] */
    float sc_alpha;                /* t_lambda::sc_alpha; 0 = no soft-core */
    int   sc_power;                /* t_lambda::sc_power: 1 or 2 */
    float sc_sigma, sc_sigma_min;  /* t_lambda::sc_sigma, sc_sigma_min (nm) */
    int   sc_coul;                 /* t_lambda::bScCoul: soft-core on Coulomb as well */
} b200nb_fep_params_t;
/* A / B state of every atom (atom order, b200nb_set_atoms' natoms): t_mdatoms::typeA/typeB/chargeA/chargeB */
int b200nb_fep_set_atoms(b200nb_t* h, const int* typeA_host, const int* typeB_host, const float* qA_host, const float* qB_host);
int b200nb_fep_upload_list(b200nb_t* h, int nri, const int* iinr, const int* shift, const int* jindex, const int* jjnr,
                           const signed char* excl_fep);
/* The same list built on the device from the gridded coordinates (after b200nb_put_on_grid, at every search step): the pairs
 * nbnxm/pairlist.cpp:1699-1872 make_fep_list cuts out of the cluster-pair list, here a search of its own over the cluster bounding
 * boxes (the cluster-pair path never sees the perturbed atoms' interactions: their charge and LJ are masked).  Perturbed atoms =
 * those whose A and B type or charge differ in b200nb_fep_set_atoms.  A pair of two perturbed atoms is listed from the lower atom
 * index; j-atoms within an entry in grid order.  Rectangular and triclinic cells, one domain; nri / nrj: entries and pairs (may be NULL). */
int b200nb_fep_build_list(b200nb_t* h, int* nri_out, int* nrj_out);
/* the current list (built or uploaded) back on the host, arrays sized from the nri / nrj of the build: iinr[nri], shift[nri],
 * jindex[nri + 1], jjnr[nrj], excl_fep[nrj] */
int b200nb_fep_get_list(b200nb_t* h, int* iinr_host, int* shift_host, int* jindex_host, int* jjnr_host, signed char* excl_fep_host);
int b200nb_fep_launch(b200nb_t* h, const b200nb_fep_params_t* p);
/* Vc, Vv, dV/dlambda_coul, dV/dlambda_vdw summed over the launches since the last call (read and reset) */
int b200nb_fep_get_outputs(b200nb_t* h, double out4_host[4]);
/* p != NULL: every b200nb_step / b200nb_compute launches the free-energy kernel with these parameters on the current list, inside
 * the captured step graph; NULL takes it out again */
int b200nb_fep_in_step(b200nb_t* h, const b200nb_fep_params_t* p);

/* ---- listed ("bonded") interactions on the nonbonded buffers (gmxapi_b200/csrc/bonded.cu) ----
 * Replaces gmx::GpuBonded (listed_forces/gpubonded.h:99-172) for the types it covers (fTypesOnGpu, gpubonded.h:84-85):
 * updateInteractionListsAndDeviceBuffers (gpubonded_impl.cu:178-310) -> b200nb_bonded_set_list, once per topology: the lists
 *   stay in atom order on the device and are mapped to the grid order inside the kernel, so search steps need no update;
 * launchKernel (gpubondedkernels.cu:823-861, exec_kernel_gpu :721-821) -> b200nb_bonded_launch, between b200nb_launch_force /
 *   b200nb_clear_outputs and b200nb_get_f: one fused kernel, a thread per interaction, forces added into the nonbonded force
 *   buffer, shift forces into the nonbonded replicas (flags & B200NB_FLAG_VIRIAL), energies per type (B200NB_FLAG_ENERGY);
 * launchEnergyTransfer / waitAccumulateEnergyTerms / clearEnergies (gpubonded_impl.cu:330-380) -> b200nb_bonded_get_energies.
 * iatoms: t_ilist rows {parameter index, atoms...} (topology/idef.h); params6: 6 floats per parameter set, the t_iparams fields
 * the type reads: bonds harmonic {rA, krA}; angles harmonic {thetaA deg, kA}; Urey-Bradley u_b {thetaA, kthetaA, r13A, kUBA};
 * proper and periodic improper dihedrals pdihs {phiA deg, cpA, mult}; Ryckaert-Bellemans rbdihs.rbcA[0..5]; improper dihedrals
 * harmonic {rA deg, krA}; 1-4 pairs lj14 {c6A, c12A} with the charges of b200nb_set_atoms.  One domain; any cell shape.
 * On a context whose atoms were given in grid order (b200nb_set_grid_atoms: a reference-built grid, no atom-order view) the atom
 * indices of iatoms are GRID SLOTS and the caller converts its lists at every search step, as the reference does. */
enum
{
    B200NB_BONDED_BONDS = 0,    /* F_BONDS */
    B200NB_BONDED_ANGLES,       /* F_ANGLES */
    B200NB_BONDED_UREY_BRADLEY, /* F_UREY_BRADLEY */
    B200NB_BONDED_PDIHS,        /* F_PDIHS */
    B200NB_BONDED_RBDIHS,       /* F_RBDIHS */
    B200NB_BONDED_IDIHS,        /* F_IDIHS */
    B200NB_BONDED_PIDIHS,       /* F_PIDIHS */
    B200NB_BONDED_LJ14,         /* F_LJ14 (+ F_COUL14) */
    B200NB_BONDED_KINDS
};
int b200nb_bonded_set_list(b200nb_t* h, int kind, int nbonds, const int* iatoms_host, int nparams, const float* params6_host);
/* GpuBonded::setPbc (gpubonded_impl.cu:312-316): the cell of the image search -- box9 the row-major GROMACS box matrix, the first
 * npbcdim dimensions periodic (setPbcAiuc) -- for callers whose context was not given the box (the Nbnxm::gpu_* interface passes
 * shift vectors only); NULL: the cell of b200nb_set_box / b200nb_set_box_triclinic (the default) */
int b200nb_bonded_set_pbc(b200nb_t* h, const float box9[9], int npbcdim);
/* epsfac_fudge: BondedCudaKernelParameters::electrostaticsScaleFactor = epsfac * fudgeQQ (gpubonded_impl.cu:105) */
int b200nb_bonded_launch(b200nb_t* h, int flags, float epsfac_fudge);
/* energies per kind, [B200NB_BONDED_KINDS] = the Coulomb part of the 1-4 pairs; summed over the launches since the last call */
int b200nb_bonded_get_energies(b200nb_t* h, double energies_host[B200NB_BONDED_KINDS + 1]);
/* enable != 0: every b200nb_step / b200nb_compute launches the bonded kernel between the force kernel and the un-sort, inside
 * the captured step graph (the reference puts its bonded kernel on the nonbonded stream the same way, mdlib/sim_util.cpp:1467-1478) */
int b200nb_bonded_in_step(b200nb_t* h, int enable, float epsfac_fudge);

typedef struct
{
    int       natoms, natoms_padded, nclusters;
    int       ncx, ncy;        /* home grid columns */
    long long ntiles_outer;    /* cluster pairs in the search list */
    long long ntiles_inner;    /* cluster pairs of the pruned list */
    long long nentries;        /* work units */
    int       comb_geometric;  /* 1 if the geometric-rule kernel is used */
    long long nlaunches;       /* kernels launched by this context so far */
    long long ntiles_packed;   /* what the force kernel evaluates on the packed list, in units of 64 pair lanes (2 per warp step) */
    long long nentries_nonlocal; /* half-entries of the packed non-local (home x halo) list the next non-local launch runs */
} b200nb_stats_t;
int b200nb_get_stats(b200nb_t* h, b200nb_stats_t* out);
/* one line describing the context (device, atoms, grid, cut-offs, kernel flavour, list sizes, launch mode) for the caller's log:
 * the reference prints its set-up the same way (nbnxm/nbnxm_setup.cpp:180-260 "Using a ... Verlet scheme", hardware report) */
int b200nb_describe(b200nb_t* h, char* buf, int cap);
/* slot -> original atom (-1 filler), natoms_padded ints: GridSet::atomIndices() */
int b200nb_get_grid_order(b200nb_t* h, int* atom_index_host, int cap);
/* the cluster pairs of the inner (or outer) list as (ci, shift, cj) triples */
long long b200nb_get_tiles(b200nb_t* h, int outer, int* tiles_host, long long cap);
/* the interacting atom pairs of the current inner list at radius r: non-excluded, r^2 < r*r, as
 * (i, j, shift) in ORIGINAL atom indices, i the shifted atom.  Returns the count; writes at most cap. */
long long b200nb_get_pairs(b200nb_t* h, float r, int* pairs_host, long long cap);
/* average duration in ms of the force kernel alone over niter launches (CUDA events on this context's
 * stream; flush_l2 != 0 writes a >L2 buffer between launches). */
int b200nb_time_force_kernel(b200nb_t* h, int locality, int flags, int nwarm, int niter, int flush_l2, float* ms_avg);
/* `niter` device-resident steps (b200nb_step) timed with CUDA events on the context's stream: average duration of the whole
 * step and of the force kernel inside it (events around its launch), L2 optionally flushed before each step. */
int b200nb_time_step(b200nb_t* h, const float* x_dev, float* f_dev, int flags, int nwarm, int niter, int flush_l2,
                     float* ms_step_avg, float* ms_force_avg);

#ifdef __cplusplus
}
#endif
#endif
