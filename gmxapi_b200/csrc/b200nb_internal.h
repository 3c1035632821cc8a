/* Internal definitions shared by the b200nb translation units (not part of the C ABI). */
#ifndef B200NB_INTERNAL_H
#define B200NB_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "b200nb.h"

#define NB_CL 8      /* atoms per cluster */
#define NB_CELL 64   /* atoms per grid cell = 8 clusters (nbnxm/pairlistparams.h:69-77) */
#define NB_MIN_RSQ 3.82e-07f /* nbnxm/pairlist.h:146 c_nbnxnMinDistanceSquared */
#define NB_MAX_ENTRY_TILES 64 /* cluster pairs (and so packed tiles) per list entry at most: k_pack staging, ordering bins */
#define NB_ORDER_BINS (NB_MAX_ENTRY_TILES + 8)
#define NB_MAX_GROUP_TILES 512 /* staging capacity of one (i-cluster, shift) group in the search */
#define NB_OUT_COPIES 32   /* replicas of the shift-force / energy accumulators: per-entry atomics spread over them */
#define NB_FSHIFT_PITCH 136 /* floats per shift-force replica (45*3 rounded up) */
#define NB_PACK_TAIL 2    /* steps of dummy atoms behind the last step of every row of the packed list: the force kernel's gather runs
                             one step ahead of the arithmetic and its slot fetch two, without bounds checks */
#define NB_DUMMY_SLOTS 512 /* far-away filler atoms appended after the grids: padding targets of the packed list */

/* Device atom layout, grid (slot) order, slot = cluster*8 + k:
 *   xq     float4 {x, y, z, q} per slot  -- the reference's nbatXYZQ (nbnxm/atomdata.cpp:659-662); one 16-byte load per atom
 *   lj     float2 {sqrt(6 C6_ii), sqrt(12 C12_ii)} per slot (geometric rule), atype int per slot (type table)
 *   f      float4 per slot (16-byte vector reductions; one 128-byte line per cluster)
 * Lane mapping shared by the search, prune, pair-extraction and force kernels: lane = jl + 8*ih handles j-atom jl of the
 * j-cluster and the i-atoms 2*ih and 2*ih+1 of the i-cluster.  A tile's 64-bit mask is two 32-bit words: bit `lane` of
 * word w says pair (i-atom 2*ih + w, j-atom jl) interacts (nbnxn_excl_t, nbnxm/pairlist.h:208-225, re-indexed to our lanes). */

struct GridDesc
{
    int   ncx, ncy, ncol;
    float lower[3], upper[3];
    float cell[2], inv_cell[2];
    int   atom_begin, atom_end;
    int   cell0;  /* first 64-atom cell of this grid in the global numbering */
    int   ncells;
    int   col0;   /* offset of this grid's columns in the column arrays */
    int   valid;
};

/* One unit of work of the force kernel: a run of cluster pairs sharing the i-cluster and the shift
 * (the role nbnxn_sci_t plays in the reference, nbnxm/pairlist.h:174-188, at cluster granularity). */
struct Entry
{
    int ci;
    int shift_nmask; /* bits 0-7 shift index, bits 8-23 number of leading tiles that carry a mask; packed list only: bit 24 the
                        entry holds the i-cluster's self tile (Coulomb self term, kernel_outer.h:408-452), bit 25 which half of
                        the i-cluster (i-atoms 4*half .. 4*half + 3) the half-entry is for */
    int start, end;  /* tile range (cluster-pair lists); STEP range (packed list: 16 j slots and one 64-bit mask per step) */
};
#define NB_ENTRY_SHIFT(v) ((v)&255)
#define NB_ENTRY_NMASK(v) (((v) >> 8) & 0xffff)
#define NB_ENTRY_SELF(v) (((v) >> 24) & 1)
#define NB_ENTRY_HALF(v) (((v) >> 25) & 1)

struct NbParamsDev
{
    float rc2, rlist_outer2, rlist_inner2;
    float epsfac, k_rf, two_k_rf, c_rf;
    float beta, beta2, beta3, sh_ewald;
    float disp_cpot, rep_cpot;
    float self_sub; /* 0.5*c_rf or beta/sqrt(pi): kernels_simd_2xmm/kernel_outer.h:418-440 */
    float self_q2;  /* self_sub / epsfac (0 when epsfac is 0) */
    int   ntypes;   /* including the filler type */
    int   eeltype;
    /* LJ modifiers / twin-range cut-off (b200nb_set_vdw; interaction_const_t rvdw, rvdw_switch, *_shift, vdw_switch) */
    int   vdw_modifier;
    float rvdw2, rvdw_switch;
    float disp_c2, disp_c3, rep_c2, rep_c3;
    float sw_c3, sw_c4, sw_c5;
    int   ljpme; /* 0 none, 1 geometric, 2 Lorentz-Berthelot grid combination rule */
    float lje_coeff2, lje_coeff6_6, sh_lj_ewald;
    /* tabulated Ewald force correction (b200nb_set_ewald_table): {F[i], F[i+1] - F[i]} per table point, null = analytical */
    const float2* ewald_tab;
    float         tab_scale, tab_max;
};

struct PairList
{
    Entry*    entries = nullptr;
    int*      cj      = nullptr;
    uint64_t* mask    = nullptr;
    long long ntiles = 0, nentries = 0;
    size_t    cap_tiles = 0, cap_entries = 0;
};

/* The list the force kernel consumes: every entry of the (pruned) cluster-pair list re-packed at j-ATOM granularity, once
 * per HALF of its i-cluster.  A half-entry is four i-atoms (4*half .. 4*half + 3 of cluster ci) + shift against a run of
 * j-atom slots (any clusters), 16 per STEP; a j-atom is kept only when at least one of its 4 pairs with the i-quad is inside
 * the list radius, which raises the share of in-range lanes from 35 % (8x8 cluster pairs) over 54 % (8 i-atoms per j-atom) to
 * ~63 % at rlist = rc = 0.9 nm.  j-atoms with an excluded / self pair come first (`nmask` leading steps carry a 64-bit mask: bit
 * 16*k + j = pair (i-atom 4*half + k, j-atom j of the step) interacts); the tail of the last step points at far-away dummy
 * atoms, NB_DUMMY_SLOTS of them shared round-robin by the entries (the kernel's zero-valued force reductions for padding lanes
 * then never pile up on one address: same-address reductions serialise in L2).
 * Half-entry p = 2*e + half of outer entry e is PACKED into row p of ja / mask, the steps [p*row, (p+1)*row) with row = pitch/2 +
 * NB_PACK_TAIL (`staged` holds its header); `entries` holds the same headers sorted by descending step count, and the force kernel gives the half-entries 2w and
 * 2w + 1 of that order to the two halves of warp w: equal step counts up to the warps that straddle a size boundary, where
 * k_pad_partner extends the shorter one's j slots with dummy atoms to the longer one's steps. */
struct PackedList
{
    Entry*    entries = nullptr; /* headers in execution order (largest first) */
    Entry*    staged  = nullptr; /* headers in packing order */
    int*      ja      = nullptr; /* 16 slots per step */
    uint64_t* mask    = nullptr; /* one per step */
    int*      dest    = nullptr; /* per half-entry in packing order: its position in `entries` */
    int*      sizes   = nullptr; /* per half-entry in packing order: its steps (the ordering key) */
    int*      order_blk = nullptr; /* scratch of the stable ordering: NB_ORDER_BINS x blocks of 256 half-entries */
    long long nentries = 0;      /* HALF-entries: 2 x the entries of the outer list */
    int       pitch = 0;         /* tiles of 8 j slots reserved per half-entry (= max_tiles_per_entry, even): pitch/2 steps */
    int       row = 0;           /* steps per row: pitch/2 + NB_PACK_TAIL */
    size_t    cap_tiles = 0, cap_entries = 0;
};

/* Peer-memory halo exchange of the domain-decomposed step (b200nb_dd_*).  A rank exchanges halos over up to NB_DD_MAX_LINKS
 * LINKS; link k of every rank belongs to the same neighbour offset (1-D: the one +x neighbour; N-D half shell: the 13 / 4
 * lexicographically positive offsets, gmxapi_b200/domdec_nd.py), so "my receive link k" and "my source's send link k" are the
 * two ends of one connection and share the flag index k.  Each rank owns a WINDOW in its device memory,
 *   [flag_x[k] @ 32*k][flag_f[k] @ 512 + 32*k][err @ 1024] ... [recv_x: 3 floats per halo atom @ 2048][recv_f: 3 floats per send entry],
 * that its neighbours write directly over NVLink (CUDA IPC mapping, or the plain device pointer inside one process):
 * the owner of halo atoms stores their coordinates into the receiver's recv_x (at the link's halo offset) and raises the
 * receiver's flag_x[k] to the step number; the rank that computed forces on them stores those into the owner's recv_f (at the
 * link's send-entry offset) and raises the owner's flag_f[k].  Consumer kernels spin on the flags in their own memory (bounded;
 * a time-out raises err).  One exchange each way per step, no host involvement, no NCCL call on the per-step path (the reference
 * pushes with cudaMemcpyAsync + event handshakes, gpuhaloexchange_impl.cu:403-444, one pulse per dimension, domdec.cpp:260-460). */
#define NB_DD_MAX_LINKS 16
#define NB_DD_MAX_PEERS 16
#define NB_DD_FLAG_X(k) (32 * (k))
#define NB_DD_FLAG_F(k) (512 + 32 * (k))
#define NB_DD_ERR 1024
#define NB_DD_DATA 2048
/* what the step's kernels need to know about the links (kernel parameter, by value) */
struct DdLinksDev
{
    int    nlinks;
    int    send_off[NB_DD_MAX_LINKS + 1]; /* send entries of link k: [send_off[k], send_off[k+1]) */
    int    halo_off[NB_DD_MAX_LINKS + 1]; /* halo atoms (0-based after the home atoms) received over link k */
    float  shift[NB_DD_MAX_LINKS][3];     /* added to the coordinates sent over link k (dd_move_x's box shift, domdec.cpp:300-318) */
    int    fshift_index[NB_DD_MAX_LINKS]; /* receive link k: shift-force slot that also gets the forces on its halo atoms, or -1 */
    float* peer_recv_x[NB_DD_MAX_LINKS];  /* send link k: where its coordinates land in the destination's window */
    int*   peer_flag_x[NB_DD_MAX_LINKS];
    float* peer_recv_f[NB_DD_MAX_LINKS];  /* receive link k: where the forces on its halo atoms go in the source's window */
    int*   peer_flag_f[NB_DD_MAX_LINKS];
};
struct DdState
{
    unsigned char* window = nullptr;
    size_t         window_bytes = 0;
    int            max_halo = 0, max_send = 0;
    size_t         off_recv_x = NB_DD_DATA, off_recv_f = 0;
    unsigned char* peer[NB_DD_MAX_PEERS] = {}; /* opened neighbour windows */
    bool           peer_is_ipc[NB_DD_MAX_PEERS] = {};
    size_t         peer_off_recv_f[NB_DD_MAX_PEERS] = {};
    int            nhome = 0, nhalo = 0, nsend = 0; /* nsend = send entries over all links (an atom can go to several) */
    DdLinksDev     links{};
    int*           d_send_atom = nullptr; /* per send entry: the home atom */
    int*           d_send_link = nullptr; /* per send entry: its link */
    int*           d_ent_off = nullptr;   /* per home atom: CSR into d_ent_idx, the send entries that carry it (force return) */
    int*           d_ent_idx = nullptr;
    unsigned char* d_halo_link = nullptr; /* per halo atom: the link it came over */
    size_t         cap_send = 0, cap_home = 0, cap_halo = 0;
    /* device-side repartitioning (dd_partition.cu): the replicated global topology, the ga2la look-up, scratch */
    int            nglobal = 0;
    int *          d_gtype = nullptr, *d_geoff = nullptr, *d_geidx = nullptr, *d_g2l = nullptr;
    float*         d_gq = nullptr;
    int*           d_part_scratch = nullptr;
    size_t         cap_part_scratch = 0;
    int*           d_count = nullptr; /* 2 last-block counters */
    int *          h_err = nullptr, *d_host_err = nullptr; /* mapped host mirror of the window's time-out flag (b200nb_dd_status) */
    int*           d_seq = nullptr;   /* device-resident step counter: the value the flags carry (graph-replayable) */
    bool           have_plan = false;
    /* the halo chain (wait, halo x -> grid, non-local kernel, push f) runs on its own high-priority stream beside
     * the local kernel, the reference's local / non-local stream split (cuda/nbnxm_cuda_data_mgmt.cu:260-291) */
    int            prio_high = 0;
    cudaStream_t   stream_nl = nullptr, stream_px = nullptr; /* stream_px: the halo x push, a branch of its own */
    bool           push_inline = false; /* B200NB_DD_PUSH_INLINE=1: halo x push on the main stream (many ranks sharing one GPU) */
    cudaEvent_t    ev_start = nullptr, ev_begin = nullptr, ev_nl_done = nullptr, ev_px_done = nullptr;
};

/* perturbed (free-energy) pairs: fep.cu */
struct FepState
{
    int          natoms = 0, nri = 0, nrj = 0, npert = 0;
    int*           d_pert = nullptr;    /* the perturbed atoms, ascending */
    unsigned char* d_is_pert = nullptr; /* per atom */
    int *        d_typeA = nullptr, *d_typeB = nullptr;
    float *      d_qA = nullptr, *d_qB = nullptr;
    int *        d_iinr = nullptr, *d_shift = nullptr, *d_jindex = nullptr, *d_jjnr = nullptr; /* t_nblist, mdtypes/nblist.h:117-137 */
    signed char* d_excl = nullptr;
    double*      d_out = nullptr; /* Vc, Vv, dvdl_coul, dvdl_vdw */
    bool                in_step = false; /* launched by b200nb_step / b200nb_compute, inside the captured graph */
    b200nb_fep_params_t step_params{};
    /* capacities of the list arrays and the count / offset scratch of b200nb_fep_build_list: grown with head-room, kept across
     * search steps (no cudaMalloc / cudaFree in a steady-state rebuild) */
    size_t cap_nri = 0, cap_nrj = 0, cap_scratch = 0;
    int*   d_scratch = nullptr;
};

/* listed interactions on the device (bonded.cu): lists in atom order, 6 floats per parameter set */
struct BondedState
{
    int     natoms = 0;
    int     count[B200NB_BONDED_KINDS] = {};
    int*    d_iatoms[B200NB_BONDED_KINDS] = {};
    float*  d_params[B200NB_BONDED_KINDS] = {};
    size_t  cap_iatoms[B200NB_BONDED_KINDS] = {}, cap_params[B200NB_BONDED_KINDS] = {}; /* elements; kept across updates */
    double* d_energy = nullptr; /* per kind + Coulomb-14 */
    bool    in_step = false;    /* launched by b200nb_step / b200nb_compute, inside the captured graph */
    float   scale14 = 0.f;
    bool    have_pbc = false;   /* b200nb_bonded_set_pbc: the cell of the image search, when it is not the context's */
    float   box9[9] = {};
    int     npbcdim = 3;
};

/* a captured step: the launches of b200nb_step / b200nb_dd_step for one set of buffers */
struct StepGraph
{
    cudaGraphExec_t exec = nullptr;
    const float*    x = nullptr;
    float*          f = nullptr;
    int             flags = -1, nkernels = 0;
    long long       generation = -1;
    cudaStream_t    stream = nullptr;
};

struct b200nb_context
{
    int          device = 0;
    cudaStream_t stream = nullptr;
    bool         own_stream = true;
    std::string  err;
    long long    nlaunches = 0;

    bool              have_params = false;
    b200nb_params_t   hp{};
    NbParamsDev       dp{};
    std::vector<float> nbfp_host; /* (ntypes+1)^2 * 2 incl. filler */
    bool              comb_geom = false;
    int               max_tiles = 16;
    float*            d_nbfp = nullptr; /* float2 per type pair */
    float*            d_nbfp_comb = nullptr; /* LJ-PME: float2 per type, NBParamGpu::nbfp_comb */
    float*            d_ewald_tab = nullptr; /* float2 per table point */
    float*            d_kconst = nullptr; /* 12 floats: rc2, beta, beta2, FD4/b, FD3/b, FN6, FN5, FD2/b, FD1/b, FD0/b, 0, 0 (force.cu KConst) */

    int    natoms = 0;
    int*   d_type = nullptr;
    float* d_q = nullptr;
    int *  d_excl_off = nullptr, *d_excl_idx = nullptr;
    std::vector<int>   h_type;
    std::vector<float> h_q;
    double             sum_q2 = 0;

    float box[3] = { 0, 0, 0 };     /* diagonal of the box matrix */
    float box_off[3] = { 0, 0, 0 }; /* triclinic cell: box[YY][XX], box[ZZ][XX], box[ZZ][YY] (b200nb_set_box_triclinic) */
    int   pbc[3] = { 1, 1, 1 };
    float* d_shift_vec = nullptr; /* 45*3 */
    float  h_shift_vec[B200NB_SHIFTS * 3];

    GridDesc grid[2]{};
    bool     grid_uploaded = false; /* slot arrays come from b200nb_set_grid_atoms (reference-built grid), no atom-order view */
    int      ncol_total = 0, ncells_total = 0, npad = 0;
    size_t   cap_atoms = 0, cap_pad = 0, cap_cols = 0;
    float*   d_x = nullptr;          /* natoms*3, original order (staging) */
    float*   d_fout = nullptr;       /* natoms*3 staging for D2H */
    int*     d_col_of_atom = nullptr;
    int*     d_pos_in_col = nullptr;  /* rank of the atom inside its column (from the counting atomic) */
    int*     d_col_count = nullptr;
    int*     d_col_cell0 = nullptr;
    int*     d_col_fill = nullptr;
    int*     d_atom_index = nullptr; /* slot -> atom */
    int*     d_slot_of_atom = nullptr;
    float*   d_xq = nullptr;
    float*   d_lj = nullptr;
    int*     d_atype = nullptr;
    float*   d_bb = nullptr;     /* 6 floats per cluster */
    float*   d_cellz = nullptr;  /* 2 floats per cell */
    float4*  d_f = nullptr;      /* per slot */
    float*   d_fshift = nullptr; /* NB_OUT_COPIES x NB_FSHIFT_PITCH accumulators */
    double*  d_energy = nullptr; /* NB_OUT_COPIES x 2 accumulators */
    float*   d_fshift_sum = nullptr; /* 45*3: replicas summed by k_reduce_outputs */
    double*  d_energy_sum = nullptr; /* 2 */
    int*     d_scratch = nullptr; /* small ints: totals, flags, counters */
    long long* d_counter = nullptr;
    int*     d_hist = nullptr; /* 80 ints: histogram + cursors of the entry ordering */

    int*  d_cnt_tiles = nullptr; /* per i-cluster counts / offsets for the two-pass search */
    int*  d_cnt_entries = nullptr;
    size_t cap_clusters = 0;

    PairList outer[2], inner[2];
    bool     inner_is_outer = true;
    bool     inner_stale[2] = { false, false }; /* inner[] (pruned cluster pairs, introspection only) is behind the packed list */
    bool     have_list = false;
    bool     search_two_pass = false; /* B200NB_SEARCH_TWO_PASS=1: always count, scan, fill (A/B switch; the first search always does) */
    PackedList packed[2];
    DdState    dd;
    FepState   fep;
    BondedState bonded;
    StepGraph  graph[3];       /* [0] single-domain step, [1] decomposed step, [2] host step through the copy engines */
    int        host_dma = -1;  /* b200nb_compute: 1 = cudaMemcpyAsync staging, 0 = zero-copy kernels, -1 = not decided yet */
    long long  generation = 0; /* bumped whenever a list or halo plan is rebuilt: invalidates the captured graphs */
    bool       use_graphs = true;
    bool       use_pdl = true; /* launch the force kernel with programmatic stream serialization (see force.cu) */
    bool       capturing = false;
    std::vector<cudaGraphNode_t> nl_nodes; /* kernel nodes captured from the non-local stream (get an explicit priority) */
    int        dummy_slot = -1; /* first of the NB_DUMMY_SLOTS far-away filler slots: the tail of the slot arrays */
    size_t     dummy_cap = 0;   /* cap_pad / type count the dummies were written for */
    int        dummy_ntypes = 0;

    float* d_flush = nullptr;
    size_t flush_bytes = 0;

    unsigned char* h_pinned = nullptr; /* mapped pinned scratch of b200nb_compute */
    size_t         pinned_bytes = 0;
    const void *   map_x_host = nullptr, *map_f_host = nullptr; /* last host buffers b200nb_compute resolved */
    float *        map_x_dev = nullptr, *map_f_dev = nullptr;   /* their device-visible addresses (null: pageable) */
};

int nb_fail(b200nb_context* h, int code, const std::string& msg);
#define NB_CUDA(h, call)                                                                         \
    do                                                                                           \
    {                                                                                            \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return nb_fail(h, B200NB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

/* force.cu */
int nb_launch_force_kernel(b200nb_context* h, int locality, int flags);
/* fep.cu */
void nb_fep_free(b200nb_context* h);
int  nb_fep_enqueue_in_step(b200nb_context* h);
/* bonded.cu */
void nb_bonded_free(b200nb_context* h);
int  nb_bonded_enqueue_in_step(b200nb_context* h, int flags);


/* ---- device helpers shared by search / prune / pair extraction ---- */
#ifdef __CUDACC__
/* r^2 with the reference's operand roles and operation order: i-atom already shifted,
 * fma(dz,dz,fma(dx,dx,dy*dy)) -- see oracle/nbnxm_oracle.c header for the derivation. */
__device__ __forceinline__ float nb_rsq(float xi, float yi, float zi, float xj, float yj, float zj)
{
    float dx = __fsub_rn(xi, xj), dy = __fsub_rn(yi, yj), dz = __fsub_rn(zi, zj);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
#endif

#endif
