/* b200nb: context management, device gridding/sorting, device pair search with exclusion masks,
 * dynamic pruning, coordinate / force buffer ops, halo pack/unpack, introspection.
 * The force kernels live in force.cu.  Hand-written for sm_100a; no CPU fallback anywhere.
 *
 * Reference behaviour replaced (paths relative to /root/reference/src/gromacs):
 *   gridding      nbnxm/grid.cpp:103-262 (setDimensions), :1173-1268 (calcColumnIndices),
 *                 :1287-1445 (setCellIndices), :292-427 (sort_atoms), :1051-1164 (sortColumnsGpuGeometry),
 *                 :860-986 (fillCell), :454-660 (bounding boxes)
 *   search        nbnxm/pairlist.cpp:3076-3582 (nbnxn_make_pairlist_part), :1090-1244 (make_cluster_list_supersub),
 *                 :1874-1972 (setExclusionsForIEntry), :2077-2194 (split_sci_entry / closeIEntry)
 *   prune         nbnxm/cuda/nbnxm_cuda_kernel_pruneonly.cuh:104-277, kernels_reference/kernel_ref_prune.cpp:45-143
 *   buffer ops    nbnxm/cuda/nbnxm_buffer_ops_kernels.cuh:65-114, mdlib/gpuforcereduction_impl.cu:70-104
 *   halo          domdec/gpuhaloexchange_impl.cu:77-131
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <nvtx3/nvToolsExt.h> /* header-only: ranges cost a null-pointer check when no profiler is attached */

#include "b200nb_internal.h"

/* NVTX range around an API call, the role of the reference's wallcycle / NVTX markers around its nonbonded calls
 * (timing/wallcycle.cpp, utility/nvtx ranges in cuda builds): shows the search, prune and step calls on a profiler timeline */
struct NvtxRange
{
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

int nb_fail(b200nb_context* h, int code, const std::string& msg)
{
    if (h) h->err = msg;
    return code;
}

#ifndef B200NB_HOST_DMA_DEFAULT
#define B200NB_HOST_DMA_DEFAULT 0 /* b200nb_compute: 0 = the step kernels read / write pinned host buffers in place (measured faster) */
#endif

#define LAUNCH_CHECK(h)                 \
    do                                  \
    {                                   \
        (h)->nlaunches++;               \
        NB_CUDA(h, cudaGetLastError()); \
    } while (0)

template<typename T>
static int ensure(b200nb_context* h, T** p, size_t* cap, size_t need, double slack = 1.2)
{
    if (need <= *cap && *p) return 0;
    if (*p) NB_CUDA(h, cudaFree(*p));
    *p        = nullptr;
    size_t nc = (size_t)(need * slack) + 64;
    NB_CUDA(h, cudaMalloc((void**)p, nc * sizeof(T)));
    *cap = nc;
    return 0;
}

template<typename T>
static int alloc_exact(b200nb_context* h, T** p, size_t n)
{
    if (*p) NB_CUDA(h, cudaFree(*p));
    *p = nullptr;
    NB_CUDA(h, cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* lifetime                                                                                                */
/* ------------------------------------------------------------------------------------------------------ */
extern "C" int b200nb_create(b200nb_t** out, int device)
{
    if (!out) return B200NB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
    {
        return B200NB_ERR_CUDA; /* no CUDA device: the product path fails loudly, there is no fallback */
    }
    b200nb_context* h = new b200nb_context;
    h->device         = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        delete h;
        return B200NB_ERR_CUDA;
    }
    cudaMalloc((void**)&h->d_shift_vec, sizeof(float) * B200NB_SHIFTS * 3);
    cudaMalloc((void**)&h->d_fshift, sizeof(float) * NB_OUT_COPIES * NB_FSHIFT_PITCH);
    cudaMalloc((void**)&h->d_energy, sizeof(double) * NB_OUT_COPIES * 2);
    cudaMalloc((void**)&h->d_fshift_sum, sizeof(float) * NB_FSHIFT_PITCH);
    cudaMalloc((void**)&h->d_energy_sum, sizeof(double) * 2);
    cudaMalloc((void**)&h->d_scratch, sizeof(int) * 64);
    cudaMalloc((void**)&h->d_kconst, sizeof(float) * 12);
    cudaMalloc((void**)&h->d_counter, sizeof(long long) * 8);
    cudaMalloc((void**)&h->d_hist, sizeof(int) * 2 * NB_ORDER_BINS);
    cudaMemsetAsync(h->d_fshift, 0, sizeof(float) * NB_OUT_COPIES * NB_FSHIFT_PITCH, h->stream);
    cudaMemsetAsync(h->d_energy, 0, sizeof(double) * NB_OUT_COPIES * 2, h->stream);
    *out = h;
    return B200NB_OK;
}

static void free_list(PairList& l)
{
    cudaFree(l.entries);
    cudaFree(l.cj);
    cudaFree(l.mask);
    l = PairList();
}

extern "C" void b200nb_destroy(b200nb_t* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    nb_fep_free(h);
    nb_bonded_free(h);
    cudaStreamSynchronize(h->stream);
    void* ptrs[] = { h->d_ewald_tab, h->d_kconst, h->d_nbfp_comb, h->d_nbfp,       h->d_type,     h->d_q,        h->d_excl_off,     h->d_excl_idx,  h->d_shift_vec,
                     h->d_x,          h->d_fout,     h->d_col_of_atom, h->d_pos_in_col, h->d_col_count, h->d_col_cell0, h->d_col_fill,
                     h->d_atom_index, h->d_slot_of_atom, h->d_xq,   h->d_lj,           h->d_atype,     h->d_bb,
                     h->d_cellz,      h->d_f,        h->d_fshift,   h->d_energy,       h->d_scratch,   h->d_counter,
                     h->d_fshift_sum, h->d_energy_sum, h->d_hist,
                     h->d_cnt_tiles,  h->d_cnt_entries, h->d_flush };
    for (void* p : ptrs) cudaFree(p);
    for (int l = 0; l < 2; l++)
    {
        if (!h->inner_is_outer) free_list(h->inner[l]);
        free_list(h->outer[l]);
        cudaFree(h->packed[l].entries);
        cudaFree(h->packed[l].staged);
        cudaFree(h->packed[l].order_blk);
        cudaFree(h->packed[l].dest);
        cudaFree(h->packed[l].sizes);
        cudaFree(h->packed[l].ja);
        cudaFree(h->packed[l].mask);
    }
    for (int side = 0; side < NB_DD_MAX_PEERS; side++)
        if (h->dd.peer[side] && h->dd.peer_is_ipc[side]) cudaIpcCloseMemHandle(h->dd.peer[side]);
    for (auto& G : h->graph)
        if (G.exec) cudaGraphExecDestroy(G.exec);
    if (h->dd.stream_nl)
    {
        cudaStreamSynchronize(h->dd.stream_nl);
        cudaStreamDestroy(h->dd.stream_nl);
        if (h->dd.stream_px)
        {
            cudaStreamSynchronize(h->dd.stream_px);
            cudaStreamDestroy(h->dd.stream_px);
        }
        cudaEventDestroy(h->dd.ev_begin);
        cudaEventDestroy(h->dd.ev_start);
        if (h->dd.ev_px_done) cudaEventDestroy(h->dd.ev_px_done);
        cudaEventDestroy(h->dd.ev_nl_done);
    }
    cudaFree(h->dd.window);
    cudaFree(h->dd.d_count);
    cudaFree(h->dd.d_send_atom);
    cudaFree(h->dd.d_send_link);
    cudaFree(h->dd.d_ent_off);
    cudaFree(h->dd.d_ent_idx);
    cudaFree(h->dd.d_halo_link);
    if (h->dd.h_err) cudaFreeHost(h->dd.h_err);
    cudaFree(h->dd.d_gtype), cudaFree(h->dd.d_gq), cudaFree(h->dd.d_geoff), cudaFree(h->dd.d_geidx), cudaFree(h->dd.d_g2l), cudaFree(h->dd.d_part_scratch);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char* b200nb_last_error(const b200nb_t* h)
{
    return h ? h->err.c_str() : "null context";
}

extern "C" void* b200nb_stream(b200nb_t* h)
{
    return h ? (void*)h->stream : nullptr;
}

/* Nbnxm::gpu_init receives its streams from the caller's DeviceStreamManager (cuda/nbnxm_cuda_data_mgmt.cu:242-291);
 * here the caller may hand over a stream it owns. All later work of the context is issued on it. */
extern "C" int b200nb_set_stream(b200nb_t* h, void* cuda_stream)
{
    if (!h) return B200NB_ERR_ARG;
    cudaSetDevice(h->device);
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    if (cuda_stream)
    {
        h->stream     = (cudaStream_t)cuda_stream;
        h->own_stream = false;
    }
    else
    {
        NB_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    return 0;
}

extern "C" int b200nb_synchronize(b200nb_t* h)
{
    if (!h) return B200NB_ERR_ARG;
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

static int ensure_pinned(b200nb_context* h, size_t bytes)
{
    if (bytes <= h->pinned_bytes) return 0;
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    h->h_pinned = nullptr;
    NB_CUDA(h, cudaHostAlloc((void**)&h->h_pinned, bytes * 2, cudaHostAllocMapped));
    h->pinned_bytes = bytes * 2;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* parameters                                                                                              */
/* ------------------------------------------------------------------------------------------------------ */
extern "C" int b200nb_set_params(b200nb_t* h, const b200nb_params_t* p)
{
    if (!h || !p || p->ntypes < 1 || !p->nbfp_host || p->rc <= 0) return nb_fail(h, B200NB_ERR_ARG, "set_params: bad argument");
    if (p->rlist_outer < p->rc) return nb_fail(h, B200NB_ERR_ARG, "set_params: rlist_outer < rc");
    cudaSetDevice(h->device);
    h->hp           = *p;
    const int nt    = p->ntypes;
    const int ntf   = nt + 1; /* + filler type with zero parameters: atomdata.cpp:456,526-534 */
    h->nbfp_host.assign((size_t)ntf * ntf * 2, 0.0f);
    for (int i = 0; i < nt; i++)
        for (int j = 0; j < nt; j++)
        {
            h->nbfp_host[(i * ntf + j) * 2]     = p->nbfp_host[(i * nt + j) * 2];
            h->nbfp_host[(i * ntf + j) * 2 + 1] = p->nbfp_host[(i * nt + j) * 2 + 1];
        }
    h->hp.nbfp_host = nullptr;
    /* geometric combination rule detection, tolerance 1e-5 (atomdata.cpp:462-525, gmx_within_tol) */
    bool geom = true;
    for (int i = 0; i < nt && geom; i++)
        for (int j = 0; j < nt && geom; j++)
        {
            double c6 = p->nbfp_host[(i * nt + j) * 2], c12 = p->nbfp_host[(i * nt + j) * 2 + 1];
            double c6ii = p->nbfp_host[(i * nt + i) * 2], c6jj = p->nbfp_host[(j * nt + j) * 2];
            double c12ii = p->nbfp_host[(i * nt + i) * 2 + 1], c12jj = p->nbfp_host[(j * nt + j) * 2 + 1];
            auto   within = [](double a, double b) { return std::fabs(a - b) <= 1e-5 * 0.5 * (std::fabs(a) + std::fabs(b)); };
            geom          = within(c6 * c6, c6ii * c6jj) && within(c12 * c12, c12ii * c12jj);
        }
    h->comb_geom = (p->comb_rule == 1) || (p->comb_rule == 0 && geom);
    /* 0 = chosen from the system size in build_pairlist */
    h->max_tiles = p->max_tiles_per_entry > 0 ? std::min(p->max_tiles_per_entry, NB_MAX_ENTRY_TILES) : 16;
    if (alloc_exact(h, &h->d_nbfp, (size_t)ntf * ntf * 2)) return B200NB_ERR_CUDA;
    NB_CUDA(h, cudaMemcpyAsync(h->d_nbfp, h->nbfp_host.data(), sizeof(float) * ntf * ntf * 2, cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(h, cudaStreamSynchronize(h->stream));

    NbParamsDev& d  = h->dp;
    d.rc2           = p->rc * p->rc;
    d.rlist_outer2  = p->rlist_outer * p->rlist_outer;
    float rin       = (p->rlist_inner > 0 && p->rlist_inner < p->rlist_outer) ? p->rlist_inner : p->rlist_outer;
    d.rlist_inner2  = rin * rin;
    d.epsfac        = p->epsfac;
    d.k_rf          = (p->eeltype == B200NB_EEL_RF) ? p->k_rf : 0.0f;
    d.two_k_rf      = 2.0f * d.k_rf;
    d.c_rf          = p->c_rf;
    d.beta          = p->ewald_beta;
    d.beta2         = p->ewald_beta * p->ewald_beta;
    d.beta3         = d.beta2 * p->ewald_beta;
    d.sh_ewald      = p->sh_ewald;
    d.disp_cpot     = p->disp_cpot;
    d.rep_cpot      = p->rep_cpot;
    d.self_sub      = (p->eeltype == B200NB_EEL_EWALD) ? (float)(0.5 * p->ewald_beta * 1.12837916709551257390) : 0.5f * p->c_rf;
    d.self_q2       = (p->epsfac != 0.0f) ? d.self_sub / p->epsfac : 0.0f;
    {
        /* leading coefficients of pmeForceCorrection (simd/simd_math.h:1609-1650) and the loop-invariant scalars */
        /* the denominator coefficients FD4..FD0 are stored divided by beta, so that the kernel's 1/denominator already carries
         * the factor beta of beta * pmecorrF (reaction field: beta = 0, coefficients unused) */
        const float ib    = d.beta != 0.0f ? 1.0f / d.beta : 0.0f;
        const float kc[12] = { d.rc2, d.beta, d.beta2, 0.0011193462567257629232f * ib, 0.014866955030185295499f * ib,
                               -1.7357322914161492954e-8f, 1.4703624142580877519e-6f, 0.11583842382862377919f * ib,
                               0.50736591960530292870f * ib, 1.0f * ib, 0.0f, 0.0f };
        NB_CUDA(h, cudaMemcpy(h->d_kconst, kc, sizeof(kc), cudaMemcpyHostToDevice));
    }
    d.ntypes        = ntf;
    d.eeltype       = p->eeltype;
    d.vdw_modifier  = B200NB_VDW_POTSHIFT;
    d.rvdw2         = d.rc2;
    d.rvdw_switch   = 0.0f;
    d.disp_c2 = d.disp_c3 = d.rep_c2 = d.rep_c3 = 0.0f;
    d.sw_c3 = d.sw_c4 = d.sw_c5 = 0.0f;
    d.ljpme      = 0;
    d.lje_coeff2 = d.lje_coeff6_6 = d.sh_lj_ewald = 0.0f;
    d.ewald_tab  = nullptr; /* analytical Ewald correction until b200nb_set_ewald_table */
    d.tab_scale = d.tab_max = 0.0f;
    h->have_params  = true;
    h->have_list    = false;
    return 0;
}

extern "C" int b200nb_set_vdw(b200nb_t* h, const b200nb_vdw_t* v)
{
    if (!h || !v) return B200NB_ERR_ARG;
    if (!h->have_params) return nb_fail(h, B200NB_ERR_STATE, "set_vdw: set_params first");
    if (v->vdw_modifier < B200NB_VDW_POTSHIFT || v->vdw_modifier > B200NB_VDW_POTSWITCH)
        return nb_fail(h, B200NB_ERR_ARG, "set_vdw: unknown modifier");
    const float rvdw = v->rvdw > 0.0f ? v->rvdw : h->hp.rc;
    if (rvdw > h->hp.rc) return nb_fail(h, B200NB_ERR_ARG, "set_vdw: rvdw > rc (the list radius follows the Coulomb cut-off)");
    /* the reference's Verlet scheme allows rvdw != rcoulomb only with PME electrostatics (kerneldispatch.cpp:175-200:
     * twin-range variants exist for the Ewald kernels alone) */
    if (rvdw < h->hp.rc && h->dp.eeltype != B200NB_EEL_EWALD)
        return nb_fail(h, B200NB_ERR_ARG, "set_vdw: rvdw < rcoulomb needs Ewald electrostatics");
    if (v->vdw_modifier != B200NB_VDW_POTSHIFT && !(v->rvdw_switch >= 0.0f && v->rvdw_switch < rvdw))
        return nb_fail(h, B200NB_ERR_ARG, "set_vdw: rvdw_switch must lie in [0, rvdw)");
    if (v->ljpme_comb_rule < 0 || v->ljpme_comb_rule > 2) return nb_fail(h, B200NB_ERR_ARG, "set_vdw: unknown LJ-PME combination rule");
    if (v->ljpme_comb_rule && v->vdw_modifier != B200NB_VDW_POTSHIFT)
        return nb_fail(h, B200NB_ERR_ARG, "set_vdw: LJ-PME goes with the potential-shift modifier only");
    NbParamsDev& d = h->dp;
    d.ljpme        = v->ljpme_comb_rule;
    if (d.ljpme)
    {
        /* nbfp_comb as set_lj_parameter_data stores it (atomdata.cpp:291-322): geometric {sqrt(6 C6_ii), sqrt(12 C12_ii)},
         * Lorentz-Berthelot {0.5 (C12/C6)^(1/6), sqrt(C6^2 / C12)} (zero for types without LJ); the filler type gets zeros */
        const int          ntf = d.ntypes;
        std::vector<float> comb((size_t)ntf * 2, 0.0f);
        for (int i = 0; i < ntf; i++)
        {
            const float c6 = h->nbfp_host[((size_t)i * ntf + i) * 2], c12 = h->nbfp_host[((size_t)i * ntf + i) * 2 + 1];
            if (d.ljpme == 1)
            {
                comb[2 * i]     = std::sqrt(c6);
                comb[2 * i + 1] = std::sqrt(c12);
            }
            else if (c6 > 0 && c12 > 0)
            {
                comb[2 * i]     = 0.5f * std::pow(c12 / c6, 1.0f / 6.0f);
                comb[2 * i + 1] = std::sqrt(c6 * c6 / c12);
            }
        }
        if (alloc_exact(h, &h->d_nbfp_comb, (size_t)ntf * 2)) return B200NB_ERR_CUDA;
        NB_CUDA(h, cudaMemcpy(h->d_nbfp_comb, comb.data(), sizeof(float) * ntf * 2, cudaMemcpyHostToDevice));
        d.lje_coeff2   = v->ewaldcoeff_lj * v->ewaldcoeff_lj;
        d.lje_coeff6_6 = d.lje_coeff2 * d.lje_coeff2 * d.lje_coeff2 / 6.0f;
        d.sh_lj_ewald  = v->sh_lj_ewald;
    }
    d.vdw_modifier = v->vdw_modifier;
    d.rvdw2        = rvdw * rvdw;
    d.rvdw_switch  = v->rvdw_switch;
    d.disp_c2 = v->disp_c2, d.disp_c3 = v->disp_c3, d.rep_c2 = v->rep_c2, d.rep_c3 = v->rep_c3;
    d.sw_c3 = v->sw_c3, d.sw_c4 = v->sw_c4, d.sw_c5 = v->sw_c5;
    h->generation++; /* captured step graphs carry the kernel parameters by value */
    return 0;
}

/* init_ewald_coulomb_force_table (nbnxm_gpu_data_mgmt.cpp:71-83): the reference's tabulated Ewald force correction,
 * EwaldCorrectionTables::tableF with its scale (tables/forcetable.cpp generateEwaldCorrectionTables).  With a table set, the plain
 * Ewald kernels interpolate it (EL_EWALD_TAB) instead of evaluating the analytical correction; n = 0 returns to the analytical form. */
extern "C" int b200nb_set_ewald_table(b200nb_t* h, const float* table_f_host, int n, float scale)
{
    if (!h || n < 0 || (n > 0 && (!table_f_host || n < 2 || !(scale > 0)))) return nb_fail(h, B200NB_ERR_ARG, "set_ewald_table: bad argument");
    if (!h->have_params) return nb_fail(h, B200NB_ERR_STATE, "set_ewald_table: set_params first");
    cudaSetDevice(h->device);
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (n == 0)
    {
        h->dp.ewald_tab = nullptr;
        h->generation++;
        return 0;
    }
    if (h->dp.eeltype != B200NB_EEL_EWALD) return nb_fail(h, B200NB_ERR_ARG, "set_ewald_table: the interaction is not Ewald");
    std::vector<float> t(2 * (size_t)n);
    for (int i = 0; i < n; i++)
    {
        t[2 * i]     = table_f_host[i];
        t[2 * i + 1] = i + 1 < n ? table_f_host[i + 1] - table_f_host[i] : 0.0f;
    }
    if (alloc_exact(h, &h->d_ewald_tab, t.size())) return B200NB_ERR_CUDA;
    NB_CUDA(h, cudaMemcpy(h->d_ewald_tab, t.data(), sizeof(float) * t.size(), cudaMemcpyHostToDevice));
    h->dp.ewald_tab = reinterpret_cast<const float2*>(h->d_ewald_tab);
    h->dp.tab_scale = scale;
    h->dp.tab_max   = (float)(n - 2) + 0.999f; /* the last interval: lanes beyond the table (outside the cut-off) read inside it */
    if (h->hp.rc * scale > (float)(n - 1)) return nb_fail(h, B200NB_ERR_ARG, "set_ewald_table: the table ends before the cut-off");
    h->generation++; /* captured step graphs carry the kernel parameters by value */
    return 0;
}

extern "C" int b200nb_set_atoms(b200nb_t* h, int natoms, const int* type_host, const float* q_host,
                                const int* excl_off_host, const int* excl_idx_host)
{
    if (!h || natoms < 1 || !type_host || !q_host) return nb_fail(h, B200NB_ERR_ARG, "set_atoms: bad argument");
    if (!h->have_params) return nb_fail(h, B200NB_ERR_STATE, "set_atoms: set_params first");
    cudaSetDevice(h->device);
    for (int a = 0; a < natoms; a++)
        if (type_host[a] < 0 || type_host[a] >= h->hp.ntypes) return nb_fail(h, B200NB_ERR_ARG, "set_atoms: atom type out of range");
    h->natoms = natoms;
    h->h_type.assign(type_host, type_host + natoms);
    h->h_q.assign(q_host, q_host + natoms);
    if (alloc_exact(h, &h->d_type, (size_t)natoms) || alloc_exact(h, &h->d_q, (size_t)natoms)) return B200NB_ERR_CUDA;
    NB_CUDA(h, cudaMemcpy(h->d_type, type_host, sizeof(int) * natoms, cudaMemcpyHostToDevice));
    NB_CUDA(h, cudaMemcpy(h->d_q, q_host, sizeof(float) * natoms, cudaMemcpyHostToDevice));
    std::vector<int> off(natoms + 1, 0);
    const int*       idx  = nullptr;
    int              nidx = 0;
    if (excl_off_host)
    {
        if (excl_off_host[0] != 0) return nb_fail(h, B200NB_ERR_ARG, "set_atoms: excl_off[0] != 0");
        off.assign(excl_off_host, excl_off_host + natoms + 1);
        nidx = off[natoms];
        idx  = excl_idx_host;
        for (int k = 0; k < nidx; k++)
            if (idx[k] < 0 || idx[k] >= natoms) return nb_fail(h, B200NB_ERR_ARG, "set_atoms: exclusion index out of range");
    }
    if (alloc_exact(h, &h->d_excl_off, (size_t)natoms + 1) || alloc_exact(h, &h->d_excl_idx, (size_t)nidx)) return B200NB_ERR_CUDA;
    NB_CUDA(h, cudaMemcpy(h->d_excl_off, off.data(), sizeof(int) * (natoms + 1), cudaMemcpyHostToDevice));
    if (nidx) NB_CUDA(h, cudaMemcpy(h->d_excl_idx, idx, sizeof(int) * nidx, cudaMemcpyHostToDevice));
    if (alloc_exact(h, &h->d_x, (size_t)natoms * 3) || alloc_exact(h, &h->d_fout, (size_t)natoms * 3)
        || alloc_exact(h, &h->d_col_of_atom, (size_t)natoms) || alloc_exact(h, &h->d_slot_of_atom, (size_t)natoms)
        || alloc_exact(h, &h->d_pos_in_col, (size_t)natoms))
        return B200NB_ERR_CUDA;
    h->grid[0].valid = h->grid[1].valid = 0;
    h->have_list                        = false;
    return 0;
}

/* b200nb_set_atoms for arrays that are BUILT ON THE DEVICE (b200nb_dd_set_local_atoms, dd_partition.cu): the same allocations,
 * the caller's kernels fill d_type / d_q / d_excl_off / d_excl_idx.  nexcl < 0: everything but the exclusion indices (their count
 * is not known yet); nexcl >= 0: the exclusion indices. */
int nb_install_atoms_dev(b200nb_context* h, int natoms, int nexcl)
{
    if (nexcl >= 0) return alloc_exact(h, &h->d_excl_idx, (size_t)nexcl) ? B200NB_ERR_CUDA : 0;
    if (!h->have_params) return nb_fail(h, B200NB_ERR_STATE, "set_atoms: set_params first");
    h->natoms = natoms;
    h->h_type.clear(); /* no host copies on this path */
    h->h_q.clear();
    if (alloc_exact(h, &h->d_type, (size_t)natoms) || alloc_exact(h, &h->d_q, (size_t)natoms) || alloc_exact(h, &h->d_excl_off, (size_t)natoms + 1)
        || alloc_exact(h, &h->d_x, (size_t)natoms * 3) || alloc_exact(h, &h->d_fout, (size_t)natoms * 3) || alloc_exact(h, &h->d_col_of_atom, (size_t)natoms)
        || alloc_exact(h, &h->d_slot_of_atom, (size_t)natoms) || alloc_exact(h, &h->d_pos_in_col, (size_t)natoms))
        return B200NB_ERR_CUDA;
    h->grid[0].valid = h->grid[1].valid = 0;
    h->have_list                        = false;
    return 0;
}

extern "C" int b200nb_set_box(b200nb_t* h, const float box[3], const int pbc_dims[3])
{
    if (!h || !box) return nb_fail(h, B200NB_ERR_ARG, "set_box: bad argument");
    cudaSetDevice(h->device);
    for (int d = 0; d < 3; d++)
    {
        h->box[d]     = box[d];
        h->box_off[d] = 0.f;
        h->pbc[d]     = pbc_dims ? pbc_dims[d] : 1;
    }
    /* pbcutil/pbc.cpp:1187-1202 calc_shifts, rectangular box */
    int n = 0;
    for (int m = -1; m <= 1; m++)
        for (int l = -1; l <= 1; l++)
            for (int k = -2; k <= 2; k++, n++)
            {
                h->h_shift_vec[3 * n]     = k * box[0];
                h->h_shift_vec[3 * n + 1] = l * box[1];
                h->h_shift_vec[3 * n + 2] = m * box[2];
            }
    NB_CUDA(h, cudaMemcpy(h->d_shift_vec, h->h_shift_vec, sizeof(h->h_shift_vec), cudaMemcpyHostToDevice));
    return 0;
}

/* Triclinic cell: the lower-triangular GROMACS box matrix, rows a = (a_x, 0, 0), b = (b_x, b_y, 0), c = (c_x, c_y, c_z), within the
 * reference's limits (pbcutil/pbc.cpp check_box: |b_x|, |c_x| <= a_x / 2, |c_y| <= b_y / 2, margin 1.001).  Atoms are expected where
 * put_atoms_in_box leaves them: in the brick [0, a_x) x [0, b_y) x [0, c_z), which is what b200nb_put_on_grid covers; the cell's
 * shape enters through the 45 shift vectors k a + l b + m c (calc_shifts, pbc.cpp:1187-1202) and the x-shift range of 2
 * (nbnxm/pairlist.cpp:3181-3188).  All three dimensions periodic; not combined with domain decomposition. */
extern "C" int b200nb_set_box_triclinic(b200nb_t* h, const float box9[9])
{
    if (!h || !box9) return nb_fail(h, B200NB_ERR_ARG, "set_box_triclinic: bad argument");
    if (box9[1] != 0.f || box9[2] != 0.f || box9[5] != 0.f) return nb_fail(h, B200NB_ERR_ARG, "set_box_triclinic: the box matrix must be lower triangular");
    if (!(box9[0] > 0.f && box9[4] > 0.f && box9[8] > 0.f)) return nb_fail(h, B200NB_ERR_ARG, "set_box_triclinic: non-positive diagonal");
    const float margin = 1.001f;
    if (fabsf(box9[3]) > 0.5f * margin * box9[0] || fabsf(box9[6]) > 0.5f * margin * box9[0] || fabsf(box9[7]) > 0.5f * margin * box9[4])
        return nb_fail(h, B200NB_ERR_ARG, "set_box_triclinic: off-diagonal elements exceed half the diagonal (pbcutil/pbc.cpp check_box)");
    cudaSetDevice(h->device);
    h->box[0] = box9[0], h->box[1] = box9[4], h->box[2] = box9[8];
    h->box_off[0] = box9[3], h->box_off[1] = box9[6], h->box_off[2] = box9[7];
    h->pbc[0] = h->pbc[1] = h->pbc[2] = 1;
    int n = 0;
    for (int m = -1; m <= 1; m++)
        for (int l = -1; l <= 1; l++)
            for (int k = -2; k <= 2; k++, n++)
                for (int d = 0; d < 3; d++) h->h_shift_vec[3 * n + d] = k * box9[d] + l * box9[3 + d] + m * box9[6 + d];
    NB_CUDA(h, cudaMemcpy(h->d_shift_vec, h->h_shift_vec, sizeof(h->h_shift_vec), cudaMemcpyHostToDevice));
    return 0;
}

/* gpu_upload_shiftvec (cuda/nbnxm_cuda_data_mgmt.cu:283-295): the 45 shift vectors as the caller computed them (nbat->shift_vec =
 * calc_shifts(box)), for callers that grid and search themselves (b200nb_upload_pairlist): any box shape the reference's
 * list was built for.  b200nb_set_box remains the call for the device search (rectangular boxes). */
extern "C" int b200nb_set_shift_vec(b200nb_t* h, const float* shift_vec_host)
{
    if (!h || !shift_vec_host) return nb_fail(h, B200NB_ERR_ARG, "set_shift_vec: bad argument");
    cudaSetDevice(h->device);
    memcpy(h->h_shift_vec, shift_vec_host, sizeof(h->h_shift_vec));
    NB_CUDA(h, cudaMemcpyAsync(h->d_shift_vec, h->h_shift_vec, sizeof(h->h_shift_vec), cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* gridding kernels                                                                                        */
/* ------------------------------------------------------------------------------------------------------ */

/* grid.cpp:1173-1268 calcColumnIndices: cx = int((x - x0) * invCell), clamped.  The atom's rank inside its column comes out of
 * the same atomic that counts the column (the scatter below then needs none); lanes of a warp that hit the same column -- the
 * normal case, atoms arrive molecule by molecule -- share ONE atomic (a million atoms on a few hundred column counters
 * otherwise serialise in L2: 228 us + 287 us at 1 M atoms, profiles/r2/i_launches_water1M.txt).  A non-finite coordinate
 * raises *lost (the reference dies in sort_atoms with "Lost particles while sorting", grid.cpp:425). */
__global__ void k_column_index(const float* __restrict__ x, GridDesc g, int* __restrict__ col_of_atom, int* __restrict__ pos_in_col,
                               int* __restrict__ col_count, int* __restrict__ lost)
{
    const int  a      = g.atom_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = a < g.atom_end;
    int        c      = -1;
    if (active)
    {
        const float px = x[3 * a], py = x[3 * a + 1], pz = x[3 * a + 2];
        if (!(isfinite(px) && isfinite(py) && isfinite(pz))) atomicExch(lost, 1);
        int cx = (int)((px - g.lower[0]) * g.inv_cell[0]);
        int cy = (int)((py - g.lower[1]) * g.inv_cell[1]);
        cx     = max(0, min(cx, g.ncx - 1));
        cy     = max(0, min(cy, g.ncy - 1));
        c      = cx * g.ncy + cy;
    }
    const unsigned lane  = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int      lead  = __ffs(peers) - 1;
    int            base  = 0;
    if (active && (int)lane == lead) base = atomicAdd(&col_count[g.col0 + c], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, lead);
    if (active)
    {
        col_of_atom[a] = c;
        pos_in_col[a]  = base + __popc(peers & ((1u << lane) - 1u));
    }
}

/* grid.cpp:1287-1445 setCellIndices: cells per column = ceil(n/64), prefix sum -> cxy_ind_.
 * totals[0] = number of cells of this grid, totals[1] = max atoms in a column. */
__global__ void k_column_scan(GridDesc g, const int* __restrict__ col_count, int* __restrict__ col_cell0, int* __restrict__ totals)
{
    __shared__ int s_part[1024];
    __shared__ int s_max[1024];
    const int      t   = threadIdx.x;
    const int      per = (g.ncol + blockDim.x - 1) / blockDim.x;
    int            sum = 0, mx = 0;
    for (int k = 0; k < per; k++)
    {
        int c = t * per + k;
        if (c < g.ncol)
        {
            int n = col_count[g.col0 + c];
            sum += (n + NB_CELL - 1) / NB_CELL;
            mx = max(mx, n);
        }
    }
    s_part[t] = sum;
    s_max[t]  = mx;
    __syncthreads();
    if (t == 0)
    {
        int run = 0, m = 0;
        for (int i = 0; i < (int)blockDim.x; i++)
        {
            int v     = s_part[i];
            s_part[i] = run;
            run += v;
            m = max(m, s_max[i]);
        }
        totals[0] = run;
        totals[1] = m;
    }
    __syncthreads();
    int run = s_part[t] + g.cell0;
    for (int k = 0; k < per; k++)
    {
        int c = t * per + k;
        if (c < g.ncol)
        {
            col_cell0[g.col0 + c] = run;
            run += (col_count[g.col0 + c] + NB_CELL - 1) / NB_CELL;
            if (c == g.ncol - 1) col_cell0[g.col0 + g.ncol] = run;
        }
    }
}

__global__ void k_column_scatter(GridDesc g, const int* __restrict__ col_of_atom, const int* __restrict__ pos_in_col,
                                 const int* __restrict__ col_cell0, int* __restrict__ atom_index)
{
    int a = g.atom_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= g.atom_end) return;
    int c   = g.col0 + col_of_atom[a];
    atom_index[col_cell0[c] * NB_CELL + pos_in_col[a]] = a;
}

__device__ __forceinline__ uint32_t orderable(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

/* In-shared-memory bitonic sort of independent aligned segments of length seg (power of two), ascending. */
__device__ void bitonic_segments(uint64_t* s, int P, int seg)
{
    for (int k = 2; k <= seg; k <<= 1)
    {
        for (int j = k >> 1; j > 0; j >>= 1)
        {
            for (int i = threadIdx.x; i < P; i += blockDim.x)
            {
                int ixj = i ^ j;
                if (ixj > i)
                {
                    bool     asc = (k == seg) ? true : ((i & k) == 0);
                    uint64_t a = s[i], b = s[ixj];
                    if ((a > b) == asc)
                    {
                        s[i]   = b;
                        s[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

/* One CTA per grid column: the exact (coordinate, atom index) sorts of sort_atoms (grid.cpp:292-427; its
 * pigeonhole+insertion procedure yields exactly that order) in the sequence of sortColumnsGpuGeometry
 * (grid.cpp:1051-1164): whole column on z; 32-atom slabs on y, direction alternating with the slab
 * parity; 16-atom halves on x, backwards on the odd half; then fills the device atom data in the
 * pair-interleaved cluster layout (fillCell, grid.cpp:860-986) and the cluster bounding boxes. */
__global__ void k_column_sort(GridDesc g, const float* __restrict__ x, const float* __restrict__ q, const int* __restrict__ type,
                              const float* __restrict__ nbfp, int ntypes, const int* __restrict__ col_count,
                              const int* __restrict__ col_cell0, int* __restrict__ atom_index, int* __restrict__ slot_of_atom,
                              float* __restrict__ xq, float* __restrict__ lj, int* __restrict__ atype, float* __restrict__ bb,
                              float* __restrict__ cellz)
{
    extern __shared__ uint64_t s_key[];
    const int c     = g.col0 + blockIdx.x;
    const int n     = col_count[c];
    const int cell0 = col_cell0[c];
    const int ncz   = col_cell0[c + 1] - cell0;
    if (ncz == 0) return;
    const int base = cell0 * NB_CELL;
    const int npos = ncz * NB_CELL;
    int       P    = 64;
    while (P < npos) P <<= 1;
    const uint64_t FILL = ~0ull;

    /* z sort of the whole column */
    for (int i = threadIdx.x; i < P; i += blockDim.x)
    {
        uint64_t k = FILL;
        if (i < n)
        {
            int a = atom_index[base + i];
            k     = ((uint64_t)orderable(x[3 * a + 2]) << 32) | (uint32_t)a;
        }
        s_key[i] = k;
    }
    __syncthreads();
    bitonic_segments(s_key, P, P);
    /* y sort within 32-atom slabs, backwards on odd slabs (complemented keys sort descending) */
    for (int i = threadIdx.x; i < P; i += blockDim.x)
    {
        uint64_t k = s_key[i];
        if (k != FILL)
        {
            int a = (int)(uint32_t)k;
            k     = ((uint64_t)orderable(x[3 * a + 1]) << 32) | (uint32_t)a;
            if ((i >> 5) & 1) k = ~k;
        }
        s_key[i] = k;
    }
    __syncthreads();
    bitonic_segments(s_key, P, 32);
    /* x sort within 16-atom halves, backwards on the odd half */
    for (int i = threadIdx.x; i < P; i += blockDim.x)
    {
        uint64_t k = s_key[i];
        if (k != FILL)
        {
            int a = ((i >> 5) & 1) ? (int)(uint32_t)(~k) : (int)(uint32_t)k;
            k     = ((uint64_t)orderable(x[3 * a]) << 32) | (uint32_t)a;
            if ((i >> 4) & 1) k = ~k;
        }
        s_key[i] = k;
    }
    __syncthreads();
    bitonic_segments(s_key, P, 16);

    /* fill atom data */
    for (int i = threadIdx.x; i < npos; i += blockDim.x)
    {
        uint64_t  k    = s_key[i];
        const int slot = base + i;
        float4    v;
        float2    l;
        int       t;
        if (k != FILL)
        {
            int a = ((i >> 4) & 1) ? (int)(uint32_t)(~k) : (int)(uint32_t)k;
            atom_index[slot] = a;
            slot_of_atom[a]  = slot;
            v                = make_float4(x[3 * a], x[3 * a + 1], x[3 * a + 2], q[a]);
            t                = type[a];
            /* sqrt(6 C6_ii), sqrt(12 C12_ii): nbfp_comb for the geometric rule (atomdata.cpp:253-330) */
            l = make_float2(sqrtf(nbfp[(t * ntypes + t) * 2]), sqrtf(nbfp[(t * ntypes + t) * 2 + 1]));
        }
        else
        {
            /* filler: no charge, zero-LJ filler type, parked far away at a unique position so that no two fillers
             * are ever at zero distance of each other (reference: -1e6, atomdata.cpp:146) */
            atom_index[slot] = -1;
            v                = make_float4(-1.0e6f - 8.0f * (float)(slot & 0xfffff), -1.0e6f - 64.0f * (float)(slot >> 20), -1.0e6f, 0.0f);
            t                = ntypes - 1;
            l                = make_float2(0.0f, 0.0f);
        }
        reinterpret_cast<float4*>(xq)[slot] = v;
        reinterpret_cast<float2*>(lj)[slot] = l;
        atype[slot]                         = t;
    }
    __syncthreads();
    /* bounding boxes of real atoms per cluster; bb[cl*6+0] > bb[cl*6+3] marks an all-filler cluster */
    for (int cl = threadIdx.x; cl < ncz * 8; cl += blockDim.x)
    {
        float lo[3] = { 3.0e38f, 3.0e38f, 3.0e38f }, hi[3] = { -3.0e38f, -3.0e38f, -3.0e38f };
        for (int kk = 0; kk < 8; kk++)
        {
            uint64_t k = s_key[cl * 8 + kk];
            if (k == FILL) continue;
            int i = cl * 8 + kk;
            int a = ((i >> 4) & 1) ? (int)(uint32_t)(~k) : (int)(uint32_t)k;
            for (int d = 0; d < 3; d++)
            {
                float v = x[3 * a + d];
                lo[d]   = fminf(lo[d], v);
                hi[d]   = fmaxf(hi[d], v);
            }
        }
        float* b = bb + (size_t)(cell0 * 8 + cl) * 6;
        for (int d = 0; d < 3; d++)
        {
            b[d]     = lo[d];
            b[3 + d] = hi[d];
        }
    }
    __syncthreads();
    /* z range per cell (bbcz_, grid.cpp:1100-1102) */
    for (int cz = threadIdx.x; cz < ncz; cz += blockDim.x)
    {
        float lo = 3.0e38f, hi = -3.0e38f;
        for (int cl = 0; cl < 8; cl++)
        {
            const float* b = bb + (size_t)((cell0 + cz) * 8 + cl) * 6;
            lo             = fminf(lo, b[2]);
            hi             = fmaxf(hi, b[5]);
        }
        cellz[2 * (cell0 + cz)]     = lo;
        cellz[2 * (cell0 + cz) + 1] = hi;
    }
}

/* grid.cpp:103-262 Grid::setDimensions, GPU geometry */
static void grid_dimensions(GridDesc& g, int natoms, const float lower[3], const float upper[3], float density, int ddZone)
{
    float size[3];
    for (int d = 0; d < 3; d++)
    {
        g.lower[d] = lower[d];
        g.upper[d] = upper[d];
        size[d]    = upper[d] - lower[d];
    }
    if (natoms > NB_CELL)
    {
        float tlen   = cbrtf((float)NB_CL / density);
        float tlen_x = tlen * 2, tlen_y = tlen * 2;
        g.ncx = std::max(1, (int)(size[0] / tlen_x));
        g.ncy = std::max(1, (int)(size[1] / tlen_y));
    }
    else
    {
        g.ncx = g.ncy = 1;
    }
    for (int d = 0; d < 2; d++)
    {
        g.cell[d]     = size[d] / (d == 0 ? g.ncx : g.ncy);
        g.inv_cell[d] = 1 / g.cell[d];
    }
    if (ddZone > 0)
    {
        g.ncx++; /* grid.cpp:199-209: extra row for atoms beyond the cut-off range */
        g.ncy++;
    }
    g.ncol = g.ncx * g.ncy;
}

extern "C" int b200nb_put_on_grid(b200nb_t* h, int gi, const float lower[3], const float upper[3], int atom_begin, int atom_end,
                                  float density, const float* x, int x_on_device)
{
    NvtxRange nvtx_("b200nb_put_on_grid");
    if (!h || gi < 0 || gi > 1 || !lower || !upper || !x) return nb_fail(h, B200NB_ERR_ARG, "put_on_grid: bad argument");
    if (!h->natoms) return nb_fail(h, B200NB_ERR_STATE, "put_on_grid: set_atoms first");
    if (atom_begin < 0 || atom_end > h->natoms || atom_begin > atom_end) return nb_fail(h, B200NB_ERR_ARG, "put_on_grid: bad atom range");
    if (gi == 1 && !h->grid[0].valid) return nb_fail(h, B200NB_ERR_STATE, "put_on_grid: grid 0 must be set before grid 1");
    cudaSetDevice(h->device);
    if (h->grid_uploaded)
    {
        /* back from the reference-built grid (b200nb_set_grid_atoms): its slot arrays lack the search data, reallocate all */
        h->grid_uploaded = false;
        h->cap_pad       = 0;
    }
    const int n = atom_end - atom_begin;
    for (int d = 0; d < 3; d++)
        if (!(upper[d] > lower[d])) return nb_fail(h, B200NB_ERR_ARG, "put_on_grid: upper <= lower");
    GridDesc& g = h->grid[gi];
    if (gi == 0)
    {
        if (density <= 0) density = (float)n / ((upper[0] - lower[0]) * (upper[1] - lower[1]) * (upper[2] - lower[2]));
        h->grid[1].valid = 0;
    }
    else if (density <= 0)
    {
        const GridDesc& g0 = h->grid[0];
        density = (float)(g0.atom_end - g0.atom_begin)
                  / ((g0.upper[0] - g0.lower[0]) * (g0.upper[1] - g0.lower[1]) * (g0.upper[2] - g0.lower[2]));
    }
    grid_dimensions(g, n, lower, upper, density, gi);
    g.atom_begin = atom_begin;
    g.atom_end   = atom_end;
    g.col0       = (gi == 0) ? 0 : h->grid[0].ncol + 1;
    g.cell0      = (gi == 0) ? 0 : h->grid[0].ncells;
    g.valid      = 0;

    /* upload coordinates of this atom range */
    if (n > 0)
    {
        if (x_on_device)
            NB_CUDA(h, cudaMemcpyAsync(h->d_x + 3 * (size_t)atom_begin, x + 3 * (size_t)atom_begin, sizeof(float) * 3 * n, cudaMemcpyDeviceToDevice, h->stream));
        else
            NB_CUDA(h, cudaMemcpyAsync(h->d_x + 3 * (size_t)atom_begin, x + 3 * (size_t)atom_begin, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, h->stream));
    }
    /* column arrays */
    size_t ncols_need = (size_t)g.col0 + g.ncol + 2;
    if (ncols_need > h->cap_cols)
    {
        /* preserve grid 0's columns when growing for grid 1 */
        int *nc = nullptr, *n0 = nullptr, *nf = nullptr;
        size_t cap = ncols_need * 2 + 64;
        NB_CUDA(h, cudaMalloc((void**)&nc, cap * sizeof(int)));
        NB_CUDA(h, cudaMalloc((void**)&n0, cap * sizeof(int)));
        NB_CUDA(h, cudaMalloc((void**)&nf, cap * sizeof(int)));
        if (h->d_col_count && g.col0 > 0)
        {
            NB_CUDA(h, cudaMemcpyAsync(nc, h->d_col_count, sizeof(int) * g.col0, cudaMemcpyDeviceToDevice, h->stream));
            NB_CUDA(h, cudaMemcpyAsync(n0, h->d_col_cell0, sizeof(int) * g.col0, cudaMemcpyDeviceToDevice, h->stream));
            NB_CUDA(h, cudaStreamSynchronize(h->stream));
        }
        cudaFree(h->d_col_count);
        cudaFree(h->d_col_cell0);
        cudaFree(h->d_col_fill);
        h->d_col_count = nc;
        h->d_col_cell0 = n0;
        h->d_col_fill  = nf;
        h->cap_cols    = cap;
    }
    NB_CUDA(h, cudaMemsetAsync(h->d_col_count + g.col0, 0, sizeof(int) * (g.ncol + 1), h->stream));
    NB_CUDA(h, cudaMemsetAsync(h->d_scratch + 2, 0, sizeof(int), h->stream)); /* the lost-atom flag */
    int totals[3] = { 0, 0, 0 };
    if (n > 0)
    {
        k_column_index<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_x, g, h->d_col_of_atom, h->d_pos_in_col, h->d_col_count, h->d_scratch + 2);
        LAUNCH_CHECK(h);
    }
    k_column_scan<<<1, 1024, 0, h->stream>>>(g, h->d_col_count, h->d_col_cell0, h->d_scratch);
    LAUNCH_CHECK(h);
    NB_CUDA(h, cudaMemcpyAsync(totals, h->d_scratch, sizeof(int) * 3, cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (totals[2]) return nb_fail(h, B200NB_ERR_LOSTATOMS, "put_on_grid: non-finite coordinates (lost particles, grid.cpp:425)");
    g.ncells = totals[0];
    if (totals[1] > 8192) return nb_fail(h, B200NB_ERR_CAPACITY, "put_on_grid: more than 8192 atoms in one grid column");
    const int ncells_total = g.cell0 + g.ncells;
    const size_t npad      = (size_t)ncells_total * NB_CELL;
    if (npad > h->cap_pad)
    {
        /* grow all slot-indexed arrays, preserving grid 0's part when adding grid 1 */
        size_t cap  = (size_t)(npad * 1.15) + 1024;
        cap         = (cap + 63) / 64 * 64;
        size_t keep = (gi == 1) ? (size_t)g.cell0 * NB_CELL : 0;
        auto grow   = [&](void** p, size_t elem) -> int {
            void* np = nullptr;
            if (cudaMalloc(&np, cap * elem) != cudaSuccess) return 1;
            if (*p && keep) cudaMemcpy(np, *p, keep * elem, cudaMemcpyDeviceToDevice);
            cudaFree(*p);
            *p = np;
            return 0;
        };
        int bad = 0;
        bad |= grow((void**)&h->d_atom_index, sizeof(int));
        bad |= grow((void**)&h->d_xq, sizeof(float) * 4);
        bad |= grow((void**)&h->d_lj, sizeof(float) * 2);
        bad |= grow((void**)&h->d_atype, sizeof(int));
        bad |= grow((void**)&h->d_f, sizeof(float4));
        {
            /* per-cluster / per-cell arrays sized from the slot capacity */
            void* np = nullptr;
            if (cudaMalloc(&np, cap / 8 * 6 * sizeof(float)) != cudaSuccess) bad = 1;
            else
            {
                if (h->d_bb && keep) cudaMemcpy(np, h->d_bb, keep / 8 * 6 * sizeof(float), cudaMemcpyDeviceToDevice);
                cudaFree(h->d_bb);
                h->d_bb = (float*)np;
            }
            np = nullptr;
            if (cudaMalloc(&np, cap / 64 * 2 * sizeof(float)) != cudaSuccess) bad = 1;
            else
            {
                if (h->d_cellz && keep) cudaMemcpy(np, h->d_cellz, keep / 64 * 2 * sizeof(float), cudaMemcpyDeviceToDevice);
                cudaFree(h->d_cellz);
                h->d_cellz = (float*)np;
            }
        }
        if (bad) return nb_fail(h, B200NB_ERR_CUDA, "put_on_grid: device allocation failed");
        h->cap_pad = cap;
        NB_CUDA(h, cudaMemsetAsync(h->d_f, 0, cap * sizeof(float4), h->stream));
    }
    if (n > 0)
    {
        k_column_scatter<<<(n + 255) / 256, 256, 0, h->stream>>>(g, h->d_col_of_atom, h->d_pos_in_col, h->d_col_cell0, h->d_atom_index);
        LAUNCH_CHECK(h);
    }
    {
        int    P    = 64;
        int    need = ((totals[1] + NB_CELL - 1) / NB_CELL) * NB_CELL;
        while (P < need) P <<= 1;
        size_t smem = (size_t)P * sizeof(uint64_t);
        NB_CUDA(h, cudaFuncSetAttribute(k_column_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        k_column_sort<<<g.ncol, 512, smem, h->stream>>>(g, h->d_x, h->d_q, h->d_type, h->d_nbfp, h->dp.ntypes, h->d_col_count,
                                                        h->d_col_cell0, h->d_atom_index, h->d_slot_of_atom, h->d_xq, h->d_lj,
                                                        h->d_atype, h->d_bb, h->d_cellz);
        LAUNCH_CHECK(h);
    }
    g.valid         = 1;
    h->ncells_total = ncells_total;
    h->npad         = (int)npad;
    h->ncol_total   = g.col0 + g.ncol + 1;
    h->have_list    = false;
    if (gi == 0)
    {
        double s = 0;
        for (int a = atom_begin; a < atom_end && a < (int)h->h_q.size(); a++) s += (double)h->h_q[a] * h->h_q[a]; /* informational only */
        h->sum_q2 = s;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* pair search                                                                                             */
/* ------------------------------------------------------------------------------------------------------ */

struct SearchArgs
{
    GridDesc gi, gj;
    int      intra;      /* i and j grid identical: half list */
    int      shp[3];     /* shift range per dimension (pairlist.cpp:3168-3190) */
    float    box[3];
    float    rlist2, rlist;
    int      max_tiles;
    int      pass;       /* 0 count, 1 fill at the scanned offsets, 2 single pass: space claimed with atomics, bounded by cap_* */
    long long cap_tiles, cap_entries;
};

__device__ __forceinline__ float bb_dist2(const float* ilo, const float* ihi, const float* jb)
{
    /* pairlist.cpp:326-347 clusterBoundingBoxDistance2 */
    float d2 = 0;
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
        float dl = ilo[d] - jb[3 + d];
        float dh = jb[d] - ihi[d];
        float dm = fmaxf(fmaxf(dl, dh), 0.0f);
        d2 += dm * dm;
    }
    return d2;
}

/* One warp per i-cluster. Lane = jl + 8*ih handles atom pairs (2*ih, jl) and (2*ih+1, jl) of a candidate tile.
 * A cluster pair enters the list iff at least one atom pair has r^2 < rlist^2 (the converged result of
 * the reference's bounding-box search, pairlist.cpp:1090-1244, followed by its list pruning,
 * nbnxm_cuda_kernel_pruneonly.cuh), so the list is the tightest superset of the in-range pairs.
 * Exclusion masks (pairlist.cpp:1874-1972): bit `lane` of mask word w is 1 when atoms (2*ih+w, jl) interact. */
#ifndef NB_SEARCH_BLOCKS
#define NB_SEARCH_BLOCKS 6 /* resident CTAs of 4 warps per SM */
#endif
__global__ void __launch_bounds__(128, NB_SEARCH_BLOCKS)
k_search(SearchArgs A, const float* __restrict__ xq, const float* __restrict__ bb, const float* __restrict__ cellz,
         const int* __restrict__ col_cell0, const int* __restrict__ atom_index, const int* __restrict__ excl_off,
         const int* __restrict__ excl_idx, const float* __restrict__ shift_vec, int* __restrict__ cnt_tiles,
         int* __restrict__ cnt_entries, Entry* __restrict__ entries, int* __restrict__ tile_cj, uint64_t* __restrict__ tile_mask,
         int* __restrict__ err_flag, unsigned long long* __restrict__ claim)
{
    __shared__ int      s_cj[4][NB_MAX_GROUP_TILES];
    __shared__ uint64_t s_mask[4][NB_MAX_GROUP_TILES];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncl_i = A.gi.ncells * 8;
    const int cil   = blockIdx.x * 4 + w; /* grid-local i-cluster */
    if (cil >= ncl_i) return;
    const int ci = A.gi.cell0 * 8 + cil;
    const float* ib = bb + (size_t)ci * 6;
    const bool   i_empty = ib[0] > ib[3];
    int tile_cursor = 0, entry_cursor = 0;
    if (A.pass == 1)
    {
        tile_cursor  = cnt_tiles[cil];
        entry_cursor = cnt_entries[cil];
    }
    int ntiles_total = 0, nentries_total = 0;
    if (!i_empty)
    {
        const int    jl = lane & 7, ih = lane >> 3;
        const float4 xa = reinterpret_cast<const float4*>(xq)[(size_t)ci * 8 + 2 * ih];
        const float4 xb = reinterpret_cast<const float4*>(xq)[(size_t)ci * 8 + 2 * ih + 1];
        const int    ai0 = atom_index[ci * 8 + 2 * ih], ai1 = atom_index[ci * 8 + 2 * ih + 1];
        int e00 = 0, e01 = 0, e10 = 0, e11 = 0;
        if (excl_off)
        {
            if (ai0 >= 0)
            {
                e00 = excl_off[ai0];
                e01 = excl_off[ai0 + 1];
            }
            if (ai1 >= 0)
            {
                e10 = excl_off[ai1];
                e11 = excl_off[ai1 + 1];
            }
        }
        /* index range each i-atom's exclusion list spans (an empty list: an empty range) */
        int xlo0 = 0x7fffffff, xhi0 = -1, xlo1 = 0x7fffffff, xhi1 = -1;
        for (int e = e00; e < e01; e++)
        {
            const int v = excl_idx[e];
            xlo0 = min(xlo0, v), xhi0 = max(xhi0, v);
        }
        for (int e = e10; e < e11; e++)
        {
            const int v = excl_idx[e];
            xlo1 = min(xlo1, v), xhi1 = max(xhi1, v);
        }
        const bool have_excl = __any_sync(0xffffffffu, e01 > e00 || e11 > e10);
        for (int tz = -A.shp[2]; tz <= A.shp[2]; tz++)
            for (int ty = -A.shp[1]; ty <= A.shp[1]; ty++)
                for (int tx = -A.shp[0]; tx <= A.shp[0]; tx++)
                {
                    const int shift = 5 * (3 * (tz + 1) + (ty + 1)) + tx + 2; /* pbcutil/ishift.h:50 */
                    if (A.intra && shift > B200NB_CENTRAL) continue;           /* pairlist.cpp:3339-3342 */
                    const float sx = shift_vec[3 * shift], sy = shift_vec[3 * shift + 1], sz = shift_vec[3 * shift + 2];
                    float ilo[3] = { ib[0] + sx, ib[1] + sy, ib[2] + sz };
                    float ihi[3] = { ib[3] + sx, ib[4] + sy, ib[5] + sz };
                    /* column range reachable within rlist; edge columns also hold atoms beyond the grid bounds */
                    int cx0 = (int)floorf((ilo[0] - A.rlist - A.gj.lower[0]) * A.gj.inv_cell[0]);
                    int cx1 = (int)floorf((ihi[0] + A.rlist - A.gj.lower[0]) * A.gj.inv_cell[0]);
                    int cy0 = (int)floorf((ilo[1] - A.rlist - A.gj.lower[1]) * A.gj.inv_cell[1]);
                    int cy1 = (int)floorf((ihi[1] + A.rlist - A.gj.lower[1]) * A.gj.inv_cell[1]);
                    if (cx1 < 0 || cy1 < 0 || cx0 > A.gj.ncx - 1 || cy0 > A.gj.ncy - 1)
                    {
                        /* entirely outside: only the clamped edge columns could hold stray atoms; they are
                         * covered because the bounding-box test below uses real atom extents */
                    }
                    cx0 = max(cx0, 0);
                    cy0 = max(cy0, 0);
                    cx1 = min(cx1, A.gj.ncx - 1);
                    cy1 = min(cy1, A.gj.ncy - 1);
                    /* the shifted i-atoms, as in the kernels */
                    const float xi0 = xa.x + sx, yi0 = xa.y + sy, zi0 = xa.z + sz, xi1 = xb.x + sx, yi1 = xb.y + sy, zi1 = xb.z + sz;
                    int n_mask = 0, n_plain = 0; /* masked tiles grow from the front, plain ones from the back */
                    for (int cx = cx0; cx <= cx1; cx++)
                        for (int cy = cy0; cy <= cy1; cy++)
                        {
                            const int col = A.gj.col0 + cx * A.gj.ncy + cy;
                            const int c0 = col_cell0[col], c1 = col_cell0[col + 1];
                            if (c1 <= c0) continue;
                            /* cells whose z range can be within rlist: cells are z-ordered (grid.cpp:1092-1103) */
                            int first = c1, last = c0 - 1;
                            for (int cb = c0; cb < c1; cb += 32)
                            {
                                int  c  = cb + lane;
                                bool ok = false;
                                if (c < c1) ok = (cellz[2 * c + 1] >= ilo[2] - A.rlist) && (cellz[2 * c] <= ihi[2] + A.rlist);
                                unsigned m = __ballot_sync(0xffffffffu, ok);
                                if (m)
                                {
                                    first = min(first, cb + __ffs(m) - 1);
                                    last  = max(last, cb + 31 - __clz(m));
                                }
                            }
                            if (last < first) continue;
                            int cj_begin = first * 8, cj_end = (last + 1) * 8;
                            if (A.intra && shift == B200NB_CENTRAL) cj_begin = max(cj_begin, ci); /* pairlist.cpp:3365-3372: j >= i */
                            for (int cjb = cj_begin; cjb < cj_end; cjb += 32)
                            {
                                const int cjc = cjb + lane;
                                bool      cand = false;
                                if (cjc < cj_end)
                                {
                                    const float* jb = bb + (size_t)cjc * 6;
                                    cand            = (jb[0] <= jb[3]) && bb_dist2(ilo, ihi, jb) < A.rlist2;
                                }
                                unsigned cm = __ballot_sync(0xffffffffu, cand);
                                /* the candidates of this batch, four at a time: their coordinate (and atom index) loads are
                                 * issued together, so the warp waits for one memory round trip per four candidates, not per one */
                                while (cm)
                                {
                                    int    cjs[4];
                                    float4 xjs[4];
                                    int    ajs[4];
#pragma unroll
                                    for (int q = 0; q < 4; q++)
                                    {
                                        cjs[q] = -1;
                                        if (cm)
                                        {
                                            cjs[q] = cjb + __ffs(cm) - 1;
                                            cm &= cm - 1;
                                        }
                                    }
#pragma unroll
                                    for (int q = 0; q < 4; q++)
                                        if (cjs[q] >= 0)
                                        {
                                            xjs[q] = reinterpret_cast<const float4*>(xq)[(size_t)cjs[q] * 8 + jl];
                                            ajs[q] = have_excl ? atom_index[cjs[q] * 8 + jl] : -1;
                                        }
#pragma unroll
                                    for (int q = 0; q < 4; q++)
                                    {
                                        if (cjs[q] < 0) continue;
                                        const int    cj  = cjs[q];
                                        const float4 xj  = xjs[q];
                                        const float  r2a = nb_rsq(xi0, yi0, zi0, xj.x, xj.y, xj.z);
                                        const float  r2b = nb_rsq(xi1, yi1, zi1, xj.x, xj.y, xj.z);
                                        const bool   in  = (r2a < A.rlist2) || (r2b < A.rlist2);
                                        if (!__any_sync(0xffffffffu, in)) continue;
                                        /* interaction bits from the topology exclusions of the two i-atoms; the lists are only
                                         * walked when the j-atom's index lies inside the index range either of them spans */
                                        uint64_t  mask = ~0ull;
                                        const int aj   = ajs[q];
                                        const bool near = have_excl && ((aj >= xlo0 && aj <= xhi0) || (aj >= xlo1 && aj <= xhi1));
                                        if (__any_sync(0xffffffffu, near))
                                        {
                                            bool ia = true, ibit = true;
                                            if (near)
                                            {
                                                for (int e = e00; e < e01; e++) ia = ia && (excl_idx[e] != aj);
                                                for (int e = e10; e < e11; e++) ibit = ibit && (excl_idx[e] != aj);
                                            }
                                            const unsigned ma = __ballot_sync(0xffffffffu, ia);
                                            const unsigned mb = __ballot_sync(0xffffffffu, ibit);
                                            mask              = ((uint64_t)mb << 32) | ma;
                                        }
                                        const bool masked = (mask != ~0ull) || (A.intra && shift == B200NB_CENTRAL && cj == ci);
                                        if (n_mask + n_plain >= NB_MAX_GROUP_TILES)
                                        {
                                            if (lane == 0) atomicOr(err_flag, 1);
                                            continue;
                                        }
                                        if (lane == 0)
                                        {
                                            int pos = masked ? n_mask : NB_MAX_GROUP_TILES - 1 - n_plain;
                                            s_cj[w][pos]   = cj;
                                            s_mask[w][pos] = mask;
                                        }
                                        if (masked) n_mask++;
                                        else n_plain++;
                                    }
                                }
                            }
                        }
                    /* close this (ci, shift) group: pairlist.cpp:2167-2194 closeIEntry + split_sci_entry */
                    const int n = n_mask + n_plain;
                    if (n == 0) continue;
                    /* equal parts instead of full parts plus a short remainder: the same number of entries, none of them tiny */
                    const int nchunks = (n + A.max_tiles - 1) / A.max_tiles;
                    const int csize   = (n + nchunks - 1) / nchunks;
                    bool write = A.pass == 1;
                    if (A.pass == 2)
                    {
                        /* single pass: this group's room in the list is claimed here; a list that outgrew the buffers of the
                         * previous search raises err_flag bit 1 and the host repeats the search with the two-pass scheme */
                        unsigned long long t0 = 0, e0 = 0;
                        if (lane == 0)
                        {
                            t0 = atomicAdd(claim, (unsigned long long)n);
                            e0 = atomicAdd(claim + 1, (unsigned long long)nchunks);
                        }
                        t0           = __shfl_sync(0xffffffffu, t0, 0);
                        e0           = __shfl_sync(0xffffffffu, e0, 0);
                        tile_cursor  = (int)t0;
                        entry_cursor = (int)e0;
                        write        = (long long)(t0 + n) <= A.cap_tiles && (long long)(e0 + nchunks) <= A.cap_entries;
                        if (!write && lane == 0) atomicOr(err_flag, 2);
                    }
                    if (write)
                    {
                        __syncwarp();
                        for (int k = lane; k < n; k += 32)
                        {
                            /* masked tiles first, then the plain ones in discovery order */
                            int src = (k < n_mask) ? k : NB_MAX_GROUP_TILES - 1 - (k - n_mask);
                            tile_cj[tile_cursor + k]   = s_cj[w][src];
                            tile_mask[tile_cursor + k] = s_mask[w][src];
                        }
                        for (int k = lane; k < nchunks; k += 32)
                        {
                            Entry e;
                            e.ci          = ci;
                            int nm        = min(max(n_mask - k * csize, 0), csize);
                            e.shift_nmask = shift | (nm << 8);
                            e.start       = tile_cursor + k * csize;
                            e.end         = tile_cursor + min((k + 1) * csize, n);
                            entries[entry_cursor + k] = e;
                        }
                        __syncwarp();
                    }
                    tile_cursor += n;
                    entry_cursor += nchunks;
                    ntiles_total += n;
                    nentries_total += nchunks;
                }
    }
    if (A.pass == 0 && lane == 0)
    {
        cnt_tiles[cil]   = ntiles_total;
        cnt_entries[cil] = nentries_total;
    }
}

/* exclusive scan of two int arrays of length n (single CTA); totals -> out_tot[0..1] as long long */
__global__ void k_scan2(int* a, int* b, int n, long long* out_tot)
{
    __shared__ long long sa[1024], sb[1024];
    const int t = threadIdx.x, per = (n + blockDim.x - 1) / blockDim.x;
    long long s1 = 0, s2 = 0;
    for (int k = 0; k < per; k++)
    {
        int i = t * per + k;
        if (i < n)
        {
            s1 += a[i];
            s2 += b[i];
        }
    }
    sa[t] = s1;
    sb[t] = s2;
    __syncthreads();
    if (t == 0)
    {
        long long r1 = 0, r2 = 0;
        for (int i = 0; i < (int)blockDim.x; i++)
        {
            long long v1 = sa[i], v2 = sb[i];
            sa[i] = r1;
            sb[i] = r2;
            r1 += v1;
            r2 += v2;
        }
        out_tot[0] = r1;
        out_tot[1] = r2;
    }
    __syncthreads();
    long long r1 = sa[t], r2 = sb[t];
    for (int k = 0; k < per; k++)
    {
        int i = t * per + k;
        if (i < n)
        {
            int v1 = a[i], v2 = b[i];
            a[i] = (int)r1;
            b[i] = (int)r2;
            r1 += v1;
            r2 += v2;
        }
    }
}

/* Dynamic pruning (cuda/nbnxm_cuda_kernel_pruneonly.cuh:104-277; kernel_ref_prune.cpp:45-143): one warp per
 * entry of the outer list; a tile stays iff any atom pair has r^2 < rlist_inner^2. The inner list reuses the
 * outer list's segment layout, compacted in place per entry (masked tiles stay in front). */
__global__ void __launch_bounds__(128)
k_prune(const Entry* __restrict__ oe, const int* __restrict__ ocj, const uint64_t* __restrict__ omask, long long nentries,
        int part, int nparts, const float* __restrict__ xq, const float* __restrict__ shift_vec, float rlist2,
        Entry* __restrict__ ie, int* __restrict__ icj, uint64_t* __restrict__ imask)
{
    const long long wid = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const long long e   = wid * nparts + part; /* rolling parts interleave entries: pruneonly.cuh:165-166 */
    if (e >= nentries) return;
    const int   lane = threadIdx.x & 31, jl = lane & 7, ih = lane >> 3;
    const Entry en   = oe[e];
    const int   shift = NB_ENTRY_SHIFT(en.shift_nmask), nmask = NB_ENTRY_NMASK(en.shift_nmask);
    const float4 xa = reinterpret_cast<const float4*>(xq)[(size_t)en.ci * 8 + 2 * ih];
    const float4 xb = reinterpret_cast<const float4*>(xq)[(size_t)en.ci * 8 + 2 * ih + 1];
    const float  sx = shift_vec[3 * shift], sy = shift_vec[3 * shift + 1], sz = shift_vec[3 * shift + 2];
    const float  xi0 = xa.x + sx, yi0 = xa.y + sy, zi0 = xa.z + sz, xi1 = xb.x + sx, yi1 = xb.y + sy, zi1 = xb.z + sz;
    int kept = 0, kept_mask = 0;
    for (int t = en.start; t < en.end; t++)
    {
        const int    cj = ocj[t];
        const float4 xj = reinterpret_cast<const float4*>(xq)[(size_t)cj * 8 + jl];
        const bool   in = (nb_rsq(xi0, yi0, zi0, xj.x, xj.y, xj.z) < rlist2) || (nb_rsq(xi1, yi1, zi1, xj.x, xj.y, xj.z) < rlist2);
        if (__any_sync(0xffffffffu, in))
        {
            if (lane == 0)
            {
                icj[en.start + kept]   = cj;
                imask[en.start + kept] = omask[t];
            }
            if (t - en.start < nmask) kept_mask++;
            kept++;
        }
    }
    if (lane == 0)
    {
        Entry o;
        o.ci          = en.ci;
        o.shift_nmask = shift | (kept_mask << 8);
        o.start       = en.start;
        o.end         = en.start + kept;
        ie[e]         = o;
    }
}

/* Prune + re-pack: entries of the OUTER cluster-pair list at j-atom granularity for the force kernel (PackedList,
 * b200nb_internal.h), TWO half-entries per entry: one for the i-atoms 0-3 of the i-cluster, one for 4-7.  One warp per entry, lane =
 * jl + 8*ih as in the search; ONE pass over the entry's cluster pairs.  A j-atom is kept for a half iff one of its pairs with the
 * half's 4 i-atoms has r^2 < rlist2 (the inner, dynamic-pruning radius) and is not removed by the j > i rule of the self tile
 * -- which subsumes the cluster-pair prune of nbnxn_kernel_prune_cuda (cuda/nbnxm_cuda_kernel_pruneonly.cuh:104-277): a cluster
 * pair without a pair in range contributes no j-atom, so no separate prune kernel and no pruned cluster-pair list are needed on
 * the per-step path (rolling pruning = this kernel on one part of the outer list).  Kept j-atoms that have an excluded pair with
 * the half (or belong to the self tile) go to a front buffer, the others to a back buffer, and are written out front first so
 * that only the leading STEPS (16 j-atoms, what a half-warp of the force kernel consumes per iteration) need masks.  Mask of a
 * step: 64 bits, bit 16*k + j says pair (i-atom 4*half + k, j-atom j of the step) interacts.  Excluded pairs inside the cut-off
 * must stay: the kernels evaluate their Ewald / reaction-field exclusion correction (kernel_inner.h:330-360).  The j list of a
 * half-entry is padded to a whole number of steps with far-away dummy atoms.
 * The reference has no such step: its GPU list stays at 8x8 cluster-pair granularity with per-pair masks
 * (nbnxm/pairlist.h:190-225). */
#define NB_PACK_FRONT (NB_MAX_ENTRY_TILES * 8 + 16)
__global__ void __launch_bounds__(128)
k_pack(const Entry* __restrict__ ie, const int* __restrict__ icj, const uint64_t* __restrict__ imask, long long nentries, int part,
       int nparts, const float* __restrict__ xq, const float* __restrict__ shift_vec, float rlist2, int intra, int dummy_slot, int pitch,
       Entry* __restrict__ staged, int* __restrict__ sizes, int* __restrict__ pja, uint64_t* __restrict__ pmask,
       const int* __restrict__ dest, Entry* __restrict__ placed)
{
    __shared__ int      s_ja[4][2][NB_PACK_FRONT];            /* front, per half: j-atoms that need masks */
    __shared__ int      s_jb[4][2][NB_MAX_ENTRY_TILES * 8];   /* back: the others */
    __shared__ unsigned s_m[4][2][NB_MAX_ENTRY_TILES + 4];    /* 2 words per step of 16 j-atoms */
    __shared__ float4   s_xi[4][8];                           /* the entry's i-atoms, shifted */
    const int       w   = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * 4 + w;
    const long long e   = wid * nparts + part; /* rolling parts interleave entries: pruneonly.cuh:165-166 */
    if (e >= nentries) return;
    const int   lane = threadIdx.x & 31, jl = lane & 7, ih = lane >> 3;
    const Entry en   = ie[e];
    const int   shift = NB_ENTRY_SHIFT(en.shift_nmask), nmask = NB_ENTRY_NMASK(en.shift_nmask);
    const int   ntile = min(en.end - en.start, NB_MAX_ENTRY_TILES);
    for (int k = lane; k < 2 * (NB_MAX_ENTRY_TILES + 4); k += 32) (&s_m[w][0][0])[k] = 0u;
    __syncwarp();
    const float4 xa = reinterpret_cast<const float4*>(xq)[(size_t)en.ci * 8 + 2 * ih];
    const float4 xb = reinterpret_cast<const float4*>(xq)[(size_t)en.ci * 8 + 2 * ih + 1];
    const float  sx = shift_vec[3 * shift], sy = shift_vec[3 * shift + 1], sz = shift_vec[3 * shift + 2];
    const float  xi0 = xa.x + sx, yi0 = xa.y + sy, zi0 = xa.z + sz, xi1 = xb.x + sx, yi1 = xb.y + sy, zi1 = xb.z + sz;
    /* this lane's two pairs (i-atoms 2*ih, 2*ih+1) belong to half ih >> 1, where they are i-atoms k = 2*(ih & 1) and k + 1:
     * word (ih & 1) of the step mask, bits j and 16 + j */
    const int hf = ih >> 1, mword = ih & 1;
    int na0 = 0, na1 = 0, nb0 = 0, nb1 = 0, has_self = 0;
    /* ---- cluster pairs that carry a mask (exclusions, the self tile): sorted to the front of the entry by the search, a few per
     * entry.  Lane = jl + 8*ih looks at j-atom jl against i-atoms 2*ih, 2*ih + 1: the layout of the tile masks. ---- */
    const int nmt = min(nmask, ntile);
    for (int t = 0; t < nmt; t++)
    {
        const int      cj   = icj[en.start + t];
        const float4   xj   = reinterpret_cast<const float4*>(xq)[(size_t)cj * 8 + jl];
        const bool     diag = intra && shift == B200NB_CENTRAL && cj == en.ci;
        if (diag) has_self = 1;
        const bool     ina  = nb_rsq(xi0, yi0, zi0, xj.x, xj.y, xj.z) < rlist2 && !(diag && jl <= 2 * ih);
        const bool     inb  = nb_rsq(xi1, yi1, zi1, xj.x, xj.y, xj.z) < rlist2 && !(diag && jl <= 2 * ih + 1);
        const unsigned kb   = __ballot_sync(0xffffffffu, ina || inb);
        const unsigned kb0 = (kb | (kb >> 8)) & 0xffu, kb1 = ((kb >> 16) | (kb >> 24)) & 0xffu;
        if (!(kb0 | kb1)) continue; /* the cluster-pair prune: nothing of this tile within the radius */
        const uint64_t m  = imask[en.start + t];
        const unsigned ba = (unsigned)(m >> lane) & 1u, bb = (unsigned)(m >> (32 + lane)) & 1u;
        const unsigned sb = __ballot_sync(0xffffffffu, !(ba && bb) || diag);
        const unsigned sb0 = (sb | (sb >> 8)) & 0xffu, sb1 = ((sb >> 16) | (sb >> 24)) & 0xffu;
        const unsigned sela0 = kb0 & sb0, selb0 = kb0 & ~sb0, sela1 = kb1 & sb1, selb1 = kb1 & ~sb1;
        const unsigned sela = hf ? sela1 : sela0, selb = hf ? selb1 : selb0, below = (1u << jl) - 1u;
        if ((sela >> jl) & 1u)
        {
            const int pos = (hf ? na1 : na0) + __popc(sela & below);
            if (mword == 0) s_ja[w][hf][pos] = cj * 8 + jl;
            atomicOr(&s_m[w][hf][(pos >> 4) * 2 + mword], (ba << (pos & 15)) | (bb << (16 + (pos & 15))));
        }
        if (mword == 0 && ((selb >> jl) & 1u)) s_jb[w][hf][(hf ? nb1 : nb0) + __popc(selb & below)] = cj * 8 + jl;
        na0 += __popc(sela0), na1 += __popc(sela1);
        nb0 += __popc(selb0), nb1 += __popc(selb1);
    }
    /* ---- the rest (all pairs interact, never the self tile): FOUR cluster pairs per iteration, lane = jl + 8*q tests j-atom jl
     * of pair t + q against the 8 i-atoms (shifted coordinates broadcast from shared memory), 0-3 and 4-7 separately: two ballots
     * and two compactions per 32 j-atoms.  Cluster indices are fetched two groups ahead, coordinates one group ahead. ---- */
    {
        if (lane < 8)
        {
            const float4 v = reinterpret_cast<const float4*>(xq)[(size_t)en.ci * 8 + lane];
            s_xi[w][lane]  = make_float4(v.x + sx, v.y + sy, v.z + sz, 0.f);
        }
        __syncwarp();
        const int      q      = lane >> 3;
        const int      tbase  = en.start + nmt + q;
        const int      nplain = ntile - nmt;
        const float4   far    = make_float4(3.0e30f, 3.0e30f, 3.0e30f, 0.f);
        const unsigned lt     = (1u << lane) - 1u;
        int            cj_cur = q < nplain ? icj[tbase] : -1, cj_next = q + 4 < nplain ? icj[tbase + 4] : -1;
        float4         xj_cur = cj_cur >= 0 ? reinterpret_cast<const float4*>(xq)[(size_t)cj_cur * 8 + jl] : far;
        for (int g = 0; g < nplain; g += 4)
        {
            const int    cj = cj_cur;
            const float4 xj = xj_cur;
            cj_cur          = cj_next;
            xj_cur          = cj_cur >= 0 ? reinterpret_cast<const float4*>(xq)[(size_t)cj_cur * 8 + jl] : far;
            cj_next         = g + 8 + q < nplain ? icj[tbase + g + 8] : -1;
            bool in0 = false, in1 = false;
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                const float4 xl = s_xi[w][i], xu = s_xi[w][i + 4];
                in0             = in0 || nb_rsq(xl.x, xl.y, xl.z, xj.x, xj.y, xj.z) < rlist2;
                in1             = in1 || nb_rsq(xu.x, xu.y, xu.z, xj.x, xj.y, xj.z) < rlist2;
            }
            in0                = in0 && cj >= 0;
            in1                = in1 && cj >= 0;
            const unsigned kb0 = __ballot_sync(0xffffffffu, in0), kb1 = __ballot_sync(0xffffffffu, in1);
            if (in0) s_jb[w][0][nb0 + __popc(kb0 & lt)] = cj * 8 + jl;
            if (in1) s_jb[w][1][nb1 + __popc(kb1 & lt)] = cj * 8 + jl;
            nb0 += __popc(kb0);
            nb1 += __popc(kb1);
        }
    }
    __syncwarp();
    /* the unmasked j-atoms (and the padding) that share the last masked step of a half interact with all 4 i-atoms: lanes 0-15
     * complete half 0's step, lanes 16-31 half 1's */
    {
        const int hh = lane >> 4, nah = hh ? na1 : na0;
        const int p  = (nah & ~15) + (lane & 15); /* position of this lane's j-atom in that step */
        if ((nah & 15) && p >= nah)
        {
            const unsigned b2 = (1u << (p & 15)) | (1u << ((p & 15) + 16));
            atomicOr(&s_m[w][hh][(p >> 4) * 2], b2), atomicOr(&s_m[w][hh][(p >> 4) * 2 + 1], b2);
        }
    }
    __syncwarp();
#pragma unroll
    for (int hh = 0; hh < 2; hh++)
    {
        const int       na = hh ? na1 : na0, n = na + (hh ? nb1 : nb0);
        const int       nsp = (n + 15) >> 4, nms = (na + 15) >> 4; /* steps, masked steps */
        const long long p   = 2 * e + hh;                          /* half-entry in packing order */
        const long long s0  = p * ((pitch >> 1) + NB_PACK_TAIL);   /* its row: pitch/2 steps + the tail of dummy atoms */
        const int       dummy0 = dummy_slot + (int)(p & (NB_DUMMY_SLOTS / 16 - 1)) * 16;
        for (int k = lane; k < (nsp + NB_PACK_TAIL) * 16; k += 32)
            pja[(size_t)s0 * 16 + k] = k < na ? s_ja[w][hh][k] : (k < n ? s_jb[w][hh][k - na] : dummy0 + (k & 15));
        unsigned* const pm = reinterpret_cast<unsigned*>(pmask + s0);
        for (int k = lane; k < nsp * 2; k += 32) pm[k] = k < nms * 2 ? s_m[w][hh][k] : ~0u;
        if (lane == 0)
        {
            Entry o;
            o.ci          = en.ci;
            o.shift_nmask = shift | (nms << 8) | (has_self << 24) | (hh << 25); /* packed list: the mask count is in STEPS */
            o.start       = (int)s0;
            o.end         = (int)s0 + nsp;
            staged[p]     = o;
            sizes[p]      = nsp;
            if (dest) placed[dest[p]] = o; /* rolling part: the execution order of the last full pack stands */
        }
    }
}

/* Positions of the packed half-entries: descending step count, and -- STABLE -- packing order within a step count (counting
 * sort on <= 33 values; the role of sort_sci, nbnxm/pairlist.cpp:3827-3873).  The force kernel runs two consecutive half-entries
 * per single-warp CTA and CTAs start in index order, so (i) the two halves of a warp run the same number of steps, (ii) the big
 * ones start first and the last wave holds only the smallest: the tail of the kernel shrinks from one full entry to one short
 * entry, (iii) the packing order is the grid order of the i-clusters, so the two half-entries of a warp, and the warps that run
 * at the same time, belong to neighbouring i-clusters and share most of their j-atoms: their gathers and their j-force
 * reductions hit the same lines.  blkcnt: NB_ORDER_BINS x nblk ints (per bin and block of 256 half-entries: count, then offset
 * inside the bin); base: NB_ORDER_BINS ints (where each bin starts). */
__global__ void __launch_bounds__(256) k_order_count(const int* __restrict__ sizes, int n, int* __restrict__ blkcnt)
{
    __shared__ int cnt[NB_ORDER_BINS];
    if (threadIdx.x < NB_ORDER_BINS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e < n) atomicAdd(&cnt[min(sizes[e], NB_ORDER_BINS - 1)], 1);
    __syncthreads();
    if (threadIdx.x < NB_ORDER_BINS) blkcnt[threadIdx.x * gridDim.x + blockIdx.x] = cnt[threadIdx.x];
}
/* one CTA per bin: exclusive scan of the bin's row of block counts (256 blocks per round, warp scans), row total -> tot[bin] */
__global__ void __launch_bounds__(256) k_order_scan_rows(int* __restrict__ blkcnt, int nblk, int* __restrict__ tot)
{
    __shared__ int ws[8];
    __shared__ int carry;
    int* const row = blkcnt + (size_t)blockIdx.x * nblk;
    const int  lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblk; b0 += 256)
    {
        const int b = b0 + threadIdx.x;
        const int v = b < nblk ? row[b] : 0;
        int       s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        if (lane == 31) ws[w] = s;
        __syncthreads();
        int before = carry;
        for (int k = 0; k < w; k++) before += ws[k];
        if (b < nblk) row[b] = before + s - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = before + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) tot[blockIdx.x] = carry;
}
/* where each bin starts: after all bins of more steps (largest first) */
__global__ void __launch_bounds__(NB_ORDER_BINS) k_order_base(const int* __restrict__ tot, int* __restrict__ base)
{
    __shared__ int t[NB_ORDER_BINS];
    const int c = threadIdx.x;
    t[c]        = tot[c];
    __syncthreads();
    int before = 0;
    for (int k = c + 1; k < NB_ORDER_BINS; k++) before += t[k];
    base[c] = before;
}
__global__ void __launch_bounds__(256) k_order_assign(const int* __restrict__ sizes, int n, const int* __restrict__ blkcnt, const int* __restrict__ base,
                                                      int* __restrict__ dest)
{
    __shared__ int wcnt[8][NB_ORDER_BINS];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < 8 * NB_ORDER_BINS; k += 256) (&wcnt[0][0])[k] = 0;
    __syncthreads();
    const int e   = blockIdx.x * 256 + threadIdx.x;
    const int bin = e < n ? min(sizes[e], NB_ORDER_BINS - 1) : -1 - lane; /* idle lanes: a bin of their own each */
    const unsigned same = __match_any_sync(0xffffffffu, bin);
    const int      rank = __popc(same & ((1u << lane) - 1u));
    if (bin >= 0 && rank == 0) wcnt[w][bin] = __popc(same);
    __syncthreads();
    if (bin >= 0)
    {
        int before = 0;
        for (int k = 0; k < w; k++) before += wcnt[k][bin];
        dest[e] = base[bin] + blkcnt[bin * gridDim.x + blockIdx.x] + before + rank;
    }
}
/* headers from packing order into execution order.  After a rolling part the positions of the last full pack are kept: the
 * part's half-entries only shrink or grow a little, the order stays nearly sorted. */
__global__ void k_place_headers(const Entry* __restrict__ staged, const int* __restrict__ dest, int n, Entry* __restrict__ entries)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) reinterpret_cast<int4*>(entries)[dest[e]] = reinterpret_cast<const int4*>(staged)[e];
}

/* The two half-entries a warp of the force kernel runs side by side (positions 2w and 2w + 1 of the execution order) are walked
 * in lockstep for the longer one's steps: the shorter one gets the difference as steps of far-away dummy atoms behind its own
 * (its row has pitch/2 >= the longer one's steps of room, plus the tail every row ends in).  Differences are rare (the order is by step count) and short. */
__global__ void k_pad_partner(const Entry* __restrict__ entries, int nwarps, int* __restrict__ pja, int dummy_slot, const int* __restrict__ dest,
                              int part, int nparts)
{
    /* full pack: one thread per warp of the force kernel.  Rolling part: one thread per half-entry of the part (those of the outer
     * entries part, part + nparts, ...), which looks after the warp its half-entry runs in; two threads may then treat the same
     * warp and write the same values */
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (dest)
    {
        const long long p = 2 * ((long long)(w >> 1) * nparts + part) + (w & 1);
        if (p >= 2LL * nwarps) return;
        w = dest[p] >> 1;
    }
    if (w >= nwarps) return;
    const int4 a = reinterpret_cast<const int4*>(entries)[2 * w], b = reinterpret_cast<const int4*>(entries)[2 * w + 1];
    const int  na = a.w - a.z, nb = b.w - b.z;
    if (na == nb) return;
    const int   s0 = na < nb ? a.z : b.z, ns = min(na, nb), nl = max(na, nb);
    const int   d0 = dummy_slot + (int)((2 * w + (na < nb ? 0 : 1)) & (NB_DUMMY_SLOTS / 16 - 1)) * 16;
    for (int s = ns; s < nl + NB_PACK_TAIL; s++)
        for (int j = 0; j < 16; j++) pja[(size_t)(s0 + s) * 16 + j] = d0 + j;
}

static int ensure_packed(b200nb_context* h, PackedList& P, size_t cap_tiles, size_t cap_entries)
{
    /* cap_tiles = outer entries x (pitch + 2 NB_PACK_TAIL): per outer entry 2 rows of pitch/2 + NB_PACK_TAIL steps x (16 slots, one
     * 64-bit mask) */
    if (cap_tiles > P.cap_tiles || !P.ja)
    {
        cudaFree(P.ja);
        cudaFree(P.mask);
        P.ja = nullptr;
        P.mask = nullptr;
        NB_CUDA(h, cudaMalloc((void**)&P.ja, std::max<size_t>(cap_tiles, 1) * 16 * sizeof(int)));
        NB_CUDA(h, cudaMalloc((void**)&P.mask, std::max<size_t>(cap_tiles, 1) * sizeof(uint64_t)));
        P.cap_tiles = cap_tiles;
    }
    if (cap_entries > P.cap_entries || !P.entries)
    {
        cudaFree(P.entries);
        cudaFree(P.staged);
        cudaFree(P.dest);
        cudaFree(P.sizes);
        P.entries = P.staged = nullptr;
        P.dest = P.sizes = nullptr;
        const size_t nh = 2 * std::max<size_t>(cap_entries, 1);
        NB_CUDA(h, cudaMalloc((void**)&P.entries, nh * sizeof(Entry)));
        NB_CUDA(h, cudaMalloc((void**)&P.staged, nh * sizeof(Entry)));
        NB_CUDA(h, cudaMalloc((void**)&P.dest, nh * sizeof(int)));
        NB_CUDA(h, cudaMalloc((void**)&P.sizes, nh * sizeof(int)));
        cudaFree(P.order_blk);
        P.order_blk = nullptr;
        NB_CUDA(h, cudaMalloc((void**)&P.order_blk, ((nh + 255) / 256 + 1) * NB_ORDER_BINS * sizeof(int)));
        P.cap_entries = cap_entries;
    }
    return 0;
}

/* prunes + packs part `part` of `nparts` of the outer list of locality loc (radius: the inner list radius) */
static int launch_pack(b200nb_context* h, int loc, int part, int nparts)
{
    const PairList& I = h->outer[loc];
    PackedList&     P = h->packed[loc];
    P.nentries        = 2 * I.nentries;
    P.pitch           = (h->max_tiles + 1) & ~1; /* whole steps of 16 j-atoms = 2 tiles */
    P.row             = P.pitch / 2 + NB_PACK_TAIL;
    if (I.nentries == 0) return 0;
    if ((size_t)I.nentries * 2 * P.row > 2000000000ull) return nb_fail(h, B200NB_ERR_CAPACITY, "build_pairlist: packed list exceeds 2^31 steps");
    if (ensure_packed(h, P, I.cap_entries * 2 * P.row, I.cap_entries)) return B200NB_ERR_CUDA;
    long long nw = (I.nentries - part + nparts - 1) / nparts;
    if (nw <= 0) return 0;
    const float    r2   = h->inner_is_outer ? h->dp.rlist_outer2 : h->dp.rlist_inner2;
    const unsigned nblk = (unsigned)((nw + 3) / 4);
    const bool rolling = nparts > 1; /* keeps the positions of the last full pack: the part's half-entries only shrink or grow a little */
    k_pack<<<nblk, 128, 0, h->stream>>>(I.entries, I.cj, I.mask, I.nentries, part, nparts, h->d_xq, h->d_shift_vec, r2, loc == 0,
                                        h->dummy_slot, P.pitch, P.staged, P.sizes, P.ja, P.mask, rolling ? P.dest : nullptr, P.entries);
    LAUNCH_CHECK(h);
    const int n = (int)P.nentries;
    if (!rolling)
    {
        /* full pack: execution order by the packed step counts */
        const int nblk = (n + 255) / 256;
        k_order_count<<<nblk, 256, 0, h->stream>>>(P.sizes, n, P.order_blk);
        LAUNCH_CHECK(h);
        k_order_scan_rows<<<NB_ORDER_BINS, 256, 0, h->stream>>>(P.order_blk, nblk, h->d_hist + NB_ORDER_BINS);
        LAUNCH_CHECK(h);
        k_order_base<<<1, NB_ORDER_BINS, 0, h->stream>>>(h->d_hist + NB_ORDER_BINS, h->d_hist);
        LAUNCH_CHECK(h);
        k_order_assign<<<nblk, 256, 0, h->stream>>>(P.sizes, n, P.order_blk, h->d_hist, P.dest);
        LAUNCH_CHECK(h);
        k_place_headers<<<(n + 255) / 256, 256, 0, h->stream>>>(P.staged, P.dest, n, P.entries);
        LAUNCH_CHECK(h);
        k_pad_partner<<<(n / 2 + 255) / 256, 256, 0, h->stream>>>(P.entries, n / 2, P.ja, h->dummy_slot, nullptr, 0, 1);
        LAUNCH_CHECK(h);
    }
    else
    {
        k_pad_partner<<<(unsigned)((2 * nw + 255) / 256), 256, 0, h->stream>>>(P.entries, n / 2, P.ja, h->dummy_slot, P.dest, part, nparts);
        LAUNCH_CHECK(h);
    }
    h->inner_stale[loc] = true; /* the pruned CLUSTER-PAIR list (introspection only) no longer matches the packed one */
    return 0;
}

static int ensure_list(b200nb_context* h, PairList& l, size_t ntiles, size_t nentries)
{
    if (ntiles > l.cap_tiles || !l.cj)
    {
        cudaFree(l.cj);
        cudaFree(l.mask);
        l.cj = nullptr;
        l.mask = nullptr;
        size_t cap = (size_t)(ntiles * 1.1) + 1024;
        NB_CUDA(h, cudaMalloc((void**)&l.cj, cap * sizeof(int)));
        NB_CUDA(h, cudaMalloc((void**)&l.mask, cap * sizeof(uint64_t)));
        l.cap_tiles = cap;
    }
    if (nentries > l.cap_entries || !l.entries)
    {
        cudaFree(l.entries);
        l.entries = nullptr;
        size_t cap = (size_t)(nentries * 1.1) + 1024;
        NB_CUDA(h, cudaMalloc((void**)&l.entries, cap * sizeof(Entry)));
        l.cap_entries = cap;
    }
    return 0;
}

static int launch_prune(b200nb_context* h, int loc, int part, int nparts)
{
    if (h->inner_is_outer) return 0;
    PairList& o = h->outer[loc];
    PairList& i = h->inner[loc];
    if (o.nentries == 0) return 0;
    long long nw = (o.nentries - part + nparts - 1) / nparts;
    if (nw <= 0) return 0;
    k_prune<<<(unsigned)((nw + 3) / 4), 128, 0, h->stream>>>(o.entries, o.cj, o.mask, o.nentries, part, nparts, h->d_xq, h->d_shift_vec,
                                                             h->dp.rlist_inner2, i.entries, i.cj, i.mask);
    LAUNCH_CHECK(h);
    return 0;
}

/* the pruned cluster-pair list, for b200nb_get_tiles / b200nb_get_stats only: the per-step path goes outer list -> packed list */
static int refresh_inner_list(b200nb_context* h)
{
    if (h->inner_is_outer) return 0;
    for (int loc = 0; loc < 2; loc++)
        if (h->inner_stale[loc])
        {
            if (launch_prune(h, loc, 0, 1)) return B200NB_ERR_CUDA;
            h->inner_stale[loc] = false;
        }
    return 0;
}

/* the far-away dummy atoms the packed list pads its last tiles with: the LAST NB_DUMMY_SLOTS slots of the slot arrays, written
 * once per allocation (a pair-search step neither moves nor rewrites them) */
static int write_dummy_atoms(b200nb_context* h)
{
    if ((size_t)h->npad + NB_DUMMY_SLOTS > h->cap_pad) return nb_fail(h, B200NB_ERR_CAPACITY, "build_pairlist: no room for the dummy atoms");
    const int slot = (int)h->cap_pad - NB_DUMMY_SLOTS;
    if (h->dummy_slot == slot && h->dummy_cap == h->cap_pad && h->dummy_ntypes == h->dp.ntypes) return 0;
    h->dummy_slot   = slot;
    h->dummy_cap    = h->cap_pad;
    h->dummy_ntypes = h->dp.ntypes;
    std::vector<float> dxq(4 * NB_DUMMY_SLOTS), dlj(2 * NB_DUMMY_SLOTS, 0.0f);
    std::vector<int>   dty(NB_DUMMY_SLOTS, h->dp.ntypes - 1);
    for (int k = 0; k < NB_DUMMY_SLOTS; k++)
    {
        dxq[4 * k] = dxq[4 * k + 1] = -3.0e6f;
        dxq[4 * k + 2]              = -3.0e6f - 64.0f * k;
        dxq[4 * k + 3]              = 0.0f;
    }
    NB_CUDA(h, cudaMemcpyAsync(h->d_xq + 4 * (size_t)slot, dxq.data(), sizeof(float) * dxq.size(), cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(h, cudaMemcpyAsync(h->d_lj + 2 * (size_t)slot, dlj.data(), sizeof(float) * dlj.size(), cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(h, cudaMemcpyAsync(h->d_atype + (size_t)slot, dty.data(), sizeof(int) * dty.size(), cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(h, cudaStreamSynchronize(h->stream)); /* the host vectors go out of scope */
    return 0;
}

extern "C" int b200nb_build_pairlist(b200nb_t* h)
{
    NvtxRange nvtx_("b200nb_build_pairlist");
    if (!h) return B200NB_ERR_ARG;
    if (!h->grid[0].valid) return nb_fail(h, B200NB_ERR_STATE, "build_pairlist: put_on_grid first");
    cudaSetDevice(h->device);
    if (const char* e = getenv("B200NB_SEARCH_TWO_PASS")) h->search_two_pass = atoi(e) != 0; /* A/B switch for profiles/ */
    const float rl = h->hp.rlist_outer;
    /* a periodic dimension must hold at least two list radii, else one pair has several images in range
     * (the reference handles that with shp[XX]=2, pairlist.cpp:3185-3188; outside our scope) */
    const bool triclinic = h->box_off[0] != 0.f || h->box_off[1] != 0.f || h->box_off[2] != 0.f;
    if (triclinic)
    {
        /* pbcutil/pbc.cpp:179-208 max_cutoff2: half the shortest box vector, and the smallest diagonal element (b_y less |c_y|) */
        const float* o   = h->box_off;
        const float  a2  = h->box[0] * h->box[0], b2 = o[0] * o[0] + h->box[1] * h->box[1], c2 = o[1] * o[1] + o[2] * o[2] + h->box[2] * h->box[2];
        const float  mss = std::min(h->box[0], std::min(h->box[1] - fabsf(o[2]), h->box[2]));
        if (rl * rl > std::min(0.25f * std::min(a2, std::min(b2, c2)), mss * mss))
            return nb_fail(h, B200NB_ERR_ARG, "build_pairlist: list radius exceeds what the triclinic cell allows (max_cutoff2)");
        if (h->grid[1].valid) return nb_fail(h, B200NB_ERR_ARG, "build_pairlist: triclinic cells are not combined with a halo grid");
    }
    else
        for (int d = 0; d < 3; d++)
            if (h->pbc[d] && h->box[d] < 2 * rl) return nb_fail(h, B200NB_ERR_ARG, "build_pairlist: box smaller than 2*rlist along a periodic dimension");
    /* list balancing granularity (the role of get_nsubpair_target, pairlist.cpp:2485-2587).  The force kernel is issue-bound
     * and an entry costs about 560 issue cycles on top of its tiles (prologue, masked first tile, i-force reduction;
     * profiles/r1/v_sweep_pair_loop_diagnostics.txt), so entries should be as long as the need for parallelism allows:
     * up to 48 k atoms 24 cluster pairs per entry keep ~2 entries per resident warp (r_sweep_sorted_entries.txt); larger
     * systems have entries to spare and take whole (i-cluster, shift) lists, up to NB_MAX_ENTRY_TILES */
    if (h->hp.max_tiles_per_entry <= 0)
        h->max_tiles = h->natoms <= 48000 ? 24 : std::min(NB_MAX_ENTRY_TILES, (int)(24.0 * h->natoms / 48000.0));
    const bool want_inner = h->dp.rlist_inner2 < h->dp.rlist_outer2;
    if (!want_inner && !h->inner_is_outer)
    {
        for (int l = 0; l < 2; l++) free_list(h->inner[l]);
    }
    if (want_inner && h->inner_is_outer)
    {
        for (int l = 0; l < 2; l++) h->inner[l] = PairList();
    }
    h->inner_is_outer = !want_inner;

    const int ncl_i = h->grid[0].ncells * 8;
    if ((size_t)ncl_i > h->cap_clusters)
    {
        cudaFree(h->d_cnt_tiles);
        cudaFree(h->d_cnt_entries);
        h->cap_clusters = (size_t)(ncl_i * 1.2) + 64;
        NB_CUDA(h, cudaMalloc((void**)&h->d_cnt_tiles, h->cap_clusters * sizeof(int)));
        NB_CUDA(h, cudaMalloc((void**)&h->d_cnt_entries, h->cap_clusters * sizeof(int)));
    }
    for (int loc = 0; loc < 2; loc++)
    {
        PairList& L = h->outer[loc];
        L.ntiles = L.nentries = 0;
        if (loc == 1 && !h->grid[1].valid)
        {
            /* no halo grid (any more): the non-local lists of an earlier search must not survive -- the packed one would be
             * re-packed against the new gridding and run by b200nb_launch_force(-1) / b200nb_dd_step */
            if (want_inner) h->inner[1].ntiles = h->inner[1].nentries = 0;
            else h->inner[1] = L;
            h->packed[1].nentries = 0;
            continue;
        }
        SearchArgs A;
        A.gi    = h->grid[0];
        A.gj    = h->grid[loc];
        A.intra = (loc == 0);
        for (int d = 0; d < 3; d++)
        {
            /* triclinic cells: images two box vectors away along x can be in range (nbnxm/pairlist.cpp:3181-3188 takes 2 when
             * a_x - |b_x| - |c_x| is below the cell-to-cell list range; always taking it costs a few empty shift iterations) */
            A.shp[d] = h->pbc[d] ? ((d == 0 && triclinic) ? 2 : 1) : 0;
            A.box[d] = h->box[d];
        }
        A.rlist     = rl;
        A.rlist2    = h->dp.rlist_outer2;
        A.max_tiles = h->max_tiles;
        NB_CUDA(h, cudaMemsetAsync(h->d_scratch + 8, 0, sizeof(int), h->stream));
        const unsigned nblk = (unsigned)((ncl_i + 3) / 4);
        long long      tot[2] = { 0, 0 };
        int            flag = 0;
        bool           done = false;
        unsigned long long* const claim = reinterpret_cast<unsigned long long*>(h->d_counter);
        if (L.cap_tiles > 0 && L.cap_entries > 0 && L.entries && !h->search_two_pass)
        {
            /* steady state: ONE pass into the buffers the previous search left (they hold 10 % more than it needed; a pair list
             * changes by well under 1 % between searches), room claimed with two atomics per (i-cluster, shift) group */
            A.pass        = 2;
            A.cap_tiles   = (long long)L.cap_tiles;
            A.cap_entries = (long long)L.cap_entries;
            NB_CUDA(h, cudaMemsetAsync(h->d_counter, 0, 2 * sizeof(long long), h->stream));
            k_search<<<nblk, 128, 0, h->stream>>>(A, h->d_xq, h->d_bb, h->d_cellz, h->d_col_cell0, h->d_atom_index, h->d_excl_off,
                                                  h->d_excl_idx, h->d_shift_vec, h->d_cnt_tiles, h->d_cnt_entries, L.entries, L.cj,
                                                  L.mask, h->d_scratch + 8, claim);
            LAUNCH_CHECK(h);
            NB_CUDA(h, cudaMemcpyAsync(tot, h->d_counter, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
            NB_CUDA(h, cudaMemcpyAsync(&flag, h->d_scratch + 8, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            NB_CUDA(h, cudaStreamSynchronize(h->stream));
            if (flag & 1) return nb_fail(h, B200NB_ERR_CAPACITY, "build_pairlist: more than 512 cluster pairs for one i-cluster and shift");
            done = !(flag & 2);
            if (!done) NB_CUDA(h, cudaMemsetAsync(h->d_scratch + 8, 0, sizeof(int), h->stream));
        }
        if (!done)
        {
            A.pass = 0;
            k_search<<<nblk, 128, 0, h->stream>>>(A, h->d_xq, h->d_bb, h->d_cellz, h->d_col_cell0, h->d_atom_index, h->d_excl_off,
                                                  h->d_excl_idx, h->d_shift_vec, h->d_cnt_tiles, h->d_cnt_entries, nullptr, nullptr,
                                                  nullptr, h->d_scratch + 8, claim);
            LAUNCH_CHECK(h);
            k_scan2<<<1, 1024, 0, h->stream>>>(h->d_cnt_tiles, h->d_cnt_entries, ncl_i, h->d_counter);
            LAUNCH_CHECK(h);
            NB_CUDA(h, cudaMemcpyAsync(tot, h->d_counter, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
            NB_CUDA(h, cudaMemcpyAsync(&flag, h->d_scratch + 8, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            NB_CUDA(h, cudaStreamSynchronize(h->stream));
            if (flag) return nb_fail(h, B200NB_ERR_CAPACITY, "build_pairlist: more than 512 cluster pairs for one i-cluster and shift");
            if (tot[0] > 2000000000LL) return nb_fail(h, B200NB_ERR_CAPACITY, "build_pairlist: pair list exceeds 2^31 cluster pairs");
            if (ensure_list(h, L, (size_t)tot[0], (size_t)tot[1])) return B200NB_ERR_CUDA;
            if (tot[0] > 0)
            {
                A.pass = 1;
                k_search<<<nblk, 128, 0, h->stream>>>(A, h->d_xq, h->d_bb, h->d_cellz, h->d_col_cell0, h->d_atom_index, h->d_excl_off,
                                                      h->d_excl_idx, h->d_shift_vec, h->d_cnt_tiles, h->d_cnt_entries, L.entries, L.cj,
                                                      L.mask, h->d_scratch + 8, claim);
                LAUNCH_CHECK(h);
            }
        }
        L.ntiles   = tot[0];
        L.nentries = tot[1];
        if (want_inner)
        {
            PairList& I = h->inner[loc];
            if (ensure_list(h, I, (size_t)tot[0], (size_t)tot[1])) return B200NB_ERR_CUDA;
            I.ntiles   = tot[0]; /* upper bound; exact count via get_stats */
            I.nentries = tot[1];
            /* the fresh-list prune (cuda/nbnxm_cuda.cu:510-517) is part of the packing below (k_pack prunes while it packs);
             * the pruned cluster-pair list itself is only materialised for introspection (refresh_inner_list) */
            h->inner_stale[loc] = true;
        }
        else
        {
            h->inner[loc] = L;
        }
    }
    if (int rc = write_dummy_atoms(h)) return rc;
    for (int loc = 0; loc < 2; loc++)
        if (launch_pack(h, loc, 0, 1)) return B200NB_ERR_CUDA;
    /* no synchronisation here: everything that follows (steps, prunes, introspection) is ordered on the stream */
    h->have_list = true;
    h->generation++;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* reference-built grid and pair list (the drop-in path behind Nbnxm::gpu_*, shim/nbnxm_b200.cpp)          */
/* ------------------------------------------------------------------------------------------------------ */
/* With mdrun / nblib in front, gridding and pair search stay where the reference does them for its CUDA backend -- on the
 * CPU (nbnxn_put_on_grid, PairlistSet::constructPairlists) -- and the backend receives the finished products: the atom
 * data in grid order (gpu_init_atomdata, nbnxm_gpu_data_mgmt.cpp) and the 8x8x8 super-cluster list (gpu_init_pairlist,
 * nbnxm_gpu_data_mgmt.cpp:251-311).  These entry points take exactly those, so the reference needs no patched call site. */

/* gpu_init_atomdata: nslots = nbat->numAtoms() (a multiple of 8), xq_host = nbat->x() in nbatXYZQ format (4 floats per slot,
 * atomdata.cpp:659-662), type_host = nbat->params().type (filler slots carry the extra zero-parameter type ntypes, atomdata.cpp:456) */
extern "C" int b200nb_set_grid_atoms(b200nb_t* h, int nslots, const float* xq_host, const int* type_host)
{
    if (!h || nslots < 8 || (nslots & 7) || !xq_host || !type_host) return nb_fail(h, B200NB_ERR_ARG, "set_grid_atoms: bad argument");
    if (!h->have_params) return nb_fail(h, B200NB_ERR_STATE, "set_grid_atoms: set_params first");
    cudaSetDevice(h->device);
    const int ntf = h->dp.ntypes;
    std::vector<float> ljv((size_t)nslots * 2);
    for (int s = 0; s < nslots; s++)
    {
        const int t = type_host[s];
        if (t < 0 || t >= ntf) return nb_fail(h, B200NB_ERR_ARG, "set_grid_atoms: atom type out of range");
        ljv[2 * (size_t)s]     = std::sqrt(h->nbfp_host[((size_t)t * ntf + t) * 2]);
        ljv[2 * (size_t)s + 1] = std::sqrt(h->nbfp_host[((size_t)t * ntf + t) * 2 + 1]);
    }
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    if ((size_t)nslots + NB_DUMMY_SLOTS > h->cap_pad)
    {
        const size_t cap = ((size_t)(nslots * 1.15) + NB_DUMMY_SLOTS + 1024 + 63) / 64 * 64;
        cudaFree(h->d_xq), cudaFree(h->d_lj), cudaFree(h->d_atype), cudaFree(h->d_f), cudaFree(h->d_atom_index);
        h->d_xq = h->d_lj = nullptr, h->d_atype = h->d_atom_index = nullptr, h->d_f = nullptr;
        h->cap_pad = 0;
        NB_CUDA(h, cudaMalloc((void**)&h->d_xq, cap * 4 * sizeof(float)));
        NB_CUDA(h, cudaMalloc((void**)&h->d_lj, cap * 2 * sizeof(float)));
        NB_CUDA(h, cudaMalloc((void**)&h->d_atype, cap * sizeof(int)));
        NB_CUDA(h, cudaMalloc((void**)&h->d_atom_index, cap * sizeof(int)));
        NB_CUDA(h, cudaMalloc((void**)&h->d_f, cap * sizeof(float4)));
        NB_CUDA(h, cudaMemset(h->d_f, 0, cap * sizeof(float4)));
        h->cap_pad = cap;
    }
    if (alloc_exact(h, &h->d_fout, (size_t)nslots * 3)) return B200NB_ERR_CUDA;
    h->npad          = nslots;
    h->natoms        = 0; /* no atom-order view in this mode: b200nb_step / b200nb_compute / get_f refuse to run */
    h->grid_uploaded = true;
    h->grid[0].valid = h->grid[1].valid = 0;
    {
        /* The reference parks all filler atoms of a grid at ONE far-away point (atomdata.cpp:146) and relies on its kernels
         * clamping r^2 (pairlist.h:146); our unmasked pair path has no clamp, so fillers get the unique far-away positions our
         * own gridding gives them (k_column_sort): no two at zero distance, none within reach of a real atom. */
        std::vector<float> xqv(xq_host, xq_host + 4 * (size_t)nslots);
        for (int s = 0; s < nslots; s++)
            if (type_host[s] == ntf - 1)
            {
                xqv[4 * (size_t)s]     = -1.0e6f - 8.0f * (float)(s & 0xfffff);
                xqv[4 * (size_t)s + 1] = -1.0e6f - 64.0f * (float)(s >> 20);
                xqv[4 * (size_t)s + 2] = -1.0e6f;
                xqv[4 * (size_t)s + 3] = 0.0f;
            }
        NB_CUDA(h, cudaMemcpy(h->d_xq, xqv.data(), sizeof(float) * 4 * (size_t)nslots, cudaMemcpyHostToDevice));
    }
    if (alloc_exact(h, &h->d_x, (size_t)nslots * 4)) return B200NB_ERR_CUDA; /* staging of b200nb_copy_xq_grid */
    NB_CUDA(h, cudaMemcpy(h->d_lj, ljv.data(), sizeof(float) * 2 * (size_t)nslots, cudaMemcpyHostToDevice));
    NB_CUDA(h, cudaMemcpy(h->d_atype, type_host, sizeof(int) * (size_t)nslots, cudaMemcpyHostToDevice));
    {
        /* slot -> "atom" for the introspection calls (b200nb_get_pairs reports SLOT indices in this mode): -1 marks fillers */
        std::vector<int> ai((size_t)nslots);
        for (int s = 0; s < nslots; s++) ai[s] = type_host[s] == ntf - 1 ? -1 : s;
        NB_CUDA(h, cudaMemcpy(h->d_atom_index, ai.data(), sizeof(int) * (size_t)nslots, cudaMemcpyHostToDevice));
    }
    h->have_list = false;
    h->generation++;
    return write_dummy_atoms(h);
}

/* gpu_init_pairlist (nbnxm_gpu_data_mgmt.cpp:251-311): takes the reference's NbnxnPairlistGpu (pairlist.h:267-299) and turns it
 * into our (i-cluster, shift) entries: every set bit (jm*8 + im) of a cj4's imask is one 8x8 cluster pair (ci = sci*8 + im,
 * cj = cj4.cj[jm]); its atom-pair interaction bits are excl[imei[jc/4].excl_ind].pair[(jc&3)*8 + ic] >> (jm*8 + im)
 * (kernel_gpu_ref.cpp:223-240), re-indexed to our lane order.  Cluster pairs with a cleared bit go to the front of their entry
 * (the format k_pack expects), entries are cut at max_tiles_per_entry.  The conversion runs on the host, like the search that
 * produced the list; from here on (prune, re-pack, force) everything is on the device. */
extern "C" int b200nb_upload_pairlist(b200nb_t* h, int locality, const b200nb_sci_t* sci, int nsci, const b200nb_cj4_t* cj4, int ncj4,
                                      const b200nb_excl_t* excl, int nexcl)
{
    NvtxRange nvtx_("b200nb_upload_pairlist");
    if (!h || locality < 0 || locality > 1 || nsci < 0 || ncj4 < 0 || nexcl < 1 || (nsci && !sci) || (ncj4 && !cj4) || !excl)
        return nb_fail(h, B200NB_ERR_ARG, "upload_pairlist: bad argument");
    if (!h->grid_uploaded) return nb_fail(h, B200NB_ERR_STATE, "upload_pairlist: set_grid_atoms first");
    cudaSetDevice(h->device);
    const int ncl = h->npad / 8;
    if (h->hp.max_tiles_per_entry <= 0)
        h->max_tiles = h->npad <= 48000 ? 24 : std::min(NB_MAX_ENTRY_TILES, (int)(24.0 * h->npad / 48000.0));
    const int maxt = h->max_tiles;
    struct Tile
    {
        int      cj;
        uint64_t mask;
    };
    std::vector<Entry>    ent;
    std::vector<int>      cjv;
    std::vector<uint64_t> mv;
    std::vector<Tile>     masked[8], plain[8];
    for (int s = 0; s < nsci; s++)
    {
        const b200nb_sci_t& S = sci[s];
        const int           shift = S.shift & 255; /* NBNXN_CI_SHIFT: the flags above bit 7 are CPU-list only */
        if (S.cj4_ind_start < 0 || S.cj4_ind_end > ncj4 || S.cj4_ind_start > S.cj4_ind_end || shift >= B200NB_SHIFTS || S.sci < 0
            || S.sci * 8 + 7 >= ncl)
            return nb_fail(h, B200NB_ERR_ARG, "upload_pairlist: sci entry out of range");
        for (int im = 0; im < 8; im++) masked[im].clear(), plain[im].clear();
        for (int c = S.cj4_ind_start; c < S.cj4_ind_end; c++)
        {
            const b200nb_cj4_t& C = cj4[c];
            const unsigned      im_all = C.imei[0].imask;
            if (!im_all) continue;
            if (C.imei[0].excl_ind < 0 || C.imei[0].excl_ind >= nexcl || C.imei[1].excl_ind < 0 || C.imei[1].excl_ind >= nexcl)
                return nb_fail(h, B200NB_ERR_ARG, "upload_pairlist: exclusion index out of range");
            const unsigned* e0 = excl[C.imei[0].excl_ind].pair;
            const unsigned* e1 = excl[C.imei[1].excl_ind].pair;
            for (int jm = 0; jm < 4; jm++)
                for (int im = 0; im < 8; im++)
                {
                    const int bit = jm * 8 + im;
                    if (!((im_all >> bit) & 1u)) continue;
                    if (C.cj[jm] < 0 || C.cj[jm] >= ncl) return nb_fail(h, B200NB_ERR_ARG, "upload_pairlist: j-cluster out of range");
                    uint64_t m = 0;
                    for (int jc = 0; jc < 8; jc++)
                    {
                        const unsigned* e = jc < 4 ? e0 : e1;
                        for (int ic = 0; ic < 8; ic++)
                            if ((e[(jc & 3) * 8 + ic] >> bit) & 1u) m |= 1ull << (32 * (ic & 1) + jc + 8 * (ic >> 1));
                    }
                    (m == ~0ull ? plain[im] : masked[im]).push_back(Tile{ C.cj[jm], m });
                }
        }
        for (int im = 0; im < 8; im++)
        {
            const size_t nm = masked[im].size(), nt = nm + plain[im].size();
            for (size_t t0 = 0; t0 < nt; t0 += (size_t)maxt)
            {
                const size_t t1 = std::min(nt, t0 + (size_t)maxt);
                Entry        E;
                E.ci          = S.sci * 8 + im;
                E.shift_nmask = shift | ((int)(t0 < nm ? std::min(nm, t1) - t0 : 0) << 8);
                E.start       = (int)cjv.size();
                for (size_t t = t0; t < t1; t++)
                {
                    const Tile& T = t < nm ? masked[im][t] : plain[im][t - nm];
                    cjv.push_back(T.cj);
                    mv.push_back(T.mask);
                }
                E.end = (int)cjv.size();
                ent.push_back(E);
            }
        }
    }
    const bool want_inner = h->dp.rlist_inner2 < h->dp.rlist_outer2;
    if (want_inner != !h->inner_is_outer)
    {
        /* switching between "inner list = outer list" (aliases) and a separate pruned list: as b200nb_build_pairlist */
        if (!want_inner) for (int l = 0; l < 2; l++) free_list(h->inner[l]);
        else for (int l = 0; l < 2; l++) h->inner[l] = PairList();
        h->inner_is_outer = !want_inner;
    }
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    PairList& L = h->outer[locality];
    if (ensure_list(h, L, cjv.size(), ent.size())) return B200NB_ERR_CUDA;
    L.ntiles   = (long long)cjv.size();
    L.nentries = (long long)ent.size();
    if (!ent.empty())
    {
        NB_CUDA(h, cudaMemcpy(L.entries, ent.data(), sizeof(Entry) * ent.size(), cudaMemcpyHostToDevice));
        NB_CUDA(h, cudaMemcpy(L.cj, cjv.data(), sizeof(int) * cjv.size(), cudaMemcpyHostToDevice));
        NB_CUDA(h, cudaMemcpy(L.mask, mv.data(), sizeof(uint64_t) * mv.size(), cudaMemcpyHostToDevice));
    }
    if (want_inner)
    {
        PairList& I = h->inner[locality];
        if (ensure_list(h, I, cjv.size(), ent.size())) return B200NB_ERR_CUDA;
        I.ntiles   = L.ntiles;
        I.nentries = L.nentries;
        h->inner_stale[locality] = true; /* fresh-list prune (cuda/nbnxm_cuda.cu:510-517): done by the packing below */
    }
    else
        h->inner[locality] = L;
    if (locality == 0 && !h->have_list)
    {
        /* a context that never had a non-local list: its packed list is empty, not stale */
        h->outer[1].ntiles = h->outer[1].nentries = 0;
        if (h->inner_is_outer) h->inner[1] = h->outer[1];
        else h->inner[1].ntiles = h->inner[1].nentries = 0;
        h->packed[1].nentries = 0;
    }
    if (launch_pack(h, locality, 0, 1)) return B200NB_ERR_CUDA;
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->have_list = true;
    h->generation++;
    return 0;
}

/* new coordinates of the real atoms of a reference-built grid; filler slots keep the positions set_grid_atoms gave them */
__global__ void k_xq_grid_update(const float4* __restrict__ in, const int* __restrict__ atype, int filler, int s0, int s1, float4* __restrict__ xq)
{
    const int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s < s1 && atype[s] != filler) xq[s] = in[s];
}

/* gpu_copy_xq_to_gpu (cuda/nbnxm_cuda.cu:395-465): slots [slot_begin, slot_end) of nbat->x(), asynchronous on the stream */
extern "C" int b200nb_copy_xq_grid(b200nb_t* h, const float* xq_host, int slot_begin, int slot_end)
{
    if (!h || !xq_host || slot_begin < 0 || slot_end < slot_begin) return nb_fail(h, B200NB_ERR_ARG, "copy_xq_grid: bad argument");
    if (!h->grid_uploaded || slot_end > h->npad) return nb_fail(h, B200NB_ERR_STATE, "copy_xq_grid: set_grid_atoms first / range beyond the grid");
    cudaSetDevice(h->device);
    const int n = slot_end - slot_begin;
    if (n == 0) return 0;
    NB_CUDA(h, cudaMemcpyAsync(h->d_x + 4 * (size_t)slot_begin, xq_host + 4 * (size_t)slot_begin, sizeof(float) * 4 * (size_t)n,
                               cudaMemcpyHostToDevice, h->stream));
    k_xq_grid_update<<<(n + 255) / 256, 256, 0, h->stream>>>(reinterpret_cast<const float4*>(h->d_x), h->d_atype, h->dp.ntypes - 1, slot_begin,
                                                             slot_end, reinterpret_cast<float4*>(h->d_xq));
    LAUNCH_CHECK(h);
    return 0;
}

__global__ void k_f_grid_to_f3(const float4* __restrict__ fg, int s0, int s1, float* __restrict__ out)
{
    const int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s1) return;
    const float4 v = fg[s];
    out[3 * (size_t)s]     = v.x;
    out[3 * (size_t)s + 1] = v.y;
    out[3 * (size_t)s + 2] = v.z;
}

/* gpu_launch_cpyback (cuda/nbnxm_cuda.cu:720-814), force part: grid-ordered forces of slots [slot_begin, slot_end) into
 * nbat->out[0].f (3 floats per slot), asynchronous on the stream -- b200nb_synchronize before reading */
extern "C" int b200nb_get_f_grid(b200nb_t* h, float* f_host, int slot_begin, int slot_end)
{
    if (!h || !f_host || slot_begin < 0 || slot_end < slot_begin) return nb_fail(h, B200NB_ERR_ARG, "get_f_grid: bad argument");
    if (!h->grid_uploaded || slot_end > h->npad) return nb_fail(h, B200NB_ERR_STATE, "get_f_grid: set_grid_atoms first / range beyond the grid");
    cudaSetDevice(h->device);
    const int n = slot_end - slot_begin;
    if (n == 0) return 0;
    k_f_grid_to_f3<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_f, slot_begin, slot_end, h->d_fout);
    LAUNCH_CHECK(h);
    NB_CUDA(h, cudaMemcpyAsync(f_host + 3 * (size_t)slot_begin, h->d_fout + 3 * (size_t)slot_begin, sizeof(float) * 3 * (size_t)n,
                               cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

extern "C" int b200nb_launch_prune(b200nb_t* h, int locality, int part, int num_parts)
{
    NvtxRange nvtx_("b200nb_launch_prune");
    if (!h || num_parts < 1 || part < 0 || part >= num_parts) return nb_fail(h, B200NB_ERR_ARG, "launch_prune: bad argument");
    if (!h->have_list) return nb_fail(h, B200NB_ERR_STATE, "launch_prune: no pair list");
    cudaSetDevice(h->device);
    for (int loc = 0; loc < 2; loc++)
        if (locality < 0 || locality == loc)
        {
            /* k_pack prunes the part's entries of the outer list to the inner radius while re-packing them */
            if (!h->inner_is_outer && launch_pack(h, loc, part, num_parts)) return B200NB_ERR_CUDA;
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* per-step buffer ops                                                                                     */
/* ------------------------------------------------------------------------------------------------------ */

/* nbnxn_gpu_x_to_nbat_x_kernel (cuda/nbnxm_buffer_ops_kernels.cuh:65-114): x (atom order) -> grid layout */
__global__ void k_x_to_grid(const float* __restrict__ x, const int* __restrict__ slot_of_atom, int a0, int a1, float* __restrict__ xq)
{
    int a = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= a1) return;
    float* xb = xq + 4 * (size_t)slot_of_atom[a];
    xb[0]     = x[3 * a];
    xb[1]     = x[3 * a + 1];
    xb[2]     = x[3 * a + 2];
}

/* reduceKernel (mdlib/gpuforcereduction_impl.cu:70-104): f[a] (+)= f_nb[cell[a]] */
__global__ void k_f_from_grid(const float4* __restrict__ fg, const int* __restrict__ slot_of_atom, int a0, int a1, int accumulate,
                              float* __restrict__ f)
{
    int a = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= a1) return;
    const float4 v = fg[slot_of_atom[a]];
    if (accumulate)
    {
        f[3 * a] += v.x;
        f[3 * a + 1] += v.y;
        f[3 * a + 2] += v.z;
    }
    else
    {
        f[3 * a]     = v.x;
        f[3 * a + 1] = v.y;
        f[3 * a + 2] = v.z;
    }
}

extern "C" int b200nb_set_x(b200nb_t* h, const float* x, int x_on_device, int a0, int a1)
{
    NvtxRange nvtx_("b200nb_set_x");
    if (!h || !x) return nb_fail(h, B200NB_ERR_ARG, "set_x: bad argument");
    if (!h->grid[0].valid) return nb_fail(h, B200NB_ERR_STATE, "set_x: put_on_grid first");
    if (a0 < 0 || a1 > h->natoms || a0 > a1) return nb_fail(h, B200NB_ERR_ARG, "set_x: bad atom range");
    if (a1 == a0) return 0;
    cudaSetDevice(h->device);
    const int n = a1 - a0;
    const float* src = x + 3 * (size_t)a0;
    if (x_on_device)
    {
        k_x_to_grid<<<(n + 255) / 256, 256, 0, h->stream>>>(x, h->d_slot_of_atom, a0, a1, h->d_xq);
    }
    else
    {
        NB_CUDA(h, cudaMemcpyAsync(h->d_x + 3 * (size_t)a0, src, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, h->stream));
        k_x_to_grid<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_x, h->d_slot_of_atom, a0, a1, h->d_xq);
    }
    LAUNCH_CHECK(h);
    return 0;
}

extern "C" int b200nb_clear_outputs(b200nb_t* h)
{
    if (!h) return B200NB_ERR_ARG;
    if (!h->grid[0].valid && !h->grid_uploaded) return nb_fail(h, B200NB_ERR_STATE, "clear_outputs: put_on_grid first");
    cudaSetDevice(h->device);
    NB_CUDA(h, cudaMemsetAsync(h->d_f, 0, sizeof(float4) * (size_t)h->npad, h->stream));
    NB_CUDA(h, cudaMemsetAsync(h->d_fshift, 0, sizeof(float) * NB_OUT_COPIES * NB_FSHIFT_PITCH, h->stream));
    NB_CUDA(h, cudaMemsetAsync(h->d_energy, 0, sizeof(double) * NB_OUT_COPIES * 2, h->stream));
    return 0;
}

extern "C" int b200nb_launch_force(b200nb_t* h, int locality, int flags)
{
    NvtxRange nvtx_("b200nb_launch_force");
    if (!h) return B200NB_ERR_ARG;
    if (!h->have_list) return nb_fail(h, B200NB_ERR_STATE, "launch_force: no pair list");
    cudaSetDevice(h->device);
    for (int loc = 0; loc < 2; loc++)
        if (locality < 0 || locality == loc)
        {
            int rc = nb_launch_force_kernel(h, loc, flags);
            if (rc) return rc;
        }
    return 0;
}

extern "C" int b200nb_get_f(b200nb_t* h, float* f, int f_on_device, int accumulate, int a0, int a1)
{
    NvtxRange nvtx_("b200nb_get_f");
    if (!h || !f) return nb_fail(h, B200NB_ERR_ARG, "get_f: bad argument");
    if (!h->grid[0].valid) return nb_fail(h, B200NB_ERR_STATE, "get_f: put_on_grid first");
    if (a0 < 0 || a1 > h->natoms || a0 > a1) return nb_fail(h, B200NB_ERR_ARG, "get_f: bad atom range");
    if (a1 == a0) return 0;
    cudaSetDevice(h->device);
    const int n = a1 - a0;
    if (f_on_device)
    {
        k_f_from_grid<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_f, h->d_slot_of_atom, a0, a1, accumulate, f);
        LAUNCH_CHECK(h);
    }
    else
    {
        if (accumulate) return nb_fail(h, B200NB_ERR_ARG, "get_f: accumulate needs a device buffer");
        k_f_from_grid<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_f, h->d_slot_of_atom, a0, a1, 0, h->d_fout);
        LAUNCH_CHECK(h);
        NB_CUDA(h, cudaMemcpyAsync(f + 3 * (size_t)a0, h->d_fout + 3 * (size_t)a0, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
        NB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}

/* The force kernel spreads its per-entry shift-force and energy atomics over NB_OUT_COPIES replicas (thousands of
 * entries adding to the same 2 + 132 addresses serialise in L2); this sums the replicas: the device half of
 * gpu_reduce_staged_outputs (gpu_common.h:249-277). */
__global__ void k_reduce_outputs(const float* __restrict__ fshift, const double* __restrict__ energy, float* __restrict__ fshift_sum,
                                 double* __restrict__ energy_sum)
{
    const int t = threadIdx.x;
    if (t < B200NB_SHIFTS * 3)
    {
        float s = 0.f;
        for (int c = 0; c < NB_OUT_COPIES; c++) s += fshift[c * NB_FSHIFT_PITCH + t];
        fshift_sum[t] = s;
    }
    else if (t < B200NB_SHIFTS * 3 + 2)
    {
        double s = 0.0;
        for (int c = 0; c < NB_OUT_COPIES; c++) s += energy[2 * c + (t - B200NB_SHIFTS * 3)];
        energy_sum[t - B200NB_SHIFTS * 3] = s;
    }
}

extern "C" int b200nb_get_outputs(b200nb_t* h, float* fshift_host, double* energies_host)
{
    if (!h) return B200NB_ERR_ARG;
    cudaSetDevice(h->device);
    float  fs[B200NB_SHIFTS * 3];
    double e[2];
    k_reduce_outputs<<<1, 160, 0, h->stream>>>(h->d_fshift, h->d_energy, h->d_fshift_sum, h->d_energy_sum);
    LAUNCH_CHECK(h);
    NB_CUDA(h, cudaMemcpyAsync(fs, h->d_fshift_sum, sizeof(fs), cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaMemcpyAsync(e, h->d_energy_sum, sizeof(e), cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (fshift_host)
        for (int k = 0; k < B200NB_SHIFTS * 3; k++) fshift_host[k] += fs[k]; /* gpu_common.h:259-275 accumulates */
    if (energies_host)
    {
        energies_host[0] += e[0];
        energies_host[1] += e[1];
    }
    return 0;
}

__global__ void k_flush(float* p, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1.0f;
}

/* ---- one whole step in three launches -------------------------------------------------------------------
 * k_step_begin: coordinates (atom order) -> grid layout (nbnxn_gpu_x_to_nbat_x_kernel) fused with the clearing of f,
 *               fshift and the energies (gpu_clear_outputs, cuda/nbnxm_cuda_data_mgmt.cu:350-377);
 * k_force;
 * k_step_end:   grid-ordered f -> atom order (reduceKernel, mdlib/gpuforcereduction_impl.cu:70-104).
 * Both buffer kernels move the atom-order arrays as coalesced 16-byte vectors through shared memory, so x / f may be
 * device memory or PINNED HOST memory mapped into the device address space: in the host case the kernels read and write
 * it over PCIe themselves (no separate cudaMemcpyAsync, no staging buffer, two fewer dependent launches per step). */
#define NB_DD_SPIN_LIMIT_NS 10000000000ll /* a flag that does not arrive within 10 s raises the window's err word */

/* flags of the peer-memory halo windows (DdState, b200nb_internal.h): system-scope release / acquire */
__device__ __forceinline__ int ld_acquire_sys(const int* p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long globaltimer_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
/* Threads k < n of the CTA each wait until their flag (32 bytes apart, starting at `flags`) has reached *seq (the step number,
 * device-resident so that a captured CUDA graph of the step can be replayed) -- only the links with active[k] != 0 -- then the
 * CTA proceeds.  Bounded: a peer that never arrives raises *err instead of hanging the GPU. */
__device__ __forceinline__ void wait_flags(const unsigned char* flags, int n, const int* active_begin, const int* seq, int* err)
{
    if ((int)threadIdx.x < n && active_begin[threadIdx.x + 1] > active_begin[threadIdx.x])
    {
        const int*      flag = reinterpret_cast<const int*>(flags + 32 * threadIdx.x);
        const int       want = *reinterpret_cast<const volatile int*>(seq) + 1; /* the step in progress: see k_dd_step_end */
        const long long t0   = globaltimer_ns();
        while (ld_acquire_sys(flag) < want)
        {
            __nanosleep(64);
            if (globaltimer_ns() - t0 > NB_DD_SPIN_LIMIT_NS)
            {
                atomicExch(err, 1);
                break;
            }
        }
    }
    __syncthreads();
}
/* after this CTA's stores to the peers: the last CTA of the grid publishes the step number in the peers' flags, one thread
 * per link with traffic */
__device__ __forceinline__ void publish_flags(int* counter, int* const* peer_flags, int n, const int* active_begin, const int* seq)
{
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const int last = atomicAdd(counter, 1) == (int)gridDim.x - 1;
        if (last) *counter = 0;
        s_last = last;
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < n && active_begin[threadIdx.x + 1] > active_begin[threadIdx.x])
    {
        __threadfence_system();
        st_release_sys(peer_flags[threadIdx.x], *reinterpret_cast<const volatile int*>(seq) + 1);
    }
}

/* what k_step_begin does on top of its single-domain work when the step is domain-decomposed */
struct DdBegin
{
    int*       seq;       /* step counter (advanced by k_dd_step_end) */
    const int* send_atom; /* per send entry: home atom */
    const int* send_link; /* per send entry: link */
    int        nsend;
    int*       counter;
};

/* The L2 prefetch of the packed list pays on small systems, where the force kernel is a few waves long and its first
 * loads would otherwise come from HBM; a list of tens of MB belongs to a kernel that hides that latency by itself, and
 * streaming it twice per step would only cost bandwidth and time in k_step_begin. */
static inline size_t nb_list_prefetch_bytes(size_t bytes) { return bytes <= ((size_t)12 << 20) ? bytes : 0; }

struct PrefetchRange
{
    const char* p[4];
    size_t      bytes[4];
};
template<bool VEC>
__global__ void __launch_bounds__(256)
k_step_begin(const float* __restrict__ x, const int* __restrict__ slot_of_atom, int a0, int a1, float* __restrict__ xq,
             float4* __restrict__ f, int nclear, float* __restrict__ fshift, double* __restrict__ energy, PrefetchRange pf)
{
    __shared__ __align__(16) float sx[768];
    const int tid = threadIdx.x, t = blockIdx.x * 256 + tid;
    /* let the force kernel behind us start launching: its prologue only reads the list (it waits for this grid's completion,
     * griddepcontrol.wait, before it touches xq or f) */
    asm volatile("griddepcontrol.launch_dependents;");
    /* pull the read-only inputs of the force kernel that is about to run (packed list, LJ parameters) into L2 while this
     * kernel streams the coordinates: its prologue is a chain of dependent loads, ~3x shorter on L2 hits than from HBM */
#pragma unroll
    for (int r = 0; r < 4; r++)
        for (size_t o = (size_t)t * 128; o < pf.bytes[r]; o += (size_t)gridDim.x * 256 * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf.p[r] + o));
    if (t < nclear) f[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < NB_OUT_COPIES * NB_FSHIFT_PITCH) fshift[t] = 0.f;
    if (t < NB_OUT_COPIES * 2) energy[t] = 0.0;
    const int base = a0 + blockIdx.x * 256;
    const int nb   = min(256, a1 - base);
    if (nb > 0)
    {
        const float* src = x + 3 * (size_t)base;
        const int    nfl = nb * 3;
        if (VEC)
        {
            if (tid * 4 + 3 < nfl) *reinterpret_cast<float4*>(&sx[tid * 4]) = reinterpret_cast<const float4*>(src)[tid];
            else
                for (int k = tid * 4; k < nfl && k < tid * 4 + 4; k++) sx[k] = src[k];
        }
        else
            for (int k = tid; k < nfl; k += 256) sx[k] = src[k];
    }
    __syncthreads();
    if (tid < nb)
    {
        float* xb = xq + 4 * (size_t)slot_of_atom[base + tid];
        xb[0]     = sx[3 * tid];
        xb[1]     = sx[3 * tid + 1];
        xb[2]     = sx[3 * tid + 2];
    }
}

/* dd_move_x, sending side (packSendBufKernel, gpuhaloexchange_impl.cu:77-100) fused with the transfer: every send entry (home
 * atom, link) goes straight from the caller's coordinates into the destination's window, shifted when the link crosses a
 * periodic edge, then the flags.  Its own kernel at the head of the NON-LOCAL stream: it needs nothing but x, so it runs beside
 * k_step_begin, and the system-scope fences of the flag protocol (an NVLink round trip) stay off the path to the local force
 * kernel (fused into k_step_begin they delayed it by ~4 us at 128 k atoms per GPU). */
__global__ void __launch_bounds__(256)
k_dd_push_x(const float* __restrict__ x, DdBegin dd, const __grid_constant__ DdLinksDev L)
{
    const int t = blockIdx.x * 256 + threadIdx.x;
    for (int e = t; e < dd.nsend; e += (int)gridDim.x * 256)
    {
        const int    a = dd.send_atom[e], l = dd.send_link[e];
        float* const d = L.peer_recv_x[l] + 3 * (size_t)(e - L.send_off[l]);
        d[0]           = x[3 * (size_t)a] + L.shift[l][0];
        d[1]           = x[3 * (size_t)a + 1] + L.shift[l][1];
        d[2]           = x[3 * (size_t)a + 2] + L.shift[l][2];
    }
    publish_flags(dd.counter, L.peer_flag_x, L.nlinks, L.send_off, dd.seq);
}

template<bool VEC>
__global__ void __launch_bounds__(256)
k_step_end(const float4* __restrict__ fg, const int* __restrict__ slot_of_atom, int a0, int a1, float* __restrict__ f)
{
    __shared__ __align__(16) float sf[768];
    const int tid  = threadIdx.x;
    const int base = a0 + blockIdx.x * 256;
    const int nb   = min(256, a1 - base);
    if (nb <= 0) return;
    if (tid < nb)
    {
        const float4 v  = fg[slot_of_atom[base + tid]];
        sf[3 * tid]     = v.x;
        sf[3 * tid + 1] = v.y;
        sf[3 * tid + 2] = v.z;
    }
    __syncthreads();
    float*    dst = f + 3 * (size_t)base;
    const int nfl = nb * 3;
    if (VEC)
    {
        if (tid * 4 + 3 < nfl) reinterpret_cast<float4*>(dst)[tid] = *reinterpret_cast<const float4*>(&sf[tid * 4]);
        else
            for (int k = tid * 4; k < nfl && k < tid * 4 + 4; k++) dst[k] = sf[k];
    }
    else
        for (int k = tid; k < nfl; k += 256) dst[k] = sf[k];
}

static int launch_step(b200nb_context* h, const float* x_dev, int flags, float* f_dev, cudaEvent_t ev_force0 = nullptr,
                       cudaEvent_t ev_force1 = nullptr)
{
    const int n      = h->natoms;
    const int nclear = (int)h->cap_pad; /* grid slots + slack + the dummy atoms of the packed list at the end */
    const unsigned nb0 = (unsigned)((std::max(std::max(n, nclear), NB_OUT_COPIES * NB_FSHIFT_PITCH) + 255) / 256), nb1 = (unsigned)((n + 255) / 256);
    PrefetchRange pf{};
    {
        const PackedList& P = h->packed[0];
        pf.p[0]     = reinterpret_cast<const char*>(P.entries);
        pf.bytes[0] = sizeof(Entry) * (size_t)P.nentries;
        pf.p[1]     = reinterpret_cast<const char*>(P.ja);
        pf.bytes[1] = nb_list_prefetch_bytes(sizeof(int) * 16 * (size_t)P.nentries * P.row);
        pf.p[2]     = reinterpret_cast<const char*>(P.mask);
        pf.bytes[2] = nb_list_prefetch_bytes(sizeof(uint64_t) * (size_t)P.nentries * P.row);
        pf.p[3]     = reinterpret_cast<const char*>(h->comb_geom ? (const void*)h->d_lj : (const void*)h->d_atype);
        pf.bytes[3] = (h->comb_geom ? 8 : 4) * (size_t)h->npad;
    }
    if ((reinterpret_cast<uintptr_t>(x_dev) & 15) == 0)
        k_step_begin<true><<<nb0, 256, 0, h->stream>>>(x_dev, h->d_slot_of_atom, 0, n, h->d_xq, h->d_f, nclear, h->d_fshift, h->d_energy, pf);
    else
        k_step_begin<false><<<nb0, 256, 0, h->stream>>>(x_dev, h->d_slot_of_atom, 0, n, h->d_xq, h->d_f, nclear, h->d_fshift, h->d_energy, pf);
    LAUNCH_CHECK(h);
    int rc;
    if (ev_force0) NB_CUDA(h, cudaEventRecord(ev_force0, h->stream));
    if ((rc = b200nb_launch_force(h, -1, flags))) return rc;
    if (ev_force1) NB_CUDA(h, cudaEventRecord(ev_force1, h->stream));
    /* listed interactions and perturbed pairs, when made part of the step (b200nb_bonded_in_step, b200nb_fep_in_step): into the
     * same grid-order forces, before the un-sort */
    if ((rc = nb_bonded_enqueue_in_step(h, flags))) return rc;
    if ((rc = nb_fep_enqueue_in_step(h))) return rc;
    if ((reinterpret_cast<uintptr_t>(f_dev) & 15) == 0) k_step_end<true><<<nb1, 256, 0, h->stream>>>(h->d_f, h->d_slot_of_atom, 0, n, f_dev);
    else k_step_end<false><<<nb1, 256, 0, h->stream>>>(h->d_f, h->d_slot_of_atom, 0, n, f_dev);
    LAUNCH_CHECK(h);
    return 0;
}

static int run_step_graph(b200nb_context* h, int which, const float* x, float* f, int flags);

/* device-resident step (the reference's GPU buffer-ops path, mdlib/sim_util.cpp:1043-1108): asynchronous on the stream */
extern "C" int b200nb_step(b200nb_t* h, const float* x_dev, int flags, float* f_dev)
{
    NvtxRange nvtx_("b200nb_step");
    if (!h || !x_dev || !f_dev) return nb_fail(h, B200NB_ERR_ARG, "step: bad argument");
    if (!h->have_list) return nb_fail(h, B200NB_ERR_STATE, "step: no pair list");
    if (h->grid[1].valid) return nb_fail(h, B200NB_ERR_STATE, "step: single-domain call on a context with a halo grid");
    cudaSetDevice(h->device);
    return run_step_graph(h, 0, x_dev, f_dev, flags);
}

/* Times `niter` device-resident steps with CUDA events on the context's stream: the whole step and, inside it, the
 * force kernel alone (events recorded right before and after its launch), optionally with L2 flushed before each step. */
extern "C" int b200nb_time_step(b200nb_t* h, const float* x_dev, float* f_dev, int flags, int nwarm, int niter, int flush_l2,
                                float* ms_step_avg, float* ms_force_avg)
{
    if (!h || !x_dev || !f_dev || niter < 1) return nb_fail(h, B200NB_ERR_ARG, "time_step: bad argument");
    if (!h->have_list) return nb_fail(h, B200NB_ERR_STATE, "time_step: no pair list");
    cudaSetDevice(h->device);
    if (flush_l2 && !h->d_flush)
    {
        h->flush_bytes = (size_t)256 << 20; /* > 126 MB L2 */
        NB_CUDA(h, cudaMalloc((void**)&h->d_flush, h->flush_bytes));
    }
    cudaEvent_t ev[4];
    for (auto& e : ev) NB_CUDA(h, cudaEventCreate(&e));
    int rc;
    for (int i = 0; i < nwarm; i++)
        if ((rc = launch_step(h, x_dev, flags, f_dev))) return rc;
    double tstep = 0, tforce = 0;
    for (int i = 0; i < niter; i++)
    {
        if (flush_l2) k_flush<<<148 * 8, 256, 0, h->stream>>>(h->d_flush, h->flush_bytes / sizeof(float));
        NB_CUDA(h, cudaEventRecord(ev[0], h->stream));
        if ((rc = launch_step(h, x_dev, flags, f_dev, ev[1], ev[2]))) return rc;
        NB_CUDA(h, cudaEventRecord(ev[3], h->stream));
        NB_CUDA(h, cudaEventSynchronize(ev[3]));
        float a = 0, b = 0;
        NB_CUDA(h, cudaEventElapsedTime(&a, ev[0], ev[3]));
        NB_CUDA(h, cudaEventElapsedTime(&b, ev[1], ev[2]));
        tstep += a;
        tforce += b;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (ms_step_avg) *ms_step_avg = (float)(tstep / niter);
    if (ms_force_avg) *ms_force_avg = (float)(tforce / niter);
    return 0;
}

/* device-visible address of a host buffer: pinned (cudaHostAlloc / cudaHostRegister) memory is used in place */
static float* mapped_host_pointer(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type == cudaMemoryTypeHost && at.devicePointer) return static_cast<float*>(at.devicePointer);
    return nullptr;
}

extern "C" int b200nb_compute(b200nb_t* h, const float* x_host, int flags, float* f_host, float* fshift_host, double* energies_host)
{
    NvtxRange nvtx_("b200nb_compute");
    if (!h || !x_host || !f_host) return nb_fail(h, B200NB_ERR_ARG, "compute: bad argument");
    if (!h->have_list) return nb_fail(h, B200NB_ERR_STATE, "compute: no pair list");
    if (h->grid[1].valid) return nb_fail(h, B200NB_ERR_STATE, "compute: single-domain call on a context with a halo grid");
    cudaSetDevice(h->device);
    const size_t bytes = sizeof(float) * 3 * (size_t)h->natoms;
    /* the device-visible aliases are looked up on every call (a pointer-attribute query costs well under a microsecond): a buffer
     * the caller freed or unregistered, and whose address came back as different memory, must not be reached through a stale alias */
    h->map_x_host = x_host;
    h->map_f_host = f_host;
    h->map_x_dev  = mapped_host_pointer(x_host);
    h->map_f_dev  = mapped_host_pointer(f_host);
    const bool want_out = fshift_host || energies_host;
    /* pinned scratch: [x staging][f staging][fshift 135 floats][energies 2 doubles] */
    const size_t off_out = 2 * ((bytes + 255) & ~(size_t)255);
    if ((!h->map_x_dev || !h->map_f_dev || want_out) && ensure_pinned(h, off_out + 1024)) return B200NB_ERR_CUDA;
    const float* xd = h->map_x_dev;
    float*       fd = h->map_f_dev;
    if (!xd)
    {
        /* pageable coordinates: one CPU copy into our own pinned buffer, which the kernel then reads in place */
        memcpy(h->h_pinned, x_host, bytes);
        xd = mapped_host_pointer(h->h_pinned);
    }
    if (!fd) fd = mapped_host_pointer(h->h_pinned + off_out / 2);
    if (!xd || !fd) return nb_fail(h, B200NB_ERR_CUDA, "compute: pinned host memory is not mapped into the device address space");
    int rc;
    if (h->host_dma < 0)
    {
        const char* e = getenv("B200NB_HOST_DMA"); /* A/B switch: 1 = copy engines + staging, 0 = kernels access the host buffers */
        h->host_dma   = e ? (atoi(e) != 0) : B200NB_HOST_DMA_DEFAULT;
    }
    if (h->host_dma)
    {
        /* the copy engines need the host pointers themselves (pinned: the caller's, or our scratch) */
        const float* xh = h->map_x_dev ? x_host : reinterpret_cast<const float*>(h->h_pinned);
        float*       fh = h->map_f_dev ? f_host : reinterpret_cast<float*>(h->h_pinned + off_out / 2);
        if ((rc = run_step_graph(h, 2, xh, fh, flags))) return rc;
    }
    else if ((rc = run_step_graph(h, 0, xd, fd, flags))) return rc;
    float*  fs_pin = reinterpret_cast<float*>(h->h_pinned + off_out);
    double* e_pin  = reinterpret_cast<double*>(h->h_pinned + off_out + 640);
    if (want_out)
    {
        k_reduce_outputs<<<1, 160, 0, h->stream>>>(h->d_fshift, h->d_energy, h->d_fshift_sum, h->d_energy_sum);
        LAUNCH_CHECK(h);
        NB_CUDA(h, cudaMemcpyAsync(fs_pin, h->d_fshift_sum, sizeof(float) * B200NB_SHIFTS * 3, cudaMemcpyDeviceToHost, h->stream));
        NB_CUDA(h, cudaMemcpyAsync(e_pin, h->d_energy_sum, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stream));
    }
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!h->map_f_dev) memcpy(f_host, h->h_pinned + off_out / 2, bytes);
    if (fshift_host) memcpy(fshift_host, fs_pin, sizeof(float) * B200NB_SHIFTS * 3);
    if (energies_host)
    {
        energies_host[0] = e_pin[0];
        energies_host[1] = e_pin[1];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* halo pack / unpack                                                                                      */
/* ------------------------------------------------------------------------------------------------------ */
/* packSendBufKernel<usePBC> (domdec/gpuhaloexchange_impl.cu:77-100) */
__global__ void k_halo_pack(const float* __restrict__ x, const int* __restrict__ index, int n, float sx, float sy, float sz,
                            float* __restrict__ out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int a          = index[k];
    out[3 * k]     = x[3 * a] + sx;
    out[3 * k + 1] = x[3 * a + 1] + sy;
    out[3 * k + 2] = x[3 * a + 2] + sz;
}
/* unpackRecvBufKernel<accumulate=true> (domdec/gpuhaloexchange_impl.cu:108-131) */
__global__ void k_halo_unpack(float* __restrict__ f, const int* __restrict__ index, int n, const float* __restrict__ in)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int a = index[k];
    f[3 * a] += in[3 * k];
    f[3 * a + 1] += in[3 * k + 1];
    f[3 * a + 2] += in[3 * k + 2];
}

extern "C" int b200nb_halo_pack_x(b200nb_t* h, const float* x_dev, const int* index_dev, int n, const float shift[3], float* out_dev)
{
    if (!h || !x_dev || !index_dev || !out_dev || n < 0) return nb_fail(h, B200NB_ERR_ARG, "halo_pack_x: bad argument");
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    k_halo_pack<<<(n + 255) / 256, 256, 0, h->stream>>>(x_dev, index_dev, n, shift ? shift[0] : 0.f, shift ? shift[1] : 0.f,
                                                       shift ? shift[2] : 0.f, out_dev);
    LAUNCH_CHECK(h);
    return 0;
}

extern "C" int b200nb_halo_unpack_f(b200nb_t* h, float* f_dev, const int* index_dev, int n, const float* in_dev)
{
    if (!h || !f_dev || !index_dev || !in_dev || n < 0) return nb_fail(h, B200NB_ERR_ARG, "halo_unpack_f: bad argument");
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    k_halo_unpack<<<(n + 255) / 256, 256, 0, h->stream>>>(f_dev, index_dev, n, in_dev);
    LAUNCH_CHECK(h);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* domain-decomposed step over peer-memory halo windows (DdState, b200nb_internal.h)                      */
/* ------------------------------------------------------------------------------------------------------ */
/* dd_move_x, receiving side, fused with nbnxn_gpu_x_to_nbat_x for the halo grid: window -> grid layout, after the
 * coordinates of every link have landed */
__global__ void __launch_bounds__(256)
k_dd_recv_x(const unsigned char* __restrict__ window, const int* __restrict__ seq, const float* __restrict__ recv_x,
            const int* __restrict__ slot_of_atom, int nhome, int nhalo, float* __restrict__ xq, const __grid_constant__ DdLinksDev L)
{
    wait_flags(window + NB_DD_FLAG_X(0), L.nlinks, L.halo_off, seq, reinterpret_cast<int*>(const_cast<unsigned char*>(window) + NB_DD_ERR));
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= nhalo) return;
    float* xb = xq + 4 * (size_t)slot_of_atom[nhome + k];
    xb[0]     = __ldcg(recv_x + 3 * k); /* written by the peer: read at L2, never from a stale L1 line */
    xb[1]     = __ldcg(recv_x + 3 * k + 1);
    xb[2]     = __ldcg(recv_x + 3 * k + 2);
}

/* dd_move_f, sending side: the forces we computed on the halo atoms go into their owners' windows; forces on atoms that
 * arrived across a periodic edge also enter the shift force of that shift (domdec/domdec.cpp:426-458: the reference adds
 * them on the owner's side; the virial sums the shift forces over all ranks, so the side does not matter) */
__global__ void __launch_bounds__(256)
k_dd_push_f(const float4* __restrict__ fg, const int* __restrict__ slot_of_atom, int nhome, int nhalo, const unsigned char* __restrict__ halo_link,
            const int* __restrict__ seq, int* __restrict__ counter, float* __restrict__ fshift, const __grid_constant__ DdLinksDev L)
{
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < nhalo)
    {
        const float4 v = fg[slot_of_atom[nhome + k]];
        const int    l = halo_link[k];
        float* const d = L.peer_recv_f[l] + 3 * (size_t)(k - L.halo_off[l]);
        d[0]           = v.x;
        d[1]           = v.y;
        d[2]           = v.z;
        if (fshift && L.fshift_index[l] >= 0)
        {
            float* fs = fshift + (blockIdx.x & (NB_OUT_COPIES - 1)) * NB_FSHIFT_PITCH + 3 * L.fshift_index[l];
            atomicAdd(fs, v.x);
            atomicAdd(fs + 1, v.y);
            atomicAdd(fs + 2, v.z);
        }
    }
    publish_flags(counter, L.peer_flag_f, L.nlinks, L.halo_off, seq);
}

/* dd_move_f, receiving side (unpackRecvBufKernel<accumulate>, gpuhaloexchange_impl.cu:108-131) fused with the force
 * un-sort (reduceKernel): f_home[a] = f_grid[slot[a]] + the forces returned for every send entry that carried a */
template<bool VEC>
__global__ void __launch_bounds__(256)
k_dd_step_end(const float4* __restrict__ fg, const int* __restrict__ slot_of_atom, int nhome, const int* __restrict__ ent_off,
              const int* __restrict__ ent_idx, const unsigned char* __restrict__ window, int* __restrict__ seq, int* __restrict__ counter,
              int* __restrict__ host_err, const float* __restrict__ recv_f, float* __restrict__ f, const __grid_constant__ DdLinksDev L)
{
    __shared__ __align__(16) float sf[768];
    const int tid  = threadIdx.x;
    const int base = blockIdx.x * 256;
    const int nb   = min(256, nhome - base);
    /* the forces on the atoms we sent have landed in our window */
    wait_flags(window + NB_DD_FLAG_F(0), L.nlinks, L.send_off, seq, reinterpret_cast<int*>(const_cast<unsigned char*>(window) + NB_DD_ERR));
    /* The step counter: during step s the device-resident counter holds s - 1 and every kernel of the step uses counter + 1 (the
     * value the flags carry), whichever branch of the graph it runs on; this kernel is the last of the step, and the last of its
     * CTAs to get here -- every CTA has read the counter in wait_flags by then -- advances it for the next replay. */
    if (tid == 0)
    {
        __threadfence();
        if (atomicAdd(counter, 1) == (int)gridDim.x - 1)
        {
            *counter = 0;
            *seq     = *seq + 1;
            /* the time-out flag of this step's flag waits, mirrored into mapped host memory: b200nb_dd_status reads it there
             * without a device call (a 4-byte cudaMemcpy per step cost ~8 us of every end-to-end step) */
            const int e = *reinterpret_cast<const volatile int*>(window + NB_DD_ERR);
            if (e) *reinterpret_cast<volatile int*>(host_err) = e;
        }
    }
    if (nb <= 0) return;
    if (tid < nb)
    {
        float4    v  = fg[slot_of_atom[base + tid]];
        const int e0 = ent_off[base + tid], e1 = ent_off[base + tid + 1];
        for (int k = e0; k < e1; k++)
        {
            const int e = ent_idx[k];
            v.x += __ldcg(recv_f + 3 * (size_t)e);
            v.y += __ldcg(recv_f + 3 * (size_t)e + 1);
            v.z += __ldcg(recv_f + 3 * (size_t)e + 2);
        }
        sf[3 * tid]     = v.x;
        sf[3 * tid + 1] = v.y;
        sf[3 * tid + 2] = v.z;
    }
    __syncthreads();
    float*    dst = f + 3 * (size_t)base;
    const int nfl = nb * 3;
    if (VEC)
    {
        if (tid * 4 + 3 < nfl) reinterpret_cast<float4*>(dst)[tid] = *reinterpret_cast<const float4*>(&sf[tid * 4]);
        else
            for (int k = tid * 4; k < nfl && k < tid * 4 + 4; k++) dst[k] = sf[k];
    }
    else
        for (int k = tid; k < nfl; k += 256) dst[k] = sf[k];
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int b200nb_dd_create_window(b200nb_t* h, int max_halo, int max_send, void* ipc_handle_out, void** window_dev_out)
{
    if (h && (h->box_off[0] != 0.f || h->box_off[1] != 0.f || h->box_off[2] != 0.f))
        return nb_fail(h, B200NB_ERR_ARG, "dd_create_window: domain decomposition of a triclinic cell is not supported");
    if (!h || max_halo < 0 || max_send < 0) return nb_fail(h, B200NB_ERR_ARG, "dd_create_window: bad argument");
    cudaSetDevice(h->device);
    DdState& D = h->dd;
    if (D.window) return nb_fail(h, B200NB_ERR_STATE, "dd_create_window: window exists (peers hold its handle)");
    D.max_halo     = max_halo;
    D.max_send     = max_send;
    D.off_recv_x   = NB_DD_DATA;
    D.off_recv_f   = NB_DD_DATA + align256(sizeof(float) * 3 * (size_t)max_halo);
    D.window_bytes = D.off_recv_f + align256(sizeof(float) * 3 * (size_t)max_send);
    NB_CUDA(h, cudaMalloc((void**)&D.window, D.window_bytes));
    NB_CUDA(h, cudaMemset(D.window, 0, D.window_bytes));
    NB_CUDA(h, cudaMalloc((void**)&D.d_count, sizeof(int) * 4));
    NB_CUDA(h, cudaMemset(D.d_count, 0, sizeof(int) * 4));
    D.d_seq = D.d_count + 2; /* the step counter the flags carry */
    {
        int lo = 0, hi = 0;
        NB_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi)); /* hi = numerically lowest = highest priority */
        if (getenv("B200NB_DD_PRIO") && atoi(getenv("B200NB_DD_PRIO")) == 0) hi = lo; /* A/B switch for profiles/ */
        D.prio_high = hi;
        NB_CUDA(h, cudaStreamCreateWithPriority(&D.stream_nl, cudaStreamNonBlocking, hi));
        NB_CUDA(h, cudaStreamCreateWithPriority(&D.stream_px, cudaStreamNonBlocking, hi));
        NB_CUDA(h, cudaHostAlloc((void**)&D.h_err, 64, cudaHostAllocMapped));
        *D.h_err = 0;
        NB_CUDA(h, cudaHostGetDevicePointer((void**)&D.d_host_err, D.h_err, 0));
        if (const char* e = getenv("B200NB_DD_PUSH_INLINE")) D.push_inline = atoi(e) != 0;
        NB_CUDA(h, cudaEventCreateWithFlags(&D.ev_px_done, cudaEventDisableTiming));
        NB_CUDA(h, cudaEventCreateWithFlags(&D.ev_begin, cudaEventDisableTiming));
        NB_CUDA(h, cudaEventCreateWithFlags(&D.ev_start, cudaEventDisableTiming));
        NB_CUDA(h, cudaEventCreateWithFlags(&D.ev_nl_done, cudaEventDisableTiming));
    }
    if (ipc_handle_out)
    {
        cudaIpcMemHandle_t hd;
        NB_CUDA(h, cudaIpcGetMemHandle(&hd, D.window));
        memcpy(ipc_handle_out, &hd, sizeof(hd));
    }
    if (window_dev_out) *window_dev_out = D.window;
    return 0;
}

/* Opens a neighbour's window as peer `peer` (0 .. NB_DD_MAX_PEERS-1; with b200nb_dd_set_plan: 0 = the -x neighbour, which
 * receives our halo coordinates, 1 = the +x neighbour, which receives the forces on its atoms).  Either an IPC handle from
 * another process, or the window's device pointer when the peer lives in this process.  A neighbour reached over several
 * links is opened ONCE (a CUDA IPC handle maps once per process).  peer_max_halo: the max_halo the peer created its window
 * with (fixes where its recv_f block starts). */
extern "C" int b200nb_dd_open_peer(b200nb_t* h, int peer, const void* ipc_handle, void* same_process_window, int peer_max_halo)
{
    if (!h || peer < 0 || peer >= NB_DD_MAX_PEERS || (!ipc_handle && !same_process_window) || peer_max_halo < 0)
        return nb_fail(h, B200NB_ERR_ARG, "dd_open_peer: bad argument");
    cudaSetDevice(h->device);
    DdState& D = h->dd;
    if (D.peer[peer] && D.peer_is_ipc[peer]) cudaIpcCloseMemHandle(D.peer[peer]);
    D.peer[peer] = nullptr;
    if (same_process_window)
    {
        D.peer[peer]        = static_cast<unsigned char*>(same_process_window);
        D.peer_is_ipc[peer] = false;
    }
    else
    {
        cudaIpcMemHandle_t hd;
        memcpy(&hd, ipc_handle, sizeof(hd));
        void* p = nullptr;
        NB_CUDA(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        D.peer[peer]        = static_cast<unsigned char*>(p);
        D.peer_is_ipc[peer] = true;
    }
    D.peer_off_recv_f[peer] = NB_DD_DATA + align256(sizeof(float) * 3 * (size_t)peer_max_halo);
    h->generation++;
    return 0;
}

/* The halo plan of the current pair-search interval, general form: local atoms [0, nhome) are home, [nhome, nhome + nhalo) halo,
 * in link order (all atoms received over link 0, then link 1, ...).  See b200nb_dd_link_t in b200nb.h. */
extern "C" int b200nb_dd_set_links(b200nb_t* h, int nhome, int nhalo, int nlinks, const b200nb_dd_link_t* links)
{
    NvtxRange nvtx_("b200nb_dd_set_links");
    if (!h || nhome < 0 || nhalo < 0 || nlinks < 0 || nlinks > NB_DD_MAX_LINKS || (nlinks && !links))
        return nb_fail(h, B200NB_ERR_ARG, "dd_set_links: bad argument");
    cudaSetDevice(h->device);
    DdState& D = h->dd;
    if (!D.window) return nb_fail(h, B200NB_ERR_STATE, "dd_set_links: create the window first");
    if (nhome + nhalo != h->natoms) return nb_fail(h, B200NB_ERR_ARG, "dd_set_links: nhome + nhalo != natoms");
    DdLinksDev L{};
    L.nlinks = nlinks;
    long long nsend = 0, nrecv = 0;
    for (int k = 0; k < nlinks; k++)
    {
        const b200nb_dd_link_t& K = links[k];
        if (K.nsend < 0 || K.nrecv < 0 || (K.nsend && !K.send_idx_host)) return nb_fail(h, B200NB_ERR_ARG, "dd_set_links: bad link");
        if (K.nsend && (K.send_peer < 0 || K.send_peer >= NB_DD_MAX_PEERS || !D.peer[K.send_peer]))
            return nb_fail(h, B200NB_ERR_STATE, "dd_set_links: the destination window of a link is not open");
        if (K.nrecv && (K.recv_peer < 0 || K.recv_peer >= NB_DD_MAX_PEERS || !D.peer[K.recv_peer]))
            return nb_fail(h, B200NB_ERR_STATE, "dd_set_links: the source window of a link is not open");
        if (K.peer_halo_offset < 0 || K.peer_entry_offset < 0 || K.fshift_index >= B200NB_SHIFTS)
            return nb_fail(h, B200NB_ERR_ARG, "dd_set_links: bad offset / shift index");
        L.send_off[k] = (int)nsend;
        L.halo_off[k] = (int)nrecv;
        nsend += K.nsend;
        nrecv += K.nrecv;
        for (int d = 0; d < 3; d++) L.shift[k][d] = K.shift[d];
        L.fshift_index[k] = K.nrecv ? K.fshift_index : -1;
        if (K.nsend)
        {
            unsigned char* w = D.peer[K.send_peer];
            L.peer_recv_x[k] = reinterpret_cast<float*>(w + NB_DD_DATA) + 3 * (size_t)K.peer_halo_offset;
            L.peer_flag_x[k] = reinterpret_cast<int*>(w + NB_DD_FLAG_X(k));
        }
        if (K.nrecv)
        {
            unsigned char* w = D.peer[K.recv_peer];
            L.peer_recv_f[k] = reinterpret_cast<float*>(w + D.peer_off_recv_f[K.recv_peer]) + 3 * (size_t)K.peer_entry_offset;
            L.peer_flag_f[k] = reinterpret_cast<int*>(w + NB_DD_FLAG_F(k));
        }
    }
    for (int k = nlinks; k <= NB_DD_MAX_LINKS; k++) L.send_off[k] = (int)nsend, L.halo_off[k] = (int)nrecv;
    if (nrecv != nhalo) return nb_fail(h, B200NB_ERR_ARG, "dd_set_links: the links' receive counts do not add up to nhalo");
    if (nhalo > D.max_halo || nsend > D.max_send) return nb_fail(h, B200NB_ERR_CAPACITY, "dd_set_links: halo larger than the window");
    /* per entry: atom and link; per home atom: the entries that carry it (CSR); per halo atom: its link */
    std::vector<int>           atom((size_t)std::max<long long>(nsend, 1)), link((size_t)std::max<long long>(nsend, 1));
    std::vector<int>           off((size_t)nhome + 2, 0), idx((size_t)std::max<long long>(nsend, 1));
    std::vector<unsigned char> hl((size_t)std::max(nhalo, 1));
    for (int k = 0, e = 0; k < nlinks; k++)
        for (int p = 0; p < links[k].nsend; p++, e++)
        {
            const int a = links[k].send_idx_host[p];
            if (a < 0 || a >= nhome) return nb_fail(h, B200NB_ERR_ARG, "dd_set_links: send index out of range");
            atom[e] = a;
            link[e] = k;
            off[a + 1]++;
        }
    for (int a = 0; a < nhome; a++) off[a + 1] += off[a];
    {
        std::vector<int> cur(off.begin(), off.end() - 1);
        for (int e = 0; e < (int)nsend; e++) idx[cur[atom[e]]++] = e;
    }
    for (int k = 0; k < nlinks; k++)
        for (int p = 0; p < links[k].nrecv; p++) hl[(size_t)L.halo_off[k] + p] = (unsigned char)k;
    auto grow = [&](void** ptr, size_t* cap, size_t need, size_t elem) -> int {
        if (*ptr && need <= *cap) return 0;
        cudaFree(*ptr);
        *ptr = nullptr;
        *cap = need + need / 4 + 64;
        return cudaMalloc(ptr, *cap * elem) != cudaSuccess;
    };
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    size_t c1 = D.cap_send, c2 = D.cap_send, c3 = D.cap_send;
    if (grow((void**)&D.d_send_atom, &c1, (size_t)nsend + 1, sizeof(int)) || grow((void**)&D.d_send_link, &c2, (size_t)nsend + 1, sizeof(int))
        || grow((void**)&D.d_ent_idx, &c3, (size_t)nsend + 1, sizeof(int)))
        return nb_fail(h, B200NB_ERR_CUDA, "dd_set_links: device allocation failed");
    D.cap_send = c1;
    if (grow((void**)&D.d_ent_off, &D.cap_home, (size_t)nhome + 2, sizeof(int)) || grow((void**)&D.d_halo_link, &D.cap_halo, (size_t)nhalo + 1, 1))
        return nb_fail(h, B200NB_ERR_CUDA, "dd_set_links: device allocation failed");
    if (nsend)
    {
        NB_CUDA(h, cudaMemcpy(D.d_send_atom, atom.data(), sizeof(int) * nsend, cudaMemcpyHostToDevice));
        NB_CUDA(h, cudaMemcpy(D.d_send_link, link.data(), sizeof(int) * nsend, cudaMemcpyHostToDevice));
        NB_CUDA(h, cudaMemcpy(D.d_ent_idx, idx.data(), sizeof(int) * nsend, cudaMemcpyHostToDevice));
    }
    NB_CUDA(h, cudaMemcpy(D.d_ent_off, off.data(), sizeof(int) * ((size_t)nhome + 1), cudaMemcpyHostToDevice));
    if (nhalo) NB_CUDA(h, cudaMemcpy(D.d_halo_link, hl.data(), (size_t)nhalo, cudaMemcpyHostToDevice));
    D.links = L;
    D.nhome = nhome;
    D.nhalo = nhalo;
    D.nsend = (int)nsend;
    D.have_plan = true;
    h->generation++;
    return 0;
}

/* The 1-D (x-slab) form of b200nb_dd_set_links: one link; peer 0 = the -x neighbour (gets our halo coordinates), peer 1 = the +x
 * neighbour (owns our halo atoms).  send_idx_host: the nsend home atoms we send, `shift` added to them; halo_fshift_index: shift-
 * force slot that also receives the forces WE compute on our halo atoms when those arrived across the periodic edge, else -1. */
extern "C" int b200nb_dd_set_plan(b200nb_t* h, int nhome, int nhalo, const int* send_idx_host, int nsend, const float shift[3],
                                  int halo_fshift_index)
{
    if (!h || nsend < 0 || nhalo < 0) return nb_fail(h, B200NB_ERR_ARG, "dd_set_plan: bad argument");
    b200nb_dd_link_t K{};
    K.send_peer     = 0;
    K.nsend         = nsend;
    K.send_idx_host = send_idx_host;
    for (int d = 0; d < 3; d++) K.shift[d] = shift ? shift[d] : 0.f;
    K.peer_halo_offset  = 0;
    K.recv_peer         = 1;
    K.nrecv             = nhalo;
    K.peer_entry_offset = 0;
    K.fshift_index      = halo_fshift_index;
    return b200nb_dd_set_links(h, nhome, nhalo, 1, &K);
}

/* While a step is being captured: remember the graph node the last launch on the non-local stream created, so that its
 * priority can be set explicitly afterwards (a captured kernel node does not inherit the priority of the stream it was
 * captured from: measured, the non-local chain then queues behind the local kernel). */
static void tag_nonlocal_node(b200nb_context* h, cudaStream_t stream = nullptr)
{
    if (!h->capturing) return;
    if (!stream) stream = h->dd.stream_nl;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    const cudaGraphNode_t*  deps = nullptr;
    size_t                  nd = 0;
    if (cudaStreamGetCaptureInfo_v2(stream, &st, nullptr, nullptr, &deps, &nd) == cudaSuccess && st == cudaStreamCaptureStatusActive)
        for (size_t k = 0; k < nd; k++) h->nl_nodes.push_back(deps[k]);
}

/* the launches of one decomposed step; x_home / f_home are device-visible addresses */
static int launch_dd_step(b200nb_context* h, const float* x_home, float* f_home, int flags)
{
    DdState&  D = h->dd;
    const int n = D.nhome, nclear = (int)h->cap_pad;
    const unsigned nb0 = (unsigned)((std::max(std::max(n, nclear), NB_OUT_COPIES * NB_FSHIFT_PITCH) + 255) / 256);
    PrefetchRange pf{};
    {
        const PackedList& P = h->packed[0];
        pf.p[0]     = reinterpret_cast<const char*>(P.entries);
        pf.bytes[0] = sizeof(Entry) * (size_t)P.nentries;
        pf.p[1]     = reinterpret_cast<const char*>(P.ja);
        pf.bytes[1] = nb_list_prefetch_bytes(sizeof(int) * 16 * (size_t)P.nentries * P.row);
        pf.p[2]     = reinterpret_cast<const char*>(P.mask);
        pf.bytes[2] = nb_list_prefetch_bytes(sizeof(uint64_t) * (size_t)P.nentries * P.row);
        pf.p[3]     = reinterpret_cast<const char*>(h->comb_geom ? (const void*)h->d_lj : (const void*)h->d_atype);
        pf.bytes[3] = (h->comb_geom ? 8 : 4) * (size_t)h->npad;
    }
    DdBegin B{};
    B.seq       = D.d_seq;
    B.send_atom = D.d_send_atom;
    B.send_link = D.d_send_link;
    B.nsend     = D.nsend;
    B.counter   = D.d_count;
    /* 1. the halo x goes out first, on a branch of its own (high priority): it needs only the caller's coordinates, nothing in
     * this step waits for it (the consumers are the neighbours), and it never waits itself -- so it can share a hardware queue
     * with anything without holding it up */
    cudaStream_t snl = D.stream_nl;
    const unsigned npx = (unsigned)std::max(1, std::min((D.nsend + 255) / 256, 2 * 148));
    if (D.push_inline)
    {
        /* B200NB_DD_PUSH_INLINE=1: the push in front of k_step_begin on the main stream.  For MANY ranks as threads of one
         * process on one GPU (tests): their graphs' branches share that device's hardware queues, and a push queued behind
         * another rank's flag wait would never run.  One process per GPU (production) has its queues to itself. */
        k_dd_push_x<<<npx, 256, 0, h->stream>>>(x_home, B, D.links);
        LAUNCH_CHECK(h);
    }
    else
    {
        NB_CUDA(h, cudaEventRecord(D.ev_start, h->stream));
        NB_CUDA(h, cudaStreamWaitEvent(D.stream_px, D.ev_start, 0));
        k_dd_push_x<<<npx, 256, 0, D.stream_px>>>(x_home, B, D.links);
        LAUNCH_CHECK(h);
        tag_nonlocal_node(h, D.stream_px);
        NB_CUDA(h, cudaEventRecord(D.ev_px_done, D.stream_px));
    }
    /* 2. home x -> grid layout, outputs cleared; then the local kernel on the main stream, the halo chain on the non-local one */
    if ((reinterpret_cast<uintptr_t>(x_home) & 15) == 0)
        k_step_begin<true><<<nb0, 256, 0, h->stream>>>(x_home, h->d_slot_of_atom, 0, n, h->d_xq, h->d_f, nclear, h->d_fshift, h->d_energy, pf);
    else
        k_step_begin<false><<<nb0, 256, 0, h->stream>>>(x_home, h->d_slot_of_atom, 0, n, h->d_xq, h->d_f, nclear, h->d_fshift, h->d_energy, pf);
    LAUNCH_CHECK(h);
    NB_CUDA(h, cudaEventRecord(D.ev_begin, h->stream));
    NB_CUDA(h, cudaStreamWaitEvent(snl, D.ev_begin, 0));
    int rc;
    if ((rc = nb_launch_force_kernel(h, 0, flags))) return rc;
    if (D.nhalo)
    {
        k_dd_recv_x<<<(unsigned)std::max(1, (D.nhalo + 255) / 256), 256, 0, snl>>>(D.window, D.d_seq, reinterpret_cast<const float*>(D.window + D.off_recv_x),
                                                                                 h->d_slot_of_atom, D.nhome, D.nhalo, h->d_xq, D.links);
        LAUNCH_CHECK(h);
        tag_nonlocal_node(h);
        cudaStream_t keep = h->stream;
        h->stream         = snl; /* the non-local kernel goes to the non-local stream */
        rc                = nb_launch_force_kernel(h, 1, flags);
        h->stream         = keep;
        if (rc) return rc;
        tag_nonlocal_node(h);
        k_dd_push_f<<<(unsigned)std::max(1, (D.nhalo + 255) / 256), 256, 0, snl>>>(h->d_f, h->d_slot_of_atom, D.nhome, D.nhalo, D.d_halo_link, D.d_seq,
                                                                                 D.d_count + 1, (flags & B200NB_FLAG_VIRIAL) ? h->d_fshift : nullptr, D.links);
        LAUNCH_CHECK(h);
        tag_nonlocal_node(h);
    }
    NB_CUDA(h, cudaEventRecord(D.ev_nl_done, snl));
    NB_CUDA(h, cudaStreamWaitEvent(h->stream, D.ev_nl_done, 0));
    if (!D.push_inline) NB_CUDA(h, cudaStreamWaitEvent(h->stream, D.ev_px_done, 0));
    /* 3. wait for the forces on the atoms we sent, add them, forces -> atom order */
    const unsigned nb1 = (unsigned)std::max(1, (n + 255) / 256);
    if ((reinterpret_cast<uintptr_t>(f_home) & 15) == 0)
        k_dd_step_end<true><<<nb1, 256, 0, h->stream>>>(h->d_f, h->d_slot_of_atom, n, D.d_ent_off, D.d_ent_idx, D.window, D.d_seq, D.d_count + 3,
                                                       D.d_host_err, reinterpret_cast<const float*>(D.window + D.off_recv_f), f_home, D.links);
    else
        k_dd_step_end<false><<<nb1, 256, 0, h->stream>>>(h->d_f, h->d_slot_of_atom, n, D.d_ent_off, D.d_ent_idx, D.window, D.d_seq, D.d_count + 3,
                                                        D.d_host_err, reinterpret_cast<const float*>(D.window + D.off_recv_f), f_home, D.links);
    LAUNCH_CHECK(h);
    return 0;
}

/* Replays (capturing it first if needed) the CUDA graph of one step: the launches depend only on the buffers, the flags and
 * the current list / plan (`generation`), so a step costs ONE graph launch: no per-kernel launch gaps on the critical path.
 * Falls back to direct launches when the stream cannot be captured (e.g. the caller is itself capturing). */
static int run_step_graph(b200nb_context* h, int which, const float* x, float* f, int flags)
{
    StepGraph& G = h->graph[which];
    auto direct = [&]() -> int {
        if (which == 0) return launch_step(h, x, flags, f);
        if (which == 1) return launch_dd_step(h, x, f, flags);
        /* which == 2: host step through the copy engines: x / f are PINNED HOST buffers, staged through d_x / d_fout */
        const size_t bytes = sizeof(float) * 3 * (size_t)h->natoms;
        NB_CUDA(h, cudaMemcpyAsync(h->d_x, x, bytes, cudaMemcpyHostToDevice, h->stream));
        int rc = launch_step(h, h->d_x, flags, h->d_fout);
        if (rc) return rc;
        NB_CUDA(h, cudaMemcpyAsync(f, h->d_fout, bytes, cudaMemcpyDeviceToHost, h->stream));
        return 0;
    };
    if (const char* e = getenv("B200NB_GRAPHS")) /* A/B switch for profiles/: 0 = direct launches */
        if (atoi(e) == 0) h->use_graphs = false;
    if (const char* e = getenv("B200NB_PDL")) /* A/B switch: 0 = no programmatic dependent launch of the force kernel */
        h->use_pdl = atoi(e) != 0;
    if (!h->use_graphs) return direct();
    if (!G.exec || G.x != x || G.f != f || G.flags != flags || G.generation != h->generation || G.stream != h->stream)
    {
        if (G.exec) cudaGraphExecDestroy(G.exec);
        G.exec = nullptr;
        const long long l0 = h->nlaunches;
        cudaGraph_t     graph = nullptr;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        {
            cudaGetLastError();
            h->use_graphs = false;
            return direct();
        }
        h->nl_nodes.clear();
        h->capturing   = true;
        const int   rc = direct();
        h->capturing   = false;
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (ce == cudaSuccess && graph)
        {
            cudaKernelNodeAttrValue v{};
            v.priority = h->dd.prio_high;
            for (cudaGraphNode_t nd : h->nl_nodes)
            {
                cudaGraphNodeType ty;
                if (cudaGraphNodeGetType(nd, &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel)
                    cudaGraphKernelNodeSetAttribute(nd, cudaKernelNodeAttributePriority, &v);
            }
            cudaGetLastError();
        }
        /* UseNodePriority: run with the per-node priorities (the non-local chain's) instead of the launch stream's for all */
        if (rc || ce != cudaSuccess || !graph
            || cudaGraphInstantiateWithFlags(&G.exec, graph, which == 1 ? cudaGraphInstantiateFlagUseNodePriority : 0) != cudaSuccess)
        {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            G.exec        = nullptr;
            h->use_graphs = false;
            h->nlaunches  = l0;
            return rc ? rc : direct();
        }
        cudaGraphDestroy(graph);
        G.nkernels   = (int)(h->nlaunches - l0);
        h->nlaunches = l0;
        G.x = x, G.f = f, G.flags = flags, G.generation = h->generation, G.stream = h->stream;
    }
    NB_CUDA(h, cudaGraphLaunch(G.exec, h->stream));
    h->nlaunches += G.nkernels;
    return 0;
}

/* One step of the decomposed calculation, asynchronous on the context's stream (the nonbonded part of do_force,
 * mdlib/sim_util.cpp:1388-1902): home x -> grid layout + clear + push halo x to the -x neighbour | local kernel || wait for
 * our halo x, -> grid layout | non-local kernel | push halo forces to the +x neighbour || wait for the forces on the atoms we
 * sent, add, un-sort.  x_home / f_home: nhome*3 floats, device or pinned host memory. */
extern "C" int b200nb_dd_step(b200nb_t* h, const float* x_home, float* f_home, int flags)
{
    NvtxRange nvtx_("b200nb_dd_step");
    if (!h || !x_home || !f_home) return nb_fail(h, B200NB_ERR_ARG, "dd_step: bad argument");
    DdState& D = h->dd;
    if (!h->have_list || !D.have_plan) return nb_fail(h, B200NB_ERR_STATE, "dd_step: needs a pair list and a halo plan");
    cudaSetDevice(h->device);
    {
        /* pinned host buffers are used in place through their device-visible address; device pointers pass through.  Queried on
         * every call: see b200nb_compute */
        cudaPointerAttributes ax, af;
        if (cudaPointerGetAttributes(&ax, x_home) != cudaSuccess || cudaPointerGetAttributes(&af, f_home) != cudaSuccess
            || !ax.devicePointer || !af.devicePointer)
        {
            cudaGetLastError();
            return nb_fail(h, B200NB_ERR_ARG, "dd_step: x_home / f_home must be device or pinned host memory");
        }
        h->map_x_host = x_home;
        h->map_f_host = f_home;
        h->map_x_dev  = static_cast<float*>(ax.devicePointer);
        h->map_f_dev  = static_cast<float*>(af.devicePointer);
    }
    return run_step_graph(h, 1, h->map_x_dev, h->map_f_dev, flags);
}

/* after synchronising: B200NB_ERR_STATE when a halo flag did not arrive in time during any step since the last call */
extern "C" int b200nb_dd_status(b200nb_t* h)
{
    if (!h) return B200NB_ERR_ARG;
    DdState& D = h->dd;
    if (!D.window) return 0;
    cudaSetDevice(h->device);
    /* the last kernel of every decomposed step mirrors the window's time-out flag into mapped host memory */
    const int e = D.h_err ? *reinterpret_cast<volatile int*>(D.h_err) : 0;
    if (e)
    {
        *reinterpret_cast<volatile int*>(D.h_err) = 0;
        cudaMemset(D.window + NB_DD_ERR, 0, sizeof(int));
        return nb_fail(h, B200NB_ERR_STATE, "dd_step: a halo exchange flag did not arrive within 10 s (peer stalled or not stepping)");
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* introspection                                                                                           */
/* ------------------------------------------------------------------------------------------------------ */
/* what the force kernel computes on the packed list, in tiles of 8 x 8 pair lanes: a warp runs the half-entries 2w and 2w + 1
 * in lockstep for the longer one's steps, 128 pair lanes (= 2 tiles) per step */
__global__ void k_count_packed(const Entry* __restrict__ e, long long nwarps, long long* out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long v = 0;
    if (i < nwarps) v = 2 * max(e[2 * i].end - e[2 * i].start, e[2 * i + 1].end - e[2 * i + 1].start);
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd((unsigned long long*)out, (unsigned long long)v);
}

__global__ void k_count_tiles(const Entry* __restrict__ e, long long n, long long* out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long v = 0;
    if (i < n) v = e[i].end - e[i].start;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd((unsigned long long*)out, (unsigned long long)v);
}

extern "C" int b200nb_get_stats(b200nb_t* h, b200nb_stats_t* out)
{
    if (!h || !out) return B200NB_ERR_ARG;
    cudaSetDevice(h->device);
    memset(out, 0, sizeof(*out));
    out->natoms         = h->natoms;
    out->natoms_padded  = h->npad;
    out->nclusters      = h->npad / 8;
    out->ncx            = h->grid[0].ncx;
    out->ncy            = h->grid[0].ncy;
    out->comb_geometric = h->comb_geom ? 1 : 0;
    if (h->have_list)
    {
        if (refresh_inner_list(h)) return B200NB_ERR_CUDA;
        out->nentries_nonlocal = h->packed[1].nentries;
        for (int l = 0; l < 2; l++)
        {
            out->ntiles_outer += h->outer[l].ntiles;
            out->nentries += h->outer[l].nentries;
            if (h->inner_is_outer) out->ntiles_inner += h->outer[l].ntiles;
            else if (h->inner[l].nentries)
            {
                long long v = 0;
                NB_CUDA(h, cudaMemsetAsync(h->d_counter + 4, 0, sizeof(long long), h->stream));
                k_count_tiles<<<(unsigned)((h->inner[l].nentries + 255) / 256), 256, 0, h->stream>>>(h->inner[l].entries, h->inner[l].nentries,
                                                                                                   h->d_counter + 4);
                LAUNCH_CHECK(h);
                NB_CUDA(h, cudaMemcpyAsync(&v, h->d_counter + 4, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
                NB_CUDA(h, cudaStreamSynchronize(h->stream));
                out->ntiles_inner += v;
            }
        }
    }
    if (h->have_list)
        for (int l = 0; l < 2; l++)
            if (h->packed[l].nentries)
            {
                long long v = 0;
                NB_CUDA(h, cudaMemsetAsync(h->d_counter + 4, 0, sizeof(long long), h->stream));
                k_count_packed<<<(unsigned)((h->packed[l].nentries / 2 + 255) / 256), 256, 0, h->stream>>>(h->packed[l].entries,
                                                                                                         h->packed[l].nentries / 2, h->d_counter + 4);
                LAUNCH_CHECK(h);
                NB_CUDA(h, cudaMemcpyAsync(&v, h->d_counter + 4, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
                NB_CUDA(h, cudaStreamSynchronize(h->stream));
                out->ntiles_packed += v;
            }
    out->nlaunches = h->nlaunches;
    return 0;
}

/* one line about the context for logs: the role of the reference's "Using a ... Verlet scheme / GPU info" setup lines
 * (nbnxm_setup.cpp:180-260, mdrunutility/printhardware) */
extern "C" int b200nb_describe(b200nb_t* h, char* buf, int cap)
{
    if (!h || !buf || cap < 1) return B200NB_ERR_ARG;
    cudaDeviceProp pr{};
    cudaGetDeviceProperties(&pr, h->device);
    const char* eel = h->dp.eeltype == B200NB_EEL_EWALD ? (h->dp.ewald_tab ? "Ewald(tab)" : "Ewald(analytical)") : (h->dp.k_rf != 0.f ? "RF" : "cut-off");
    const long long ne0 = h->packed[0].nentries, ne1 = h->packed[1].nentries;
    snprintf(buf, (size_t)cap,
             "b200nb: %s (sm_%d%d, %d SMs) device %d | %d atoms in %d slots, grid %d x %d columns | rc %.3g rlist %.3g/%.3g | LJ %s + %s | "
             "list: %lld + %lld cluster pairs -> %lld + %lld half-entries, <= %d cluster pairs per entry | force kernel: 2 half-entries per "
             "warp, 1 warp per CTA, %s | %s%s",
             pr.name, pr.major, pr.minor, pr.multiProcessorCount, h->device, h->natoms, h->npad, h->grid[0].ncx, h->grid[0].ncy,
             h->have_params ? sqrtf(h->dp.rc2) : 0.f, h->have_params ? sqrtf(h->dp.rlist_outer2) : 0.f, h->have_params ? sqrtf(h->dp.rlist_inner2) : 0.f,
             h->comb_geom ? "geometric" : "type table", eel, h->outer[0].ntiles, h->outer[1].ntiles, ne0, ne1, h->max_tiles,
             h->use_pdl ? "programmatic dependent launch" : "serialized launch", h->use_graphs ? "step = 1 CUDA graph" : "direct launches",
             h->dd.have_plan ? " | halo over peer-memory windows" : "");
    return 0;
}

extern "C" int b200nb_get_grid_order(b200nb_t* h, int* atom_index_host, int cap)
{
    if (!h || !atom_index_host) return B200NB_ERR_ARG;
    if (!h->grid[0].valid) return nb_fail(h, B200NB_ERR_STATE, "get_grid_order: put_on_grid first");
    if (cap < h->npad) return nb_fail(h, B200NB_ERR_CAPACITY, "get_grid_order: buffer too small");
    cudaSetDevice(h->device);
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    NB_CUDA(h, cudaMemcpy(atom_index_host, h->d_atom_index, sizeof(int) * (size_t)h->npad, cudaMemcpyDeviceToHost));
    return h->npad;
}

extern "C" long long b200nb_get_tiles(b200nb_t* h, int outer, int* tiles_host, long long cap)
{
    if (!h) return B200NB_ERR_ARG;
    if (!h->have_list) return nb_fail(h, B200NB_ERR_STATE, "get_tiles: no pair list");
    cudaSetDevice(h->device);
    if (!outer && refresh_inner_list(h)) return B200NB_ERR_CUDA;
    cudaStreamSynchronize(h->stream);
    long long n = 0;
    for (int l = 0; l < 2; l++)
    {
        const PairList& L = (outer || h->inner_is_outer) ? h->outer[l] : h->inner[l];
        if (L.nentries == 0) continue;
        std::vector<Entry> e((size_t)L.nentries);
        std::vector<int>   cj((size_t)h->outer[l].ntiles);
        cudaMemcpy(e.data(), L.entries, sizeof(Entry) * e.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(cj.data(), L.cj, sizeof(int) * cj.size(), cudaMemcpyDeviceToHost);
        for (const Entry& en : e)
            for (int t = en.start; t < en.end; t++)
            {
                if (tiles_host && n < cap)
                {
                    tiles_host[3 * n]     = en.ci;
                    tiles_host[3 * n + 1] = en.shift_nmask & 255;
                    tiles_host[3 * n + 2] = cj[t];
                }
                n++;
            }
    }
    return n;
}

/* every interacting atom pair of the PACKED list (what the force kernel consumes): mask bit set, not on/below the
 * diagonal of the i-cluster's own atoms, both atoms real, r^2 < r2 -- the same predicate and step layout the force kernel
 * applies.  One warp per half-entry: lane = j16 + 16*kp looks at j-atom j16 of the step against the i-atoms 4*half + 2*kp, + 1 */
__global__ void __launch_bounds__(128)
k_pairs(const Entry* __restrict__ ent, const int* __restrict__ pja, const uint64_t* __restrict__ tmask, long long nentries,
        const float* __restrict__ xq, const float* __restrict__ shift_vec, const int* __restrict__ atom_index, int nslots, float r2,
        int intra, int* __restrict__ out, long long cap, unsigned long long* __restrict__ counter)
{
    const long long e = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= nentries) return;
    const int   lane = threadIdx.x & 31, j16 = lane & 15, kp = lane >> 4;
    const Entry en   = ent[e];
    const int   shift = NB_ENTRY_SHIFT(en.shift_nmask), nmask = NB_ENTRY_NMASK(en.shift_nmask), half = NB_ENTRY_HALF(en.shift_nmask);
    const float sx = shift_vec[3 * shift], sy = shift_vec[3 * shift + 1], sz = shift_vec[3 * shift + 2];
    const int   nstep = en.end - en.start;
    for (int s = 0; s < nstep; s++)
    {
        const int      js   = pja[(size_t)(en.start + s) * 16 + j16];
        const uint64_t m    = s < nmask ? tmask[en.start + s] : ~(uint64_t)0;
        const bool     diag = intra && shift == B200NB_CENTRAL && (js >> 3) == en.ci;
        const float4   xj   = reinterpret_cast<const float4*>(xq)[js];
        const int      aj   = js < nslots ? atom_index[js] : -1;
        for (int kk = 0; kk < 2; kk++)
        {
            const int    k  = 2 * kp + kk, i = 4 * half + k;
            const float4 xi = reinterpret_cast<const float4*>(xq)[(size_t)en.ci * 8 + i];
            const float  r  = nb_rsq(xi.x + sx, xi.y + sy, xi.z + sz, xj.x, xj.y, xj.z);
            const int    ai = atom_index[en.ci * 8 + i];
            bool ok = (r < r2) && ((m >> (16 * k + j16)) & 1u) && ai >= 0 && aj >= 0 && !(diag && (js & 7) <= i);
            if (ok)
            {
                unsigned long long pos = atomicAdd(counter, 1ull);
                if ((long long)pos < cap)
                {
                    out[3 * pos]     = ai;
                    out[3 * pos + 1] = aj;
                    out[3 * pos + 2] = shift;
                }
            }
        }
    }
}

extern "C" long long b200nb_get_pairs(b200nb_t* h, float r, int* pairs_host, long long cap)
{
    if (!h) return B200NB_ERR_ARG;
    if (!h->have_list) return nb_fail(h, B200NB_ERR_STATE, "get_pairs: no pair list");
    cudaSetDevice(h->device);
    int* d_out = nullptr;
    if (pairs_host && cap > 0)
        if (cudaMalloc((void**)&d_out, sizeof(int) * 3 * (size_t)cap) != cudaSuccess) return nb_fail(h, B200NB_ERR_CUDA, "get_pairs: cudaMalloc");
    cudaMemsetAsync(h->d_counter + 5, 0, sizeof(long long), h->stream);
    for (int l = 0; l < 2; l++)
    {
        const PackedList& L = h->packed[l];
        if (L.nentries == 0) continue;
        k_pairs<<<(unsigned)((L.nentries + 3) / 4), 128, 0, h->stream>>>(L.entries, L.ja, L.mask, L.nentries, h->d_xq, h->d_shift_vec,
                                                                         h->d_atom_index, h->npad, r * r, l == 0, d_out, d_out ? cap : 0,
                                                                         (unsigned long long*)(h->d_counter + 5));
        h->nlaunches++;
    }
    long long n = 0;
    cudaMemcpyAsync(&n, h->d_counter + 5, sizeof(n), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    if (d_out)
    {
        cudaMemcpy(pairs_host, d_out, sizeof(int) * 3 * (size_t)std::min(n, cap), cudaMemcpyDeviceToHost);
        cudaFree(d_out);
    }
    if (cudaGetLastError() != cudaSuccess) return nb_fail(h, B200NB_ERR_CUDA, "get_pairs: kernel failed");
    return n;
}

extern "C" int b200nb_time_force_kernel(b200nb_t* h, int locality, int flags, int nwarm, int niter, int flush_l2, float* ms_avg)
{
    if (!h || !ms_avg || niter < 1) return nb_fail(h, B200NB_ERR_ARG, "time_force_kernel: bad argument");
    if (!h->have_list) return nb_fail(h, B200NB_ERR_STATE, "time_force_kernel: no pair list");
    cudaSetDevice(h->device);
    if (flush_l2 && !h->d_flush)
    {
        h->flush_bytes = (size_t)256 << 20; /* > 126 MB L2 */
        NB_CUDA(h, cudaMalloc((void**)&h->d_flush, h->flush_bytes));
    }
    cudaEvent_t e0, e1;
    NB_CUDA(h, cudaEventCreate(&e0));
    NB_CUDA(h, cudaEventCreate(&e1));
    int rc;
    for (int i = 0; i < nwarm; i++)
        if ((rc = b200nb_launch_force(h, locality, flags))) return rc;
    double total = 0;
    for (int i = 0; i < niter; i++)
    {
        if (flush_l2)
        {
            k_flush<<<148 * 8, 256, 0, h->stream>>>(h->d_flush, h->flush_bytes / sizeof(float));
        }
        NB_CUDA(h, cudaEventRecord(e0, h->stream));
        if ((rc = b200nb_launch_force(h, locality, flags))) return rc;
        NB_CUDA(h, cudaEventRecord(e1, h->stream));
        NB_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0;
        NB_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_avg = (float)(total / niter);
    return 0;
}
