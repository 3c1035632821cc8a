/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
 *
 * Plain-C CPU restatement of the reference short-range nonbonded (nbnxm) hot path of
 * kassonlab/gmxapi (GROMACS 2021), written from scratch as the parity oracle for the CUDA
 * implementation under gmxapi_b200/csrc.  Each function cites the reference file:line whose
 * behaviour it restates (paths relative to /root/reference/src/gromacs).
 *
 * PARITY PINNING: this oracle is checked (tests/test_oracle_vs_reference.py, tests/golden/) against
 *  (a) the reference's own golden vectors api/nblib/tests/refdata/NBlibTest_ArgonForcesAreCorrect.xml
 *      and NBlibTest_SpcMethanolForcesAreCorrect.xml, and
 *  (b) the UNMODIFIED reference code compiled here by oracle/build_ref.sh (oracle/_ref): grid atom
 *      order, in-range pair set (bit-exact), forces / shift forces / energies of the CPU SIMD kernels.
 *
 * Conventions restated:
 *  - shift vectors: pbcutil/ishift.h:40-54 (D_BOX 2,1,1 -> 45 shifts, CENTRAL 22), pbcutil/pbc.cpp:1187
 *  - only shifts <= CENTRAL are listed for an intra-grid search and the i-atom is the shifted one
 *    (nbnxm/pairlist.cpp:3339-3342, kernels_simd_2xmm/kernel_outer.h:482-489)
 *  - r^2 = fma(dz,dz,fma(dx,dx,dy*dy)) with dx = (xi+shift) - xj: simd/vector_operations.h:106-115
 *    (norm2: ax*ax, ay*ay+ret, az*az+ret) AS COMPILED by gcc 13 -O3 -mfma for the 2xMM kernels: the
 *    disassembly of kernels_simd_2xmm/kernel_ElecEw_VdwLJCombGeom_F.cpp shows vmulps(dy,dy),
 *    vfmadd231ps(dx,dx), vfmadd231ps(dz,dz).  Verified bit-for-bit against oracle/_ref by
 *    tests/test_oracle_vs_reference.py::test_rsq_formula.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_CENTRAL 22
#define ORC_SHIFTS 45
#define ORC_CL 8        /* atoms per cluster, nbnxm/pairlistparams.h:66 */
#define ORC_CELL 64     /* atoms per grid cell (super-cluster), pairlistparams.h:69-77 */

enum { ORC_EEL_CUT = 0, ORC_EEL_RF = 1, ORC_EEL_EWALD = 2 };

typedef struct
{
    float rc;         /* rvdw = rcoulomb */
    int   eeltype;    /* ORC_EEL_*; CUT is evaluated as RF with k_rf=0 (nbnxm/kerneldispatch.cpp:168-171) */
    float epsfac;     /* interaction_const_t::epsfac */
    float k_rf, c_rf;
    float beta;       /* ewaldcoeff_q */
    float sh_ewald;
    float disp_cpot;  /* dispersion_shift.cpot */
    float rep_cpot;   /* repulsion_shift.cpot */
    int   ntypes;
    const float* nbfp; /* ntypes*ntypes*2: 6*C6, 12*C12 (nbnxm/atomdata.cpp:498-506) */
    /* Van der Waals modifiers and the twin-range cut-off (interaction_const_t, mdtypes/interaction_const.h:107-172) */
    float rvdw;        /* <= rc; 0 means rc.  rvdw < rc: VDW_CUTOFF_CHECK, kernel_ref_inner.h:252-262 */
    int   vdw_modifier; /* ORC_VDW_* */
    float rvdw_switch;
    float disp_c2, disp_c3, rep_c2, rep_c3; /* dispersion_shift / repulsion_shift .c2 .c3 (force switch) */
    float sw_c3, sw_c4, sw_c5;              /* vdw_switch (potential switch) */
    /* LJ-PME real-space grid correction (vdwtype = evdwPME; kernel_ref_inner.h:207-250, kernel_ref_outer.h:182-189,316-321) */
    int   ljpme;          /* 0 none, 1 geometric (eljpmeGEOM), 2 Lorentz-Berthelot (eljpmeLB) grid combination rule */
    float ewaldcoeff_lj;  /* interaction_const_t::ewaldcoeff_lj */
    float sh_lj_ewald;    /* interaction_const_t::sh_lj_ewald (mdlib/forcerec.cpp:709-717) */
} orc_params;

enum { ORC_VDW_POTSHIFT = 0, ORC_VDW_FORCESWITCH = 1, ORC_VDW_POTSWITCH = 2 };

/* mdlib/forcerec.cpp:787-801 force_switch_constants for a potential r^-p: out = {c2, c3, cpot} */
void orc_force_switch_constants(double pw, double rsw, double rc, float out[3])
{
    const double c2 = ((pw + 1) * rsw - (pw + 4) * rc) / (pow(rc, pw + 2) * (rc - rsw) * (rc - rsw));
    const double c3 = -((pw + 1) * rsw - (pw + 3) * rc) / (pow(rc, pw + 2) * (rc - rsw) * (rc - rsw) * (rc - rsw));
    /* the reference evaluates these in `real` = float: round the stored coefficients the same way */
    const float c2f = (float)c2, c3f = (float)c3;
    const double d  = rc - rsw;
    out[0]          = c2f;
    out[1]          = c3f;
    out[2]          = (float)(-pow(rc, -pw) + pw * c2f / 3 * d * d * d + pw * c3f / 4 * d * d * d * d);
}

/* mdlib/forcerec.cpp:803-816 potential_switch_constants: out = {c3, c4, c5} */
void orc_potential_switch_constants(double rsw, double rc, float out[3])
{
    const double d = rc - rsw;
    out[0]         = (float)(-10 / (d * d * d));
    out[1]         = (float)(15 / (d * d * d * d));
    out[2]         = (float)(-6 / (d * d * d * d * d));
}

static int shift_index(int tx, int ty, int tz)
{
    return 5 * (3 * (tz + 1) + (ty + 1)) + tx + 2; /* pbcutil/ishift.h:50 XYZ2IS */
}

/* Triclinic cells: the functions below take the DIAGONAL of the box matrix (box[3] = {a_x, b_y, c_z}); the off-diagonal elements
 * of a lower-triangular GROMACS box, {b_x, c_x, c_y} = box[YY][XX], box[ZZ][XX], box[ZZ][YY], are a mode of the oracle set with
 * orc_set_triclinic (all zero = rectangular, the default).  Atoms are expected in the brick [0, a_x) x [0, b_y) x [0, c_z), where
 * put_atoms_in_box leaves them (pbcutil/pbc.cpp). */
static float g_tric[3] = { 0.f, 0.f, 0.f };
void orc_set_triclinic(const float offdiag[3])
{
    for (int d = 0; d < 3; d++) g_tric[d] = offdiag ? offdiag[d] : 0.f;
}
static int orc_is_triclinic(void) { return g_tric[0] != 0.f || g_tric[1] != 0.f || g_tric[2] != 0.f; }

/* pbcutil/pbc.cpp:1187-1202 calc_shifts: shift_vec[n] = k a + l b + m c, in float like the reference */
void orc_shift_vectors(const float box[3], float* shift_vec /* 45*3 */)
{
    const float a[3] = { box[0], 0.f, 0.f }, b[3] = { g_tric[0], box[1], 0.f }, c[3] = { g_tric[1], g_tric[2], box[2] };
    int         n = 0;
    for (int m = -1; m <= 1; m++)
        for (int l = -1; l <= 1; l++)
            for (int k = -2; k <= 2; k++, n++)
                for (int d = 0; d < 3; d++) shift_vec[3 * n + d] = k * a[d] + l * b[d] + m * c[d];
}

/* r^2 with the reference's operand roles and operation order (see header). */
float orc_rsq(float xi, float yi, float zi, float sx, float sy, float sz, float xj, float yj, float zj)
{
    float dx = (xi + sx) - xj;
    float dy = (yi + sy) - yj;
    float dz = (zi + sz) - zj;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* ------------------------------------------------------------------------------------------------
 * Gridding and sorting
 * ------------------------------------------------------------------------------------------------ */

/* nbnxm/grid.cpp:103-262 Grid::setDimensions, GPU (hierarchical 8x8x8) geometry, home zone. */
void orc_grid_dims(int natoms, const float lower[3], const float upper[3], float density, int* ncx, int* ncy,
                   float cell_size[2], float inv_cell_size[2])
{
    float size[3];
    for (int d = 0; d < 3; d++) size[d] = upper[d] - lower[d];
    if (density <= 0) density = (float)natoms / (size[0] * size[1] * size[2]);
    if (natoms > ORC_CELL)
    {
        float tlen   = cbrtf((float)ORC_CL / density);
        float tlen_x = tlen * 2, tlen_y = tlen * 2;
        int   nx = (int)(size[0] / tlen_x), ny = (int)(size[1] / tlen_y);
        *ncx = nx > 1 ? nx : 1;
        *ncy = ny > 1 ? ny : 1;
    }
    else
    {
        *ncx = 1;
        *ncy = 1;
    }
    cell_size[0]     = size[0] / *ncx;
    cell_size[1]     = size[1] / *ncy;
    inv_cell_size[0] = 1 / cell_size[0];
    inv_cell_size[1] = 1 / cell_size[1];
}

typedef struct
{
    float key;
    int   idx;
} sort_item;

/* nbnxm/grid.cpp:292-427 sort_atoms: the pigeonhole + insertion procedure yields the exact order by
 * (coordinate, atom index) ascending; "Backwards" emits the same sequence reversed. */
static int cmp_item(const void* a, const void* b)
{
    const sort_item *p = (const sort_item*)a, *q = (const sort_item*)b;
    if (p->key < q->key) return -1;
    if (p->key > q->key) return 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}

static void sort_atoms(int dim, int backwards, int* a, int n, const float* x)
{
    if (n <= 1) return;
    sort_item* it = (sort_item*)malloc(sizeof(sort_item) * (size_t)n);
    for (int i = 0; i < n; i++)
    {
        it[i].key = x[3 * a[i] + dim];
        it[i].idx = a[i];
    }
    qsort(it, (size_t)n, sizeof(sort_item), cmp_item);
    for (int i = 0; i < n; i++) a[i] = backwards ? it[n - 1 - i].idx : it[i].idx;
    free(it);
}

/* nbnxm/grid.cpp:1173-1268 calcColumnIndices (home zone), :1287-1445 setCellIndices,
 * :1051-1164 sortColumnsGpuGeometry.  Outputs:
 *   col_cell0[ncol+1]  first cell of each column (cxy_ind_)
 *   atom_index[npad]   original atom at each grid slot, -1 for fillers (gridSetData.atomIndices)
 *   slot_of_atom[n]    grid slot of each atom (gridSetData.cells)
 * Returns npad (= 64 * number of cells) or -1 if cap is too small. */
int orc_put_on_grid(int natoms, const float* x, const float lower[3], const float upper[3], float density,
                    int* ncx_out, int* ncy_out, int* col_cell0, int col_cap, int* atom_index, int cap,
                    int* slot_of_atom)
{
    int   ncx, ncy;
    float cs[2], ics[2];
    orc_grid_dims(natoms, lower, upper, density, &ncx, &ncy, cs, ics);
    *ncx_out = ncx;
    *ncy_out = ncy;
    const int ncol = ncx * ncy;
    if (ncol + 1 > col_cap) return -1;
    int* col_of = (int*)malloc(sizeof(int) * (size_t)natoms);
    int* cnt    = (int*)calloc((size_t)ncol + 1, sizeof(int));
    for (int i = 0; i < natoms; i++)
    {
        int cx = (int)((x[3 * i + 0] - lower[0]) * ics[0]);
        int cy = (int)((x[3 * i + 1] - lower[1]) * ics[1]);
        if (cx > ncx - 1) cx = ncx - 1;
        if (cy > ncy - 1) cy = ncy - 1;
        if (cx < 0) cx = 0; /* atoms a few bits below the lower bound, grid.cpp:1207-1213 */
        if (cy < 0) cy = 0;
        col_of[i] = cx * ncy + cy;
        cnt[col_of[i]]++;
    }
    col_cell0[0] = 0;
    for (int c = 0; c < ncol; c++) col_cell0[c + 1] = col_cell0[c] + (cnt[c] + ORC_CELL - 1) / ORC_CELL;
    const int npad = col_cell0[ncol] * ORC_CELL;
    if (npad > cap)
    {
        free(col_of);
        free(cnt);
        return -1;
    }
    for (int s = 0; s < npad; s++) atom_index[s] = -1;
    int* fill = (int*)calloc((size_t)ncol, sizeof(int));
    for (int i = 0; i < natoms; i++) /* grid.cpp:1390-1400: atoms enter their column in index order */
    {
        int c                                          = col_of[i];
        atom_index[col_cell0[c] * ORC_CELL + fill[c]++] = i;
    }
    for (int c = 0; c < ncol; c++)
    {
        const int n   = cnt[c];
        int*      a   = atom_index + col_cell0[c] * ORC_CELL;
        const int ncz = col_cell0[c + 1] - col_cell0[c];
        sort_atoms(2, 0, a, n, x);
        for (int sub_z = 0; sub_z < ncz * 2; sub_z++)
        {
            const int offz = sub_z * 32;
            int       nz   = n - offz < 32 ? n - offz : 32;
            sort_atoms(1, (sub_z & 1) != 0, a + offz, nz, x);
            for (int sub_y = 0; sub_y < 2; sub_y++)
            {
                const int offy = offz + sub_y * 16;
                int       ny   = n - offy < 16 ? n - offy : 16;
                /* grid.cpp:1133: direction ((cz*2+sub_y)&1), cz=-1 on odd sub_z: parity of sub_y */
                sort_atoms(0, (sub_y & 1) != 0, a + offy, ny, x);
            }
        }
    }
    for (int s = 0; s < npad; s++)
        if (atom_index[s] >= 0) slot_of_atom[atom_index[s]] = s;
    free(col_of);
    free(cnt);
    free(fill);
    return npad;
}

/* Bounding boxes of 8-atom clusters (nbnxm/grid.cpp:454-520 calc_bounding_box*, fillers ignored;
 * all-filler clusters get the filler coordinate, atomdata.cpp:136-146). bb[c*6] = lx,ly,lz,ux,uy,uz. */
void orc_bounding_boxes(int npad, const int* atom_index, const float* x, float* bb)
{
    for (int c = 0; c < npad / ORC_CL; c++)
    {
        float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
        int   any   = 0;
        for (int k = 0; k < ORC_CL; k++)
        {
            int a = atom_index[c * ORC_CL + k];
            if (a < 0) continue;
            any = 1;
            for (int d = 0; d < 3; d++)
            {
                if (x[3 * a + d] < lo[d]) lo[d] = x[3 * a + d];
                if (x[3 * a + d] > hi[d]) hi[d] = x[3 * a + d];
            }
        }
        for (int d = 0; d < 3; d++)
        {
            bb[6 * c + d]     = any ? lo[d] : -1000000.0f;
            bb[6 * c + 3 + d] = any ? hi[d] : -1000000.0f;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * In-range atom pairs (the observable of search + prune) via a uniform cell list
 * ------------------------------------------------------------------------------------------------ */

typedef struct
{
    int* head;
    int* next;
    int  nc[3];
    float inv[3];
} cell_list;

static void cl_build(cell_list* cl, int n, const float* x, const float box[3], float r)
{
    for (int d = 0; d < 3; d++)
    {
        cl->nc[d] = (int)(box[d] / r);
        if (cl->nc[d] < 1) cl->nc[d] = 1;
        if (cl->nc[d] > 256) cl->nc[d] = 256;
        cl->inv[d] = cl->nc[d] / box[d];
    }
    size_t ncell = (size_t)cl->nc[0] * cl->nc[1] * cl->nc[2];
    cl->head     = (int*)malloc(sizeof(int) * ncell);
    cl->next     = (int*)malloc(sizeof(int) * (size_t)n);
    for (size_t c = 0; c < ncell; c++) cl->head[c] = -1;
    for (int i = n - 1; i >= 0; i--)
    {
        int c[3];
        for (int d = 0; d < 3; d++)
        {
            double f = (double)x[3 * i + d] * cl->inv[d];
            c[d]     = (int)floor(f);
            c[d]     = ((c[d] % cl->nc[d]) + cl->nc[d]) % cl->nc[d];
        }
        size_t ci    = ((size_t)c[0] * cl->nc[1] + c[1]) * cl->nc[2] + c[2];
        cl->next[i]  = cl->head[ci];
        cl->head[ci] = i;
    }
}

static void cl_free(cell_list* cl)
{
    free(cl->head);
    free(cl->next);
}

typedef void (*pair_cb)(void* ctx, int ai, int aj, int shift, float rsq);

/* Enumerates every unordered atom pair {a,b}, a != b, whose minimum-image distance evaluated with the
 * reference's roles (i = the atom shifted by a shift vector with index <= CENTRAL,
 * nbnxm/pairlist.cpp:3339-3342) satisfies r^2 < r2max.  For the CENTRAL shift i<j by atom index. */
static void for_each_pair(int n, const float* x, const float box[3], float rmax, pair_cb cb, void* ctx)
{
    float sv[ORC_SHIFTS * 3];
    orc_shift_vectors(box, sv);
    if (orc_is_triclinic())
    {
        /* triclinic cell: brute force over all atom pairs and all 23 shifts with index <= CENTRAL (nbnxm/pairlist.cpp:3339-3342:
         * the i-atom carries the shift; CENTRAL: each pair once) -- small systems only */
        const float r2m = rmax * rmax;
        for (int is = 0; is <= ORC_CENTRAL; is++)
            for (int a = 0; a < n; a++)
                for (int b = (is == ORC_CENTRAL ? a + 1 : 0); b < n; b++)
                {
                    if (a == b) continue;
                    const float r2 = orc_rsq(x[3 * a], x[3 * a + 1], x[3 * a + 2], sv[3 * is], sv[3 * is + 1], sv[3 * is + 2], x[3 * b],
                                             x[3 * b + 1], x[3 * b + 2]);
                    if (r2 < r2m) cb(ctx, a, b, is, r2);
                }
        return;
    }
    cell_list cl;
    /* cells at least rmax*(1+eps) wide so that +-1 neighbour cells suffice */
    cl_build(&cl, n, x, box, rmax * 1.0001f + 1e-6f);
    const float r2max = rmax * rmax;
    for (int a = 0; a < n; a++)
    {
        int c[3];
        for (int d = 0; d < 3; d++)
        {
            c[d] = (int)floor((double)x[3 * a + d] * cl.inv[d]);
            c[d] = ((c[d] % cl.nc[d]) + cl.nc[d]) % cl.nc[d];
        }
        int lo[3], hi[3];
        for (int d = 0; d < 3; d++)
        {
            if (cl.nc[d] >= 3)
            {
                lo[d] = -1;
                hi[d] = 1;
            }
            else
            {
                lo[d] = 0;
                hi[d] = cl.nc[d] - 1; /* visit every cell once */
            }
        }
        for (int ox = lo[0]; ox <= hi[0]; ox++)
            for (int oy = lo[1]; oy <= hi[1]; oy++)
                for (int oz = lo[2]; oz <= hi[2]; oz++)
                {
                    int cc[3] = { cl.nc[0] >= 3 ? (c[0] + ox + cl.nc[0]) % cl.nc[0] : ox,
                                  cl.nc[1] >= 3 ? (c[1] + oy + cl.nc[1]) % cl.nc[1] : oy,
                                  cl.nc[2] >= 3 ? (c[2] + oz + cl.nc[2]) % cl.nc[2] : oz };
                    size_t ci = ((size_t)cc[0] * cl.nc[1] + cc[1]) * cl.nc[2] + cc[2];
                    for (int b = cl.head[ci]; b >= 0; b = cl.next[b])
                    {
                        if (b <= a) continue;
                        /* minimum image shift of a relative to b, in double to choose the image */
                        int t[3];
                        for (int d = 0; d < 3; d++)
                        {
                            double dd = (double)x[3 * a + d] - (double)x[3 * b + d];
                            t[d]      = -(int)lrint(dd / box[d]);
                        }
                        if (abs(t[0]) > 2 || abs(t[1]) > 1 || abs(t[2]) > 1) continue;
                        int is = shift_index(t[0], t[1], t[2]);
                        int ai = a, aj = b;
                        if (is > ORC_CENTRAL)
                        {
                            /* the other atom carries the (negated) shift */
                            is = shift_index(-t[0], -t[1], -t[2]);
                            ai = b;
                            aj = a;
                        }
                        float r2 = orc_rsq(x[3 * ai], x[3 * ai + 1], x[3 * ai + 2], sv[3 * is], sv[3 * is + 1],
                                           sv[3 * is + 2], x[3 * aj], x[3 * aj + 1], x[3 * aj + 2]);
                        if (r2 < r2max) cb(ctx, ai, aj, is, r2);
                    }
                }
    }
    cl_free(&cl);
}

/* exclusion lookup in the CSR topology exclusions (utility/listoflists.h as used by
 * nbnxm/pairlist.cpp:1874-1972 setExclusionsForIEntry) */
static int is_excluded(const int* excl_off, const int* excl_idx, int a, int b)
{
    if (!excl_off) return 0;
    for (int k = excl_off[a]; k < excl_off[a + 1]; k++)
        if (excl_idx[k] == b) return 1;
    return 0;
}

typedef struct
{
    int*        out;
    long long   cap, n;
    const int * excl_off, *excl_idx;
    int         include_excluded;
} pairset_ctx;

static void pairset_cb(void* vctx, int ai, int aj, int shift, float rsq)
{
    (void)rsq;
    pairset_ctx* c = (pairset_ctx*)vctx;
    if (!c->include_excluded && is_excluded(c->excl_off, c->excl_idx, ai, aj)) return;
    if (c->out && c->n < c->cap)
    {
        c->out[3 * c->n]     = ai;
        c->out[3 * c->n + 1] = aj;
        c->out[3 * c->n + 2] = shift;
    }
    c->n++;
}

/* The set of interacting atom pairs after search and prune at radius r: every non-excluded pair with
 * r^2 < r*r, as (i, j, shift) with the reference's roles. Returns the count (writes at most cap). */
long long orc_pair_set(int n, const float* x, const float box[3], float r, const int* excl_off,
                       const int* excl_idx, int* pairs, long long cap)
{
    pairset_ctx c = { pairs, cap, 0, excl_off, excl_idx, 0 };
    for_each_pair(n, x, box, r, pairset_cb, &c);
    return c.n;
}

/* ------------------------------------------------------------------------------------------------
 * Cluster-pair (tile) list: every (ci, cj, shift) of 8-atom clusters of the grid with at least one
 * atom pair within rlist -- what the reference list converges to after dynamic pruning
 * (kernels_reference/kernel_ref_prune.cpp:45-143: keep a cluster pair iff any r^2 < rlist^2;
 * cuda/nbnxm_cuda_kernel_pruneonly.cuh:104-277), with the half-list rules of
 * nbnxm/pairlist.cpp:3339-3342,3365-3372 (shift <= CENTRAL; cj >= ci for CENTRAL).
 * ------------------------------------------------------------------------------------------------ */
typedef struct
{
    uint64_t* keys;
    long long cap, n;
    const int* slot;
} tile_ctx;

static void tile_cb(void* vctx, int ai, int aj, int shift, float rsq)
{
    (void)rsq;
    tile_ctx* c  = (tile_ctx*)vctx;
    int       ci = c->slot[ai] / ORC_CL, cj = c->slot[aj] / ORC_CL;
    if (shift == ORC_CENTRAL && cj < ci)
    {
        int t = ci;
        ci    = cj;
        cj    = t;
    }
    if (c->n < c->cap) c->keys[c->n] = ((uint64_t)ci << 38) | ((uint64_t)shift << 32) | (uint64_t)cj;
    c->n++;
}

static int cmp_u64(const void* a, const void* b)
{
    uint64_t p = *(const uint64_t*)a, q = *(const uint64_t*)b;
    return (p > q) - (p < q);
}

/* tiles[3*k] = ci, shift, cj sorted by (ci, shift, cj). Returns count or -1 on overflow. */
long long orc_tile_list(int n, const float* x, const float box[3], float rlist, const int* slot_of_atom,
                        int* tiles, long long cap)
{
    long long kcap = (1 << 20) + n;
    for (;;)
    {
        tile_ctx c = { (uint64_t*)malloc(sizeof(uint64_t) * (size_t)kcap), kcap, 0, slot_of_atom };
        /* every non-empty cluster pairs with itself (distance 0): the reference always lists the diagonal
         * cluster pair, which also carries the self-energy term (kernel_outer.h:408-452) */
        for (int a = 0; a < n; a++) tile_cb(&c, a, a, ORC_CENTRAL, 0.0f);
        for_each_pair(n, x, box, rlist, tile_cb, &c);
        if (c.n > kcap)
        {
            free(c.keys);
            kcap = c.n + 16;
            continue;
        }
        qsort(c.keys, (size_t)c.n, sizeof(uint64_t), cmp_u64);
        long long m = 0;
        for (long long k = 0; k < c.n; k++)
        {
            if (k > 0 && c.keys[k] == c.keys[k - 1]) continue;
            if (m < cap)
            {
                tiles[3 * m]     = (int)(c.keys[k] >> 38);
                tiles[3 * m + 1] = (int)((c.keys[k] >> 32) & 63);
                tiles[3 * m + 2] = (int)(c.keys[k] & 0xffffffffu);
            }
            m++;
        }
        free(c.keys);
        return m <= cap ? m : -1;
    }
}

/* kernels_reference/kernel_ref_prune.cpp:45-143 restated for 8x8 tiles in grid order: keep[k]=1 iff any
 * atom pair of tile k has r^2 < rlist_inner^2. x in ORIGINAL atom order, atom_index maps slot->atom. */
void orc_prune_tiles(long long ntiles, const int* tiles, const int* atom_index, const float* x,
                     const float box[3], float rlist_inner, unsigned char* keep)
{
    float sv[ORC_SHIFTS * 3];
    orc_shift_vectors(box, sv);
    const float r2 = rlist_inner * rlist_inner;
    for (long long k = 0; k < ntiles; k++)
    {
        int ci = tiles[3 * k], is = tiles[3 * k + 1], cj = tiles[3 * k + 2];
        int in = 0;
        for (int i = 0; i < ORC_CL && !in; i++)
        {
            int a = atom_index[ci * ORC_CL + i];
            if (a < 0) continue;
            for (int j = 0; j < ORC_CL; j++)
            {
                int b = atom_index[cj * ORC_CL + j];
                if (b < 0) continue;
                if (orc_rsq(x[3 * a], x[3 * a + 1], x[3 * a + 2], sv[3 * is], sv[3 * is + 1], sv[3 * is + 2],
                            x[3 * b], x[3 * b + 1], x[3 * b + 2])
                    < r2)
                {
                    in = 1;
                    break;
                }
            }
        }
        keep[k] = (unsigned char)in;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Pair interactions
 * ------------------------------------------------------------------------------------------------ */

/* simd/simd_math.h:1609-1650 pmeForceCorrection (single precision rational minimax), scalar */
static float pme_force_correction(float z2)
{
    const float FN6 = -1.7357322914161492954e-8f, FN5 = 1.4703624142580877519e-6f,
                FN4 = -0.000053401640219807709149f, FN3 = 0.0010054721316683106153f,
                FN2 = -0.019278317264888380590f, FN1 = 0.069670166153766424023f,
                FN0 = -0.75225204789749321333f;
    const float FD4 = 0.0011193462567257629232f, FD3 = 0.014866955030185295499f,
                FD2 = 0.11583842382862377919f, FD1 = 0.50736591960530292870f, FD0 = 1.0f;
    float z4 = z2 * z2;
    float d0 = fmaf(FD4, z4, FD2), d1 = fmaf(FD3, z4, FD1);
    d0       = fmaf(d0, z4, FD0);
    d0       = fmaf(d1, z2, d0);
    d0       = 1.0f / d0;
    float n0 = fmaf(FN6, z4, FN4), n1 = fmaf(FN5, z4, FN3);
    n0       = fmaf(n0, z4, FN2);
    n1       = fmaf(n1, z4, FN1);
    n0       = fmaf(n0, z4, FN0);
    n0       = fmaf(n1, z2, n0);
    return n0 * d0;
}

/* simd/simd_math.h:1687-1722 pmePotentialCorrection */
static float pme_potential_correction(float z2)
{
    const float VN6 = 1.9296833005951166339e-8f, VN5 = -1.4213390571557850962e-6f,
                VN4 = 0.000041603292906656984871f, VN3 = -0.00013134036773265025626f,
                VN2 = 0.038657983986041781264f, VN1 = 0.11285044772717598220f, VN0 = 1.1283802385263030286f;
    const float VD3 = 0.0066752224023576045451f, VD2 = 0.078647795836373922256f,
                VD1 = 0.43336185284710920150f, VD0 = 1.0f;
    float z4 = z2 * z2;
    float d1 = fmaf(VD3, z4, VD1), d0 = fmaf(VD2, z4, VD0);
    d0       = fmaf(d1, z2, d0);
    d0       = 1.0f / d0;
    float n0 = fmaf(VN6, z4, VN4), n1 = fmaf(VN5, z4, VN3);
    n0       = fmaf(n0, z4, VN2);
    n1       = fmaf(n1, z4, VN1);
    n0       = fmaf(n0, z4, VN0);
    n0       = fmaf(n1, z2, n0);
    return n0 * d0;
}

typedef struct
{
    const orc_params* p;
    const float*      x;
    const float*      q;
    const int*        type;
    const int *       excl_off, *excl_idx;
    float             sv[ORC_SHIFTS * 3];
    double*           f;      /* 3n */
    double*           fshift; /* 45*3 */
    double            evdw, ecoul;
    long long         npairs; /* non-excluded pairs within the cut-off */
    int               want_energy;
} force_ctx;

/* One atom pair, the arithmetic of kernels_simd_2xmm/kernel_inner.h:226-880 (LJ cut-off with potential
 * shift; reaction-field / plain cut-off :376-383; analytical Ewald :386-400,462-474) evaluated in
 * single precision; excluded pairs within the cut-off keep only the reaction-field / Ewald exclusion
 * correction (EXCL_FORCES, :358-366,520-523).  Contributions are accumulated in double. */
static void force_cb(void* vctx, int ai, int aj, int is, float rsq)
{
    force_ctx*        c = (force_ctx*)vctx;
    const orc_params* p = c->p;
    const int   excluded = is_excluded(c->excl_off, c->excl_idx, ai, aj);
    const float interact = excluded ? 0.0f : 1.0f;
    if (!excluded) c->npairs++;

    const float dx = (c->x[3 * ai] + c->sv[3 * is]) - c->x[3 * aj];
    const float dy = (c->x[3 * ai + 1] + c->sv[3 * is + 1]) - c->x[3 * aj + 1];
    const float dz = (c->x[3 * ai + 2] + c->sv[3 * is + 2]) - c->x[3 * aj + 2];

    if (rsq < 3.82e-07f) rsq = 3.82e-07f; /* nbnxm/pairlist.h:146 c_nbnxnMinDistanceSquared */
    const float rinv    = 1.0f / sqrtf(rsq);
    const float rinvsq  = rinv * rinv;
    const float rinv_ex = rinv * interact;

    const float qq = (p->epsfac * c->q[ai]) * c->q[aj];
    float       frcoul, vcoul = 0;
    if (p->eeltype == ORC_EEL_EWALD)
    {
        const float brsq   = p->beta * p->beta * rsq;
        const float ewcorr = p->beta * pme_force_correction(brsq);
        frcoul             = qq * fmaf(ewcorr, brsq, rinv_ex);
        if (c->want_energy)
        {
            float vc_sub = p->beta * pme_potential_correction(brsq) + p->sh_ewald * interact;
            vcoul        = qq * (rinv_ex - vc_sub);
        }
    }
    else
    {
        const float k_rf = (p->eeltype == ORC_EEL_RF) ? p->k_rf : 0.0f;
        frcoul           = qq * fmaf(rsq, -2.0f * k_rf, rinv_ex);
        if (c->want_energy) vcoul = qq * (rinv_ex + fmaf(rsq, k_rf, -p->c_rf));
    }

    const int   ti = c->type[ai], tj = c->type[aj];
    const float c6 = p->nbfp[(ti * p->ntypes + tj) * 2], c12 = p->nbfp[(ti * p->ntypes + tj) * 2 + 1];
    const float rinvsix = rinvsq * rinvsq * rinvsq * interact;
    const float frlj6 = c6 * rinvsix, frlj12 = c12 * rinvsix * rinvsix;
    float       frlj  = frlj12 - frlj6;
    /* kernel_ref_inner.h:157-164: the LJ energy is also needed by the potential switch */
    float vlj = (1.0f / 12.0f) * fmaf(c12, p->rep_cpot, frlj12) - (1.0f / 6.0f) * fmaf(c6, p->disp_cpot, frlj6);
    if (p->vdw_modifier != ORC_VDW_POTSHIFT)
    {
        /* kernel_ref_inner.h:166-172: force or potential switching from rvdw_switch */
        const float r   = rsq * rinv;
        float       rsw = r - p->rvdw_switch;
        rsw             = (rsw >= 0.0f ? rsw : 0.0f);
        if (p->vdw_modifier == ORC_VDW_FORCESWITCH)
        {
            /* kernel_ref_inner.h:173-182 */
            frlj += -c6 * (p->disp_c2 + p->disp_c3 * rsw) * rsw * rsw * r + c12 * (p->rep_c2 + p->rep_c3 * rsw) * rsw * rsw * r;
            vlj += -c6 * (-p->disp_c2 / 3 - p->disp_c3 / 4 * rsw) * rsw * rsw * rsw
                   + c12 * (-p->rep_c2 / 3 - p->rep_c3 / 4 * rsw) * rsw * rsw * rsw;
            vlj = vlj * interact; /* :184-191 masking after force switching */
        }
        else
        {
            /* kernel_ref_inner.h:184-205: mask, then sw = 1 + c3 rsw^3 + c4 rsw^4 + c5 rsw^5, dsw = its derivative */
            vlj = vlj * interact;
            const float sw  = 1.0f + (p->sw_c3 + (p->sw_c4 + p->sw_c5 * rsw) * rsw) * rsw * rsw * rsw;
            const float dsw = (3 * p->sw_c3 + (4 * p->sw_c4 + 5 * p->sw_c5 * rsw) * rsw) * rsw * rsw;
            frlj            = frlj * sw - r * vlj * dsw;
            vlj *= sw;
        }
    }
    else
    {
        vlj = vlj * interact;
    }
    if (p->ljpme)
    {
        /* kernel_ref_inner.h:207-250: subtract the grid (mesh) part of the dispersion from the real-space LJ; the per-type
         * grid parameters are nbfp_comb as set_lj_parameter_data stores them (nbnxm/atomdata.cpp:291-322) */
        const float c6ii = p->nbfp[(ti * p->ntypes + ti) * 2], c12ii = p->nbfp[(ti * p->ntypes + ti) * 2 + 1];
        const float c6jj = p->nbfp[(tj * p->ntypes + tj) * 2], c12jj = p->nbfp[(tj * p->ntypes + tj) * 2 + 1];
        float       c6grid;
        if (p->ljpme == 1)
        {
            c6grid = sqrtf(c6ii) * sqrtf(c6jj);
        }
        else
        {
            const float si = (c6ii > 0 && c12ii > 0) ? 0.5f * powf(c12ii / c6ii, 1.0f / 6.0f) : 0.0f;
            const float ei = (c6ii > 0 && c12ii > 0) ? sqrtf(c6ii * c6ii / c12ii) : 0.0f;
            const float sj = (c6jj > 0 && c12jj > 0) ? 0.5f * powf(c12jj / c6jj, 1.0f / 6.0f) : 0.0f;
            const float ej = (c6jj > 0 && c12jj > 0) ? sqrtf(c6jj * c6jj / c12jj) : 0.0f;
            const float sigma = si + sj, sigma2 = sigma * sigma;
            c6grid            = ei * ej * sigma2 * sigma2 * sigma2;
        }
        const float lje_coeff2   = p->ewaldcoeff_lj * p->ewaldcoeff_lj;
        const float lje_coeff6_6 = lje_coeff2 * lje_coeff2 * lje_coeff2 / 6.0f;
        const float rinvsix_nm   = rinvsq * rinvsq * rinvsq; /* without the exclusion mask */
        const float cr2          = lje_coeff2 * rsq;
        const float expmcr2      = expf(-cr2);
        const float poly         = 1 + cr2 + 0.5f * cr2 * cr2;
        frlj += c6grid * (rinvsix_nm - expmcr2 * (rinvsix_nm * poly + lje_coeff6_6));
        vlj += c6grid / 6 * (rinvsix_nm * (1 - expmcr2 * poly) + p->sh_lj_ewald * interact);
    }
    if (p->rvdw > 0.0f && p->rvdw < p->rc)
    {
        /* kernel_ref_inner.h:252-262 VDW_CUTOFF_CHECK: VdW cut-off shorter than the Coulomb cut-off */
        const float skip = (rsq < p->rvdw * p->rvdw) ? 1.0f : 0.0f;
        frlj *= skip;
        vlj *= skip;
    }
    if (c->want_energy)
    {
        c->evdw += (double)vlj;
        c->ecoul += (double)vcoul;
    }
    const float fscal = rinvsq * (frcoul + frlj);
    const float tx = fscal * dx, ty = fscal * dy, tz = fscal * dz;
    c->f[3 * ai] += tx;
    c->f[3 * ai + 1] += ty;
    c->f[3 * ai + 2] += tz;
    c->f[3 * aj] -= tx;
    c->f[3 * aj + 1] -= ty;
    c->f[3 * aj + 2] -= tz;
    c->fshift[3 * is] += tx; /* kernels_simd_2xmm/kernel_outer.h:620-640: i-forces summed per shift */
    c->fshift[3 * is + 1] += ty;
    c->fshift[3 * is + 2] += tz;
}

/* Forces (original atom order), shift forces [45*3], energies {Vvdw, Vcoul}, pair count.
 * Self terms: kernels_simd_2xmm/kernel_outer.h:408-452 (RF: -0.5*c_rf*facel*q^2, Ewald:
 * -facel*q^2*beta/sqrt(pi)). Returns the number of non-excluded pairs within the cut-off. */
long long orc_forces(int n, const float* x, const float box[3], const float* q, const int* type,
                     const int* excl_off, const int* excl_idx, const orc_params* p, int want_energy,
                     double* f, double* fshift, double* energies)
{
    force_ctx c;
    memset(&c, 0, sizeof(c));
    c.p = p;
    c.x = x;
    c.q = q;
    c.type = type;
    c.excl_off = excl_off;
    c.excl_idx = excl_idx;
    c.f = f;
    c.fshift = fshift;
    c.want_energy = want_energy;
    orc_shift_vectors(box, c.sv);
    memset(f, 0, sizeof(double) * 3 * (size_t)n);
    memset(fshift, 0, sizeof(double) * 3 * ORC_SHIFTS);
    for_each_pair(n, x, box, p->rc, force_cb, &c);
    if (want_energy)
    {
        double sub = (p->eeltype == ORC_EEL_EWALD) ? 0.5 * p->beta * 1.12837916709551257390 /* M_2_SQRTPI */
                                                   : 0.5 * p->c_rf;
        for (int a = 0; a < n; a++) c.ecoul -= (double)p->epsfac * q[a] * q[a] * sub;
        if (p->ljpme)
        {
            /* kernel_ref_outer.h:316-321: LJ Ewald self interaction, 0.5 * (6 C6_ii) / 6 * coeff^6 / 6 per atom */
            const double c2 = (double)p->ewaldcoeff_lj * p->ewaldcoeff_lj, c6_6 = c2 * c2 * c2 / 6.0;
            for (int a = 0; a < n; a++) c.evdw += 0.5 * (double)p->nbfp[(type[a] * p->ntypes + type[a]) * 2] / 6.0 * c6_6;
        }
    }
    energies[0] = c.evdw;
    energies[1] = c.ecoul;
    return c.npairs;
}

/* mdlib/sim_util.cpp:157-183 calc_virial's shift-force part: vir = -0.5 * sum_s shift_vec[s] (x) fshift[s]
 * (mdlib/calcvir.cpp calc_vir); returned as 9 doubles row-major. */
void orc_virial_from_fshift(const float box[3], const double* fshift, double* vir)
{
    float sv[ORC_SHIFTS * 3];
    orc_shift_vectors(box, sv);
    for (int k = 0; k < 9; k++) vir[k] = 0;
    for (int s = 0; s < ORC_SHIFTS; s++)
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) vir[3 * a + b] += -0.5 * (double)sv[3 * s + a] * fshift[3 * s + b];
}


/* ------------------------------------------------------------------------------------------------
 * Perturbed (free-energy) pairs: gmxlib/nonbonded/nb_free_energy.cpp:203-860 restated for the flavours
 * our FEP kernel covers -- reaction-field / plain cut-off or Ewald electrostatics, cut-off LJ with potential shift,
 * soft-core with r-power 6 (lambda power 1 or 2) or none -- on a pair list in t_nblist form (mdtypes/nblist.h:117-137):
 * nri i-entries {iinr, shift, jindex[nri+1]}, jjnr, excl_fep (1: the pair interacts, 0: excluded, only its
 * reaction-field correction is evaluated; an atom listed with itself counts half).
 * Arithmetic in float like the reference's ScalarDataTypes instantiation (GMX_DOUBLE = 0), sums in float in
 * the reference's order of accumulation; out4 = {Vc, Vv, dvdl_coul, dvdl_vdw}.
 * ------------------------------------------------------------------------------------------------ */
typedef struct
{
    float rc, epsfac, k_rf, c_rf, disp_cpot, rep_cpot;
    float lambda_coul, lambda_vdw;
    float alpha_coul, alpha_vdw; /* interaction_const_t::SoftCoreParameters: alphaCoulomb = bScCoul ? sc_alpha : 0 */
    int   lam_power;
    float sigma6_def, sigma6_min;
    float beta, sh_ewald; /* beta > 0: Ewald electrostatics */
    float rvdw_switch;    /* > 0: LJ potential switch from rvdw_switch to rc (eintmodPOTSWITCH; disp_cpot = rep_cpot = 0 then) */
    int   ljpme;          /* 0: cut-off LJ; 1 / 2: LJ-PME with the geometric / Lorentz-Berthelot grid rule */
    float beta_lj, sh_lj_ewald; /* interaction_const_t::ewaldcoeff_lj, sh_lj_ewald */
    float rvdw;           /* > 0: rvdw < rcoulomb = rc (disp_cpot / rep_cpot / sh_lj_ewald are for rvdw then) */
} orc_fep_params;

void orc_fep_kernel(int natoms, const float* x, const float* shift_vec, int ntype, const float* nbfp, const int* typeA, const int* typeB,
                    const float* chargeA, const float* chargeB, int nri, const int* iinr, const int* shift, const int* jindex, const int* jjnr,
                    const signed char* excl_fep, const orc_fep_params* p, float* f, float* fshift, float* out4)
{
    const float facel = p->epsfac, krf = p->k_rf, crf = p->c_rf, rcoulomb = p->rc, rvdw = p->rvdw > 0.f ? p->rvdw : p->rc;
    const float alpha_coul = p->alpha_coul, alpha_vdw = p->alpha_vdw, sigma6_def = p->sigma6_def, sigma6_min = p->sigma6_min;
    const float lam_power = (float)p->lam_power;
    const int   useSoftCore = !(alpha_coul == 0.f && alpha_vdw == 0.f);                             /* :946-958 */
    const int   scDiffer = useSoftCore && !(p->lambda_coul == p->lambda_vdw && alpha_coul == alpha_vdw); /* :980-992 */
    const int   ewald = p->beta > 0.f, ljpme = p->ljpme != 0;
    const float rcutoff_max2 = (rcoulomb > rvdw ? rcoulomb : rvdw) * (rcoulomb > rvdw ? rcoulomb : rvdw); /* :300-301 */
    /* LJ-PME: the grid C6 per type pair at the positions of C6 in nbfp (fr->ljpme_c6grid: mdlib/forcerec.cpp:157-195
     * make_ljpme_c6grid, in real = float, from the types' own C6 / C12; nbfp holds 6 C6 and 12 C12) */
    float* c6grid = 0;
    if (ljpme)
    {
        c6grid = (float*)calloc(2 * (size_t)ntype * ntype, sizeof(float));
        for (int i = 0; i < ntype; i++)
            for (int j = 0; j < ntype; j++)
            {
                const float c6i = nbfp[2 * (i * ntype + i)] / 6.0f, c12i = nbfp[2 * (i * ntype + i) + 1] / 12.0f;
                const float c6j = nbfp[2 * (j * ntype + j)] / 6.0f, c12j = nbfp[2 * (j * ntype + j) + 1] / 12.0f;
                float       c6  = sqrtf(c6i * c6j);
                if (p->ljpme == 2 && c6 != 0.f && c12i != 0.f && c12j != 0.f)
                {
                    const float sigmai = (float)pow((double)(c12i / c6i), 1.0 / 6.0), sigmaj = (float)pow((double)(c12j / c6j), 1.0 / 6.0);
                    const float epsi = c6i * c6i / c12i, epsj = c6j * c6j / c12j, sm = 0.5f * (sigmai + sigmaj);
                    c6 = sqrtf(epsi * epsj) * (sm * sm * sm * sm * sm * sm);
                }
                c6grid[2 * (ntype * i + j)] = c6 * 6.0f;
            }
    }
    float LFC[2] = { 1.f - p->lambda_coul, p->lambda_coul }, LFV[2] = { 1.f - p->lambda_vdw, p->lambda_vdw }, DLF[2] = { -1.f, 1.f };
    float lfac_coul[2], dlfac_coul[2], lfac_vdw[2], dlfac_vdw[2];
    for (int i = 0; i < 2; i++) /* :363-370 */
    {
        lfac_coul[i]  = (p->lam_power == 2 ? (1 - LFC[i]) * (1 - LFC[i]) : (1 - LFC[i]));
        dlfac_coul[i] = DLF[i] * lam_power / 6.0f * (p->lam_power == 2 ? (1 - LFC[i]) : 1);
        lfac_vdw[i]   = (p->lam_power == 2 ? (1 - LFV[i]) * (1 - LFV[i]) : (1 - LFV[i]));
        dlfac_vdw[i]  = DLF[i] * lam_power / 6.0f * (p->lam_power == 2 ? (1 - LFV[i]) : 1);
    }
    for (int k = 0; k < 3 * natoms; k++) f[k] = 0.f;
    for (int k = 0; k < 3 * ORC_SHIFTS; k++) fshift[k] = 0.f;
    float dvdl_coul = 0.f, dvdl_vdw = 0.f, Vc = 0.f, Vv = 0.f;
    for (int n = 0; n < nri; n++)
    {
        int         npair_within_cutoff = 0;
        const int   is3 = 3 * shift[n], ii = iinr[n];
        const float ix = shift_vec[is3] + x[3 * ii], iy = shift_vec[is3 + 1] + x[3 * ii + 1], iz = shift_vec[is3 + 2] + x[3 * ii + 2];
        const float iqA = facel * chargeA[ii], iqB = facel * chargeB[ii];
        const int   ntiA = 2 * ntype * typeA[ii], ntiB = 2 * ntype * typeB[ii];
        float       vctot = 0, vvtot = 0, fix = 0, fiy = 0, fiz = 0;
        for (int k = jindex[n]; k < jindex[n + 1]; k++)
        {
            const int   jnr = jjnr[k];
            const float dx = ix - x[3 * jnr], dy = iy - x[3 * jnr + 1], dz = iz - x[3 * jnr + 2];
            const float rsq = dx * dx + dy * dy + dz * dz;
            const int   included = excl_fep == 0 || excl_fep[k];
            if (rsq >= rcutoff_max2 && included) continue; /* :421-434 */
            npair_within_cutoff++;
            float rinv = 0.f, r = 0.f, rp, rpm2;
            if (rsq > 0)
            {
                rinv = 1.0f / sqrtf(rsq);
                r    = rsq * rinv;
            }
            if (useSoftCore)
            {
                rpm2 = rsq * rsq;
                rp   = rpm2 * rsq;
            }
            else
            {
                rpm2 = rinv * rinv;
                rp   = 1;
            }
            float Fscal = 0;
            float qq[2] = { iqA * chargeA[jnr], iqB * chargeB[jnr] };
            int   tj[2] = { ntiA + 2 * typeA[jnr], ntiB + 2 * typeB[jnr] };
            if (included)
            {
                float c6[2], c12[2], sigma6[2] = { 0, 0 }, alpha_vdw_eff = 0, alpha_coul_eff = 0;
                float FscalC[2], FscalV[2], Vcoul[2], Vvdw[2];
                for (int i = 0; i < 2; i++)
                {
                    c6[i]  = nbfp[tj[i]];
                    c12[i] = nbfp[tj[i] + 1];
                    if (useSoftCore)
                    {
                        if (c6[i] > 0 && c12[i] > 0)
                        {
                            sigma6[i] = 0.5f * c12[i] / c6[i];
                            if (sigma6[i] < sigma6_min) sigma6[i] = sigma6_min;
                        }
                        else sigma6[i] = sigma6_def;
                    }
                }
                if (useSoftCore && !(c12[0] > 0 && c12[1] > 0)) /* :498-509: soft-core only if an end state has no repulsion */
                {
                    alpha_vdw_eff  = alpha_vdw;
                    alpha_coul_eff = alpha_coul;
                }
                for (int i = 0; i < 2; i++)
                {
                    FscalC[i] = FscalV[i] = Vcoul[i] = Vvdw[i] = 0;
                    float rinvC, rinvV, rC, rV, rpinvC, rpinvV;
                    if (qq[i] != 0 || c6[i] != 0 || c12[i] != 0)
                    {
                        if (useSoftCore)
                        {
                            rpinvC = 1.0f / (alpha_coul_eff * lfac_coul[i] * sigma6[i] + rp);
                            rC     = 1.0f / sqrtf(cbrtf(rpinvC)); /* pthRoot: invPthRoot = invsqrt(cbrt(.)), the effective r */
                            rinvC  = 1.0f / rC;
                            if (scDiffer)
                            {
                                rpinvV = 1.0f / (alpha_vdw_eff * lfac_vdw[i] * sigma6[i] + rp);
                                rV     = 1.0f / sqrtf(cbrtf(rpinvV));
                                rinvV  = 1.0f / rV;
                            }
                            else
                            {
                                rpinvV = rpinvC;
                                rinvV  = rinvC;
                                rV     = rC;
                            }
                        }
                        else
                        {
                            rpinvC = rpinvV = 1;
                            rinvC = rinvV = rinv;
                            rC = rV = r;
                        }
                        if (qq[i] != 0 && (ewald ? r < rcoulomb : rC < rcoulomb)) /* :565-581 */
                        {
                            if (ewald) /* plain (soft-cored) 1/r: the long-range part is subtracted below */
                            {
                                Vcoul[i]  = qq[i] * (rinvC - p->sh_ewald);
                                FscalC[i] = qq[i] * rinvC;
                            }
                            else
                            {
                                Vcoul[i]  = qq[i] * (rinvC + krf * rC * rC - crf);
                                FscalC[i] = qq[i] * (rinvC - 2.0f * krf * rC * rC);
                            }
                        }
                        if ((c6[i] != 0 || c12[i] != 0) && (ljpme ? r < rvdw : rV < rvdw)) /* :586-607 */
                        {
                            float rinv6;
                            if (useSoftCore) rinv6 = rpinvV;
                            else
                            {
                                rinv6 = rinvV * rinvV;
                                rinv6 = rinv6 * rinv6 * rinv6;
                            }
                            const float Vvdw6 = c6[i] * rinv6, Vvdw12 = c12[i] * rinv6 * rinv6;
                            Vvdw[i]   = (Vvdw12 + c12[i] * p->rep_cpot) * (1.0f / 12.0f) - (Vvdw6 + c6[i] * p->disp_cpot) * (1.0f / 6.0f);
                            FscalV[i] = Vvdw12 - Vvdw6;
                            if (ljpme) Vvdw[i] += c6grid[tj[i]] * p->sh_lj_ewald * (1.0f / 6.0f); /* :606-611: the grid potential at the cut-off */
                            if (p->rvdw_switch > 0.f) /* :613-625 potential switch on the (soft-cored) distance; constants :273-285 */
                            {
                                const float dsw_ = rvdw - p->rvdw_switch;
                                const float swV3 = -10.0f / (dsw_ * dsw_ * dsw_), swV4 = 15.0f / (dsw_ * dsw_ * dsw_ * dsw_),
                                            swV5 = -6.0f / (dsw_ * dsw_ * dsw_ * dsw_ * dsw_);
                                float d = rV - p->rvdw_switch;
                                d       = d > 0.f ? d : 0.f;
                                const float d2 = d * d;
                                const float sw = 1.0f + d2 * d * (swV3 + d * (swV4 + d * swV5));
                                const float dsw = d2 * (3.0f * swV3 + d * (4.0f * swV4 + d * 5.0f * swV5));
                                FscalV[i] = FscalV[i] * sw - rV * Vvdw[i] * dsw; /* rV < rvdw holds here */
                                Vvdw[i]   = Vvdw[i] * sw;
                            }
                        }
                        FscalC[i] *= rpinvC;
                        FscalV[i] *= rpinvV;
                    }
                }
                for (int i = 0; i < 2; i++) /* :644-667 */
                {
                    vctot += LFC[i] * Vcoul[i];
                    vvtot += LFV[i] * Vvdw[i];
                    Fscal += LFC[i] * FscalC[i] * rpm2;
                    Fscal += LFV[i] * FscalV[i] * rpm2;
                    if (useSoftCore)
                    {
                        dvdl_coul += Vcoul[i] * DLF[i] + LFC[i] * alpha_coul_eff * dlfac_coul[i] * FscalC[i] * sigma6[i];
                        dvdl_vdw += Vvdw[i] * DLF[i] + LFV[i] * alpha_vdw_eff * dlfac_vdw[i] * FscalV[i] * sigma6[i];
                    }
                    else
                    {
                        dvdl_coul += Vcoul[i] * DLF[i];
                        dvdl_vdw += Vvdw[i] * DLF[i];
                    }
                }
            }
            else if (!ewald) /* :669-691: excluded pair, reaction-field correction only (eelCUT or RF) */
            {
                const float FF = -2.0f * krf;
                float       VV = krf * rsq - crf;
                if (ii == jnr) VV *= 0.5f;
                for (int i = 0; i < 2; i++)
                {
                    vctot += LFC[i] * qq[i] * VV;
                    Fscal += LFC[i] * qq[i] * FF;
                    dvdl_coul += DLF[i] * qq[i] * VV;
                }
            }
            if (ewald && (r < rcoulomb || !included))
            {
                /* :693-737: the reciprocal-space part of the pair, subtracted unsoftened.  The reference interpolates erf(beta r)/r
                 * and its derivative from the cubic-spline table coulombEwaldTables (tableFDV0, spacing ~5e-4 nm); here they are
                 * evaluated directly -- the test against the reference kernel measures the difference (forces 1e-6) */
                float v_lr, f_lr;
                if (rsq > 0)
                {
                    const double br = (double)p->beta * r, er = erf(br);
                    v_lr = (float)(er / r);
                    f_lr = (float)((er / r - 2.0 * p->beta / 1.7724538509055159 * exp(-br * br)) / ((double)r * r)); /* -(dv/dr) / r */
                }
                else
                {
                    v_lr = (float)(2.0 * p->beta / 1.7724538509055159);
                    f_lr = 0.f;
                }
                if (ii == jnr) v_lr *= 0.5f;
                for (int i = 0; i < 2; i++)
                {
                    vctot -= LFC[i] * qq[i] * v_lr;
                    Fscal -= LFC[i] * qq[i] * f_lr;
                    dvdl_coul -= (DLF[i] * qq[i]) * v_lr;
                }
            }
            if (ljpme && r < rvdw)
            {
                /* :725-770: the grid (reciprocal-space) part of the dispersion, (1 - exp(-x)(1 + x + x^2/2)) / r^6 with x = (beta_lj r)^2
                 * (tables/forcetable.cpp v_lj_ewald_lr), taken off unsoftened, for excluded pairs and the atom with itself (half) too.
                 * The reference interpolates it and its derivative from the cubic-spline table vdwEwaldTables and divides by six; here
                 * evaluated directly in double, by its series where the closed form cancels (x < 0.1) */
                const double b2 = (double)p->beta_lj * p->beta_lj, xx = b2 * (double)r * r, ex = exp(-xx);
                double       v_lr, f_lr;
                if (rsq > 0)
                {
                    const double r2d = (double)r * r, r6d = r2d * r2d * r2d;
                    const double ser = xx * (1.0 / 4 + xx * (1.0 / 20 + xx * (1.0 / 120 + xx * (1.0 / 840 + xx * (1.0 / 6720 + xx * (1.0 / 60480 + xx / 604800.0))))));
                    const double g   = xx < 0.1 ? ex * xx * xx * xx / 6.0 * (1.0 + ser) : 1.0 - ex * (1.0 + xx + 0.5 * xx * xx);
                    v_lr = g / r6d;
                    /* -(dv/dr) / r = 6 g / r^8 - beta^6 exp(-x) / r^2 */
                    f_lr = xx < 0.1 ? b2 * b2 * b2 * ex * ser / r2d : 6.0 * g / (r6d * r2d) - b2 * b2 * b2 * ex / r2d;
                }
                else
                {
                    v_lr = b2 * b2 * b2 / 6.0;
                    f_lr = 0.0;
                }
                const float FF = (float)(f_lr / 6.0);
                float       VV = (float)(v_lr / 6.0);
                if (ii == jnr) VV *= 0.5f;
                for (int i = 0; i < 2; i++)
                {
                    const float c6g = c6grid[tj[i]];
                    vvtot += LFV[i] * c6g * VV;
                    Fscal += LFV[i] * c6g * FF;
                    dvdl_vdw += (DLF[i] * c6g) * VV;
                }
            }
            const float tx = Fscal * dx, ty = Fscal * dy, tz = Fscal * dz;
            fix += tx, fiy += ty, fiz += tz;
            f[3 * jnr] -= tx, f[3 * jnr + 1] -= ty, f[3 * jnr + 2] -= tz;
        }
        if (npair_within_cutoff > 0)
        {
            f[3 * ii] += fix, f[3 * ii + 1] += fiy, f[3 * ii + 2] += fiz;
            fshift[is3] += fix, fshift[is3 + 1] += fiy, fshift[is3 + 2] += fiz;
            Vc += vctot;
            Vv += vvtot;
        }
    }
    out4[0] = Vc, out4[1] = Vv, out4[2] = dvdl_coul, out4[3] = dvdl_vdw;
    free(c6grid);
}

/* ---- listed ("bonded") interactions: the types the reference runs on the GPU (listed_forces/gpubonded.h:84-85) ----
 * TEST INFRASTRUCTURE.  Restates listed_forces/gpubondedkernels.cu (bonds :93-141, angles :164-232, Urey-Bradley :234-335,
 * dih_angle / do_dih_fup :337-436, proper :438-481, Ryckaert-Bellemans :483-585, improper :600-656, 1-4 pairs :658-718) and the
 * minimum-image rule of pbcutil/pbc_aiuc_cuda.cuh:60-125: displacement and image decision in float, everything after it in double,
 * sums in double.  Pinned to the reference's CPU functions for the same types (bonded.cpp / pairs.cpp through oracle/_ref,
 * tests/test_oracle_cpu.py).  kind / iatoms / params6 as gmxref_bonded (oracle/ref_harness.h).  Shift forces as the reference's
 * kernels book them: the force on an atom goes to the shift of its image relative to the interaction's reference atom. */
#define ORC_CENTRAL 22
static int orc_pbc_dx(const float* b, const float* x1, const float* x2, double* dr)
{
    float d0 = x1[0] - x2[0], d1 = x1[1] - x2[1], d2 = x1[2] - x2[2];
    float shz = 0.f, shy = 0.f, shx = 0.f;
    if (b[8] > 0.f)
    {
        shz = rintf(d2 * (1.0f / b[8]));
        d0 -= shz * b[6], d1 -= shz * b[7], d2 -= shz * b[8];
    }
    if (b[4] > 0.f)
    {
        shy = rintf(d1 * (1.0f / b[4]));
        d0 -= shy * b[3], d1 -= shy * b[4];
    }
    if (b[0] > 0.f)
    {
        shx = rintf(d0 * (1.0f / b[0]));
        d0 -= shx * b[0];
    }
    dr[0] = d0, dr[1] = d1, dr[2] = d2;
    return 5 * (3 * (-(int)shz + 1) + (-(int)shy + 1)) + (-(int)shx + 2); /* pbcutil/ishift.h:50 */
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void   cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1], c[1] = a[2] * b[0] - a[0] * b[2], c[2] = a[0] * b[1] - a[1] * b[0];
}
static void add_f(double* f, double* fshift, int a, int s, const double* v, double sign)
{
    for (int d = 0; d < 3; d++)
    {
        f[3 * a + d] += sign * v[d];
        fshift[3 * s + d] += sign * v[d];
    }
}
/* forces of a torsion from dV/dphi: gpubondedkernels.cu:378-436 */
static void orc_dih_forces(const float* box9, const float* x, const int* a, double ddphi, const double* r_ij, const double* r_kj,
                           const double* r_kl, const double* m, const double* n, int t1, int t2, double* f, double* fshift)
{
    const double iprm = dot3(m, m), iprn = dot3(n, n), nrkj2 = dot3(r_kj, r_kj);
    const double toler = nrkj2 * 1.1920928955078125e-07; /* GMX_REAL_EPS, mixed precision */
    if (!(iprm > toler && iprn > toler)) return;
    const double nrkj = sqrt(nrkj2);
    double       f_i[3], f_l[3], f_j[3], f_k[3];
    const double ca = -ddphi * nrkj / iprm, cb = ddphi * nrkj / iprn;
    const double p = dot3(r_ij, r_kj) / nrkj2, q = dot3(r_kl, r_kj) / nrkj2;
    for (int d = 0; d < 3; d++)
    {
        f_i[d]         = ca * m[d];
        f_l[d]         = cb * n[d];
        const double s = p * f_i[d] - q * f_l[d];
        f_j[d]         = f_i[d] - s;
        f_k[d]         = f_l[d] + s;
    }
    double    dx_jl[3];
    const int t3 = orc_pbc_dx(box9, x + 3 * a[3], x + 3 * a[1], dx_jl);
    add_f(f, fshift, a[0], t1, f_i, 1.0);
    add_f(f, fshift, a[1], ORC_CENTRAL, f_j, -1.0);
    add_f(f, fshift, a[2], t2, f_k, -1.0);
    add_f(f, fshift, a[3], t3, f_l, 1.0);
}

int orc_bonded(int kind, int nbonds, const int* iatoms, const float* params6, int natoms, const float* x, const float* q,
               const float* box9, float epsfac_fudge, double* f, double* fshift, double* energy2)
{
    static const int nral[8] = { 2, 3, 3, 4, 4, 4, 4, 2 };
    const double     deg2rad = 3.14159265358979323846 / 180.0, pi = 3.14159265358979323846;
    if (kind < 0 || kind > 7) return -1;
    const int stride = nral[kind] + 1;
    for (int i = 0; i < nbonds; i++)
    {
        const int*   ia = iatoms + (size_t)stride * i;
        const float* p  = params6 + 6 * ia[0];
        const int*   a  = ia + 1;
        for (int k = 0; k < nral[kind]; k++)
            if (a[k] < 0 || a[k] >= natoms) return -2;
        if (kind == 0 || kind == 7)
        {
            double    dx[3];
            const int ki  = orc_pbc_dx(box9, x + 3 * a[0], x + 3 * a[1], dx);
            const double r2 = dot3(dx, dx);
            double       fs;
            if (kind == 0)
            {
                const double r = sqrt(r2), dr = r - p[0];
                energy2[0] += 0.5 * p[1] * dr * dr;
                if (r2 == 0.0) continue;
                fs = -p[1] * dr / r;
            }
            else
            {
                const double rinv2 = 1.0 / r2, rinv6 = rinv2 * rinv2 * rinv2, velec = (double)epsfac_fudge * q[a[0]] * q[a[1]] * sqrt(rinv2);
                fs = ((12.0 * p[1] * rinv6 - 6.0 * p[0]) * rinv6 + velec) * rinv2;
                energy2[0] += (p[1] * rinv6 - p[0]) * rinv6;
                energy2[1] += velec;
            }
            double fij[3] = { fs * dx[0], fs * dx[1], fs * dx[2] };
            add_f(f, fshift, a[0], ki, fij, 1.0);
            add_f(f, fshift, a[1], ORC_CENTRAL, fij, -1.0);
        }
        else if (kind == 1 || kind == 2)
        {
            double    r_ij[3], r_kj[3];
            const int t1 = orc_pbc_dx(box9, x + 3 * a[0], x + 3 * a[1], r_ij), t2 = orc_pbc_dx(box9, x + 3 * a[2], x + 3 * a[1], r_kj);
            const double nij2 = dot3(r_ij, r_ij), nkj2 = dot3(r_kj, r_kj);
            double       c = dot3(r_ij, r_kj) / sqrt(nij2 * nkj2);
            c              = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
            const double th = acos(c), dth = th - p[0] * deg2rad;
            energy2[0] += 0.5 * p[1] * dth * dth;
            const double c2 = c * c;
            if (c2 < 1.0)
            {
                const double st = -p[1] * dth / sqrt(1.0 - c2), sth = st * c;
                const double cik = st / sqrt(nij2 * nkj2), cii = sth / nij2, ckk = sth / nkj2;
                double       f_i[3], f_k[3], f_j[3];
                for (int d = 0; d < 3; d++)
                {
                    f_i[d] = -(cik * r_kj[d] - cii * r_ij[d]);
                    f_k[d] = -(cik * r_ij[d] - ckk * r_kj[d]);
                    f_j[d] = -f_i[d] - f_k[d];
                }
                add_f(f, fshift, a[0], t1, f_i, 1.0);
                add_f(f, fshift, a[1], ORC_CENTRAL, f_j, 1.0);
                add_f(f, fshift, a[2], t2, f_k, 1.0);
            }
            if (kind == 2) /* the 1-3 bond of Urey-Bradley */
            {
                double    r_ik[3];
                const int ki  = orc_pbc_dx(box9, x + 3 * a[0], x + 3 * a[2], r_ik);
                const double r2 = dot3(r_ik, r_ik), r = sqrt(r2), dr = r - p[2];
                energy2[0] += 0.5 * p[3] * dr * dr;
                if (r2 != 0.0)
                {
                    const double fs = -p[3] * dr / r;
                    double       fik[3] = { fs * r_ik[0], fs * r_ik[1], fs * r_ik[2] };
                    add_f(f, fshift, a[0], ki, fik, 1.0);
                    add_f(f, fshift, a[2], ORC_CENTRAL, fik, -1.0);
                }
            }
        }
        else
        {
            double    r_ij[3], r_kj[3], r_kl[3], m[3], n[3], mxn[3];
            const int t1 = orc_pbc_dx(box9, x + 3 * a[0], x + 3 * a[1], r_ij), t2 = orc_pbc_dx(box9, x + 3 * a[2], x + 3 * a[1], r_kj);
            (void)orc_pbc_dx(box9, x + 3 * a[2], x + 3 * a[3], r_kl);
            cross3(r_ij, r_kj, m);
            cross3(r_kj, r_kl, n);
            cross3(m, n, mxn);
            double phi = atan2(sqrt(dot3(mxn, mxn)), dot3(m, n)); /* gmx_angle */
            if (dot3(r_ij, n) < 0.0) phi = -phi;
            double ddphi;
            if (kind == 3 || kind == 6)
            {
                const double mult = (double)(int)p[2], mdphi = mult * phi - p[0] * deg2rad;
                energy2[0] += p[1] * (1.0 + cos(mdphi));
                ddphi = -p[1] * mult * sin(mdphi);
            }
            else if (kind == 4)
            {
                phi += phi < 0.0 ? pi : -pi; /* polymer convention */
                const double cp = cos(phi), sp = sin(phi);
                double       v = p[0], dd = 0.0, cf = 1.0;
                for (int k = 1; k < 6; k++)
                {
                    dd += k * p[k] * cf;
                    cf *= cp;
                    v += cf * p[k];
                }
                energy2[0] += v;
                ddphi = -dd * sp;
            }
            else
            {
                double dp = phi - p[0] * deg2rad;
                if (dp >= pi) dp -= 2.0 * pi;
                else if (dp < -pi) dp += 2.0 * pi;
                energy2[0] += 0.5 * p[1] * dp * dp;
                ddphi = p[1] * dp;
            }
            orc_dih_forces(box9, x, a, ddphi, r_ij, r_kj, r_kl, m, n, t1, t2, f, fshift);
        }
    }
    return 0;
}
