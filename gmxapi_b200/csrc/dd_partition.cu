/* b200nb: the coordinate-dependent work of a repartitioning pair-search step as device kernels -- which home atoms left
 * the domain, to which neighbour they go, the new home set in global-index order, which home atoms the neighbours need
 * as their halo, and the local topology (types, charges, exclusions in local indices) of home + halo.  The coordinates
 * never leave the device; the host sees a handful of counts.
 *
 * Reference behaviour replaced (paths relative to /root/reference/src/gromacs; all of it CPU code there):
 *   domdec/redistribute.cpp:505-760   dd_redistribute_cg: per atom the cell it moved to, move flags, compaction of the
 *                                     stayers, packing of the leavers per direction ("moved more than one cell" is fatal)
 *   pbcutil/pbc.cpp put_atoms_in_box  the wrap into the unit cell before the decision
 *   domdec/domdec.cpp:1900-2100 /     setup_dd_communication: the send lists = home atoms within the cut-off of a
 *   domdec/partition.cpp:2058-2440    domain face (get_zone_pulse_cgs), the index maps GpuHaloExchange::reinitHalo uploads
 *                                     (domdec/gpuhaloexchange_impl.cu:133-213)
 *   domdec/localtopology.cpp          the exclusions of the local atoms, renumbered from global to local indices
 *                                     (make_exclusions_zone), with the global -> local look-up of ga2la
 * Transport of the leavers / halo indices between ranks stays with the caller (torch.distributed on device tensors, or the
 * in-process loopback of the tests): these are search-step messages of a few thousand atoms.
 */
#include <algorithm>
#include <cstdio>
#include <vector>

#include "b200nb_internal.h"

#define PART_BLOCK 256
#define PART_MAX_CODES 32 /* the 27 neighbour offsets of an N-D grid + 'moved more than one domain' */

namespace
{

/* put_atoms_in_box for a rectangular box, in float32 like the reference: x - box while x >= box, x + box while x < 0 */
__device__ __forceinline__ float wrap1(float v, float b)
{
    while (v < 0.f) v += b;
    while (v >= b) v -= b;
    return v;
}

/* Slab decomposition along x: wraps the atom into the box and says where it belongs now: 0 it stays, 1 it goes to the -x
 * neighbour, 2 to the +x neighbour, 3 it moved more than one domain (fatal, as in redistribute.cpp).  bounds[k] = float32(k *
 * box_x / nranks): the owner is the slab whose [bounds[o], bounds[o+1]) holds x -- numpy searchsorted(bounds, x, "right") - 1,
 * clipped, bit for bit (gmxapi_b200/domdec.py DomainPlan.owner_of). */
__global__ void k_wrap_classify(float* __restrict__ x, int n, float bx, float by, float bz, const float* __restrict__ bounds, int nranks, int rank,
                                int* __restrict__ code)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const float xx = wrap1(x[3 * a], bx), yy = wrap1(x[3 * a + 1], by), zz = wrap1(x[3 * a + 2], bz);
    x[3 * a] = xx, x[3 * a + 1] = yy, x[3 * a + 2] = zz;
    int o = (int)(xx * ((float)nranks / bx));
    o     = max(0, min(o, nranks - 1));
    while (o > 0 && xx < bounds[o]) o--;
    while (o < nranks - 1 && xx >= bounds[o + 1]) o++;
    const int left = (rank + nranks - 1) % nranks, right = (rank + 1) % nranks;
    int       c = 3;
    if (o == rank) c = 0;
    else if (o == left) c = 1; /* two ranks: both faces lead to the same neighbour, everything that leaves goes "left" */
    else if (o == right) c = 2;
    code[a] = c;
}

/* halo selection of the slab decomposition: home atoms within rlist of the lower face go to the -x neighbour
 * (DomainPlan._send_list: x - lo < rlist in float32) */
__global__ void k_select_lower_face(const float* __restrict__ x, int n, float lo, float rlist, int* __restrict__ code)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n) code[a] = __fsub_rn(x[3 * a], lo) < rlist ? 1 : 0;
}

/* N-D grids (gmxapi_b200/domdec_nd.py): per dimension the owner cell of the wrapped coordinate, as the offset -1 / 0 / +1 from
 * this rank's cell (periodic; with two cells along a dimension the other one is offset +1); code = 9 (ox+1) + 3 (oy+1) + (oz+1),
 * 13 = stays, 27 = moved more than one domain.  Cell boundaries are float32(i * box_d / n_d), owner = searchsorted(bounds, x,
 * "right") - 1 clipped: DomainPlanND.owner_of, bit for bit. */
struct NdGeom
{
    float box[3];
    int   grid[3], coords[3];
};
__device__ __forceinline__ int owner_cell(float v, float b, int n)
{
    if (n == 1) return 0;
    const double w = (double)b / n;
    int          o = max(0, min((int)(v * ((float)n / b)), n - 1));
    while (o > 0 && v < (float)(o * w)) o--;
    while (o < n - 1 && v >= (float)((o + 1) * w)) o++;
    return o;
}
__global__ void k_wrap_classify_nd(float* __restrict__ x, int n, NdGeom G, int* __restrict__ code)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    int c = 0;
    bool lost = false;
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
        const float v = wrap1(x[3 * a + d], G.box[d]);
        x[3 * a + d]  = v;
        const int nd = G.grid[d], o = owner_cell(v, G.box[d], nd);
        int       off = 0;
        if (o != G.coords[d])
        {
            const int fwd = (o - G.coords[d] + nd) % nd, bwd = (G.coords[d] - o + nd) % nd;
            if (fwd == 1) off = 1; /* also the only other cell of a two-cell dimension */
            else if (bwd == 1) off = -1;
            else lost = true;
        }
        c = 3 * c + (off + 1);
    }
    code[a] = lost ? 27 : c;
}
/* boundary atoms of this domain for the rank that sees it at half-shell offset o (DomainPlanND.boundary_atoms): within rlist of
 * the lower face along dimensions with o = +1 (x - lo < r), of the upper face where o = -1 (hi - x <= r), float32 */
__global__ void k_select_boundary(const float* __restrict__ x, int n, float lo0, float lo1, float lo2, float hi0, float hi1, float hi2, int o0, int o1,
                                  int o2, float rlist, int* __restrict__ code)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const float lo[3] = { lo0, lo1, lo2 }, hi[3] = { hi0, hi1, hi2 };
    const int   o[3]  = { o0, o1, o2 };
    bool        m     = true;
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
        const float v = x[3 * a + d];
        if (o[d] == 1) m = m && (__fsub_rn(v, lo[d]) < rlist);
        else if (o[d] == -1) m = m && (__fsub_rn(hi[d], v) <= rlist);
    }
    code[a] = m ? 1 : 0;
}

/* ---- stable multi-way partition of the indices 0..n-1 by code: counts per block, scan over the blocks, scatter ---- */
__global__ void __launch_bounds__(PART_BLOCK)
k_part_count(const int* __restrict__ code, int n, int ncodes, int* __restrict__ block_count)
{
    __shared__ int cnt[PART_MAX_CODES];
    if (threadIdx.x < PART_MAX_CODES) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int a = blockIdx.x * PART_BLOCK + threadIdx.x;
    const int c = a < n ? code[a] : -1;
    for (int k = 0; k < ncodes; k++)
    {
        const unsigned b = __ballot_sync(0xffffffffu, c == k);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&cnt[k], __popc(b));
    }
    __syncthreads();
    if (threadIdx.x < ncodes) block_count[threadIdx.x * gridDim.x + blockIdx.x] = cnt[threadIdx.x];
}
/* one thread per code walks its row of block counts: row -> exclusive offsets, total[k] = atoms with code k; base[k] = where the
 * list of code k starts in the output (lists are concatenated in code order) */
__global__ void k_part_scan(int* __restrict__ block_count, int nblk, int ncodes, int* __restrict__ total)
{
    __shared__ int tot[PART_MAX_CODES];
    const int k = threadIdx.x;
    if (k < ncodes)
    {
        int run = 0;
        for (int b = 0; b < nblk; b++)
        {
            const int v             = block_count[k * nblk + b];
            block_count[k * nblk + b] = run;
            run += v;
        }
        tot[k]   = run;
        total[k] = run;
    }
    __syncthreads();
    if (k < ncodes)
    {
        int base = 0;
        for (int j = 0; j < k; j++) base += tot[j];
        total[PART_MAX_CODES + k] = base;
    }
}
__global__ void __launch_bounds__(PART_BLOCK)
k_part_scatter(const int* __restrict__ code, int n, int ncodes, const int* __restrict__ block_offset, const int* __restrict__ total, int* __restrict__ out)
{
    __shared__ int wcnt[PART_MAX_CODES][PART_BLOCK / 32];
    const int a = blockIdx.x * PART_BLOCK + threadIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = a < n ? code[a] : -1;
    int       rank_in_warp = 0;
    for (int k = 0; k < ncodes; k++)
    {
        const unsigned b = __ballot_sync(0xffffffffu, c == k);
        if (lane == 0) wcnt[k][w] = __popc(b);
        if (c == k) rank_in_warp = __popc(b & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (c >= 0 && c < ncodes)
    {
        int before = 0;
        for (int j = 0; j < w; j++) before += wcnt[c][j];
        out[total[PART_MAX_CODES + c] + block_offset[c * gridDim.x + blockIdx.x] + before + rank_in_warp] = a;
    }
}

/* leavers as the messages the ranks exchange: 4 ints per atom {global index, x, y, z bits} */
__global__ void k_pack_atoms(const int* __restrict__ idx, int m, const int* __restrict__ gid, const float* __restrict__ x, int* __restrict__ out4)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int a      = idx[k];
    out4[4 * k]     = gid[a];
    out4[4 * k + 1] = __float_as_int(x[3 * a]);
    out4[4 * k + 2] = __float_as_int(x[3 * a + 1]);
    out4[4 * k + 3] = __float_as_int(x[3 * a + 2]);
}

/* The new home set = stayers (already ascending in global index: the home set is kept sorted and the partition is stable)
 * merged with the arrivals (a few thousand, any order).  First the arrivals are put in order by ranking (every arrival counts
 * the arrivals with a smaller index: they are few); then every atom computes its own position with one binary search -- a stayer
 * its rank among the stayers plus the arrivals with a smaller global index, an arrival its rank among the arrivals plus the
 * stayers with a smaller index. */
__global__ void k_rank_arrivals(const int* __restrict__ arr4, int m, int* __restrict__ sorted4)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int g = arr4[4 * j];
    int       r = 0;
    for (int i = 0; i < m; i++) r += arr4[4 * i] < g || (arr4[4 * i] == g && i < j);
    reinterpret_cast<int4*>(sorted4)[r] = reinterpret_cast<const int4*>(arr4)[j];
}
__global__ void k_merge_home(const int* __restrict__ stay_idx, int ns, const int* __restrict__ gid, const float* __restrict__ x, const int* __restrict__ sorted4,
                             int m, int* __restrict__ gid_out, float* __restrict__ x_out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ns)
    {
        const int a = stay_idx[t], g = gid[a];
        int       lo = 0, hi = m; /* arrivals with a smaller global index */
        while (lo < hi)
        {
            const int mid = (lo + hi) >> 1;
            if (sorted4[4 * mid] < g) lo = mid + 1;
            else hi = mid;
        }
        const int p   = t + lo;
        gid_out[p]    = g;
        x_out[3 * p]  = x[3 * a], x_out[3 * p + 1] = x[3 * a + 1], x_out[3 * p + 2] = x[3 * a + 2];
    }
    else if (t < ns + m)
    {
        const int j = t - ns, g = sorted4[4 * j];
        int       lo = 0, hi = ns; /* stayers with a smaller global index */
        while (lo < hi)
        {
            const int mid = (lo + hi) >> 1;
            if (gid[stay_idx[mid]] < g) lo = mid + 1;
            else hi = mid;
        }
        const int p  = j + lo;
        gid_out[p]   = g;
        x_out[3 * p] = __int_as_float(sorted4[4 * j + 1]), x_out[3 * p + 1] = __int_as_float(sorted4[4 * j + 2]), x_out[3 * p + 2] = __int_as_float(sorted4[4 * j + 3]);
    }
}

__global__ void k_gather_int(const int* __restrict__ idx, int m, const int* __restrict__ in, int* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) out[k] = in[idx[k]];
}

/* ---- local topology: ga2la look-up, types / charges of the local atoms, exclusions renumbered to local indices ---- */
__global__ void k_g2l_fill(const int* __restrict__ local_gid, int nlocal, int nglobal, int* __restrict__ g2l, int* __restrict__ err)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nlocal) return;
    const int g = local_gid[k];
    if (g < 0 || g >= nglobal) { atomicExch(err, 1); return; }
    g2l[g] = k;
}
__global__ void k_topo_count(const int* __restrict__ local_gid, int nlocal, const int* __restrict__ g2l, const int* __restrict__ type_g, const float* __restrict__ q_g,
                             const int* __restrict__ eoff_g, const int* __restrict__ eidx_g, int* __restrict__ type_l, float* __restrict__ q_l, int* __restrict__ cnt)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nlocal) return;
    const int g = local_gid[k];
    type_l[k]   = type_g[g];
    q_l[k]      = q_g[g];
    int c       = 0;
    for (int e = eoff_g[g]; e < eoff_g[g + 1]; e++) c += g2l[eidx_g[e]] >= 0; /* partners that are not local cannot be in range here */
    cnt[k] = c;
}
__global__ void k_topo_fill(const int* __restrict__ local_gid, int nlocal, const int* __restrict__ g2l, const int* __restrict__ eoff_g, const int* __restrict__ eidx_g,
                            const int* __restrict__ off_l, int* __restrict__ idx_l)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nlocal) return;
    const int g = local_gid[k];
    int       o = off_l[k];
    for (int e = eoff_g[g]; e < eoff_g[g + 1]; e++)
    {
        const int l = g2l[eidx_g[e]];
        if (l >= 0) idx_l[o++] = l;
    }
}
__global__ void k_g2l_clear(const int* __restrict__ local_gid, int nlocal, int* __restrict__ g2l)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nlocal) g2l[local_gid[k]] = -1;
}

/* exclusive scan of n ints in place (n up to a few million): block sums, one thread over the block sums, block-local scan */
#define SCAN_BLOCK 1024
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_block_sums(const int* __restrict__ v, int n, int* __restrict__ bsum)
{
    __shared__ int ws[32];
    const int a = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    int       s = a < n ? v[a] : 0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        int t = ws[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) bsum[blockIdx.x] = t;
    }
}
__global__ void k_scan_sums(int* __restrict__ bsum, int nblk, int* __restrict__ total)
{
    if (threadIdx.x || blockIdx.x) return;
    int run = 0;
    for (int b = 0; b < nblk; b++)
    {
        const int v = bsum[b];
        bsum[b]     = run;
        run += v;
    }
    *total = run;
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(int* __restrict__ v, int n, const int* __restrict__ bsum, const int* __restrict__ total)
{
    __shared__ int ws[32];
    const int a = blockIdx.x * SCAN_BLOCK + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int x = a < n ? v[a] : 0;
    int       s = x;
    for (int o = 1; o < 32; o <<= 1)
    {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) ws[w] = s;
    __syncthreads();
    if (w == 0)
    {
        int t = ws[lane];
        for (int o = 1; o < 32; o <<= 1)
        {
            const int u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        ws[lane] = t;
    }
    __syncthreads();
    const int excl = s - x + (w ? ws[w - 1] : 0) + bsum[blockIdx.x];
    if (a < n) v[a] = excl;
    if (a == n) v[a] = *total; /* the CSR's closing offset */
}

int ensure_scratch(b200nb_context* h, size_t ints)
{
    DdState& D = h->dd;
    if (ints <= D.cap_part_scratch && D.d_part_scratch) return 0;
    cudaFree(D.d_part_scratch);
    D.d_part_scratch = nullptr;
    NB_CUDA(h, cudaMalloc((void**)&D.d_part_scratch, sizeof(int) * ints));
    D.cap_part_scratch = ints;
    return 0;
}

} // namespace

#define PART_LAUNCH_CHECK(h)            \
    do                                  \
    {                                   \
        (h)->nlaunches++;               \
        NB_CUDA(h, cudaGetLastError()); \
    } while (0)

extern "C" int b200nb_dd_wrap_classify(b200nb_t* h, float* x_dev, int n, const float box[3], const float* bounds_host, int nranks, int rank, int* code_dev)
{
    if (!h || !x_dev || !code_dev || !box || !bounds_host || n < 0 || nranks < 1 || rank < 0 || rank >= nranks)
        return nb_fail(h, B200NB_ERR_ARG, "dd_wrap_classify: bad argument");
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    if (ensure_scratch(h, (size_t)nranks + 1 + 64)) return B200NB_ERR_CUDA;
    float* d_bounds = reinterpret_cast<float*>(h->dd.d_part_scratch);
    NB_CUDA(h, cudaMemcpyAsync(d_bounds, bounds_host, sizeof(float) * (nranks + 1), cudaMemcpyHostToDevice, h->stream));
    k_wrap_classify<<<(n + 255) / 256, 256, 0, h->stream>>>(x_dev, n, box[0], box[1], box[2], d_bounds, nranks, rank, code_dev);
    PART_LAUNCH_CHECK(h);
    NB_CUDA(h, cudaStreamSynchronize(h->stream)); /* bounds_host may be a temporary of the caller */
    return 0;
}

extern "C" int b200nb_dd_select_lower_face(b200nb_t* h, const float* x_dev, int n, float lo, float rlist, int* code_dev)
{
    if (!h || !x_dev || !code_dev || n < 0) return nb_fail(h, B200NB_ERR_ARG, "dd_select_lower_face: bad argument");
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    k_select_lower_face<<<(n + 255) / 256, 256, 0, h->stream>>>(x_dev, n, lo, rlist, code_dev);
    PART_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int b200nb_dd_wrap_classify_nd(b200nb_t* h, float* x_dev, int n, const float box[3], const int grid[3], const int coords[3], int* code_dev)
{
    if (!h || !x_dev || !code_dev || !box || !grid || !coords || n < 0) return nb_fail(h, B200NB_ERR_ARG, "dd_wrap_classify_nd: bad argument");
    NdGeom G;
    for (int d = 0; d < 3; d++)
    {
        if (grid[d] < 1 || coords[d] < 0 || coords[d] >= grid[d] || !(box[d] > 0)) return nb_fail(h, B200NB_ERR_ARG, "dd_wrap_classify_nd: bad grid");
        G.box[d] = box[d], G.grid[d] = grid[d], G.coords[d] = coords[d];
    }
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    k_wrap_classify_nd<<<(n + 255) / 256, 256, 0, h->stream>>>(x_dev, n, G, code_dev);
    PART_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int b200nb_dd_select_boundary(b200nb_t* h, const float* x_dev, int n, const float lo[3], const float hi[3], const int offset[3], float rlist,
                                         int* code_dev)
{
    if (!h || !x_dev || !code_dev || !lo || !hi || !offset || n < 0) return nb_fail(h, B200NB_ERR_ARG, "dd_select_boundary: bad argument");
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    k_select_boundary<<<(n + 255) / 256, 256, 0, h->stream>>>(x_dev, n, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], offset[0], offset[1], offset[2], rlist,
                                                             code_dev);
    PART_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int b200nb_dd_partition_indices(b200nb_t* h, const int* code_dev, int n, int ncodes, int* idx_dev, int* counts_host)
{
    if (!h || !code_dev || !idx_dev || !counts_host || n < 0 || ncodes < 1 || ncodes > PART_MAX_CODES)
        return nb_fail(h, B200NB_ERR_ARG, "dd_partition_indices: bad argument");
    for (int k = 0; k < ncodes; k++) counts_host[k] = 0;
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    const int nblk = (n + PART_BLOCK - 1) / PART_BLOCK;
    if (ensure_scratch(h, (size_t)PART_MAX_CODES * nblk + 2 * PART_MAX_CODES)) return B200NB_ERR_CUDA;
    int* d_blk = h->dd.d_part_scratch;
    int* d_tot = d_blk + (size_t)PART_MAX_CODES * nblk;
    k_part_count<<<nblk, PART_BLOCK, 0, h->stream>>>(code_dev, n, ncodes, d_blk);
    PART_LAUNCH_CHECK(h);
    k_part_scan<<<1, PART_MAX_CODES, 0, h->stream>>>(d_blk, nblk, ncodes, d_tot);
    PART_LAUNCH_CHECK(h);
    k_part_scatter<<<nblk, PART_BLOCK, 0, h->stream>>>(code_dev, n, ncodes, d_blk, d_tot, idx_dev);
    PART_LAUNCH_CHECK(h);
    NB_CUDA(h, cudaMemcpyAsync(counts_host, d_tot, sizeof(int) * ncodes, cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200nb_dd_pack_atoms(b200nb_t* h, const int* idx_dev, int m, const int* gid_dev, const float* x_dev, int* out4_dev)
{
    if (!h || m < 0 || (m && (!idx_dev || !gid_dev || !x_dev || !out4_dev))) return nb_fail(h, B200NB_ERR_ARG, "dd_pack_atoms: bad argument");
    if (m == 0) return 0;
    cudaSetDevice(h->device);
    k_pack_atoms<<<(m + 255) / 256, 256, 0, h->stream>>>(idx_dev, m, gid_dev, x_dev, out4_dev);
    PART_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int b200nb_dd_merge_home(b200nb_t* h, const int* stay_idx_dev, int nstay, const int* gid_dev, const float* x_dev, const int* arrived4_dev,
                                    int narrived, int* gid_out_dev, float* x_out_dev)
{
    if (!h || nstay < 0 || narrived < 0 || (nstay && (!stay_idx_dev || !gid_dev || !x_dev)) || (narrived && !arrived4_dev)
        || ((nstay + narrived) && (!gid_out_dev || !x_out_dev)))
        return nb_fail(h, B200NB_ERR_ARG, "dd_merge_home: bad argument");
    const int n = nstay + narrived;
    if (n == 0) return 0;
    cudaSetDevice(h->device);
    if (ensure_scratch(h, 4 * (size_t)narrived + 16)) return B200NB_ERR_CUDA;
    int* sorted4 = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(h->dd.d_part_scratch) + 15) & ~(uintptr_t)15);
    if (narrived)
    {
        k_rank_arrivals<<<(narrived + 255) / 256, 256, 0, h->stream>>>(arrived4_dev, narrived, sorted4);
        PART_LAUNCH_CHECK(h);
    }
    k_merge_home<<<(n + 255) / 256, 256, 0, h->stream>>>(stay_idx_dev, nstay, gid_dev, x_dev, sorted4, narrived, gid_out_dev, x_out_dev);
    PART_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int b200nb_dd_gather_int(b200nb_t* h, const int* idx_dev, int m, const int* in_dev, int* out_dev)
{
    if (!h || m < 0 || (m && (!idx_dev || !in_dev || !out_dev))) return nb_fail(h, B200NB_ERR_ARG, "dd_gather_int: bad argument");
    if (m == 0) return 0;
    cudaSetDevice(h->device);
    k_gather_int<<<(m + 255) / 256, 256, 0, h->stream>>>(idx_dev, m, in_dev, out_dev);
    PART_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int b200nb_dd_set_global_topology(b200nb_t* h, int nglobal, const int* type_host, const float* q_host, const int* excl_off_host,
                                             const int* excl_idx_host)
{
    if (!h || nglobal < 1 || !type_host || !q_host) return nb_fail(h, B200NB_ERR_ARG, "dd_set_global_topology: bad argument");
    if (!h->have_params) return nb_fail(h, B200NB_ERR_STATE, "dd_set_global_topology: set_params first");
    for (int a = 0; a < nglobal; a++)
        if (type_host[a] < 0 || type_host[a] >= h->hp.ntypes) return nb_fail(h, B200NB_ERR_ARG, "dd_set_global_topology: atom type out of range");
    cudaSetDevice(h->device);
    DdState& D = h->dd;
    std::vector<int> zero(nglobal + 1, 0);
    const int* off  = excl_off_host ? excl_off_host : zero.data();
    const int  nidx = off[nglobal];
    if (off[0] != 0) return nb_fail(h, B200NB_ERR_ARG, "dd_set_global_topology: excl_off[0] != 0");
    for (int k = 0; k < nidx; k++)
        if (excl_idx_host[k] < 0 || excl_idx_host[k] >= nglobal) return nb_fail(h, B200NB_ERR_ARG, "dd_set_global_topology: exclusion index out of range");
    cudaFree(D.d_gtype), cudaFree(D.d_gq), cudaFree(D.d_geoff), cudaFree(D.d_geidx), cudaFree(D.d_g2l);
    D.d_gtype = D.d_geoff = D.d_geidx = D.d_g2l = nullptr;
    D.d_gq                                     = nullptr;
    NB_CUDA(h, cudaMalloc((void**)&D.d_gtype, sizeof(int) * nglobal));
    NB_CUDA(h, cudaMalloc((void**)&D.d_gq, sizeof(float) * nglobal));
    NB_CUDA(h, cudaMalloc((void**)&D.d_geoff, sizeof(int) * (nglobal + 1)));
    NB_CUDA(h, cudaMalloc((void**)&D.d_geidx, sizeof(int) * std::max(nidx, 1)));
    NB_CUDA(h, cudaMalloc((void**)&D.d_g2l, sizeof(int) * nglobal));
    NB_CUDA(h, cudaMemcpy(D.d_gtype, type_host, sizeof(int) * nglobal, cudaMemcpyHostToDevice));
    NB_CUDA(h, cudaMemcpy(D.d_gq, q_host, sizeof(float) * nglobal, cudaMemcpyHostToDevice));
    NB_CUDA(h, cudaMemcpy(D.d_geoff, off, sizeof(int) * (nglobal + 1), cudaMemcpyHostToDevice));
    if (nidx) NB_CUDA(h, cudaMemcpy(D.d_geidx, excl_idx_host, sizeof(int) * nidx, cudaMemcpyHostToDevice));
    NB_CUDA(h, cudaMemset(D.d_g2l, 0xff, sizeof(int) * nglobal)); /* -1 everywhere: set_local_atoms restores that after itself */
    D.nglobal = nglobal;
    return 0;
}

int nb_install_atoms_dev(b200nb_context* h, int natoms, int nexcl); /* b200nb.cu: set_atoms' allocations for device-built arrays */

extern "C" int b200nb_dd_set_local_atoms(b200nb_t* h, const int* local_gid_dev, int nlocal)
{
    if (!h || !local_gid_dev || nlocal < 1) return nb_fail(h, B200NB_ERR_ARG, "dd_set_local_atoms: bad argument");
    DdState& D = h->dd;
    if (!D.d_g2l) return nb_fail(h, B200NB_ERR_STATE, "dd_set_local_atoms: dd_set_global_topology first");
    cudaSetDevice(h->device);
    const int nblk = (nlocal + 1 + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (ensure_scratch(h, (size_t)nblk + 8)) return B200NB_ERR_CUDA;
    int* d_bsum = D.d_part_scratch;
    int* d_tot  = d_bsum + nblk;
    int* d_err  = d_tot + 1;
    NB_CUDA(h, cudaMemsetAsync(d_err, 0, sizeof(int), h->stream));
    /* types, charges and the exclusion counts go straight into the context's atom arrays */
    int rc;
    if ((rc = nb_install_atoms_dev(h, nlocal, -1))) return rc;
    const unsigned nb = (unsigned)((nlocal + 255) / 256);
    k_g2l_fill<<<nb, 256, 0, h->stream>>>(local_gid_dev, nlocal, D.nglobal, D.d_g2l, d_err);
    PART_LAUNCH_CHECK(h);
    int err = 0;
    NB_CUDA(h, cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (err) return nb_fail(h, B200NB_ERR_ARG, "dd_set_local_atoms: global index out of range");
    k_topo_count<<<nb, 256, 0, h->stream>>>(local_gid_dev, nlocal, D.d_g2l, D.d_gtype, D.d_gq, D.d_geoff, D.d_geidx, h->d_type, h->d_q, h->d_excl_off);
    PART_LAUNCH_CHECK(h);
    k_scan_block_sums<<<nblk, SCAN_BLOCK, 0, h->stream>>>(h->d_excl_off, nlocal, d_bsum);
    PART_LAUNCH_CHECK(h);
    k_scan_sums<<<1, 32, 0, h->stream>>>(d_bsum, nblk, d_tot);
    PART_LAUNCH_CHECK(h);
    k_scan_apply<<<nblk, SCAN_BLOCK, 0, h->stream>>>(h->d_excl_off, nlocal, d_bsum, d_tot);
    PART_LAUNCH_CHECK(h);
    int nexcl = 0;
    NB_CUDA(h, cudaMemcpyAsync(&nexcl, d_tot, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    if ((rc = nb_install_atoms_dev(h, nlocal, nexcl))) return rc;
    k_topo_fill<<<nb, 256, 0, h->stream>>>(local_gid_dev, nlocal, D.d_g2l, D.d_geoff, D.d_geidx, h->d_excl_off, h->d_excl_idx);
    PART_LAUNCH_CHECK(h);
    k_g2l_clear<<<nb, 256, 0, h->stream>>>(local_gid_dev, nlocal, D.d_g2l);
    PART_LAUNCH_CHECK(h);
    return 0;
}
