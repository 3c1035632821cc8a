/* b200nb: the listed ("bonded") interactions the reference runs on the GPU, on the nonbonded module's own buffers.
 *
 * Replaces listed_forces/gpubonded.h:84-85 (fTypesOnGpu: bonds, angles, Urey-Bradley, proper / Ryckaert-Bellemans / improper /
 * periodic improper dihedrals, 1-4 pairs), gpubonded_impl.cu:178-310 (updateInteractionListsAndDeviceBuffers: the lists converted to
 * the nonbonded atom order on the host at every search step) and gpubondedkernels.cu:721-821 (one fused kernel, a thread per
 * interaction).  As there the kernel reads the nonbonded xq buffer in grid order and adds into the nonbonded force buffer, so the
 * bonded forces ride along with the nonbonded ones through the un-sort and the copy back; shift forces go to the same replicas,
 * energies to one accumulator per type (+ the Coulomb part of the 1-4 pairs).
 * Differences by design: the lists stay in ATOM order on the device and every thread maps its atoms through slot_of_atom, so a
 * search step costs the bonded module nothing (no host pass over the lists, no re-upload); each type's thread range is padded to
 * whole warps, so a warp evaluates one type and the energy reduces by shuffles; the shift-force accumulators of a block live in
 * shared memory and only the touched ones go out.  The central shift carries no virial and is not accumulated (as in the
 * cluster-pair kernels).
 * Minimum image as pbcutil/pbc_aiuc_cuda.cuh:60-125 (z, y, x in turn, general triclinic). */
#include <cstdio>
#include <cstring>
#include <vector>

#include "b200nb_internal.h"

namespace
{

struct BondedPbc
{
    float inv_zz, zx, zy, zz, inv_yy, yx, yy, inv_xx, xx;
};

struct BondedDev
{
    int          first[B200NB_BONDED_KINDS + 1]; /* thread ranges, each padded to whole warps */
    int          count[B200NB_BONDED_KINDS];
    const int*   iatoms[B200NB_BONDED_KINDS];
    const float* params[B200NB_BONDED_KINDS];
    BondedPbc    pbc;
    float        scale14; /* epsfac * fudgeQQ */
};

__device__ __forceinline__ float3 operator*(float s, float3 v) { return make_float3(s * v.x, s * v.y, s * v.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float  dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(float3 a, float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

/* r1 - r2 at the minimum image; returns the shift index of r1's image relative to r2 */
__device__ __forceinline__ int pbc_dx(const BondedPbc& P, const float4 r1, const float4 r2, float3& d)
{
    d               = make_float3(r1.x - r2.x, r1.y - r2.y, r1.z - r2.z);
    const float shz = rintf(d.z * P.inv_zz);
    d.x -= shz * P.zx, d.y -= shz * P.zy, d.z -= shz * P.zz;
    const float shy = rintf(d.y * P.inv_yy);
    d.x -= shy * P.yx, d.y -= shy * P.yy;
    const float shx = rintf(d.x * P.inv_xx);
    d.x -= shx * P.xx;
    return 5 * (3 * (1 - (int)shz) + (1 - (int)shy)) + (2 - (int)shx); /* pbcutil/ishift.h:50 of (-shx, -shy, -shz) */
}

template<bool VE>
__device__ __forceinline__ void add_force(float4* __restrict__ f, float* s_fs, int slot, int shift, float3 v)
{
    atomicAdd(f + slot, make_float4(v.x, v.y, v.z, 0.f));
    if (VE && shift != B200NB_CENTRAL)
    {
        atomicAdd(s_fs + 3 * shift, v.x);
        atomicAdd(s_fs + 3 * shift + 1, v.y);
        atomicAdd(s_fs + 3 * shift + 2, v.z);
    }
}

/* the four forces of a torsion from dV/dphi */
template<bool VE>
__device__ __forceinline__ void torsion_forces(const BondedPbc& P, const float4* x4, const int* s, float ddphi, float3 r_ij, float3 r_kj, float3 r_kl,
                                               float3 m, float3 n, int t1, int t2, float4* __restrict__ f, float* s_fs)
{
    const float iprm = dot(m, m), iprn = dot(n, n), nrkj2 = dot(r_kj, r_kj);
    const float toler = nrkj2 * 1.1920928955078125e-07f;
    if (!(iprm > toler && iprn > toler)) return;
    const float  nrkj_1 = rsqrtf(nrkj2), nrkj_2 = nrkj_1 * nrkj_1, nrkj = nrkj2 * nrkj_1;
    const float3 f_i = (-ddphi * nrkj / iprm) * m, f_l = (ddphi * nrkj / iprn) * n;
    const float  p = dot(r_ij, r_kj) * nrkj_2, q = dot(r_kl, r_kj) * nrkj_2;
    const float3 sv = p * f_i - q * f_l;
    const float3 f_j = f_i - sv, f_k = f_l + sv;
    int          t3 = B200NB_CENTRAL;
    if (VE)
    {
        float3 dx_jl;
        t3 = pbc_dx(P, x4[3], x4[1], dx_jl);
    }
    add_force<VE>(f, s_fs, s[0], t1, f_i);
    add_force<VE>(f, s_fs, s[1], B200NB_CENTRAL, -f_j);
    add_force<VE>(f, s_fs, s[2], t2, -f_k);
    add_force<VE>(f, s_fs, s[3], t3, f_l);
}

template<bool VE>
__global__ void __launch_bounds__(128)
k_bonded(const __grid_constant__ BondedDev B, const float4* __restrict__ xq, const int* __restrict__ slot_of_atom, float4* __restrict__ f,
         float* __restrict__ fshift, double* __restrict__ energy)
{
    __shared__ float s_fs[B200NB_SHIFTS * 3];
    constexpr float  deg2rad = 0.017453292519943295f, pi = 3.14159265358979323846f;
    if (VE)
    {
        for (int k = threadIdx.x; k < B200NB_SHIFTS * 3; k += blockDim.x) s_fs[k] = 0.f;
        __syncthreads();
    }
    const int tid  = blockIdx.x * blockDim.x + threadIdx.x;
    int       kind = -1;
#pragma unroll
    for (int k = 0; k < B200NB_BONDED_KINDS; k++)
        if (tid >= B.first[k] && tid < B.first[k + 1]) kind = k;
    float e0 = 0.f, e1 = 0.f;
    const int i = kind >= 0 ? tid - B.first[kind] : 0;
    if (kind >= 0 && i < B.count[kind])
    {
        const int    nral = (kind == B200NB_BONDED_BONDS || kind == B200NB_BONDED_LJ14) ? 2 : (kind <= B200NB_BONDED_UREY_BRADLEY ? 3 : 4);
        const int*   ia   = B.iatoms[kind] + (size_t)(nral + 1) * i;
        const float* p    = B.params[kind] + 6 * ia[0];
        int          s[4];
        float4       x4[4];
        for (int k = 0; k < nral; k++)
        {
            s[k]  = slot_of_atom ? slot_of_atom[ia[1 + k]] : ia[1 + k];
            x4[k] = xq[s[k]];
        }
        if (kind == B200NB_BONDED_BONDS)
        {
            float3      dx;
            const int   ki = pbc_dx(B.pbc, x4[0], x4[1], dx);
            const float r2 = dot(dx, dx), r = sqrtf(r2), dr = r - p[0];
            if (VE) e0 = 0.5f * p[1] * dr * dr;
            if (r2 != 0.f)
            {
                const float3 fij = (-p[1] * dr * rsqrtf(r2)) * dx;
                add_force<VE>(f, s_fs, s[0], ki, fij);
                add_force<VE>(f, s_fs, s[1], B200NB_CENTRAL, -fij);
            }
        }
        else if (kind == B200NB_BONDED_LJ14)
        {
            float3      dx;
            const int   ki    = pbc_dx(B.pbc, x4[0], x4[1], dx);
            const float rinv  = rsqrtf(dot(dx, dx)), rinv2 = rinv * rinv, rinv6 = rinv2 * rinv2 * rinv2;
            const float velec = B.scale14 * x4[0].w * x4[1].w * rinv;
            const float3 fij  = (((12.0f * p[1] * rinv6 - 6.0f * p[0]) * rinv6 + velec) * rinv2) * dx;
            add_force<VE>(f, s_fs, s[0], ki, fij);
            add_force<VE>(f, s_fs, s[1], B200NB_CENTRAL, -fij);
            if (VE) e0 = (p[1] * rinv6 - p[0]) * rinv6, e1 = velec;
        }
        else if (nral == 3)
        {
            float3      r_ij, r_kj;
            const int   t1 = pbc_dx(B.pbc, x4[0], x4[1], r_ij), t2 = pbc_dx(B.pbc, x4[2], x4[1], r_kj);
            const float nij2 = dot(r_ij, r_ij), nkj2 = dot(r_kj, r_kj), nij_1 = rsqrtf(nij2), nkj_1 = rsqrtf(nkj2);
            const float c   = fminf(1.f, fmaxf(-1.f, dot(r_ij, r_kj) * nij_1 * nkj_1));
            const float dth = acosf(c) - p[0] * deg2rad;
            if (VE) e0 = 0.5f * p[1] * dth * dth;
            const float c2 = c * c;
            if (c2 < 1.f)
            {
                const float  st = -p[1] * dth * rsqrtf(1.f - c2), sth = st * c;
                const float  cik = st * nij_1 * nkj_1, cii = sth * nij_1 * nij_1, ckk = sth * nkj_1 * nkj_1;
                const float3 f_i = cii * r_ij - cik * r_kj, f_k = ckk * r_kj - cik * r_ij;
                add_force<VE>(f, s_fs, s[0], t1, f_i);
                add_force<VE>(f, s_fs, s[1], B200NB_CENTRAL, -(f_i + f_k));
                add_force<VE>(f, s_fs, s[2], t2, f_k);
            }
            if (kind == B200NB_BONDED_UREY_BRADLEY)
            {
                float3      r_ik;
                const int   ki = pbc_dx(B.pbc, x4[0], x4[2], r_ik);
                const float r2 = dot(r_ik, r_ik), dr = sqrtf(r2) - p[2];
                if (VE) e0 += 0.5f * p[3] * dr * dr;
                if (r2 != 0.f)
                {
                    const float3 fik = (-p[3] * dr * rsqrtf(r2)) * r_ik;
                    add_force<VE>(f, s_fs, s[0], ki, fik);
                    add_force<VE>(f, s_fs, s[2], B200NB_CENTRAL, -fik);
                }
            }
        }
        else
        {
            float3    r_ij, r_kj, r_kl;
            const int t1 = pbc_dx(B.pbc, x4[0], x4[1], r_ij), t2 = pbc_dx(B.pbc, x4[2], x4[1], r_kj);
            (void)pbc_dx(B.pbc, x4[2], x4[3], r_kl);
            const float3 m = cross(r_ij, r_kj), n = cross(r_kj, r_kl), mxn = cross(m, n);
            float        phi = atan2f(sqrtf(dot(mxn, mxn)), dot(m, n));
            if (dot(r_ij, n) < 0.f) phi = -phi;
            float ddphi;
            if (kind == B200NB_BONDED_PDIHS || kind == B200NB_BONDED_PIDIHS)
            {
                const float mult = (float)(int)p[2], mdphi = mult * phi - p[0] * deg2rad;
                float       sn, cs;
                sincosf(mdphi, &sn, &cs);
                if (VE) e0 = p[1] * (1.f + cs);
                ddphi = -p[1] * mult * sn;
            }
            else if (kind == B200NB_BONDED_RBDIHS)
            {
                phi += phi < 0.f ? pi : -pi; /* polymer convention */
                float sp, cp;
                sincosf(phi, &sp, &cp);
                float v = p[0], dd = 0.f, cf = 1.f;
#pragma unroll
                for (int k = 1; k < 6; k++)
                {
                    dd += (float)k * p[k] * cf;
                    cf *= cp;
                    v += cf * p[k];
                }
                if (VE) e0 = v;
                ddphi = -dd * sp;
            }
            else
            {
                float dp = phi - p[0] * deg2rad;
                if (dp >= pi) dp -= 2.f * pi;
                else if (dp < -pi) dp += 2.f * pi;
                if (VE) e0 = 0.5f * p[1] * dp * dp;
                ddphi = p[1] * dp;
            }
            torsion_forces<VE>(B.pbc, x4, s, ddphi, r_ij, r_kj, r_kl, m, n, t1, t2, f, s_fs);
        }
    }
    if (VE)
    {
        /* a warp holds one type (ranges are padded to warps) */
        for (int o = 16; o > 0; o >>= 1)
        {
            e0 += __shfl_xor_sync(0xffffffffu, e0, o);
            e1 += __shfl_xor_sync(0xffffffffu, e1, o);
        }
        if ((threadIdx.x & 31) == 0 && kind >= 0)
        {
            atomicAdd(energy + kind, (double)e0);
            if (kind == B200NB_BONDED_LJ14) atomicAdd(energy + B200NB_BONDED_KINDS, (double)e1);
        }
        __syncthreads();
        float* fs = fshift + (blockIdx.x & (NB_OUT_COPIES - 1)) * NB_FSHIFT_PITCH;
        for (int k = threadIdx.x; k < B200NB_SHIFTS * 3; k += blockDim.x)
            if (s_fs[k] != 0.f) atomicAdd(fs + k, s_fs[k]);
    }
}

} // namespace

extern "C" int b200nb_bonded_set_list(b200nb_t* h, int kind, int nbonds, const int* iatoms_host, int nparams, const float* params6_host)
{
    if (!h) return B200NB_ERR_ARG;
    if (kind < 0 || kind >= B200NB_BONDED_KINDS || nbonds < 0 || nparams < 0 || (nbonds && (!iatoms_host || !params6_host || nparams < 1)))
        return nb_fail(h, B200NB_ERR_ARG, "bonded_set_list: bad argument");
    /* atom indices, or -- on a context whose atoms came in grid order (b200nb_set_grid_atoms: the reference-built grid of the
     * Nbnxm::gpu_* shim, no atom-order view) -- grid slots: the caller converts its lists with the grid's atom order at every
     * search step, as gpubonded_impl.cu:178-310 does */
    const int limit = h->natoms > 0 ? h->natoms : h->npad;
    if (limit < 1) return nb_fail(h, B200NB_ERR_STATE, "bonded_set_list: set_atoms (or set_grid_atoms) first");
    const int nral = (kind == B200NB_BONDED_BONDS || kind == B200NB_BONDED_LJ14) ? 2 : (kind <= B200NB_BONDED_UREY_BRADLEY ? 3 : 4);
    for (int i = 0; i < nbonds; i++)
    {
        const int* ia = iatoms_host + (size_t)(nral + 1) * i;
        if (ia[0] < 0 || ia[0] >= nparams) return nb_fail(h, B200NB_ERR_ARG, "bonded_set_list: parameter index out of range");
        for (int k = 1; k <= nral; k++)
            if (ia[k] < 0 || ia[k] >= limit) return nb_fail(h, B200NB_ERR_ARG, "bonded_set_list: atom index out of range");
    }
    cudaSetDevice(h->device);
    BondedState& S = h->bonded;
    S.count[kind] = 0;
    if (nbonds > 0)
    {
        /* the arrays are kept and only grown (with head-room): callers that convert their lists at every search step -- the
         * Nbnxm::gpu_* shim -- update all eight types each time */
        const size_t ni = (size_t)(nral + 1) * nbonds, np = 6 * (size_t)nparams;
        if (ni > S.cap_iatoms[kind])
        {
            cudaFree(S.d_iatoms[kind]);
            S.d_iatoms[kind] = nullptr, S.cap_iatoms[kind] = 0;
            NB_CUDA(h, cudaMalloc((void**)&S.d_iatoms[kind], sizeof(int) * (ni + ni / 8 + 64)));
            S.cap_iatoms[kind] = ni + ni / 8 + 64;
        }
        if (np > S.cap_params[kind])
        {
            cudaFree(S.d_params[kind]);
            S.d_params[kind] = nullptr, S.cap_params[kind] = 0;
            NB_CUDA(h, cudaMalloc((void**)&S.d_params[kind], sizeof(float) * np));
            S.cap_params[kind] = np;
        }
        NB_CUDA(h, cudaMemcpyAsync(S.d_iatoms[kind], iatoms_host, sizeof(int) * ni, cudaMemcpyHostToDevice, h->stream));
        NB_CUDA(h, cudaMemcpyAsync(S.d_params[kind], params6_host, sizeof(float) * np, cudaMemcpyHostToDevice, h->stream));
        NB_CUDA(h, cudaStreamSynchronize(h->stream)); /* the host arrays are the caller's */
        S.count[kind] = nbonds;
    }
    S.natoms = h->natoms;
    h->generation++; /* a captured step that includes the bonded kernel carries the list pointers by value */
    if (!S.d_energy)
    {
        NB_CUDA(h, cudaMalloc((void**)&S.d_energy, sizeof(double) * (B200NB_BONDED_KINDS + 1)));
        NB_CUDA(h, cudaMemset(S.d_energy, 0, sizeof(double) * (B200NB_BONDED_KINDS + 1)));
    }
    return 0;
}

extern "C" int b200nb_bonded_launch(b200nb_t* h, int flags, float epsfac_fudge)
{
    if (!h) return B200NB_ERR_ARG;
    BondedState& S = h->bonded;
    if (!S.d_energy) return 0; /* no lists: haveInteractions() == false */
    if (S.natoms != h->natoms) return nb_fail(h, B200NB_ERR_STATE, "bonded_launch: lists were set for another set of atoms");
    if (h->natoms > 0 && !h->grid[0].valid) return nb_fail(h, B200NB_ERR_STATE, "bonded_launch: put_on_grid first");
    if (h->dd.window) return nb_fail(h, B200NB_ERR_ARG, "bonded_launch: not built for decomposed runs");
    cudaSetDevice(h->device);
    BondedDev B{};
    int       nthreads = 0;
    for (int k = 0; k < B200NB_BONDED_KINDS; k++)
    {
        B.first[k]  = nthreads;
        B.count[k]  = S.count[k];
        B.iatoms[k] = S.d_iatoms[k];
        B.params[k] = S.d_params[k];
        nthreads += (S.count[k] + 31) / 32 * 32;
    }
    B.first[B200NB_BONDED_KINDS] = nthreads;
    if (nthreads == 0) return 0;
    /* setPbcAiuc, pbcutil/pbc_aiuc.h:99-131: the first npbcdim dimensions are periodic */
    float bx, by, bz, yx, zx, zy;
    if (S.have_pbc)
    {
        bx = S.npbcdim > 0 ? S.box9[0] : 0.f, by = S.npbcdim > 1 ? S.box9[4] : 0.f, bz = S.npbcdim > 2 ? S.box9[8] : 0.f;
        yx = S.box9[3], zx = S.box9[6], zy = S.box9[7];
    }
    else
    {
        bx = h->pbc[0] ? h->box[0] : 0.f, by = h->pbc[1] ? h->box[1] : 0.f, bz = h->pbc[2] ? h->box[2] : 0.f;
        yx = h->box_off[0], zx = h->box_off[1], zy = h->box_off[2];
    }
    B.pbc.xx = bx, B.pbc.inv_xx = bx > 0.f ? 1.0f / bx : 0.f;
    B.pbc.yy = by, B.pbc.inv_yy = by > 0.f ? 1.0f / by : 0.f, B.pbc.yx = by > 0.f ? yx : 0.f;
    B.pbc.zz = bz, B.pbc.inv_zz = bz > 0.f ? 1.0f / bz : 0.f, B.pbc.zx = bz > 0.f ? zx : 0.f, B.pbc.zy = bz > 0.f ? zy : 0.f;
    B.scale14 = epsfac_fudge;
    const unsigned nblk = (unsigned)((nthreads + 127) / 128);
    const float4*  xq   = reinterpret_cast<const float4*>(h->d_xq);
    const int*     slot_map = h->natoms > 0 ? h->d_slot_of_atom : nullptr; /* grid-order atoms: the lists hold slots already */
    if (flags & (B200NB_FLAG_ENERGY | B200NB_FLAG_VIRIAL))
        k_bonded<true><<<nblk, 128, 0, h->stream>>>(B, xq, slot_map, h->d_f, h->d_fshift, S.d_energy);
    else
        k_bonded<false><<<nblk, 128, 0, h->stream>>>(B, xq, slot_map, h->d_f, h->d_fshift, S.d_energy);
    h->nlaunches++;
    NB_CUDA(h, cudaGetLastError());
    return 0;
}

extern "C" int b200nb_bonded_get_energies(b200nb_t* h, double energies_host[B200NB_BONDED_KINDS + 1])
{
    if (!h || !energies_host) return nb_fail(h, B200NB_ERR_ARG, "bonded_get_energies: bad argument");
    BondedState& S = h->bonded;
    for (int k = 0; k <= B200NB_BONDED_KINDS; k++) energies_host[k] = 0.0;
    if (!S.d_energy) return 0;
    cudaSetDevice(h->device);
    NB_CUDA(h, cudaMemcpyAsync(energies_host, S.d_energy, sizeof(double) * (B200NB_BONDED_KINDS + 1), cudaMemcpyDeviceToHost, h->stream));
    NB_CUDA(h, cudaMemsetAsync(S.d_energy, 0, sizeof(double) * (B200NB_BONDED_KINDS + 1), h->stream)); /* read and reset */
    NB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

/* GpuBonded::setPbc (gpubonded_impl.cu:312-316 -> setPbcAiuc): the cell of the image search, for callers whose context was not
 * given the box (the Nbnxm::gpu_* interface passes shift vectors only); box9 = NULL returns to the context's cell. */
extern "C" int b200nb_bonded_set_pbc(b200nb_t* h, const float box9[9], int npbcdim)
{
    if (!h || npbcdim < 0 || npbcdim > 3) return nb_fail(h, B200NB_ERR_ARG, "bonded_set_pbc: bad argument");
    BondedState& S = h->bonded;
    const bool   changed = (box9 != nullptr) != S.have_pbc || (box9 && (memcmp(S.box9, box9, sizeof(S.box9)) != 0 || S.npbcdim != npbcdim));
    S.have_pbc = box9 != nullptr;
    if (box9) memcpy(S.box9, box9, sizeof(S.box9)), S.npbcdim = npbcdim;
    if (changed) h->generation++; /* the cell is a kernel parameter of a captured step */
    return 0;
}

/* The bonded kernel as part of b200nb_step / b200nb_compute: launched between the force kernel and the un-sort, inside the
 * captured step graph (the reference launches its bonded kernel on the nonbonded stream between the nonbonded kernel and the
 * force copy-back in the same way, mdlib/sim_util.cpp:1467-1478). */
extern "C" int b200nb_bonded_in_step(b200nb_t* h, int enable, float epsfac_fudge)
{
    if (!h) return B200NB_ERR_ARG;
    h->bonded.in_step = enable != 0;
    h->bonded.scale14 = epsfac_fudge;
    h->generation++;
    return 0;
}

int nb_bonded_enqueue_in_step(b200nb_context* h, int flags)
{
    if (!h->bonded.in_step) return 0;
    return b200nb_bonded_launch(h, flags, h->bonded.scale14);
}

void nb_bonded_free(b200nb_context* h)
{
    BondedState& S = h->bonded;
    for (int k = 0; k < B200NB_BONDED_KINDS; k++) cudaFree(S.d_iatoms[k]), cudaFree(S.d_params[k]);
    cudaFree(S.d_energy);
    S = BondedState{};
}
