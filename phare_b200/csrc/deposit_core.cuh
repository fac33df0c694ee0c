// Shared pieces of the deposit path (K3) used by deposit.cu and by the fused push+deposit kernels of move.cu:
// parameters, the box-list selector, the per-particle atomic scatter (ParticleToMesh<dim>,
// src/core/numerics/interpolator/interpolator.hpp:278-363) and the register reduce-scatter of a cell group.
#pragma once
#include "particle_math.cuh"
#include "pipeline.cuh"

namespace phb
{
struct MomentViews
{
    double* f[5]; // rho_n, rho_q, Fx, Fy, Fz : all primal, same shape
    int n[3];
    __device__ __forceinline__ size_t at(int i, int j, int k) const
    {
        return (size_t(i) * n[1] + j) * n[2] + k;
    }
};

template<int DIM>
struct DepositParams
{
    DevLayout L;
    PartView P;
    MomentViews M;
    size_t first, last;
    double coef;
    BoxList sel;               // n == 0: everything selected
    DevBox keybox;             // cell-ordered kernel: key -> cell
    const uint32_t* cell_start;
    unsigned nkeys;
    // cell-ordered kernels: indices of the particles that left their cell, handled afterwards by deposit_list_kernel.
    // MOVER_LISTS sub-lists (chosen by CTA), each with its own counter on its own 128-byte line: one shared counter
    // serialises ~0.4 % of all particles on a single L2 atomic unit (0.4 ms per pass at config 5)
    uint32_t* mover_list;  // MOVER_LISTS x mover_cap entries
    unsigned* mover_count; // MOVER_LISTS counters, 32 words apart
    unsigned mover_cap;
};

template<int DIM>
__device__ __forceinline__ bool selected(const BoxList& sel, const int* c)
{
    if (sel.n == 0)
        return true;
    for (int b = 0; b < sel.n; ++b)
        if (in_box<DIM>(c, sel.b[b]))
            return true;
    return false;
}

// per-particle atomic scatter (ParticleToMesh<DIM>, interpolator.hpp:278-363)
template<int DIM, int ORDER>
__device__ __forceinline__ void scatter_atomic(const DevLayout& L, const MomentViews& M, const int* icell,
                                               const double* delta, const double (&dep)[5])
{
    int start[DIM];
    double w[DIM][ORDER + 1];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        start[d] = index_and_weights<ORDER, PRIMAL>(icell[d] - (L.amr_lower[d] - L.g), delta[d], w[d]);
#pragma unroll
    for (int f = 0; f < 5; ++f)
    {
        if constexpr (DIM == 1)
        {
#pragma unroll
            for (int ix = 0; ix <= ORDER; ++ix)
                atomicAdd(M.f[f] + M.at(start[0] + ix, 0, 0), dep[f] * w[0][ix]);
        }
        else if constexpr (DIM == 2)
        {
#pragma unroll
            for (int ix = 0; ix <= ORDER; ++ix)
#pragma unroll
                for (int iy = 0; iy <= ORDER; ++iy)
                    atomicAdd(M.f[f] + M.at(start[0] + ix, start[1] + iy, 0), dep[f] * w[0][ix] * w[1][iy]);
        }
        else
        {
#pragma unroll
            for (int ix = 0; ix <= ORDER; ++ix)
#pragma unroll
                for (int iy = 0; iy <= ORDER; ++iy)
#pragma unroll
                    for (int iz = 0; iz <= ORDER; ++iz)
                        atomicAdd(M.f[f] + M.at(start[0] + ix, start[1] + iy, start[2] + iz),
                                  dep[f] * w[0][ix] * w[1][iy] * w[2][iz]);
        }
    }
}

// union of the primal supports of all particles of one cell, per direction:
// order 1: {l, l+1}; order 2: {l-1 .. l+2} (start = l-1 or l); order 3: {l-1 .. l+2}
template<int ORDER> constexpr int cell_support() { return ORDER == 1 ? 2 : 4; }
template<int ORDER> constexpr int cell_base_shift() { return ORDER == 1 ? 0 : 1; }
constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

template<int NV, int GS, int NCHUNK, int MASK>
struct GroupReduce
{
    // a[0 .. NCHUNK*5) valid on entry; after all steps the lane owns chunks [base, base+nleft)
    __device__ static __forceinline__ void run(double (&a)[NV], int lane, int& base, int& nleft)
    {
        if constexpr (MASK < GS)
        {
            if constexpr (NCHUNK > 1)
            {
                constexpr int half = NCHUNK / 2;
                bool const upper   = (lane & MASK) != 0;
#pragma unroll
                for (int i = 0; i < half * 5; ++i)
                {
                    double const send = upper ? a[i] : a[i + half * 5];
                    double const keep = upper ? a[i + half * 5] : a[i];
                    a[i]              = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
                }
                base += upper ? half : 0;
                GroupReduce<NV, GS, half, MASK * 2>::run(a, lane, base, nleft);
            }
            else
            {
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    a[i] += __shfl_xor_sync(0xffffffffu, a[i], MASK);
                GroupReduce<NV, GS, 1, MASK * 2>::run(a, lane, base, nleft);
            }
        }
        else
            nleft = NCHUNK;
    }
};

template<int DIM>
struct Loaded
{
    int icell[DIM];
    double delta[DIM], v[3], weight, charge;
};
template<int DIM>
__device__ __forceinline__ Loaded<DIM> load_particle(const PartView& P, size_t p)
{
    Loaded<DIM> r;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        r.icell[d] = __ldcs(P.icell[d] + p);
        r.delta[d] = __ldcs(P.delta[d] + p);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        r.v[c] = __ldcs(P.v[c] + p);
    r.weight = __ldcs(P.weight + p);
    r.charge = __ldcs(P.charge + p);
    return r;
}

// Particles that left their cell since the store was ordered are not scattered from inside the cell-ordered
// kernel (one active lane doing 5*(o+1)^d atomics stalls its warp): the kernel only appends their index to
// a list; this kernel then handles the list densely, one thread per listed particle.
constexpr unsigned MOVER_LISTS = 1024;
// words of scratch for a pass over n particles: [LISTS+1 counters, 32 words apart | LISTS sub-lists | overflow list]
inline size_t mover_scratch_words(size_t n) { return size_t(MOVER_LISTS + 1) * 32 + (n + MOVER_LISTS) + n; }
template<int DIM>
inline void set_mover_lists(DepositParams<DIM>& A, uint32_t* base, size_t n)
{
    A.mover_count = base;
    A.mover_list  = base + size_t(MOVER_LISTS + 1) * 32;
    A.mover_cap   = unsigned(n / MOVER_LISTS + 1);
}
inline size_t mover_counter_bytes() { return size_t(MOVER_LISTS + 1) * 32 * sizeof(unsigned); }
// append particle p to this CTA's sub-list, or, when that is full, to the shared overflow list (which holds any number)
template<int DIM>
__device__ __forceinline__ void mover_append(const DepositParams<DIM>& A, size_t p)
{
    unsigned const l = blockIdx.x % MOVER_LISTS;
    unsigned const r = atomicAdd(A.mover_count + l * 32, 1u);
    if (r < A.mover_cap)
        A.mover_list[size_t(l) * A.mover_cap + r] = uint32_t(p);
    else
        A.mover_list[size_t(MOVER_LISTS) * A.mover_cap + atomicAdd(A.mover_count + MOVER_LISTS * 32, 1u)] = uint32_t(p);
}
template<int DIM, int ORDER>
__device__ __forceinline__ void deposit_one(const DepositParams<DIM>& A, size_t i)
{
    int icell[DIM];
    double delta[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        icell[d] = A.P.icell[d][i];
        delta[d] = A.P.delta[d][i];
    }
    double const weight = A.P.weight[i];
    double const dep[5] = {1. * weight * A.coef, A.P.charge[i] * weight * A.coef, A.P.v[0][i] * weight * A.coef,
                           A.P.v[1][i] * weight * A.coef, A.P.v[2][i] * weight * A.coef};
    scatter_atomic<DIM, ORDER>(A.L, A.M, icell, delta, dep);
}
// the listed particles, one thread each: CTA b walks the sub-lists b, b + gridDim.x, ... then its share of the overflow
template<int DIM, int ORDER>
__global__ void __launch_bounds__(256) deposit_list_kernel(const __grid_constant__ DepositParams<DIM> A)
{
    for (unsigned l = blockIdx.x; l < MOVER_LISTS; l += gridDim.x)
    {
        unsigned const total = A.mover_count[l * 32];
        unsigned const n     = total < A.mover_cap ? total : A.mover_cap;
        for (unsigned t = threadIdx.x; t < n; t += blockDim.x)
            deposit_one<DIM, ORDER>(A, A.mover_list[size_t(l) * A.mover_cap + t]);
    }
    unsigned const over      = A.mover_count[MOVER_LISTS * 32];
    const uint32_t* overflow = A.mover_list + size_t(MOVER_LISTS) * A.mover_cap;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < over; t += gridDim.x * blockDim.x)
        deposit_one<DIM, ORDER>(A, overflow[t]);
}

// host side: fill the parameter block of a deposit on layout L
template<int DIM>
void prepare_deposit(const phb_layout* L, const phb_particles* P, size_t first, size_t last, double* rho_n,
                     double* rho_q, const phb_vecfield* flux, double coef, const phb_box* sel, int nsel,
                     const phb_box* domain, const uint32_t* cell_start, DepositParams<DIM>& A)
{
    A.L      = make_dev_layout(*L);
    A.P      = make_part(*P);
    A.M.f[0] = rho_n;
    A.M.f[1] = rho_q;
    for (int c = 0; c < 3; ++c)
        A.M.f[2 + c] = flux->comp[c];
    for (int d = 0; d < 3; ++d)
        A.M.n[d] = alloc_extent(A.L, PHB_RHO, d);
    A.first = first;
    A.last  = last;
    A.coef  = coef;
    A.sel.n = nsel;
    for (int b = 0; b < nsel; ++b)
        A.sel.b[b] = make_box(sel[b], DIM);
    if (nsel == 0)
    {
        // "everything" = every cell whose (order+1)^d primal stencil fits the arrays (nodes -g .. n+g): order 1 touches
        // the nodes c, c+1 -> the patch grown by g cells; orders 2 and 3 touch c-1 .. c+2 -> grown by g-1.  A particle
        // left far outside by a move-two-cells error (reported at the next poll) is skipped instead of scattered out of
        // bounds; everything the reference could deposit is still deposited.
        int const grow = A.L.interp == 1 ? A.L.g : A.L.g - 1;
        A.sel.n        = 1;
        for (int d = 0; d < 3; ++d)
        {
            A.sel.b[0].lo[d] = d < DIM ? A.L.amr_lower[d] - grow : 0;
            A.sel.b[0].hi[d] = d < DIM ? A.L.amr_lower[d] + A.L.ncells[d] - 1 + grow : 0;
        }
    }
    A.cell_start  = cell_start;
    A.nkeys       = 0;
    A.mover_list  = nullptr;
    A.mover_count = nullptr;
    A.mover_cap   = 0;
    if (cell_start != nullptr && domain != nullptr)
    {
        A.keybox   = make_box(*domain, DIM);
        size_t vol = 1;
        for (int d = 0; d < DIM; ++d)
            vol *= size_t(domain->upper[d] - domain->lower[d] + 1);
        A.nkeys = unsigned(vol);
    }
}
} // namespace phb
