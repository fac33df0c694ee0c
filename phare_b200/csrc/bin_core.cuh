// Shared pieces of the binning path (K2) used by bin.cu and by the fused deposit+scatter pass of sortdep.cu:
// the key space of phb_bin (include/phare_b200.h) and the scan entry points.
#pragma once
#include "common.cuh"

namespace phb
{
template<int DIM>
struct KeySpace
{
    DevBox domain, ghost;
    BoxList keep;
    unsigned ext_d[3], ext_g[3];
    unsigned Nd, Ng;
};

template<int DIM>
__device__ __forceinline__ unsigned bin_key(const KeySpace<DIM>& K, const int* c)
{
    if (in_box<DIM>(c, K.domain))
    {
        unsigned k = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            k = k * K.ext_d[d] + unsigned(c[d] - K.domain.lo[d]);
        return k;
    }
    if (in_box<DIM>(c, K.ghost))
    {
        bool kept = false;
        for (int b = 0; b < K.keep.n; ++b)
            kept = kept || in_box<DIM>(c, K.keep.b[b]);
        if (kept)
        {
            unsigned k = 0;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
                k = k * K.ext_g[d] + unsigned(c[d] - K.ghost.lo[d]);
            return K.Nd + k;
        }
    }
    return K.Nd + K.Ng;
}

template<int DIM>
KeySpace<DIM> make_keyspace(const phb_layout* L, const phb_box* domain, const phb_box* keep, int nkeep)
{
    KeySpace<DIM> K;
    K.domain = make_box(*domain, DIM);
    K.ghost  = K.domain;
    int const pg = particle_ghosts(L->interp);
    K.Nd = K.Ng = 1;
    for (int d = 0; d < 3; ++d)
    {
        if (d < DIM)
        {
            K.ghost.lo[d] -= pg;
            K.ghost.hi[d] += pg;
        }
        K.ext_d[d] = unsigned(K.domain.hi[d] - K.domain.lo[d] + 1);
        K.ext_g[d] = unsigned(K.ghost.hi[d] - K.ghost.lo[d] + 1);
        if (d < DIM)
        {
            K.Nd *= K.ext_d[d];
            K.Ng *= K.ext_g[d];
        }
    }
    K.keep.n = nkeep;
    for (int b = 0; b < nkeep; ++b)
        K.keep.b[b] = make_box(keep[b], DIM);
    return K;
}

template<int DIM>
__global__ void __launch_bounds__(256)
    bin_count_kernel(const __grid_constant__ KeySpace<DIM> K, PartView P, size_t n, uint32_t* __restrict__ count,
                     uint32_t* __restrict__ slot)
{
    size_t const i   = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    bool const live  = i < n;
    unsigned key     = 0xffffffffu;
    if (live)
    {
        int c[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            c[d] = __ldcs(P.icell[d] + i);
        key = bin_key<DIM>(K, c);
    }
    // lanes of a warp that share a key reserve their slots with one atomic
    unsigned const peers  = __match_any_sync(0xffffffffu, key);
    unsigned const lane   = threadIdx.x & 31;
    int const leader      = __ffs(peers) - 1;
    unsigned const before = __popc(peers & ((1u << lane) - 1));
    unsigned base         = 0;
    if (live && int(lane) == leader)
        base = atomicAdd(count + key, unsigned(__popc(peers)));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live)
        slot[i] = base + before;
}

template<int DIM>
__global__ void __launch_bounds__(256)
    bin_scatter_kernel(const __grid_constant__ KeySpace<DIM> K, PartView in, PartView out, size_t first, size_t n,
                       const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ slot)
{
    size_t const i = first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    int c[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        c[d] = __ldcs(in.icell[d] + i);
    unsigned const key = bin_key<DIM>(K, c);
    size_t const j     = size_t(__ldg(cell_start + key)) + __ldcs(slot + i);
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        out.icell[d][j] = c[d];
        out.delta[d][j] = __ldcs(in.delta[d] + i);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        out.v[k][j] = __ldcs(in.v[k] + i);
    out.weight[j] = __ldcs(in.weight + i);
    out.charge[j] = __ldcs(in.charge + i);
}

size_t scan_scratch_words(size_t n);
int exclusive_scan(phb_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch);
} // namespace phb
