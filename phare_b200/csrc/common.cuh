// Shared device-side definitions for libphare_b200 (sm_100a).
// Index space restated from GridLayout<Yee> (src/core/data/grid/gridlayout.hpp:746-763,849-864,
// 1386-1494) and the centering table of gridlayout_hybrid_yee.hpp:54-85 as constexpr tables.
#pragma once
#include "../../include/phare_b200.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <string>

namespace phb
{
enum : int { PRIMAL = 0, DUAL = 1 };

// centering per quantity and direction (x,y,z)
__host__ __device__ constexpr int centering(int qty, int dir)
{
    // Bx(p,d,d) By(d,p,d) Bz(d,d,p); E,J: x(d,p,p) y(p,d,p) z(p,p,d); moments primal
    return (qty <= PHB_BZ) ? ((qty - PHB_BX) == dir ? PRIMAL : DUAL)
           : (qty <= PHB_EZ) ? ((qty - PHB_EX) == dir ? DUAL : PRIMAL)
           : (qty <= PHB_JZ) ? ((qty - PHB_JX) == dir ? DUAL : PRIMAL)
                             : PRIMAL;
}
__host__ __device__ constexpr int field_ghosts(int interp) { return interp == 1 ? 2 : 4; }
__host__ __device__ constexpr int particle_ghosts(int interp) { return interp == 1 ? 1 : 2; }

// device copy of the layout, with derived constants
struct DevLayout
{
    int dim, interp, g, level;
    int amr_lower[3];
    int ncells[3];
    double dx[3], inv_dx[3];
};

inline DevLayout make_dev_layout(const phb_layout& L)
{
    DevLayout D{};
    D.dim    = L.dim;
    D.interp = L.interp;
    D.g      = field_ghosts(L.interp);
    D.level  = L.level;
    for (int d = 0; d < 3; ++d)
    {
        D.amr_lower[d] = d < L.dim ? L.amr_lower[d] : 0;
        D.ncells[d]    = d < L.dim ? int(L.ncells[d]) : 1;
        D.dx[d]        = d < L.dim ? L.dx[d] : 1.;
        D.inv_dx[d]    = 1. / D.dx[d]; // GridLayout ctor: inverseMeshSize_ = 1./meshSize (gridlayout.hpp:137)
    }
    return D;
}

// allocSize(qty) per direction
__host__ __device__ inline int alloc_extent(const DevLayout& L, int qty, int dir)
{
    return dir < L.dim ? L.ncells[dir] + (centering(qty, dir) == PRIMAL ? 1 : 0) + 2 * L.g : 1;
}

// C-ordered view of one field component; extents of the unused trailing directions are 1
struct FieldView
{
    double* p;
    int n[3];
    __host__ __device__ inline size_t at(int i, int j, int k) const
    {
        return (size_t(i) * n[1] + j) * n[2] + k;
    }
};
inline FieldView make_view(const DevLayout& L, const double* p, int qty)
{
    FieldView f;
    f.p = const_cast<double*>(p);
    for (int d = 0; d < 3; ++d)
        f.n[d] = alloc_extent(L, qty, d);
    return f;
}
struct VecView
{
    FieldView c[3];
};
inline VecView make_vec(const DevLayout& L, const phb_vecfield* v, int qty0)
{
    VecView r;
    for (int c = 0; c < 3; ++c)
        r.c[c] = make_view(L, v->comp[c], qty0 + c);
    return r;
}

struct DevBox
{
    int lo[3], hi[3];
};
inline DevBox make_box(const phb_box& b, int dim)
{
    DevBox r;
    for (int d = 0; d < 3; ++d)
    {
        r.lo[d] = d < dim ? b.lower[d] : 0;
        r.hi[d] = d < dim ? b.upper[d] : 0;
    }
    return r;
}
template<int DIM>
__device__ __forceinline__ bool in_box(const int* c, const DevBox& b)
{
    bool ok = true;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        ok = ok && c[d] >= b.lo[d] && c[d] <= b.hi[d];
    return ok;
}

constexpr int MAX_BOXES = 28; // a patch has at most 3^3-1 neighbours (+ itself)
struct BoxList
{
    int n;
    DevBox b[MAX_BOXES];
};

// SoA particle columns as kernel parameter
struct PartView
{
    int* icell[3];
    double* delta[3];
    double* v[3];
    double* weight;
    double* charge;
};
inline PartView make_part(const phb_particles& P)
{
    PartView r;
    for (int d = 0; d < 3; ++d)
    {
        r.icell[d] = P.icell[d];
        r.delta[d] = P.delta[d];
        r.v[d]     = P.v[d];
    }
    r.weight = P.weight;
    r.charge = P.charge;
    return r;
}

// kernel-side error record (device memory, polled by phb_poll_error)
struct DevError
{
    int code;
    int pad;
    double delta, vel;
    unsigned long long index;
};

} // namespace phb

struct phb_ctx
{
    int device = 0, dim = 0, interp = 0;
    bool exact            = true;
    bool no_tma           = false; // PHB_NO_TMA=1: force the plain (non bulk-copy) kernels
    bool no_fused_cells   = false; // PHB_NO_FUSED_CELLS=1: phb_push_deposit uses the per-particle kernel only
    cudaStream_t stream   = nullptr;
    bool own_stream       = false;
    phb::DevError* d_err  = nullptr;
    phb::DevError* h_err  = nullptr; // pinned
    uint64_t launches     = 0;
    int sm_count          = 148;
    std::string last_error;
    // scratch for binning (grown on demand)
    void* scratch         = nullptr;
    size_t scratch_bytes  = 0;
    uint32_t* h_counts    = nullptr; // pinned, 8 entries
    uint32_t* h_bounce    = nullptr; // pinned, SMALL_D2H_BYTES: small results on their way to pageable host memory
    double* em_pack       = nullptr; // node-interleaved copy of E,B used by the push kernels
    size_t em_bytes       = 0;
    size_t plan_n         = size_t(-1); // particles covered by the pending phb_bin_plan (slots live in scratch)
    int plan_kind         = 0;     // 0: none / phb_bin_plan (slots in scratch); 2: tile plan (stay / arrivals / slots in plan_buf)
    void* plan_buf        = nullptr; // tile plan: [stay nk+1 | arrivals nk+1 | slot n | scan scratch]
    size_t plan_bytes     = 0;
    void* strip_counter   = nullptr; // work counter of the strip kernel (strip.cuh)
    bool no_strip         = true;  // PHB_STRIP=1 routes K1 of a cell-ordered store through the strip kernel (strip.cuh)
    bool no_tile          = false; // PHB_NO_TILE=1: the cell-ordered passes use the round-1 kernels (E,B through L1)
    double predict_eps    = 1. / 4096.; // predicted re-binning: a predicted delta within eps of a cell face is left to the resolve kernel
};

namespace phb
{
int set_error(phb_ctx* ctx, int code, const std::string& msg);
int cuda_check(phb_ctx* ctx, cudaError_t e, const char* what);
int ensure_scratch(phb_ctx* ctx, size_t bytes);
// Small results (class counts, totals of a scan, the error record) travel to the host through a store from a one-CTA kernel
// into mapped pinned memory, NOT through a copy engine: an engine serves its queue in order, so a 4-byte count enqueued
// behind another stream's field / moment read-back (solver.HostStaging: hundreds of MB) would hold the host for the whole
// transfer.  `pinned_dst` is cudaMallocHost memory (h_counts, h_err, h_bounce); valid after the stream is synchronised.
constexpr size_t SMALL_D2H_BYTES = 4096;
int words_to_host(phb_ctx* ctx, void* pinned_dst, const void* d_src, size_t bytes);
#define PHB_CUDA(ctx, call)                                                                              \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return phb::cuda_check(ctx, e__, #call);                                                     \
    } while (0)
#define PHB_LAUNCH_CHECK(ctx)                                                                            \
    do                                                                                                   \
    {                                                                                                    \
        ++(ctx)->launches;                                                                               \
        cudaError_t e__ = cudaGetLastError();                                                            \
        if (e__ != cudaSuccess)                                                                          \
            return phb::cuda_check(ctx, e__, "kernel launch");                                           \
    } while (0)
inline bool valid_layout(const phb_ctx* ctx, const phb_layout* L)
{
    return ctx && L && L->dim == ctx->dim && L->interp == ctx->interp;
}
} // namespace phb
