// Shared pieces of the push path (K1) used by push.cu and by the fused push+deposit kernels of move.cu:
// parameters, the node-interleaved E/B pack, the packed gather, and the per-particle move arithmetic
// (BorisPusher::move without the stores, src/core/numerics/pusher/boris.hpp:93-138).
#pragma once
#include "particle_math.cuh"
#include "pipeline.cuh"

#include <algorithm>

namespace phb
{
template<int DIM>
struct PushParams
{
    DevLayout L;
    FieldView E[3], B[3];
    // E and B re-packed per call into ONE array of nodes {Ex,Ey,Ez,Bx,By,Bz} on the common (primal-sized)
    // index space: one 32-bit node index per component, shared row strides, immediate offsets along the
    // fastest direction, and the six components of a node share cache lines
    const double* em;
    int ps0, ps1;       // node strides of the packed array along x and y (ps2 = 1)
    long long rs0, rs1; // the same strides in bytes (48 bytes per node)
    PartView in, out;
    size_t n;
    double h[3];   // 0.5*dt/dx  (Pusher::setMeshAndTimeStep, boris.hpp:143-148)
    double dto2m;  // 0.5*dt/mass (boris.hpp:108)
    DevBox first;  // first selector box
    DevError* err;
    bool copy_weight_charge;
};

// re-pack the six field components into the node-interleaved array (one thread per packed node; reads
// coalesced along the fastest index of every component, writes 48 contiguous bytes per thread)
template<int DIM>
struct PackParams
{
    FieldView f[6];
    double* em;
    int pn[3]; // packed extents
};
template<int DIM>
__global__ void __launch_bounds__(256) pack_em_kernel(const __grid_constant__ PackParams<DIM> A)
{
    size_t const t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= size_t(A.pn[0]) * A.pn[1] * A.pn[2])
        return;
    int const k = int(t % A.pn[2]);
    int const j = int((t / A.pn[2]) % A.pn[1]);
    int const i = int(t / (size_t(A.pn[2]) * A.pn[1]));
    double v[6];
#pragma unroll
    for (int c = 0; c < 6; ++c)
    {
        const FieldView& f = A.f[c];
        v[c] = (i < f.n[0] && j < f.n[1] && k < f.n[2]) ? f.p[f.at(i, j, k)] : 0.;
    }
    double2* out = reinterpret_cast<double2*>(A.em + t * 6);
    out[0]       = make_double2(v[0], v[1]);
    out[1]       = make_double2(v[2], v[3]);
    out[2]       = make_double2(v[4], v[5]);
}

// MeshToParticle on the packed array: same nested z -> y -> x accumulation and operation order as gather().
// Addresses are built as byte pointers: one IMAD.WIDE per component (node index * 48 + base) and one 64-bit
// add of a warp-uniform byte stride per (ix,iy) row; the offsets along the fastest direction are immediates.
// Each chain starts from its first product instead of `0. + product` (identical value: x + 0 == x; only the
// sign of an exact zero could differ, which no later operation observes).
struct EmPtr
{
    const char* p;
    __device__ __forceinline__ double operator[](int node) const
    {
        return __ldg(reinterpret_cast<const double*>(p + node * 48));
    }
};
template<int DIM, int ORDER, int QTY, int COMP, bool EXACT>
__device__ __forceinline__ double gather_packed(const IndexWeights<DIM, ORDER>& iw, const double* em, long long rs0,
                                                long long rs1, int ps0, int ps1)
{
    constexpr int cx = centering(QTY, 0), cy = centering(QTY, 1), cz = centering(QTY, 2);
    auto chain = [](double acc, double f, double w, bool first) { return first ? f * w : mad<EXACT>(f, w, acc); };
    double F = 0.;
    if constexpr (DIM == 1)
    {
        EmPtr const row{reinterpret_cast<const char*>(em) + (long long)(iw.start[cx][0]) * 48 + COMP * 8};
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
            F = chain(F, row[ix], iw.w[cx][0][ix], ix == 0);
    }
    else if constexpr (DIM == 2)
    {
        const char* base
            = reinterpret_cast<const char*>(em) + (long long)(iw.start[cx][0] * ps0 + iw.start[cy][1]) * 48 + COMP * 8;
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
        {
            EmPtr const row{base + ix * rs0};
            double Y = 0.;
#pragma unroll
            for (int iy = 0; iy <= ORDER; ++iy)
                Y = chain(Y, row[iy], iw.w[cy][1][iy], iy == 0);
            F = chain(F, Y, iw.w[cx][0][ix], ix == 0);
        }
    }
    else
    {
        const char* base = reinterpret_cast<const char*>(em)
                           + (long long)(iw.start[cx][0] * ps0 + iw.start[cy][1] * ps1 + iw.start[cz][2]) * 48 + COMP * 8;
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
        {
            double Y = 0.;
#pragma unroll
            for (int iy = 0; iy <= ORDER; ++iy)
            {
                EmPtr const row{base + (ix * rs0 + iy * rs1)};
                double Z = 0.;
#pragma unroll
                for (int iz = 0; iz <= ORDER; ++iz)
                    Z = chain(Z, row[iz], iw.w[cz][2][iz], iz == 0);
                Y = chain(Y, Z, iw.w[cy][1][iy], iy == 0);
            }
            F = chain(F, Y, iw.w[cx][0][ix], ix == 0);
        }
    }
    return F;
}

// pre-push, (first selector), gather, Boris, post-push on the particle held in registers; no memory traffic
// except the E,B nodes.  `ok` goes false on a move of more than two cells (boris.hpp:164-165).
template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST>
__device__ __forceinline__ void move_particle(const PushParams<DIM>& P, int (&icell)[DIM], double (&delta)[DIM],
                                              double (&v)[3], double charge, bool& ok, double& bad_delta,
                                              double& bad_vel)
{
    advance_position<DIM>(P.h, icell, delta, v, ok, bad_delta, bad_vel);
    // The reference's sweep ends with an exception at the first particle that moves more than two cells; here the
    // error is reported at the next poll and the sweep goes on, so the offender must not index anything: no gather
    // (its cell may lie anywhere).  It is stored where the half step left it; the deposit and re-binning kernels only
    // touch particles inside their selection boxes (by default the cells whose stencil fits the arrays), so they skip it.
    if (!ok)
        return;

    bool selected = true;
    if constexpr (HAS_FIRST)
        selected = in_box<DIM>(icell, P.first);

    if (selected)
    {
        IndexWeights<DIM, ORDER> iw;
        both_centerings<DIM, ORDER>(P.L, icell, delta, iw);
        double E[3], B[3];
        E[0] = gather_packed<DIM, ORDER, PHB_EX, 0, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        E[1] = gather_packed<DIM, ORDER, PHB_EY, 1, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        E[2] = gather_packed<DIM, ORDER, PHB_EZ, 2, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        B[0] = gather_packed<DIM, ORDER, PHB_BX, 3, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        B[1] = gather_packed<DIM, ORDER, PHB_BY, 4, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        B[2] = gather_packed<DIM, ORDER, PHB_BZ, 5, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        boris<EXACT>(v, charge, P.dto2m, E, B);
        advance_position<DIM>(P.h, icell, delta, v, ok, bad_delta, bad_vel);
    }
}

__device__ __forceinline__ void report_move_error(DevError* err, bool ok, double bad_delta, double bad_vel, size_t i)
{
    if (!ok && atomicCAS(&err->code, 0, int(PHB_ERR_MOVE_TWO_CELL)) == 0)
    {
        err->delta = bad_delta;
        err->vel   = bad_vel;
        err->index = i;
    }
}

// host side: fill the parameter block and re-pack E,B (one small launch) for a push on layout L
template<int DIM>
int prepare_push(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, double mass,
                 double dt, const phb_box* first, PushParams<DIM>& P)
{
    P.L = make_dev_layout(*L);
    for (int c = 0; c < 3; ++c)
    {
        P.E[c] = make_view(P.L, E->comp[c], PHB_EX + c);
        P.B[c] = make_view(P.L, B->comp[c], PHB_BX + c);
    }
    for (int d = 0; d < 3; ++d)
        P.h[d] = d < DIM ? 0.5 * dt / L->dx[d] : 0.;
    P.dto2m = 0.5 * dt / mass;
    if (first)
        P.first = make_box(*first, DIM);
    P.err = ctx->d_err;
    PackParams<DIM> K;
    size_t nodes = 1;
    for (int d = 0; d < 3; ++d)
    {
        K.pn[d] = d < DIM ? P.L.ncells[d] + 1 + 2 * P.L.g : 1;
        nodes *= size_t(K.pn[d]);
    }
    for (int c = 0; c < 3; ++c)
    {
        K.f[c]     = P.E[c];
        K.f[3 + c] = P.B[c];
    }
    if (nodes * 6 * sizeof(double) > ctx->em_bytes)
    {
        if (ctx->em_pack)
        {
            PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            PHB_CUDA(ctx, cudaFree(ctx->em_pack));
            ctx->em_pack = nullptr;
        }
        PHB_CUDA(ctx, cudaMalloc(&ctx->em_pack, nodes * 6 * sizeof(double)));
        ctx->em_bytes = nodes * 6 * sizeof(double);
    }
    K.em = ctx->em_pack;
    pack_em_kernel<DIM><<<unsigned((nodes + 255) / 256), 256, 0, ctx->stream>>>(K);
    PHB_LAUNCH_CHECK(ctx);
    P.em  = ctx->em_pack;
    P.ps0 = K.pn[1] * K.pn[2];
    P.ps1 = K.pn[2];
    if (DIM == 2)
        P.ps0 = K.pn[1]; // 2-D: index = i * pn[1] + j
    P.rs0 = (long long)P.ps0 * 48;
    P.rs1 = (long long)P.ps1 * 48;
    return PHB_OK;
}
} // namespace phb
