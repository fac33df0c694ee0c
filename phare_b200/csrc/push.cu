// K1 — fused interpolate + Boris push.  Replaces BorisPusher::move (src/core/numerics/pusher/
// boris.hpp:93-138) with its prePushStep_/postPushStep_ (:180-234), Interpolator::operator()
// (particle, em, layout) (interpolator/interpolator.hpp:420-456) and accelerate_ (:240-300).
//
// One thread per particle, SoA columns read and written fully coalesced (a warp touches
// 32 x 8 B = 2 full 128-B lines per double column).  HBM traffic per particle (algorithmic):
//   read  iCell 4d + delta 8d + v 24 + charge 8 ; write iCell 4d + delta 8d + v 24
//   = 80 / 104 / 128 B for d = 1 / 2 / 3 (weight is only copied when out != in).
// The E,B nodes a particle needs ((o+1)^d per component) are read through the read-only L1 path:
// with the cell-sorted store a warp sits in one or two cells, so every gather load is a one- or
// two-address broadcast that hits L1/L2; field arrays are (o+1)^d-fold reused and never the
// HBM bound.
#include "particle_math.cuh"

namespace phb
{
template<int DIM>
struct PushParams
{
    DevLayout L;
    FieldView E[3], B[3];
    PartView in, out;
    size_t n;
    double h[3];   // 0.5*dt/dx  (Pusher::setMeshAndTimeStep, boris.hpp:143-148)
    double dto2m;  // 0.5*dt/mass (boris.hpp:108)
    DevBox first;  // first selector box
    DevError* err;
    bool copy_weight_charge;
};

template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST>
__global__ void __launch_bounds__(256) push_kernel(const __grid_constant__ PushParams<DIM> P)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= P.n)
        return;

    int icell[DIM];
    double delta[DIM], v[3];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        icell[d] = __ldcs(P.in.icell[d] + i);
        delta[d] = __ldcs(P.in.delta[d] + i);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        v[c] = __ldcs(P.in.v[c] + i);
    double const charge = __ldcs(P.in.charge + i);
    if (P.copy_weight_charge)
    {
        __stcs(P.out.charge + i, charge);
        __stcs(P.out.weight + i, __ldcs(P.in.weight + i));
    }

    double bad_delta = 0, bad_vel = 0;
    bool ok = advance_position<DIM>(P.h, icell, delta, v, bad_delta, bad_vel);

    bool selected = true;
    if constexpr (HAS_FIRST)
        selected = in_box<DIM>(icell, P.first);

    if (selected)
    {
        IndexWeights<DIM, ORDER> iw;
        both_centerings<DIM, ORDER>(P.L, icell, delta, iw);
        double E[3], B[3];
        auto ld = [](const FieldView& f) {
            return [&f](int a, int b, int c) { return __ldg(f.p + f.at(a, b, c)); };
        };
        E[0] = gather<DIM, ORDER, PHB_EX, EXACT>(iw, ld(P.E[0]));
        E[1] = gather<DIM, ORDER, PHB_EY, EXACT>(iw, ld(P.E[1]));
        E[2] = gather<DIM, ORDER, PHB_EZ, EXACT>(iw, ld(P.E[2]));
        B[0] = gather<DIM, ORDER, PHB_BX, EXACT>(iw, ld(P.B[0]));
        B[1] = gather<DIM, ORDER, PHB_BY, EXACT>(iw, ld(P.B[1]));
        B[2] = gather<DIM, ORDER, PHB_BZ, EXACT>(iw, ld(P.B[2]));
        boris<EXACT>(v, charge, P.dto2m, E, B);
        ok = advance_position<DIM>(P.h, icell, delta, v, bad_delta, bad_vel) && ok;
    }

#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        __stcs(P.out.icell[d] + i, icell[d]);
        __stcs(P.out.delta[d] + i, delta[d]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        __stcs(P.out.v[c] + i, v[c]);

    if (!ok && atomicCAS(&P.err->code, 0, int(PHB_ERR_MOVE_TWO_CELL)) == 0)
    {
        P.err->delta = bad_delta;
        P.err->vel   = bad_vel;
        P.err->index = i;
    }
}

template<int DIM, int ORDER>
int launch_push(phb_ctx* ctx, const PushParams<DIM>& P, bool has_first)
{
    if (P.n == 0)
        return PHB_OK;
    constexpr int BS = 256;
    unsigned const grid = unsigned((P.n + BS - 1) / BS);
    if (ctx->exact)
    {
        if (has_first)
            push_kernel<DIM, ORDER, true, true><<<grid, BS, 0, ctx->stream>>>(P);
        else
            push_kernel<DIM, ORDER, true, false><<<grid, BS, 0, ctx->stream>>>(P);
    }
    else
    {
        if (has_first)
            push_kernel<DIM, ORDER, false, true><<<grid, BS, 0, ctx->stream>>>(P);
        else
            push_kernel<DIM, ORDER, false, false><<<grid, BS, 0, ctx->stream>>>(P);
    }
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

template<int DIM>
int push_dim(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
             const phb_particles* in, phb_particles* out, double mass, double dt, const phb_box* first)
{
    PushParams<DIM> P;
    P.L = make_dev_layout(*L);
    for (int c = 0; c < 3; ++c)
    {
        P.E[c] = make_view(P.L, E->comp[c], PHB_EX + c);
        P.B[c] = make_view(P.L, B->comp[c], PHB_BX + c);
    }
    P.in  = make_part(*in);
    P.out = make_part(*out);
    P.n   = in->n;
    for (int d = 0; d < 3; ++d)
        P.h[d] = d < DIM ? 0.5 * dt / L->dx[d] : 0.;
    P.dto2m = 0.5 * dt / mass;
    if (first)
        P.first = make_box(*first, DIM);
    P.err                = ctx->d_err;
    P.copy_weight_charge = in->weight != out->weight;
    switch (L->interp)
    {
        case 1: return launch_push<DIM, 1>(ctx, P, first != nullptr);
        case 2: return launch_push<DIM, 2>(ctx, P, first != nullptr);
        default: return launch_push<DIM, 3>(ctx, P, first != nullptr);
    }
}
} // namespace phb

extern "C" int phb_push(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                        const phb_particles* in, phb_particles* out, double mass, double dt,
                        const phb_box* first)
{
    if (!phb::valid_layout(ctx, L) || !E || !B || !in || !out)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push: invalid argument");
    if (out->capacity < in->n)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_push: out.capacity < in.n");
    int rc;
    switch (L->dim)
    {
        case 1: rc = phb::push_dim<1>(ctx, L, E, B, in, out, mass, dt, first); break;
        case 2: rc = phb::push_dim<2>(ctx, L, E, B, in, out, mass, dt, first); break;
        default: rc = phb::push_dim<3>(ctx, L, E, B, in, out, mass, dt, first); break;
    }
    if (rc == PHB_OK)
        out->n = in->n;
    return rc;
}
