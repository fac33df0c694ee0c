// K4-K7 — Faraday, Ampere, Ohm, Electrons::update, ion totals, average.
// Bandwidth-bound stencils on the Yee grid, one thread per node, consecutive threads along the
// fastest (last) index so every load and store is coalesced; neighbours come from L1/L2.
// Arithmetic restated in the reference's order (compiled -fmad=false => bit-identical):
//   Faraday  src/core/numerics/faraday/faraday.hpp:28-97
//   Ampere   src/core/numerics/ampere/ampere.hpp:26-95
//   Ohm      src/core/numerics/ohm/ohm.hpp:48-270
//   deriv / laplacian / project   src/core/data/grid/gridlayout.hpp:550-700,784-796
//   Yee projection stencils       src/core/data/grid/impl/yee/gridlayout_hybrid_yee.hpp:350-749
//   Electrons::update             src/core/data/electrons/electrons.hpp:102-128,202-212
//   Ions totals                   src/core/data/ions/ions.hpp:75-145
//   average                       src/core/utilities/algorithm.hpp:68-77
#include "yee.cuh"

namespace phb
{
// ---------------------------------------------------------------- Faraday
template<int DIM>
struct FaradayParams
{
    DevLayout L;
    VecView B, E, Bnew;
    IterBox box[3];
    double dt;
};
template<int DIM>
__global__ void __launch_bounds__(256) faraday_kernel(const __grid_constant__ FaradayParams<DIM> A)
{
    int const c = blockIdx.y;
    int i, j, k;
    if (!unravel(A.box[c], size_t(blockIdx.x) * blockDim.x + threadIdx.x, i, j, k))
        return;
    const DevLayout& L = A.L;
    double const dt    = A.dt;
    const FieldView &Ex = A.E.c[0], &Ey = A.E.c[1], &Ez = A.E.c[2];
    const FieldView& b = A.B.c[c];
    size_t const p     = b.at(i, j, k);
    double r;
    if (c == 0)
    {
        if constexpr (DIM == 1)
            r = b.p[p];
        else if constexpr (DIM == 2)
            r = b.p[p] - dt * deriv<1>(L, Ez, PHB_EZ, i, j, k);
        else
            r = b.p[p] - dt * deriv<1>(L, Ez, PHB_EZ, i, j, k) + dt * deriv<2>(L, Ey, PHB_EY, i, j, k);
    }
    else if (c == 1)
    {
        if constexpr (DIM < 3)
            r = b.p[p] + dt * deriv<0>(L, Ez, PHB_EZ, i, j, k);
        else
            r = b.p[p] - dt * deriv<2>(L, Ex, PHB_EX, i, j, k) + dt * deriv<0>(L, Ez, PHB_EZ, i, j, k);
    }
    else
    {
        if constexpr (DIM == 1)
            r = b.p[p] - dt * deriv<0>(L, Ey, PHB_EY, i, j, k);
        else
            r = b.p[p] - dt * deriv<0>(L, Ey, PHB_EY, i, j, k) + dt * deriv<1>(L, Ex, PHB_EX, i, j, k);
    }
    A.Bnew.c[c].p[p] = r;
}

// ---------------------------------------------------------------- Ampere
template<int DIM>
struct AmpereParams
{
    DevLayout L;
    VecView B, J;
    IterBox box[3];
};
template<int DIM>
__global__ void __launch_bounds__(256) ampere_kernel(const __grid_constant__ AmpereParams<DIM> A)
{
    int const c = blockIdx.y;
    int i, j, k;
    if (!unravel(A.box[c], size_t(blockIdx.x) * blockDim.x + threadIdx.x, i, j, k))
        return;
    const DevLayout& L = A.L;
    const FieldView &Bx = A.B.c[0], &By = A.B.c[1], &Bz = A.B.c[2];
    double r;
    if (c == 0)
    {
        if constexpr (DIM == 1)
            r = 0.0;
        else if constexpr (DIM == 2)
            r = deriv<1>(L, Bz, PHB_BZ, i, j, k);
        else
            r = deriv<1>(L, Bz, PHB_BZ, i, j, k) - deriv<2>(L, By, PHB_BY, i, j, k);
    }
    else if (c == 1)
    {
        if constexpr (DIM < 3)
            r = -deriv<0>(L, Bz, PHB_BZ, i, j, k);
        else
            r = deriv<2>(L, Bx, PHB_BX, i, j, k) - deriv<0>(L, Bz, PHB_BZ, i, j, k);
    }
    else
    {
        if constexpr (DIM == 1)
            r = deriv<0>(L, By, PHB_BY, i, j, k);
        else
            r = deriv<0>(L, By, PHB_BY, i, j, k) - deriv<1>(L, Bx, PHB_BX, i, j, k);
    }
    A.J.c[c].p[A.J.c[c].at(i, j, k)] = r;
}

// ---------------------------------------------------------------- Ohm
template<int DIM>
struct OhmParams
{
    DevLayout L;
    FieldView n, Pe;
    VecView Ve, B, J, E;
    IterBox box[3];
    double eta, nu, lvlCoeff;
    int hyper_mode;
};

// one E component: C = 0,1,2.  Stencil kinds from gridlayout_hybrid_yee.hpp:450-749
template<int DIM, int C>
__device__ __forceinline__ double ohm_component(const OhmParams<DIM>& A, int i, int j, int k)
{
    // momentsToE<C>: PrimalToDual along C
    constexpr int MX = C == 0 ? 1 : 0, MY = C == 1 ? 1 : 0, MZ = C == 2 ? 1 : 0;
    constexpr int C1 = (C + 1) % 3, C2 = (C + 2) % 3;
    // B<X>ToE<C> kinds: diagonal (X==C): P2D along C, D2P along the two others;
    // off-diagonal: D2P along the third direction (the one that is neither X nor C)
    auto bproj = [&](auto xtag) {
        constexpr int X = decltype(xtag)::value;
        if constexpr (X == C)
            return project<DIM, (C == 0 ? 1 : 2), (C == 1 ? 1 : 2), (C == 2 ? 1 : 2)>(A.B.c[X], i, j, k);
        else
        {
            constexpr int T = 3 - X - C;
            return project<DIM, (T == 0 ? 2 : 0), (T == 1 ? 2 : 0), (T == 2 ? 2 : 0)>(A.B.c[X], i, j, k);
        }
    };
    // ideal_ (ohm.hpp:92-143):  E_C = -v_{C1} * b_{C2} + v_{C2} * b_{C1}
    double const v1    = project<DIM, MX, MY, MZ>(A.Ve.c[C1], i, j, k);
    double const v2    = project<DIM, MX, MY, MZ>(A.Ve.c[C2], i, j, k);
    double const b1    = bproj(std::integral_constant<int, C1>{});
    double const b2    = bproj(std::integral_constant<int, C2>{});
    double const ideal = -v1 * b2 + v2 * b1;
    // pressure_ (ohm.hpp:145-187)
    double pressure = 0.;
    if constexpr (C < DIM)
    {
        double const nOnE  = project<DIM, MX, MY, MZ>(A.n, i, j, k);
        double const gradP = deriv<C>(A.L, A.Pe, PHB_P, i, j, k);
        pressure           = -gradP / nOnE;
    }
    // resistive_ (ohm.hpp:189-209): J<C>ToE<C> is the identity stencil (0. + 1.0*J)
    const FieldView& Jc    = A.J.c[C];
    double const jOnE      = 0. + 1.0 * Jc.p[Jc.at(i, j, k)];
    double const resistive = A.eta * jOnE;
    // hyperresistive_ (ohm.hpp:211-268)
    double hyper;
    if (A.hyper_mode == 0)
        hyper = -A.nu * laplacian<DIM>(A.L, Jc, i, j, k);
    else
    {
        double const bx   = bproj(std::integral_constant<int, 0>{});
        double const by   = bproj(std::integral_constant<int, 1>{});
        double const bz   = bproj(std::integral_constant<int, 2>{});
        double const nOnE = project<DIM, MX, MY, MZ>(A.n, i, j, k);
        double const b    = sqrt(bx * bx + by * by + bz * bz);
        hyper             = -A.nu * (b / (nOnE + 0.1) + 1) * A.lvlCoeff * laplacian<DIM>(A.L, Jc, i, j, k);
    }
    return ideal + pressure + resistive + hyper;
}

template<int DIM>
__global__ void __launch_bounds__(256) ohm_kernel(const __grid_constant__ OhmParams<DIM> A)
{
    int const c = blockIdx.y;
    int i, j, k;
    if (!unravel(A.box[c], size_t(blockIdx.x) * blockDim.x + threadIdx.x, i, j, k))
        return;
    double r;
    if (c == 0)
        r = ohm_component<DIM, 0>(A, i, j, k);
    else if (c == 1)
        r = ohm_component<DIM, 1>(A, i, j, k);
    else
        r = ohm_component<DIM, 2>(A, i, j, k);
    A.E.c[c].p[A.E.c[c].at(i, j, k)] = r;
}

// ---------------------------------------------------------------- Electrons::update
template<int DIM>
struct ElectronParams
{
    DevLayout L;
    FieldView Ne, Pe;
    VecView Vi, J, Ve;
    IterBox box;
    size_t nnodes;
    double Te;
};
template<int DIM>
__global__ void __launch_bounds__(256) electrons_kernel(const __grid_constant__ ElectronParams<DIM> A)
{
    size_t const t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    // Pe = Ne*Te over the whole allocation (electrons.hpp:202-212)
    if (t < A.nnodes)
        A.Pe.p[t] = A.Ne.p[t] * A.Te;
    int i, j, k;
    if (!unravel(A.box, t, i, j, k))
        return;
    size_t const p  = A.Ne.at(i, j, k);
    double const ne = A.Ne.p[p];
    // J{x,y,z}ToMoments: DualToPrimal along the component's own direction
    double const jx = project<DIM, 2, 0, 0>(A.J.c[0], i, j, k);
    double const jy = project<DIM, 0, 2, 0>(A.J.c[1], i, j, k);
    double const jz = project<DIM, 0, 0, 2>(A.J.c[2], i, j, k);
    A.Ve.c[0].p[p]  = A.Vi.c[0].p[p] - jx / ne;
    A.Ve.c[1].p[p]  = A.Vi.c[1].p[p] - jy / ne;
    A.Ve.c[2].p[p]  = A.Vi.c[2].p[p] - jz / ne;
}

// ---------------------------------------------------------------- ion totals, average
constexpr int MAX_POP = 8;
struct TotalsParams
{
    size_t n;
    int npop;
    const double* rho_n[MAX_POP];
    const double* rho_q[MAX_POP];
    const double* flux[MAX_POP][3];
    double mass[MAX_POP];
    double *rho_q_tot, *rho_m_tot, *V[3];
};
__global__ void __launch_bounds__(256) totals_kernel(const __grid_constant__ TotalsParams A)
{
    size_t const p = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= A.n)
        return;
    double q = 0., m = 0., vx = 0., vy = 0., vz = 0.;
    for (int s = 0; s < A.npop; ++s)
    {
        q  = q + A.rho_q[s][p];
        m  = m + A.rho_n[s][p] * A.mass[s];
        vx = vx + A.flux[s][0][p] * A.mass[s];
        vy = vy + A.flux[s][1][p] * A.mass[s];
        vz = vz + A.flux[s][2][p] * A.mass[s];
    }
    // outputs may be absent: Ions::computeChargeDensity / computeMassDensity / computeBulkVelocity called on their own
    if (A.rho_q_tot)
        A.rho_q_tot[p] = q;
    if (A.rho_m_tot)
        A.rho_m_tot[p] = m;
    if (A.V[0])
    {
        A.V[0][p] = vx / m;
        A.V[1][p] = vy / m;
        A.V[2][p] = vz / m;
    }
}
__global__ void __launch_bounds__(256)
    average_kernel(size_t n, const double* a, const double* b, double* avg)
{
    size_t const p = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p < n)
        avg[p] = (a[p] + b[p]) * .5;
}

// several averages in one launch (the six components of average_, solver_ppc.hpp:484-510): array in blockIdx.y
struct AverageMany
{
    size_t n[8];
    const double* a[8];
    const double* b[8];
    double* avg[8];
};
__global__ void __launch_bounds__(256) average_many_kernel(const __grid_constant__ AverageMany A)
{
    int const k    = blockIdx.y;
    size_t const p = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p < A.n[k])
        A.avg[k][p] = (A.a[k][p] + A.b[k][p]) * .5;
}

inline unsigned blocks_for(size_t n) { return unsigned((n + 255) / 256); }

template<int DIM>
int faraday_dim(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* B, const phb_vecfield* E, phb_vecfield* Bnew,
                double dt)
{
    FaradayParams<DIM> A;
    A.L    = make_dev_layout(*L);
    A.B    = make_vec(A.L, B, PHB_BX);
    A.E    = make_vec(A.L, E, PHB_EX);
    A.Bnew = make_vec(A.L, Bnew, PHB_BX);
    A.dt   = dt;
    size_t vmax = 0;
    for (int c = 0; c < 3; ++c)
    {
        A.box[c] = phys_box(A.L, PHB_BX + c);
        vmax     = std::max(vmax, A.box[c].volume());
    }
    faraday_kernel<DIM><<<dim3(blocks_for(vmax), 3), 256, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
template<int DIM>
int ampere_dim(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* B, phb_vecfield* J)
{
    AmpereParams<DIM> A;
    A.L = make_dev_layout(*L);
    A.B = make_vec(A.L, B, PHB_BX);
    A.J = make_vec(A.L, J, PHB_JX);
    size_t vmax = 0;
    for (int c = 0; c < 3; ++c)
    {
        A.box[c] = shrunk_ghost_box(A.L, PHB_JX + c);
        vmax     = std::max(vmax, A.box[c].volume());
    }
    ampere_kernel<DIM><<<dim3(blocks_for(vmax), 3), 256, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
template<int DIM>
int ohm_dim(phb_ctx* ctx, const phb_layout* L, const double* n, const phb_vecfield* Ve, const double* Pe,
            const phb_vecfield* B, const phb_vecfield* J, phb_vecfield* Enew, double eta, double nu, int hyper_mode)
{
    OhmParams<DIM> A;
    A.L          = make_dev_layout(*L);
    A.n          = make_view(A.L, n, PHB_RHO);
    A.Pe         = make_view(A.L, Pe, PHB_P);
    A.Ve         = make_vec(A.L, Ve, PHB_VX);
    A.B          = make_vec(A.L, B, PHB_BX);
    A.J          = make_vec(A.L, J, PHB_JX);
    A.E          = make_vec(A.L, Enew, PHB_EX);
    A.eta        = eta;
    A.nu         = nu;
    A.hyper_mode = hyper_mode;
    A.lvlCoeff   = 1. / std::pow(4, L->level);
    size_t vmax  = 0;
    for (int c = 0; c < 3; ++c)
    {
        A.box[c] = phys_box(A.L, PHB_EX + c);
        vmax     = std::max(vmax, A.box[c].volume());
    }
    ohm_kernel<DIM><<<dim3(blocks_for(vmax), 3), 256, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
template<int DIM>
int electrons_dim(phb_ctx* ctx, const phb_layout* L, const double* Ne, const phb_vecfield* Vi, const phb_vecfield* J,
                  double Te, phb_vecfield* Ve, double* Pe)
{
    ElectronParams<DIM> A;
    A.L      = make_dev_layout(*L);
    A.Ne     = make_view(A.L, Ne, PHB_RHO);
    A.Pe     = make_view(A.L, Pe, PHB_P);
    A.Vi     = make_vec(A.L, Vi, PHB_VX);
    A.J      = make_vec(A.L, J, PHB_JX);
    A.Ve     = make_vec(A.L, Ve, PHB_VX);
    A.box    = phys_box(A.L, PHB_RHO);
    A.nnodes = size_t(A.Ne.n[0]) * A.Ne.n[1] * A.Ne.n[2];
    A.Te     = Te;
    electrons_kernel<DIM><<<blocks_for(A.nnodes), 256, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
} // namespace phb

#define PHB_DIM_SWITCH(L, fn, ...)                                                                       \
    switch ((L)->dim)                                                                                    \
    {                                                                                                    \
        case 1: return phb::fn<1>(__VA_ARGS__);                                                          \
        case 2: return phb::fn<2>(__VA_ARGS__);                                                          \
        default: return phb::fn<3>(__VA_ARGS__);                                                         \
    }

extern "C" {
int phb_faraday(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* B, const phb_vecfield* E, phb_vecfield* Bnew,
                double dt)
{
    // "Error - Faraday - not all VecField parameters are usable" (faraday.hpp:30-31)
    if (!phb::valid_layout(ctx, L) || !B || !E || !Bnew)
        return phb::set_error(ctx, PHB_ERR_INVALID, "Error - Faraday - not all VecField parameters are usable");
    for (int c = 0; c < 3; ++c)
        if (!B->comp[c] || !E->comp[c] || !Bnew->comp[c])
            return phb::set_error(ctx, PHB_ERR_INVALID, "Error - Faraday - not all VecField parameters are usable");
    PHB_DIM_SWITCH(L, faraday_dim, ctx, L, B, E, Bnew, dt)
}
int phb_ampere(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* B, phb_vecfield* J)
{
    if (!phb::valid_layout(ctx, L) || !B || !J)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_ampere: invalid argument");
    PHB_DIM_SWITCH(L, ampere_dim, ctx, L, B, J)
}
int phb_ohm(phb_ctx* ctx, const phb_layout* L, const double* n, const phb_vecfield* Ve, const double* Pe,
            const phb_vecfield* B, const phb_vecfield* J, phb_vecfield* Enew, double eta, double nu, int hyper_mode)
{
    if (!phb::valid_layout(ctx, L) || !n || !Ve || !Pe || !B || !J || !Enew)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_ohm: invalid argument");
    if (hyper_mode != 0 && hyper_mode != 1)
        return phb::set_error(ctx, PHB_ERR_INVALID, "Error - Ohm - unknown hyper_mode");
    PHB_DIM_SWITCH(L, ohm_dim, ctx, L, n, Ve, Pe, B, J, Enew, eta, nu, hyper_mode)
}
int phb_electrons_update(phb_ctx* ctx, const phb_layout* L, const double* Ne, const phb_vecfield* Vi,
                         const phb_vecfield* J, double Te, phb_vecfield* Ve, double* Pe)
{
    if (!phb::valid_layout(ctx, L) || !Ne || !Vi || !J || !Ve || !Pe)
        return phb::set_error(ctx, PHB_ERR_INVALID, "Error - Electron  is not usable");
    PHB_DIM_SWITCH(L, electrons_dim, ctx, L, Ne, Vi, J, Te, Ve, Pe)
}
int phb_ions_totals(phb_ctx* ctx, size_t nnodes, int npop, const double* const* h_rho_n, const double* const* h_rho_q,
                    const phb_vecfield* h_flux, const double* h_mass, double* rho_q_tot, double* rho_m_tot,
                    phb_vecfield* V)
{
    if (!ctx || npop < 1 || npop > phb::MAX_POP || !h_rho_n || !h_rho_q || !h_flux || !h_mass
        || (!rho_q_tot && !rho_m_tot && !V))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_ions_totals: invalid argument (npop <= 8)");
    phb::TotalsParams A;
    A.n    = nnodes;
    A.npop = npop;
    for (int s = 0; s < npop; ++s)
    {
        A.rho_n[s] = h_rho_n[s];
        A.rho_q[s] = h_rho_q[s];
        for (int c = 0; c < 3; ++c)
            A.flux[s][c] = h_flux[s].comp[c];
        A.mass[s] = h_mass[s];
    }
    A.rho_q_tot = rho_q_tot;
    A.rho_m_tot = rho_m_tot;
    for (int c = 0; c < 3; ++c)
        A.V[c] = V ? V->comp[c] : nullptr;
    phb::totals_kernel<<<phb::blocks_for(nnodes), 256, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
int phb_average_many(phb_ctx* ctx, int count, const size_t* n, const double* const* a, const double* const* b,
                     double* const* avg)
{
    if (!ctx || count < 1 || count > 8 || !n || !a || !b || !avg)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_average_many: invalid argument");
    phb::AverageMany A{};
    size_t most = 0;
    for (int k = 0; k < count; ++k)
    {
        if (!a[k] || !b[k] || !avg[k])
            return phb::set_error(ctx, PHB_ERR_INVALID, "phb_average_many: invalid argument");
        A.n[k] = n[k], A.a[k] = a[k], A.b[k] = b[k], A.avg[k] = avg[k];
        most = n[k] > most ? n[k] : most;
    }
    if (most)
    {
        phb::average_many_kernel<<<dim3(phb::blocks_for(most), unsigned(count)), 256, 0, ctx->stream>>>(A);
        PHB_LAUNCH_CHECK(ctx);
    }
    return PHB_OK;
}
int phb_average(phb_ctx* ctx, size_t n, const double* a, const double* b, double* avg)
{
    if (!ctx || !a || !b || !avg)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_average: invalid argument");
    phb::average_kernel<<<phb::blocks_for(n), 256, 0, ctx->stream>>>(n, a, b, avg);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
}
