// GridLayout index space and finite-difference primitives on the Yee lattice, shared by the field kernels (fields.cu) and
// by the probe entry point that the GridLayout golden vectors are checked through (probe.cu):
//   evalOnBox / evalOnShrinkedGhostBox     src/core/data/grid/gridlayout.hpp:1198-1231
//   deriv / laplacian / project            src/core/data/grid/gridlayout.hpp:550-700,784-796
//   Yee projection stencils                src/core/data/grid/impl/yee/gridlayout_hybrid_yee.hpp:350-749
#pragma once
#include "common.cuh"

namespace phb
{
struct IterBox
{
    int lo[3], n[3]; // first index and extent per direction (extent 1 in unused directions)
    __host__ __device__ size_t volume() const { return size_t(n[0]) * n[1] * n[2]; }
};

// physical box of a quantity (evalOnBox, gridlayout.hpp:1198-1206)
inline IterBox phys_box(const DevLayout& L, int qty)
{
    IterBox b;
    for (int d = 0; d < 3; ++d)
    {
        b.lo[d] = d < L.dim ? L.g : 0;
        b.n[d]  = d < L.dim ? L.ncells[d] + (centering(qty, d) == PRIMAL ? 1 : 0) : 1;
    }
    return b;
}
// ghost box shrunk by one (evalOnShrinkedGhostBox, gridlayout.hpp:1220-1231)
inline IterBox shrunk_ghost_box(const DevLayout& L, int qty)
{
    IterBox b;
    for (int d = 0; d < 3; ++d)
    {
        b.lo[d] = d < L.dim ? 1 : 0;
        b.n[d]  = d < L.dim ? alloc_extent(L, qty, d) - 2 : 1;
    }
    return b;
}

__device__ __forceinline__ bool unravel(const IterBox& b, size_t t, int& i, int& j, int& k)
{
    size_t const vol = b.volume();
    if (t >= vol)
        return false;
    if (vol <= 0xffffffffull)
    {
        // (every patch in practice) 32-bit divisions: a pair of 64-bit div/mod per thread costs more than a stencil
        unsigned const u = unsigned(t), n2 = unsigned(b.n[2]), n1 = unsigned(b.n[1]);
        unsigned const q = u / n2, q1 = q / n1;
        k = b.lo[2] + int(u - q * n2);
        j = b.lo[1] + int(q - q1 * n1);
        i = b.lo[0] + int(q1);
        return true;
    }
    k = b.lo[2] + int(t % b.n[2]);
    t /= b.n[2];
    j = b.lo[1] + int(t % b.n[1]);
    i = b.lo[0] + int(t / b.n[1]);
    return true;
}

// deriv<DIR>: primal operand -> (next, prev) = (idx+1, idx); dual operand -> (idx, idx-1)
template<int DIR>
__device__ __forceinline__ double deriv(const DevLayout& L, const FieldView& f, int qty, int i, int j, int k)
{
    int const up = centering(qty, DIR) == PRIMAL ? 1 : 0;
    int const dn = up - 1;
    double next, prev;
    if constexpr (DIR == 0)
    {
        next = f.p[f.at(i + up, j, k)];
        prev = f.p[f.at(i + dn, j, k)];
    }
    else if constexpr (DIR == 1)
    {
        next = f.p[f.at(i, j + up, k)];
        prev = f.p[f.at(i, j + dn, k)];
    }
    else
    {
        next = f.p[f.at(i, j, k + up)];
        prev = f.p[f.at(i, j, k + dn)];
    }
    return L.inv_dx[DIR] * (next - prev);
}

template<int DIM>
__device__ __forceinline__ double laplacian(const DevLayout& L, const FieldView& f, int i, int j, int k)
{
    double const here = f.p[f.at(i, j, k)];
    double lap = L.inv_dx[0] * L.inv_dx[0] * (f.p[f.at(i + 1, j, k)] - 2.0 * here + f.p[f.at(i - 1, j, k)]);
    if constexpr (DIM >= 2)
        lap = lap + L.inv_dx[1] * L.inv_dx[1] * (f.p[f.at(i, j + 1, k)] - 2.0 * here + f.p[f.at(i, j - 1, k)]);
    if constexpr (DIM >= 3)
        lap = lap + L.inv_dx[2] * L.inv_dx[2] * (f.p[f.at(i, j, k + 1)] - 2.0 * here + f.p[f.at(i, j, k - 1)]);
    return lap;
}

// project: kinds per direction 0 none, 1 PrimalToDual {0,+1}, 2 DualToPrimal {-1,0};
// directions >= DIM degenerate to the identity (directionalInterp, gridlayout_hybrid_yee.hpp:350-356)
template<int DIM, int KX, int KY, int KZ>
__device__ __forceinline__ double project(const FieldView& f, int i, int j, int k)
{
    constexpr int kx = KX, ky = DIM >= 2 ? KY : 0, kz = DIM >= 3 ? KZ : 0;
    constexpr int nx = kx ? 2 : 1, ny = ky ? 2 : 1, nz = kz ? 2 : 1;
    constexpr double coef = (kx ? .5 : 1.) * (ky ? .5 : 1.) * (kz ? .5 : 1.);
    double result = 0.;
#pragma unroll
    for (int a = 0; a < nx; ++a)
#pragma unroll
        for (int b = 0; b < ny; ++b)
#pragma unroll
            for (int c = 0; c < nz; ++c)
            {
                int const oi = kx ? (kx == 1 ? a : a - 1) : 0;
                int const oj = ky ? (ky == 1 ? b : b - 1) : 0;
                int const ok = kz ? (kz == 1 ? c : c - 1) : 0;
                result += coef * f.p[f.at(i + oi, j + oj, k + ok)];
            }
    return result;
}

} // namespace phb
