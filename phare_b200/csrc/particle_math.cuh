// Per-particle arithmetic of the hot path, in the reference's operation order.
// Compiled with -fmad=false: `a*b+c` is two roundings, exactly like the reference's default
// (no -march) build.  EXACT=false replaces the multiply-add chains of the gather and of the Boris
// rotation by explicit fma() (one rounding) — positions and cell indices stay unfused in both modes.
#pragma once
#include "common.cuh"

namespace phb
{
template<bool EXACT>
__device__ __forceinline__ double mad(double a, double b, double c)
{
    if constexpr (EXACT)
        return a * b + c;
    else
        return fma(a, b, c);
}

// computeStartLeftShift, interpolator.hpp:519-549
template<int ORDER, int CENTER>
__device__ __forceinline__ int start_shift(double delta)
{
    if constexpr (ORDER == 1)
        return CENTER == PRIMAL ? 0 : (delta < .5 ? 1 : 0);
    else if constexpr (ORDER == 2)
        return CENTER == PRIMAL ? (delta < .5 ? 1 : 0) : 1;
    else
        return CENTER == PRIMAL ? 1 : (delta < .5 ? 2 : 1);
}

// Weighter<ORDER>::computeWeight (interpolator.hpp:54-125) + indexAndWeights_ (:381-406).
// l = local (ghost-offset) cell index, dl = double(l) (the reference converts its uint32 iCell);
// returns the start index, fills w[ORDER+1].  double(start) is formed as dl - double(shift): both are
// small integers, so the subtraction is exact and equals the reference's int -> double conversion
// (saves the slow I2F.F64 per centering and direction).
template<int ORDER, int CENTER>
__device__ __forceinline__ int index_and_weights(int l, double dl, double delta, double (&w)[ORDER + 1])
{
    int const shift    = start_shift<ORDER, CENTER>(delta);
    int const start    = l - shift;
    double const dstart = dl - double(shift); // shift in {0,1,2}: a select of constants, no conversion
    double x           = dl + delta;
    if constexpr (CENTER == DUAL)
        x -= .5;
    if constexpr (ORDER == 1)
    {
        w[1] = x - dstart;
        w[0] = 1. - w[1];
    }
    else if constexpr (ORDER == 2)
    {
        double const d     = (dstart + 1.) - x; // double(start + 1), exact
        double const coef1 = 0.5 + d, coef2 = d, coef3 = 0.5 - d;
        w[0] = 0.5 * coef1 * coef1;
        w[1] = 0.75 - coef2 * coef2;
        w[2] = 0.5 * coef3 * coef3;
    }
    else
    {
        constexpr double _4_over_3 = 4. / 3., _2_over_3 = 2. / 3.;
        double const index = dstart - x;
        double const coef1 = 1. + 0.5 * index, coef2 = index + 1, coef3 = index + 2;
        double const coef4 = 1. - 0.5 * (index + 3);
        double const coef2_sq = coef2 * coef2, coef2_cub = coef2_sq * coef2;
        double const coef3_sq = coef3 * coef3, coef3_cub = coef3_sq * coef3;
        w[0] = _4_over_3 * coef1 * coef1 * coef1;
        w[1] = _2_over_3 - coef2_sq - 0.5 * coef2_cub;
        w[2] = _2_over_3 - coef3_sq + 0.5 * coef3_cub;
        w[3] = _4_over_3 * coef4 * coef4 * coef4;
    }
    return start;
}
template<int ORDER, int CENTER>
__device__ __forceinline__ int index_and_weights(int l, double delta, double (&w)[ORDER + 1])
{
    return index_and_weights<ORDER, CENTER>(l, double(unsigned(l)), delta, w);
}

template<int DIM, int ORDER>
struct IndexWeights
{
    int start[2][DIM];             // [centering][dir]
    double w[2][DIM][ORDER + 1];
};

template<int DIM, int ORDER>
__device__ __forceinline__ void both_centerings(const DevLayout& L, const int* icell, const double* delta,
                                                IndexWeights<DIM, ORDER>& iw)
{
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        int const l     = icell[d] - (L.amr_lower[d] - L.g); // AMRToLocal, gridlayout.hpp:746-763
        double const dl = double(unsigned(l));
        iw.start[DUAL][d]   = index_and_weights<ORDER, DUAL>(l, dl, delta[d], iw.w[DUAL][d]);
        iw.start[PRIMAL][d] = index_and_weights<ORDER, PRIMAL>(l, dl, delta[d], iw.w[PRIMAL][d]);
    }
}

// MeshToParticle<DIM>::operator(), interpolator.hpp:152-264: nested z -> y -> x accumulation.
// One base pointer per component, one row pointer per (ix,iy), immediate offsets along the row.
template<int DIM, int ORDER, int QTY, bool EXACT>
__device__ __forceinline__ double gather(const IndexWeights<DIM, ORDER>& iw, const FieldView& f)
{
    constexpr int cx = centering(QTY, 0), cy = centering(QTY, 1), cz = centering(QTY, 2);
    double F = 0.;
    if constexpr (DIM == 1)
    {
        const double* row = f.p + iw.start[cx][0];
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
            F = mad<EXACT>(__ldg(row + ix), iw.w[cx][0][ix], F);
    }
    else if constexpr (DIM == 2)
    {
        const double* base = f.p + (iw.start[cx][0] * f.n[1] + iw.start[cy][1]);
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
        {
            const double* row = base + ix * f.n[1];
            double Y          = 0.;
#pragma unroll
            for (int iy = 0; iy <= ORDER; ++iy)
                Y = mad<EXACT>(__ldg(row + iy), iw.w[cy][1][iy], Y);
            F = mad<EXACT>(Y, iw.w[cx][0][ix], F);
        }
    }
    else
    {
        int const s1 = f.n[2], s0 = f.n[1] * f.n[2];
        const double* base = f.p + (iw.start[cx][0] * s0 + iw.start[cy][1] * s1 + iw.start[cz][2]);
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
        {
            double Y = 0.;
#pragma unroll
            for (int iy = 0; iy <= ORDER; ++iy)
            {
                const double* row = base + (ix * s0 + iy * s1);
                double Z          = 0.;
#pragma unroll
                for (int iz = 0; iz <= ORDER; ++iz)
                    Z = mad<EXACT>(__ldg(row + iz), iw.w[cz][2][iz], Z);
                Y = mad<EXACT>(Z, iw.w[cy][1][iy], Y);
            }
            F = mad<EXACT>(Y, iw.w[cx][0][ix], F);
        }
    }
    return F;
}

// BorisPusher::accelerate_, boris.hpp:240-300
template<bool EXACT>
__device__ __forceinline__ void boris(double (&v)[3], double charge, double dto2m, const double (&E)[3],
                                      const double (&B)[3])
{
    double const coef1 = charge * dto2m;
    double velx1 = mad<EXACT>(coef1, E[0], v[0]);
    double vely1 = mad<EXACT>(coef1, E[1], v[1]);
    double velz1 = mad<EXACT>(coef1, E[2], v[2]);
    double const rx = coef1 * B[0], ry = coef1 * B[1], rz = coef1 * B[2];
    double const rx2 = rx * rx, ry2 = ry * ry, rz2 = rz * rz;
    double const rxry = rx * ry, rxrz = rx * rz, ryrz = ry * rz;
    double const invDet = 1. / (1. + rx2 + ry2 + rz2);
    double const mxx = 1. + rx2 - ry2 - rz2;
    double const mxy = 2. * (rxry + rz);
    double const mxz = 2. * (rxrz - ry);
    double const myx = 2. * (rxry - rz);
    double const myy = 1. + ry2 - rx2 - rz2;
    double const myz = 2. * (ryrz + rx);
    double const mzx = 2. * (rxrz + ry);
    double const mzy = 2. * (ryrz - rx);
    double const mzz = 1. + rz2 - rx2 - ry2;
    double velx2, vely2, velz2;
    if constexpr (EXACT)
    {
        velx2 = (mxx * velx1 + mxy * vely1 + mxz * velz1) * invDet;
        vely2 = (myx * velx1 + myy * vely1 + myz * velz1) * invDet;
        velz2 = (mzx * velx1 + mzy * vely1 + mzz * velz1) * invDet;
    }
    else
    {
        velx2 = fma(mxz, velz1, fma(mxy, vely1, mxx * velx1)) * invDet;
        vely2 = fma(myz, velz1, fma(myy, vely1, myx * velx1)) * invDet;
        velz2 = fma(mzz, velz1, fma(mzy, vely1, mzx * velx1)) * invDet;
    }
    v[0] = mad<EXACT>(coef1, E[0], velx2);
    v[1] = mad<EXACT>(coef1, E[1], vely2);
    v[2] = mad<EXACT>(coef1, E[2], velz2);
}

// BorisPusher::advancePosition_, boris.hpp:156-172 (always unfused: it decides the cell index)
template<int DIM>
__device__ __forceinline__ void advance_position(const double* h, int* icell, double* delta, const double* v,
                                                 bool& ok, double& bad_delta, double& bad_vel)
{
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        double const t = __dadd_rn(delta[d], __dmul_rn(h[d], v[d]));
        if (fabs(t) > 2 && ok) // the reference throws at the first offending direction
        {
            ok        = false;
            bad_delta = t;
            bad_vel   = v[d];
        }
        int const s = int(floor(t));
        delta[d]    = t - double(s);
        icell[d] += s;
    }
}

} // namespace phb
