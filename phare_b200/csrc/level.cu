// Coarse <-> fine level operators (refinement ratio 2): what SAMRAI's RefineSchedule / CoarsenSchedule execute
// through PHARE's operator policies between two levels.
//   DefaultFieldRefiner        src/amr/data/field/refine/field_refiner.hpp:32-170 (+ field_linear_refine.hpp:29-129,
//                              linear_weighter.cpp:9-56)
//   MagneticFieldRefiner       src/amr/data/field/refine/magnetic_field_refiner.hpp:28-195
//   MagneticFieldInitRefiner   src/amr/data/field/refine/magnetic_field_init_refiner.hpp:27-190
//   ElectricFieldRefiner       src/amr/data/field/refine/electric_field_refiner.hpp:30-345
//   postprocessRefine          src/amr/data/field/refine/magnetic_refine_patch_strategy.hpp:66-372 (Toth & Roe 2002)
//   ElectricFieldCoarsener     src/amr/data/field/coarsening/electric_field_coarsener.hpp:38-150
//   MomentsCoarsener           src/amr/data/field/coarsening/moments_coarsener.hpp:30-84
//   setNaNsOnFieldGhosts       src/amr/messengers/hybrid_hybrid_messenger_strategy.hpp:924-955
//   PlusEqualsProduct          src/core/utilities/types.hpp:584-588 (SolverPPC::accumulateFluxSum, solver_ppc.hpp:263-276)
// One thread per node of the destination box, consecutive threads along the fastest index; pure streaming
// (HBM-bound, a few MB per call).  Same operation order as the reference, -fmad=false: bit-identical results.
// Which boxes are refined / coarsened is decided on the host (phare_b200/amr.py).
#include "common.cuh"

#include <cmath>

namespace phb
{
struct LevelView
{
    double* p;
    int n[3];  // extents, 1 in unused trailing directions
    int lo[3]; // AMR field index of element 0
    __device__ __forceinline__ size_t at(int i, int j, int k) const
    {
        return (size_t(i - lo[0]) * n[1] + (j - lo[1])) * n[2] + (k - lo[2]);
    }
};
inline LevelView make_level_view(const phb_field_view& v, int dim)
{
    LevelView r;
    r.p = v.data;
    for (int d = 0; d < 3; ++d)
    {
        r.n[d]  = d < dim ? int(v.shape[d]) : 1;
        r.lo[d] = d < dim ? v.lo[d] : 0;
    }
    return r;
}

struct LevelOpParams
{
    LevelView src, dst;
    int lo[3], ext[3]; // destination box (AMR field indices)
    int cen[3];
    int op;
};

// toCoarseIndex (amr_utils.hpp:128-135)
__device__ __forceinline__ int to_coarse(int i) { return (i >= 0) ? i / 2 : i / 2 + i % 2; }

// Thread -> node of a box without integer division: threadIdx.x runs along the fastest used direction, threadIdx.y /
// blockIdx.y along the next one, blockIdx.z along the slowest (1-D: 256 x 1 threads, 2-D / 3-D: 32 x 8).
template<int DIM>
__device__ __forceinline__ bool box_index(const int (&ext)[3], int (&idx)[3])
{
    idx[0] = idx[1] = idx[2] = 0;
    if constexpr (DIM == 1)
    {
        idx[0] = int(blockIdx.x * blockDim.x + threadIdx.x);
        return idx[0] < ext[0];
    }
    else if constexpr (DIM == 2)
    {
        idx[1] = int(blockIdx.x * blockDim.x + threadIdx.x);
        idx[0] = int(blockIdx.y * blockDim.y + threadIdx.y);
        return idx[1] < ext[1] && idx[0] < ext[0];
    }
    else
    {
        idx[2] = int(blockIdx.x * blockDim.x + threadIdx.x);
        idx[1] = int(blockIdx.y * blockDim.y + threadIdx.y);
        idx[0] = int(blockIdx.z);
        return idx[2] < ext[2] && idx[1] < ext[1];
    }
}
inline void box_launch_dims(int dim, const int ext[3], dim3& grid, dim3& block)
{
    if (dim == 1)
    {
        block = dim3(256, 1, 1);
        grid  = dim3(unsigned(ext[0] + 255) / 256, 1, 1);
    }
    else if (dim == 2)
    {
        block = dim3(32, 8, 1);
        grid  = dim3(unsigned(ext[1] + 31) / 32, unsigned(ext[0] + 7) / 8, 1);
    }
    else
    {
        block = dim3(32, 8, 1);
        grid  = dim3(unsigned(ext[2] + 31) / 32, unsigned(ext[1] + 7) / 8, unsigned(ext[0]));
    }
}

template<int DIM>
__device__ __forceinline__ bool node_of_thread(const LevelOpParams& A, int f[3])
{
    int idx[3];
    if (!box_index<DIM>(A.ext, idx))
        return false;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        f[d] = d < DIM ? A.lo[d] + idx[d] : 0;
    return true;
}

// LinearWeighter, ratio 2 (linear_weighter.cpp:9-56): weights {1-d, d}; primal d = {0, 1/2}; dual d = {3/4, 1/4}
__device__ __forceinline__ void linear_weights(int centering, int iw, double w[2])
{
    double const small = 1. / 2;
    double dist;
    if (centering == PRIMAL)
        dist = double(iw) / 2;
    else
        dist = iw == 0 ? (0.5 + double(1)) * small : (0.5 + double(0)) * small;
    w[0] = 1. - dist;
    w[1] = dist;
}

template<int DIM>
__global__ void __launch_bounds__(256) refine_kernel(const __grid_constant__ LevelOpParams A)
{
    int f[3];
    if (!node_of_thread<DIM>(A, f))
        return;
    LevelView const& C = A.src;
    LevelView const& F = A.dst;
    size_t const pf    = F.at(f[0], f[1], f[2]);
    if (A.op == PHB_REFINE_DEFAULT)
    {
        // DefaultFieldRefiner::operator() (field_refiner.hpp:58-163)
        if (!isnan(F.p[pf]))
            return;
        int start[3] = {0, 0, 0};
        double w[3][2];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            double const shift = A.cen[d] == PRIMAL ? 0. : 0.5;
            start[d]           = int(floor(double(f[d] + shift) / 2 - shift)); // coarseStartIndex
            linear_weights(A.cen[d], abs(f[d]) % 2, w[d]);                    // computeWeightIndex
        }
        double value = 0.;
        if constexpr (DIM == 1)
        {
            for (int sx = 0; sx < 2; ++sx)
                value += C.p[C.at(start[0] + sx, 0, 0)] * w[0][sx];
        }
        else if constexpr (DIM == 2)
        {
            for (int sx = 0; sx < 2; ++sx)
            {
                double Y = 0.;
                for (int sy = 0; sy < 2; ++sy)
                    Y += C.p[C.at(start[0] + sx, start[1] + sy, 0)] * w[1][sy];
                value += Y * w[0][sx];
            }
        }
        else
        {
            for (int sx = 0; sx < 2; ++sx)
            {
                double Y = 0.;
                for (int sy = 0; sy < 2; ++sy)
                {
                    double Z = 0.;
                    for (int sz = 0; sz < 2; ++sz)
                        Z += C.p[C.at(start[0] + sx, start[1] + sy, start[2] + sz)] * w[2][sz];
                    Y += Z * w[1][sy];
                }
                value += Y * w[0][sx];
            }
        }
        F.p[pf] = value;
        return;
    }
    int c[3] = {0, 0, 0};
    bool on[3]; // the fine index lies on a coarse face of direction d
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        c[d]  = to_coarse(f[d]);
        on[d] = f[d] % 2 == 0;
    }
    if (A.op == PHB_REFINE_MAGNETIC || A.op == PHB_REFINE_MAGNETIC_INIT)
    {
        // a fine face on top of a coarse face takes its value; new fine faces are left to the post-process
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            if (A.cen[d] == PRIMAL && !on[d])
                return;
        if (A.op == PHB_REFINE_MAGNETIC_INIT || isnan(F.p[pf]))
            F.p[pf] = C.p[C.at(c[0], c[1], c[2])];
        return;
    }
    // ElectricFieldRefiner
    if (!isnan(F.p[pf]))
        return;
    auto CV = [&](int dx, int dy, int dz) { return C.p[C.at(c[0] + dx, c[1] + dy, c[2] + dz)]; };
    double value;
    if constexpr (DIM == 1)
        value = CV(0, 0, 0); // refine1D_ :75-82: the coarse value whatever the centering
    else if constexpr (DIM == 2)
    {
        if (A.cen[0] == DUAL && A.cen[1] == PRIMAL) // Ex
            value = on[1] ? CV(0, 0, 0) : 0.5 * (CV(0, 0, 0) + CV(0, 1, 0));
        else if (A.cen[0] == PRIMAL && A.cen[1] == DUAL) // Ey
            value = on[0] ? CV(0, 0, 0) : 0.5 * (CV(0, 0, 0) + CV(1, 0, 0));
        else if (A.cen[0] == PRIMAL && A.cen[1] == PRIMAL) // Ez
        {
            if (on[0] && on[1])
                value = CV(0, 0, 0);
            else if (on[0])
                value = 0.5 * (CV(0, 0, 0) + CV(0, 1, 0));
            else if (on[1])
                value = 0.5 * (CV(0, 0, 0) + CV(1, 0, 0));
            else
                value = 0.25 * (CV(0, 0, 0) + CV(1, 0, 0) + CV(0, 1, 0) + CV(1, 1, 0));
        }
        else
            return;
    }
    else
    {
        if (A.cen[0] == DUAL && A.cen[1] == PRIMAL && A.cen[2] == PRIMAL) // Ex :166-196
        {
            if (on[1] && on[2])
                value = CV(0, 0, 0);
            else if (on[1])
                value = 0.5 * (CV(0, 0, 0) + CV(0, 0, 1));
            else if (on[2])
                value = 0.5 * (CV(0, 0, 0) + CV(0, 1, 0));
            else
                value = 0.25 * (CV(0, 0, 0) + CV(0, 1, 0)) + 0.25 * (CV(0, 0, 1) + CV(0, 1, 1));
        }
        else if (A.cen[0] == PRIMAL && A.cen[1] == DUAL && A.cen[2] == PRIMAL) // Ey :199-231
        {
            if (on[0] && on[2])
                value = CV(0, 0, 0);
            else if (on[0])
                value = 0.5 * (CV(0, 0, 0) + CV(0, 0, 1));
            else if (on[2])
                value = 0.5 * (CV(0, 0, 0) + CV(1, 0, 0));
            else
                value = 0.25 * (CV(0, 0, 0) + CV(1, 0, 0) + CV(0, 0, 1) + CV(1, 0, 1));
        }
        else if (A.cen[0] == PRIMAL && A.cen[1] == PRIMAL && A.cen[2] == DUAL) // Ez :234-334
        {
            if (on[0] && on[1])
                value = CV(0, 0, 0);
            else if (on[0])
                value = 0.5 * (CV(0, 0, 0) + CV(0, 1, 0));
            else if (on[1])
                value = 0.5 * (CV(0, 0, 0) + CV(1, 0, 0));
            else
                value = 0.25 * (CV(0, 0, 0) + CV(1, 0, 0) + CV(0, 1, 0) + CV(1, 1, 0));
        }
        else
            return;
    }
    F.p[pf] = value;
}

template<int DIM>
__global__ void __launch_bounds__(256) coarsen_kernel(const __grid_constant__ LevelOpParams A, int dual_dir)
{
    int c[3];
    if (!node_of_thread<DIM>(A, c))
        return;
    LevelView const& F = A.src;
    int f0[3]          = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        f0[d] = c[d] * 2; // fineStartIndex = coarseIndex * refinementRatio
    double value;
    if (dual_dir < 0)
        value = F.p[F.at(f0[0], f0[1], f0[2])]; // MomentsCoarsener / primal-primal electric component: injection
    else
    {
        int f1[3] = {f0[0], f0[1], f0[2]};
        f1[dual_dir] += 1;
        double const a = F.p[F.at(f0[0], f0[1], f0[2])], b = F.p[F.at(f1[0], f1[1], f1[2])];
        value = DIM == 1 ? 0.5 * (b + a) : 0.5 * (a + b); // operand order of :69-72 (1-D) and :83-138 (2-D, 3-D)
    }
    A.dst.p[A.dst.at(c[0], c[1], c[2])] = value;
}

// ---- magnetic post-process: new fine faces from the coarse faces around them
struct PostParams
{
    FieldView bx, by, bz;
    int lo[3], ext[3]; // field box of the component (AMR indices)
    int shift[3];      // local = AMR - shift  (GridLayout::AMRToLocal)
    int comp;
    double dx[3];
    BoxList excluded; // cell boxes whose faces are not touched (the patches of the level)
};

__device__ __forceinline__ bool post_excluded(const PostParams& A, int i, int j, int k)
{
    for (int b = 0; b < A.excluded.n; ++b)
    {
        DevBox const& e = A.excluded.b[b];
        if (i >= e.lo[0] && i <= e.hi[0] && j >= e.lo[1] && j <= e.hi[1] && k >= e.lo[2] && k <= e.hi[2])
            return true;
    }
    return false;
}

__device__ __forceinline__ int p_plus(int i, int o) { return i + 2 - o; }
__device__ __forceinline__ int p_minus(int i, int o) { return i - o; }
__device__ __forceinline__ int d_plus(int i, int o) { return i + 1 - o; }
__device__ __forceinline__ int d_minus(int i, int o) { return i - o; }

template<int DIM>
__global__ void __launch_bounds__(256) magnetic_postprocess_kernel(const __grid_constant__ PostParams A)
{
    int idx[3];
    if (!box_index<DIM>(A.ext, idx))
        return;
    int const i = A.lo[0] + idx[0];
    int const j = A.lo[1] + idx[1];
    if ((A.comp == 0 ? i : j) % 2 == 0) // isNewFineFace :127-132 (AMR indices can be negative: != 0, not == 1)
        return;
    if (post_excluded(A, i, j, 0))
        return;
    int const ix = i - A.shift[0], iy = DIM > 1 ? j - A.shift[1] : 0;
    double* const X = A.bx.p;
    double* const Y = A.by.p;
    if constexpr (DIM == 1)
        X[A.bx.at(ix, 0, 0)] = 0.5 * (X[A.bx.at(ix - 1, 0, 0)] + X[A.bx.at(ix + 1, 0, 0)]); // postprocessBx1d
    else if (A.comp == 0)
    {
        // postprocessBx2d :143-166
        int const xo = 1, yo = (j % 2 == 0) ? 0 : 1;
        X[A.bx.at(ix, iy, 0)]
            = 0.5 * (X[A.bx.at(ix - 1, iy, 0)] + X[A.bx.at(ix + 1, iy, 0)])
              + 0.25
                    * (Y[A.by.at(d_minus(ix, xo), p_minus(iy, yo), 0)] - Y[A.by.at(d_minus(ix, xo), p_plus(iy, yo), 0)]
                       - Y[A.by.at(d_plus(ix, xo), p_minus(iy, yo), 0)] + Y[A.by.at(d_plus(ix, xo), p_plus(iy, yo), 0)]);
    }
    else
    {
        // postprocessBy2d :168-190
        int const xo = (i % 2 == 0) ? 0 : 1, yo = 1;
        Y[A.by.at(ix, iy, 0)]
            = 0.5 * (Y[A.by.at(ix, iy - 1, 0)] + Y[A.by.at(ix, iy + 1, 0)])
              + 0.25
                    * (X[A.bx.at(p_minus(ix, xo), d_minus(iy, yo), 0)] - X[A.bx.at(p_plus(ix, xo), d_minus(iy, yo), 0)]
                       - X[A.bx.at(p_minus(ix, xo), d_plus(iy, yo), 0)] + X[A.bx.at(p_plus(ix, xo), d_plus(iy, yo), 0)]);
    }
}


// 3-D: postprocessBx3d :192-262, postprocessBy3d :264-322, postprocessBz3d :324-372.  The sums are written out in the
// reference's order; BX/BY/BZ(sx, sy, sz) pick the minus (0) or plus (1) neighbour per direction: p_* along the
// component's own (primal) direction, d_* along the two others.
__global__ void __launch_bounds__(256) magnetic_postprocess_3d_kernel(const __grid_constant__ PostParams A)
{
    int idx[3];
    if (!box_index<3>(A.ext, idx))
        return;
    int const i = A.lo[0] + idx[0], j = A.lo[1] + idx[1], k = A.lo[2] + idx[2];
    if ((A.comp == 0 ? i : A.comp == 1 ? j : k) % 2 == 0)
        return;
    if (post_excluded(A, i, j, k))
        return;
    int const ix = i - A.shift[0], iy = j - A.shift[1], iz = k - A.shift[2];
    int const xo = A.comp == 0 ? 1 : ((i % 2 == 0) ? 0 : 1);
    int const yo = A.comp == 1 ? 1 : ((j % 2 == 0) ? 0 : 1);
    int const zo = A.comp == 2 ? 1 : ((k % 2 == 0) ? 0 : 1);
    double const Dx = A.dx[0], Dy = A.dx[1], Dz = A.dx[2];
    auto fac = [](int o) { return o == 0 ? -1 : 1; }; // ijk_factor_ :383
    auto BX  = [&](int sx, int sy, int sz) {
        return A.bx.p[A.bx.at(sx ? p_plus(ix, xo) : p_minus(ix, xo), sy ? d_plus(iy, yo) : d_minus(iy, yo),
                              sz ? d_plus(iz, zo) : d_minus(iz, zo))];
    };
    auto BY = [&](int sx, int sy, int sz) {
        return A.by.p[A.by.at(sx ? d_plus(ix, xo) : d_minus(ix, xo), sy ? p_plus(iy, yo) : p_minus(iy, yo),
                              sz ? d_plus(iz, zo) : d_minus(iz, zo))];
    };
    auto BZ = [&](int sx, int sy, int sz) {
        return A.bz.p[A.bz.at(sx ? d_plus(ix, xo) : d_minus(ix, xo), sy ? d_plus(iy, yo) : d_minus(iy, yo),
                              sz ? p_plus(iz, zo) : p_minus(iz, zo))];
    };
    if (A.comp == 0)
    {
        double* const X = A.bx.p;
        X[A.bx.at(ix, iy, iz)]
            = 0.5 * (X[A.bx.at(ix - 1, iy, iz)] + X[A.bx.at(ix + 1, iy, iz)])
              + 0.125 * (BY(0, 0, 0) - BY(0, 1, 0) - BY(1, 0, 0) + BY(1, 1, 0) + BY(0, 0, 1) - BY(0, 1, 1) - BY(1, 0, 1) + BY(1, 1, 1))
              + 0.125 * (BZ(0, 0, 0) + BZ(0, 1, 0) - BZ(1, 0, 0) - BZ(1, 1, 0) - BZ(0, 0, 1) - BZ(0, 1, 1) + BZ(1, 0, 1) + BZ(1, 1, 1))
              + (0.125 * fac(zo) * Dz * Dz / (Dx * Dx + Dz * Dz))
                    * (BY(1, 1, 1) - BY(0, 1, 1) - BY(1, 0, 1) - BY(1, 1, 0) + BY(1, 0, 0) + BY(0, 1, 0) + BY(0, 0, 1) - BY(0, 0, 0))
              + (0.125 * fac(yo) * Dy * Dy / (Dx * Dx + Dy * Dy))
                    * (BZ(1, 1, 1) - BZ(0, 1, 1) - BZ(1, 0, 1) - BZ(1, 1, 0) + BZ(1, 0, 0) + BZ(0, 1, 0) + BZ(0, 0, 1) - BZ(0, 0, 0));
    }
    else if (A.comp == 1)
    {
        double* const Y = A.by.p;
        Y[A.by.at(ix, iy, iz)]
            = 0.5 * (Y[A.by.at(ix, iy - 1, iz)] + Y[A.by.at(ix, iy + 1, iz)])
              + 0.125 * (BX(0, 0, 0) - BX(0, 1, 0) - BX(1, 0, 0) + BX(1, 1, 0) + BX(0, 0, 1) - BX(0, 1, 1) - BX(1, 0, 1) + BX(1, 1, 1))
              + 0.125 * (BZ(0, 0, 0) - BZ(0, 1, 0) + BZ(1, 0, 0) - BZ(1, 1, 0) - BZ(0, 0, 1) + BZ(0, 1, 1) - BZ(1, 0, 1) + BZ(1, 1, 1))
              + (0.125 * fac(xo) * Dx * Dx / (Dx * Dx + Dy * Dy))
                    * (BZ(1, 1, 1) - BZ(0, 1, 1) - BZ(1, 0, 1) - BZ(1, 1, 0) + BZ(1, 0, 0) + BZ(0, 1, 0) + BZ(0, 0, 1) - BZ(0, 0, 0))
              + (0.125 * fac(zo) * Dz * Dz / (Dy * Dy + Dz * Dz))
                    * (BX(1, 1, 1) - BX(0, 1, 1) - BX(1, 0, 1) - BX(1, 1, 0) + BX(1, 0, 0) + BX(0, 1, 0) + BX(0, 0, 1) - BX(0, 0, 0));
    }
    else
    {
        double* const Z = A.bz.p;
        Z[A.bz.at(ix, iy, iz)]
            = 0.5 * (Z[A.bz.at(ix, iy, iz - 1)] + Z[A.bz.at(ix, iy, iz + 1)])
              + 0.125 * (BX(0, 0, 0) + BX(0, 1, 0) - BX(1, 0, 0) - BX(1, 1, 0) - BX(0, 0, 1) - BX(0, 1, 1) + BX(1, 0, 1) + BX(1, 1, 1))
              + 0.125 * (BY(0, 0, 0) - BY(0, 1, 0) + BY(1, 0, 0) - BY(1, 1, 0) - BY(0, 0, 1) + BY(0, 1, 1) - BY(1, 0, 1) + BY(1, 1, 1))
              + (0.125 * fac(yo) * Dy * Dy / (Dy * Dy + Dz * Dz))
                    * (BX(1, 1, 1) - BX(0, 1, 1) - BX(1, 0, 1) - BX(1, 1, 0) + BX(1, 0, 0) + BX(0, 1, 0) + BX(0, 0, 1) - BX(0, 0, 0))
              + (0.125 * fac(xo) * Dx * Dx / (Dx * Dx + Dz * Dz))
                    * (BY(1, 1, 1) - BY(0, 1, 1) - BY(1, 0, 1) - BY(1, 1, 0) + BY(1, 0, 0) + BY(0, 1, 0) + BY(0, 0, 1) - BY(0, 0, 0));
    }
}

struct FillParams
{
    double* dst;
    int dn[3], dlo[3], ext[3]; // left-aligned: unused trailing directions are 1 / 0 / 1
    double value;
};
template<int DIM>
__global__ void __launch_bounds__(256) box_fill_kernel(const __grid_constant__ FillParams A)
{
    int idx[3];
    if (!box_index<DIM>(A.ext, idx))
        return;
    A.dst[(size_t(A.dlo[0] + idx[0]) * A.dn[1] + (A.dlo[1] + idx[1])) * A.dn[2] + (A.dlo[2] + idx[2])] = A.value;
}

__global__ void __launch_bounds__(256) axpy_kernel(size_t n, double* __restrict__ dst, const double* __restrict__ src,
                                                   double coef)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = __dadd_rn(dst[i], __dmul_rn(src[i], coef)); // d += d0 * o, separately rounded
}

static bool fill_level_params(LevelOpParams& A, int dim, int qty, const phb_field_view* src, const phb_field_view* dst,
                              const phb_box* box, size_t& n)
{
    A.src = make_level_view(*src, dim);
    A.dst = make_level_view(*dst, dim);
    n     = 1;
    for (int d = 0; d < 3; ++d)
    {
        A.lo[d]  = d < dim ? box->lower[d] : 0;
        A.ext[d] = d < dim ? box->upper[d] - box->lower[d] + 1 : 1;
        A.cen[d] = d < dim ? centering(qty, d) : PRIMAL;
        if (A.ext[d] <= 0)
        {
            n = 0;
            return true;
        }
        n *= size_t(A.ext[d]);
        if (d < dim && (A.lo[d] < A.dst.lo[d] || A.lo[d] + A.ext[d] > A.dst.lo[d] + A.dst.n[d]))
            return false; // destination box outside the destination array
    }
    return true;
}
} // namespace phb

extern "C" {
int phb_field_refine(phb_ctx* ctx, int dim, int op, int qty, const phb_field_view* coarse, const phb_field_view* fine,
                     const phb_box* fine_box)
{
    using namespace phb;
    if (!ctx || dim < 1 || dim > 3 || op < 0 || op > PHB_REFINE_ELECTRIC || qty < 0 || qty >= PHB_NQTY || !coarse || !fine
        || !fine_box || !coarse->data || !fine->data)
        return set_error(ctx, PHB_ERR_INVALID, "phb_field_refine: invalid argument");
    LevelOpParams A;
    size_t n;
    if (!fill_level_params(A, dim, qty, coarse, fine, fine_box, n))
        return set_error(ctx, PHB_ERR_INVALID, "phb_field_refine: box outside the fine array");
    A.op = op;
    if (n == 0)
        return PHB_OK;
    // the coarse view must hold every coarse node the box reads (+1 for the two-point stencils)
    int const reach = (op == PHB_REFINE_MAGNETIC || op == PHB_REFINE_MAGNETIC_INIT) ? 0 : 1;
    for (int d = 0; d < dim; ++d)
    {
        int const flo = A.lo[d], fhi = A.lo[d] + A.ext[d] - 1;
        int clo = int(std::floor(flo / 2.)), chi = int(std::floor(fhi / 2.)) + reach;
        if (op == PHB_REFINE_DEFAULT && A.cen[d] == DUAL)
        {
            clo = int(std::floor((flo + 0.5) / 2 - 0.5));
            chi = int(std::floor((fhi + 0.5) / 2 - 0.5)) + 1;
        }
        if (clo < A.src.lo[d] || chi > A.src.lo[d] + A.src.n[d] - 1)
            return set_error(ctx, PHB_ERR_INVALID, "phb_field_refine: the coarse view does not cover the stencil");
    }
    dim3 grid, block;
    box_launch_dims(dim, A.ext, grid, block);
    if (dim == 1)
        refine_kernel<1><<<grid, block, 0, ctx->stream>>>(A);
    else if (dim == 2)
        refine_kernel<2><<<grid, block, 0, ctx->stream>>>(A);
    else
        refine_kernel<3><<<grid, block, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

int phb_field_coarsen(phb_ctx* ctx, int dim, int op, int qty, const phb_field_view* fine, const phb_field_view* coarse,
                      const phb_box* coarse_box)
{
    using namespace phb;
    if (!ctx || dim < 1 || dim > 3 || op < 0 || op > PHB_COARSEN_MOMENTS || qty < 0 || qty >= PHB_NQTY || !coarse || !fine
        || !coarse_box || !coarse->data || !fine->data)
        return set_error(ctx, PHB_ERR_INVALID, "phb_field_coarsen: invalid argument");
    LevelOpParams A;
    size_t n;
    if (!fill_level_params(A, dim, qty, fine, coarse, coarse_box, n))
        return set_error(ctx, PHB_ERR_INVALID, "phb_field_coarsen: box outside the coarse array");
    A.op         = op;
    int dual_dir = -1, ndual = 0;
    for (int d = 0; d < dim; ++d)
        if (A.cen[d] == DUAL)
        {
            dual_dir = d;
            ++ndual;
        }
    if ((op == PHB_COARSEN_MOMENTS && ndual) || ndual > 1)
        return set_error(ctx, PHB_ERR_INVALID, "phb_field_coarsen: centering not handled by this coarsener");
    if (n == 0)
        return PHB_OK;
    for (int d = 0; d < dim; ++d)
    {
        int const flo = 2 * A.lo[d], fhi = 2 * (A.lo[d] + A.ext[d] - 1) + (d == dual_dir ? 1 : 0);
        if (flo < A.src.lo[d] || fhi > A.src.lo[d] + A.src.n[d] - 1)
            return set_error(ctx, PHB_ERR_INVALID, "phb_field_coarsen: the fine view does not cover the box");
    }
    dim3 grid, block;
    box_launch_dims(dim, A.ext, grid, block);
    if (dim == 1)
        coarsen_kernel<1><<<grid, block, 0, ctx->stream>>>(A, dual_dir);
    else if (dim == 2)
        coarsen_kernel<2><<<grid, block, 0, ctx->stream>>>(A, dual_dir);
    else
        coarsen_kernel<3><<<grid, block, 0, ctx->stream>>>(A, dual_dir);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

int phb_magnetic_postprocess(phb_ctx* ctx, const phb_layout* fine, const phb_vecfield* B, const phb_box* cells,
                             const phb_box* excluded, int nexcluded)
{
    using namespace phb;
    if (!ctx || !valid_layout(ctx, fine) || !B || !cells || !B->comp[0] || !B->comp[1] || (fine->dim == 3 && !B->comp[2])
        || nexcluded < 0 || nexcluded > MAX_BOXES || (nexcluded && !excluded))
        return set_error(ctx, PHB_ERR_INVALID, "phb_magnetic_postprocess: invalid argument");
    DevLayout const L = make_dev_layout(*fine);
    PostParams A;
    A.bx = make_view(L, B->comp[0], PHB_BX);
    A.by = make_view(L, B->comp[1], PHB_BY);
    A.bz = make_view(L, B->comp[2], PHB_BZ);
    for (int d = 0; d < 3; ++d)
        A.dx[d] = L.dx[d];
    A.excluded.n = nexcluded;
    for (int b = 0; b < nexcluded; ++b)
        A.excluded.b[b] = make_box(excluded[b], L.dim);
    for (int comp = 0; comp < L.dim; ++comp)
    {
        size_t n = 1;
        for (int d = 0; d < 3; ++d)
        {
            // toFieldBox: one more node on the upper side of a primal direction (field_geometry.hpp:139-181)
            A.lo[d]    = d < L.dim ? cells->lower[d] : 0;
            A.ext[d]   = d < L.dim ? cells->upper[d] - cells->lower[d] + 1 + (centering(PHB_BX + comp, d) == PRIMAL ? 1 : 0) : 1;
            A.shift[d] = d < L.dim ? L.amr_lower[d] - L.g : 0;
            if (A.ext[d] <= 0)
                n = 0;
            else
                n *= size_t(A.ext[d]);
            // the stencil reaches one node beyond an odd face: the box must stay inside the ghost box
            if (d < L.dim && (A.lo[d] < A.shift[d] || A.lo[d] + A.ext[d] > A.shift[d] + alloc_extent(L, PHB_BX + comp, d)))
                return set_error(ctx, PHB_ERR_INVALID, "phb_magnetic_postprocess: box outside the ghost box");
        }
        if (n == 0)
            continue;
        A.comp = comp;
        dim3 grid, block;
        box_launch_dims(L.dim, A.ext, grid, block);
        if (L.dim == 1)
            magnetic_postprocess_kernel<1><<<grid, block, 0, ctx->stream>>>(A);
        else if (L.dim == 2)
            magnetic_postprocess_kernel<2><<<grid, block, 0, ctx->stream>>>(A);
        else
            magnetic_postprocess_3d_kernel<<<grid, block, 0, ctx->stream>>>(A);
        PHB_LAUNCH_CHECK(ctx);
    }
    return PHB_OK;
}

int phb_box_fill(phb_ctx* ctx, int dim, double* dst, const uint32_t dst_shape[3], const uint32_t dst_lo[3],
                 const uint32_t extent[3], double value)
{
    using namespace phb;
    if (!ctx || dim < 1 || dim > 3 || !dst)
        return set_error(ctx, PHB_ERR_INVALID, "phb_box_fill: invalid argument");
    FillParams A;
    A.dst    = dst;
    size_t n = 1;
    for (int d = 0; d < 3; ++d)
    {
        A.dn[d]  = d < dim ? int(dst_shape[d]) : 1;
        A.dlo[d] = d < dim ? int(dst_lo[d]) : 0;
        A.ext[d] = d < dim ? int(extent[d]) : 1;
        n *= size_t(A.ext[d]);
    }
    A.value = value;
    if (n == 0)
        return PHB_OK;
    dim3 grid, block;
    box_launch_dims(dim, A.ext, grid, block);
    if (dim == 1)
        box_fill_kernel<1><<<grid, block, 0, ctx->stream>>>(A);
    else if (dim == 2)
        box_fill_kernel<2><<<grid, block, 0, ctx->stream>>>(A);
    else
        box_fill_kernel<3><<<grid, block, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

int phb_axpy(phb_ctx* ctx, size_t n, double* dst, const double* src, double coef)
{
    using namespace phb;
    if (!ctx || (n && (!dst || !src)))
        return set_error(ctx, PHB_ERR_INVALID, "phb_axpy: invalid argument");
    if (n == 0)
        return PHB_OK;
    axpy_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(n, dst, src, coef);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
}
