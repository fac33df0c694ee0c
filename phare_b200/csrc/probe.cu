// phb_gridlayout_probe: the GridLayout primitives the field kernels are built from (yee.cuh: deriv, laplacian, the
// linear-combination projections) applied on their own, so that the reference's GridLayout golden vectors
// (tests/core/data/gridlayout/{test_deriv,test_laplacian,test_linear_combinations_yee}.py) check THE device functions
// Faraday / Ampere / Ohm call, not a restatement of them.  Test support; nothing in the step calls it.
#include "yee.cuh"

namespace phb
{
struct ProbeParams
{
    DevLayout L;
    FieldView in, out;
    IterBox box; // indices of `out` that are written
    int qty, dir;
};

template<int DIM>
__global__ void __launch_bounds__(256) probe_deriv_kernel(const __grid_constant__ ProbeParams A)
{
    int i, j, k;
    if (!unravel(A.box, size_t(blockIdx.x) * blockDim.x + threadIdx.x, i, j, k))
        return;
    double r;
    if (A.dir == 0)
        r = deriv<0>(A.L, A.in, A.qty, i, j, k);
    else if (A.dir == 1)
        r = deriv<1>(A.L, A.in, A.qty, i, j, k);
    else
        r = deriv<2>(A.L, A.in, A.qty, i, j, k);
    A.out.p[A.out.at(i, j, k)] = r;
}

template<int DIM>
__global__ void __launch_bounds__(256) probe_laplacian_kernel(const __grid_constant__ ProbeParams A)
{
    int i, j, k;
    if (!unravel(A.box, size_t(blockIdx.x) * blockDim.x + threadIdx.x, i, j, k))
        return;
    A.out.p[A.out.at(i, j, k)] = laplacian<DIM>(A.L, A.in, i, j, k);
}

template<int DIM, int KX, int KY, int KZ>
__global__ void __launch_bounds__(256) probe_project_kernel(const __grid_constant__ ProbeParams A)
{
    int i, j, k;
    if (!unravel(A.box, size_t(blockIdx.x) * blockDim.x + threadIdx.x, i, j, k))
        return;
    A.out.p[A.out.at(i, j, k)] = project<DIM, KX, KY, KZ>(A.in, i, j, k);
}

template<int DIM, int KX, int KY>
int launch_project_z(phb_ctx* ctx, const ProbeParams& A, int kz, unsigned grid)
{
    switch (kz)
    {
        case 0: probe_project_kernel<DIM, KX, KY, 0><<<grid, 256, 0, ctx->stream>>>(A); break;
        case 1: probe_project_kernel<DIM, KX, KY, 1><<<grid, 256, 0, ctx->stream>>>(A); break;
        default: probe_project_kernel<DIM, KX, KY, 2><<<grid, 256, 0, ctx->stream>>>(A); break;
    }
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
template<int DIM, int KX>
int launch_project_y(phb_ctx* ctx, const ProbeParams& A, int ky, int kz, unsigned grid)
{
    switch (ky)
    {
        case 0: return launch_project_z<DIM, KX, 0>(ctx, A, kz, grid);
        case 1: return launch_project_z<DIM, KX, 1>(ctx, A, kz, grid);
        default: return launch_project_z<DIM, KX, 2>(ctx, A, kz, grid);
    }
}
template<int DIM>
int launch_project(phb_ctx* ctx, const ProbeParams& A, int kx, int ky, int kz, unsigned grid)
{
    switch (kx)
    {
        case 0: return launch_project_y<DIM, 0>(ctx, A, ky, kz, grid);
        case 1: return launch_project_y<DIM, 1>(ctx, A, ky, kz, grid);
        default: return launch_project_y<DIM, 2>(ctx, A, ky, kz, grid);
    }
}

template<int DIM>
int probe_dim(phb_ctx* ctx, const phb_layout* L, int op, int qty, int arg, const double* in, double* out)
{
    ProbeParams A;
    A.L   = make_dev_layout(*L);
    A.in  = make_view(A.L, in, qty);
    A.out = A.in;
    A.out.p = out;
    A.qty = qty;
    A.dir = arg;
    A.box = phys_box(A.L, qty);
    if (op == 0)
    {
        // the derivative lives on the other centering along `dir`: its array and its physical range are those of a
        // quantity primal where `qty` is dual and the reverse (allocSizeDerived, gridlayout.hpp:866-880)
        if (arg < 0 || arg >= DIM)
            return set_error(ctx, PHB_ERR_INVALID, "phb_gridlayout_probe: direction out of range");
        bool const src_primal = centering(qty, arg) == PRIMAL;
        A.out.n[arg]          = A.L.ncells[arg] + 2 * A.L.g + (src_primal ? 0 : 1);
        A.box.n[arg]          = A.L.ncells[arg] + (src_primal ? 0 : 1);
    }
    else if (op == 2)
    {
        // every index whose stencil (offsets -1 .. +1) stays inside the array
        for (int d = 0; d < DIM; ++d)
        {
            A.box.lo[d] = 1;
            A.box.n[d]  = A.in.n[d] - 2;
        }
    }
    unsigned const grid = unsigned((A.box.volume() + 255) / 256);
    if (op == 0)
        probe_deriv_kernel<DIM><<<grid, 256, 0, ctx->stream>>>(A);
    else if (op == 1)
        probe_laplacian_kernel<DIM><<<grid, 256, 0, ctx->stream>>>(A);
    else if (op == 2)
        return launch_project<DIM>(ctx, A, arg & 3, (arg >> 2) & 3, (arg >> 4) & 3, grid);
    else
        return set_error(ctx, PHB_ERR_INVALID, "phb_gridlayout_probe: unknown op");
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
} // namespace phb

extern "C" int phb_gridlayout_probe(phb_ctx* ctx, const phb_layout* L, int op, int qty, int arg, const double* d_in,
                                    double* d_out)
{
    if (!phb::valid_layout(ctx, L) || !d_in || !d_out || qty < 0 || qty >= PHB_NQTY)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_gridlayout_probe: invalid argument");
    switch (L->dim)
    {
        case 1: return phb::probe_dim<1>(ctx, L, op, qty, arg, d_in, d_out);
        case 2: return phb::probe_dim<2>(ctx, L, op, qty, arg, d_in, d_out);
        default: return phb::probe_dim<3>(ctx, L, op, qty, arg, d_in, d_out);
    }
}
