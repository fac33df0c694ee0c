// K8 — box copy / += / max between field arrays, and pack / unpack of a box to a contiguous
// buffer (NCCL staging).  These are the data movers behind the same-level messenger semantics:
//   ghost fill  : FieldData::copy / packStream / unpackStream  (src/amr/data/field/field_data.hpp:212-290)
//   border sum  : FieldBorderOp with PlusEquals               (src/amr/data/field/field_data.hpp:446-462,
//                                                              src/core/utilities/types.hpp:570-575)
//   border max  : SetMax                                       (src/core/utilities/types.hpp:577-581)
// Which boxes are exchanged is decided on the host (phare_b200/messenger.py + boxes.py restate
// field_geometry.hpp:139-304 and field_variable_fill_pattern.hpp:30-313).
#include "common.cuh"

namespace phb
{
struct BoxOpParams
{
    double* dst;
    const double* src;
    int dn[3], dlo[3], sn[3], slo[3], ext[3];
    int op;
};

__device__ __forceinline__ double apply_op(int op, double d, double s)
{
    // copy / PlusEquals / SetMax: d = std::max(d, s) = (d < s) ? s : d  (a NaN destination stays NaN, a NaN
    // source is ignored: utilities/types.hpp:577-581)
    return op == 0 ? s : op == 1 ? d + s : ((d < s) ? s : d);
}

__global__ void __launch_bounds__(256) box_op_kernel(const __grid_constant__ BoxOpParams A)
{
    size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= size_t(A.ext[0]) * A.ext[1] * A.ext[2])
        return;
    int const k = int(t % A.ext[2]);
    t /= A.ext[2];
    int const j = int(t % A.ext[1]);
    int const i = int(t / A.ext[1]);
    size_t const pd = (size_t(A.dlo[0] + i) * A.dn[1] + (A.dlo[1] + j)) * A.dn[2] + (A.dlo[2] + k);
    size_t const ps = (size_t(A.slo[0] + i) * A.sn[1] + (A.slo[1] + j)) * A.sn[2] + (A.slo[2] + k);
    A.dst[pd]       = apply_op(A.op, A.dst[pd], A.src[ps]);
}

inline void fill3(int dim, const uint32_t* src, int* dst, int dflt)
{
    for (int d = 0; d < 3; ++d)
        dst[d] = d < dim ? int(src[d]) : dflt;
}

int box_op(phb_ctx* ctx, int dim, double* dst, const uint32_t* ds, const uint32_t* dlo, const double* src,
           const uint32_t* ss, const uint32_t* slo, const uint32_t* ext, int op)
{
    BoxOpParams A;
    A.dst = dst;
    A.src = src;
    fill3(dim, ds, A.dn, 1);
    fill3(dim, dlo, A.dlo, 0);
    fill3(dim, ss, A.sn, 1);
    fill3(dim, slo, A.slo, 0);
    fill3(dim, ext, A.ext, 1);
    A.op           = op;
    size_t const n = size_t(A.ext[0]) * A.ext[1] * A.ext[2];
    if (n == 0)
        return PHB_OK;
    box_op_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

// ---- batched variant: one launch runs a whole exchange phase (every face/edge/corner box of every
// component of every patch pair).  Descriptors live in device memory and are built once per plan.
__global__ void __launch_bounds__(256)
    box_op_batch_kernel(const phb_box_desc* __restrict__ ops, int nops, unsigned long long total,
                        const DevError* __restrict__ err)
{
    unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total)
        return;
    // a neighbour GPU never delivered (phb_peer_wait timed out): its receive area holds the data of an older phase, so
    // nothing is packed or unpacked any more; the host sees PHB_ERR_PEER_TIMEOUT at its next phb_poll_error
    if (err->code == int(PHB_ERR_PEER_TIMEOUT))
        return;
    // binary search: last op whose first element index is <= t
    int lo = 0, hi = nops - 1;
    while (lo < hi)
    {
        int const mid = (lo + hi + 1) >> 1;
        if (ops[mid].first <= t)
            lo = mid;
        else
            hi = mid - 1;
    }
    const phb_box_desc& A = ops[lo];
    unsigned long long r  = t - A.first;
    unsigned const k      = unsigned(r % A.ext[2]);
    r /= A.ext[2];
    unsigned const j = unsigned(r % A.ext[1]);
    unsigned const i = unsigned(r / A.ext[1]);
    size_t const pd  = (size_t(A.dst_lo[0] + i) * A.dst_shape[1] + (A.dst_lo[1] + j)) * A.dst_shape[2] + (A.dst_lo[2] + k);
    size_t const ps  = (size_t(A.src_lo[0] + i) * A.src_shape[1] + (A.src_lo[1] + j)) * A.src_shape[2] + (A.src_lo[2] + k);
    // several overlaps of one phase may cover the same destination node (faces, edges and corners of the
    // ghost box): += and max must therefore be atomic; copies never overlap (the plan deduplicates them)
    if (A.op == 3)
    {
        // setNaNsOnFieldGhosts (hybrid_hybrid_messenger_strategy.hpp:924-955): the source is not read
        A.dst[pd] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    double const sv = A.src[ps];
    if (A.op == 0)
        A.dst[pd] = sv;
    else if (A.op == 1)
        atomicAdd(A.dst + pd, sv);
    else
    {
        auto* addr                 = reinterpret_cast<unsigned long long*>(A.dst + pd);
        unsigned long long old     = *addr;
        while (true)
        {
            double const d = __longlong_as_double((long long)old);
            if (!(d < sv))
                break;
            unsigned long long const seen = atomicCAS(addr, old, (unsigned long long)__double_as_longlong(sv));
            if (seen == old)
                break;
            old = seen;
        }
    }
}
} // namespace phb

extern "C" {
int phb_box_op(phb_ctx* ctx, int dim, double* dst, const uint32_t dst_shape[3], const uint32_t dst_lo[3],
               const double* src, const uint32_t src_shape[3], const uint32_t src_lo[3], const uint32_t extent[3],
               int op)
{
    if (!ctx || dim < 1 || dim > 3 || !dst || !src || op < 0 || op > 2)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_box_op: invalid argument");
    return phb::box_op(ctx, dim, dst, dst_shape, dst_lo, src, src_shape, src_lo, extent, op);
}
int phb_box_pack(phb_ctx* ctx, int dim, const double* src, const uint32_t src_shape[3], const uint32_t src_lo[3],
                 const uint32_t extent[3], double* buf)
{
    if (!ctx || dim < 1 || dim > 3 || !buf || !src)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_box_pack: invalid argument");
    uint32_t const zero[3] = {0, 0, 0};
    return phb::box_op(ctx, dim, buf, extent, zero, src, src_shape, src_lo, extent, 0);
}
int phb_box_unpack(phb_ctx* ctx, int dim, double* dst, const uint32_t dst_shape[3], const uint32_t dst_lo[3],
                   const uint32_t extent[3], const double* buf, int op)
{
    if (!ctx || dim < 1 || dim > 3 || !buf || !dst || op < 0 || op > 2)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_box_unpack: invalid argument");
    uint32_t const zero[3] = {0, 0, 0};
    return phb::box_op(ctx, dim, dst, dst_shape, dst_lo, buf, extent, zero, extent, op);
}
int phb_box_op_batch(phb_ctx* ctx, const phb_box_desc* d_ops, int nops, uint64_t total_elements)
{
    if (!ctx || (nops > 0 && !d_ops))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_box_op_batch: invalid argument");
    if (nops <= 0 || total_elements == 0)
        return PHB_OK;
    phb::box_op_batch_kernel<<<unsigned((total_elements + 255) / 256), 256, 0, ctx->stream>>>(d_ops, nops,
                                                                                             total_elements, ctx->d_err);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
}
