// Device-resident SoA particle store: allocation, host interop (AoS records of the reference's
// Particle<dim>, src/core/data/particles/particle.hpp:38-73, and the ContiguousParticles SoA
// layout of particle_array.hpp:250-354), and column-wise copies.
#include "common.cuh"

#include <vector>

namespace phb
{
// byte layout of Particle<dim>: weight, charge, iCell[dim] (padded to 8), delta[dim], v[3]
__host__ __device__ inline size_t aos_stride(int dim) { return dim == 1 ? 56 : dim == 2 ? 64 : 80; }
__host__ __device__ inline size_t aos_delta_offset(int dim) { return dim == 2 ? 24 : (dim == 1 ? 24 : 32); }

template<int DIM, bool TO_SOA>
__global__ void __launch_bounds__(256) aos_kernel(unsigned char* aos, PartView P, size_t n)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    unsigned char* rec = aos + i * aos_stride(DIM);
    double* hd         = reinterpret_cast<double*>(rec);
    int* ic            = reinterpret_cast<int*>(rec + 16);
    double* de         = reinterpret_cast<double*>(rec + aos_delta_offset(DIM));
    double* v          = de + DIM;
    if constexpr (TO_SOA)
    {
        P.weight[i] = hd[0];
        P.charge[i] = hd[1];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            P.icell[d][i] = ic[d];
            P.delta[d][i] = de[d];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            P.v[c][i] = v[c];
    }
    else
    {
        hd[0] = P.weight[i];
        hd[1] = P.charge[i];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            ic[d] = P.icell[d][i];
            de[d] = P.delta[d][i];
        }
        if (DIM != 2)
            ic[DIM] = 0; // padding
#pragma unroll
        for (int c = 0; c < 3; ++c)
            v[c] = P.v[c][i];
    }
}

// interleaved host SoA (iCell[n*dim], delta[n*dim], v[n*3]) <-> one column per component
template<typename T, bool TO_COLUMNS>
__global__ void __launch_bounds__(256) interleave_kernel(T* packed, T* c0, T* c1, T* c2, int ncomp, size_t n)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    T* cols[3] = {c0, c1, c2};
    for (int c = 0; c < ncomp; ++c)
    {
        if constexpr (TO_COLUMNS)
            cols[c][i] = packed[i * ncomp + c];
        else
            packed[i * ncomp + c] = cols[c][i];
    }
}
} // namespace phb

extern "C" {

size_t phb_aos_stride(int dim) { return phb::aos_stride(dim); }

int phb_particles_alloc(phb_ctx* ctx, size_t capacity, phb_particles* out)
{
    if (!ctx || !out)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_particles_alloc: invalid argument");
    *out          = phb_particles{};
    size_t const cap = capacity ? capacity : 1;
    // one allocation, columns 256-byte aligned
    size_t const col8 = ((cap * 8 + 255) / 256) * 256, col4 = ((cap * 4 + 255) / 256) * 256;
    size_t const bytes = col8 * (ctx->dim + 3 + 2) + col4 * ctx->dim;
    unsigned char* base = nullptr;
    PHB_CUDA(ctx, cudaSetDevice(ctx->device));
    PHB_CUDA(ctx, cudaMalloc(&base, bytes));
    unsigned char* p = base;
    out->weight = reinterpret_cast<double*>(p); // the allocation is owned through `weight`
    p += col8;
    out->charge = reinterpret_cast<double*>(p);
    p += col8;
    for (int d = 0; d < ctx->dim; ++d, p += col8)
        out->delta[d] = reinterpret_cast<double*>(p);
    for (int c = 0; c < 3; ++c, p += col8)
        out->v[c] = reinterpret_cast<double*>(p);
    for (int d = 0; d < ctx->dim; ++d, p += col4)
        out->icell[d] = reinterpret_cast<int*>(p);
    out->n        = 0;
    out->capacity = cap;
    return PHB_OK;
}

int phb_particles_free(phb_ctx* ctx, phb_particles* P)
{
    if (!ctx || !P)
        return PHB_ERR_INVALID;
    if (P->weight)
    {
        PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PHB_CUDA(ctx, cudaFree(P->weight));
    }
    *P = phb_particles{};
    return PHB_OK;
}

int phb_particles_from_aos(phb_ctx* ctx, const void* h_aos, size_t n, phb_particles* dst)
{
    if (!ctx || !dst || (n && !h_aos))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_particles_from_aos: invalid argument");
    if (n > dst->capacity)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_particles_from_aos: capacity");
    dst->n = n;
    if (!n)
        return PHB_OK;
    size_t const bytes = n * phb::aos_stride(ctx->dim);
    if (int rc = phb::ensure_scratch(ctx, bytes))
        return rc;
    PHB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, h_aos, bytes, cudaMemcpyHostToDevice, ctx->stream));
    auto* aos           = static_cast<unsigned char*>(ctx->scratch);
    unsigned const grid = unsigned((n + 255) / 256);
    phb::PartView V     = phb::make_part(*dst);
    if (ctx->dim == 1)
        phb::aos_kernel<1, true><<<grid, 256, 0, ctx->stream>>>(aos, V, n);
    else if (ctx->dim == 2)
        phb::aos_kernel<2, true><<<grid, 256, 0, ctx->stream>>>(aos, V, n);
    else
        phb::aos_kernel<3, true><<<grid, 256, 0, ctx->stream>>>(aos, V, n);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

int phb_particles_to_aos(phb_ctx* ctx, const phb_particles* src, void* h_aos)
{
    if (!ctx || !src || (src->n && !h_aos))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_particles_to_aos: invalid argument");
    size_t const n = src->n;
    if (!n)
        return PHB_OK;
    size_t const bytes = n * phb::aos_stride(ctx->dim);
    if (int rc = phb::ensure_scratch(ctx, bytes))
        return rc;
    auto* aos           = static_cast<unsigned char*>(ctx->scratch);
    unsigned const grid = unsigned((n + 255) / 256);
    phb::PartView V     = phb::make_part(*src);
    if (ctx->dim == 1)
        phb::aos_kernel<1, false><<<grid, 256, 0, ctx->stream>>>(aos, V, n);
    else if (ctx->dim == 2)
        phb::aos_kernel<2, false><<<grid, 256, 0, ctx->stream>>>(aos, V, n);
    else
        phb::aos_kernel<3, false><<<grid, 256, 0, ctx->stream>>>(aos, V, n);
    PHB_LAUNCH_CHECK(ctx);
    PHB_CUDA(ctx, cudaMemcpyAsync(h_aos, aos, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PHB_OK;
}

int phb_particles_from_soa(phb_ctx* ctx, const int* h_icell, const double* h_delta, const double* h_weight,
                           const double* h_charge, const double* h_v, size_t n, phb_particles* dst)
{
    if (!ctx || !dst || (n && (!h_icell || !h_delta || !h_weight || !h_charge || !h_v)))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_particles_from_soa: invalid argument");
    if (n > dst->capacity)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_particles_from_soa: capacity");
    dst->n = n;
    if (!n)
        return PHB_OK;
    int const dim = ctx->dim;
    if (int rc = phb::ensure_scratch(ctx, n * 8 * 3))
        return rc;
    unsigned const grid = unsigned((n + 255) / 256);
    PHB_CUDA(ctx, cudaMemcpyAsync(dst->weight, h_weight, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    PHB_CUDA(ctx, cudaMemcpyAsync(dst->charge, h_charge, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    auto* tmp_i = static_cast<int*>(ctx->scratch);
    auto* tmp_d = static_cast<double*>(ctx->scratch);
    PHB_CUDA(ctx, cudaMemcpyAsync(tmp_i, h_icell, n * 4 * dim, cudaMemcpyHostToDevice, ctx->stream));
    phb::interleave_kernel<int, true><<<grid, 256, 0, ctx->stream>>>(tmp_i, dst->icell[0], dst->icell[1],
                                                                       dst->icell[2], dim, n);
    PHB_LAUNCH_CHECK(ctx);
    PHB_CUDA(ctx, cudaMemcpyAsync(tmp_d, h_delta, n * 8 * dim, cudaMemcpyHostToDevice, ctx->stream));
    phb::interleave_kernel<double, true><<<grid, 256, 0, ctx->stream>>>(tmp_d, dst->delta[0], dst->delta[1],
                                                                          dst->delta[2], dim, n);
    PHB_LAUNCH_CHECK(ctx);
    PHB_CUDA(ctx, cudaMemcpyAsync(tmp_d, h_v, n * 8 * 3, cudaMemcpyHostToDevice, ctx->stream));
    phb::interleave_kernel<double, true><<<grid, 256, 0, ctx->stream>>>(tmp_d, dst->v[0], dst->v[1], dst->v[2], 3, n);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

int phb_particles_to_soa(phb_ctx* ctx, const phb_particles* src, int* h_icell, double* h_delta, double* h_weight,
                         double* h_charge, double* h_v)
{
    if (!ctx || !src)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_particles_to_soa: invalid argument");
    size_t const n = src->n;
    if (!n)
        return PHB_OK;
    int const dim = ctx->dim;
    if (int rc = phb::ensure_scratch(ctx, n * 8 * 3))
        return rc;
    unsigned const grid = unsigned((n + 255) / 256);
    PHB_CUDA(ctx, cudaMemcpyAsync(h_weight, src->weight, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PHB_CUDA(ctx, cudaMemcpyAsync(h_charge, src->charge, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    auto* tmp_i = static_cast<int*>(ctx->scratch);
    auto* tmp_d = static_cast<double*>(ctx->scratch);
    phb::interleave_kernel<int, false><<<grid, 256, 0, ctx->stream>>>(tmp_i, src->icell[0], src->icell[1],
                                                                        src->icell[2], dim, n);
    PHB_LAUNCH_CHECK(ctx);
    PHB_CUDA(ctx, cudaMemcpyAsync(h_icell, tmp_i, n * 4 * dim, cudaMemcpyDeviceToHost, ctx->stream));
    phb::interleave_kernel<double, false><<<grid, 256, 0, ctx->stream>>>(tmp_d, src->delta[0], src->delta[1],
                                                                           src->delta[2], dim, n);
    PHB_LAUNCH_CHECK(ctx);
    PHB_CUDA(ctx, cudaMemcpyAsync(h_delta, tmp_d, n * 8 * dim, cudaMemcpyDeviceToHost, ctx->stream));
    phb::interleave_kernel<double, false><<<grid, 256, 0, ctx->stream>>>(tmp_d, src->v[0], src->v[1], src->v[2], 3, n);
    PHB_LAUNCH_CHECK(ctx);
    PHB_CUDA(ctx, cudaMemcpyAsync(h_v, tmp_d, n * 8 * 3, cudaMemcpyDeviceToHost, ctx->stream));
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PHB_OK;
}

int phb_particles_copy(phb_ctx* ctx, const phb_particles* src, size_t src_first, size_t count, phb_particles* dst,
                       size_t dst_first)
{
    if (!ctx || !src || !dst || src_first + count > src->capacity)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_particles_copy: invalid argument");
    if (dst_first + count > dst->capacity)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_particles_copy: capacity");
    if (count)
    {
        auto cp = [&](void* d, const void* s, size_t esz) {
            return cudaMemcpyAsync(static_cast<char*>(d) + dst_first * esz,
                                   static_cast<const char*>(s) + src_first * esz, count * esz,
                                   cudaMemcpyDeviceToDevice, ctx->stream);
        };
        for (int d = 0; d < ctx->dim; ++d)
        {
            PHB_CUDA(ctx, cp(dst->icell[d], src->icell[d], 4));
            PHB_CUDA(ctx, cp(dst->delta[d], src->delta[d], 8));
        }
        for (int c = 0; c < 3; ++c)
            PHB_CUDA(ctx, cp(dst->v[c], src->v[c], 8));
        PHB_CUDA(ctx, cp(dst->weight, src->weight, 8));
        PHB_CUDA(ctx, cp(dst->charge, src->charge, 8));
    }
    if (dst_first + count > dst->n)
        dst->n = dst_first + count;
    return PHB_OK;
}
}

// ---- flat packing of a store range for the migration messages (ParticlesData::packStream / unpackStream,
// src/amr/data/particles/particles_data.hpp:702-784): ONE launch per store instead of one copy per column.
// A message of `total` particles is column-major: delta[d][total] | v[3][total] | weight[total] | charge[total] |
// icell[d][total]; a store range of n particles occupies entries [off, off+n) of every column.
namespace phb
{
template<bool PACK>
__global__ void __launch_bounds__(256)
    particles_flat_kernel(PartView P, size_t first, size_t n, int dim, unsigned char* buf, size_t total, size_t off)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    double* b8 = reinterpret_cast<double*>(buf) + off;
    int* b4    = reinterpret_cast<int*>(buf + size_t(dim + 5) * total * 8) + off;
    int c      = 0;
    auto mv8   = [&](double* col) {
        if (PACK)
            b8[size_t(c) * total + i] = col[first + i];
        else
            col[first + i] = b8[size_t(c) * total + i];
        ++c;
    };
    for (int d = 0; d < dim; ++d)
        mv8(P.delta[d]);
    for (int k = 0; k < 3; ++k)
        mv8(P.v[k]);
    mv8(P.weight);
    mv8(P.charge);
    for (int d = 0; d < dim; ++d)
    {
        if (PACK)
            b4[size_t(d) * total + i] = P.icell[d][first + i];
        else
            P.icell[d][first + i] = b4[size_t(d) * total + i];
    }
}
} // namespace phb

extern "C" {
size_t phb_particles_flat_bytes(int dim, size_t total) { return total * size_t(8 * (dim + 5) + 4 * dim); }

int phb_particles_pack(phb_ctx* ctx, const phb_particles* src, size_t first, size_t count, void* d_buf, size_t total,
                       size_t off)
{
    if (!ctx || !src || first + count > src->capacity || (count && !d_buf) || off + count > total)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_particles_pack: invalid argument");
    if (count == 0)
        return PHB_OK;
    phb::particles_flat_kernel<true><<<unsigned((count + 255) / 256), 256, 0, ctx->stream>>>(
        phb::make_part(*src), first, count, ctx->dim, static_cast<unsigned char*>(d_buf), total, off);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

int phb_particles_unpack(phb_ctx* ctx, const void* d_buf, size_t total, size_t off, size_t count, phb_particles* dst)
{
    if (!ctx || !dst || (count && !d_buf) || off + count > total)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_particles_unpack: invalid argument");
    if (dst->n + count > dst->capacity)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_particles_unpack: capacity");
    if (count)
    {
        phb::particles_flat_kernel<false><<<unsigned((count + 255) / 256), 256, 0, ctx->stream>>>(
            phb::make_part(*dst), dst->n, count, ctx->dim,
            const_cast<unsigned char*>(static_cast<const unsigned char*>(d_buf)), total, off);
        PHB_LAUNCH_CHECK(ctx);
    }
    dst->n += count;
    return PHB_OK;
}
}

