// Context, memory and geometry entry points of the C ABI (include/phare_b200.h).
#include "common.cuh"

#include <cstdio>
#include <cstring>
#include <cstdlib>

namespace phb
{
int set_error(phb_ctx* ctx, int code, const std::string& msg)
{
    if (ctx)
        ctx->last_error = msg;
    return code;
}
int cuda_check(phb_ctx* ctx, cudaError_t e, const char* what)
{
    return set_error(ctx, PHB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
int ensure_scratch(phb_ctx* ctx, size_t bytes)
{
    // whoever asks for the scratch is about to overwrite it: a pending phb_bin_plan (slots in the scratch) is void
    // (phb_bin_plan sets plan_n after its own call; a tile plan keeps its slots in plan_buf)
    if (ctx->plan_kind != 2)
        ctx->plan_n = size_t(-1);
    if (bytes <= ctx->scratch_bytes)
        return 0;
    if (ctx->scratch)
    {
        PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PHB_CUDA(ctx, cudaFree(ctx->scratch));
        ctx->scratch       = nullptr;
        ctx->scratch_bytes = 0;
    }
    size_t const want = bytes + bytes / 4 + 4096;
    PHB_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return 0;
}
} // namespace phb

static thread_local std::string g_create_error;

namespace phb
{
__global__ void __launch_bounds__(128) words_to_host_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, unsigned n)
{
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
        dst[i] = src[i];
    __threadfence_system();
}
int words_to_host(phb_ctx* ctx, void* pinned_dst, const void* d_src, size_t bytes)
{
    void* dev = nullptr;
    PHB_CUDA(ctx, cudaHostGetDevicePointer(&dev, pinned_dst, 0));
    words_to_host_kernel<<<1, 128, 0, ctx->stream>>>(static_cast<uint32_t*>(dev), static_cast<const uint32_t*>(d_src),
                                                    unsigned(bytes / 4));
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
} // namespace phb

extern "C" {

const char* phb_version(void) { return "phare_b200 0.1 (sm_100a)"; }

int phb_create(int device, int dim, int interp, phb_ctx** out)
{
    if (!out || dim < 1 || dim > 3 || interp < 1 || interp > 3)
    {
        g_create_error = "phb_create: dim and interp must be in 1..3";
        return PHB_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev)
    {
        g_create_error = "phb_create: no CUDA device (there is no CPU fallback)";
        return PHB_ERR_NO_DEVICE;
    }
    auto* ctx   = new phb_ctx;
    ctx->device = device;
    ctx->dim    = dim;
    ctx->interp = interp;
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess
        || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess
        || (e = cudaMalloc(&ctx->d_err, sizeof(phb::DevError))) != cudaSuccess
        || (e = cudaMemset(ctx->d_err, 0, sizeof(phb::DevError))) != cudaSuccess
        || (e = cudaMallocHost(&ctx->h_err, sizeof(phb::DevError))) != cudaSuccess
        || (e = cudaMallocHost(&ctx->h_counts, 8 * sizeof(uint32_t))) != cudaSuccess
        || (e = cudaMallocHost(&ctx->h_bounce, phb::SMALL_D2H_BYTES)) != cudaSuccess)
    {
        g_create_error = std::string("phb_create: ") + cudaGetErrorString(e);
        delete ctx;
        return PHB_ERR_CUDA;
    }
    ctx->own_stream = true;
    if (const char* e = getenv("PHB_PREDICT_EPS"))
        ctx->predict_eps = atof(e);
    if (const char* e = getenv("PHB_NO_TMA"))
        ctx->no_tma = e[0] == '1';
    if (const char* e = getenv("PHB_NO_FUSED_CELLS"))
        ctx->no_fused_cells = e[0] == '1';
    if (const char* e = getenv("PHB_NO_TILE"))
        ctx->no_tile = e[0] == '1';
    if (const char* e = getenv("PHB_STRIP")) // the strip kernel is opt-in: measured no faster than the streaming K1
        ctx->no_strip = e[0] != '1';
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return PHB_OK;
}

void phb_destroy(phb_ctx* ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch)
        cudaFree(ctx->scratch);
    if (ctx->em_pack)
        cudaFree(ctx->em_pack);
    if (ctx->plan_buf)
        cudaFree(ctx->plan_buf);
    if (ctx->strip_counter)
        cudaFree(ctx->strip_counter);
    if (ctx->d_err)
        cudaFree(ctx->d_err);
    if (ctx->h_err)
        cudaFreeHost(ctx->h_err);
    if (ctx->h_counts)
        cudaFreeHost(ctx->h_counts);
    if (ctx->h_bounce)
        cudaFreeHost(ctx->h_bounce);
    if (ctx->own_stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* phb_last_error(phb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int phb_set_stream(phb_ctx* ctx, void* s)
{
    if (!ctx)
        return PHB_ERR_INVALID;
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream)
        cudaStreamDestroy(ctx->stream);
    ctx->stream     = static_cast<cudaStream_t>(s);
    ctx->own_stream = false;
    return PHB_OK;
}
void* phb_get_stream(phb_ctx* ctx) { return ctx ? ctx->stream : nullptr; }

int phb_sync(phb_ctx* ctx)
{
    if (!ctx)
        return PHB_ERR_INVALID;
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PHB_OK;
}

int phb_set_predict_eps(phb_ctx* ctx, double eps)
{
    if (!ctx || !(eps >= 0.) || eps > 0.5)
        return PHB_ERR_INVALID;
    ctx->predict_eps = eps;
    return PHB_OK;
}
double phb_get_predict_eps(phb_ctx* ctx) { return ctx ? ctx->predict_eps : 0.; }

int phb_set_exact(phb_ctx* ctx, int exact)
{
    if (!ctx)
        return PHB_ERR_INVALID;
    ctx->exact = exact != 0;
    return PHB_OK;
}

uint64_t phb_launch_count(phb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int phb_poll_error(phb_ctx* ctx)
{
    if (!ctx)
        return PHB_ERR_INVALID;
    static_assert(sizeof(phb::DevError) % 4 == 0, "DevError travels as words");
    if (int rc = phb::words_to_host(ctx, ctx->h_err, ctx->d_err, sizeof(phb::DevError)))
        return rc;
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int const code = ctx->h_err->code;
    if (code == 0)
        return PHB_OK;
    char buf[256];
    if (code == PHB_ERR_MOVE_TWO_CELL)
        // same text as the reference exception, boris.hpp:207-214
        snprintf(buf, sizeof buf, "Particle moved 2 cells with delta/vel: %g/%g (particle index %llu)",
                 ctx->h_err->delta, ctx->h_err->vel, ctx->h_err->index);
    else if (code == PHB_ERR_PEER_TIMEOUT)
        snprintf(buf, sizeof buf, "peer halo exchange timed out: flag %llu at %g, expected %g", ctx->h_err->index,
                 ctx->h_err->delta, ctx->h_err->vel);
    else
        snprintf(buf, sizeof buf, "Updater::outsideGhostBox (particle index %llu)", ctx->h_err->index);
    ctx->last_error = buf;
    PHB_CUDA(ctx, cudaMemsetAsync(ctx->d_err, 0, sizeof(phb::DevError), ctx->stream));
    return code;
}

int phb_malloc(phb_ctx* ctx, size_t bytes, void** d_out)
{
    if (!ctx || !d_out)
        return PHB_ERR_INVALID;
    PHB_CUDA(ctx, cudaSetDevice(ctx->device));
    PHB_CUDA(ctx, cudaMalloc(d_out, bytes ? bytes : 8));
    return PHB_OK;
}
int phb_free(phb_ctx* ctx, void* d)
{
    if (!ctx)
        return PHB_ERR_INVALID;
    if (d)
    {
        PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PHB_CUDA(ctx, cudaFree(d));
    }
    return PHB_OK;
}
int phb_memset(phb_ctx* ctx, void* d, int byte, size_t bytes)
{
    if (!ctx)
        return PHB_ERR_INVALID;
    PHB_CUDA(ctx, cudaMemsetAsync(d, byte, bytes, ctx->stream));
    return PHB_OK;
}
int phb_h2d(phb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes)
{
    if (!ctx)
        return PHB_ERR_INVALID;
    PHB_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PHB_OK;
}
int phb_d2h(phb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes)
{
    if (!ctx)
        return PHB_ERR_INVALID;
    if (bytes && bytes <= phb::SMALL_D2H_BYTES && bytes % 4 == 0 && reinterpret_cast<uintptr_t>(d_src) % 4 == 0)
    {
        // small: not through a copy engine (see words_to_host)
        if (int rc = phb::words_to_host(ctx, ctx->h_bounce, d_src, bytes))
            return rc;
        PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(h_dst, ctx->h_bounce, bytes);
        return PHB_OK;
    }
    PHB_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PHB_OK;
}
int phb_d2d(phb_ctx* ctx, void* d_dst, const void* d_src, size_t bytes)
{
    if (!ctx)
        return PHB_ERR_INVALID;
    PHB_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return PHB_OK;
}
int phb_host_alloc(size_t bytes, void** h_out)
{
    return cudaMallocHost(h_out, bytes ? bytes : 8) == cudaSuccess ? PHB_OK : PHB_ERR_CUDA;
}
int phb_host_free(void* h) { return cudaFreeHost(h) == cudaSuccess ? PHB_OK : PHB_ERR_CUDA; }

size_t phb_field_shape(const phb_layout* L, int qty, uint32_t shape[3])
{
    if (!L || qty < 0 || qty >= PHB_NQTY)
        return 0;
    phb::DevLayout D = phb::make_dev_layout(*L);
    size_t n         = 1;
    for (int d = 0; d < 3; ++d)
    {
        shape[d] = uint32_t(phb::alloc_extent(D, qty, d));
        n *= shape[d];
    }
    return n;
}
int phb_field_ghosts(int interp) { return phb::field_ghosts(interp); }
int phb_particle_ghosts(int interp) { return phb::particle_ghosts(interp); }
}
