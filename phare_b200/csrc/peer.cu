// Peer-memory halo exchange over NVLink: the messenger's pack kernel (K8, phb_box_op_batch) writes
// the boxes a neighbour needs DIRECTLY into that neighbour's receive area — device memory of another GPU of
// the node, mapped through CUDA IPC — and a pair of tiny kernels orders the two streams:
//   phb_peer_signal : after everything already enqueued on this stream, publish `value` in a flag word that
//                     lives in the peer's memory (release, system scope)
//   phb_peer_wait   : hold this stream until each listed local flag word has reached `value` (acquire, system
//                     scope); bounded by a timeout so that a missing peer is an error, not a hang
// One exchange phase is then  pack(remote stores) -> signal | wait -> unpack  with no library collective and no
// staging copy: latency is a few kernel launches instead of a grouped NCCL send/recv.  Replaces, for the
// fixed-size field phases, the SAMRAI RefineSchedule / MPI messages behind
// HybridHybridMessengerStrategy::fill*Ghosts / fill*Borders (hybrid_hybrid_messenger_strategy.hpp:376-497).
#include "common.cuh"

#include <cstring>

namespace phb
{
constexpr int MAX_PEER_FLAGS = 32;
struct FlagList
{
    unsigned long long* p[MAX_PEER_FLAGS];
    unsigned long long v[MAX_PEER_FLAGS]; // value published to / expected in each flag word
    int n;
};

__global__ void peer_signal_kernel(const __grid_constant__ FlagList F)
{
    int const t = threadIdx.x;
    if (t >= F.n)
        return;
    __threadfence_system(); // the remote stores of the kernels before us are ordered before the flag
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(F.p[t]), "l"(F.v[t]) : "memory");
}

__global__ void peer_wait_kernel(const __grid_constant__ FlagList F, long long max_cycles, DevError* err)
{
    int const t = threadIdx.x;
    if (t >= F.n)
        return;
    long long const t0             = clock64();
    unsigned long long const value = F.v[t];
    unsigned long long seen;
    while (true)
    {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(F.p[t]) : "memory");
        if (seen >= value)
            break;
        if (clock64() - t0 > max_cycles)
        {
            if (atomicCAS(&err->code, 0, int(PHB_ERR_PEER_TIMEOUT)) == 0)
            {
                err->index = (unsigned long long)t;
                err->delta = double(seen);
                err->vel   = double(value);
            }
            break;
        }
        __nanosleep(200);
    }
}
// ---- device-side phase counters: the value a signal publishes / a wait expects lives in device memory and is advanced by
// the kernel itself (one thread per flag, stream order), so the launch arguments of a phase never change: one C call per
// phase with precomputed arguments (and a phase that could be captured in a CUDA graph)
struct AutoFlagList
{
    unsigned long long* flag[MAX_PEER_FLAGS];    // signal: word in the PEER's arena; wait: word in mine
    unsigned long long* counter[MAX_PEER_FLAGS]; // my own count of phases delivered to / received from that peer
    int n;
};

__global__ void peer_signal_auto_kernel(const __grid_constant__ AutoFlagList F)
{
    int const t = threadIdx.x;
    if (t >= F.n)
        return;
    unsigned long long const v = *F.counter[t] + 1ull;
    *F.counter[t]              = v;
    __threadfence_system(); // the remote stores of the kernels before us are ordered before the flag
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(F.flag[t]), "l"(v) : "memory");
}

__global__ void peer_wait_auto_kernel(const __grid_constant__ AutoFlagList F, long long max_cycles, DevError* err)
{
    int const t = threadIdx.x;
    if (t >= F.n)
        return;
    unsigned long long const value = *F.counter[t] + 1ull;
    *F.counter[t]                  = value;
    long long const t0             = clock64();
    unsigned long long seen;
    while (true)
    {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(F.flag[t]) : "memory");
        if (seen >= value)
            break;
        if (clock64() - t0 > max_cycles)
        {
            if (atomicCAS(&err->code, 0, int(PHB_ERR_PEER_TIMEOUT)) == 0)
            {
                err->index = (unsigned long long)t;
                err->delta = double(seen);
                err->vel   = double(value);
            }
            break;
        }
        __nanosleep(200);
    }
}
} // namespace phb

extern "C" {
int phb_peer_phase(phb_ctx* ctx, const phb_peer_phase_desc* ph)
{
    if (!ctx || !ph || ph->n_signal < 0 || ph->n_signal > phb::MAX_PEER_FLAGS || ph->n_wait < 0
        || ph->n_wait > phb::MAX_PEER_FLAGS)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_peer_phase: invalid argument");
    if (int rc = phb_box_op_batch(ctx, ph->pre, ph->n_pre, ph->total_pre))
        return rc;
    if (ph->n_signal)
    {
        phb::AutoFlagList F;
        F.n = ph->n_signal;
        for (int i = 0; i < F.n; ++i)
        {
            F.flag[i]    = reinterpret_cast<unsigned long long*>(ph->signal_flag[i]);
            F.counter[i] = reinterpret_cast<unsigned long long*>(ph->signal_counter[i]);
        }
        phb::peer_signal_auto_kernel<<<1, 32, 0, ctx->stream>>>(F);
        PHB_LAUNCH_CHECK(ctx);
    }
    if (int rc = phb_box_op_batch(ctx, ph->local, ph->n_local, ph->total_local))
        return rc;
    if (ph->n_wait)
    {
        phb::AutoFlagList F;
        F.n = ph->n_wait;
        for (int i = 0; i < F.n; ++i)
        {
            F.flag[i]    = reinterpret_cast<unsigned long long*>(ph->wait_flag[i]);
            F.counter[i] = reinterpret_cast<unsigned long long*>(ph->wait_counter[i]);
        }
        phb::peer_wait_auto_kernel<<<1, 32, 0, ctx->stream>>>(F, (long long)(ph->timeout_s * 2.0e9), ctx->d_err);
        PHB_LAUNCH_CHECK(ctx);
    }
    return phb_box_op_batch(ctx, ph->post, ph->n_post, ph->total_post);
}

int phb_ipc_export(phb_ctx* ctx, void* d_ptr, unsigned char h_handle[64])
{
    if (!ctx || !d_ptr || !h_handle)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_ipc_export: invalid argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    PHB_CUDA(ctx, cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(h_handle, &h, 64);
    return PHB_OK;
}

int phb_ipc_open(phb_ctx* ctx, const unsigned char h_handle[64], void** d_peer_ptr)
{
    if (!ctx || !h_handle || !d_peer_ptr)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_ipc_open: invalid argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, h_handle, 64);
    PHB_CUDA(ctx, cudaIpcOpenMemHandle(d_peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PHB_OK;
}

int phb_ipc_close(phb_ctx* ctx, void* d_peer_ptr)
{
    if (!ctx || !d_peer_ptr)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_ipc_close: invalid argument");
    PHB_CUDA(ctx, cudaIpcCloseMemHandle(d_peer_ptr));
    return PHB_OK;
}

int phb_peer_signal(phb_ctx* ctx, int n, uint64_t* const* h_flag_ptrs, const uint64_t* h_values)
{
    if (!ctx || n < 0 || n > phb::MAX_PEER_FLAGS || (n > 0 && (!h_flag_ptrs || !h_values)))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_peer_signal: invalid argument");
    if (n == 0)
        return PHB_OK;
    phb::FlagList F;
    F.n = n;
    for (int i = 0; i < n; ++i)
    {
        F.p[i] = reinterpret_cast<unsigned long long*>(h_flag_ptrs[i]);
        F.v[i] = h_values[i];
    }
    phb::peer_signal_kernel<<<1, 32, 0, ctx->stream>>>(F);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

int phb_peer_wait(phb_ctx* ctx, int n, uint64_t* const* h_flag_ptrs, const uint64_t* h_values, double timeout_s)
{
    if (!ctx || n < 0 || n > phb::MAX_PEER_FLAGS || (n > 0 && (!h_flag_ptrs || !h_values)))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_peer_wait: invalid argument");
    if (n == 0)
        return PHB_OK;
    phb::FlagList F;
    F.n = n;
    for (int i = 0; i < n; ++i)
    {
        F.p[i] = reinterpret_cast<unsigned long long*>(h_flag_ptrs[i]);
        F.v[i] = h_values[i];
    }
    // clock64 ticks at the SM clock; 2 GHz is an upper bound on B200 (1.965 GHz boost), so the timeout is at least
    // timeout_s.  (cudaDevAttrClockRate is NOT queried here: it is a slow, driver-serialised attribute.)
    long long const cycles = (long long)(timeout_s * 2.0e9);
    phb::peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(F, cycles, ctx->d_err);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
}
