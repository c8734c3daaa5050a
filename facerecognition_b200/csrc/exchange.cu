// exchange.cu — the cross-GPU step of a sharded search, fused into ONE kernel over NVLink peer memory.
//
// After the local search every rank holds its best k candidates per query ([Q, k] scores + global row ids).
// topk_exchange_merge_kernel does the rest without a library collective:
//   1. CTA b (128 queries) stores its part of the local record straight into slot `rank` of EVERY rank's
//      exchange buffer (peer stores through NVLink; the buffers are cudaMalloc'ed by this library and mapped
//      into each process with CUDA IPC),
//   2. fences to system scope and release-stores the step's epoch into flag[rank][b] on every rank,
//   3. acquire-spins on its own flag[p][b] for every rank p — it needs exactly the CTAs b of the other GPUs,
//      nothing else, so there is no grid-wide or device-wide barrier —
//   4. merges the R lists of its queries under the total order (key, then lowest global row).
// Buffers are double-buffered by epoch parity: a rank can only be one step ahead of a peer (it had to see the
// peer's flag of the previous step), and the peer's kernel of two steps ago has finished by then.
// The epoch lives in DEVICE memory (ctl[0] of the rank's own buffer): every CTA reads it on entry and the last CTA
// to finish advances it, so a launch carries no per-call host state and the step can be replayed from a CUDA graph.
// The spin is bounded (~10 s): a rank that never arrives is counted in ctl[2] (frb_exchange_status) and the kernel
// returns instead of hanging the GPU; after any failure the ranks call frb_exchange_reset in lockstep to restart
// the epochs from zero.
//
// Replaces: torch.distributed all_gather_into_tensor + frb_topk_merge_strided (facerecognition_b200/sharded.py),
// which stay as the portable path (gloo tests, hosts without IPC).  There is no reference counterpart: the
// reference is single-device.
#include "frb_common.cuh"

#include <new>

struct frb_exchange {
    int world, rank, device;
    int64_t max_query;
    int max_k, max_ctas;
    size_t rec_bytes;    // one rank's record for max_query x max_k: ids then scores, padded to 16
    size_t flags_off;    // byte offset of the flags: u32 [2 parities][world][max_ctas]
    size_t ctl_off;      // byte offset of the control words: u32 {epoch of the last finished step, CTAs finished, timeouts}
    size_t total_bytes;
    unsigned char *local;            // this rank's buffer (cudaMalloc)
    unsigned char *peer[FRB_EXCHANGE_MAX_WORLD];  // every rank's buffer as mapped here (peer[rank] == local)
    bool opened[FRB_EXCHANGE_MAX_WORLD];
};

namespace frb {

constexpr int kExThreads = 128;

struct ExParams {
    unsigned char *peer[FRB_EXCHANGE_MAX_WORLD];
    int world, rank;
    size_t rec_bytes, flags_off, ctl_off;
    int max_ctas;
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// record layout inside a slot, for THIS call's (n_query, k): ids i64 [n_query, k] | scores f32 [n_query, k]
__device__ __forceinline__ int64_t *slot_idx(unsigned char *buf, size_t rec_bytes, int world, unsigned parity, int r)
{
    return reinterpret_cast<int64_t *>(buf + ((size_t)parity * world + r) * rec_bytes);
}

// EMULATE: all ranks live on this device and blockIdx.y is the rank (single-launch test of the protocol).
template <bool LARGEST, bool EMULATE>
__global__ void __launch_bounds__(kExThreads) topk_exchange_merge_kernel(const float *__restrict__ ls, const int64_t *__restrict__ li,
                                                                         int64_t n_query, int k, ExParams ex,
                                                                         float *__restrict__ os, int64_t *__restrict__ oi)
{
    const int rank = EMULATE ? (int)blockIdx.y : ex.rank;
    if (EMULATE) {
        ls += (int64_t)rank * n_query * k;
        li += (int64_t)rank * n_query * k;
        os += (int64_t)rank * n_query * k;
        oi += (int64_t)rank * n_query * k;
    }
    // this step's epoch: one more than the last finished step of THIS rank (advanced below by the last CTA to finish)
    unsigned *ctl = reinterpret_cast<unsigned *>(ex.peer[rank] + ex.ctl_off);
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned *>(ctl) + 1u;
    __syncthreads();
    const unsigned epoch = s_epoch;
    const unsigned parity = epoch & 1u;
    const int64_t q = (int64_t)blockIdx.x * kExThreads + threadIdx.x;
    const bool live = q < n_query;

    // 1. my candidates -> slot `rank` of every rank's buffer
    if (live) {
        for (int p = 0; p < ex.world; p++) {
            int64_t *di = slot_idx(ex.peer[p], ex.rec_bytes, ex.world, parity, rank);
            float *ds = reinterpret_cast<float *>(di + n_query * k);
            for (int j = 0; j < k; j++) {
                di[q * k + j] = li[q * k + j];
                ds[q * k + j] = ls[q * k + j];
            }
        }
    }
    // 2. publish: everything this CTA wrote is visible system-wide before the flags are
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < ex.world) {
        unsigned *flags = reinterpret_cast<unsigned *>(ex.peer[threadIdx.x] + ex.flags_off);
        st_release_sys(flags + ((size_t)parity * ex.world + rank) * ex.max_ctas + blockIdx.x, epoch);
    }
    // 3. wait for the same CTA of every rank
    if (threadIdx.x < ex.world) {
        const unsigned *flags = reinterpret_cast<const unsigned *>(ex.peer[rank] + ex.flags_off);
        const unsigned *f = flags + ((size_t)parity * ex.world + threadIdx.x) * ex.max_ctas + blockIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != epoch) {
            if (clock64() - t0 > 20000000000LL) {          // ~10 s: a rank never arrived; report, do not hang
                atomicAdd(ctl + 2, 1u);
                break;
            }
        }
    }
    __syncthreads();
    // the last CTA of this rank to get here advances the epoch (every CTA has read it by then)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ctl + 1, 1u) == gridDim.x - 1) {
            ctl[1] = 0;
            __threadfence();
            *reinterpret_cast<volatile unsigned *>(ctl) = epoch;
        }
    }
    if (!live) return;

    // 4. merge (L1 is bypassed: these lines were written by other GPUs)
    float s[FRB_MAX_K];
    int64_t id[FRB_MAX_K];
    list_init<LARGEST>(s, id, k);
    for (int p = 0; p < ex.world; p++) {
        const int64_t *pi = slot_idx(ex.peer[rank], ex.rec_bytes, ex.world, parity, p);
        const float *ps = reinterpret_cast<const float *>(pi + n_query * k);
        for (int j = 0; j < k; j++) {
            const float v = __ldcg(ps + q * k + j);
            const int64_t idx = __ldcg(pi + q * k + j);
            if (idx < 0 || !better<LARGEST>(v, idx, s[k - 1], id[k - 1])) continue;
            int pos = k - 1;
            while (pos > 0 && better<LARGEST>(v, idx, s[pos - 1], id[pos - 1])) {
                s[pos] = s[pos - 1];
                id[pos] = id[pos - 1];
                --pos;
            }
            s[pos] = v;
            id[pos] = idx;
        }
    }
    for (int j = 0; j < k; j++) {
        os[q * k + j] = s[j];
        oi[q * k + j] = id[j];
    }
}

static int check_call(const char *fn, const frb_exchange *ex, int64_t n_query, int k)
{
    FRB_CHECK_ARG(ex, "%s: null exchange", fn);
    FRB_CHECK_ARG(n_query >= 0 && n_query <= ex->max_query && k >= 1 && k <= ex->max_k, "%s: n_query=%lld k=%d exceed the exchange's %lld x %d",
                  fn, (long long)n_query, k, (long long)ex->max_query, ex->max_k);
    return FRB_OK;
}

static void fill_params(const frb_exchange *ex, ExParams *p)
{
    for (int i = 0; i < FRB_EXCHANGE_MAX_WORLD; i++) p->peer[i] = ex->peer[i];
    p->world = ex->world;
    p->rank = ex->rank;
    p->rec_bytes = ex->rec_bytes;
    p->flags_off = ex->flags_off;
    p->ctl_off = ex->ctl_off;
    p->max_ctas = ex->max_ctas;
}

}  // namespace frb

using namespace frb;

extern "C" {

int frb_exchange_create(int world, int rank, int64_t max_query, int max_k, frb_exchange **out, unsigned char *ipc_handle_out)
{
    FRB_CHECK_ARG(out, "frb_exchange_create: null out");
    FRB_CHECK_ARG(world >= 1 && world <= FRB_EXCHANGE_MAX_WORLD && rank >= 0 && rank < world, "frb_exchange_create: world=%d rank=%d (world <= %d)",
                  world, rank, FRB_EXCHANGE_MAX_WORLD);
    FRB_CHECK_ARG(max_query >= 1 && max_k >= 1 && max_k <= FRB_MAX_K, "frb_exchange_create: max_query=%lld max_k=%d", (long long)max_query, max_k);
    frb_exchange *ex = new (std::nothrow) frb_exchange();
    FRB_CHECK_ARG(ex, "frb_exchange_create: out of host memory");
    ex->world = world;
    ex->rank = rank;
    ex->max_query = max_query;
    ex->max_k = max_k;
    ex->max_ctas = (int)((max_query + kExThreads - 1) / kExThreads);
    ex->rec_bytes = align_up((size_t)max_query * max_k * 12, 16);
    ex->flags_off = align_up(2 * (size_t)world * ex->rec_bytes, 256);
    ex->ctl_off = align_up(ex->flags_off + 2 * (size_t)world * ex->max_ctas * sizeof(unsigned), 256);
    ex->total_bytes = ex->ctl_off + 256;
    for (int i = 0; i < FRB_EXCHANGE_MAX_WORLD; i++) {
        ex->peer[i] = nullptr;
        ex->opened[i] = false;
    }
    cudaError_t e = cudaGetDevice(&ex->device);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ex->local, ex->total_bytes);
    if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->total_bytes);
    if (e == cudaSuccess && ipc_handle_out) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, ex->local);
        if (e == cudaSuccess) memcpy(ipc_handle_out, &h, sizeof(h));
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error("frb_exchange_create: %s (%zu bytes)", cudaGetErrorString(e), ex->total_bytes);
        if (ex->local) cudaFree(ex->local);
        delete ex;
        return FRB_ERR_CUDA;
    }
    ex->peer[rank] = ex->local;
    *out = ex;
    return FRB_OK;
}

int frb_exchange_open(frb_exchange *ex, const unsigned char *ipc_handles)
{
    FRB_CHECK_ARG(ex && ipc_handles, "frb_exchange_open: null pointer");
    for (int p = 0; p < ex->world; p++) {
        if (p == ex->rank || ex->peer[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_handles + (size_t)p * FRB_IPC_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("frb_exchange_open: cudaIpcOpenMemHandle for rank %d failed: %s", p, cudaGetErrorString(e));
            return FRB_ERR_CUDA;
        }
        ex->peer[p] = (unsigned char *)ptr;
        ex->opened[p] = true;
    }
    return FRB_OK;
}

int frb_exchange_destroy(frb_exchange *ex)
{
    if (!ex) return FRB_OK;
    for (int p = 0; p < ex->world; p++)
        if (ex->opened[p] && ex->peer[p]) cudaIpcCloseMemHandle(ex->peer[p]);
    if (ex->local) cudaFree(ex->local);
    delete ex;
    return FRB_OK;
}

int frb_exchange_topk_merge(frb_exchange *ex, const float *local_scores, const int64_t *local_idx, int64_t n_query, int k,
                            int largest, float *out_scores, int64_t *out_idx, void *stream)
{
    int rc = check_call("frb_exchange_topk_merge", ex, n_query, k);
    if (rc != FRB_OK) return rc;
    for (int p = 0; p < ex->world; p++) FRB_CHECK_ARG(ex->peer[p], "frb_exchange_topk_merge: rank %d's buffer is not mapped (frb_exchange_open)", p);
    if (n_query == 0) return FRB_OK;   // no launch, no epoch: every rank sees the same n_query
    FRB_CHECK_ARG(local_scores && local_idx && out_scores && out_idx, "frb_exchange_topk_merge: null pointer");
    ExParams p;
    fill_params(ex, &p);
    const unsigned grid = (unsigned)((n_query + kExThreads - 1) / kExThreads);
    if (largest)
        topk_exchange_merge_kernel<true, false><<<grid, kExThreads, 0, (cudaStream_t)stream>>>(local_scores, local_idx, n_query, k, p, out_scores, out_idx);
    else
        topk_exchange_merge_kernel<false, false><<<grid, kExThreads, 0, (cudaStream_t)stream>>>(local_scores, local_idx, n_query, k, p, out_scores, out_idx);
    FRB_LAUNCH_OK("topk_exchange_merge_kernel");
    return FRB_OK;
}

int frb_exchange_status(frb_exchange *ex, int *timeouts, unsigned *epoch)
{
    FRB_CHECK_ARG(ex, "frb_exchange_status: null exchange");
    unsigned ctl[3] = {0, 0, 0};
    FRB_CUDA_OK(cudaMemcpy(ctl, ex->local + ex->ctl_off, sizeof(ctl), cudaMemcpyDeviceToHost));   // waits for the stream work before it
    if (epoch) *epoch = ctl[0];
    if (timeouts) *timeouts = (int)ctl[2];
    return FRB_OK;
}

int frb_exchange_reset(frb_exchange *ex)
{
    FRB_CHECK_ARG(ex, "frb_exchange_reset: null exchange");
    FRB_CUDA_OK(cudaDeviceSynchronize());
    FRB_CUDA_OK(cudaMemset(ex->local + ex->flags_off, 0, ex->total_bytes - ex->flags_off));   // flags + control words
    FRB_CUDA_OK(cudaDeviceSynchronize());
    return FRB_OK;
}

int frb_exchange_emulate(frb_exchange *const *ranks, int world, const float *local_scores, const int64_t *local_idx, int64_t n_query,
                         int k, int largest, float *out_scores, int64_t *out_idx, void *stream)
{
    FRB_CHECK_ARG(ranks && world >= 1 && world <= FRB_EXCHANGE_MAX_WORLD, "frb_exchange_emulate: world=%d", world);
    for (int r = 0; r < world; r++) {
        int rc = check_call("frb_exchange_emulate", ranks[r], n_query, k);
        if (rc != FRB_OK) return rc;
        FRB_CHECK_ARG(ranks[r]->world == world && ranks[r]->rank == r, "frb_exchange_emulate: context %d is rank %d of %d", r, ranks[r]->rank,
                      ranks[r]->world);
    }
    const unsigned grid_x = (unsigned)((n_query + kExThreads - 1) / kExThreads);
    // every CTA of every emulated rank must be resident at once (they wait for one another)
    FRB_CHECK_ARG((int64_t)grid_x * world <= (int64_t)sm_count() * 8, "frb_exchange_emulate: %u x %d CTAs cannot all be resident", grid_x, world);
    ExParams p;
    fill_params(ranks[0], &p);
    for (int r = 0; r < world; r++) p.peer[r] = ranks[r]->local;  // same process: no IPC mapping needed
    if (n_query == 0) return FRB_OK;
    dim3 grid(grid_x, (unsigned)world);
    if (largest)
        topk_exchange_merge_kernel<true, true><<<grid, kExThreads, 0, (cudaStream_t)stream>>>(local_scores, local_idx, n_query, k, p, out_scores, out_idx);
    else
        topk_exchange_merge_kernel<false, true><<<grid, kExThreads, 0, (cudaStream_t)stream>>>(local_scores, local_idx, n_query, k, p, out_scores, out_idx);
    FRB_LAUNCH_OK("topk_exchange_merge_kernel");
    return FRB_OK;
}

}  // extern "C"
