// peer.cu -- collectives of the path written directly over NVLink peer memory (one process per GPU, CUDA IPC windows).
//
// Why not NCCL here: the two exchanges of the path are tiny and sit on the critical path of a launch-bound step --
//   * training: one all-reduce of the flat MLP gradient buffer (6k .. 30k floats) between the backward and the optimizer
//     of a 0.12 ms step (reference: the optimizer step of nnmd/nn/bpnn.py:366-368, made data-parallel);
//   * spatial sharding: <= 1 MB of halo positions per step (nnmd_b200/dist.py:SlabShard).
// A NCCL call costs a host-side enqueue and cannot (on this stack) be captured into the CUDA graph of the step, so the step
// splits into graph | collective | graph with eager glue kernels around it.  The kernels below are ordinary kernels: they
// are captured into the SAME graph as the compute they follow, read and write the peers' windows with plain loads and
// stores over NVLink, and synchronise through monotonically increasing step numbers (no reset, no host involvement).
//
// Window layout (every rank allocates the same): [ data: 2 parities x cap bytes ][ signals ].  Data is double-buffered by
// the step parity: a rank that runs ahead writes the OTHER half, and it cannot run two steps ahead because it needs every
// peer's signal of the step in between.
#include <algorithm>
#include <new>

#include "common.cuh"

namespace nnmd {
constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_MAX_BLOCKS = 64;  // blocks of the all-reduce kernel (each owns a slice and its own signal row)
constexpr long long PEER_SPIN_LIMIT = 4000000000ll;  // clock cycles (~2 s): a lost peer raises the status word, never a hang
}  // namespace nnmd

struct nnmd_peer {
    int rank = 0, world = 1;
    int64_t cap_bytes = 0;           // one parity of the data area
    char *base = nullptr;            // local window
    char *peer_base[nnmd::PEER_MAX_WORLD] = {};
    bool opened[nnmd::PEER_MAX_WORLD] = {};
    // device tables (copied once at connect)
    char **d_peer_base = nullptr;
    uint32_t *d_status = nullptr;    // sticky: 1 = a wait timed out
};

namespace nnmd {

static inline int64_t signal_offset(int64_t cap_bytes) { return 2 * cap_bytes; }
// signals: uint32 [kind][PEER_MAX_BLOCKS][PEER_MAX_WORLD] arrival words, then uint32 [kind][PEER_MAX_BLOCKS] local step counters
constexpr int PEER_KINDS = 2;  // 0 all-reduce, 1 halo
static inline int64_t signal_bytes() {
    return int64_t(PEER_KINDS) * PEER_MAX_BLOCKS * PEER_MAX_WORLD * 4 + int64_t(PEER_KINDS) * PEER_MAX_BLOCKS * 4;
}
__device__ __forceinline__ uint32_t *arrival(char *base, int64_t cap_bytes, int kind, int block) {
    return reinterpret_cast<uint32_t *>(base + 2 * cap_bytes) + (size_t(kind) * PEER_MAX_BLOCKS + block) * PEER_MAX_WORLD;
}
__device__ __forceinline__ uint32_t *counter(char *base, int64_t cap_bytes, int kind, int block) {
    return reinterpret_cast<uint32_t *>(base + 2 * cap_bytes) + size_t(PEER_KINDS) * PEER_MAX_BLOCKS * PEER_MAX_WORLD +
           size_t(kind) * PEER_MAX_BLOCKS + block;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer1(const float *p) {  // peer data: never from a stale L1 line
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// Threads 0..world-1 of the block: tell peer t that this block's slice of step `step` is in place, then wait for peer t's.
__device__ __forceinline__ void signal_and_wait(char *const *peer_base, int64_t cap_bytes, int kind, int block, int rank,
                                                int world, uint32_t step, uint32_t *status) {
    __syncthreads();  // the block's stores precede the fence of the signalling threads
    const int t = threadIdx.x;
    if (t < world) {
        // release at system scope, cumulative over the block's stores through the barrier above (no separate fence needed)
        st_release_sys(arrival(peer_base[t], cap_bytes, kind, block) + rank, step);
        const uint32_t *mine = arrival(peer_base[rank], cap_bytes, kind, block) + t;
        const long long t0 = clock64();
        // signed distance: step numbers wrap after 2^32 launches
        while (int32_t(ld_acquire_sys(mine) - step) < 0) {
            if (clock64() - t0 > PEER_SPIN_LIMIT) { atomicOr(status, 1u); break; }
        }
    }
    __syncthreads();
}

// buf[0..n): per-rank batch-mean gradient, buf[n-1] = the rank's frame count b_r (set by the caller).
// After the kernel: buf[i] = sum_r b_r buf_r[i] / sum_r b_r  (i < n-1),  buf[n-1] = sum_r b_r  -- identical on all ranks
// (fixed summation order over ranks).  A rank without a batch passes b_r = 0.  Block b owns the slice [b*per, (b+1)*per)
// and its own signal row, so no block waits for another block of its own rank; the weight b_r travels once per block
// (slot n + b of the mailbox) because a block's wait orders it only after the SAME block of the peers.
// WEIGHTED = false: plain sum of buf[0..n) over the ranks (total energy of a spatially sharded structure).
// The reads of the peers' mailboxes are the latency of the kernel (one NVLink round trip each): a thread first issues the
// loads of ALL ranks for its elements, then adds them up in rank order.
constexpr int PEER_AR_THREADS = 512;
template <bool WEIGHTED>
// weight_arg >= 0: the rank's weight comes as an argument (and is what ends up summed in buf[n-1]) instead of being read from
// buf[n-1]; loss_in / loss_acc (optional, 3 floats): loss_acc += loss_in * b_r / sum_r b_r, this rank's share of the
// global-batch loss -- both fold the two one-element kernels around the all-reduce of the training step into it.
__global__ void __launch_bounds__(PEER_AR_THREADS) peer_allreduce_kernel(char *const *peer_base, int64_t cap_bytes, int rank, int world,
                                                                         float *buf, int64_t n, int64_t per, uint32_t *status,
                                                                         float weight_arg, const float *loss_in, float *loss_acc) {
    const int b = blockIdx.x;
    __shared__ uint32_t s_step;
    __shared__ const float *s_src[PEER_MAX_WORLD];
    char *base = peer_base[rank];
    if (threadIdx.x == 0) {
        uint32_t *c = counter(base, cap_bytes, 0, b);
        s_step = *c + 1;
        *c = s_step;
    }
    __syncthreads();
    const uint32_t step = s_step;
    if (threadIdx.x < PEER_MAX_WORLD)
        s_src[threadIdx.x] = threadIdx.x < world ? reinterpret_cast<const float *>(peer_base[threadIdx.x] + (step & 1) * cap_bytes) : nullptr;
    const int64_t lo = int64_t(b) * per, hi = min(n, lo + per);
    const float bw = WEIGHTED ? (weight_arg >= 0.f ? weight_arg : buf[n - 1]) : 1.f;
    float *mine = reinterpret_cast<float *>(base + (step & 1) * cap_bytes);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) mine[i] = (WEIGHTED && i == n - 1) ? bw : buf[i] * bw;
    if (WEIGHTED && threadIdx.x == 0) mine[n + b] = bw;
    signal_and_wait(peer_base, cap_bytes, 0, b, rank, world, step, status);
    float v[PEER_MAX_WORLD], w[PEER_MAX_WORLD];
    const int64_t i0 = lo + threadIdx.x;
#pragma unroll
    for (int r = 0; r < PEER_MAX_WORLD; ++r) {
        v[r] = (r < world && i0 < hi) ? ld_peer1(s_src[r] + i0) : 0.f;
        w[r] = (WEIGHTED && r < world) ? ld_peer1(s_src[r] + n + b) : 0.f;
    }
    float inv = 1.f;
    if (WEIGHTED) {
        float wsum = 0.f;
#pragma unroll
        for (int r = 0; r < PEER_MAX_WORLD; ++r) wsum += w[r];
        inv = wsum > 0.f ? 1.f / wsum : 0.f;
        if (loss_acc != nullptr && b == 0 && threadIdx.x < 3) loss_acc[threadIdx.x] += loss_in[threadIdx.x] * bw * inv;
    }
    for (int64_t i = i0; i < hi; i += blockDim.x) {
        if (i != i0) {
#pragma unroll
            for (int r = 0; r < PEER_MAX_WORLD; ++r) v[r] = r < world ? ld_peer1(s_src[r] + i) : 0.f;
        }
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < PEER_MAX_WORLD; ++r) acc += v[r];  // fixed order: the same bits on every rank
        buf[i] = (WEIGHTED && i == n - 1) ? acc : acc * inv;
    }
}

// Halo exchange of the slab decomposition (rows of `width` 4-byte words: 3 for fp32 positions, 6 for fp64).
// Push: own[send_idx[k]] -> row dst_slot[k] of the staging area of peer dst_peer[k].  The parity is the one the
// receivers' sync kernel is about to advance to: every rank runs one exchange per step, so the step counters move in lockstep.
__global__ void __launch_bounds__(256) peer_halo_push_kernel(char *const *peer_base, int64_t cap_bytes, int rank, const uint32_t *own,
                                                             const int64_t *send_idx, const int32_t *dst_peer, const int64_t *dst_slot,
                                                             int64_t n_send, int width) {
    const uint32_t parity = (*counter(peer_base[rank], cap_bytes, 1, 0) + 1) & 1;
    for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n_send; k += int64_t(gridDim.x) * blockDim.x) {
        const uint32_t *src = own + send_idx[k] * width;
        uint32_t *dst = reinterpret_cast<uint32_t *>(peer_base[dst_peer[k]] + parity * cap_bytes) + dst_slot[k] * width;
        for (int c = 0; c < width; ++c) dst[c] = src[c];
    }
}
// One block: publish "my pushes of this step are done" to every peer and wait for theirs (kernel boundaries order the
// pushes before this kernel and the unpack after it).  The step number lives in the window: graph replays advance it.
__global__ void peer_halo_sync_kernel(char *const *peer_base, int64_t cap_bytes, int rank, int world, uint32_t *status) {
    __shared__ uint32_t s_step;
    if (threadIdx.x == 0) {
        uint32_t *c = counter(peer_base[rank], cap_bytes, 1, 0);
        s_step = *c + 1;
        *c = s_step;
    }
    __syncthreads();
    signal_and_wait(peer_base, cap_bytes, 1, 0, rank, world, s_step, status);
}
// local[0..n_own) = own, local[n_own..n_own+n_halo) = the staging area of this step's parity
__global__ void __launch_bounds__(256) peer_halo_unpack_kernel(char *const *peer_base, int64_t cap_bytes, int rank, const uint32_t *own,
                                                               int64_t n_own, int64_t n_halo, int width, uint32_t *local) {
    char *base = peer_base[rank];
    const uint32_t step = *counter(base, cap_bytes, 1, 0);  // advanced by the sync kernel of this step
    const uint32_t *stage = reinterpret_cast<const uint32_t *>(base + (step & 1) * cap_bytes);
    const int64_t n0 = n_own * width, n1 = n_halo * width;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n0 + n1; i += int64_t(gridDim.x) * blockDim.x)
        local[i] = i < n0 ? own[i] : __ldcv(stage + (i - n0));
}

}  // namespace nnmd

using namespace nnmd;

extern "C" int nnmd_peer_handle_bytes(void) { return int(sizeof(cudaIpcMemHandle_t)); }

extern "C" int nnmd_peer_create(int rank, int world, int64_t cap_bytes, nnmd_peer **out) {
    if (!out || world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || cap_bytes < 16)
        return set_err(NNMD_ERR_INVALID, "peer_create: bad arguments (world <= %d)", PEER_MAX_WORLD);
    nnmd_peer *p = new (std::nothrow) nnmd_peer();
    if (!p) return set_err(NNMD_ERR_INVALID, "peer_create: out of host memory");
    p->rank = rank; p->world = world;
    p->cap_bytes = (cap_bytes + 255) & ~int64_t(255);
    const int64_t total = signal_offset(p->cap_bytes) + signal_bytes();
    cudaError_t e = cudaMalloc((void **)&p->base, size_t(total));
    if (e == cudaSuccess) e = cudaMemset(p->base, 0, size_t(total));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_peer_base, sizeof(char *) * PEER_MAX_WORLD);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_status, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(p->d_status, 0, sizeof(uint32_t));
    if (e != cudaSuccess) {
        cudaFree(p->base); cudaFree(p->d_peer_base); cudaFree(p->d_status);
        delete p;
        return set_err(NNMD_ERR_CUDA, "peer_create: %s", cudaGetErrorString(e));
    }
    p->peer_base[rank] = p->base;
    *out = p;
    return NNMD_OK;
}

extern "C" int nnmd_peer_export(nnmd_peer *p, void *handle_out) {
    if (!p || !handle_out) return set_err(NNMD_ERR_INVALID, "peer_export: null argument");
    cudaIpcMemHandle_t h;
    NNMD_CUDA_CHECK(cudaIpcGetMemHandle(&h, p->base));
    memcpy(handle_out, &h, sizeof(h));
    return NNMD_OK;
}

// handles: world x nnmd_peer_handle_bytes() bytes in rank order (the local entry is ignored)
extern "C" int nnmd_peer_connect(nnmd_peer *p, const void *handles) {
    if (!p || !handles) return set_err(NNMD_ERR_INVALID, "peer_connect: null argument");
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + size_t(r) * sizeof(h), sizeof(h));
        void *ptr = nullptr;
        NNMD_CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        p->peer_base[r] = (char *)ptr;
        p->opened[r] = true;
    }
    NNMD_CUDA_CHECK(cudaMemcpy(p->d_peer_base, p->peer_base, sizeof(char *) * PEER_MAX_WORLD, cudaMemcpyHostToDevice));
    return NNMD_OK;
}

extern "C" void nnmd_peer_destroy(nnmd_peer *p) {
    if (!p) return;
    for (int r = 0; r < p->world; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->peer_base[r]);
    cudaFree(p->base); cudaFree(p->d_peer_base); cudaFree(p->d_status);
    delete p;
}

extern "C" int64_t nnmd_peer_capacity(const nnmd_peer *p) { return p ? p->cap_bytes : 0; }

// sticky status of the window: 0 fine, 1 a peer did not arrive within the spin limit (the results of that call are garbage)
extern "C" int nnmd_peer_status(nnmd_peer *p, int *status) {
    if (!p || !status) return set_err(NNMD_ERR_INVALID, "peer_status: null argument");
    uint32_t v = 0;
    NNMD_CUDA_CHECK(cudaMemcpy(&v, p->d_status, sizeof(v), cudaMemcpyDeviceToHost));
    *status = int(v);
    return NNMD_OK;
}

static int peer_allreduce(nnmd_peer *p, float *buf, int64_t n, bool weighted, void *stream, float weight = -1.f,
                          const float *loss_in = nullptr, float *loss_acc = nullptr) {
    if (!p || !buf || n < (weighted ? 2 : 1)) return set_err(NNMD_ERR_INVALID, "peer_allreduce: bad arguments");
    // mailbox: payload n floats, then the weight once per block (see the kernel)
    int blocks = int(std::min<int64_t>(PEER_MAX_BLOCKS, (n + PEER_AR_THREADS - 1) / PEER_AR_THREADS));
    if (blocks < 1) blocks = 1;
    if ((n + PEER_MAX_BLOCKS) * int64_t(sizeof(float)) > p->cap_bytes)
        return set_err(NNMD_ERR_CAPACITY, "peer_allreduce: %lld floats exceed the window (%lld bytes per parity)", (long long)n, (long long)p->cap_bytes);
    const int64_t per = (n + blocks - 1) / blocks;
    cudaStream_t st = (cudaStream_t)stream;
    if (weighted) peer_allreduce_kernel<true><<<blocks, PEER_AR_THREADS, 0, st>>>(p->d_peer_base, p->cap_bytes, p->rank, p->world, buf, n, per, p->d_status, weight, loss_in, loss_acc);
    else peer_allreduce_kernel<false><<<blocks, PEER_AR_THREADS, 0, st>>>(p->d_peer_base, p->cap_bytes, p->rank, p->world, buf, n, per, p->d_status, -1.f, nullptr, nullptr);
    NNMD_LAUNCH_CHECK("peer_allreduce_kernel");
    return NNMD_OK;
}
extern "C" int nnmd_peer_allreduce_weighted_f32(nnmd_peer *p, float *buf, int64_t n, void *stream) { return peer_allreduce(p, buf, n, true, stream); }
// the all-reduce of one training step: weight = frames of this rank's batch (>= 0); loss_in (3 floats: this rank's batch-mean
// losses) is added to loss_acc with the rank's share of the global batch; both may be NULL
extern "C" int nnmd_peer_allreduce_step_f32(nnmd_peer *p, float *buf, int64_t n, float weight, const float *loss_in, float *loss_acc,
                                            void *stream) {
    if (weight < 0.f || (loss_in == nullptr) != (loss_acc == nullptr)) return set_err(NNMD_ERR_INVALID, "peer_allreduce_step: bad arguments");
    return peer_allreduce(p, buf, n, true, stream, weight, loss_in, loss_acc);
}
extern "C" int nnmd_peer_allreduce_sum_f32(nnmd_peer *p, float *buf, int64_t n, void *stream) { return peer_allreduce(p, buf, n, false, stream); }

// One halo exchange: push the n_send rows named by send_idx into the peers' staging areas, synchronise, then assemble
// local = [own (n_own rows) | halo (n_halo rows, in the order fixed by dst_slot on the senders)].  Rows are `width`
// 4-byte words.  Three kernels on `stream`, capturable into a CUDA graph; every rank must call it once per step.
extern "C" int nnmd_peer_halo_exchange(nnmd_peer *p, const void *own, int64_t n_own, const int64_t *send_idx, const int32_t *dst_peer,
                                       const int64_t *dst_slot, int64_t n_send, int64_t n_halo, int width, void *local, void *stream) {
    if (!p || !own || !local || width < 1 || n_own < 0 || n_send < 0 || n_halo < 0)
        return set_err(NNMD_ERR_INVALID, "peer_halo_exchange: bad arguments");
    if (n_halo * width * int64_t(sizeof(uint32_t)) > p->cap_bytes)
        return set_err(NNMD_ERR_CAPACITY, "peer_halo_exchange: %lld halo rows exceed the window", (long long)n_halo);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_send > 0) {
        const int blocks = int(std::min<int64_t>((n_send + 255) / 256, int64_t(sm_count()) * 4));
        peer_halo_push_kernel<<<blocks, 256, 0, st>>>(p->d_peer_base, p->cap_bytes, p->rank, (const uint32_t *)own, send_idx, dst_peer,
                                                      dst_slot, n_send, width);
        NNMD_LAUNCH_CHECK("peer_halo_push_kernel");
    }
    peer_halo_sync_kernel<<<1, 32, 0, st>>>(p->d_peer_base, p->cap_bytes, p->rank, p->world, p->d_status);
    NNMD_LAUNCH_CHECK("peer_halo_sync_kernel");
    const int64_t words = (n_own + n_halo) * width;
    if (words > 0) {
        const int blocks = int(std::min<int64_t>((words + 255) / 256, int64_t(sm_count()) * 8));
        peer_halo_unpack_kernel<<<blocks, 256, 0, st>>>(p->d_peer_base, p->cap_bytes, p->rank, (const uint32_t *)own, n_own, n_halo, width,
                                                        (uint32_t *)local);
        NNMD_LAUNCH_CHECK("peer_halo_unpack_kernel");
    }
    return NNMD_OK;
}
