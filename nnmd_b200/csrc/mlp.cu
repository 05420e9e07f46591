// mlp.cu -- K6 atomic MLP: per-atom energy e = MLP(x) and dE/dx in one pass (sm_100a).
//
// Replaces  nnmd/nn/atomic_nn.py:43-48 (Linear -> sigmoid -> ... -> Linear, no activation on the last layer)
// and the torch.autograd.grad(e.sum(), g) of nnmd/nn/bpnn.py:347-349 / :537-538.
//
// One CTA owns a tile of 32 atoms (lane = atom) and keeps every layer's input activations of the tile in
// shared memory, feature-major with a 33-real pitch (conflict-free for lane = atom).  The backward sweep
// reuses those buffers in place: delta_{l-1} overwrites h_l, dE/dx overwrites the staged x.  Each thread
// produces 4 neurons of one atom per step, so a weight fetched once (warp-uniform, L1-resident) feeds 32 atoms.
// NNMD_MLP_EXACT: fp32 / fp64 FMA arithmetic (parity path).  The tensor-core (TF32) variant lives in mlp_tc.cu.
#include <algorithm>

#include "common.cuh"

namespace nnmd {

constexpr int MLP_MAX_LAYERS = 8;
constexpr int MLP_TILE = 32;
constexpr int MLP_PITCH = 33;
constexpr int MLP_THREADS = 256;

template <typename T>
struct MlpArgs {
    int L;
    int sizes[MLP_MAX_LAYERS + 1];
    int hoff[MLP_MAX_LAYERS];  // offset (in reals) of layer l's input buffer
    const T *W[MLP_MAX_LAYERS];
    const T *b[MLP_MAX_LAYERS];
    const T *x;
    int64_t n_rows;
    T *e;
    T *dEdx;
};

template <typename T> __device__ __forceinline__ T sigmoid_(T z) { return T(1) / (T(1) + M<T>::exp_(-z)); }

template <typename T>
__global__ void __launch_bounds__(MLP_THREADS) mlp_kernel(MlpArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, tid = threadIdx.x;
    const int L = a.L;
    const int K0 = a.sizes[0];
    const int64_t n_tiles = (a.n_rows + MLP_TILE - 1) / MLP_TILE;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * MLP_TILE;
        const int valid = int(min(int64_t(MLP_TILE), a.n_rows - m0));
        // stage x (row-major, coalesced) as h_0[k][m]
        T *h0 = sm + a.hoff[0];
        for (int i = tid; i < MLP_TILE * K0; i += blockDim.x) {
            const int m = i / K0, k = i - m * K0;
            h0[k * MLP_PITCH + m] = m < valid ? a.x[(m0 + m) * K0 + k] : T(0);
        }
        __syncthreads();
        // ---- forward
        for (int l = 0; l < L; ++l) {
            const int K = a.sizes[l], N = a.sizes[l + 1];
            const T *hin = sm + a.hoff[l];
            const T *W = a.W[l], *b = a.b[l];
            for (int g = warp; g * 4 < N; g += nwarps) {
                const int n0 = 4 * g;
                const int n1 = min(n0 + 1, N - 1), n2 = min(n0 + 2, N - 1), n3 = min(n0 + 3, N - 1);
                T a0 = __ldg(b + n0), a1 = __ldg(b + n1), a2 = __ldg(b + n2), a3 = __ldg(b + n3);
                const T *w0 = W + size_t(n0) * K, *w1 = W + size_t(n1) * K, *w2 = W + size_t(n2) * K, *w3 = W + size_t(n3) * K;
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    const T hv = hin[k * MLP_PITCH + lane];
                    a0 = M<T>::fma_(__ldg(w0 + k), hv, a0);
                    a1 = M<T>::fma_(__ldg(w1 + k), hv, a1);
                    a2 = M<T>::fma_(__ldg(w2 + k), hv, a2);
                    a3 = M<T>::fma_(__ldg(w3 + k), hv, a3);
                }
                if (l < L - 1) {
                    T *hout = sm + a.hoff[l + 1];
                    hout[n0 * MLP_PITCH + lane] = sigmoid_(a0);
                    if (n0 + 1 < N) hout[(n0 + 1) * MLP_PITCH + lane] = sigmoid_(a1);
                    if (n0 + 2 < N) hout[(n0 + 2) * MLP_PITCH + lane] = sigmoid_(a2);
                    if (n0 + 3 < N) hout[(n0 + 3) * MLP_PITCH + lane] = sigmoid_(a3);
                } else if (n0 == 0 && lane < valid) {
                    a.e[m0 + lane] = a0;  // output size 1 (bpnn.py:128-133)
                }
            }
            __syncthreads();
        }
        // ---- backward: dE/dx, in place
        if (a.dEdx) {
            {   // gradient w.r.t. the input of the last layer is its single weight row
                const int l = L - 1;
                const int K = a.sizes[l];
                T *buf = sm + a.hoff[l];
                const T *W = a.W[l];
                for (int i = tid; i < K * MLP_TILE; i += blockDim.x) {
                    const int k = i >> 5, m = i & 31;
                    const T wv = __ldg(W + k);
                    if (l > 0) {
                        const T h = buf[k * MLP_PITCH + m];
                        buf[k * MLP_PITCH + m] = wv * h * (T(1) - h);  // delta of layer l-1 (sigmoid' = h(1-h))
                    } else {
                        buf[k * MLP_PITCH + m] = wv;
                    }
                }
                __syncthreads();
            }
            for (int l = L - 1; l >= 1; --l) {
                // buf holds delta_{l-1} (width sizes[l]); propagate through W_{l-1} (sizes[l] x sizes[l-1])
                const int K = a.sizes[l], Kp = a.sizes[l - 1];
                const T *buf = sm + a.hoff[l];
                T *prev = sm + a.hoff[l - 1];
                const T *W = a.W[l - 1];
                for (int g = warp; g * 4 < Kp; g += nwarps) {
                    const int c0 = 4 * g;
                    const int c1 = min(c0 + 1, Kp - 1), c2 = min(c0 + 2, Kp - 1), c3 = min(c0 + 3, Kp - 1);
                    T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
                    for (int k = 0; k < K; ++k) {
                        const T dv = buf[k * MLP_PITCH + lane];
                        const T *wr = W + size_t(k) * Kp;
                        a0 = M<T>::fma_(__ldg(wr + c0), dv, a0);
                        a1 = M<T>::fma_(__ldg(wr + c1), dv, a1);
                        a2 = M<T>::fma_(__ldg(wr + c2), dv, a2);
                        a3 = M<T>::fma_(__ldg(wr + c3), dv, a3);
                    }
                    if (l - 1 >= 1) {
                        T h;
                        h = prev[c0 * MLP_PITCH + lane]; prev[c0 * MLP_PITCH + lane] = a0 * h * (T(1) - h);
                        if (c0 + 1 < Kp) { h = prev[(c0 + 1) * MLP_PITCH + lane]; prev[(c0 + 1) * MLP_PITCH + lane] = a1 * h * (T(1) - h); }
                        if (c0 + 2 < Kp) { h = prev[(c0 + 2) * MLP_PITCH + lane]; prev[(c0 + 2) * MLP_PITCH + lane] = a2 * h * (T(1) - h); }
                        if (c0 + 3 < Kp) { h = prev[(c0 + 3) * MLP_PITCH + lane]; prev[(c0 + 3) * MLP_PITCH + lane] = a3 * h * (T(1) - h); }
                    } else {
                        prev[c0 * MLP_PITCH + lane] = a0;
                        if (c0 + 1 < Kp) prev[(c0 + 1) * MLP_PITCH + lane] = a1;
                        if (c0 + 2 < Kp) prev[(c0 + 2) * MLP_PITCH + lane] = a2;
                        if (c0 + 3 < Kp) prev[(c0 + 3) * MLP_PITCH + lane] = a3;
                    }
                }
                __syncthreads();
            }
            for (int i = tid; i < MLP_TILE * K0; i += blockDim.x) {
                const int m = i / K0, k = i - m * K0;
                if (m < valid) a.dEdx[(m0 + m) * K0 + k] = h0[k * MLP_PITCH + m];
            }
        }
        __syncthreads();
    }
}

template <typename T>
static int mlp_run(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases,
                   const void *x, int64_t n_rows, void *e, void *dEdx, cudaStream_t st) {
    MlpArgs<T> a;
    a.L = n_layers;
    int off = 0;
    for (int l = 0; l <= n_layers; ++l) a.sizes[l] = sizes[l];
    for (int l = 0; l < n_layers; ++l) {
        a.hoff[l] = off;
        off += sizes[l] * MLP_PITCH;
        a.W[l] = (const T *)weights[l];
        a.b[l] = (const T *)biases[l];
    }
    a.x = (const T *)x; a.n_rows = n_rows; a.e = (T *)e; a.dEdx = (T *)dEdx;
    const size_t smem = size_t(off) * sizeof(T);
    if (smem > 220 * 1024) return set_err(NNMD_ERR_CAPACITY, "mlp: layer widths need %zu B of shared memory per tile", smem);
    auto kern = mlp_kernel<T>;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 1;
    NNMD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MLP_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t n_tiles = (n_rows + MLP_TILE - 1) / MLP_TILE;
    const int64_t blocks = std::min<int64_t>(n_tiles, int64_t(sm_count()) * occ);
    ProfScope prof(PROF_MLP, st);
    kern<<<unsigned(blocks), MLP_THREADS, smem, st>>>(a);
    NNMD_LAUNCH_CHECK("mlp_kernel");
    return NNMD_OK;
}

int mlp_tf32_run(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases,
                 const void *x, int64_t n_rows, void *e, void *dEdx, cudaStream_t st);

}  // namespace nnmd

using namespace nnmd;

extern "C" int nnmd_mlp_energy_grad(int dtype, int precision, int n_layers, const int32_t *sizes,
                                    const void *const *weights, const void *const *biases, const void *x,
                                    int64_t n_rows, void *e, void *dEdx, void *stream) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "mlp: bad dtype %d", dtype);
    if (n_layers < 1 || n_layers > MLP_MAX_LAYERS || !sizes) return set_err(NNMD_ERR_INVALID, "mlp: 1..%d layers supported", MLP_MAX_LAYERS);
    for (int l = 0; l <= n_layers; ++l)
        if (sizes[l] < 1) return set_err(NNMD_ERR_INVALID, "mlp: layer size must be positive");
    if (sizes[n_layers] != 1) return set_err(NNMD_ERR_INVALID, "mlp: the atomic network has a single output (got %d)", sizes[n_layers]);
    if (n_rows < 0) return set_err(NNMD_ERR_INVALID, "mlp: negative row count");
    if (n_rows == 0) return NNMD_OK;
    if (!weights || !biases || !x || !e) return set_err(NNMD_ERR_INVALID, "mlp: null pointer");
    for (int l = 0; l < n_layers; ++l)
        if (!weights[l] || !biases[l]) return set_err(NNMD_ERR_INVALID, "mlp: null parameter pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == NNMD_MLP_TF32) {
        if (dtype != NNMD_F32) return set_err(NNMD_ERR_INVALID, "mlp: TF32 tensor-core path needs fp32 data");
        return mlp_tf32_run(n_layers, sizes, weights, biases, x, n_rows, e, dEdx, st);
    }
    if (precision != NNMD_MLP_EXACT) return set_err(NNMD_ERR_INVALID, "mlp: bad precision %d", precision);
    if (dtype == NNMD_F32) return mlp_run<float>(n_layers, sizes, weights, biases, x, n_rows, e, dEdx, st);
    return mlp_run<double>(n_layers, sizes, weights, biases, x, n_rows, e, dEdx, st);
}
