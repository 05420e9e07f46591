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
#include <cstdlib>

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

// ---------------------------------------------------------------------------------------------------------
// v2: register-tiled kernel (fp32).  A CTA of 256 threads owns 64 atoms; every thread produces a 4 atoms x 4 neurons
// tile per layer, so one 16 B load of activations and one 16 B load of weights feed 16 FMAs.  Weights live in shared
// memory in BOTH layouts: transposed [k][n] for the forward sweep (4 neurons contiguous) and row-major [n][k] for
// the backward sweep (4 inputs contiguous).  Used when everything fits in shared memory; otherwise v1 above.
// ---------------------------------------------------------------------------------------------------------
constexpr int MLP2_TILE = 64;
constexpr int MLP2_PITCH = 68;   // activations [k][68]: 16 B aligned rows, staggered banks
constexpr int MLP2_THREADS = 256;

struct Mlp2Args {
    int L;
    int sizes[MLP_MAX_LAYERS + 1];
    int hoff[MLP_MAX_LAYERS];   // activations of layer l's input (floats)
    int wtoff[MLP_MAX_LAYERS];  // W^T  [K][Np]
    int woff[MLP_MAX_LAYERS];   // W    [N][Kp]
    int boff[MLP_MAX_LAYERS];   // bias [Np]
    const float *W[MLP_MAX_LAYERS];
    const float *b[MLP_MAX_LAYERS];
    const float *x;
    int64_t n_rows;
    float *e;
    float *dEdx;
};

__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + expf(-z)); }

__global__ void __launch_bounds__(MLP2_THREADS, 1) mlp2_kernel(Mlp2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sm = reinterpret_cast<float *>(smem_raw);
    const int tid = threadIdx.x;
    const int L = a.L, K0 = a.sizes[0];
    // ---- weights into shared memory, both layouts, zero padded to multiples of 4
    for (int l = 0; l < L; ++l) {
        const int K = a.sizes[l], N = a.sizes[l + 1], Np = (N + 3) & ~3, Kp = (K + 3) & ~3;
        float *wt = sm + a.wtoff[l], *w = sm + a.woff[l], *bb = sm + a.boff[l];
        for (int i = tid; i < K * Np; i += blockDim.x) {
            const int k = i / Np, n = i - k * Np;
            wt[i] = n < N ? __ldg(a.W[l] + size_t(n) * K + k) : 0.f;
        }
        for (int i = tid; i < N * Kp; i += blockDim.x) {
            const int n = i / Kp, k = i - n * Kp;
            w[i] = k < K ? __ldg(a.W[l] + size_t(n) * K + k) : 0.f;
        }
        for (int i = tid; i < Np; i += blockDim.x) bb[i] = i < N ? __ldg(a.b[l] + i) : 0.f;
    }
    __syncthreads();
    const int aq = tid & 15;  // atom quad inside the tile
    const int64_t n_tiles = (a.n_rows + MLP2_TILE - 1) / MLP2_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * MLP2_TILE;
        const int valid = int(min(int64_t(MLP2_TILE), a.n_rows - m0));
        float *h0 = sm + a.hoff[0];
        for (int i = tid; i < MLP2_TILE * K0; i += blockDim.x) {
            const int m = i / K0, k = i - m * K0;
            h0[k * MLP2_PITCH + m] = m < valid ? a.x[(m0 + m) * K0 + k] : 0.f;
        }
        __syncthreads();
        // ---- forward
        for (int l = 0; l < L; ++l) {
            const int K = a.sizes[l], N = a.sizes[l + 1], Np = (N + 3) & ~3;
            const float *hin = sm + a.hoff[l], *wt = sm + a.wtoff[l], *bb = sm + a.boff[l];
            for (int nq = tid >> 4; nq * 4 < Np; nq += MLP2_THREADS / 16) {
                const float4 b4 = *reinterpret_cast<const float4 *>(bb + 4 * nq);
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc[i][0] = b4.x; acc[i][1] = b4.y; acc[i][2] = b4.z; acc[i][3] = b4.w; }
                const float *hp = hin + 4 * aq, *wp = wt + 4 * nq;
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    const float4 h4 = *reinterpret_cast<const float4 *>(hp + k * MLP2_PITCH);
                    const float4 w4 = *reinterpret_cast<const float4 *>(wp + k * Np);
                    const float hv[4] = {h4.x, h4.y, h4.z, h4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(hv[i], wv[j], acc[i][j]);
                }
                if (l < L - 1) {
                    float *hout = sm + a.hoff[l + 1];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (4 * nq + j < N)
                            *reinterpret_cast<float4 *>(hout + (4 * nq + j) * MLP2_PITCH + 4 * aq) =
                                make_float4(sigmoidf_(acc[0][j]), sigmoidf_(acc[1][j]), sigmoidf_(acc[2][j]), sigmoidf_(acc[3][j]));
                } else if (nq == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (4 * aq + i < valid) a.e[m0 + 4 * aq + i] = acc[i][0];
                }
            }
            __syncthreads();
        }
        // ---- backward (dE/dx), in place as in v1
        if (a.dEdx) {
            {
                const int l = L - 1, K = a.sizes[l];
                float *buf = sm + a.hoff[l];
                const float *w = sm + a.woff[l];  // single row
                for (int i = tid; i < K * MLP2_TILE; i += blockDim.x) {
                    const int k = i >> 6, m = i & 63;
                    const float wv = w[k];
                    if (l > 0) { const float h = buf[k * MLP2_PITCH + m]; buf[k * MLP2_PITCH + m] = wv * h * (1.f - h); }
                    else buf[k * MLP2_PITCH + m] = wv;
                }
                __syncthreads();
            }
            for (int l = L - 1; l >= 1; --l) {
                const int N = a.sizes[l], Kc = a.sizes[l - 1], Kp = (Kc + 3) & ~3;  // W_{l-1} is (N x Kc)
                const float *buf = sm + a.hoff[l], *w = sm + a.woff[l - 1];
                float *prev = sm + a.hoff[l - 1];
                for (int kq = tid >> 4; kq * 4 < Kp; kq += MLP2_THREADS / 16) {
                    float acc[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
                    const float *dp = buf + 4 * aq, *wp = w + 4 * kq;
#pragma unroll 4
                    for (int n = 0; n < N; ++n) {
                        const float4 d4 = *reinterpret_cast<const float4 *>(dp + n * MLP2_PITCH);
                        const float4 w4 = *reinterpret_cast<const float4 *>(wp + n * Kp);
                        const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dv[i], wv[j], acc[i][j]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = 4 * kq + j;
                        if (k >= Kc) continue;
                        float4 *pp = reinterpret_cast<float4 *>(prev + k * MLP2_PITCH + 4 * aq);
                        if (l - 1 >= 1) {
                            const float4 h = *pp;
                            *pp = make_float4(acc[0][j] * h.x * (1.f - h.x), acc[1][j] * h.y * (1.f - h.y),
                                              acc[2][j] * h.z * (1.f - h.z), acc[3][j] * h.w * (1.f - h.w));
                        } else {
                            *pp = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                        }
                    }
                }
                __syncthreads();
            }
            for (int i = tid; i < MLP2_TILE * K0; i += blockDim.x) {
                const int m = i / K0, k = i - m * K0;
                if (m < valid) a.dEdx[(m0 + m) * K0 + k] = h0[k * MLP2_PITCH + m];
            }
        }
        __syncthreads();
    }
}

// returns NNMD_OK after launching, or -1 when the network does not fit (caller falls back to v1)
// mlp_fast.cu: tile kernel for two hidden layers of <= 64 units (negative return: shape not covered)
int mlp_fwd_fast(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases, const void *x,
                 int64_t n_rows, void *e, void *dEdx, cudaStream_t st);
constexpr int64_t MLP_FAST_MAX_ROWS = INT64_MAX;  // measured faster than mlp2_kernel at 13.5k and at 1M rows

static int mlp2_try(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases,
                    const void *x, int64_t n_rows, void *e, void *dEdx, cudaStream_t st) {
    Mlp2Args a;
    a.L = n_layers;
    int off = 0;
    for (int l = 0; l <= n_layers; ++l) a.sizes[l] = sizes[l];
    for (int l = 0; l < n_layers; ++l) {
        const int K = sizes[l], N = sizes[l + 1], Np = (N + 3) & ~3, Kp = (K + 3) & ~3;
        a.hoff[l] = off;  off += K * MLP2_PITCH;
        a.wtoff[l] = off; off += K * Np;
        a.woff[l] = off;  off += N * Kp;
        a.boff[l] = off;  off += Np;
        off = (off + 3) & ~3;
        a.W[l] = (const float *)weights[l];
        a.b[l] = (const float *)biases[l];
    }
    const size_t smem = size_t(off) * sizeof(float);
    if (smem > 220 * 1024) return -1;
    a.x = (const float *)x; a.n_rows = n_rows; a.e = (float *)e; a.dEdx = (float *)dEdx;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(mlp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const int64_t n_tiles = (n_rows + MLP2_TILE - 1) / MLP2_TILE;
    const int64_t blocks = std::min<int64_t>(n_tiles, int64_t(sm_count()));
    ProfScope prof(PROF_MLP, st);
    mlp2_kernel<<<unsigned(blocks), MLP2_THREADS, smem, st>>>(a);
    NNMD_LAUNCH_CHECK("mlp2_kernel");
    return NNMD_OK;
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
    if (dtype == NNMD_F32) {
        // first choice: the tile kernel of mlp_fast.cu (weights in shared memory, register-tiled GEMMs); NNMD_MLP_FWD=v skips it
        const char *env = tuning_env("NNMD_MLP_FWD");
        const bool fast = env ? env[0] == 'f' : n_rows <= MLP_FAST_MAX_ROWS;
        if (fast) {
            const int rf = mlp_fwd_fast(n_layers, sizes, weights, biases, x, n_rows, e, dEdx, st);
            if (rf >= 0) return rf;
        }
        const int r = mlp2_try(n_layers, sizes, weights, biases, x, n_rows, e, dEdx, st);
        if (r >= 0) return r;
        return mlp_run<float>(n_layers, sizes, weights, biases, x, n_rows, e, dEdx, st);
    }
    return mlp_run<double>(n_layers, sizes, weights, biases, x, n_rows, e, dEdx, st);
}
