// mlp_train.cu -- K7: weight gradients of the Behler-Parrinello loss through BOTH outputs of the atomic MLP
// (the energy e and the input gradient w = dE/dx that the force uses), and K4c', the backward of the force contraction.
//
// Replaces what torch.autograd does for   nnmd/nn/bpnn.py:342-366
//     e  = net(g);  dE = autograd.grad(e.sum(), g, create_graph=True);  f = -einsum("ijk,ijkl->ijl", dE, dG)
//     loss(e.sum(1), E, f, F).backward()
// i.e. a DOUBLE backward through the sigmoid MLP (atomic_nn.py:43-48).  Nothing is stored between forward and
// backward: the kernel recomputes the activations of its 32-atom tile in shared memory (the nets are tiny).
//
// Math (per atom; h_0 = x, z_l = W_l h_l + b_l, h_{l+1} = sigmoid(z_l) for l < L-1, e = z_{L-1};
//       s_l = h_{l+1}(1-h_{l+1}), r_l = de/dh_l : r_{L-1} = W_{L-1}, r_l = (r_{l+1} * s_l) W_l, w = r_0):
//   given ge = dLoss/de and u = dLoss/dw
//   rho_0 = u;  for l = 0..L-2:  gW_l += (r_{l+1} * s_l) (x) rho_l ;  v_l = W_l rho_l ;  rho_{l+1} = v_l * s_l ;
//                                S_l = v_l * r_{l+1} * s_l (1 - 2 h_{l+1})                  (through sigmoid'')
//   gW_{L-1} += rho_{L-1}
//   zeta_{L-1} = ge;  for l = L-1..0:  gW_l += zeta_l (x) h_l ;  gb_l += zeta_l ;
//                                      zeta_{l-1} = (W_l^T zeta_l) * s_{l-1} + S_{l-1}
// Sums over atoms go to global memory with one atomicAdd per weight and tile.
#include <algorithm>

#include "common.cuh"

namespace nnmd {

constexpr int MT_MAX_LAYERS = 8;
constexpr int MT_THREADS = 256;

template <typename T>
struct TrainArgs {
    int L;
    int sizes[MT_MAX_LAYERS + 1];
    int hoff[MT_MAX_LAYERS];  // h_l        (width sizes[l])        l = 0..L-1
    int roff[MT_MAX_LAYERS];  // r_l -> S_{l-1} (width sizes[l])    l = 1..L-1
    int poff[MT_MAX_LAYERS];  // rho_l -> zeta_{l-1} (width sizes[l])
    const T *W[MT_MAX_LAYERS];
    const T *b[MT_MAX_LAYERS];
    T *gW[MT_MAX_LAYERS];
    T *gb[MT_MAX_LAYERS];
    const T *x;
    const T *ge;  // (n_rows)
    const T *u;   // (n_rows, sizes[0]) or nullptr (no force term)
    int64_t n_rows;
};

template <typename T> __device__ __forceinline__ T sigm(T z) { return T(1) / (T(1) + M<T>::exp_(-z)); }

// out[n][m] = (bias[n]) + sum_k Wm[n*ldw + k*sk] * in[k][m]   for n < N, lane m   (generic small mat-vec per atom)
template <typename T, int TM, typename Epi>
__device__ __forceinline__ void tile_matvec(const T *__restrict__ Wm, int ldn, int ldk, const T *__restrict__ in, int K, int N,
                                            Epi epi) {
    constexpr int P = TM + 1;
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int m = lane % TM;
    for (int g = warp; g * 4 < N; g += nwarps) {
        const int n0 = 4 * g, n1 = min(n0 + 1, N - 1), n2 = min(n0 + 2, N - 1), n3 = min(n0 + 3, N - 1);
        T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const T v = in[k * P + m];
            a0 = M<T>::fma_(__ldg(Wm + size_t(n0) * ldn + size_t(k) * ldk), v, a0);
            a1 = M<T>::fma_(__ldg(Wm + size_t(n1) * ldn + size_t(k) * ldk), v, a1);
            a2 = M<T>::fma_(__ldg(Wm + size_t(n2) * ldn + size_t(k) * ldk), v, a2);
            a3 = M<T>::fma_(__ldg(Wm + size_t(n3) * ldn + size_t(k) * ldk), v, a3);
        }
        if (lane < TM) {
            epi(n0, m, a0);
            if (n0 + 1 < N) epi(n0 + 1, m, a1);
            if (n0 + 2 < N) epi(n0 + 2, m, a2);
            if (n0 + 3 < N) epi(n0 + 3, m, a3);
        }
    }
}

// gW[n][k] += sum_m A[n][m] * B[k][m]   (rows of the tile beyond `valid` hold zeros in A or B)
template <typename T, int TM>
__device__ __forceinline__ void tile_outer_acc(T *__restrict__ gW, const T *__restrict__ A, int N, const T *__restrict__ B, int K) {
    constexpr int P = TM + 1;
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i - n * K;
        T s = 0;
#pragma unroll 8
        for (int m = 0; m < TM; ++m) s = M<T>::fma_(A[n * P + m], B[k * P + m], s);
        atomicAdd(gW + i, s);
    }
}

template <typename T, int TM>
__global__ void __launch_bounds__(MT_THREADS) mlp_train_bwd_kernel(TrainArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    constexpr int P = TM + 1;
    const int tid = threadIdx.x, L = a.L, K0 = a.sizes[0];
    const int64_t n_tiles = (a.n_rows + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * TM;
        const int valid = int(min(int64_t(TM), a.n_rows - m0));
        // ---- stage x and u (zero rows beyond the end: they then contribute nothing to any sum)
        T *h0 = sm + a.hoff[0], *rho0 = sm + a.poff[0];
        for (int i = tid; i < TM * K0; i += blockDim.x) {
            const int m = i / K0, k = i - m * K0;
            h0[k * P + m] = m < valid ? a.x[(m0 + m) * K0 + k] : T(0);
            rho0[k * P + m] = (m < valid && a.u) ? a.u[(m0 + m) * K0 + k] : T(0);
        }
        __syncthreads();
        // ---- A: forward, h_1 .. h_{L-1}
        for (int l = 0; l + 1 < L; ++l) {
            const T *b = a.b[l];
            T *hout = sm + a.hoff[l + 1];
            tile_matvec<T, TM>(a.W[l], a.sizes[l], 1, sm + a.hoff[l], a.sizes[l], a.sizes[l + 1],
                               [&](int n, int m, T acc) { hout[n * P + m] = sigm<T>(acc + __ldg(b + n)); });
            __syncthreads();
        }
        // ---- B: r_{L-1} .. r_1   (r_l = (r_{l+1} * s_l) W_l ; the product r_{l+1}*s_l is formed on the fly)
        if (L >= 2) {
            T *rtop = sm + a.roff[L - 1];
            const T *wl = a.W[L - 1];
            for (int i = tid; i < a.sizes[L - 1] * TM; i += blockDim.x) {
                const int k = i / TM, m = i - k * TM;
                rtop[k * P + m] = __ldg(wl + k);
            }
            __syncthreads();
            for (int l = L - 2; l >= 1; --l) {
                // delta_l = r_{l+1} * s_l into the rho_{l+1} buffer (free until phase C reaches it)
                const int Kn = a.sizes[l + 1];
                T *dl = sm + a.poff[l + 1];
                const T *rn = sm + a.roff[l + 1], *hn = sm + a.hoff[l + 1];
                for (int i = tid; i < Kn * TM; i += blockDim.x) {
                    const int k = i / TM, m = i - k * TM;
                    const T h = hn[k * P + m];
                    dl[k * P + m] = rn[k * P + m] * h * (T(1) - h);
                }
                __syncthreads();
                T *rl = sm + a.roff[l];
                // r_l[k] = sum_n delta_l[n] W_l[n][k]  -> "rows" are k (stride 1), reduction over n (stride K_l)
                tile_matvec<T, TM>(a.W[l], 1, a.sizes[l], dl, Kn, a.sizes[l], [&](int k, int m, T acc) { rl[k * P + m] = acc; });
                __syncthreads();
            }
        }
        // ---- C: rho sweep, bottom-up
        for (int l = 0; l + 1 < L; ++l) {
            const int K = a.sizes[l], Kn = a.sizes[l + 1];
            T *rho = sm + a.poff[l], *rn = sm + a.roff[l + 1], *hn = sm + a.hoff[l + 1], *rho_n = sm + a.poff[l + 1];
            // delta_l = r_{l+1} * s_l, staged in the rho_{l+1} buffer for the outer product
            for (int i = tid; i < Kn * TM; i += blockDim.x) {
                const int k = i / TM, m = i - k * TM;
                const T h = hn[k * P + m];
                rho_n[k * P + m] = rn[k * P + m] * h * (T(1) - h);
            }
            __syncthreads();
            if (a.u) tile_outer_acc<T, TM>(a.gW[l], rho_n, Kn, rho, K);
            __syncthreads();
            // v_l = W_l rho_l ; rho_{l+1} = v_l * s_l ; S_l = v_l * r_{l+1} * s_l (1 - 2h) -> overwrites r_{l+1}
            tile_matvec<T, TM>(a.W[l], K, 1, rho, K, Kn, [&](int n, int m, T v) {
                const T h = hn[n * P + m], s = h * (T(1) - h);
                const T r = rn[n * P + m];
                rho_n[n * P + m] = v * s;
                rn[n * P + m] = v * r * s * (T(1) - T(2) * h);
            });
            __syncthreads();
        }
        if (a.u) {  // gW_{L-1} += sum_m rho_{L-1}
            const int K = a.sizes[L - 1];
            const T *rho = sm + a.poff[L - 1];
            for (int k = tid; k < K; k += blockDim.x) {
                T s = 0;
                for (int m = 0; m < TM; ++m) s += rho[k * P + m];
                atomicAdd(a.gW[L - 1] + k, s);
            }
        }
        __syncthreads();
        // ---- D: zeta sweep, top-down.  zeta_l (width sizes[l+1]) lives in the rho_{l+1} buffer; zeta_{L-1} is the scalar ge
        {
            const int l = L - 1, K = a.sizes[l];
            const T *h = sm + a.hoff[l];
            // gW_{L-1}[k] += sum_m ge[m] h[k][m] ; gb_{L-1} += sum_m ge[m] ; zeta_{L-2} = W_{L-1}[k] ge * s + S
            for (int k = tid; k < K; k += blockDim.x) {
                T s = 0;
                for (int m = 0; m < valid; ++m) s = M<T>::fma_(__ldg(a.ge + m0 + m), h[k * P + m], s);
                atomicAdd(a.gW[l] + k, s);
            }
            if (tid == 0) {
                T s = 0;
                for (int m = 0; m < valid; ++m) s += __ldg(a.ge + m0 + m);
                atomicAdd(a.gb[l], s);
            }
            if (l >= 1) {
                T *zl = sm + a.poff[l];  // zeta_{l-1}, width sizes[l]
                const T *S = sm + a.roff[l];
                const T *wl = a.W[l];
                for (int i = tid; i < K * TM; i += blockDim.x) {
                    const int k = i / TM, m = i - k * TM;
                    const T hh = h[k * P + m];
                    const T g = m < valid ? __ldg(a.ge + m0 + m) : T(0);
                    zl[k * P + m] = __ldg(wl + k) * g * hh * (T(1) - hh) + S[k * P + m];
                }
            }
            __syncthreads();
        }
        for (int l = L - 2; l >= 0; --l) {
            const int K = a.sizes[l], Kn = a.sizes[l + 1];
            const T *z = sm + a.poff[l + 1];  // zeta_l
            const T *h = sm + a.hoff[l];
            tile_outer_acc<T, TM>(a.gW[l], z, Kn, h, K);
            for (int n = tid; n < Kn; n += blockDim.x) {
                T s = 0;
                for (int m = 0; m < TM; ++m) s += z[n * P + m];
                atomicAdd(a.gb[l] + n, s);
            }
            if (l >= 1) {
                // zeta_{l-1}[k] = (sum_n W_l[n][k] zeta_l[n]) * s_{l-1}[k] + S_{l-1}[k]
                T *zp = sm + a.poff[l];
                const T *S = sm + a.roff[l];
                __syncthreads();
                tile_matvec<T, TM>(a.W[l], 1, K, z, Kn, K, [&](int k, int m, T acc) {
                    const T hh = h[k * P + m];
                    zp[k * P + m] = acc * hh * (T(1) - hh) + S[k * P + m];
                });
            }
            __syncthreads();
        }
    }
}

// K4c': u[a,f] = - sum_c gf[a,c] dG[a,f,c]     (backward of force = -einsum(dE, dG) w.r.t. dE)
template <typename T>
__global__ void __launch_bounds__(256) contract_bwd_kernel(const T *__restrict__ gf, const T *__restrict__ dG, int64_t n_rows,
                                                           int nf, T *__restrict__ u) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n_rows * nf) return;
    const int64_t row = i / nf;
    const T *g = dG + 3 * i;
    u[i] = -(gf[3 * row] * g[0] + gf[3 * row + 1] * g[1] + gf[3 * row + 2] * g[2]);
}

template <typename T, int TM>
static int train_bwd_run(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases,
                         const void *x, int64_t n_rows, const void *ge, const void *u, void *const *gW, void *const *gb,
                         cudaStream_t st) {
    TrainArgs<T> a;
    a.L = n_layers;
    int off = 0;
    for (int l = 0; l <= n_layers; ++l) a.sizes[l] = sizes[l];
    for (int l = 0; l < n_layers; ++l) {
        a.hoff[l] = off; off += sizes[l] * (TM + 1);
        a.poff[l] = off; off += sizes[l] * (TM + 1);
        a.roff[l] = off; if (l >= 1) off += sizes[l] * (TM + 1);
        a.W[l] = (const T *)weights[l]; a.b[l] = (const T *)biases[l];
        a.gW[l] = (T *)gW[l]; a.gb[l] = (T *)gb[l];
    }
    a.x = (const T *)x; a.ge = (const T *)ge; a.u = (const T *)u; a.n_rows = n_rows;
    const size_t smem = size_t(off) * sizeof(T);
    if (smem > 220 * 1024) return set_err(NNMD_ERR_CAPACITY, "mlp_train: layer widths need %zu B of shared memory per tile", smem);
    auto kern = mlp_train_bwd_kernel<T, TM>;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 1;
    NNMD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MT_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t n_tiles = (n_rows + TM - 1) / TM;
    const int64_t blocks = std::min<int64_t>(n_tiles, int64_t(sm_count()) * occ);
    ProfScope prof(PROF_MLP, st);
    kern<<<unsigned(blocks), MT_THREADS, smem, st>>>(a);
    NNMD_LAUNCH_CHECK("mlp_train_bwd_kernel");
    return NNMD_OK;
}

// mlp_fast.cu: register-tiled kernel for two hidden layers of <= 64 units (negative return: shape not covered)
int train_bwd_fast(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases, const void *x,
                   int64_t n_rows, const void *ge, const void *u, void *const *gW, void *const *gb, cudaStream_t st);

}  // namespace nnmd

using namespace nnmd;

extern "C" int nnmd_mlp_train_backward(int dtype, int n_layers, const int32_t *sizes, const void *const *weights,
                                       const void *const *biases, const void *x, int64_t n_rows, const void *ge,
                                       const void *u, void *const *grad_weights, void *const *grad_biases, void *stream) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "mlp_train: bad dtype %d", dtype);
    if (n_layers < 1 || n_layers > MT_MAX_LAYERS || !sizes) return set_err(NNMD_ERR_INVALID, "mlp_train: 1..%d layers supported", MT_MAX_LAYERS);
    for (int l = 0; l <= n_layers; ++l)
        if (sizes[l] < 1) return set_err(NNMD_ERR_INVALID, "mlp_train: layer size must be positive");
    if (sizes[n_layers] != 1) return set_err(NNMD_ERR_INVALID, "mlp_train: the atomic network has a single output");
    if (n_rows < 0) return set_err(NNMD_ERR_INVALID, "mlp_train: negative row count");
    if (n_rows == 0) return NNMD_OK;
    if (!weights || !biases || !x || !ge || !grad_weights || !grad_biases) return set_err(NNMD_ERR_INVALID, "mlp_train: null pointer");
    for (int l = 0; l < n_layers; ++l)
        if (!weights[l] || !biases[l] || !grad_weights[l] || !grad_biases[l]) return set_err(NNMD_ERR_INVALID, "mlp_train: null parameter pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NNMD_F32) {
        const int rc = train_bwd_fast(n_layers, sizes, weights, biases, x, n_rows, ge, u, grad_weights, grad_biases, st);
        if (rc >= 0) return rc;
    }
    if (dtype == NNMD_F32) return train_bwd_run<float, 32>(n_layers, sizes, weights, biases, x, n_rows, ge, u, grad_weights, grad_biases, st);
    return train_bwd_run<double, 16>(n_layers, sizes, weights, biases, x, n_rows, ge, u, grad_weights, grad_biases, st);
}

extern "C" int nnmd_force_contract_backward(int dtype, const void *gf, const void *dG, int64_t n_rows, int nf, void *u, void *stream) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "contract_bwd: bad dtype %d", dtype);
    if (n_rows < 0 || nf <= 0) return set_err(NNMD_ERR_INVALID, "contract_bwd: bad sizes");
    if (n_rows == 0) return NNMD_OK;
    if (!gf || !dG || !u) return set_err(NNMD_ERR_INVALID, "contract_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = n_rows * nf;
    const unsigned blocks = unsigned((n + 255) / 256);
    ProfScope prof(PROF_CONTRACT, st);
    if (dtype == NNMD_F32) contract_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float *)gf, (const float *)dG, n_rows, nf, (float *)u);
    else contract_bwd_kernel<double><<<blocks, 256, 0, st>>>((const double *)gf, (const double *)dG, n_rows, nf, (double *)u);
    NNMD_LAUNCH_CHECK("contract_bwd_kernel");
    return NNMD_OK;
}
