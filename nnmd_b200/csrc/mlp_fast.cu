// mlp_fast.cu -- K7 for the shipped shapes of the atomic network: two hidden layers of at most 64 units, at most 128
// descriptors, float32 (LJ 3-25-25-1, Li 14-64-64-1, binary 32-64-64-1, 1M-atom table 96-64-64-1).
//
// Same mathematics as the generic kernel in mlp_train.cu (see the derivation in its header; reference:
// nnmd/nn/bpnn.py:342-366, atomic_nn.py:43-48), organised as register-tiled shared-memory GEMMs:
//   * every per-atom quantity of a tile of TM atoms lives in shared memory as [unit][atom] with a row pitch of TM + 4
//     floats, so that "atoms x units" products read one 16 B vector of atoms and one 16 B vector of weights per
//     reduction step for 16 (TM = 64) FMAs, and the weight-gradient outer products (reduction over the atoms of the
//     tile) read eight 16 B vectors for 64 FMAs;
//   * the weights are staged once per block (W0^T, W1^T and W1, zero-padded to 64 units; the input axis is padded
//     to a multiple of 16 only, which lets two blocks share an SM for <= 32 descriptors), the weight gradients are
//     kept in registers across all the tiles of a block and leave with ONE atomicAdd per weight and block;
//   * padded units carry zero weights, which makes every padded contribution vanish (sigmoid(0) only ever meets a
//     zero weight), and atoms past the end of the last tile carry x = u = ge = 0 and contribute nothing.
// The generic kernel remains the path for every other shape and for float64.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace nnmd {

#ifndef NNMD_FT_THREADS
#define NNMD_FT_THREADS 256
#endif
constexpr int FT_THREADS = NNMD_FT_THREADS;  // 256 or 512
constexpr int FT_TNS = FT_THREADS / 16;       // groups of 16 threads (one group = all atoms of a few units)
constexpr int FT_UN = 64 / FT_TNS;            // units (GEMM) / rows (outer product) per thread: 4 or 2
constexpr int FT_MINB = 2;                    // two blocks per SM whenever shared memory allows (caps registers at 128 / 64)
constexpr int FT_H = 64;  // padded hidden width

struct FastTrainArgs {
    int K0, H1, H2;
    int K0R;  // input rows held in shared memory: K0 rounded up to a multiple of 16
    const float *W0, *b0, *W1, *b1, *W2;
    float *gW0, *gb0, *gW1, *gb1, *gW2, *gb2;
    const float *x, *ge, *u;
    int64_t n_rows;
};

__device__ __forceinline__ float sigm_f(float z) { return 1.f / (1.f + M<float>::exp_(-z)); }

template <int MM> struct AtomVec;
template <> struct AtomVec<4> {
    static __device__ __forceinline__ void ld(const float *p, float (&v)[4]) {
        const float4 q = *reinterpret_cast<const float4 *>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void st(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct AtomVec<2> {
    static __device__ __forceinline__ void ld(const float *p, float (&v)[2]) {
        const float2 q = *reinterpret_cast<const float2 *>(p);
        v[0] = q.x; v[1] = q.y;
    }
    static __device__ __forceinline__ void st(float *p, const float (&v)[2]) { *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]); }
};

// C[n][m] = sum_k A[k][m] * Bt[k][n]   for the 64 units n and the TM atoms m of the tile.
// Thread (tm, tn): atoms MM*tm .. MM*tm+MM-1, units FT_UN*tn .. FT_UN*tn+FT_UN-1.  epi(n, m0, v[MM]) receives one unit at a time.
// Threads whose units lie at or beyond `ncols` do nothing (the function contains no barrier).
template <int TM, typename Epi>
__device__ __forceinline__ void gemm_atoms(const float *__restrict__ A, int K, const float *__restrict__ Bt, int ldb, Epi epi,
                                           int ncols = FT_H) {
    constexpr int P = TM + 4, MM = TM / 16;
    const int tm = threadIdx.x & 15, tn = threadIdx.x >> 4;
    if (FT_UN * tn >= ncols) return;
    float acc[FT_UN][MM];
#pragma unroll
    for (int j = 0; j < FT_UN; ++j)
#pragma unroll
        for (int i = 0; i < MM; ++i) acc[j][i] = 0.f;
    const float *a = A + MM * tm, *b = Bt + FT_UN * tn;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float av[MM], bv[FT_UN];
        AtomVec<MM>::ld(a + k * P, av);
        AtomVec<FT_UN>::ld(b + size_t(k) * ldb, bv);
#pragma unroll
        for (int j = 0; j < FT_UN; ++j)
#pragma unroll
            for (int i = 0; i < MM; ++i) acc[j][i] = fmaf(av[i], bv[j], acc[j][i]);
    }
#pragma unroll
    for (int j = 0; j < FT_UN; ++j) epi(FT_UN * tn + j, MM * tm, acc[j]);
}

// acc[i][j] += sum_m A[tn + FT_TNS i][m] * B[kbase + tk + 16 j][m]   for j < nj  (strided rows: conflict-free 16 B loads)
template <int TM>
__device__ __forceinline__ void outer_acc(const float *__restrict__ A, const float *__restrict__ B, int kbase, int nj,
                                          float (&acc)[FT_UN][4]) {
    constexpr int P = TM + 4;
    const int tk = threadIdx.x & 15, tn = threadIdx.x >> 4;
    const float *a = A + tn * P, *b = B + (kbase + tk) * P;
#pragma unroll 2
    for (int m = 0; m < TM; m += 4) {
        float4 av[FT_UN], bv[4];
#pragma unroll
        for (int i = 0; i < FT_UN; ++i) av[i] = *reinterpret_cast<const float4 *>(a + FT_TNS * i * P + m);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            bv[j] = j < nj ? *reinterpret_cast<const float4 *>(b + 16 * j * P + m) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= nj) continue;  // block-uniform
#pragma unroll
            for (int i = 0; i < FT_UN; ++i) {
                float s = acc[i][j];
                s = fmaf(av[i].x, bv[j].x, s); s = fmaf(av[i].y, bv[j].y, s);
                s = fmaf(av[i].z, bv[j].z, s); s = fmaf(av[i].w, bv[j].w, s);
                acc[i][j] = s;
            }
        }
    }
}

// sum over the TM atoms of a unit: the 16 lanes that share tn hold them (MM each)
template <int MM>
__device__ __forceinline__ float unit_sum(const float (&v)[MM]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MM; ++i) s += v[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

template <int TM, int NKG>
__global__ void __launch_bounds__(FT_THREADS, FT_MINB) mlp_train_fast_kernel(FastTrainArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int P = TM + 4, MM = TM / 16;
    const int K0R = a.K0R;
    float *sm = reinterpret_cast<float *>(smem_raw);
    // weights
    float *W0T = sm;                   // [K0R][64]   W0T[k][n] = W0[n][k]
    float *W1T = W0T + K0R * FT_H;     // [64][64]    W1T[k][n] = W1[n][k]
    float *W1N = W1T + FT_H * FT_H;    // [64][64]    W1N[n][k] = W1[n][k]
    float *W2s = W1N + FT_H * FT_H;    // [64]
    float *b0s = W2s + FT_H, *b1s = b0s + FT_H;
    float *gb0s = b1s + FT_H, *gb1s = gb0s + FT_H, *gw2s = gb1s + FT_H;  // block-level sums
    float *ges = gw2s + FT_H;          // [TM] dL/de of the tile's atoms
    // per-atom quantities of the tile, [unit][atom]
    float *X = ges + TM;               // [K0R]
    float *U = X + K0R * P;            // [K0R]  rho_0
    float *H1 = U + K0R * P;           // [64]
    float *R1 = H1 + FT_H * P;         // [64]   r_1, then S_0
    float *D0 = R1 + FT_H * P;         // [64]   delta_0, then zeta_0
    float *RHO1 = D0 + FT_H * P;       // [64]
    float *H2 = RHO1 + FT_H * P;       // [64]
    float *D1 = H2 + FT_H * P;         // [64]   delta_1, then zeta_1
    float *sm_end = D1 + FT_H * P;
    const int tid = threadIdx.x;
    const int K0 = a.K0, H1n = a.H1, H2n = a.H2;

    for (float *p = sm + tid; p < sm_end; p += FT_THREADS) *p = 0.f;
    __syncthreads();
    for (int i = tid; i < H1n * K0; i += FT_THREADS) { const int n = i / K0, k = i - n * K0; W0T[k * FT_H + n] = __ldg(a.W0 + i); }
    for (int i = tid; i < H2n * H1n; i += FT_THREADS) {
        const int n = i / H1n, k = i - n * H1n;
        const float w = __ldg(a.W1 + i);
        W1T[k * FT_H + n] = w; W1N[n * FT_H + k] = w;
    }
    for (int i = tid; i < H2n; i += FT_THREADS) { W2s[i] = __ldg(a.W2 + i); b1s[i] = __ldg(a.b1 + i); }
    for (int i = tid; i < H1n; i += FT_THREADS) b0s[i] = __ldg(a.b0 + i);
    float gb2_acc = 0.f;  // thread 0 only
    float g1[FT_UN][4], g0[NKG][FT_UN][4];
#pragma unroll
    for (int i = 0; i < FT_UN; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            g1[i][j] = 0.f;
#pragma unroll
            for (int g = 0; g < NKG; ++g) g0[g][i][j] = 0.f;
        }
    __syncthreads();

    const int64_t n_tiles = (a.n_rows + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * TM;
        const int valid = int(min(int64_t(TM), a.n_rows - m0));
        // ---- 1. stage x, u, ge (zeros past the end)
        for (int i = tid; i < TM * K0; i += FT_THREADS) {
            const int m = i / K0, k = i - m * K0;
            const bool ok = m < valid;
            X[k * P + m] = ok ? __ldg(a.x + (m0 + m) * K0 + k) : 0.f;
            U[k * P + m] = (ok && a.u) ? __ldg(a.u + (m0 + m) * K0 + k) : 0.f;
        }
        for (int m = tid; m < TM; m += FT_THREADS) ges[m] = m < valid ? __ldg(a.ge + m0 + m) : 0.f;
        __syncthreads();
        // ---- 2. h1 = sigmoid(W0 x + b0)
        gemm_atoms<TM>(X, K0, W0T, FT_H, [&](int n, int mm, const float (&v)[MM]) {
            float o[MM];
#pragma unroll
            for (int i = 0; i < MM; ++i) o[i] = sigm_f(v[i] + b0s[n]);
            AtomVec<MM>::st(H1 + n * P + mm, o);
        });
        __syncthreads();
        // ---- 3. h2 = sigmoid(W1 h1 + b1);  4. delta_1 = W2 * s(h2)
        gemm_atoms<TM>(H1, H1n, W1T, FT_H, [&](int n, int mm, const float (&v)[MM]) {
            float o[MM], d[MM];
            const float w2 = W2s[n];
#pragma unroll
            for (int i = 0; i < MM; ++i) { o[i] = sigm_f(v[i] + b1s[n]); d[i] = w2 * o[i] * (1.f - o[i]); }
            AtomVec<MM>::st(H2 + n * P + mm, o);
            AtomVec<MM>::st(D1 + n * P + mm, d);
        });
        __syncthreads();
        // ---- 5. r_1[k] = sum_n delta_1[n] W1[n][k];  6. delta_0 = r_1 * s(h1)
        gemm_atoms<TM>(D1, H2n, W1N, FT_H, [&](int k, int mm, const float (&v)[MM]) {
            float h[MM], d[MM];
            AtomVec<MM>::ld(H1 + k * P + mm, h);
#pragma unroll
            for (int i = 0; i < MM; ++i) d[i] = v[i] * h[i] * (1.f - h[i]);
            AtomVec<MM>::st(R1 + k * P + mm, v);
            AtomVec<MM>::st(D0 + k * P + mm, d);
        });
        __syncthreads();
        if (a.u) {
            // ---- 7. gW0 += delta_0 (x) rho_0
#pragma unroll
            for (int g = 0; g < NKG; ++g) outer_acc<TM>(D0, U, 64 * g, min(4, (K0R - 64 * g) / 16), g0[g]);
            // ---- 8. v_0 = W0 rho_0;  rho_1 = v_0 * s(h1);  S_0 = v_0 * r_1 * s(h1) (1 - 2 h1)  (overwrites r_1)
            gemm_atoms<TM>(U, K0, W0T, FT_H, [&](int n, int mm, const float (&v)[MM]) {
                float h[MM], r[MM], rho[MM], S[MM];
                AtomVec<MM>::ld(H1 + n * P + mm, h);
                AtomVec<MM>::ld(R1 + n * P + mm, r);
#pragma unroll
                for (int i = 0; i < MM; ++i) {
                    const float s = h[i] * (1.f - h[i]);
                    rho[i] = v[i] * s;
                    S[i] = v[i] * r[i] * s * (1.f - 2.f * h[i]);
                }
                AtomVec<MM>::st(RHO1 + n * P + mm, rho);
                AtomVec<MM>::st(R1 + n * P + mm, S);
            });
            __syncthreads();
            // ---- 9. gW1 += delta_1 (x) rho_1
            outer_acc<TM>(D1, RHO1, 0, 4, g1);
            __syncthreads();  // step 10 overwrites delta_1
        } else {
            for (float *p = R1 + tid; p < R1 + FT_H * P; p += FT_THREADS) *p = 0.f;  // S_0 = 0, rho_1 = 0 (RHO1 stays zero)
            __syncthreads();
        }
        // ---- 10. v_1 = W1 rho_1;  rho_2 = v_1 * s(h2);  zeta_1 = W2 s(h2) (ge + v_1 (1 - 2 h2))  (overwrites delta_1)
        //      11. gW2 += sum_m rho_2 + ge h2 ;  gb1 += sum_m zeta_1   (each unit has one owner thread: no race)
        gemm_atoms<TM>(RHO1, H1n, W1T, FT_H, [&](int n, int mm, const float (&v)[MM]) {
            float h[MM], q[MM], z[MM];
            AtomVec<MM>::ld(H2 + n * P + mm, h);
            const float w2 = W2s[n];
#pragma unroll
            for (int i = 0; i < MM; ++i) {
                const float s = h[i] * (1.f - h[i]);
                q[i] = v[i] * s + ges[mm + i] * h[i];
                z[i] = w2 * s * (ges[mm + i] + v[i] * (1.f - 2.f * h[i]));
            }
            AtomVec<MM>::st(D1 + n * P + mm, z);
            const float sq = unit_sum<MM>(q), sz = unit_sum<MM>(z);
            if ((tid & 15) == 0) { gw2s[n] += sq; gb1s[n] += sz; }
        });
        if (tid == 0) {  // gb2 += sum_m ge
            float s = 0.f;
            for (int m = 0; m < TM; ++m) s += ges[m];
            gb2_acc += s;
        }
        __syncthreads();
        // ---- 12. gW1 += zeta_1 (x) h1
        outer_acc<TM>(D1, H1, 0, 4, g1);
        // ---- 13. zeta_0[k] = (sum_n W1[n][k] zeta_1[n]) s(h1) + S_0   (overwrites delta_0; all readers of delta_0 are done)
        gemm_atoms<TM>(D1, H2n, W1N, FT_H, [&](int k, int mm, const float (&v)[MM]) {
            float h[MM], S[MM], z[MM];
            AtomVec<MM>::ld(H1 + k * P + mm, h);
            AtomVec<MM>::ld(R1 + k * P + mm, S);
#pragma unroll
            for (int i = 0; i < MM; ++i) z[i] = v[i] * h[i] * (1.f - h[i]) + S[i];
            AtomVec<MM>::st(D0 + k * P + mm, z);
            const float sz = unit_sum<MM>(z);  // gb0 += sum_m zeta_0
            if ((tid & 15) == 0) gb0s[k] += sz;
        });
        __syncthreads();
        // ---- 14. gW0 += zeta_0 (x) x
#pragma unroll
        for (int g = 0; g < NKG; ++g) outer_acc<TM>(D0, X, 64 * g, min(4, (K0R - 64 * g) / 16), g0[g]);
        __syncthreads();
    }
    // ---- write-back: one atomicAdd per weight and block
    {
        const int tk = tid & 15, tn = tid >> 4;
#pragma unroll
        for (int i = 0; i < FT_UN; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = tn + FT_TNS * i, k = tk + 16 * j;
                if (n < H2n && k < H1n) atomicAdd(a.gW1 + n * H1n + k, g1[i][j]);
#pragma unroll
                for (int g = 0; g < NKG; ++g) {
                    const int k0 = 64 * g + k;
                    if (n < H1n && k0 < K0) atomicAdd(a.gW0 + n * K0 + k0, g0[g][i][j]);
                }
            }
        for (int n = tid; n < FT_H; n += FT_THREADS) {
            if (n < H1n) atomicAdd(a.gb0 + n, gb0s[n]);
            if (n < H2n) { atomicAdd(a.gb1 + n, gb1s[n]); atomicAdd(a.gW2 + n, gw2s[n]); }
        }
        if (tid == 0) atomicAdd(a.gb2, gb2_acc);
    }
}

template <int TM>
static size_t fast_smem_bytes(int K0R) {
    constexpr int P = TM + 4;
    return sizeof(float) * (size_t(K0R) * FT_H + 2 * FT_H * FT_H + 6 * FT_H + TM + 2 * size_t(K0R) * P + 6 * FT_H * P);
}

template <int TM, int NKG>
static int fast_launch(const FastTrainArgs &a, cudaStream_t st) {
    const size_t smem = fast_smem_bytes<TM>(a.K0R);
    auto kern = mlp_train_fast_kernel<TM, NKG>;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 1;
    NNMD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FT_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t n_tiles = (a.n_rows + TM - 1) / TM;
    const int64_t blocks = std::min<int64_t>(n_tiles, int64_t(sm_count()) * occ);
    ProfScope prof(PROF_MLP, st);
    kern<<<unsigned(blocks), FT_THREADS, smem, st>>>(a);
    NNMD_LAUNCH_CHECK("mlp_train_fast_kernel");
    return NNMD_OK;
}

// Returns NNMD_OK when the fast kernel ran, a negative value when the shape is not covered (the caller then uses the
// generic kernel), or an error code.
int train_bwd_fast(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases, const void *x,
                   int64_t n_rows, const void *ge, const void *u, void *const *gW, void *const *gb, cudaStream_t st) {
    if (n_layers != 3 || sizes[1] > FT_H || sizes[2] > FT_H || sizes[0] > 128 || sizes[3] != 1) return -1;
    if (const char *env = tuning_env("NNMD_MLP_TRAIN_GENERIC")) if (env[0] == '1') return -1;
    FastTrainArgs a;
    a.K0 = sizes[0]; a.H1 = sizes[1]; a.H2 = sizes[2]; a.K0R = (sizes[0] + 15) / 16 * 16;
    a.W0 = (const float *)weights[0]; a.W1 = (const float *)weights[1]; a.W2 = (const float *)weights[2];
    a.b0 = (const float *)biases[0]; a.b1 = (const float *)biases[1];
    a.gW0 = (float *)gW[0]; a.gW1 = (float *)gW[1]; a.gW2 = (float *)gW[2];
    a.gb0 = (float *)gb[0]; a.gb1 = (float *)gb[1]; a.gb2 = (float *)gb[2];
    a.x = (const float *)x; a.ge = (const float *)ge; a.u = (const float *)u; a.n_rows = n_rows;
    // tile height: 32-atom tiles let two blocks share an SM while the input axis is short; otherwise take the height
    // with the shorter critical path (rounds x atoms per tile); inputs wider than 64 only fit with 32-atom tiles
    if (a.K0R > 64) return fast_launch<32, 2>(a, st);
    const int64_t sms = sm_count();
    const int64_t occ32 = fast_smem_bytes<32>(a.K0R) <= 112 * 1024 ? 2 : 1;
    const int64_t path64 = ((n_rows + 63) / 64 + sms - 1) / sms * 64;
    const int64_t path32 = ((n_rows + 31) / 32 + sms * occ32 - 1) / (sms * occ32) * 32;
    return path32 <= path64 ? fast_launch<32, 1>(a, st) : fast_launch<64, 1>(a, st);
}

// ---------------------------------------------------------------------------------------------------------------
// K6 for the same shapes: e = net(x) and dE/dx (reference: bpnn.py:343-349) with the tile machinery above.
struct FastFwdArgs {
    int K0, H1, H2, K0R;
    const float *W0, *b0, *W1, *b1, *W2, *b2;
    const float *x;
    float *e, *dE;  // dE may be nullptr
    int64_t n_rows;
};

template <int TM>
__global__ void __launch_bounds__(FT_THREADS, FT_MINB) mlp_fwd_fast_kernel(FastFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int P = TM + 4, MM = TM / 16;
    const int K0R = a.K0R;
    float *sm = reinterpret_cast<float *>(smem_raw);
    float *W0T = sm;                    // [K0R][64]
    float *W0N = W0T + K0R * FT_H;      // [64][K0R]   W0N[n][k] = W0[n][k]
    float *W1T = W0N + FT_H * K0R;      // [64][64]
    float *W1N = W1T + FT_H * FT_H;     // [64][64]
    float *W2s = W1N + FT_H * FT_H, *b0s = W2s + FT_H, *b1s = b0s + FT_H;
    float *ep = b1s + FT_H;             // [FT_TNS][TM] partial energies (one row per tn)
    float *X = ep + FT_TNS * TM;            // [K0R], later dE/dx of the tile
    float *H1 = X + K0R * P, *H2 = H1 + FT_H * P, *D1 = H2 + FT_H * P, *D0 = D1 + FT_H * P;
    float *sm_end = D0 + FT_H * P;
    const int tid = threadIdx.x, tn = tid >> 4;
    const int K0 = a.K0, H1n = a.H1, H2n = a.H2;
    for (float *p = sm + tid; p < sm_end; p += FT_THREADS) *p = 0.f;
    __syncthreads();
    for (int i = tid; i < H1n * K0; i += FT_THREADS) {
        const int n = i / K0, k = i - n * K0;
        const float w = __ldg(a.W0 + i);
        W0T[k * FT_H + n] = w; W0N[n * K0R + k] = w;
    }
    for (int i = tid; i < H2n * H1n; i += FT_THREADS) {
        const int n = i / H1n, k = i - n * H1n;
        const float w = __ldg(a.W1 + i);
        W1T[k * FT_H + n] = w; W1N[n * FT_H + k] = w;
    }
    for (int i = tid; i < H2n; i += FT_THREADS) { W2s[i] = __ldg(a.W2 + i); b1s[i] = __ldg(a.b1 + i); }
    for (int i = tid; i < H1n; i += FT_THREADS) b0s[i] = __ldg(a.b0 + i);
    const float b2 = __ldg(a.b2);
    __syncthreads();
    const int64_t n_tiles = (a.n_rows + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * TM;
        const int valid = int(min(int64_t(TM), a.n_rows - m0));
        for (int i = tid; i < TM * K0; i += FT_THREADS) {
            const int m = i / K0, k = i - m * K0;
            X[k * P + m] = m < valid ? __ldg(a.x + (m0 + m) * K0 + k) : 0.f;
        }
        __syncthreads();
        gemm_atoms<TM>(X, K0, W0T, FT_H, [&](int n, int mm, const float (&v)[MM]) {
            float o[MM];
#pragma unroll
            for (int i = 0; i < MM; ++i) o[i] = sigm_f(v[i] + b0s[n]);
            AtomVec<MM>::st(H1 + n * P + mm, o);
        });
        __syncthreads();
        {
            float pe[MM];
#pragma unroll
            for (int i = 0; i < MM; ++i) pe[i] = 0.f;
            gemm_atoms<TM>(H1, H1n, W1T, FT_H, [&](int n, int mm, const float (&v)[MM]) {
                float o[MM], d[MM];
                const float w2 = W2s[n];
#pragma unroll
                for (int i = 0; i < MM; ++i) {
                    o[i] = sigm_f(v[i] + b1s[n]);
                    d[i] = w2 * o[i] * (1.f - o[i]);
                    pe[i] = fmaf(w2, o[i], pe[i]);
                }
                AtomVec<MM>::st(D1 + n * P + mm, d);
            });
            AtomVec<MM>::st(ep + tn * TM + MM * (tid & 15), pe);
        }
        __syncthreads();
        for (int m = tid; m < valid; m += FT_THREADS) {
            float s = b2;
#pragma unroll
            for (int r = 0; r < FT_TNS; ++r) s += ep[r * TM + m];
            a.e[m0 + m] = s;
        }
        if (a.dE) {
            // delta_0 = ((delta_1 W1) * s(h1));  dE/dx[k] = sum_n delta_0[n] W0[n][k]
            gemm_atoms<TM>(D1, H2n, W1N, FT_H, [&](int k, int mm, const float (&v)[MM]) {
                float h[MM], d[MM];
                AtomVec<MM>::ld(H1 + k * P + mm, h);
#pragma unroll
                for (int i = 0; i < MM; ++i) d[i] = v[i] * h[i] * (1.f - h[i]);
                AtomVec<MM>::st(D0 + k * P + mm, d);
            });
            __syncthreads();
            for (int c0 = 0; c0 < K0R; c0 += FT_H)
                gemm_atoms<TM>(D0, H1n, W0N + c0, K0R, [&](int k, int mm, const float (&v)[MM]) {
                    AtomVec<MM>::st(X + (c0 + k) * P + mm, v);
                }, K0R - c0);
            __syncthreads();
            for (int i = tid; i < valid * K0; i += FT_THREADS) {
                const int m = i / K0, k = i - m * K0;
                a.dE[m0 * K0 + i] = X[k * P + m];
            }
        }
        __syncthreads();
    }
}

template <int TM>
static size_t fwd_smem_bytes(int K0R) {
    constexpr int P = TM + 4;
    return sizeof(float) * (2 * size_t(K0R) * FT_H + 2 * FT_H * FT_H + 3 * FT_H + FT_TNS * TM + size_t(K0R) * P + 4 * FT_H * P);
}

template <int TM>
static int fwd_launch(const FastFwdArgs &a, cudaStream_t st) {
    const size_t smem = fwd_smem_bytes<TM>(a.K0R);
    auto kern = mlp_fwd_fast_kernel<TM>;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 1;
    NNMD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FT_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t n_tiles = (a.n_rows + TM - 1) / TM;
    const int64_t blocks = std::min<int64_t>(n_tiles, int64_t(sm_count()) * occ);
    ProfScope prof(PROF_MLP, st);
    kern<<<unsigned(blocks), FT_THREADS, smem, st>>>(a);
    NNMD_LAUNCH_CHECK("mlp_fwd_fast_kernel");
    return NNMD_OK;
}

// Negative return: shape not covered (the caller falls back to the other fp32 kernels).
int mlp_fwd_fast(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases, const void *x,
                 int64_t n_rows, void *e, void *dEdx, cudaStream_t st) {
    if (n_layers != 3 || sizes[1] > FT_H || sizes[2] > FT_H || sizes[0] > 128 || sizes[3] != 1) return -1;
    FastFwdArgs a;
    a.K0 = sizes[0]; a.H1 = sizes[1]; a.H2 = sizes[2]; a.K0R = (sizes[0] + 15) / 16 * 16;
    a.W0 = (const float *)weights[0]; a.W1 = (const float *)weights[1]; a.W2 = (const float *)weights[2];
    a.b0 = (const float *)biases[0]; a.b1 = (const float *)biases[1]; a.b2 = (const float *)biases[2];
    a.x = (const float *)x; a.e = (float *)e; a.dE = (float *)dEdx; a.n_rows = n_rows;
    if (fwd_smem_bytes<64>(a.K0R) > 200 * 1024) return fwd_launch<32>(a, st);
    const int64_t sms = sm_count();
    const int64_t occ32 = std::max<int64_t>(1, (227 * 1024) / int64_t(fwd_smem_bytes<32>(a.K0R) + 1024));
    const int64_t occ64 = std::max<int64_t>(1, (227 * 1024) / int64_t(fwd_smem_bytes<64>(a.K0R) + 1024));
    const int64_t path64 = ((n_rows + 63) / 64 + sms * occ64 - 1) / (sms * occ64) * 64;
    const int64_t path32 = ((n_rows + 31) / 32 + sms * occ32 - 1) / (sms * occ32) * 32;
    return path32 <= path64 ? fwd_launch<32>(a, st) : fwd_launch<64>(a, st);
}

}  // namespace nnmd
