// mlp_tc.cu -- tensor-core variant of K6 (NNMD_MLP_TF32): e = MLP(x) and dE/dx for two-hidden-layer atomic networks.
//
// Replaces nnmd/nn/atomic_nn.py:43-48 + the autograd.grad of nnmd/nn/bpnn.py:347-349 like mlp.cu does, but runs the four
// dense contractions (two forward, two backward) on the tensor cores with TF32 operands and fp32 accumulation.
// Plain TF32 (10-bit mantissa) cannot meet the 1e-5 fp32 parity bar, so every product is error-compensated
// ("3xTF32"): a = a_hi + a_lo, b = b_hi + b_lo, a*b ~ a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, three MMAs per tile.
//
// One warp owns 16 atoms and never leaves registers between layers: the accumulator fragment of layer l (rows g, g+8;
// columns 2t, 2t+1 of an 8-wide tile) IS the A fragment of layer l+1 once the k-slots of that MMA are read as
// (slot t -> neuron 8j+2t, slot t+4 -> neuron 8j+2t+1); the B fragments are laid out in shared memory to match, already
// split into hi/lo, one 16 B load per (k-step, n-tile, lane).  The backward sweep reuses the same trick with the
// transposed weight fragments.  Shapes are template parameters (tiles of 8): <input tiles, hidden-1 tiles, hidden-2 tiles>.
#include <algorithm>

#include "common.cuh"

namespace nnmd {

struct TcArgs {
    const float *W0, *b0, *W1, *b1, *W2, *b2;  // W0 (H1,K0)  W1 (H2,H1)  W2 (1,H2), torch.nn.Linear layout
    int K0, H1, H2;
    const float *x;
    int64_t n_rows;
    float *e, *dEdx;
    // fused form (K5 + K6 + K4c in one pass): x holds the UN-normalised descriptors G; the kernel normalises them itself
    // (calculate_sf.py:100-101), optionally writes g_hat back, and contracts dE/dg_hat with dG into the force (bpnn.py:353)
    float *ghat;         // (n_rows, K0) normalised descriptors out, or NULL (may alias x)
    const float *dG;     // (n_rows, K0, 3)
    float *force;        // (n_rows, 3)
    uint32_t *flag;      // NNMD_FLAG_NAN_G on 0/0
};

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// c[NT][4] += A (16 x 8*KT, fragments a[KT][4]) * B, B fragments at bf[(j*NT + i)*32 + lane] = (b0_hi, b1_hi, b0_lo, b1_lo)
template <int KT, int NT>
__device__ __forceinline__ void gemm3x(float (&c)[NT][4], const float (&a)[KT][4], const float4 *__restrict__ bf, int lane) {
#pragma unroll
    for (int j = 0; j < KT; ++j) {
        float ah[4], al[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { ah[r] = tf32_hi(a[j][r]); al[r] = a[j][r] - ah[r]; }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const float4 b = bf[(j * NT + i) * 32 + lane];
            mma_tf32(c[i], al, b.x, b.y);
            mma_tf32(c[i], ah, b.z, b.w);
            mma_tf32(c[i], ah, b.x, b.y);
        }
    }
}

// Fill B fragments for out[m][n] = sum_k in[m][k] * Wsrc(k, n):  element (k, n) = W[n*ld_n + k*ld_k] (zero outside kmax, nmax)
__device__ __forceinline__ void fill_frags(float4 *dst, const float *W, int ld_n, int ld_k, int KT, int NT, int kmax, int nmax) {
    for (int idx = threadIdx.x; idx < KT * NT * 32; idx += blockDim.x) {
        const int lane = idx & 31, tile = idx >> 5, i = tile % NT, j = tile / NT;
        const int g = lane >> 2, t = lane & 3;
        const int n = 8 * i + g, k0 = 8 * j + 2 * t, k1 = k0 + 1;
        const float v0 = (n < nmax && k0 < kmax) ? __ldg(W + size_t(n) * ld_n + size_t(k0) * ld_k) : 0.f;
        const float v1 = (n < nmax && k1 < kmax) ? __ldg(W + size_t(n) * ld_n + size_t(k1) * ld_k) : 0.f;
        const float h0 = tf32_hi(v0), h1 = tf32_hi(v1);
        dst[idx] = make_float4(h0, h1, v0 - h0, v1 - h1);
    }
}

// warps per block: they share one copy of the weight fragments (160 kB for 96-64-64-1), so the warps of ONE block are all the
// memory-level parallelism an SM gets for the dG rows of the fused epilogue
#ifndef NNMD_TC_WARPS
#define NNMD_TC_WARPS 16  // 8: 1.50 ms, 12: 1.23, 16: 1.17, 20: 1.16 (spills), 24: 1.29 per 1M atoms (fused tail, 96-64-64-1)
#endif
constexpr int TC_WARPS = NNMD_TC_WARPS;
template <int K0T, int H1T, int H2T, bool FUSED = false>
__global__ void __launch_bounds__(32 * TC_WARPS, 1) mlp_tc_kernel(TcArgs a) {
    if (FUSED && a.flag != nullptr && (*a.flag & NNMD_FLAG_OVERFLOW)) return;  // incomplete neighbour list: G / dG are not valid
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *f_fwd0 = reinterpret_cast<float4 *>(smem_raw);   // x  (K0) -> z1 (H1):  k = input feature, n = neuron of layer 0
    float4 *f_fwd1 = f_fwd0 + K0T * H1T * 32;                 // h1 (H1) -> z2 (H2)
    float4 *f_bwd1 = f_fwd1 + H1T * H2T * 32;                 // d2 (H2) -> g1 (H1):  k = neuron of layer 1, n = neuron of layer 0 output
    float4 *f_bwd0 = f_bwd1 + H2T * H1T * 32;                 // d1 (H1) -> dE/dx (K0)
    float *bias0 = reinterpret_cast<float *>(f_bwd0 + H1T * K0T * 32);
    float *bias1 = bias0 + 8 * H1T;
    float *w2 = bias1 + 8 * H2T;
    // forward: element (k, n) = W[n][k]; backward: element (k = out neuron, n = in feature) = W[k][n]
    fill_frags(f_fwd0, a.W0, a.K0, 1, K0T, H1T, a.K0, a.H1);
    fill_frags(f_fwd1, a.W1, a.H1, 1, H1T, H2T, a.H1, a.H2);
    fill_frags(f_bwd1, a.W1, 1, a.H1, H2T, H1T, a.H2, a.H1);
    fill_frags(f_bwd0, a.W0, 1, a.K0, H1T, K0T, a.H1, a.K0);
    for (int i = threadIdx.x; i < 8 * H1T; i += blockDim.x) bias0[i] = i < a.H1 ? __ldg(a.b0 + i) : 0.f;
    for (int i = threadIdx.x; i < 8 * H2T; i += blockDim.x) { bias1[i] = i < a.H2 ? __ldg(a.b1 + i) : 0.f; w2[i] = i < a.H2 ? __ldg(a.W2 + i) : 0.f; }
    __syncthreads();
    const float b2 = __ldg(a.b2);
    const int lane = lane_id(), warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int64_t n_groups = (a.n_rows + 15) / 16;
    bool bad = false;
    for (int64_t grp = int64_t(blockIdx.x) * TC_WARPS + warp; grp < n_groups; grp += int64_t(gridDim.x) * TC_WARPS) {
        const int64_t r0 = grp * 16 + g, r1 = r0 + 8;
        const bool v0 = r0 < a.n_rows, v1 = r1 < a.n_rows;
        if (FUSED) {
            // the 16 rows of dG this warp contracts at the END of the pass (16 x K0 x 12 B, contiguous): ask L2 for them now
            const char *base = reinterpret_cast<const char *>(a.dG + size_t(grp) * 16 * a.K0 * 3);
            const int64_t rows_here = (a.n_rows - grp * 16) < 16 ? (a.n_rows - grp * 16) : 16;
            const int64_t bytes = rows_here * a.K0 * 12;
            for (int64_t o = int64_t(lane) * 128; o < bytes; o += 32 * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
        }
        // ---- x as A fragments
        float xa[K0T][4];
#pragma unroll
        for (int j = 0; j < K0T; ++j) {
            const int k = 8 * j + 2 * t;
            xa[j][0] = (v0 && k < a.K0) ? __ldg(a.x + r0 * a.K0 + k) : 0.f;
            xa[j][2] = (v0 && k + 1 < a.K0) ? __ldg(a.x + r0 * a.K0 + k + 1) : 0.f;
            xa[j][1] = (v1 && k < a.K0) ? __ldg(a.x + r1 * a.K0 + k) : 0.f;
            xa[j][3] = (v1 && k + 1 < a.K0) ? __ldg(a.x + r1 * a.K0 + k + 1) : 0.f;
        }
        if (FUSED) {
            // K5: g_hat = G / ||G||_2 per atom; a row is spread over the four lanes of a quad
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < K0T; ++j) {
                s0 = fmaf(xa[j][0], xa[j][0], fmaf(xa[j][2], xa[j][2], s0));
                s1 = fmaf(xa[j][1], xa[j][1], fmaf(xa[j][3], xa[j][3], s1));
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            const float n0 = sqrtf(s0), n1 = sqrtf(s1);
#pragma unroll
            for (int j = 0; j < K0T; ++j) {
                const int k = 8 * j + 2 * t;
                // 0/0 -> NaN for an atom without neighbours, like calculate_sf.py:101 (padding columns stay 0)
                xa[j][0] = (k < a.K0) ? xa[j][0] / n0 : 0.f;  xa[j][2] = (k + 1 < a.K0) ? xa[j][2] / n0 : 0.f;
                xa[j][1] = (k < a.K0) ? xa[j][1] / n1 : 0.f;  xa[j][3] = (k + 1 < a.K0) ? xa[j][3] / n1 : 0.f;
                bad |= (v0 && (xa[j][0] != xa[j][0] || xa[j][2] != xa[j][2])) || (v1 && (xa[j][1] != xa[j][1] || xa[j][3] != xa[j][3]));
                if (a.ghat) {
                    if (v0 && k < a.K0) a.ghat[r0 * a.K0 + k] = xa[j][0];
                    if (v0 && k + 1 < a.K0) a.ghat[r0 * a.K0 + k + 1] = xa[j][2];
                    if (v1 && k < a.K0) a.ghat[r1 * a.K0 + k] = xa[j][1];
                    if (v1 && k + 1 < a.K0) a.ghat[r1 * a.K0 + k + 1] = xa[j][3];
                }
            }
        }
        // ---- layer 0
        float h1[H1T][4];
#pragma unroll
        for (int i = 0; i < H1T; ++i) { h1[i][0] = h1[i][2] = bias0[8 * i + 2 * t]; h1[i][1] = h1[i][3] = bias0[8 * i + 2 * t + 1]; }
        gemm3x<K0T, H1T>(h1, xa, f_fwd0, lane);
        // accumulator (row g: c0 c1 | row g+8: c2 c3) -> A fragment (a0 = row g slot t, a1 = row g+8 slot t, a2 = row g slot t+4, a3 = row g+8 slot t+4)
#pragma unroll
        for (int i = 0; i < H1T; ++i) {
            const float s0 = 1.f / (1.f + expf(-h1[i][0])), s1 = 1.f / (1.f + expf(-h1[i][1]));
            const float s2 = 1.f / (1.f + expf(-h1[i][2])), s3 = 1.f / (1.f + expf(-h1[i][3]));
            h1[i][0] = s0; h1[i][1] = s2; h1[i][2] = s1; h1[i][3] = s3;
        }
        // ---- layer 1
        float h2[H2T][4];
#pragma unroll
        for (int i = 0; i < H2T; ++i) { h2[i][0] = h2[i][2] = bias1[8 * i + 2 * t]; h2[i][1] = h2[i][3] = bias1[8 * i + 2 * t + 1]; }
        gemm3x<H1T, H2T>(h2, h1, f_fwd1, lane);
        // ---- output layer (one neuron) on the FMA pipe, and d2 = w2 * h2 (1 - h2) in A-fragment order
        float e0 = 0.f, e1 = 0.f;
#pragma unroll
        for (int i = 0; i < H2T; ++i) {
            const float wa = w2[8 * i + 2 * t], wb = w2[8 * i + 2 * t + 1];
            const float s0 = 1.f / (1.f + expf(-h2[i][0])), s1 = 1.f / (1.f + expf(-h2[i][1]));
            const float s2 = 1.f / (1.f + expf(-h2[i][2])), s3 = 1.f / (1.f + expf(-h2[i][3]));
            e0 = fmaf(wa, s0, fmaf(wb, s1, e0));
            e1 = fmaf(wa, s2, fmaf(wb, s3, e1));
            h2[i][0] = wa * s0 * (1.f - s0); h2[i][2] = wb * s1 * (1.f - s1);
            h2[i][1] = wa * s2 * (1.f - s2); h2[i][3] = wb * s3 * (1.f - s3);
        }
        e0 += __shfl_xor_sync(0xffffffffu, e0, 1); e0 += __shfl_xor_sync(0xffffffffu, e0, 2);
        e1 += __shfl_xor_sync(0xffffffffu, e1, 1); e1 += __shfl_xor_sync(0xffffffffu, e1, 2);
        if (t == 0) {
            if (v0) a.e[r0] = e0 + b2;
            if (v1) a.e[r1] = e1 + b2;
        }
        if (FUSED || a.dEdx) {
            // ---- g1 = d2 W1, d1 = g1 * h1 (1 - h1)
            float d1[H1T][4];
#pragma unroll
            for (int i = 0; i < H1T; ++i) d1[i][0] = d1[i][1] = d1[i][2] = d1[i][3] = 0.f;
            gemm3x<H2T, H1T>(d1, h2, f_bwd1, lane);
#pragma unroll
            for (int i = 0; i < H1T; ++i) {
                // d1 is an accumulator (c0 c1 | c2 c3), h1 already sits in A-fragment order (a0 a1 a2 a3) = (c0 c2 c1 c3)
                const float q0 = d1[i][0] * h1[i][0] * (1.f - h1[i][0]), q1 = d1[i][1] * h1[i][2] * (1.f - h1[i][2]);
                const float q2 = d1[i][2] * h1[i][1] * (1.f - h1[i][1]), q3 = d1[i][3] * h1[i][3] * (1.f - h1[i][3]);
                d1[i][0] = q0; d1[i][1] = q2; d1[i][2] = q1; d1[i][3] = q3;
            }
            // ---- dE/dx = d1 W0
            float gx[K0T][4];
#pragma unroll
            for (int i = 0; i < K0T; ++i) gx[i][0] = gx[i][1] = gx[i][2] = gx[i][3] = 0.f;
            gemm3x<H1T, K0T>(gx, d1, f_bwd0, lane);
            if (a.dEdx) {
#pragma unroll
                for (int i = 0; i < K0T; ++i) {
                    const int k = 8 * i + 2 * t;
                    if (v0 && k < a.K0) a.dEdx[r0 * a.K0 + k] = gx[i][0];
                    if (v0 && k + 1 < a.K0) a.dEdx[r0 * a.K0 + k + 1] = gx[i][1];
                    if (v1 && k < a.K0) a.dEdx[r1 * a.K0 + k] = gx[i][2];
                    if (v1 && k + 1 < a.K0) a.dEdx[r1 * a.K0 + k + 1] = gx[i][3];
                }
            }
            if (FUSED) {
                // K4c: F[a,:] = - sum_f dE/dg_hat[a,f] dG[a,f,:].  A lane holds two adjacent functions of rows r0 and r1:
                // their six gradient components are 24 contiguous bytes of dG.
                float f0[3] = {0.f, 0.f, 0.f}, f1[3] = {0.f, 0.f, 0.f};
                const bool pairs_ok = (a.K0 & 1) == 0;  // rows of an even number of functions keep the 8 B alignment of a pair
#pragma unroll
                for (int i = 0; i < K0T; ++i) {
                    const int k = 8 * i + 2 * t;
                    if (pairs_ok && k + 1 < a.K0) {
                        if (v0) {
                            const float2 *p = reinterpret_cast<const float2 *>(a.dG + (r0 * a.K0 + k) * 3);
                            const float2 q0 = __ldcs(p), q1 = __ldcs(p + 1), q2 = __ldcs(p + 2);   // (x y) (z | x) (y z)
                            f0[0] = fmaf(gx[i][0], q0.x, fmaf(gx[i][1], q1.y, f0[0]));
                            f0[1] = fmaf(gx[i][0], q0.y, fmaf(gx[i][1], q2.x, f0[1]));
                            f0[2] = fmaf(gx[i][0], q1.x, fmaf(gx[i][1], q2.y, f0[2]));
                        }
                        if (v1) {
                            const float2 *p = reinterpret_cast<const float2 *>(a.dG + (r1 * a.K0 + k) * 3);
                            const float2 q0 = __ldcs(p), q1 = __ldcs(p + 1), q2 = __ldcs(p + 2);
                            f1[0] = fmaf(gx[i][2], q0.x, fmaf(gx[i][3], q1.y, f1[0]));
                            f1[1] = fmaf(gx[i][2], q0.y, fmaf(gx[i][3], q2.x, f1[1]));
                            f1[2] = fmaf(gx[i][2], q1.x, fmaf(gx[i][3], q2.y, f1[2]));
                        }
                    } else {  // odd number of functions: scalar loads
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            if (k + q >= a.K0) continue;
                            if (v0) {
                                const float *p = a.dG + (r0 * a.K0 + k + q) * 3;
                                const float w = gx[i][q];
                                f0[0] = fmaf(w, __ldg(p), f0[0]); f0[1] = fmaf(w, __ldg(p + 1), f0[1]); f0[2] = fmaf(w, __ldg(p + 2), f0[2]);
                            }
                            if (v1) {
                                const float *p = a.dG + (r1 * a.K0 + k + q) * 3;
                                const float w = gx[i][2 + q];
                                f1[0] = fmaf(w, __ldg(p), f1[0]); f1[1] = fmaf(w, __ldg(p + 1), f1[1]); f1[2] = fmaf(w, __ldg(p + 2), f1[2]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    f0[c] += __shfl_xor_sync(0xffffffffu, f0[c], 1); f0[c] += __shfl_xor_sync(0xffffffffu, f0[c], 2);
                    f1[c] += __shfl_xor_sync(0xffffffffu, f1[c], 1); f1[c] += __shfl_xor_sync(0xffffffffu, f1[c], 2);
                }
                if (t == 0) {
                    if (v0) { a.force[3 * r0] = -f0[0]; a.force[3 * r0 + 1] = -f0[1]; a.force[3 * r0 + 2] = -f0[2]; }
                    if (v1) { a.force[3 * r1] = -f1[0]; a.force[3 * r1 + 1] = -f1[1]; a.force[3 * r1 + 2] = -f1[2]; }
                }
            }
        }
    }
    if (FUSED && bad && a.flag) atomicOr(a.flag, NNMD_FLAG_NAN_G);
}

template <int K0T, int H1T, int H2T, bool FUSED = false>
static int tc_launch(const TcArgs &a, cudaStream_t st) {
    const size_t smem = size_t(K0T * H1T + H1T * H2T + H2T * H1T + H1T * K0T) * 32 * sizeof(float4) + size_t(8 * H1T + 16 * H2T) * sizeof(float);
    if (smem > 220 * 1024) return set_err(NNMD_ERR_CAPACITY, "mlp(tf32): weights need %zu B of shared memory", smem);
    auto kern = mlp_tc_kernel<K0T, H1T, H2T, FUSED>;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 1;
    NNMD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * TC_WARPS, smem));
    if (occ < 1) occ = 1;
    const int64_t n_groups = (a.n_rows + 15) / 16;
    const int64_t blocks = std::min<int64_t>((n_groups + TC_WARPS - 1) / TC_WARPS, int64_t(sm_count()) * occ);
    ProfScope prof(PROF_MLP, st);
    kern<<<unsigned(blocks), 32 * TC_WARPS, smem, st>>>(a);
    NNMD_LAUNCH_CHECK("mlp_tc_kernel");
    return NNMD_OK;
}

int mlp_tf32_run(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases, const void *x,
                 int64_t n_rows, void *e, void *dEdx, cudaStream_t st) {
    if (n_layers != 3) return set_err(NNMD_ERR_INVALID, "mlp(tf32): the tensor-core path covers networks with two hidden layers");
    TcArgs a;
    a.W0 = (const float *)weights[0]; a.b0 = (const float *)biases[0];
    a.W1 = (const float *)weights[1]; a.b1 = (const float *)biases[1];
    a.W2 = (const float *)weights[2]; a.b2 = (const float *)biases[2];
    a.K0 = sizes[0]; a.H1 = sizes[1]; a.H2 = sizes[2];
    a.x = (const float *)x; a.n_rows = n_rows; a.e = (float *)e; a.dEdx = (float *)dEdx;
    a.ghat = nullptr; a.dG = nullptr; a.force = nullptr; a.flag = nullptr;
    const int k0t = (a.K0 + 7) / 8, h1t = (a.H1 + 7) / 8, h2t = (a.H2 + 7) / 8;
    if (k0t <= 2 && h1t <= 4 && h2t <= 4) return tc_launch<2, 4, 4>(a, st);     // e.g. LJ 3-25-25-1
    if (k0t <= 2 && h1t <= 8 && h2t <= 8) return tc_launch<2, 8, 8>(a, st);     // e.g. Li 14-64-64-1
    if (k0t <= 7 && h1t <= 8 && h2t <= 8) return tc_launch<7, 8, 8>(a, st);     // up to 56 descriptors
    if (k0t <= 12 && h1t <= 8 && h2t <= 8) return tc_launch<12, 8, 8>(a, st);   // the 96-function bench table
    return set_err(NNMD_ERR_CAPACITY, "mlp(tf32): shape %d-%d-%d-1 has no tensor-core instantiation", a.K0, a.H1, a.H2);
}

// K5 + K6 + K4c in one pass over G / dG.  Returns a negative value when the shape has no tensor-core instantiation.
int energy_force_fused_run(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases,
                           const void *G, const void *dG, int64_t n_rows, void *ghat, void *e, void *dEdx, void *force,
                           uint32_t *flag, cudaStream_t st) {
    if (n_layers != 3) return -1;
    TcArgs a;
    a.W0 = (const float *)weights[0]; a.b0 = (const float *)biases[0];
    a.W1 = (const float *)weights[1]; a.b1 = (const float *)biases[1];
    a.W2 = (const float *)weights[2]; a.b2 = (const float *)biases[2];
    a.K0 = sizes[0]; a.H1 = sizes[1]; a.H2 = sizes[2];
    a.x = (const float *)G; a.n_rows = n_rows; a.e = (float *)e; a.dEdx = (float *)dEdx;
    a.ghat = (float *)ghat; a.dG = (const float *)dG; a.force = (float *)force; a.flag = flag;
    const int k0t = (a.K0 + 7) / 8, h1t = (a.H1 + 7) / 8, h2t = (a.H2 + 7) / 8;
    if (k0t <= 2 && h1t <= 8 && h2t <= 8) return tc_launch<2, 8, 8, true>(a, st);
    if (k0t <= 7 && h1t <= 8 && h2t <= 8) return tc_launch<7, 8, 8, true>(a, st);
    if (k0t <= 12 && h1t <= 8 && h2t <= 8) return tc_launch<12, 8, 8, true>(a, st);
    return -1;
}

}  // namespace nnmd

extern "C" int nnmd_energy_force_fused(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases,
                                       const void *G, const void *dG, int64_t n_rows, void *ghat, void *e, void *dEdx, void *force,
                                       uint32_t *flag, void *stream) {
    using namespace nnmd;
    if (n_layers != 3 || !sizes) return set_err(NNMD_ERR_CAPACITY, "fused energy+force: two hidden layers only");
    if (sizes[3] != 1) return set_err(NNMD_ERR_INVALID, "fused energy+force: the atomic network has a single output");
    if (n_rows < 0) return set_err(NNMD_ERR_INVALID, "fused energy+force: negative row count");
    if (n_rows == 0) return NNMD_OK;
    if (!weights || !biases || !G || !dG || !e || !force || !flag) return set_err(NNMD_ERR_INVALID, "fused energy+force: null pointer");
    const int r = energy_force_fused_run(n_layers, sizes, weights, biases, G, dG, n_rows, ghat, e, dEdx, force, flag, (cudaStream_t)stream);
    if (r < 0) return set_err(NNMD_ERR_CAPACITY, "fused energy+force: shape %d-%d-%d-1 has no tensor-core instantiation", sizes[0], sizes[1], sizes[2]);
    return r;
}
