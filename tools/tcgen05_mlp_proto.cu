// tcgen05_mlp_proto.cu -- prototype: the first contraction of the atomic network (K6), H = sigmoid(X W0^T + b0), on the
// 5th-generation tensor cores of sm_100a: tcgen05.mma kind::tf32 issued by one thread, operands in shared memory in the
// canonical K-major "interleave" (no-swizzle) layout, accumulator in tensor memory (TMEM), read back with tcgen05.ld.
//
// Why a prototype and not the product kernel: the product K6 (csrc/mlp_tc.cu) keeps all four
// contractions of the network (two forward, two backward) in registers with warp-level mma.sync and never touches shared
// memory between layers; a UMMA form has to round-trip every layer through TMEM -> registers -> shared memory (the next A
// operand).  This tool measures what ONE such contraction costs on the UMMA path, with the same 3xTF32 error compensation
// (a = a_hi + a_lo, b = b_hi + b_lo; a_lo b_hi + a_hi b_lo + a_hi b_hi) the parity bar of 1e-5 needs, so that the decision
// rests on numbers: it prints the time per million atoms, the achieved tensor throughput and the worst error against fp64.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/tcgen05_mlp_proto tools/tcgen05_mlp_proto.cu
//   tools/tcgen05_mlp_proto [rows]           (ncu --set full on it gives sm__pipe_tensor_* of the UMMA path)
//
// Tile: 128 atoms x 64 neurons x 96 descriptors (12 k-steps of 8), one CTA of 128 threads per SM, persistent.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int K0 = 96, H1 = 64, TM = 128;          // descriptors, neurons, atoms per tile
constexpr int KSTEPS = K0 / 8;                     // one tcgen05.mma kind::tf32 consumes K = 8
// canonical K-major interleave layout, in bytes: element (r, c) of an (rows x K0) fp32 operand sits at
//   (r / 8) * SBO + (c / 4) * LBO + (r % 8) * 16 + (c % 4) * 4        (8 x 16 B core matrices, stored contiguously)
constexpr int LBO = 128;                           // next core matrix along K
constexpr int SBO = (K0 / 4) * 128;                // next group of 8 rows
constexpr int A_BYTES = TM * K0 * 4, B_BYTES = H1 * K0 * 4;
constexpr int TMEM_COLS = 64;                      // fp32 accumulator 128 lanes x 64 columns

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO in 16 B units, version 1 (sm_100),
// layout type 0 = no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= uint64_t((saddr & 0x3ffff) >> 4);
    d |= uint64_t(LBO >> 4) << 16;
    d |= uint64_t(SBO >> 4) << 32;
    d |= uint64_t(1) << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c = f32 (1 << 4), a = b = tf32 (2 << 7, 2 << 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(H1 >> 3) << 17) | (uint32_t(TM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) umma_layer0_kernel(const float *__restrict__ X, const float *__restrict__ W0,
                                                             const float *__restrict__ b0, float *__restrict__ Hout,
                                                             long long n_rows) {
    extern __shared__ __align__(128) unsigned char smem[];
    float *a_hi = reinterpret_cast<float *>(smem), *a_lo = reinterpret_cast<float *>(smem + A_BYTES);
    float *b_hi = reinterpret_cast<float *>(smem + 2 * A_BYTES), *b_lo = reinterpret_cast<float *>(smem + 2 * A_BYTES + B_BYTES);
    float *bias = reinterpret_cast<float *>(smem + 2 * A_BYTES + 2 * B_BYTES);
    uint64_t *bar = reinterpret_cast<uint64_t *>(bias + H1);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time: weights (hi / lo) in the canonical layout, bias, barrier, TMEM
    for (int e = tid; e < H1 * (K0 / 4); e += 128) {
        const int n = e / (K0 / 4), c4 = e % (K0 / 4);
        const float4 w = *reinterpret_cast<const float4 *>(W0 + size_t(n) * K0 + 4 * c4);
        const int off = ((n >> 3) * SBO + c4 * LBO + (n & 7) * 16) >> 2;
        const float4 h = make_float4(tf32_hi(w.x), tf32_hi(w.y), tf32_hi(w.z), tf32_hi(w.w));
        *reinterpret_cast<float4 *>(b_hi + off) = h;
        *reinterpret_cast<float4 *>(b_lo + off) = make_float4(w.x - h.x, w.y - h.y, w.z - h.z, w.w - h.w);
    }
    if (tid < H1) bias[tid] = b0[tid];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);

    const long long n_tiles = (n_rows + TM - 1) / TM;
    uint32_t parity = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- stage the 128 x 96 tile of X, split into hi / lo, coalesced 16 B loads
        const long long r0 = tile * TM;
        for (int e = tid; e < TM * (K0 / 4); e += 128) {
            const int r = e / (K0 / 4), c4 = e % (K0 / 4);
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < n_rows) x = __ldg(reinterpret_cast<const float4 *>(X + (r0 + r) * K0) + c4);
            const int off = ((r >> 3) * SBO + c4 * LBO + (r & 7) * 16) >> 2;
            const float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
            *reinterpret_cast<float4 *>(a_hi + off) = h;
            *reinterpret_cast<float4 *>(a_lo + off) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        // ---- one thread issues 12 k-steps x 3 MMAs (3xTF32) into the same TMEM accumulator
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int k = 0; k < KSTEPS; ++k) {
                const uint32_t ko = k * 2 * LBO;  // K = 8 tf32 = two core matrices along K
                umma_tf32(tmem, make_desc(a_lo_s + ko), make_desc(b_hi_s + ko), k > 0);
                umma_tf32(tmem, make_desc(a_hi_s + ko), make_desc(b_lo_s + ko), 1);
                umma_tf32(tmem, make_desc(a_hi_s + ko), make_desc(b_hi_s + ko), 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: thread t owns row 32 * warp + lane of the accumulator: 64 columns in four chunks of 16
        const long long row = r0 + tid;
#pragma unroll
        for (int c0 = 0; c0 < H1; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem + (uint32_t(32 * warp) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < n_rows) {
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                    float4 o;
                    o.x = 1.f / (1.f + expf(-(__uint_as_float(v[q]) + bias[c0 + q])));
                    o.y = 1.f / (1.f + expf(-(__uint_as_float(v[q + 1]) + bias[c0 + q + 1])));
                    o.z = 1.f / (1.f + expf(-(__uint_as_float(v[q + 2]) + bias[c0 + q + 2])));
                    o.w = 1.f / (1.f + expf(-(__uint_as_float(v[q + 3]) + bias[c0 + q + 3])));
                    *reinterpret_cast<float4 *>(Hout + row * H1 + c0 + q) = o;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // every thread has drained its TMEM rows and its smem reads before the next tile overwrites them
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

int main(int argc, char **argv) {
    const long long n_rows = argc > 1 ? atoll(argv[1]) : 1000188;
    std::vector<float> hX(size_t(n_rows) * K0), hW(size_t(H1) * K0), hb(H1);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return float((s >> 11) * (1.0 / 9007199254740992.0)); };
    for (long long r = 0; r < n_rows; ++r) {  // unit-norm rows like the normalised descriptors
        double nrm = 0;
        for (int k = 0; k < K0; ++k) { hX[r * K0 + k] = rnd(); nrm += double(hX[r * K0 + k]) * hX[r * K0 + k]; }
        for (int k = 0; k < K0; ++k) hX[r * K0 + k] = float(hX[r * K0 + k] / sqrt(nrm));
    }
    const float bound = sqrtf(6.f / (K0 + H1));  // Xavier uniform like bpnn.py:149-157
    for (auto &w : hW) w = (2.f * rnd() - 1.f) * bound;
    for (auto &b : hb) b = 0.1f * (2.f * rnd() - 1.f);
    float *dX, *dW, *db, *dH;
    CHECK(cudaMalloc(&dX, hX.size() * 4)); CHECK(cudaMalloc(&dW, hW.size() * 4)); CHECK(cudaMalloc(&db, hb.size() * 4));
    CHECK(cudaMalloc(&dH, size_t(n_rows) * H1 * 4));
    CHECK(cudaMemcpy(dX, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    int sms = 148;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t smem = 2 * A_BYTES + 2 * B_BYTES + H1 * 4 + 64;
    CHECK(cudaFuncSetAttribute(umma_layer0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const long long n_tiles = (n_rows + TM - 1) / TM;
    const unsigned blocks = unsigned(n_tiles < sms ? n_tiles : sms);
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    for (int w = 0; w < 3; ++w) umma_layer0_kernel<<<blocks, 128, smem>>>(dX, dW, db, dH, n_rows);
    CHECK(cudaDeviceSynchronize());
    const int reps = 10;
    CHECK(cudaEventRecord(a));
    for (int w = 0; w < reps; ++w) umma_layer0_kernel<<<blocks, 128, smem>>>(dX, dW, db, dH, n_rows);
    CHECK(cudaEventRecord(b));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, a, b));
    ms /= reps;
    std::vector<float> hH(size_t(n_rows) * H1);
    CHECK(cudaMemcpy(hH.data(), dH, hH.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, scale = 0;
    for (long long r = 0; r < n_rows; r += (n_rows > 4096 ? n_rows / 4096 : 1)) {
        for (int n = 0; n < H1; ++n) {
            double z = hb[n];
            for (int k = 0; k < K0; ++k) z += double(hX[r * K0 + k]) * double(hW[size_t(n) * K0 + k]);
            const double ref = 1.0 / (1.0 + exp(-z));
            worst = fmax(worst, fabs(ref - double(hH[r * H1 + n])));
            scale = fmax(scale, fabs(ref));
        }
    }
    const double flops = 2.0 * double(n_rows) * K0 * H1;  // algorithmic; the tensor pipe does 3x that (error compensation)
    printf("{\"kernel\": \"umma_layer0 (tcgen05.mma kind::tf32, 3xTF32, TMEM accumulator)\", \"rows\": %lld, \"ms\": %.4f, "
           "\"ms_per_million_rows\": %.4f, \"algorithmic_tflops\": %.2f, \"tensor_pipe_tflops\": %.2f, "
           "\"hbm_gbs\": %.1f, \"max_abs_err_vs_fp64\": %.3e, \"rel_to_max\": %.3e}\n",
           n_rows, ms, ms * 1e6 / double(n_rows), flops / (ms * 1e-3) / 1e12, 3 * flops / (ms * 1e-3) / 1e12,
           double(n_rows) * (K0 + H1) * 4 / (ms * 1e-3) / 1e9, worst, worst / scale);
    return worst / scale < 1e-5 ? 0 : 2;
}
