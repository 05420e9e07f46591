// loss.cu -- K8: the training loss of the Behler-Parrinello step and its gradient in ONE launch.
//
// Replaces, for nnmd/nn/bpnn.py:196-215 and :357-366,
//     loss = e_coeff * MSELoss(e.sum(1), E) + f_coeff / 3 * MSELoss(f, F);   loss.backward() down to e and f
// i.e. ~35 elementwise / reduction launches of torch.autograd (the step is launch-bound) by one kernel that reads the
// per-atom energies and forces of every species once and writes dL/de, dL/df and the three reported scalars.
// Per-frame partial sums are reduced in a fixed order by the last block to finish: the result does not depend on the
// order in which blocks run.
#include "common.cuh"

namespace nnmd {

constexpr int LOSS_MAX_SPECIES = 8;
constexpr int LOSS_THREADS = 128;

template <typename T>
struct LossArgs {
    int n_species;
    int n_atoms[LOSS_MAX_SPECIES];
    int atom_off[LOSS_MAX_SPECIES];  // first atom of the species in the frame's force rows
    int n_total;
    const T *e[LOSS_MAX_SPECIES];
    const T *f[LOSS_MAX_SPECIES];
    T *ge[LOSS_MAX_SPECIES];
    T *gf[LOSS_MAX_SPECIES];
    const T *E, *F;
    int64_t n_frames;
    double e_coeff, f_coeff;
    double *partials;          // (n_frames, 2): squared energy error, summed squared force error
    unsigned int *counter;     // blocks finished (reset by the last one)
    T *out;                    // (E term, F term, total)
};

template <typename V> __device__ __forceinline__ V block_sum(V v, V *scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    V s = 0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) s += scratch[w];  // fixed order, every thread gets the sum
    return s;
}

template <typename T>
__global__ void __launch_bounds__(LOSS_THREADS) loss_mse_kernel(LossArgs<T> a) {
    __shared__ double scratch[LOSS_THREADS / 32];
    __shared__ bool last;
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    // frame energy
    double es = 0;
    for (int s = 0; s < a.n_species; ++s) {
        const T *e = a.e[s] + b * a.n_atoms[s];
        for (int i = tid; i < a.n_atoms[s]; i += LOSS_THREADS) es += double(e[i]);
    }
    es = block_sum<double>(es, scratch);
    const double de = es - double(a.E[b]);
    const T g_e = T(2.0 * a.e_coeff * de / double(a.n_frames));
    const double f_scale = 2.0 * (a.f_coeff / 3.0) / (double(a.n_frames) * double(a.n_total) * 3.0);
    double fs = 0;
    for (int s = 0; s < a.n_species; ++s) {
        const int n = a.n_atoms[s];
        T *ge = a.ge[s] + b * n;
        for (int i = tid; i < n; i += LOSS_THREADS) ge[i] = g_e;
        const T *f = a.f[s] + b * n * 3;
        const T *F = a.F + (b * a.n_total + a.atom_off[s]) * 3;
        T *gf = a.gf[s] + b * n * 3;
        for (int i = tid; i < 3 * n; i += LOSS_THREADS) {
            const double d = double(f[i]) - double(F[i]);
            fs += d * d;
            gf[i] = T(f_scale * d);
        }
    }
    fs = block_sum<double>(fs, scratch);
    if (tid == 0) {
        a.partials[2 * b] = de * de;
        a.partials[2 * b + 1] = fs;
        __threadfence();
        last = atomicAdd(a.counter, 1u) == unsigned(gridDim.x) - 1u;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double se = 0, sf = 0;
    for (int64_t i = tid; i < a.n_frames; i += LOSS_THREADS) {
        se += __ldcg(a.partials + 2 * i);
        sf += __ldcg(a.partials + 2 * i + 1);
    }
    se = block_sum<double>(se, scratch);
    sf = block_sum<double>(sf, scratch);
    if (tid == 0) {
        const double le = a.e_coeff * se / double(a.n_frames);
        const double lf = (a.f_coeff / 3.0) * sf / (double(a.n_frames) * double(a.n_total) * 3.0);
        a.out[0] = T(le); a.out[1] = T(lf); a.out[2] = T(le + lf);
        *a.counter = 0u;
    }
}

template <typename T>
static int loss_run(int n_species, const int32_t *n_atoms, const void *const *e, const void *const *f, const void *E, const void *F,
                    int64_t n_frames, double e_coeff, double f_coeff, void *const *ge, void *const *gf, void *ws, void *out,
                    cudaStream_t st) {
    LossArgs<T> a;
    a.n_species = n_species;
    int off = 0;
    for (int s = 0; s < n_species; ++s) {
        a.n_atoms[s] = n_atoms[s]; a.atom_off[s] = off; off += n_atoms[s];
        a.e[s] = (const T *)e[s]; a.f[s] = (const T *)f[s]; a.ge[s] = (T *)ge[s]; a.gf[s] = (T *)gf[s];
    }
    a.n_total = off;
    a.E = (const T *)E; a.F = (const T *)F; a.n_frames = n_frames; a.e_coeff = e_coeff; a.f_coeff = f_coeff;
    a.counter = reinterpret_cast<unsigned int *>(ws);
    a.partials = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(ws) + 16);
    a.out = (T *)out;
    ProfScope prof(PROF_CONTRACT, st);
    loss_mse_kernel<T><<<unsigned(n_frames), LOSS_THREADS, 0, st>>>(a);
    NNMD_LAUNCH_CHECK("loss_mse_kernel");
    return NNMD_OK;
}

}  // namespace nnmd

using namespace nnmd;

extern "C" size_t nnmd_loss_workspace_bytes(int64_t n_frames) { return 16 + size_t(n_frames > 0 ? n_frames : 0) * 2 * sizeof(double); }

extern "C" int nnmd_loss_mse(int dtype, int n_species, const int32_t *n_atoms, const void *const *e, const void *const *f,
                             const void *E, const void *F, int64_t n_frames, double e_coeff, double f_coeff, void *const *ge,
                             void *const *gf, void *ws, void *out, void *stream) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "loss: bad dtype %d", dtype);
    if (n_species < 1 || n_species > LOSS_MAX_SPECIES) return set_err(NNMD_ERR_INVALID, "loss: 1..%d species supported", LOSS_MAX_SPECIES);
    if (n_frames < 1 || n_frames > 0x7fffffffLL) return set_err(NNMD_ERR_INVALID, "loss: bad frame count");
    if (!n_atoms || !e || !f || !E || !F || !ge || !gf || !ws || !out) return set_err(NNMD_ERR_INVALID, "loss: null pointer");
    for (int s = 0; s < n_species; ++s) {
        if (n_atoms[s] < 1) return set_err(NNMD_ERR_INVALID, "loss: species %d has no atoms", s);
        if (!e[s] || !f[s] || !ge[s] || !gf[s]) return set_err(NNMD_ERR_INVALID, "loss: null pointer for species %d", s);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NNMD_F32) return loss_run<float>(n_species, n_atoms, e, f, E, F, n_frames, e_coeff, f_coeff, ge, gf, ws, out, st);
    return loss_run<double>(n_species, n_atoms, e, f, E, F, n_frames, e_coeff, f_coeff, ge, gf, ws, out, st);
}

// ---------------------------------------------------------------------------------------------------------------
// Batch assembly: rows `sel` of several arrays (descriptors, descriptor gradients, energies, forces of the cached
// training set) are copied into the static buffers of the training step by ONE launch instead of one index_select
// per array (reference: the DataLoader collation of bpnn.py:234-236).
namespace nnmd {

constexpr int GATHER_MAX_ARRAYS = 20;

struct GatherArgs {
    int n_arrays;
    const unsigned char *src[GATHER_MAX_ARRAYS];
    unsigned char *dst[GATHER_MAX_ARRAYS];
    long long row_bytes[GATHER_MAX_ARRAYS];
    const long long *sel;
    const long long *cursor;  // optional (device): the rows are sel[*cursor + k], the position inside an epoch's permutation
};

__global__ void __launch_bounds__(256) gather_rows_kernel(GatherArgs a) {
    const long long row = a.sel[(a.cursor ? *a.cursor : 0) + blockIdx.x];
    for (int k = 0; k < a.n_arrays; ++k) {
        const long long nb = a.row_bytes[k];
        const unsigned char *s = a.src[k] + row * nb;
        unsigned char *d = a.dst[k] + (long long)blockIdx.x * nb;
        if (((nb | (long long)(size_t)s | (long long)(size_t)d) & 15) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
            uint4 *d4 = reinterpret_cast<uint4 *>(d);
            for (long long i = threadIdx.x; i < nb / 16; i += blockDim.x) d4[i] = __ldg(s4 + i);
        } else {
            const unsigned int *s1 = reinterpret_cast<const unsigned int *>(s);
            unsigned int *d1 = reinterpret_cast<unsigned int *>(d);
            for (long long i = threadIdx.x; i < nb / 4; i += blockDim.x) d1[i] = __ldg(s1 + i);
        }
    }
}

__global__ void advance_cursor_kernel(long long *cursor, long long by) { *cursor += by; }

}  // namespace nnmd

static int gather_rows_impl(int n_arrays, const void *const *src, void *const *dst, const int64_t *row_bytes,
                            const int64_t *sel, int64_t *cursor, int64_t n_sel, void *stream);

extern "C" int nnmd_gather_rows(int n_arrays, const void *const *src, void *const *dst, const int64_t *row_bytes,
                                const int64_t *sel, int64_t n_sel, void *stream) {
    return gather_rows_impl(n_arrays, src, dst, row_bytes, sel, nullptr, n_sel, stream);
}

// The batch loader of a whole epoch without the host: `perm` holds the epoch's frame order on the device, `cursor` (device
// int64) the position of the next batch; the call gathers rows perm[*cursor .. *cursor + n_sel) and advances the cursor, so a
// captured training step (gather included) is replayed once per batch with no per-batch arguments.
// (reference: the DataLoader iteration of nnmd/nn/bpnn.py:321-329)
extern "C" int nnmd_gather_rows_cursor(int n_arrays, const void *const *src, void *const *dst, const int64_t *row_bytes,
                                       const int64_t *perm, int64_t *cursor, int64_t n_sel, void *stream) {
    if (!cursor) return set_err(NNMD_ERR_INVALID, "gather: null cursor");
    int r = gather_rows_impl(n_arrays, src, dst, row_bytes, perm, cursor, n_sel, stream);
    if (r != NNMD_OK || n_sel == 0) return r;
    advance_cursor_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long *)cursor, (long long)n_sel);
    NNMD_LAUNCH_CHECK("advance_cursor_kernel");
    return NNMD_OK;
}

static int gather_rows_impl(int n_arrays, const void *const *src, void *const *dst, const int64_t *row_bytes,
                            const int64_t *sel, int64_t *cursor, int64_t n_sel, void *stream) {
    if (n_arrays < 1 || n_arrays > GATHER_MAX_ARRAYS) return set_err(NNMD_ERR_INVALID, "gather: 1..%d arrays supported", GATHER_MAX_ARRAYS);
    if (n_sel < 0 || n_sel > 0x7fffffffLL) return set_err(NNMD_ERR_INVALID, "gather: bad row count");
    if (n_sel == 0) return NNMD_OK;
    if (!src || !dst || !row_bytes || !sel) return set_err(NNMD_ERR_INVALID, "gather: null pointer");
    GatherArgs a;
    a.n_arrays = n_arrays;
    for (int k = 0; k < n_arrays; ++k) {
        if (!src[k] || !dst[k]) return set_err(NNMD_ERR_INVALID, "gather: null pointer for array %d", k);
        if (row_bytes[k] <= 0 || row_bytes[k] % 4) return set_err(NNMD_ERR_INVALID, "gather: row size of array %d must be a positive multiple of 4 bytes", k);
        a.src[k] = (const unsigned char *)src[k]; a.dst[k] = (unsigned char *)dst[k]; a.row_bytes[k] = row_bytes[k];
    }
    a.sel = (const long long *)sel;
    a.cursor = (const long long *)cursor;
    cudaStream_t st = (cudaStream_t)stream;
    gather_rows_kernel<<<unsigned(n_sel), 256, 0, st>>>(a);
    NNMD_LAUNCH_CHECK("gather_rows_kernel");
    return NNMD_OK;
}
