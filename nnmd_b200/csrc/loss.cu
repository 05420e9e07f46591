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
