// sf.cu -- K2 radial pair kernel, K3 angular triplet kernel, K4a summed-gradient output,
//          K4b fused force, K4c force contraction, K5 normalisation (sm_100a).
//
// What is computed (reference: nnmd/features/symm_funcs.py:107-277, calculate_sf.py:85-104; math as
// implemented = SURVEY.md Appendix A):
//   fc(d)  = 1/2 (cos(pi d / rc) + 1),  d < rc
//   G1_a   = sum_j fc(d_aj)                       G2_a = sum_j exp(-eta (d_aj - rs)^2) fc(d_aj)
//   G4_a   = 2^(1-zeta) sum_{j != k} P(c_ajk) exp(-eta (x^2 + y^2 + z^2)) fc(x) fc(y) fc(z)
//   G5_a   = 2^(1-zeta) sum_{j != k} P(c_ajk) exp(-eta (x^2 + y^2))       fc(x) fc(y) fc(z)
//            x = d_aj, y = d_ak, z = d_jk, ALL three inside (0, rc)  
//            P(c) = (1 + lambda |c|)^zeta if |c| > 1e-3 else 0        
//   dG[a,f,:] = d (sum_i G_i,f) / d r_a                               
//
// Because every contributing triplet is a closed triangle, both G_a and dG[a] are gathered from
// atom a's own neighbour row: no atomics, fixed summation order.
//
// K3 work decomposition (one warp per centre atom, one launch per (kind, rc, eta) group):
//   phase 0  lanes over the neighbour row: u, d, 1/d, fc, fc', psi = exp(-eta d^2) fc, psi' -> shared memory
//   phase 1  lanes over closed triangles (j<k, z<rc): everything that does not depend on the function
//            (three cosines, their derivatives, the radial products, log2 of the two bases 1+-|c|)
//            is reduced to a 38-real record in shared memory
//   phase 2  lanes over FUNCTIONS (G lanes x FPL functions per lane, 32/G triangles in flight):
//            per centre one ex2 and five FMA-pipe ops, six more per triangle for the direction vectors
// The inner loop is bound by the FP32 FMA pipe and the SFU (ex2), not by HBM: see DESIGN.md.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <new>
#include <tuple>

#include "sf_plan.h"

namespace nnmd {

// =============================================================================================
// plan
// =============================================================================================
template <typename T>
static int upload(const std::vector<double> &h, void **dptr) {
    std::vector<T> tmp(h.size());
    for (size_t i = 0; i < h.size(); ++i) tmp[i] = T(h[i]);
    NNMD_CUDA_CHECK(cudaMalloc(dptr, std::max<size_t>(tmp.size(), 1) * sizeof(T)));
    if (!tmp.empty()) NNMD_CUDA_CHECK(cudaMemcpy(*dptr, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
    return NNMD_OK;
}
static int upload_int(const std::vector<int> &h, int **dptr) {
    NNMD_CUDA_CHECK(cudaMalloc((void **)dptr, std::max<size_t>(h.size(), 1) * sizeof(int)));
    if (!h.empty()) NNMD_CUDA_CHECK(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
    return NNMD_OK;
}
static int upload_real(int dtype, const std::vector<double> &h, void **dptr) {
    return dtype == NNMD_F32 ? upload<float>(h, dptr) : upload<double>(h, dptr);
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

static int build_group(int dtype, int kind, double rc, double eta, const std::vector<int> &members,
                       const double *params, std::vector<AngGroup> &out) {
    // members: function indices sharing (kind, rc, eta).  Split so that one launch holds <= 32 lanes x 4.
    bool pm1 = true;
    for (int f : members) {
        const double lam = params[4 * f + 3];
        if (!(lam == 1.0 || lam == -1.0)) pm1 = false;
    }
    // order: lambda class (+ first), then zeta -- lanes of one class share the record block they read
    std::vector<int> ord(members);
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
        const double la = params[4 * a + 3], lb = params[4 * b + 3];
        if (pm1 && la != lb) return la > lb;
        return params[4 * a + 2] < params[4 * b + 2];
    });
    size_t pos = 0;
    while (pos < ord.size()) {
        const size_t chunk = std::min<size_t>(ord.size() - pos, 128);
        std::vector<int> part(ord.begin() + pos, ord.begin() + pos + chunk);
        pos += chunk;
        // classes inside this part
        std::vector<int> plus, minus;
        for (int f : part) ((pm1 && params[4 * f + 3] < 0) ? minus : plus).push_back(f);
        int fpl = 1;
        auto slots_for = [&](int l) { return int((plus.size() + l - 1) / l + (minus.size() + l - 1) / l); };
        while (fpl < ANG_MAX_FPL && slots_for(fpl) > 32) fpl <<= 1;
        // fp32, lambda = +-1, 33..64 functions: 4 per lane on 16 lanes, so the two half-warps of the edge-centric
        // kernel work on two records at once (fewer shared-memory wavefronts and SFU ops per record)
        if (dtype == NNMD_F32 && pm1 && fpl == 2 && slots_for(2) > 16 && slots_for(4) <= 16) fpl = 4;
        if (const char *env = tuning_env("NNMD_ANG_FPL")) {  // tuning knob: force the number of functions per lane
            const int want = atoi(env);
            if (want == 1 || want == 2 || want == 4) {
                fpl = want;
                while (fpl < ANG_MAX_FPL && slots_for(fpl) > 32) fpl <<= 1;
            }
        }
        if (slots_for(fpl) > 32) return set_err(NNMD_ERR_INVALID, "internal: angular group does not fit a warp");
        AngGroup g;
        g.kind = kind; g.rc = rc; g.eta = eta; g.pm1 = pm1 ? 1 : 0; g.fpl = fpl;
        g.lanes = next_pow2(std::max(slots_for(fpl), 1));
        g.n_slots = g.lanes * fpl;
        std::vector<double> zeta(g.n_slots, 1.0), lam(g.n_slots, 1.0), coef(g.n_slots, 0.0);
        std::vector<int> col(g.n_slots, -1);
        int slot = 0;
        auto place = [&](const std::vector<int> &cls, double lam_pad) {
            const int lanes_cls = int((cls.size() + fpl - 1) / fpl);
            for (int l = 0; l < lanes_cls; ++l)
                for (int i = 0; i < fpl; ++i, ++slot) {
                    const size_t k = size_t(l) * fpl + i;
                    lam[slot] = lam_pad;
                    if (k < cls.size()) {
                        const int f = cls[k];
                        zeta[slot] = params[4 * f + 2];
                        lam[slot] = params[4 * f + 3];
                        // 2^(1-zeta), times 2 because the reference sums ORDERED (j,k) (symm_funcs.py:216-219)
                        coef[slot] = 2.0 * std::pow(2.0, 1.0 - params[4 * f + 2]);
                        col[slot] = f;
                    }
                }
        };
        place(plus, 1.0);
        place(minus, -1.0);
        // zeta ladder (what calculate_params emits): when every lane's two functions are separated by the same
        // step (to within the 6-decimal rounding of the table, <= 2e-6), the second power is the first one times
        // 2^(step * L), and 2^(step * L) does not depend on the lane: it moves into the triangle record.
        const char *no_ladder = tuning_env("NNMD_NO_LADDER");  // tuning knob
        if (pm1 && fpl >= 2 && !(no_ladder && no_ladder[0] == '1')) {
            double sum = 0, lo = 1e300, hi = -1e300;
            int n = 0;
            for (int l = 0; l < g.lanes; ++l)
                for (int i = 0; i + 1 < fpl; ++i)
                    if (col[fpl * l + i] >= 0 && col[fpl * l + i + 1] >= 0) {
                        const double d = zeta[fpl * l + i + 1] - zeta[fpl * l + i];
                        sum += d; lo = std::min(lo, d); hi = std::max(hi, d); ++n;
                    }
            if (n > 0 && hi - lo <= 2e-6) { g.ladder = 1; g.ladder_step = sum / n; }
        }
        const char *no_edge = tuning_env("NNMD_NO_EDGE");  // tuning knob
        g.edge_path = (dtype == NNMD_F32 && pm1 && fpl == 4 && g.lanes == 16 && !(no_edge && no_edge[0] == '1')) ? 1 : 0;
        int r;
        if ((r = upload_real(dtype, zeta, &g.d_zeta))) return r;
        if ((r = upload_real(dtype, lam, &g.d_lam))) return r;
        if ((r = upload_real(dtype, coef, &g.d_coef))) return r;
        if ((r = upload_int(col, &g.d_col))) return r;
        g.cols = col;
        out.push_back(g);
    }
    return NNMD_OK;
}

template <typename T>
__device__ __forceinline__ void st4(T *p, T a, T b, T c, T d);
template <> __device__ __forceinline__ void st4<float>(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void st4<double>(double *p, double a, double b, double c, double d) {
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(c, d);
}
template <typename T> struct Q4 { T a, b, c, d; };
template <typename T> __device__ __forceinline__ Q4<T> ld4(const T *p);
template <> __device__ __forceinline__ Q4<float> ld4<float>(const float *p) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    return {v.x, v.y, v.z, v.w};
}
template <> __device__ __forceinline__ Q4<double> ld4<double>(const double *p) {
    const double2 u = *reinterpret_cast<const double2 *>(p), v = *reinterpret_cast<const double2 *>(p + 2);
    return {u.x, u.y, v.x, v.y};
}

// =============================================================================================
// shared-memory staging of one neighbour row (phase 0, used by K2 and K3)
// =============================================================================================
// Per neighbour: u4 = (ux, uy, uz, d) as one 16 B (fp32) / 32 B (fp64) entry, then 1/d, fc, fc', psi, psi' as SoA.
template <typename T>
struct NbView {
    T *u4, *invd, *fc, *dfc, *psi, *dpsi;
    int *epos;  // edge slot of the neighbour inside the centre's owned-edge range, or -1 when the neighbour owns the edge
    __device__ __forceinline__ void bind(T *base, int cap) {
        u4 = base; invd = u4 + 4 * cap; fc = invd + cap; dfc = fc + cap; psi = dfc + cap; dpsi = psi + cap;
        epos = reinterpret_cast<int *>(dpsi + cap);
    }
};

// Gathers the neighbours of a centre that lie inside rc into shared memory; returns their number.
// FULL: also stage fc, fc', psi = exp(-eta d^2) fc, psi' for the given (rc, eta).
// first_gt >= 0: also record each kept neighbour's edge slot (row position - first_gt; negative = owned by the neighbour).
// first_gt == STAGE_SPECIES: record each kept neighbour's species index (posw.w) in the same array instead (multi-species plans).
constexpr int STAGE_SPECIES = -2;
template <typename T, bool FULL>
__device__ __forceinline__ int stage_row(const T *__restrict__ posw, const int32_t *__restrict__ idx, int64_t o0,
                                         int64_t o1, T xa, T ya, T za, T rc, T inv_rc, T eta, NbView<T> nb, int cap,
                                         int first_gt = -1) {
    const int lane = lane_id();
    int nn = 0;
    for (int64_t base = o0; base < o1; base += 32) {
        const int64_t e = base + lane;
        bool keep = false;
        T ux = 0, uy = 0, uz = 0, d = 0, wj = 0;
        if (e < o1) {
            const int j = __ldg(idx + e);
            T xj, yj, zj;
            load_pos4<T>(posw, j, xj, yj, zj, wj);
            ux = xj - xa; uy = yj - ya; uz = zj - za;
            d = norm3<T>(ux, uy, uz);
            keep = (d < rc) && (d > T(0));
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int p = nn + __popc(m & ((1u << lane) - 1u));
            if (p < cap) {
                st4<T>(nb.u4 + 4 * p, ux, uy, uz, d);
                nb.invd[p] = M<T>::rcp(d);
                if (first_gt >= 0) nb.epos[p] = int(e - o0) - first_gt;
                else if (first_gt == STAGE_SPECIES) nb.epos[p] = int(wj);
                if (FULL) {
                    T s, c;
                    M<T>::sincospi_(d * inv_rc, &s, &c);
                    const T fc = T(0.5) * (c + T(1));
                    const T dfc = T(-0.5) * M<T>::pi() * inv_rc * s;
                    const T ex = M<T>::exp_(-eta * d * d);
                    nb.fc[p] = fc; nb.dfc[p] = dfc;
                    nb.psi[p] = ex * fc;
                    nb.dpsi[p] = ex * (dfc - T(2) * eta * d * fc);
                }
            }
        }
        nn += __popc(m);
    }
    __syncwarp();
    return min(nn, cap);
}

// =============================================================================================
// K2 radial
// =============================================================================================
enum { MODE_G = 0, MODE_DG = 1, MODE_FORCE = 2 };

template <typename T>
struct RadArgs {
    const T *posw;
    int64_t n_rows, n_atoms, n_centres;
    const int64_t *offsets;
    const int32_t *idx;
    int cap;
    int n;  // radial functions
    T rc_max;
    const T *rc, *eta, *rs;
    const int *col;
    int nf;  // row stride of the outputs
    T *G, *dG;
    const T *w;
    T *force;
    uint32_t *flag;
    const T *wr;  // multi-species: weight of a neighbour of species s is wr[s]; NULL otherwise
};

// One warp per centre atom, lanes over radial functions, the staged row broadcast from shared memory.
// UNI: every radial function has the same cutoff, so fc and fc' are staged once per neighbour and the
// per-function work is one exponential and a handful of FMAs.
// MS (multi-species, extension without a reference counterpart: SURVEY Q9 / 8f-3): G_i = sum_j wr[s_j] phi(d_ij), and the
// summed gradient d(sum_i G_i)/d r_a collects (wr[s_j] + wr[s_a]) phi'(d_aj) from the pair (a, j).
template <typename T, int MODE, bool UNI, bool MS = false>
__global__ void __launch_bounds__(256) radial_kernel(RadArgs<T> a) {
    if (a.flag != nullptr && (*a.flag & NNMD_FLAG_OVERFLOW)) return;  // incomplete neighbour list (capacity-bound build)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id(), warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    T *base = reinterpret_cast<T *>(smem_raw) + size_t(warp) * NB_FIELDS * a.cap;
    NbView<T> nb;
    nb.bind(base, a.cap);
    const T inv_rc_max = M<T>::rcp(a.rc_max);
    bool bad = false;
    for (int64_t row = int64_t(blockIdx.x) * wpb + warp; row < a.n_rows; row += int64_t(gridDim.x) * wpb) {
        const int64_t frame = row / a.n_centres;
        const int64_t gi = frame * a.n_atoms + (row - frame * a.n_centres);
        T xa, ya, za, sa;
        load_pos4<T>(a.posw, gi, xa, ya, za, sa);
        const int nn = stage_row<T, UNI>(a.posw, a.idx, a.offsets[row], a.offsets[row + 1], xa, ya, za, a.rc_max,
                                         inv_rc_max, T(0), nb, a.cap, MS ? STAGE_SPECIES : -1);
        if (MS) {  // psi is not used by the radial functions: it carries the neighbour's weight
            for (int t = lane; t < nn; t += 32) nb.psi[t] = a.wr[nb.epos[t]];
            __syncwarp();
        }
        if (MODE != MODE_G) {  // fold the -2/d of the gradient into the staged vectors: u <- -2 u / d (d itself stays in .d)
            const T wa = MS ? a.wr[int(sa)] : T(1);
            for (int t = lane; t < nn; t += 32) {
                const Q4<T> u = ld4<T>(nb.u4 + 4 * t);
                const T sc = (MS ? -(wa + nb.psi[t]) : T(-2)) * nb.invd[t];
                st4<T>(nb.u4 + 4 * t, sc * u.a, sc * u.b, sc * u.c, u.d);
            }
            __syncwarp();
        }
        T fx = 0, fy = 0, fz = 0;  // MODE_FORCE partial sums
        // fp32, one cutoff, gradients wanted: TWO neighbours per step on packed FP32 (fma.rn.f32x2, sm_100) -- 10 packed
        // instructions, 2 ex2 and 6 shared-memory loads per pair instead of 28 instructions for two scalar steps.
        constexpr bool PACKED = sizeof(T) == 4 && UNI && !MS && MODE != MODE_G;
        if constexpr (PACKED) {
            // structure-of-arrays copies of the staged row in the fields the radial functions do not use:
            // psi <- ux', dpsi <- uy', invd <- uz' (u' = -2 u / d), epos <- d; padded to an even count with a null entry
            float *sx = reinterpret_cast<float *>(nb.psi), *sy = reinterpret_cast<float *>(nb.dpsi), *sz = reinterpret_cast<float *>(nb.invd);
            float *sd = reinterpret_cast<float *>(nb.epos), *sfc = reinterpret_cast<float *>(nb.fc), *sdfc = reinterpret_cast<float *>(nb.dfc);
            for (int t = lane; t < nn; t += 32) {
                const Q4<T> u = ld4<T>(nb.u4 + 4 * t);  // already scaled by -2 / d above
                sx[t] = u.a; sy[t] = u.b; sz[t] = u.c; sd[t] = u.d;
            }
            if ((nn & 1) && lane == 0) { sx[nn] = sy[nn] = sz[nn] = 0.f; sd[nn] = 0.f; sfc[nn] = 0.f; sdfc[nn] = 0.f; }
            __syncwarp();
            const int np = (nn + 1) >> 1;
            for (int f0 = 0; f0 < a.n; f0 += 32) {
                const int f = f0 + lane;
                const bool act = f < a.n;
                const float eta = act ? a.eta[f] : 0.f, rs = act ? a.rs[f] : 0.f;
                const float nel2 = -eta * 1.4426950408889634074f, m2eta = -2.f * eta;
                const float2 mrs = make_float2(-rs, -rs), k2 = make_float2(nel2, nel2), me2 = make_float2(m2eta, m2eta);
                float2 val = make_float2(0.f, 0.f), gx = val, gy = val, gz = val;
                if (act) {
#pragma unroll 2
                    for (int q = 0; q < np; ++q) {
                        const float2 d2 = *reinterpret_cast<const float2 *>(sd + 2 * q);
                        const float2 fc2 = *reinterpret_cast<const float2 *>(sfc + 2 * q);
                        const float2 dfc2 = *reinterpret_cast<const float2 *>(sdfc + 2 * q);
                        const float2 dr = __fadd2_rn(d2, mrs);
                        const float2 ee = __fmul2_rn(__fmul2_rn(dr, dr), k2);
                        const float2 ex = make_float2(M<float>::exp2_fast(ee.x), M<float>::exp2_fast(ee.y));  // G1: eta = 0 -> 1
                        val = __ffma2_rn(ex, fc2, val);
                        const float2 dphi = __fmul2_rn(ex, __ffma2_rn(__fmul2_rn(me2, dr), fc2, dfc2));
                        gx = __ffma2_rn(dphi, *reinterpret_cast<const float2 *>(sx + 2 * q), gx);
                        gy = __ffma2_rn(dphi, *reinterpret_cast<const float2 *>(sy + 2 * q), gy);
                        gz = __ffma2_rn(dphi, *reinterpret_cast<const float2 *>(sz + 2 * q), gz);
                    }
                    const float v = val.x + val.y, g0 = gx.x + gx.y, g1 = gy.x + gy.y, g2 = gz.x + gz.y;
                    const int col = a.col[f];
                    const int64_t o = row * a.nf + col;
                    if (MODE == MODE_DG) {
                        a.G[o] = v;
                        a.dG[3 * o] = g0; a.dG[3 * o + 1] = g1; a.dG[3 * o + 2] = g2;
                        bad |= !(isfinite(g0) && isfinite(g1) && isfinite(g2));
                    }
                    if (MODE == MODE_FORCE) {
                        const float wv = a.w[o];
                        fx -= wv * g0; fy -= wv * g1; fz -= wv * g2;
                    }
                }
            }
        } else
        for (int f0 = 0; f0 < a.n; f0 += 32) {
            const int f = f0 + lane;
            const bool act = f < a.n;
            const T rc = act ? a.rc[f] : T(1), eta = act ? a.eta[f] : T(0), rs = act ? a.rs[f] : T(0);
            const T inv_rc = M<T>::rcp(rc);
            const T nel2 = -eta * T(1.4426950408889634074);  // exp(-eta x) = 2^(nel2 x)
            const T m2eta = T(-2) * eta;
            T val = 0, gx = 0, gy = 0, gz = 0;
            if (act) {
                for (int t = 0; t < nn; ++t) {
                    const Q4<T> u = ld4<T>(nb.u4 + 4 * t);
                    const T d = u.d;
                    T fc, dfc;
                    if (UNI) {
                        fc = nb.fc[t]; dfc = nb.dfc[t];
                    } else {
                        if (!(d < rc)) continue;
                        T s, c;
                        M<T>::sincospi_(d * inv_rc, &s, &c);
                        fc = T(0.5) * (c + T(1));
                        dfc = T(-0.5) * M<T>::pi() * inv_rc * s;
                    }
                    const T dr = d - rs;
                    const T ex = M<T>::exp2_fast(nel2 * dr * dr);  // G1: eta = 0 -> 1
                    val = MS ? M<T>::fma_(ex * fc, nb.psi[t], val) : M<T>::fma_(ex, fc, val);
                    if (MODE != MODE_G) {
                        // own-atom gradient plus the identical term inside every neighbour's G_j: the staged
                        // vector already carries -2/d
                        const T dphi = ex * M<T>::fma_(m2eta * dr, fc, dfc);
                        gx = M<T>::fma_(dphi, u.a, gx);
                        gy = M<T>::fma_(dphi, u.b, gy);
                        gz = M<T>::fma_(dphi, u.c, gz);
                    }
                }
                const int col = a.col[f];
                const int64_t o = row * a.nf + col;
                if (MODE == MODE_G || MODE == MODE_DG) a.G[o] = val;
                if (MODE == MODE_DG) {
                    a.dG[3 * o] = gx; a.dG[3 * o + 1] = gy; a.dG[3 * o + 2] = gz;
                    bad |= !(isfinite(gx) && isfinite(gy) && isfinite(gz));
                }
                if (MODE == MODE_FORCE) {
                    const T wv = a.w[o];
                    fx -= wv * gx; fy -= wv * gy; fz -= wv * gz;
                }
            }
        }
        if (MODE == MODE_FORCE) {
            fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
            if (lane == 0) { a.force[3 * row] += fx; a.force[3 * row + 1] += fy; a.force[3 * row + 2] += fz; }
        }
        __syncwarp();
    }
    if (MODE == MODE_DG && bad) atomicOr(a.flag, NNMD_FLAG_NAN_DG);
}

// =============================================================================================
// K3 angular
// =============================================================================================
template <typename T>
struct AngArgs {
    const T *posw;
    int64_t n_rows, n_atoms, n_centres;
    const int64_t *offsets;
    const int32_t *idx;
    int cap;
    int kind, pm1, lanes, log_lanes, ladder;
    T rc, eta, step;
    const T *zeta, *lam, *coef;
    const int *col;
    int nf;
    T *G, *dG;
    const T *w;
    T *force;
    uint32_t *flag;
    const T *pw;  // multi-species: symmetric pair weights pw[s_j * ns + s_k] of the two neighbours of a centre; NULL otherwise
    int ns;
};

// Triangle record (REC_STRIDE reals in shared memory), one block of REC_CLASS per lambda class (+ then -):
//   [ L0  L1  L2  V0 | A0x A0y B0x B0y | A1x A1y B1x B1y | A2x A2y B2x B2y | R0 R1 R2 . ]   R_X = 2^(step L_X), ladder only
//   L_X = log2(1 +- |c_X|), V0 = (1 +- |c_0|) E_0, A_X = +-sgn(c_X) E_X dc_X/d(x,y), B_X = (1 +- |c_X|) dE_X/d(x,y)
// so that for a function (zeta, lambda = +-1), with q_X = 2^((zeta-1) L_X):
//   G += q_0 V0 ;  dS/dx = sum_X q_X (zeta A_Xx + B_Xx) ;  dS/dy likewise
// tail at REC_TAIL: [ ujx ukx ujy uky | ujz ukz . . ] unit vectors towards j and k, interleaved so that
// (dS/dx, dS/dy) * (uj_c, uk_c) is one packed FMA per component.
// General lambda: block 0 holds [ |c_0| |c_1| |c_2| E_0 | sgn E dc pairs, dE pairs ] and the lane finishes base, log, lambda.
constexpr int RL = 0, RV = 3, RA = 4, RR = 16;  // offsets inside a class block: L_X at RL+X, V0 at RV, centre X quad at RA+4X, R_X at RR+X

// cosine at a centre with adjacent sides (p, q), opposite side r (squares given), thresholds of symm_funcs.py:193-197
template <typename T>
__device__ __forceinline__ T cos_law(T p2, T q2, T r2, T p, T q, T ip, T iq, bool &live) {
    live = (T(2) * p * q) >= T(1e-3);
    return live ? (p2 + q2 - r2) * (T(0.5) * ip * iq) : T(0);
}

// phase 1: one lane reduces one closed triangle (tj < tk) to its record.
// fp32 uses the SFU approximations (rsqrt, cos, ex2, lg2: each a few ulp, see DESIGN.md "fp32 error budget");
// fp64 uses the accurate library routines.
template <typename T, int MODE>
__device__ __forceinline__ void make_record(const NbView<T> &nb, int tj, int tk, int kind, int pm1, int ladder, T step,
                                            T rc, T inv_rc, T eta, T *__restrict__ rec, const T *__restrict__ pw = nullptr,
                                            int ns = 0, int s_a = 0) {
    const Q4<T> uj = ld4<T>(nb.u4 + 4 * tj), uk = ld4<T>(nb.u4 + 4 * tk);
    const T x = uj.d, y = uk.d, ix = nb.invd[tj], iy = nb.invd[tk];
    const T wx = uj.a - uk.a, wy = uj.b - uk.b, wz = uj.c - uk.c;
    const T z2 = wx * wx + wy * wy + wz * wz;
    const T iz = M<T>::rsqrt_fast(z2), z = z2 * iz;
    const T x2 = x * x, y2 = y * y;
    const T fz = T(0.5) * (M<T>::cospi_fast(z * inv_rc) + T(1));
    const T psz = M<T>::exp_fast(-eta * z2) * fz;
    const T psx = nb.psi[tj], psy = nb.psi[tk], dpsx = nb.dpsi[tj], dpsy = nb.dpsi[tk];

    T c[3], dcx[3], dcy[3], E[3], Ex[3], Ey[3];
    bool live[3];
    c[0] = cos_law<T>(x2, y2, z2, x, y, ix, iy, live[0]);  // centre a : sides x, y   opposite z
    c[1] = cos_law<T>(x2, z2, y2, x, z, ix, iz, live[1]);  // centre j : sides x, z   opposite y
    c[2] = cos_law<T>(y2, z2, x2, y, z, iy, iz, live[2]);  // centre k : sides y, z   opposite x
    dcx[0] = live[0] ? iy - c[0] * ix : T(0);  dcy[0] = live[0] ? ix - c[0] * iy : T(0);
    dcx[1] = live[1] ? iz - c[1] * ix : T(0);  dcy[1] = live[1] ? -y * ix * iz : T(0);
    dcx[2] = live[2] ? -x * iy * iz : T(0);    dcy[2] = live[2] ? iz - c[2] * iy : T(0);
    if (kind == NNMD_G4) {
        const T e = psx * psy * psz;
        E[0] = E[1] = E[2] = e;
        Ex[0] = Ex[1] = Ex[2] = dpsx * psy * psz;
        Ey[0] = Ey[1] = Ey[2] = psx * dpsy * psz;
    } else {  // G5 keeps fc of the opposite side
        const T fcx = nb.fc[tj], fcy = nb.fc[tk], dfcx = nb.dfc[tj], dfcy = nb.dfc[tk];
        E[0] = psx * psy * fz;   Ex[0] = dpsx * psy * fz;   Ey[0] = psx * dpsy * fz;
        E[1] = psx * psz * fcy;  Ex[1] = dpsx * psz * fcy;  Ey[1] = psx * psz * dfcy;
        E[2] = psy * psz * fcx;  Ex[2] = psy * psz * dfcx;  Ey[2] = dpsy * psz * fcx;
    }
    if (pw != nullptr) {
        // multi-species: the term of a centre is weighted by the species pair of its two neighbours: centre a sees (j, k),
        // centre j sees (a, k), centre k sees (a, j)   (the staged epos array carries the species index)
        const int s_j = nb.epos[tj], s_k = nb.epos[tk];
        const T W[3] = {pw[s_j * ns + s_k], pw[s_a * ns + s_k], pw[s_a * ns + s_j]};
#pragma unroll
        for (int X = 0; X < 3; ++X) { E[X] *= W[X]; Ex[X] *= W[X]; Ey[X] *= W[X]; }
    }
    T blk[2][REC_CLASS];
    blk[0][RR + 3] = blk[1][RR + 3] = T(0);
#pragma unroll
    for (int X = 0; X < 3; ++X) {
        const T ac = M<T>::abs_(c[X]);
        const bool on = ac > T(1e-3);  // symm_funcs.py:210-211: the angular factor is 0 for |cos| <= 1e-3
        const T sg = c[X] > T(0) ? T(1) : T(-1);
        const T ax = on ? sg * E[X] * dcx[X] : T(0), ay = on ? sg * E[X] * dcy[X] : T(0);
        const T ex = on ? Ex[X] : T(0), ey = on ? Ey[X] : T(0);
        if (pm1) {
            const T bp = T(1) + ac, bm = M<T>::max_(T(1) - ac, T(0));
            blk[0][RL + X] = on ? M<T>::log2_fast(bp) : T(0);
            blk[1][RL + X] = on ? M<T>::max_(M<T>::log2_fast(bm), -M<T>::big()) : T(0);
            if (ladder) {
                blk[0][RR + X] = M<T>::exp2_fast(step * blk[0][RL + X]);
                blk[1][RR + X] = M<T>::exp2_fast(step * blk[1][RL + X]);
            }
            blk[0][RA + 4 * X + 0] = ax;       blk[0][RA + 4 * X + 1] = ay;
            blk[0][RA + 4 * X + 2] = bp * ex;  blk[0][RA + 4 * X + 3] = bp * ey;
            blk[1][RA + 4 * X + 0] = -ax;      blk[1][RA + 4 * X + 1] = -ay;
            blk[1][RA + 4 * X + 2] = bm * ex;  blk[1][RA + 4 * X + 3] = bm * ey;
            if (X == 0) { blk[0][RV] = on ? bp * E[0] : T(0); blk[1][RV] = on ? bm * E[0] : T(0); }
        } else {  // general lambda: the lane finishes base, log and the lambda factors itself
            blk[0][RL + X] = on ? ac : T(0);
            blk[0][RA + 4 * X + 0] = ax;  blk[0][RA + 4 * X + 1] = ay;
            blk[0][RA + 4 * X + 2] = ex;  blk[0][RA + 4 * X + 3] = ey;
            if (X == 0) blk[0][RV] = on ? E[0] : T(0);
        }
    }
    if (ladder) {
        st4<T>(rec + RR, blk[0][RR], blk[0][RR + 1], blk[0][RR + 2], T(0));
        st4<T>(rec + REC_CLASS + RR, blk[1][RR], blk[1][RR + 1], blk[1][RR + 2], T(0));
    }
    if (MODE == MODE_G) {
        st4<T>(rec, blk[0][0], blk[0][1], blk[0][2], blk[0][3]);
        if (pm1) st4<T>(rec + REC_CLASS, blk[1][0], blk[1][1], blk[1][2], blk[1][3]);
        return;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) st4<T>(rec + 4 * q, blk[0][4 * q], blk[0][4 * q + 1], blk[0][4 * q + 2], blk[0][4 * q + 3]);
    if (pm1) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            st4<T>(rec + REC_CLASS + 4 * q, blk[1][4 * q], blk[1][4 * q + 1], blk[1][4 * q + 2], blk[1][4 * q + 3]);
    }
    st4<T>(rec + REC_TAIL, uj.a * ix, uk.a * iy, uj.b * ix, uk.b * iy);
    st4<T>(rec + REC_TAIL + 4, uj.c * ix, uk.c * iy, T(0), T(0));
}

// Per-lane accumulators of phase 2.  g*[i] hold the (towards-j, towards-k) partial sums of one component.
template <typename T, int FPL>
struct AngAcc {
    T val[FPL];
    T gxa[FPL], gxb[FPL], gya[FPL], gyb[FPL], gza[FPL], gzb[FPL];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < FPL; ++i) val[i] = gxa[i] = gxb[i] = gya[i] = gyb[i] = gza[i] = gzb[i] = T(0);
    }
};

// phase 2, generic arithmetic (fp64, and fp32 with a general lambda)
template <typename T, int FPL, int MODE>
__device__ __forceinline__ void phase2_generic(const T *__restrict__ recs, int cnt, int grp, int ngrp, int cls_off, int pm1,
                                               const T (&zeta)[FPL], const T (&zm1)[FPL], const T (&lam)[FPL],
                                               AngAcc<T, FPL> &acc) {
    for (int t = grp; t < cnt; t += ngrp) {
        const T *r = recs + t * REC_STRIDE + cls_off;
        const Q4<T> l = ld4<T>(r);
        Q4<T> c0{0, 0, 0, 0}, c1{0, 0, 0, 0}, c2{0, 0, 0, 0}, ua{0, 0, 0, 0}, ub{0, 0, 0, 0};
        if (MODE != MODE_G) {
            c0 = ld4<T>(r + RA); c1 = ld4<T>(r + RA + 4); c2 = ld4<T>(r + RA + 8);
            ua = ld4<T>(recs + t * REC_STRIDE + REC_TAIL);
            ub = ld4<T>(recs + t * REC_STRIDE + REC_TAIL + 4);
        }
#pragma unroll
        for (int i = 0; i < FPL; ++i) {
            T L0 = l.a, L1 = l.b, L2 = l.c, V0 = l.d;
            T A0x = c0.a, A0y = c0.b, B0x = c0.c, B0y = c0.d;
            T A1x = c1.a, A1y = c1.b, B1x = c1.c, B1y = c1.d;
            T A2x = c2.a, A2y = c2.b, B2x = c2.c, B2y = c2.d;
            if (!pm1) {  // record holds |c|, sgn E dc, dE: finish base, log and the lambda factors here
                const T lm = lam[i];
                const T b0 = M<T>::max_(M<T>::fma_(lm, L0, T(1)), T(0));
                const T b1 = M<T>::max_(M<T>::fma_(lm, L1, T(1)), T(0));
                const T b2 = M<T>::max_(M<T>::fma_(lm, L2, T(1)), T(0));
                L0 = M<T>::max_(M<T>::log2_fast(b0), -M<T>::big());
                L1 = M<T>::max_(M<T>::log2_fast(b1), -M<T>::big());
                L2 = M<T>::max_(M<T>::log2_fast(b2), -M<T>::big());
                A0x *= lm; A0y *= lm; A1x *= lm; A1y *= lm; A2x *= lm; A2y *= lm;
                B0x *= b0; B0y *= b0; B1x *= b1; B1y *= b1; B2x *= b2; B2y *= b2;
                V0 *= b0;
            }
            const T q0 = M<T>::exp2_fast(zm1[i] * L0);
            acc.val[i] = M<T>::fma_(q0, V0, acc.val[i]);
            if (MODE != MODE_G) {
                const T q1 = M<T>::exp2_fast(zm1[i] * L1);
                const T q2 = M<T>::exp2_fast(zm1[i] * L2);
                const T zt = zeta[i];
                T sa = q0 * M<T>::fma_(zt, A0x, B0x);
                T sb = q0 * M<T>::fma_(zt, A0y, B0y);
                sa = M<T>::fma_(q1, M<T>::fma_(zt, A1x, B1x), sa);
                sb = M<T>::fma_(q1, M<T>::fma_(zt, A1y, B1y), sb);
                sa = M<T>::fma_(q2, M<T>::fma_(zt, A2x, B2x), sa);
                sb = M<T>::fma_(q2, M<T>::fma_(zt, A2y, B2y), sb);
                acc.gxa[i] = M<T>::fma_(sa, ua.a, acc.gxa[i]); acc.gxb[i] = M<T>::fma_(sb, ua.b, acc.gxb[i]);
                acc.gya[i] = M<T>::fma_(sa, ua.c, acc.gya[i]); acc.gyb[i] = M<T>::fma_(sb, ua.d, acc.gyb[i]);
                acc.gza[i] = M<T>::fma_(sa, ub.a, acc.gza[i]); acc.gzb[i] = M<T>::fma_(sb, ub.b, acc.gzb[i]);
            }
        }
    }
}

// phase 2, fp32 with lambda = +-1: packed FP32 (fma.rn.f32x2, sm_100) on everything that comes in (x,y) pairs.
// Per function and triangle: 2 issue slots for the three exponents, 3 ex2, 3 packed FMAs for zeta*A+B, 3 packed FMAs
// for q*(...), 1 for the value, 3 packed FMAs for the direction vectors: 15 issue slots / 22 FMA-pipe cycles
// (the scalar form needs 25 issue slots).
template <int FPL, int MODE>
__device__ __forceinline__ void phase2_packed(const float *__restrict__ recs, int cnt, int grp, int ngrp, int cls_off,
                                              const float (&zeta)[FPL], const float (&zm1)[FPL], AngAcc<float, FPL> &acc) {
    for (int t = grp; t < cnt; t += ngrp) {
        const float *r = recs + t * REC_STRIDE + cls_off;
        const float4 l = *reinterpret_cast<const float4 *>(r);
        float4 c0, c1, c2, ua, ub;
        if (MODE != MODE_G) {
            c0 = *reinterpret_cast<const float4 *>(r + RA);
            c1 = *reinterpret_cast<const float4 *>(r + RA + 4);
            c2 = *reinterpret_cast<const float4 *>(r + RA + 8);
            ua = *reinterpret_cast<const float4 *>(recs + t * REC_STRIDE + REC_TAIL);
            ub = *reinterpret_cast<const float4 *>(recs + t * REC_STRIDE + REC_TAIL + 4);
        }
#pragma unroll
        for (int i = 0; i < FPL; ++i) {
            if (MODE == MODE_G) {
                const float q0 = M<float>::exp2_fast(zm1[i] * l.x);
                acc.val[i] = fmaf(q0, l.w, acc.val[i]);
            } else {
                const float2 zm = make_float2(zm1[i], zm1[i]), zz = make_float2(zeta[i], zeta[i]);
                const float2 e01 = __fmul2_rn(zm, make_float2(l.x, l.y));
                const float q0 = M<float>::exp2_fast(e01.x), q1 = M<float>::exp2_fast(e01.y);
                const float q2 = M<float>::exp2_fast(zm1[i] * l.z);
                const float2 t0 = __ffma2_rn(zz, make_float2(c0.x, c0.y), make_float2(c0.z, c0.w));
                const float2 t1 = __ffma2_rn(zz, make_float2(c1.x, c1.y), make_float2(c1.z, c1.w));
                const float2 t2 = __ffma2_rn(zz, make_float2(c2.x, c2.y), make_float2(c2.z, c2.w));
                acc.val[i] = fmaf(q0, l.w, acc.val[i]);
                // FFMA2 takes a scalar register as a broadcast operand, so q_X * (t_Xx, t_Xy) is one instruction
                float2 s = __fmul2_rn(make_float2(q0, q0), t0);
                s = __ffma2_rn(make_float2(q1, q1), t1, s);
                s = __ffma2_rn(make_float2(q2, q2), t2, s);
                float2 gx = make_float2(acc.gxa[i], acc.gxb[i]), gy = make_float2(acc.gya[i], acc.gyb[i]),
                       gz = make_float2(acc.gza[i], acc.gzb[i]);
                gx = __ffma2_rn(s, make_float2(ua.x, ua.y), gx);
                gy = __ffma2_rn(s, make_float2(ua.z, ua.w), gy);
                gz = __ffma2_rn(s, make_float2(ub.x, ub.y), gz);
                acc.gxa[i] = gx.x; acc.gxb[i] = gx.y; acc.gya[i] = gy.x; acc.gyb[i] = gy.y; acc.gza[i] = gz.x; acc.gzb[i] = gz.y;
            }
        }
    }
}

#ifndef NNMD_ANG_MINB
#define NNMD_ANG_MINB 4  // resident CTAs per SM the register allocation must allow
#endif
// phase 2, fp32, lambda = +-1, two functions per lane on a zeta ladder: the lane's second power is the first one
// times the record's R_X = 2^(step L_X), so the SFU issues 3 ex2 per triangle instead of 6.
template <int MODE>
__device__ __forceinline__ void phase2_ladder(const float *__restrict__ recs, int cnt, int grp, int ngrp, int cls_off,
                                              const float (&zeta)[2], const float (&zm1)[2], AngAcc<float, 2> &acc) {
    for (int t = grp; t < cnt; t += ngrp) {
        const float *r = recs + t * REC_STRIDE + cls_off;
        const float4 l = *reinterpret_cast<const float4 *>(r);
        const float4 rr = *reinterpret_cast<const float4 *>(r + RR);
        const float2 zm = make_float2(zm1[0], zm1[0]);
        const float2 e01 = __fmul2_rn(zm, make_float2(l.x, l.y));
        float2 qa01, qb01;
        qa01.x = M<float>::exp2_fast(e01.x); qa01.y = M<float>::exp2_fast(e01.y);
        const float qa2 = M<float>::exp2_fast(zm1[0] * l.z);
        qb01 = __fmul2_rn(qa01, make_float2(rr.x, rr.y));
        const float qb2 = qa2 * rr.z;
        acc.val[0] = fmaf(qa01.x, l.w, acc.val[0]);
        acc.val[1] = fmaf(qb01.x, l.w, acc.val[1]);
        if (MODE != MODE_G) {
            const float4 c0 = *reinterpret_cast<const float4 *>(r + RA);
            const float4 c1 = *reinterpret_cast<const float4 *>(r + RA + 4);
            const float4 c2 = *reinterpret_cast<const float4 *>(r + RA + 8);
            const float4 ua = *reinterpret_cast<const float4 *>(recs + t * REC_STRIDE + REC_TAIL);
            const float4 ub = *reinterpret_cast<const float4 *>(recs + t * REC_STRIDE + REC_TAIL + 4);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float q0 = i ? qb01.x : qa01.x, q1 = i ? qb01.y : qa01.y, q2 = i ? qb2 : qa2;
                const float2 zz = make_float2(zeta[i], zeta[i]);
                const float2 t0 = __ffma2_rn(zz, make_float2(c0.x, c0.y), make_float2(c0.z, c0.w));
                const float2 t1 = __ffma2_rn(zz, make_float2(c1.x, c1.y), make_float2(c1.z, c1.w));
                const float2 t2 = __ffma2_rn(zz, make_float2(c2.x, c2.y), make_float2(c2.z, c2.w));
                float2 s = __fmul2_rn(make_float2(q0, q0), t0);
                s = __ffma2_rn(make_float2(q1, q1), t1, s);
                s = __ffma2_rn(make_float2(q2, q2), t2, s);
                float2 gx = make_float2(acc.gxa[i], acc.gxb[i]), gy = make_float2(acc.gya[i], acc.gyb[i]),
                       gz = make_float2(acc.gza[i], acc.gzb[i]);
                gx = __ffma2_rn(s, make_float2(ua.x, ua.y), gx);
                gy = __ffma2_rn(s, make_float2(ua.z, ua.w), gy);
                gz = __ffma2_rn(s, make_float2(ub.x, ub.y), gz);
                acc.gxa[i] = gx.x; acc.gxb[i] = gx.y; acc.gya[i] = gy.x; acc.gyb[i] = gy.y; acc.gza[i] = gz.x; acc.gzb[i] = gz.y;
            }
        }
    }
}

// fp64 values take two registers each: with the fp32 bound of 128 registers the double instantiations spilled 0.7-1.6 kB per
// thread (round-1 VERDICT); two resident blocks per SM give them 255 registers and no spills.
template <typename T, int FPL, int MODE>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? 2 : NNMD_ANG_MINB) angular_kernel(AngArgs<T> a) {
    if (a.flag != nullptr && (*a.flag & NNMD_FLAG_OVERFLOW)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id(), warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    // per-warp shared memory: records | neighbour fields | pair queue
    const size_t per_warp = size_t(ANG_BATCH) * REC_STRIDE * sizeof(T) + size_t(NB_FIELDS) * a.cap * sizeof(T) + 64 * sizeof(int);
    unsigned char *wbase = smem_raw + per_warp * warp;
    T *recs = reinterpret_cast<T *>(wbase);
    NbView<T> nb;
    nb.bind(recs + ANG_BATCH * REC_STRIDE, a.cap);
    int *queue = reinterpret_cast<int *>(recs + ANG_BATCH * REC_STRIDE + size_t(NB_FIELDS) * a.cap);

    const int G = a.lanes, grp = lane >> a.log_lanes, fl = lane & (G - 1), ngrp = 32 >> a.log_lanes;
    const T inv_rc = M<T>::rcp(a.rc), rc2 = a.rc * a.rc;
    // this lane's functions
    T zeta[FPL], zm1[FPL], lam[FPL], coef[FPL];
    int col[FPL];
#pragma unroll
    for (int i = 0; i < FPL; ++i) {
        const int s = fl * FPL + i;
        zeta[i] = a.zeta[s]; zm1[i] = zeta[i] - T(1); lam[i] = a.lam[s]; coef[i] = a.coef[s]; col[i] = a.col[s];
    }
    const int cls_off = (a.pm1 && lam[0] < T(0)) ? REC_CLASS : 0;
    bool bad = false;

    for (int64_t row = int64_t(blockIdx.x) * wpb + warp; row < a.n_rows; row += int64_t(gridDim.x) * wpb) {
        const int64_t frame = row / a.n_centres;
        const int64_t gi = frame * a.n_atoms + (row - frame * a.n_centres);
        T xa, ya, za, sa;
        load_pos4<T>(a.posw, gi, xa, ya, za, sa);
        const int nn = stage_row<T, true>(a.posw, a.idx, a.offsets[row], a.offsets[row + 1], xa, ya, za, a.rc, inv_rc,
                                          a.eta, nb, a.cap, a.pw != nullptr ? STAGE_SPECIES : -1);
        const int s_a = int(sa);
        AngAcc<T, FPL> acc;
        acc.clear();

        int head = 0, qn = 0;  // ring buffer of closed (tj,tk) pairs, warp-uniform
        // process `cnt` queued triangles: phase 1 (records) then phase 2 (functions)
        auto flush = [&](int cnt) {
            __syncwarp();
            if (lane < cnt) {
                const int pk = queue[(head + lane) & 63];
                make_record<T, MODE>(nb, pk & 0xffff, pk >> 16, a.kind, a.pm1, a.ladder, a.step, a.rc, inv_rc, a.eta,
                                     recs + lane * REC_STRIDE, a.pw, a.ns, s_a);
            }
            __syncwarp();
            if constexpr (sizeof(T) == 4) {
                if constexpr (FPL == 2) {
                    if (a.ladder) { phase2_ladder<MODE>(recs, cnt, grp, ngrp, cls_off, zeta, zm1, acc); __syncwarp(); return; }
                }
                if (a.pm1) phase2_packed<FPL, MODE>(recs, cnt, grp, ngrp, cls_off, zeta, zm1, acc);
                else phase2_generic<T, FPL, MODE>(recs, cnt, grp, ngrp, cls_off, a.pm1, zeta, zm1, lam, acc);
            } else {
                phase2_generic<T, FPL, MODE>(recs, cnt, grp, ngrp, cls_off, a.pm1, zeta, zm1, lam, acc);
            }
            __syncwarp();
        };

        for (int tj = 0; tj + 1 < nn; ++tj) {
            const Q4<T> uj = ld4<T>(nb.u4 + 4 * tj);
            for (int kb = tj + 1; kb < nn; kb += 32) {
                const int tk = kb + lane;
                bool ok = false;
                if (tk < nn) {
                    const Q4<T> uk = ld4<T>(nb.u4 + 4 * tk);
                    const T wx = uj.a - uk.a, wy = uj.b - uk.b, wz = uj.c - uk.c;
                    const T z2 = wx * wx + wy * wy + wz * wz;
                    ok = (z2 < rc2) && (z2 > T(0));  // the third side of the triangle (symm_funcs.py:34-50)
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) queue[(head + qn + __popc(m & ((1u << lane) - 1u))) & 63] = tj | (tk << 16);
                qn += __popc(m);
                if (qn >= ANG_BATCH) {
                    flush(ANG_BATCH);
                    head = (head + ANG_BATCH) & 63;
                    qn -= ANG_BATCH;
                }
            }
        }
        if (qn > 0) flush(qn);

        // combine the triangle groups, then lanes < G own their functions' totals
        T val[FPL], gx[FPL], gy[FPL], gz[FPL];
#pragma unroll
        for (int i = 0; i < FPL; ++i) {
            val[i] = acc.val[i]; gx[i] = acc.gxa[i] + acc.gxb[i]; gy[i] = acc.gya[i] + acc.gyb[i]; gz[i] = acc.gza[i] + acc.gzb[i];
            for (int o = G; o < 32; o <<= 1) {
                val[i] += __shfl_xor_sync(0xffffffffu, val[i], o);
                if (MODE != MODE_G) {
                    gx[i] += __shfl_xor_sync(0xffffffffu, gx[i], o);
                    gy[i] += __shfl_xor_sync(0xffffffffu, gy[i], o);
                    gz[i] += __shfl_xor_sync(0xffffffffu, gz[i], o);
                }
            }
        }
        T fx = 0, fy = 0, fz = 0;
        if (grp == 0) {
#pragma unroll
            for (int i = 0; i < FPL; ++i) {
                if (col[i] < 0) continue;
                const int64_t o = row * a.nf + col[i];
                const T cf = coef[i];
                if (MODE == MODE_G || MODE == MODE_DG) a.G[o] = cf * val[i];
                if (MODE == MODE_DG) {
                    const T dx = -cf * gx[i], dy = -cf * gy[i], dz = -cf * gz[i];
                    a.dG[3 * o] = dx; a.dG[3 * o + 1] = dy; a.dG[3 * o + 2] = dz;
                    bad |= !(isfinite(dx) && isfinite(dy) && isfinite(dz));
                }
                if (MODE == MODE_FORCE) {
                    const T wv = a.w[o] * cf;  // F = -sum_f w dG = +sum_f w coef g
                    fx = M<T>::fma_(wv, gx[i], fx); fy = M<T>::fma_(wv, gy[i], fy); fz = M<T>::fma_(wv, gz[i], fz);
                }
            }
        }
        if (MODE == MODE_FORCE) {
            fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
            if (lane == 0) { a.force[3 * row] += fx; a.force[3 * row + 1] += fy; a.force[3 * row + 2] += fz; }
        }
        __syncwarp();
    }
    if (MODE == MODE_DG && bad) atomicOr(a.flag, NNMD_FLAG_NAN_DG);
}

#include "sf_edge.cuh"

// =============================================================================================
// K5 normalise, K4c contraction
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) normalize_kernel(const T *G, int64_t n_rows, int nf, T *ghat,
                                                        uint32_t *__restrict__ flag) {
    const int lane = lane_id(), warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    bool bad = false;
    for (int64_t row = int64_t(blockIdx.x) * wpb + warp; row < n_rows; row += int64_t(gridDim.x) * wpb) {
        const T *g = G + row * nf;
        T ss = 0;
        for (int f = lane; f < nf; f += 32) { const T v = g[f]; ss = M<T>::fma_(v, v, ss); }
        ss = warp_sum(ss);
        const T nrm = M<T>::sqrt_(ss);
        for (int f = lane; f < nf; f += 32) {
            const T v = g[f] / nrm;  // 0/0 -> NaN for an atom without neighbours, like calculate_sf.py:101
            ghat[row * nf + f] = v;
            bad |= !(v == v);
        }
    }
    if (bad) atomicOr(flag, NNMD_FLAG_NAN_G);
}

template <typename T>
__global__ void __launch_bounds__(256) contract_kernel(const T *__restrict__ dE, const T *__restrict__ dG, int64_t n_rows,
                                                       int nf, T *__restrict__ force) {
    const int lane = lane_id(), warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int64_t row = int64_t(blockIdx.x) * wpb + warp; row < n_rows; row += int64_t(gridDim.x) * wpb) {
        const T *w = dE + row * nf;
        const T *g = dG + row * nf * 3;
        T acc[3] = {0, 0, 0};
        // 96 contiguous elements per step: lane l covers component (l % 3) in every step
        for (int e0 = 0; e0 < 3 * nf; e0 += 96) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int e = e0 + 32 * k + lane;
                if (e < 3 * nf) {
                    const T p = w[e / 3] * g[e];
                    const int c = e % 3;
                    acc[0] += c == 0 ? p : T(0); acc[1] += c == 1 ? p : T(0); acc[2] += c == 2 ? p : T(0);
                }
            }
        }
        const T fx = warp_sum(acc[0]), fy = warp_sum(acc[1]), fz = warp_sum(acc[2]);
        if (lane == 0) { force[3 * row] = -fx; force[3 * row + 1] = -fy; force[3 * row + 2] = -fz; }
    }
}

// =============================================================================================
// virial of a frame from per-atom forces:  W = - sum_a (p_a - o) (x) F_a
// The reference never adds periodic images, so every frame is a finite cluster and this IS the
// pair-decomposed virial -1/2 sum_{a != b} r_ab (x) f_ab whenever the forces are those of a translation-invariant
// energy (sum_a F_a = 0).  o = centroid of the frame's centre atoms (or a caller-supplied origin), which fixes the
// otherwise origin-dependent value for the reference-semantic forces, whose sum need not vanish.
// =============================================================================================
template <typename T>
__global__ void centroid_kernel(const T *__restrict__ posw, int64_t n_rows, int64_t n_atoms, int64_t n_centres, double *__restrict__ csum) {
    for (int64_t row = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; row < n_rows; row += int64_t(gridDim.x) * blockDim.x) {
        const int64_t frame = row / n_centres;
        T x, y, z;
        load_pos<T>(posw, frame * n_atoms + (row - frame * n_centres), x, y, z);
        atomicAdd(csum + 3 * frame, double(x)); atomicAdd(csum + 3 * frame + 1, double(y)); atomicAdd(csum + 3 * frame + 2, double(z));
    }
}

template <typename T>
__global__ void virial_kernel(const T *__restrict__ posw, const T *__restrict__ force, int64_t n_rows, int64_t n_atoms,
                              int64_t n_centres, const double *__restrict__ csum, double ox, double oy, double oz, int use_origin,
                              double *__restrict__ wsum) {
    for (int64_t row = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; row < n_rows; row += int64_t(gridDim.x) * blockDim.x) {
        const int64_t frame = row / n_centres;
        T x, y, z;
        load_pos<T>(posw, frame * n_atoms + (row - frame * n_centres), x, y, z);
        double r[3] = {double(x), double(y), double(z)};
        if (use_origin) { r[0] -= ox; r[1] -= oy; r[2] -= oz; }
        else { r[0] -= csum[3 * frame] / double(n_centres); r[1] -= csum[3 * frame + 1] / double(n_centres); r[2] -= csum[3 * frame + 2] / double(n_centres); }
        const double f[3] = {double(force[3 * row]), double(force[3 * row + 1]), double(force[3 * row + 2])};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) atomicAdd(wsum + 9 * frame + 3 * i + j, -r[i] * f[j]);
    }
}

template <typename T>
__global__ void virial_store_kernel(const double *__restrict__ wsum, int64_t n, T *__restrict__ out) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = T(wsum[i]);
}

// =============================================================================================
// launch helpers
// =============================================================================================
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

template <typename T, int MODE, bool UNI, bool MS = false>
static int launch_radial_t(RadArgs<T> a, cudaStream_t st) {
    int wpb = 8;
    while (wpb > 1 && size_t(wpb) * NB_FIELDS * a.cap * sizeof(T) > 160 * 1024) wpb >>= 1;
    const size_t smem = size_t(wpb) * NB_FIELDS * a.cap * sizeof(T);
    if (smem > 220 * 1024) return set_err(NNMD_ERR_CAPACITY, "radial: neighbour row of %d entries exceeds shared memory", a.cap);
    auto kern = radial_kernel<T, MODE, UNI, MS>;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int64_t blocks = (a.n_rows + wpb - 1) / wpb;
#ifndef NNMD_RADIAL_BPS
#define NNMD_RADIAL_BPS 64  // blocks per SM of the radial grid; 8: 2.05 ms, 16: 1.96, 64 and uncapped: 1.90 per 1M atoms (row costs vary: finer blocks balance better)
#endif
    blocks = std::min<int64_t>(blocks, int64_t(sm_count()) * NNMD_RADIAL_BPS);
    ProfScope prof(PROF_RADIAL, st);
    kern<<<unsigned(blocks), wpb * 32, smem, st>>>(a);
    NNMD_LAUNCH_CHECK("radial_kernel");
    return NNMD_OK;
}

template <typename T, int MODE>
static int launch_radial(const RadialTab &rad, RadArgs<T> a, cudaStream_t st) {
    if (a.wr != nullptr) return rad.uniform_rc ? launch_radial_t<T, MODE, true, true>(a, st) : launch_radial_t<T, MODE, false, true>(a, st);
    return rad.uniform_rc ? launch_radial_t<T, MODE, true>(a, st) : launch_radial_t<T, MODE, false>(a, st);
}

template <typename T, int FPL, int MODE>
static int launch_angular_t(AngArgs<T> a, cudaStream_t st) {
    const size_t per_warp = size_t(ANG_BATCH) * REC_STRIDE * sizeof(T) + size_t(NB_FIELDS) * a.cap * sizeof(T) + 64 * sizeof(int);
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > 220 * 1024) return set_err(NNMD_ERR_CAPACITY, "angular: neighbour row of %d entries exceeds shared memory", a.cap);
    auto kern = angular_kernel<T, FPL, MODE>;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 1;
    NNMD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
    if (occ < 1) occ = 1;
    int64_t blocks = (a.n_rows + wpb - 1) / wpb;
    blocks = std::min<int64_t>(blocks, int64_t(sm_count()) * occ);  // persistent: one resident wave, rows strided
    ProfScope prof(PROF_ANGULAR, st);
    kern<<<unsigned(blocks), wpb * 32, smem, st>>>(a);
    NNMD_LAUNCH_CHECK("angular_kernel");
    return NNMD_OK;
}

template <typename T, int MODE>
static int launch_angular(const AngGroup &g, AngArgs<T> a, cudaStream_t st) {
    a.kind = g.kind; a.pm1 = g.pm1; a.lanes = g.lanes; a.log_lanes = ilog2(g.lanes);
    a.ladder = (g.ladder && sizeof(T) == 4) ? 1 : 0;
    a.rc = T(g.rc); a.eta = T(g.eta); a.step = T(g.ladder_step);
    a.zeta = (const T *)g.d_zeta; a.lam = (const T *)g.d_lam; a.coef = (const T *)g.d_coef; a.col = g.d_col;
    switch (g.fpl) {
        case 1: return launch_angular_t<T, 1, MODE>(a, st);
        case 2: return launch_angular_t<T, 2, MODE>(a, st);
        case 4: return launch_angular_t<T, 4, MODE>(a, st);
    }
    return set_err(NNMD_ERR_INVALID, "internal: bad fpl %d", g.fpl);
}

template <typename T, int MODE>
static int sf_run(const nnmd_sf_plan *p, const T *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres,
                  const int64_t *offsets, const int32_t *idx, int64_t max_row, const int64_t *edge_offsets,
                  int64_t n_edges, void *edge_ws, int64_t edge_ws_bytes, T *G, T *dG, const T *w, T *force,
                  uint32_t *flag, cudaStream_t st) {
    const int64_t n_rows = n_frames * n_centres;
    if (n_rows == 0) return NNMD_OK;
    int cap = int((std::max<int64_t>(max_row, 1) + 3) & ~int64_t(3));
    if (cap >= 32768) return set_err(NNMD_ERR_CAPACITY, "neighbour row of %lld entries is not supported", (long long)max_row);
    auto run_radial = [&](const RadialTab &rad) -> int {
        RadArgs<T> a;
        a.posw = posw; a.n_rows = n_rows; a.n_atoms = n_atoms; a.n_centres = n_centres; a.offsets = offsets; a.idx = idx;
        a.cap = cap; a.n = rad.n; a.rc_max = T(rad.rc_max);
        a.rc = (const T *)rad.d_rc; a.eta = (const T *)rad.d_eta; a.rs = (const T *)rad.d_rs; a.col = rad.d_col;
        a.nf = p->nf; a.G = G; a.dG = dG; a.w = w; a.force = force; a.flag = flag; a.wr = (const T *)rad.d_wr;
        return launch_radial<T, MODE>(rad, a, st);
    };
    if (p->rad.n > 0) {
        int r = run_radial(p->rad);
        if (r) return r;
    }
    for (const RadialTab &rad : p->ms_rads) {  // multi-species plans: one launch per neighbour-weight vector
        int r = run_radial(rad);
        if (r) return r;
    }
    // the first EDGE_WS_HEADER bytes of the scratch hold the row counter of the dynamically scheduled kernel: it belongs to
    // the CALL (two calls that share a plan on different streams bring their own scratch, hence their own counter)
    const bool edge_ok = edge_offsets && edge_ws && p->edge_slots > 0 &&
                         edge_ws_bytes >= EDGE_WS_HEADER + n_edges * 2 * int64_t(p->edge_slots) * int64_t(sizeof(float));
    for (const AngGroup &g : p->groups) {
        if constexpr (sizeof(T) == 4) {
            if (g.edge_path && edge_ok) {
                EdgeArgs e;
                e.posw = posw; e.n_rows = n_rows; e.n_atoms = n_atoms; e.n_centres = n_centres; e.offsets = offsets; e.idx = idx;
                e.edge_offsets = edge_offsets; e.cap = cap; e.kind = g.kind; e.ladder = g.ladder;
                e.rc = float(g.rc); e.eta = float(g.eta); e.step = float(g.ladder_step);
                e.zeta = (const float *)g.d_zeta; e.lam = (const float *)g.d_lam; e.coef = (const float *)g.d_coef; e.col = g.d_col;
                e.nf = p->nf; e.nfa = p->edge_slots; e.ebase = g.edge_base; e.E = (float *)((char *)edge_ws + EDGE_WS_HEADER);
                e.G = G; e.dG = dG; e.w = w; e.force = force; e.flag = flag; e.row_counter = (unsigned long long *)edge_ws;
#ifdef NNMD_TUNING
                e.dbg = 0;
#endif
                int r = launch_edge<MODE>(g, e, st);
                if (r) return r;
                continue;
            }
        }
        AngArgs<T> a;
        a.posw = posw; a.n_rows = n_rows; a.n_atoms = n_atoms; a.n_centres = n_centres; a.offsets = offsets; a.idx = idx;
        a.cap = cap; a.nf = p->nf; a.G = G; a.dG = dG; a.w = w; a.force = force; a.flag = flag;
        a.pw = (const T *)g.d_pw; a.ns = p->n_species;
        int r = launch_angular<T, MODE>(g, a, st);
        if (r) return r;
    }
    return NNMD_OK;
}

}  // namespace nnmd

using namespace nnmd;

// =============================================================================================
// C ABI
// =============================================================================================
// weights == NULL: single-species plan (the reference's semantics).  Otherwise n_species > 0 and weights holds, per
// function, an n_species x n_species block: radial functions read its first row as the neighbour weights wr[s_j];
// angular functions read all of it as the symmetric pair weights pw[s_j][s_k].
static int plan_create(int dtype, int nf, const int32_t *kinds, const double *params, int n_species, const double *weights,
                       nnmd_sf_plan **out) {
    if (!out) return set_err(NNMD_ERR_INVALID, "plan: null output");
    *out = nullptr;
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "plan: bad dtype %d", dtype);
    if (nf <= 0 || !kinds || !params) return set_err(NNMD_ERR_INVALID, "plan: empty symmetry-function table");
    if (weights && (n_species < 1 || n_species > 64)) return set_err(NNMD_ERR_INVALID, "plan: 1..64 species supported");
    for (int f = 0; f < nf; ++f) {
        const int k = kinds[f];
        if (k != NNMD_G1 && k != NNMD_G2 && k != NNMD_G4 && k != NNMD_G5)
            return set_err(NNMD_ERR_INVALID, "Unknown symmetry function number: %d", k);
        if (!(params[4 * f] > 0.0)) return set_err(NNMD_ERR_INVALID, "plan: function %d has a non-positive cutoff", f);
    }
    nnmd_sf_plan *p = new (std::nothrow) nnmd_sf_plan();
    if (!p) return set_err(NNMD_ERR_INVALID, "plan: out of host memory");
    p->dtype = dtype; p->nf = nf; p->rc_max = 0;
    p->n_species = weights ? n_species : 0;
    const int S2 = weights ? n_species * n_species : 0;
    using WKey = std::vector<double>;
    auto wkey = [&](int f, int len) { return weights ? WKey(weights + size_t(f) * S2, weights + size_t(f) * S2 + len) : WKey(); };
    struct RadBuild { std::vector<double> rc, eta, rs; std::vector<int> col; double rc_max = 0; };
    std::map<WKey, RadBuild> rads;
    std::vector<WKey> rad_order;
    std::map<std::tuple<int, double, double, WKey>, std::vector<int>> groups;
    std::vector<std::tuple<int, double, double, WKey>> order;  // first-appearance order keeps launches deterministic
    for (int f = 0; f < nf; ++f) {
        p->rc_max = std::max(p->rc_max, params[4 * f]);
        if (kinds[f] == NNMD_G1 || kinds[f] == NNMD_G2) {
            const WKey k = wkey(f, weights ? n_species : 0);
            if (!rads.count(k)) rad_order.push_back(k);
            RadBuild &b = rads[k];
            b.rc.push_back(params[4 * f]);
            b.eta.push_back(kinds[f] == NNMD_G2 ? params[4 * f + 1] : 0.0);
            b.rs.push_back(kinds[f] == NNMD_G2 ? params[4 * f + 2] : 0.0);
            b.col.push_back(f);
            b.rc_max = std::max(b.rc_max, params[4 * f]);
        } else {
            auto key = std::make_tuple(int(kinds[f]), params[4 * f], params[4 * f + 1], wkey(f, S2));
            if (!groups.count(key)) order.push_back(key);
            groups[key].push_back(f);
        }
    }
    int r = NNMD_OK;
    for (const WKey &k : rad_order) {
        if (r) break;
        const RadBuild &b = rads[k];
        RadialTab t;
        t.n = int(b.col.size());
        t.rc_max = b.rc_max;
        t.uniform_rc = 1;
        for (double v : b.rc) if (v != b.rc_max) t.uniform_rc = 0;
        if (!r) r = upload_real(dtype, b.rc, &t.d_rc);
        if (!r) r = upload_real(dtype, b.eta, &t.d_eta);
        if (!r) r = upload_real(dtype, b.rs, &t.d_rs);
        if (!r) r = upload_int(b.col, &t.d_col);
        if (!r && weights) r = upload_real(dtype, k, &t.d_wr);
        if (weights) p->ms_rads.push_back(t); else p->rad = t;
    }
    for (auto &key : order) {
        if (r) break;
        const size_t first = p->groups.size();
        r = build_group(dtype, std::get<0>(key), std::get<1>(key), std::get<2>(key), groups[key], params, p->groups);
        if (!r && weights)
            for (size_t gi = first; gi < p->groups.size() && !r; ++gi) {
                p->groups[gi].edge_path = 0;  // the edge-centric kernel relies on every atom weighting a triangle alike
                r = upload_real(dtype, std::get<3>(key), &p->groups[gi].d_pw);
            }
    }
    if (r) { nnmd_sf_plan_destroy(p); return r; }
    for (AngGroup &g : p->groups)
        if (g.edge_path) { g.edge_base = p->edge_slots; p->edge_slots += g.n_slots; }
    *out = p;
    return NNMD_OK;
}

extern "C" int nnmd_sf_plan_create(int dtype, int nf, const int32_t *kinds, const double *params, nnmd_sf_plan **out) {
    return plan_create(dtype, nf, kinds, params, 0, nullptr, out);
}

extern "C" int nnmd_sf_plan_create_species(int dtype, int nf, const int32_t *kinds, const double *params, int n_species,
                                           const double *weights, nnmd_sf_plan **out) {
    if (!weights) return set_err(NNMD_ERR_INVALID, "plan: null species weights");
    for (int f = 0; f < nf && kinds; ++f)
        if (kinds[f] == NNMD_G4 || kinds[f] == NNMD_G5)
            for (int a = 0; a < n_species; ++a)
                for (int b = 0; b < a; ++b)
                    if (weights[(size_t(f) * n_species + a) * n_species + b] != weights[(size_t(f) * n_species + b) * n_species + a])
                        return set_err(NNMD_ERR_INVALID, "plan: the pair weights of angular function %d are not symmetric", f);
    return plan_create(dtype, nf, kinds, params, n_species, weights, out);
}

template <typename T>
__global__ void set_species_kernel(T *__restrict__ posw, const int32_t *__restrict__ species, int64_t n_total, int64_t n_atoms) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n_total) posw[4 * i + 3] = T(species[i % n_atoms]);
}

extern "C" int nnmd_set_species(int dtype, void *posw, const int32_t *species, int64_t n_frames, int64_t n_atoms, void *stream) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "set_species: bad dtype %d", dtype);
    if (n_frames < 0 || n_atoms < 0) return set_err(NNMD_ERR_INVALID, "set_species: bad sizes");
    const int64_t n_total = n_frames * n_atoms;
    if (n_total == 0) return NNMD_OK;
    if (!posw || !species) return set_err(NNMD_ERR_INVALID, "set_species: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = unsigned((n_total + 255) / 256);
    if (dtype == NNMD_F32) set_species_kernel<float><<<blocks, 256, 0, st>>>((float *)posw, species, n_total, n_atoms);
    else set_species_kernel<double><<<blocks, 256, 0, st>>>((double *)posw, species, n_total, n_atoms);
    NNMD_LAUNCH_CHECK("set_species_kernel");
    return NNMD_OK;
}

extern "C" void nnmd_sf_plan_destroy(nnmd_sf_plan *p) {
    if (!p) return;
    cudaFree(p->rad.d_rc); cudaFree(p->rad.d_eta); cudaFree(p->rad.d_rs); cudaFree(p->rad.d_col);
    for (nnmd::RadialTab &t : p->ms_rads) { cudaFree(t.d_rc); cudaFree(t.d_eta); cudaFree(t.d_rs); cudaFree(t.d_col); cudaFree(t.d_wr); }
    for (AngGroup &g : p->groups) { cudaFree(g.d_zeta); cudaFree(g.d_lam); cudaFree(g.d_coef); cudaFree(g.d_col); cudaFree(g.d_pw); }
    delete p;
}

extern "C" int nnmd_set_option(const char *name, int value) {
    if (name && strcmp(name, "edge_form") == 0) { nnmd::g_edge_form = value ? 1 : 0; return NNMD_OK; }
    return set_err(NNMD_ERR_INVALID, "set_option: unknown option %s", name ? name : "(null)");
}

extern "C" double nnmd_sf_plan_rc_max(const nnmd_sf_plan *p) { return p ? p->rc_max : 0.0; }

static int sf_args_ok(const nnmd_sf_plan *plan, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres,
                      const int64_t *offsets, const int32_t *idx, int64_t max_row) {
    if (!plan) return set_err(NNMD_ERR_INVALID, "sf: null plan");
    if (n_frames < 0 || n_atoms < 0 || n_centres < 0 || n_centres > n_atoms || max_row < 0)
        return set_err(NNMD_ERR_INVALID, "sf: bad sizes");
    if (n_frames * n_centres > 0 && (!posw || !offsets)) return set_err(NNMD_ERR_INVALID, "sf: null pointer");
    if (max_row > 0 && !idx) return set_err(NNMD_ERR_INVALID, "sf: null neighbour indices");
    return NNMD_OK;
}

extern "C" int64_t nnmd_sf_edge_scratch_bytes(const nnmd_sf_plan *plan, int64_t n_edges) {
    if (!plan || n_edges < 0 || plan->dtype != NNMD_F32) return 0;
    if (n_edges >= (int64_t(1) << 31)) return 0;  // slots are int32: beyond that the vertex-centric kernel is used
    if (plan->edge_slots == 0) return 0;
    return EDGE_WS_HEADER + n_edges * 2 * int64_t(plan->edge_slots) * int64_t(sizeof(float));
}

extern "C" int nnmd_sf_eval(const nnmd_sf_plan *plan, const void *posw, int64_t n_frames, int64_t n_atoms,
                            int64_t n_centres, const int64_t *offsets, const int32_t *idx, int64_t max_row,
                            const int64_t *edge_offsets, int64_t n_edges, void *edge_ws, int64_t edge_ws_bytes, void *G,
                            void *dG, uint32_t *flag, void *stream) {
    int r = sf_args_ok(plan, posw, n_frames, n_atoms, n_centres, offsets, idx, max_row);
    if (r) return r;
    if (n_frames * n_centres > 0 && (!G || !flag)) return set_err(NNMD_ERR_INVALID, "sf_eval: null output");
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->dtype == NNMD_F32) {
        if (dG) return sf_run<float, MODE_DG>(plan, (const float *)posw, n_frames, n_atoms, n_centres, offsets, idx, max_row, edge_offsets, n_edges, edge_ws, edge_ws_bytes, (float *)G, (float *)dG, nullptr, nullptr, flag, st);
        return sf_run<float, MODE_G>(plan, (const float *)posw, n_frames, n_atoms, n_centres, offsets, idx, max_row, edge_offsets, n_edges, edge_ws, edge_ws_bytes, (float *)G, nullptr, nullptr, nullptr, flag, st);
    }
    if (dG) return sf_run<double, MODE_DG>(plan, (const double *)posw, n_frames, n_atoms, n_centres, offsets, idx, max_row, nullptr, 0, nullptr, 0, (double *)G, (double *)dG, nullptr, nullptr, flag, st);
    return sf_run<double, MODE_G>(plan, (const double *)posw, n_frames, n_atoms, n_centres, offsets, idx, max_row, nullptr, 0, nullptr, 0, (double *)G, nullptr, nullptr, nullptr, flag, st);
}

extern "C" int nnmd_sf_force(const nnmd_sf_plan *plan, const void *posw, int64_t n_frames, int64_t n_atoms,
                             int64_t n_centres, const int64_t *offsets, const int32_t *idx, int64_t max_row,
                             const int64_t *edge_offsets, int64_t n_edges, void *edge_ws, int64_t edge_ws_bytes,
                             const void *w, void *force, void *virial, void *stream) {
    return nnmd_sf_force_flag(plan, posw, n_frames, n_atoms, n_centres, offsets, idx, max_row, edge_offsets, n_edges, edge_ws,
                              edge_ws_bytes, w, force, virial, nullptr, stream);
}

extern "C" int nnmd_sf_force_flag(const nnmd_sf_plan *plan, const void *posw, int64_t n_frames, int64_t n_atoms,
                                  int64_t n_centres, const int64_t *offsets, const int32_t *idx, int64_t max_row,
                                  const int64_t *edge_offsets, int64_t n_edges, void *edge_ws, int64_t edge_ws_bytes,
                                  const void *w, void *force, void *virial, uint32_t *flag, void *stream) {
    int r = sf_args_ok(plan, posw, n_frames, n_atoms, n_centres, offsets, idx, max_row);
    if (r) return r;
    const int64_t n_rows = n_frames * n_centres;
    if (n_rows > 0 && (!w || !force)) return set_err(NNMD_ERR_INVALID, "sf_force: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = plan->dtype == NNMD_F32 ? 4 : 8;
    if (n_rows > 0) NNMD_CUDA_CHECK(cudaMemsetAsync(force, 0, size_t(n_rows) * 3 * es, st));
    if (plan->dtype == NNMD_F32)
        r = sf_run<float, MODE_FORCE>(plan, (const float *)posw, n_frames, n_atoms, n_centres, offsets, idx, max_row, edge_offsets, n_edges, edge_ws, edge_ws_bytes, nullptr, nullptr, (const float *)w, (float *)force, flag, st);
    else
        r = sf_run<double, MODE_FORCE>(plan, (const double *)posw, n_frames, n_atoms, n_centres, offsets, idx, max_row, nullptr, 0, nullptr, 0, nullptr, nullptr, (const double *)w, (double *)force, flag, st);
    if (r || !virial || n_frames == 0) return r;
    // virial of the forces just computed (frame centroid as origin), scratch from the stream-ordered allocator
    void *ws = nullptr;
    NNMD_CUDA_CHECK(cudaMallocAsync(&ws, size_t(nnmd_virial_workspace_bytes(n_frames)), st));
    r = nnmd_virial(plan->dtype, posw, force, n_frames, n_atoms, n_centres, nullptr, ws, virial, stream);
    cudaFreeAsync(ws, st);
    return r;
}

extern "C" int64_t nnmd_virial_workspace_bytes(int64_t n_frames) { return n_frames < 0 ? -1 : n_frames * 12 * int64_t(sizeof(double)); }

extern "C" int nnmd_virial(int dtype, const void *posw, const void *force, int64_t n_frames, int64_t n_atoms, int64_t n_centres,
                           const double *origin, void *ws, void *virial, void *stream) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "virial: bad dtype %d", dtype);
    if (n_frames < 0 || n_atoms < 0 || n_centres < 0 || n_centres > n_atoms) return set_err(NNMD_ERR_INVALID, "virial: bad sizes");
    if (n_frames == 0) return NNMD_OK;
    if (!posw || !force || !ws || !virial) return set_err(NNMD_ERR_INVALID, "virial: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    double *csum = (double *)ws, *wsum = csum + 3 * n_frames;
    NNMD_CUDA_CHECK(cudaMemsetAsync(ws, 0, size_t(n_frames) * 12 * sizeof(double), st));
    const int64_t n_rows = n_frames * n_centres;
    if (n_rows > 0) {
        const unsigned blocks = unsigned(std::min<int64_t>((n_rows + 255) / 256, int64_t(sm_count()) * 8));
        const double o[3] = {origin ? origin[0] : 0.0, origin ? origin[1] : 0.0, origin ? origin[2] : 0.0};
        if (dtype == NNMD_F32) {
            if (!origin) centroid_kernel<float><<<blocks, 256, 0, st>>>((const float *)posw, n_rows, n_atoms, n_centres, csum);
            virial_kernel<float><<<blocks, 256, 0, st>>>((const float *)posw, (const float *)force, n_rows, n_atoms, n_centres, csum, o[0], o[1], o[2], origin ? 1 : 0, wsum);
        } else {
            if (!origin) centroid_kernel<double><<<blocks, 256, 0, st>>>((const double *)posw, n_rows, n_atoms, n_centres, csum);
            virial_kernel<double><<<blocks, 256, 0, st>>>((const double *)posw, (const double *)force, n_rows, n_atoms, n_centres, csum, o[0], o[1], o[2], origin ? 1 : 0, wsum);
        }
        NNMD_LAUNCH_CHECK("virial_kernel");
    }
    const int64_t n = n_frames * 9;
    if (dtype == NNMD_F32) virial_store_kernel<float><<<unsigned((n + 255) / 256), 256, 0, st>>>(wsum, n, (float *)virial);
    else virial_store_kernel<double><<<unsigned((n + 255) / 256), 256, 0, st>>>(wsum, n, (double *)virial);
    NNMD_LAUNCH_CHECK("virial_store_kernel");
    return NNMD_OK;
}

extern "C" int nnmd_sf_normalize(int dtype, const void *G, int64_t n_rows, int nf, void *ghat, uint32_t *flag, void *stream) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "normalize: bad dtype %d", dtype);
    if (n_rows < 0 || nf <= 0) return set_err(NNMD_ERR_INVALID, "normalize: bad sizes");
    if (n_rows == 0) return NNMD_OK;
    if (!G || !ghat || !flag) return set_err(NNMD_ERR_INVALID, "normalize: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int wpb = 8;
    int64_t blocks = std::min<int64_t>((n_rows + wpb - 1) / wpb, int64_t(sm_count()) * 16);
    ProfScope prof(PROF_NORMALIZE, st);
    if (dtype == NNMD_F32) normalize_kernel<float><<<unsigned(blocks), wpb * 32, 0, st>>>((const float *)G, n_rows, nf, (float *)ghat, flag);
    else normalize_kernel<double><<<unsigned(blocks), wpb * 32, 0, st>>>((const double *)G, n_rows, nf, (double *)ghat, flag);
    NNMD_LAUNCH_CHECK("normalize_kernel");
    return NNMD_OK;
}

extern "C" int nnmd_force_contract(int dtype, const void *dE, const void *dG, int64_t n_rows, int nf, void *force, void *stream) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "contract: bad dtype %d", dtype);
    if (n_rows < 0 || nf <= 0) return set_err(NNMD_ERR_INVALID, "contract: bad sizes");
    if (n_rows == 0) return NNMD_OK;
    if (!dE || !dG || !force) return set_err(NNMD_ERR_INVALID, "contract: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int wpb = 8;
    int64_t blocks = std::min<int64_t>((n_rows + wpb - 1) / wpb, int64_t(sm_count()) * 16);
    ProfScope prof(PROF_CONTRACT, st);
    if (dtype == NNMD_F32) contract_kernel<float><<<unsigned(blocks), wpb * 32, 0, st>>>((const float *)dE, (const float *)dG, n_rows, nf, (float *)force);
    else contract_kernel<double><<<unsigned(blocks), wpb * 32, 0, st>>>((const double *)dE, (const double *)dG, n_rows, nf, (double *)force);
    NNMD_LAUNCH_CHECK("contract_kernel");
    return NNMD_OK;
}
