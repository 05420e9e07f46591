// sf_edge_m.cuh -- moment form of the edge-centric angular kernel K3e (included by sf_edge.cuh, inside namespace nnmd).
//
// Same decomposition as angular_edge4_kernel (every unordered edge evaluated once by its owner, 4 functions on each of 16
// lanes, the two half-warps on alternate records), with the per-record work cut from 11 to 9 FMA-pipe lane-ops per function.
//
// For the edge (a, j) with length x, every centre term of a closing triangle factorises as
//     E_X = e_X(x) * W_X ,      dE_X/dx = e_X'(x) * W_X ,
// where e_X depends on the EDGE only (G4: psi(x) for all three centres; G5: psi(x) for the centres a and j, fc(x) for the
// third atom) and W_X collects everything else.  With b = 1 +- |c_X| and q = b^(zeta - 1):
//     d/dx [ b^zeta E_X ] = q * ( zeta * (+-sgn c_X) E_X dc_X/dx ) + (q b W_X) * e_X'(x)
// so that, per function, the sum over the third atoms needs only two kinds of power sums ("moments"),
//     SA  = sum_r sum_X q_X A_X        A_X = +-sgn(c_X) E_X dc_X/dx
//     MW_X = sum_r q_X (b_X W_X)       X = a, j, k
// and, when the edge retires,
//     Phi   = zeta * SA + e_a'(x) (MW_a + MW_j) + e_k'(x) MW_k
//     value of the centre a  += e_a(x) MW_a ,      Gamma (the value share of j) = e_a(x) MW_j .
// The old record carried B_X = b_X dE_X/dx and V_X = b_X E_X separately (5 + 3 + 3 FMA-pipe ops per function and record);
// here B and V come out of the same moment (3 ladder multiplications + 3 + 3).
//
// record, per lambda class (EM_CLASS floats):  [ L0 L1 R0 R1 | A0 A1 W0 W1 | L2 R2 A2 W2 ]
//     L_X = log2 b_X,  R_X = 2^(step L_X) (zeta ladder),  W_X here = b_X W_X (0 where |c_X| <= 1e-3)
#pragma once

constexpr int EM_CLASS = 12;
#ifndef NNMD_EDGEM_STRIDE
#define NNMD_EDGEM_STRIDE 24
#endif
constexpr int EM_STRIDE = NNMD_EDGEM_STRIDE;  // 24 used.  28 would make the 16 B record stores conflict-free (24: two-way), but the
                                              // 1 kB per warp it costs is what lets a fifth block fit on the SM (measured: 47.5 vs 49.4 ms)
#ifndef NNMD_EDGEM_EBATCH
#define NNMD_EDGEM_EBATCH 64
#endif
constexpr int EMBATCH = NNMD_EDGEM_EBATCH;  // (edge, third atom) records built per flush: up to two per lane
// queue entry (tj, tk): 8 + 8 bits while rows have at most 256 entries, else 15 + 15 bits
template <typename QT> struct QPack;
template <> struct QPack<uint16_t> {
    static __device__ __forceinline__ uint16_t pack(int tj, int tk) { return uint16_t(tj | (tk << 8)); }
    static __device__ __forceinline__ int tj(uint16_t v) { return v & 0xff; }
    static __device__ __forceinline__ int tk(uint16_t v) { return v >> 8; }
};
template <> struct QPack<uint32_t> {
    static __device__ __forceinline__ uint32_t pack(int tj, int tk) { return uint32_t(tj | (tk << 15)); }
    static __device__ __forceinline__ int tj(uint32_t v) { return v & 0x7fff; }
    static __device__ __forceinline__ int tk(uint32_t v) { return (v >> 15) & 0x7fff; }
};
static inline size_t edgem_per_warp_bytes(int cap) {
    return size_t(EMBATCH) * EM_STRIDE * sizeof(float) + size_t(NB_FIELDS) * cap * sizeof(float) + 256 * (cap <= 256 ? sizeof(uint16_t) : sizeof(uint32_t));
}

// `store` = false: the lane has no record in this slot; it still runs the arithmetic (on clamped indices) so that the two
// records of a lane are built in ONE instruction stream with twice the instruction-level parallelism.
template <bool WITH_PHI, bool LADDER>
__device__ __forceinline__ void make_record_m(const NbView<float> &nb, int tj, int tk, int kind, float step, float inv_rc,
                                              float eta, float *__restrict__ rec, bool store) {
    using MF = M<float>;
    const Q4<float> uj = ld4<float>(nb.u4 + 4 * tj), uk = ld4<float>(nb.u4 + 4 * tk);
    const float x = uj.d, y = uk.d, ix = nb.invd[tj], iy = nb.invd[tk];
    const float wx = uj.a - uk.a, wy = uj.b - uk.b, wz = uj.c - uk.c;
    const float z2 = wx * wx + wy * wy + wz * wz;
    const float iz = MF::rsqrt_fast(z2), z = z2 * iz;
    const float x2 = x * x, y2 = y * y;
    const float fz = 0.5f * (MF::cospi_fast(z * inv_rc) + 1.f);
    const float psz = MF::exp_fast(-eta * z2) * fz;
    const float psy = nb.psi[tk];
    float c[3], dcx[3], Wb[3], ef[3];
    bool live[3];
    c[0] = cos_law<float>(x2, y2, z2, x, y, ix, iy, live[0]);  // centre a : sides x, y   opposite z
    c[1] = cos_law<float>(x2, z2, y2, x, z, ix, iz, live[1]);  // centre j : sides x, z   opposite y
    c[2] = cos_law<float>(y2, z2, x2, y, z, iy, iz, live[2]);  // centre k : sides y, z   opposite x
    if (WITH_PHI) {
        dcx[0] = live[0] ? iy - c[0] * ix : 0.f;
        dcx[1] = live[1] ? iz - c[1] * ix : 0.f;
        dcx[2] = live[2] ? -x * iy * iz : 0.f;
    }
    if (kind == NNMD_G4) {
        Wb[0] = Wb[1] = Wb[2] = psy * psz;
        if (WITH_PHI) ef[0] = ef[1] = ef[2] = nb.psi[tj];
    } else {  // G5 keeps fc of the opposite side
        Wb[0] = psy * fz;
        Wb[1] = psz * nb.fc[tk];
        Wb[2] = psy * psz;
        if (WITH_PHI) { ef[0] = ef[1] = nb.psi[tj]; ef[2] = nb.fc[tj]; }
    }
    float L[2][3], R[2][3], A[3], W[2][3];
#pragma unroll
    for (int X = 0; X < 3; ++X) {
        const float ac = fabsf(c[X]);
        const float wb = ac > 1e-3f ? Wb[X] : 0.f;  // symm_funcs.py:210-211: the angular factor is 0 for |cos| <= 1e-3
        const float bp = 1.f + ac, bm = fmaxf(1.f - ac, 0.f);
        L[0][X] = MF::log2_fast(bp);
        L[1][X] = fmaxf(MF::log2_fast(bm), -MF::big());
        if (WITH_PHI) A[X] = (c[X] > 0.f ? wb : -wb) * ef[X] * dcx[X];
        W[0][X] = bp * wb;
        W[1][X] = bm * wb;
        if (LADDER) { R[0][X] = MF::exp2_fast(step * L[0][X]); R[1][X] = MF::exp2_fast(step * L[1][X]); }
        else R[0][X] = R[1][X] = 0.f;
    }
    if (!WITH_PHI) A[0] = A[1] = A[2] = 0.f;
    if (!store) return;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        float *r = rec + s * EM_CLASS;
        const float sgn = s ? -1.f : 1.f;
        st4<float>(r, L[s][0], L[s][1], R[s][0], R[s][1]);
        st4<float>(r + 4, sgn * A[0], sgn * A[1], W[s][0], W[s][1]);
        if (WITH_PHI) st4<float>(r + 8, L[s][2], R[s][2], sgn * A[2], W[s][2]);
    }
}

#ifndef NNMD_EDGEM_MINB
#define NNMD_EDGEM_MINB 5
#endif
#ifndef NNMD_EDGEM_UNROLL
#define NNMD_EDGEM_UNROLL 2
#endif
constexpr int NNMD_EDGEM_UNROLL_N = NNMD_EDGEM_UNROLL;

template <bool WITH_PHI, bool LADDER, typename QT>
__global__ void __launch_bounds__(128, NNMD_EDGEM_MINB) angular_edgem_kernel(EdgeArgs a) {
    if (a.flag != nullptr && (*a.flag & NNMD_FLAG_OVERFLOW)) return;  // incomplete neighbour list (capacity-bound build)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using MF = M<float>;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    constexpr int EBATCH = EMBATCH;
    const size_t per_warp = size_t(EBATCH) * EM_STRIDE * sizeof(float) + size_t(NB_FIELDS) * a.cap * sizeof(float) + 256 * sizeof(QT);
    unsigned char *wbase = smem_raw + per_warp * warp;
    float *recs = reinterpret_cast<float *>(wbase);
    NbView<float> nb;
    nb.bind(recs + EBATCH * EM_STRIDE, a.cap);
    QT *queue = reinterpret_cast<QT *>(recs + EBATCH * EM_STRIDE + size_t(NB_FIELDS) * a.cap);  // ring of 256 packed (tj, tk)

    const float inv_rc = MF::rcp(a.rc), rc2 = a.rc * a.rc;
    const int half = lane >> 4, l16 = lane & 15;
    float zeta[4], zm1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { zeta[i] = a.zeta[4 * l16 + i]; zm1[i] = zeta[i] - 1.f; }
    const int class_off = (a.lam[4 * l16] < 0.f) ? EM_CLASS : 0;  // lambda class of this lane's four functions
    const size_t nfa = size_t(a.nfa);
    const int kind = a.kind;
    const float eta = a.eta, step = a.step;
    // the factor of the third atom's moment in Phi: psi'(x) for G4, fc'(x) for G5
    const float *__restrict__ dek = kind == NNMD_G4 ? nb.dpsi : nb.dfc;

    for (;;) {
        unsigned long long next = 0;
        if (lane == 0) next = atomicAdd(a.row_counter, 1ull);
        const int64_t row = int64_t(__shfl_sync(0xffffffffu, next, 0));
        if (row >= a.n_rows) break;
        const int64_t frame = row / a.n_centres;
        const int64_t gi = frame * a.n_atoms + (row - frame * a.n_centres);
        float xa, ya, za;
        load_pos<float>(a.posw, gi, xa, ya, za);
        const int64_t o0 = a.offsets[row], o1 = a.offsets[row + 1];
        const int64_t e0 = a.edge_offsets[row], e1 = a.edge_offsets[row + 1];
        const int first_gt = int((o1 - o0) - (e1 - e0));
        const int nn = stage_row<float, true>(a.posw, a.idx, o0, o1, xa, ya, za, a.rc, inv_rc, eta, nb, a.cap, first_gt);
        // owned edges: replace the own-row slot by the slot in the OTHER endpoint's row (-1 not owned, -2 owned without slot).
        // Rows are ascending, so the owned edges are the last nn - t_first staged entries.
        int t_first = 0;
        for (int t0 = 0; t0 < nn; t0 += 32) {
            const int t = t0 + lane;
            const int ep = t < nn ? nb.epos[t] : 0;
            t_first += __popc(__ballot_sync(0xffffffffu, t < nn && ep < 0));
            if (t < nn && ep >= 0) {
                const int64_t j = __ldg(a.idx + o0 + first_gt + ep);
                const int64_t lj = j - frame * a.n_atoms;
                int code = -2;
                if (lj < a.n_centres) {
                    const int64_t rj = frame * a.n_centres + lj;
                    const int64_t p0 = a.offsets[rj], q0 = a.edge_offsets[rj];
                    int64_t lo = p0, hi = p0 + ((a.offsets[rj + 1] - p0) - (a.edge_offsets[rj + 1] - q0));
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if (int64_t(__ldg(a.idx + mid)) < gi) lo = mid + 1; else hi = mid;
                    }
                    code = int((p0 - q0) + (lo - p0));
                }
                nb.epos[t] = code;
            }
        }
        __syncwarp();

        // per edge: sa = packed partial sums of SA (centre k folded into .x), mw = (MW_a, MW_j), m2 = MW_k
        float2 sa[4], mw[4];
        float m2[4], rowv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { sa[i] = make_float2(0.f, 0.f); mw[i] = sa[i]; m2[i] = 0.f; rowv[i] = 0.f; }
        // the owner's own share of dG: sum over its edges of Phi * u.  After the halves are combined both hold the same Phi:
        // half 0 keeps the share of its functions 0 and 1, half 1 that of 2 and 3.
        float go[2][3];
#pragma unroll
        for (int i = 0; i < 2; ++i) go[i][0] = go[i][1] = go[i][2] = 0.f;
        int cur_tj = -1;  // row slot of the edge being accumulated (-1: none yet)
        auto retire = [&]() {
            const float ea = nb.psi[cur_tj];
            float ph[4], ga[4];
            if (WITH_PHI) {
                const float dea = nb.dpsi[cur_tj], de2 = dek[cur_tj];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float p = zeta[i] * (sa[i].x + sa[i].y);
                    p = fmaf(dea, mw[i].x + mw[i].y, p);
                    p = fmaf(de2, m2[i], p);
                    ph[i] = p + __shfl_xor_sync(0xffffffffu, p, 16);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                rowv[i] = fmaf(ea, mw[i].x, rowv[i]);
                const float g = ea * mw[i].y;
                ga[i] = g + __shfl_xor_sync(0xffffffffu, g, 16);
            }
            if (WITH_PHI) {
                const Q4<float> u = ld4<float>(nb.u4 + 4 * cur_tj);
                const float inv = nb.invd[cur_tj];
                const float ux = u.a * inv, uy = u.b * inv, uz = u.c * inv;
                const float p0 = half ? ph[2] : ph[0], p1 = half ? ph[3] : ph[1];
                go[0][0] = fmaf(p0, ux, go[0][0]); go[0][1] = fmaf(p0, uy, go[0][1]); go[0][2] = fmaf(p0, uz, go[0][2]);
                go[1][0] = fmaf(p1, ux, go[1][0]); go[1][1] = fmaf(p1, uy, go[1][1]); go[1][2] = fmaf(p1, uz, go[1][2]);
            }
            const int eid = nb.epos[cur_tj];
            if (half == 0 && eid >= 0) {
                float *dst = a.E + (size_t(eid) * 2) * nfa + a.ebase + 4 * l16;
                // write-once, read-once data: streaming stores keep positions and neighbour rows in L2
                if (WITH_PHI) __stcs(reinterpret_cast<float4 *>(dst), make_float4(ph[0], ph[1], ph[2], ph[3]));
                __stcs(reinterpret_cast<float4 *>(dst + nfa), make_float4(ga[0], ga[1], ga[2], ga[3]));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { sa[i] = make_float2(0.f, 0.f); mw[i] = sa[i]; m2[i] = 0.f; }
        };
        int head = 0, qn = 0;
        int last_tj = -1;  // edge (row slot) of the last record already processed

        // process `cnt` queued (edge, third atom) pairs: phase 1 (records, up to two per lane) then phase 2 (functions)
        auto flush = [&](int cnt) {
            __syncwarp();
            const bool has0 = lane < cnt, has1 = EBATCH > 32 && lane + 32 < cnt;
            const QT pk0 = has0 ? queue[(head + lane) & 255] : QT(0), pk1 = has1 ? queue[(head + lane + 32) & 255] : QT(0);
            const int tj0 = has0 ? QPack<QT>::tj(pk0) : -2, tj1 = has1 ? QPack<QT>::tj(pk1) : -2;
#ifndef NNMD_EDGEM_MERGE
#define NNMD_EDGEM_MERGE 1
#endif
            int prev0 = __shfl_up_sync(0xffffffffu, tj0, 1), prev1 = __shfl_up_sync(0xffffffffu, tj1, 1);
            const int tj0_last = __shfl_sync(0xffffffffu, tj0, 31);
            if (lane == 0) { prev0 = last_tj; prev1 = tj0_last; }
            // the queue is edge-major: a change of tj opens the next edge
            const bool first0 = has0 && tj0 != prev0, first1 = has1 && tj1 != prev1;
#if NNMD_EDGEM_MERGE
            // pk = 0 names the staged neighbours (0, 0): finite garbage that is never stored
            make_record_m<WITH_PHI, LADDER>(nb, QPack<QT>::tj(pk0), QPack<QT>::tk(pk0), kind, step, inv_rc, eta, recs + lane * EM_STRIDE, has0);
            if (EBATCH > 32)
                make_record_m<WITH_PHI, LADDER>(nb, QPack<QT>::tj(pk1), QPack<QT>::tk(pk1), kind, step, inv_rc, eta, recs + (lane + 32) * EM_STRIDE, has1);
#else
            if (has0) make_record_m<WITH_PHI, LADDER>(nb, tj0, QPack<QT>::tk(pk0), kind, step, inv_rc, eta, recs + lane * EM_STRIDE, true);
            if (EBATCH > 32 && has1)
                make_record_m<WITH_PHI, LADDER>(nb, tj1, QPack<QT>::tk(pk1), kind, step, inv_rc, eta, recs + (lane + 32) * EM_STRIDE, true);
#endif
            last_tj = cnt > 32 ? __shfl_sync(0xffffffffu, tj1, cnt - 33) : __shfl_sync(0xffffffffu, tj0, cnt - 1);
            const unsigned long long firsts = (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, first1)) << 32) |
                                              __ballot_sync(0xffffffffu, first0);
            __syncwarp();
            // walk the batch edge by edge: `rest` holds the edge openings after record t (everything here is warp-uniform)
            unsigned long long rest = firsts & ~1ull;
            bool opens = (firsts & 1ull) != 0;
            int t = 0;
            while (t < cnt) {
                if (opens) {  // record t opens the next edge: retire the previous one
                    if (cur_tj >= 0) retire();
                    cur_tj = __shfl_sync(0xffffffffu, t < 32 ? tj0 : tj1, t & 31);
                }
                const int t_end = rest ? __ffsll(static_cast<long long>(rest)) - 1 : cnt;
                const float *r = recs + (t + half) * EM_STRIDE + class_off;
#pragma unroll NNMD_EDGEM_UNROLL_N
                for (int tt = t + half; tt < t_end; tt += 2, r += 2 * EM_STRIDE) {
                    const float4 v0 = *reinterpret_cast<const float4 *>(r);      // L0 L1 R0 R1
                    const float4 v1 = *reinterpret_cast<const float4 *>(r + 4);  // A0 A1 W0 W1
                    float2 q01;
                    float q2 = 0.f;
                    {
                        const float2 e01 = __fmul2_rn(make_float2(zm1[0], zm1[0]), make_float2(v0.x, v0.y));
                        q01.x = MF::exp2_fast(e01.x); q01.y = MF::exp2_fast(e01.y);
                    }
                    if (WITH_PHI) {
                        const float4 v2 = *reinterpret_cast<const float4 *>(r + 8);  // L2 R2 A2 W2
                        q2 = MF::exp2_fast(zm1[0] * v2.x);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (i > 0) {
                                if (LADDER) {
                                    q01 = __fmul2_rn(q01, make_float2(v0.z, v0.w));
                                    q2 *= v2.y;
                                } else {
                                    const float2 f01 = __fmul2_rn(make_float2(zm1[i], zm1[i]), make_float2(v0.x, v0.y));
                                    q01.x = MF::exp2_fast(f01.x); q01.y = MF::exp2_fast(f01.y);
                                    q2 = MF::exp2_fast(zm1[i] * v2.x);
                                }
                            }
                            sa[i] = __ffma2_rn(q01, make_float2(v1.x, v1.y), sa[i]);
                            sa[i].x = fmaf(q2, v2.z, sa[i].x);
                            mw[i] = __ffma2_rn(q01, make_float2(v1.z, v1.w), mw[i]);
                            m2[i] = fmaf(q2, v2.w, m2[i]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (i > 0) {
                                if (LADDER) {
                                    q01 = __fmul2_rn(q01, make_float2(v0.z, v0.w));
                                } else {
                                    const float2 f01 = __fmul2_rn(make_float2(zm1[i], zm1[i]), make_float2(v0.x, v0.y));
                                    q01.x = MF::exp2_fast(f01.x); q01.y = MF::exp2_fast(f01.y);
                                }
                            }
                            mw[i] = __ffma2_rn(q01, make_float2(v1.z, v1.w), mw[i]);
                        }
                    }
                }
                t = t_end;
                rest &= rest - 1ull;
                opens = true;
            }
            __syncwarp();
        };

        for (int tj = t_first; tj < nn; ++tj) {
            const int ep = nb.epos[tj];
            const Q4<float> uj = ld4<float>(nb.u4 + 4 * tj);
            int pushed_edge = 0;
            // 96 candidates per trip: three independent closure tests per lane (branch-free: the three loads are in
            // flight together), then three compactions into the ring; the ring (256) holds a partial batch (< 64) plus
            // one trip (<= 96)
            for (int kb = 0; kb < nn; kb += 96) {
                bool ok[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int tk = kb + 32 * c + lane;
                    const Q4<float> uk = ld4<float>(nb.u4 + 4 * min(tk, nn - 1));
                    const float wx = uj.a - uk.a, wy = uj.b - uk.b, wz = uj.c - uk.c;
                    const float z2 = wx * wx + wy * wy + wz * wz;
                    ok[c] = (tk < nn) && (z2 < rc2) && (z2 > 0.f);  // third side inside (0, rc) (symm_funcs.py:34-50); z2 = 0 for tk == tj
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const unsigned m = __ballot_sync(0xffffffffu, ok[c]);
                    if (ok[c]) queue[(head + qn + __popc(m & ((1u << lane) - 1u))) & 255] = QPack<QT>::pack(tj, kb + 32 * c + lane);
                    const int n_ok = __popc(m);
                    pushed_edge += n_ok;
                    qn += n_ok;
                }
                while (qn >= EBATCH) {
                    flush(EBATCH);
                    head = (head + EBATCH) & 255;
                    qn -= EBATCH;
                }
            }
            if (pushed_edge == 0 && half == 0 && ep >= 0) {  // an edge that closes no triangle still has a (zero) slot
                float *dst = a.E + (size_t(ep) * 2) * nfa + a.ebase + 4 * l16;
                if (WITH_PHI) *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4 *>(dst + nfa) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (qn > 0) flush(qn);
        if (cur_tj >= 0) retire();
        // the centre's own share of its value (raw sum; scaled and completed by the gather kernel)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v = rowv[i] + __shfl_xor_sync(0xffffffffu, rowv[i], 16);
            const int col = a.col[4 * l16 + i];
            if (a.G && half == 0 && col >= 0) a.G[row * a.nf + col] = v;
        }
        if (WITH_PHI) {
            float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int s = 4 * l16 + 2 * half + i;
                const int col = a.col[s];
                if (col < 0) continue;
                const float coef = a.coef[s];
                const int64_t o = row * a.nf + col;
                if (a.dG) { a.dG[3 * o] = -coef * go[i][0]; a.dG[3 * o + 1] = -coef * go[i][1]; a.dG[3 * o + 2] = -coef * go[i][2]; }
                if (a.force) {
                    const float wv = a.w[o] * coef;
                    fx = fmaf(wv, go[i][0], fx); fy = fmaf(wv, go[i][1], fy); fz = fmaf(wv, go[i][2], fz);
                }
            }
            if (a.force) {
                fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
                if (lane == 0) { a.force[3 * row] += fx; a.force[3 * row + 1] += fy; a.force[3 * row + 2] += fz; }
            }
        }
        __syncwarp();
    }
}
