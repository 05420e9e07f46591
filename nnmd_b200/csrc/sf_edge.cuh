// sf_edge.cuh -- edge-centric form of the angular kernel K3 (fp32, lambda = +-1, 33..64 functions: 4 per lane on 16 lanes).
// Included by sf.cu after the shared staging helpers.
//
// Why: with S_f(a,j,k) = sum over the three centres of P_f(c_X) E_X (one closed triangle),
//     dG[a,f] = -coef_f sum_{j in N(a)} Phi_f(a,j) u_aj ,   Phi_f(a,j) = sum_{k closes (a,j)} dS_f(a,j,k) / d d_aj ,
// and Phi_f(a,j) = Phi_f(j,a).  The vertex-centric kernel evaluates both derivatives (d/dx and d/dy) of every triangle
// from each of its three vertices; here every unordered edge is evaluated ONCE, by the endpoint with the smaller
// global index (its "owner"), for one derivative only: 11 FMA-pipe cycles per (edge, third atom, function) instead
// of 22, no direction-vector work in the inner loop.  The owner adds its own Phi*u terms to dG when an edge retires;
// Phi also goes to HBM as an edge array from which a second, bandwidth-bound kernel serves the other endpoint.  The value G_a needs the centre-a term of every ordered pair (j,k): the
// owner accumulates it in registers for its own edges and hands the other endpoint its share (Gamma) through the
// same edge array.  Still gather-only and atomic-free: the summation order is fixed.
//
// Edge scratch layout: E[slot][2][nfa] floats, [0] = Phi, [1] = Gamma (share of the NON-owner), function axis in lane
// order.  The slot is the position of the edge in the NON-owner's row: rows are ascending, so the entries of row j that
// j does not own (neighbours with a smaller index) are its first first_gt(j) entries, and
//     slot(a -> j) = (offsets[j] - edge_offsets[j]) + position of a in row j.
// The owner finds that position once per edge with a binary search while it stages its row; the gather of atom j then
// streams one contiguous block of its row's slots (no search, no scattered reads).  Edges towards halo atoms (which have
// no row of their own) get no slot.
#pragma once
// (included inside namespace nnmd)

#ifdef NNMD_TUNING
#define NNMD_EDGE_DBG(a) ((a).dbg)
#else
#define NNMD_EDGE_DBG(a) 0
#endif
struct EdgeArgs {
    const float *posw;
    int64_t n_rows, n_atoms, n_centres;
    const int64_t *offsets;
    const int32_t *idx;
    const int64_t *edge_offsets;
    int cap;
    int kind, ladder;
    float rc, eta, step;
    const float *zeta, *lam, *coef;
    const int *col;
    int nf;     // row stride of G / dG / w
    int nfa;    // function-axis length of the edge scratch
    int ebase;  // first slot of this group
    float *E;
    float *G, *dG;
    const float *w;
    float *force;
    uint32_t *flag;
    unsigned long long *row_counter;  // dynamic row scheduling (zeroed before the launch)
#ifdef NNMD_TUNING
    int dbg;                          // profiling knob (NNMD_DBG, tuning builds only): 1 = skip phase 2, 2 = skip phase 1 as well
#endif
};

// record: per lambda class (EREC_CLASS floats)  [ L0 L1 L2 eid | A0 A1 B0 B1 | A2 B2 V0a V0j | R0 R1 R2 . ]
template <bool WITH_PHI, bool LADDER>
__device__ __forceinline__ void make_record_edge(const NbView<float> &nb, int tj, int tk, int kind, float step,
                                                 float inv_rc, float eta, int eid, float *__restrict__ rec) {
    using MF = M<float>;
    const Q4<float> uj = ld4<float>(nb.u4 + 4 * tj), uk = ld4<float>(nb.u4 + 4 * tk);
    const float x = uj.d, y = uk.d, ix = nb.invd[tj], iy = nb.invd[tk];
    const float wx = uj.a - uk.a, wy = uj.b - uk.b, wz = uj.c - uk.c;
    const float z2 = wx * wx + wy * wy + wz * wz;
    const float iz = MF::rsqrt_fast(z2), z = z2 * iz;
    const float x2 = x * x, y2 = y * y;
    const float fz = 0.5f * (MF::cospi_fast(z * inv_rc) + 1.f);
    const float psz = MF::exp_fast(-eta * z2) * fz;
    const float psx = nb.psi[tj], psy = nb.psi[tk], dpsx = nb.dpsi[tj];
    float c[3], dcx[3], E[3], Ex[3];
    bool live[3];
    c[0] = cos_law<float>(x2, y2, z2, x, y, ix, iy, live[0]);  // centre a : sides x, y   opposite z
    c[1] = cos_law<float>(x2, z2, y2, x, z, ix, iz, live[1]);  // centre j : sides x, z   opposite y
    c[2] = cos_law<float>(y2, z2, x2, y, z, iy, iz, live[2]);  // centre k : sides y, z   opposite x
    dcx[0] = live[0] ? iy - c[0] * ix : 0.f;
    dcx[1] = live[1] ? iz - c[1] * ix : 0.f;
    dcx[2] = live[2] ? -x * iy * iz : 0.f;
    if (kind == NNMD_G4) {
        E[0] = E[1] = E[2] = psx * psy * psz;
        Ex[0] = Ex[1] = Ex[2] = dpsx * psy * psz;
    } else {  // G5 keeps fc of the opposite side
        const float fcx = nb.fc[tj], fcy = nb.fc[tk], dfcx = nb.dfc[tj];
        E[0] = psx * psy * fz;   Ex[0] = dpsx * psy * fz;
        E[1] = psx * psz * fcy;  Ex[1] = dpsx * psz * fcy;
        E[2] = psy * psz * fcx;  Ex[2] = psy * psz * dfcx;
    }
    float L[2][3], A[3], B[2][3], V[2][2];
#pragma unroll
    for (int X = 0; X < 3; ++X) {
        const float ac = fabsf(c[X]);
        const bool on = ac > 1e-3f;  // symm_funcs.py:210-211
        const float sg = c[X] > 0.f ? 1.f : -1.f;
        const float bp = 1.f + ac, bm = fmaxf(1.f - ac, 0.f);
        L[0][X] = on ? MF::log2_fast(bp) : 0.f;
        L[1][X] = on ? fmaxf(MF::log2_fast(bm), -MF::big()) : 0.f;
        A[X] = on ? sg * E[X] * dcx[X] : 0.f;
        B[0][X] = on ? bp * Ex[X] : 0.f;
        B[1][X] = on ? bm * Ex[X] : 0.f;
        if (X < 2) { V[0][X] = on ? bp * E[X] : 0.f; V[1][X] = on ? bm * E[X] : 0.f; }
    }
    const float eb = __int_as_float(eid);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        float *r = rec + s * EREC_CLASS;
        const float sgn = s ? -1.f : 1.f;
        st4<float>(r, L[s][0], L[s][1], L[s][2], eb);
        if (WITH_PHI) st4<float>(r + 4, sgn * A[0], sgn * A[1], B[s][0], B[s][1]);
        st4<float>(r + 8, sgn * A[2], B[s][2], V[s][0], V[s][1]);
        if (LADDER) st4<float>(r + 12, MF::exp2_fast(step * L[s][0]), MF::exp2_fast(step * L[s][1]), MF::exp2_fast(step * L[s][2]), 0.f);
    }
}

#ifndef NNMD_EDGE_PACK2
#define NNMD_EDGE_PACK2 0
#endif
#ifndef NNMD_EDGE_UNROLL
#define NNMD_EDGE_UNROLL 2
#endif
constexpr int NNMD_EDGE_UNROLL_N = NNMD_EDGE_UNROLL;
#ifndef NNMD_EDGE_EBATCH
#define NNMD_EDGE_EBATCH 64
#endif
constexpr int EBATCH = NNMD_EDGE_EBATCH;  // (edge, third atom) records built per flush: up to two per lane

// K3e, 4 functions per lane: 16 lanes cover the 64 functions, so the two half-warps work on two different records at
// once.  Per record pair the warp issues ~40 instructions, 3 ex2 and 44 FMA-pipe cycles: the FMA pipe alone is the
// limiter (the 2-per-lane form co-saturates FMA pipe, SFU and issue).  LADDER: q_{i+1} = q_i * R_X along the lane's 4 zetas.
#ifndef NNMD_EDGE_MINB
#define NNMD_EDGE_MINB 4
#endif
template <bool WITH_PHI, bool LADDER>
__global__ void __launch_bounds__(128, NNMD_EDGE_MINB) angular_edge4_kernel(EdgeArgs a) {
    if (a.flag != nullptr && (*a.flag & NNMD_FLAG_OVERFLOW)) return;  // incomplete neighbour list (capacity-bound build)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using MF = M<float>;
    const int lane = lane_id(), warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const size_t per_warp = size_t(EBATCH) * EREC_STRIDE * sizeof(float) + size_t(NB_FIELDS) * a.cap * sizeof(float) + 256 * sizeof(int);
    unsigned char *wbase = smem_raw + per_warp * warp;
    float *recs = reinterpret_cast<float *>(wbase);
    NbView<float> nb;
    nb.bind(recs + EBATCH * EREC_STRIDE, a.cap);
    int *queue = reinterpret_cast<int *>(recs + EBATCH * EREC_STRIDE + size_t(NB_FIELDS) * a.cap);  // ring of 256 packed (tj, tk)

    const float inv_rc = MF::rcp(a.rc), rc2 = a.rc * a.rc;
    const int half = lane >> 4, l16 = lane & 15;
    float zeta[4], zm1[4], coef[4];
    int col[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        zeta[i] = a.zeta[4 * l16 + i]; zm1[i] = zeta[i] - 1.f; col[i] = a.col[4 * l16 + i]; coef[i] = a.coef[4 * l16 + i];
    }
    const int class_off = (a.lam[4 * l16] < 0.f) ? EREC_CLASS : 0;  // lambda class of this lane's four functions
    const size_t nfa = size_t(a.nfa);
    const int kind = a.kind;
    const float eta = a.eta, step = a.step;

    // Rows cost between 0 and ~2x the mean (an atom owns only the edges towards larger indices), so warps pull
    // rows from a global counter instead of striding.
    (void)wpb;
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

        // phi*: two partial sums of Phi per function; vg*: (centre-a value of the row, Gamma of the current edge)
        float2 phi[4], vg[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { phi[i] = make_float2(0.f, 0.f); vg[i] = phi[i]; }
#if NNMD_EDGE_PACK2
        // centre-k terms of the lane's functions (0,1) and (2,3), accumulated as packed pairs: two FFMA2 instead of four FFMA
        float2 phik[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        const float2 zz01 = make_float2(zeta[0], zeta[1]), zz23 = make_float2(zeta[2], zeta[3]);
#endif
        // the owner's own share of dG: sum over its edges of Phi * u (the gather then only visits the other edges)
        float gox[4], goy[4], goz[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) gox[i] = goy[i] = goz[i] = 0.f;
        int cur_tj = -1;  // row slot of the edge being accumulated (-1: none yet)
        // the two half-warps hold partial sums of the same edge: combine, then half 0 writes 16 B per lane
        auto retire = [&](int eid) {
            float ph[4], ga[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ph[i] = phi[i].x + phi[i].y;
#if NNMD_EDGE_PACK2
                ph[i] += (i & 1) ? phik[i >> 1].y : phik[i >> 1].x;
#endif
                ph[i] += __shfl_xor_sync(0xffffffffu, ph[i], 16);
                ga[i] = vg[i].y + __shfl_xor_sync(0xffffffffu, vg[i].y, 16);
            }
            if (WITH_PHI) {
                const Q4<float> u = ld4<float>(nb.u4 + 4 * cur_tj);
                const float inv = nb.invd[cur_tj];
                const float ux = u.a * inv, uy = u.b * inv, uz = u.c * inv;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    gox[i] = fmaf(ph[i], ux, gox[i]); goy[i] = fmaf(ph[i], uy, goy[i]); goz[i] = fmaf(ph[i], uz, goz[i]);
                }
            }
            if (half == 0 && eid >= 0) {
                float *dst = a.E + (size_t(eid) * 2) * nfa + a.ebase + 4 * l16;
                // write-once, read-once data: streaming stores keep positions and neighbour rows in L2
                if (WITH_PHI) __stcs(reinterpret_cast<float4 *>(dst), make_float4(ph[0], ph[1], ph[2], ph[3]));
                __stcs(reinterpret_cast<float4 *>(dst + nfa), make_float4(ga[0], ga[1], ga[2], ga[3]));
            }
        };
        int cur_eid = -1;
        int head = 0, qn = 0;
        int last_tj = -1;  // edge (row slot) of the last record already processed

        // process `cnt` queued (edge, third atom) pairs: phase 1 (records, one per lane) then phase 2 (functions)
        auto flush = [&](int cnt) {
            __syncwarp();
            // phase 1: every lane builds up to two records (slots lane and lane + 32)
            const bool has0 = lane < cnt, has1 = EBATCH > 32 && lane + 32 < cnt;
            const int pk0 = has0 ? queue[(head + lane) & 255] : 0, pk1 = has1 ? queue[(head + lane + 32) & 255] : 0;
            const int tj0 = has0 ? (pk0 & 0x7fff) : -2, tj1 = has1 ? (pk1 & 0x7fff) : -2;
            int prev0 = __shfl_up_sync(0xffffffffu, tj0, 1), prev1 = __shfl_up_sync(0xffffffffu, tj1, 1);
            const int tj0_last = __shfl_sync(0xffffffffu, tj0, 31);
            if (lane == 0) { prev0 = last_tj; prev1 = tj0_last; }
            // the queue is edge-major: a change of tj opens the next edge
            const bool first0 = has0 && tj0 != prev0, first1 = has1 && tj1 != prev1;
            if (!(NNMD_EDGE_DBG(a) & 2)) {
                if (has0) make_record_edge<WITH_PHI, LADDER>(nb, tj0, (pk0 >> 15) & 0x7fff, kind, step, inv_rc, eta, nb.epos[tj0],
                                                             recs + lane * EREC_STRIDE);
                if (EBATCH > 32 && has1)
                    make_record_edge<WITH_PHI, LADDER>(nb, tj1, (pk1 >> 15) & 0x7fff, kind, step, inv_rc, eta, nb.epos[tj1],
                                                       recs + (lane + 32) * EREC_STRIDE);
            }
            last_tj = cnt > 32 ? __shfl_sync(0xffffffffu, tj1, cnt - 33) : __shfl_sync(0xffffffffu, tj0, cnt - 1);
            const unsigned long long firsts = (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, first1)) << 32) |
                                              __ballot_sync(0xffffffffu, first0);
            __syncwarp();
            // walk the batch edge by edge: `rest` holds the edge openings after record t (everything here is warp-uniform)
            unsigned long long rest = firsts & ~1ull;
            bool opens = (firsts & 1ull) != 0;
            int t = (NNMD_EDGE_DBG(a) & 1) ? cnt : 0;
            while (t < cnt) {
                if (opens) {  // record t opens the next edge: retire the previous one
                    if (cur_tj >= 0) retire(cur_eid);
                    cur_eid = __float_as_int(recs[t * EREC_STRIDE + 3]);
                    cur_tj = __shfl_sync(0xffffffffu, t < 32 ? tj0 : tj1, t & 31);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { phi[i] = make_float2(0.f, 0.f); vg[i].y = 0.f; }
#if NNMD_EDGE_PACK2
                    phik[0] = phik[1] = make_float2(0.f, 0.f);
#endif
                }
                const int t_end = rest ? __ffsll(static_cast<long long>(rest)) - 1 : cnt;
                const float *r = recs + (t + half) * EREC_STRIDE + class_off;
#pragma unroll NNMD_EDGE_UNROLL_N
                for (int tt = t + half; tt < t_end; tt += 2, r += 2 * EREC_STRIDE) {
                    const float4 l = *reinterpret_cast<const float4 *>(r);       // L0 L1 L2 .
                    const float4 c2 = *reinterpret_cast<const float4 *>(r + 8);  // A2 B2 V0a V0j
                    float2 q01[4];
                    float q2[4];
                    const float2 e01 = __fmul2_rn(make_float2(zm1[0], zm1[0]), make_float2(l.x, l.y));
                    q01[0].x = MF::exp2_fast(e01.x); q01[0].y = MF::exp2_fast(e01.y);
                    q2[0] = MF::exp2_fast(zm1[0] * l.z);
                    if (LADDER) {
                        const float4 rr = *reinterpret_cast<const float4 *>(r + 12);  // R0 R1 R2 .
#pragma unroll
                        for (int i = 1; i < 4; ++i) {
                            q01[i] = __fmul2_rn(q01[i - 1], make_float2(rr.x, rr.y));
#if !NNMD_EDGE_PACK2
                            q2[i] = q2[i - 1] * rr.z;
#endif
                        }
#if NNMD_EDGE_PACK2
                        q2[1] = q2[0] * rr.z;
                        const float rz2 = rr.z * rr.z;
                        const float2 qk23 = __fmul2_rn(make_float2(q2[0], q2[1]), make_float2(rz2, rz2));
                        q2[2] = qk23.x; q2[3] = qk23.y;
#endif
                    } else {
#pragma unroll
                        for (int i = 1; i < 4; ++i) {
                            const float2 f01 = __fmul2_rn(make_float2(zm1[i], zm1[i]), make_float2(l.x, l.y));
                            q01[i].x = MF::exp2_fast(f01.x); q01[i].y = MF::exp2_fast(f01.y);
                            q2[i] = MF::exp2_fast(zm1[i] * l.z);
                        }
                    }
                    if (WITH_PHI) {
                        const float4 c1 = *reinterpret_cast<const float4 *>(r + 4);  // A0 A1 B0 B1
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 t01 = __ffma2_rn(make_float2(zeta[i], zeta[i]), make_float2(c1.x, c1.y), make_float2(c1.z, c1.w));
                            phi[i] = __ffma2_rn(q01[i], t01, phi[i]);
#if !NNMD_EDGE_PACK2
                            const float t2 = fmaf(zeta[i], c2.x, c2.y);
                            phi[i].x = fmaf(q2[i], t2, phi[i].x);
#endif
                        }
#if NNMD_EDGE_PACK2
                        const float2 a2 = make_float2(c2.x, c2.x), b2 = make_float2(c2.y, c2.y);
                        phik[0] = __ffma2_rn(make_float2(q2[0], q2[1]), __ffma2_rn(zz01, a2, b2), phik[0]);
                        phik[1] = __ffma2_rn(make_float2(q2[2], q2[3]), __ffma2_rn(zz23, a2, b2), phik[1]);
#endif
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) vg[i] = __ffma2_rn(q01[i], make_float2(c2.z, c2.w), vg[i]);
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
                    if (ok[c]) queue[(head + qn + __popc(m & ((1u << lane) - 1u))) & 255] = tj | ((kb + 32 * c + lane) << 15);
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
        if (cur_tj >= 0) retire(cur_eid);
        // the centre's own share of its value (raw sum; scaled and completed by the gather kernel)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v = vg[i].x + __shfl_xor_sync(0xffffffffu, vg[i].x, 16);
            if (a.G && half == 0 && col[i] >= 0) a.G[row * a.nf + col[i]] = v;
        }
        if (WITH_PHI) {
            float fx = 0.f, fy = 0.f, fz = 0.f;
            if (half == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (col[i] < 0) continue;
                    const int64_t o = row * a.nf + col[i];
                    if (a.dG) { a.dG[3 * o] = -coef[i] * gox[i]; a.dG[3 * o + 1] = -coef[i] * goy[i]; a.dG[3 * o + 2] = -coef[i] * goz[i]; }
                    if (a.force) {
                        const float wv = a.w[o] * coef[i];
                        fx = fmaf(wv, gox[i], fx); fy = fmaf(wv, goy[i], fy); fz = fmaf(wv, goz[i], fz);
                    }
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

// K4e: every centre collects Phi (and Gamma) of the edges it does NOT own -- the first first_gt entries of its row,
// whose slots are one contiguous block of E.  HBM-bound streaming.  The unit of work is one 256 B row of E (the Phi or
// the Gamma row of an edge): a half-warp reads it with one 16 B load per lane, the two half-warps take alternate rows
// (MODE_DG: half 0 the Phi row, half 1 the Gamma row of the same edge), GATHER_ROWS rows per half are requested before
// any is used -- the kernel lives on memory-level parallelism.  The centre's own edges were already added to
// dG / force by K3e.
#ifndef NNMD_GATHER_ROWS
#define NNMD_GATHER_ROWS 4
#endif
constexpr int GATHER_ROWS = NNMD_GATHER_ROWS;
template <int MODE>
#ifndef NNMD_GATHER_MINB
#define NNMD_GATHER_MINB 1
#endif
__global__ void __launch_bounds__(256, NNMD_GATHER_MINB) edge_gather_kernel(EdgeArgs a) {
    if (a.flag != nullptr && (*a.flag & NNMD_FLAG_OVERFLOW)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id(), warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    float *ubuf = reinterpret_cast<float *>(smem_raw) + size_t(warp) * 4 * a.cap;  // (ux/d, uy/d, uz/d, slot bits)
    constexpr int PER = (MODE == MODE_DG) ? 2 : 1;  // rows of E read per edge
    const size_t nfa = size_t(a.nfa);
    // this lane's row of an edge slot: Phi at +0, Gamma at +nfa
    const size_t lane_off = size_t(a.ebase) + 4 * l16 + ((MODE == MODE_G || (MODE == MODE_DG && half == 1)) ? nfa : 0);
    const bool phi_lane = MODE == MODE_FORCE || (MODE == MODE_DG && half == 0);
    bool bad = false;
    for (int64_t row = int64_t(blockIdx.x) * wpb + warp; row < a.n_rows; row += int64_t(gridDim.x) * wpb) {
        const int64_t frame = row / a.n_centres;
        const int64_t gi = frame * a.n_atoms + (row - frame * a.n_centres);
        float xa, ya, za;
        load_pos<float>(a.posw, gi, xa, ya, za);
        const int64_t o0 = a.offsets[row], o1 = a.offsets[row + 1];
        const int64_t e0 = a.edge_offsets[row], e1 = a.edge_offsets[row + 1];
        const int64_t first_gt = (o1 - o0) - (e1 - e0);
        const int64_t slot0 = o0 - e0;  // first slot of this row's block in E
        int nn = 0;
        for (int64_t base = 0; base < first_gt; base += 32) {
            const int64_t p = base + lane;
            bool keep = false;
            float ux = 0, uy = 0, uz = 0, d = 0;
            if (p < first_gt) {
                const int j = __ldg(a.idx + o0 + p);
                float xj, yj, zj;
                load_pos<float>(a.posw, j, xj, yj, zj);
                ux = xj - xa; uy = yj - ya; uz = zj - za;
                d = norm3<float>(ux, uy, uz);  // the same expression as stage_row: both kernels see the same edge set
                keep = (d < a.rc) && (d > 0.f);
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int q = nn + __popc(m & ((1u << lane) - 1u));
                if (q < a.cap) {
                    const float inv = 1.0f / d;
                    *reinterpret_cast<float4 *>(ubuf + 4 * q) = make_float4(ux * inv, uy * inv, uz * inv, __int_as_float(int(slot0 + p)));
                }
            }
            nn += __popc(m);
        }
        __syncwarp();
        nn = min(nn, a.cap);
        float gx[4] = {0, 0, 0, 0}, gy[4] = {0, 0, 0, 0}, gz[4] = {0, 0, 0, 0}, gam[4] = {0, 0, 0, 0};
        const int n_items = nn * PER;
        for (int i0 = 0; i0 < n_items; i0 += 2 * GATHER_ROWS) {
            float4 v[GATHER_ROWS];
#pragma unroll
            for (int k = 0; k < GATHER_ROWS; ++k) {
                const int item = min(i0 + 2 * k + half, n_items - 1);
                const int slot = __float_as_int(ubuf[4 * (item / PER) + 3]);
                v[k] = __ldcs(reinterpret_cast<const float4 *>(a.E + (size_t(slot) * 2) * nfa + lane_off));
            }
#pragma unroll
            for (int k = 0; k < GATHER_ROWS; ++k) {
                const int item = i0 + 2 * k + half;
                if (item < n_items) {
                    if (phi_lane) {
                        const float4 u = *reinterpret_cast<const float4 *>(ubuf + 4 * (item / PER));
                        gx[0] = fmaf(v[k].x, u.x, gx[0]); gy[0] = fmaf(v[k].x, u.y, gy[0]); gz[0] = fmaf(v[k].x, u.z, gz[0]);
                        gx[1] = fmaf(v[k].y, u.x, gx[1]); gy[1] = fmaf(v[k].y, u.y, gy[1]); gz[1] = fmaf(v[k].y, u.z, gz[1]);
                        gx[2] = fmaf(v[k].z, u.x, gx[2]); gy[2] = fmaf(v[k].z, u.y, gy[2]); gz[2] = fmaf(v[k].z, u.z, gz[2]);
                        gx[3] = fmaf(v[k].w, u.x, gx[3]); gy[3] = fmaf(v[k].w, u.y, gy[3]); gz[3] = fmaf(v[k].w, u.z, gz[3]);
                    } else {
                        gam[0] += v[k].x; gam[1] += v[k].y; gam[2] += v[k].z; gam[3] += v[k].w;
                    }
                }
            }
        }
        // bring the two half-warps together: half 0 ends up with the complete sums of its four functions
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (MODE == MODE_FORCE) {
                gx[i] += __shfl_xor_sync(0xffffffffu, gx[i], 16);
                gy[i] += __shfl_xor_sync(0xffffffffu, gy[i], 16);
                gz[i] += __shfl_xor_sync(0xffffffffu, gz[i], 16);
            } else {
                gam[i] += __shfl_xor_sync(0xffffffffu, gam[i], 16);
            }
        }
        float fx = 0, fy = 0, fz = 0;
        if (half == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int col = a.col[4 * l16 + i];
                if (col < 0) continue;
                const float coef = a.coef[4 * l16 + i];
                const int64_t o = row * a.nf + col;
                if (MODE != MODE_FORCE) a.G[o] = 0.5f * coef * (a.G[o] + gam[i]);  // ordered (j,k): both edges of a pair counted
                if (MODE == MODE_DG) {  // add to the share K3e wrote for the centre's own edges
                    const float dx = a.dG[3 * o] - coef * gx[i], dy = a.dG[3 * o + 1] - coef * gy[i], dz = a.dG[3 * o + 2] - coef * gz[i];
                    a.dG[3 * o] = dx; a.dG[3 * o + 1] = dy; a.dG[3 * o + 2] = dz;
                    bad |= !(isfinite(dx) && isfinite(dy) && isfinite(dz));
                }
                if (MODE == MODE_FORCE) {
                    const float wv = a.w[o] * coef;
                    fx = fmaf(wv, gx[i], fx); fy = fmaf(wv, gy[i], fy); fz = fmaf(wv, gz[i], fz);
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

#include "sf_edge_m.cuh"

// Form of K3e: 1 (default) = moment form (sf_edge_m.cuh, 9 FMA-pipe lane-ops per record and function), 0 = the round-1
// derivative-record form above (11), kept for A/B measurements.  nnmd_set_option("edge_form", 0|1).
static int g_edge_form = 1;

template <int MODE>
static int launch_edge(const AngGroup &g, EdgeArgs a, cudaStream_t st) {
    constexpr bool WITH_PHI = MODE != MODE_G;
    int form = g_edge_form;
    if (const char *env = tuning_env("NNMD_EDGE_FORM")) form = atoi(env);  // tuning builds only
    NNMD_CUDA_CHECK(cudaMemsetAsync(a.row_counter, 0, sizeof(unsigned long long), st));
    {
        const size_t per_warp = form ? edgem_per_warp_bytes(a.cap)
                                     : size_t(EBATCH) * EREC_STRIDE * sizeof(float) + size_t(NB_FIELDS) * a.cap * sizeof(float) + 256 * sizeof(int);
        int wpb = 4;
        while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
        const size_t smem = per_warp * wpb;
        if (smem > 220 * 1024) return set_err(NNMD_ERR_CAPACITY, "angular: neighbour row of %d entries exceeds shared memory", a.cap);
        void (*kern)(EdgeArgs);
        if (form && a.cap <= 256) kern = a.ladder ? angular_edgem_kernel<WITH_PHI, true, uint16_t> : angular_edgem_kernel<WITH_PHI, false, uint16_t>;
        else if (form) kern = a.ladder ? angular_edgem_kernel<WITH_PHI, true, uint32_t> : angular_edgem_kernel<WITH_PHI, false, uint32_t>;
        else kern = a.ladder ? angular_edge4_kernel<WITH_PHI, true> : angular_edge4_kernel<WITH_PHI, false>;
        NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
#ifdef NNMD_TUNING
        if (const char *env = tuning_env("NNMD_CARVEOUT")) NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(env)));
#endif
        int occ = 1;
        NNMD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
        if (occ < 1) occ = 1;
        int64_t blocks = std::min<int64_t>((a.n_rows + wpb - 1) / wpb, int64_t(sm_count()) * occ);
#ifdef NNMD_TUNING
        if (const char *env = tuning_env("NNMD_DBG")) a.dbg = atoi(env);
        if (tuning_env("NNMD_VERBOSE")) fprintf(stderr, "[nnmd] K3e form %d cap %d smem %zu B/block, %d blocks/SM\n", form, a.cap, smem, occ);
#endif
        ProfScope prof(PROF_ANGULAR, st);
        kern<<<unsigned(blocks), wpb * 32, smem, st>>>(a);
        NNMD_LAUNCH_CHECK("angular_edge_kernel");
    }
    {
#ifndef NNMD_GATHER_WPB
#define NNMD_GATHER_WPB 8
#endif
        int wpb = NNMD_GATHER_WPB;
        while (wpb > 1 && size_t(wpb) * 16 * a.cap > 160 * 1024) wpb >>= 1;
        const size_t smem = size_t(wpb) * 16 * a.cap;
        auto kern = edge_gather_kernel<MODE>;
        NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
#ifndef NNMD_GATHER_BPS
#define NNMD_GATHER_BPS 16
#endif
        int64_t blocks = std::min<int64_t>((a.n_rows + wpb - 1) / wpb, int64_t(sm_count()) * NNMD_GATHER_BPS);
        ProfScope prof(PROF_GATHER, st);
        kern<<<unsigned(blocks), wpb * 32, smem, st>>>(a);
        NNMD_LAUNCH_CHECK("edge_gather_kernel");
    }
    (void)g;
    return NNMD_OK;
}
