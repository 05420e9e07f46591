// sf_edge_ws.cuh -- warp-specialised form of the edge-centric angular kernel K3e (included by sf.cu after sf_edge.cuh).
//
// Why: in angular_edge4_kernel every warp alternates between two kinds of work -- integer / SFU / shared-memory heavy
// bookkeeping (row staging, slot search, triangle enumeration, record construction: ~45 % of the instructions, next to no
// FMA-pipe use) and the FMA-pipe-bound inner loop over (record, function).  With four such warps per scheduler the two
// phases overlap only by chance: ncu shows the FMA pipe busy 57 % of the time and the issue slots 69 % (profiles/r1_k3e_*).
// Here the two kinds of work live in DIFFERENT warps of a block that run concurrently:
//
//   producer warp p (warps 4..7)   pulls rows from the global counter, stages the neighbour row, finds the edge slots,
//                                  enumerates the closing third atoms and builds the records of a batch (64 (edge, third
//                                  atom) pairs) into one of two shared-memory buffers
//   consumer warp p (warps 0..3)   walks the batch edge by edge: the FMA inner loop (16 lanes x 4 functions, two records in
//                                  flight), retires finished edges to the edge scratch and writes the row's own share of
//                                  G / dG / force when the row ends
//
// A producer/consumer pair hands batches over through a two-slot ring guarded by mbarriers (full / empty per slot): the
// consumer never touches the staged row, the record is self-contained (it carries the edge's slot id and unit vector), so
// the producer stages the NEXT row while the consumer still works on the last batches of the current one.  Each scheduler
// of an SM then holds two consumers (FMA pipe) and two producers (ALU / SFU / LSU) at all times.
// Arithmetic, record layout, summation order within a row and the edge-scratch layout are those of angular_edge4_kernel:
// the results are bit-identical to it.
#pragma once
// (included inside namespace nnmd)

constexpr int WS_PAIRS = 4;        // producer/consumer pairs per block (8 warps)
constexpr int WS_FLAG_ROW_END = 1;  // the row ends with this batch: retire the open edge, write the row's outputs
constexpr int WS_FLAG_TERMINATE = 2;
constexpr int WS_TAIL = 2 * EREC_CLASS;  // record floats [32..35]: unit vector of the edge (x, y, z), spare

struct WsBatchHdr {
    int cnt;
    int flags;
    unsigned long long firsts;  // bit t: record t opens a new edge
    long long row;
};

__device__ __forceinline__ uint32_t ws_smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ws_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem_addr(bar)) : "memory");
}
#ifndef NNMD_WS_SLEEP
#define NNMD_WS_SLEEP 0   // ns of back-off between failed probes of a barrier (0: none)
#endif
#ifndef NNMD_WS_UNROLL
#define NNMD_WS_UNROLL 2  // record pairs in flight in the consumer's inner loop
#endif
constexpr int WS_UNROLL = NNMD_WS_UNROLL;
__device__ __forceinline__ void ws_mbar_wait(uint64_t *bar, unsigned parity) {
    const uint32_t addr = ws_smem_addr(bar);
    uint32_t done = 0;
    while (!done) {  // try_wait suspends the warp for a hardware-chosen time before it reports failure: not a hot spin
        if (NNMD_WS_SLEEP > 0) __nanosleep(NNMD_WS_SLEEP);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}

__host__ __device__ inline size_t ws_pair_bytes(int cap) {
    return size_t(2) * EBATCH * EREC_STRIDE * sizeof(float) + size_t(NB_FIELDS) * cap * sizeof(float) + 256 * sizeof(int) +
           2 * sizeof(WsBatchHdr) + 4 * sizeof(uint64_t);
}

// ------------------------------------------------------------------------------------------------- producer
template <bool WITH_PHI, bool LADDER>
__device__ __forceinline__ void ws_producer(const EdgeArgs &a, float *__restrict__ recs2, NbView<float> nb, int *__restrict__ queue,
                                            WsBatchHdr *__restrict__ hdr, uint64_t *full, uint64_t *empty) {
    using MF = M<float>;
    const int lane = lane_id(), half = lane >> 4, l16 = lane & 15;
    const float inv_rc = MF::rcp(a.rc), rc2 = a.rc * a.rc;
    const size_t nfa = size_t(a.nfa);
    const int kind = a.kind;
    const float eta = a.eta, step = a.step;
    int buf = 0;
    unsigned phase = 0;  // flips every time the ring wraps
    int last_tj = -1;    // edge (row slot) of the last record handed over for the current row

    // hand `cnt` queued (edge, third atom) pairs (queue positions head ..) over as one batch
    auto emit = [&](int cnt, int head, long long row, int flags) {
        ws_mbar_wait(&empty[buf], phase ^ 1u);  // the consumer has finished with this slot
        float *recs = recs2 + size_t(buf) * EBATCH * EREC_STRIDE;
        unsigned long long firsts = 0ull;
        if (cnt > 0) {
            const bool has0 = lane < cnt, has1 = EBATCH > 32 && lane + 32 < cnt;
            const int pk0 = has0 ? queue[(head + lane) & 255] : 0, pk1 = has1 ? queue[(head + lane + 32) & 255] : 0;
            const int tj0 = has0 ? (pk0 & 0x7fff) : -2, tj1 = has1 ? (pk1 & 0x7fff) : -2;
            int prev0 = __shfl_up_sync(0xffffffffu, tj0, 1), prev1 = __shfl_up_sync(0xffffffffu, tj1, 1);
            const int tj0_last = __shfl_sync(0xffffffffu, tj0, 31);
            if (lane == 0) { prev0 = last_tj; prev1 = tj0_last; }
            // the queue is edge-major: a change of tj opens the next edge
            const bool first0 = has0 && tj0 != prev0, first1 = has1 && tj1 != prev1;
            if (has0) {
                float *r = recs + lane * EREC_STRIDE;
                make_record_edge<WITH_PHI, LADDER>(nb, tj0, (pk0 >> 15) & 0x7fff, kind, step, inv_rc, eta, nb.epos[tj0], r);
                const Q4<float> u = ld4<float>(nb.u4 + 4 * tj0);
                const float inv = nb.invd[tj0];
                st4<float>(r + WS_TAIL, u.a * inv, u.b * inv, u.c * inv, 0.f);
            }
            if (EBATCH > 32 && has1) {
                float *r = recs + (lane + 32) * EREC_STRIDE;
                make_record_edge<WITH_PHI, LADDER>(nb, tj1, (pk1 >> 15) & 0x7fff, kind, step, inv_rc, eta, nb.epos[tj1], r);
                const Q4<float> u = ld4<float>(nb.u4 + 4 * tj1);
                const float inv = nb.invd[tj1];
                st4<float>(r + WS_TAIL, u.a * inv, u.b * inv, u.c * inv, 0.f);
            }
            last_tj = cnt > 32 ? __shfl_sync(0xffffffffu, tj1, cnt - 33) : __shfl_sync(0xffffffffu, tj0, cnt - 1);
            firsts = (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, first1)) << 32) | __ballot_sync(0xffffffffu, first0);
        }
        if (lane == 0) { hdr[buf].cnt = cnt; hdr[buf].flags = flags; hdr[buf].firsts = firsts; hdr[buf].row = row; }
        __syncwarp();
        if (lane == 0) ws_mbar_arrive(&full[buf]);
        buf ^= 1;
        if (buf == 0) phase ^= 1u;
    };

    for (;;) {
        unsigned long long next = 0;
        if (lane == 0) next = atomicAdd(a.row_counter, 1ull);
        const int64_t row = int64_t(__shfl_sync(0xffffffffu, next, 0));
        if (row >= a.n_rows) {
            emit(0, 0, -1, WS_FLAG_TERMINATE);
            break;
        }
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
        last_tj = -1;
        int head = 0, qn = 0;
        for (int tj = t_first; tj < nn; ++tj) {
            const int ep = nb.epos[tj];
            const Q4<float> uj = ld4<float>(nb.u4 + 4 * tj);
            int pushed_edge = 0;
            // 96 candidates per trip: three branch-free closure tests per lane, then three compactions into the ring
            // (256 entries: a partial batch (< 64) plus one trip (<= 96))
            for (int kb = 0; kb < nn; kb += 96) {
                bool ok[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int tk = kb + 32 * c + lane;
                    const Q4<float> uk = ld4<float>(nb.u4 + 4 * min(tk, nn - 1));
                    const float wx = uj.a - uk.a, wy = uj.b - uk.b, wz = uj.c - uk.c;
                    const float z2 = wx * wx + wy * wy + wz * wz;
                    ok[c] = (tk < nn) && (z2 < rc2) && (z2 > 0.f);  // third side inside (0, rc) (symm_funcs.py:34-50)
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const unsigned m = __ballot_sync(0xffffffffu, ok[c]);
                    if (ok[c]) queue[(head + qn + __popc(m & ((1u << lane) - 1u))) & 255] = tj | ((kb + 32 * c + lane) << 15);
                    const int n_ok = __popc(m);
                    pushed_edge += n_ok;
                    qn += n_ok;
                }
                __syncwarp();
                while (qn >= EBATCH) {
                    emit(EBATCH, head, row, 0);
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
        __syncwarp();
        emit(qn, head, row, WS_FLAG_ROW_END);  // the last (possibly empty) batch of the row
        // the records of that batch still reference the staged row: it must not be overwritten before they are BUILT,
        // which `emit` has done by now (the consumer reads only the records)
    }
}

// ------------------------------------------------------------------------------------------------- consumer
template <bool WITH_PHI, bool LADDER>
__device__ __forceinline__ void ws_consumer(const EdgeArgs &a, const float *__restrict__ recs2, const WsBatchHdr *__restrict__ hdr,
                                            uint64_t *full, uint64_t *empty) {
    using MF = M<float>;
    const int lane = lane_id(), half = lane >> 4, l16 = lane & 15;
    float zeta[4], zm1[4], coef[4];
    int col[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        zeta[i] = a.zeta[4 * l16 + i]; zm1[i] = zeta[i] - 1.f; col[i] = a.col[4 * l16 + i]; coef[i] = a.coef[4 * l16 + i];
    }
    const int class_off = (a.lam[4 * l16] < 0.f) ? EREC_CLASS : 0;  // lambda class of this lane's four functions
    const size_t nfa = size_t(a.nfa);

    // phi: two partial sums of Phi per function; vg: (centre-a value of the row, Gamma of the current edge)
    float2 phi[4], vg[4];
    float gox[4], goy[4], goz[4];  // the owner's own share of dG: sum over its edges of Phi * u
#pragma unroll
    for (int i = 0; i < 4; ++i) { phi[i] = make_float2(0.f, 0.f); vg[i] = phi[i]; gox[i] = goy[i] = goz[i] = 0.f; }
    bool has_edge = false;
    int cur_eid = -1;
    float ux = 0.f, uy = 0.f, uz = 0.f;  // unit vector of the edge being accumulated

    // the two half-warps hold partial sums of the same edge: combine, then half 0 writes 16 B per lane
    auto retire = [&]() {
        float ph[4], ga[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ph[i] = phi[i].x + phi[i].y;
            ph[i] += __shfl_xor_sync(0xffffffffu, ph[i], 16);
            ga[i] = vg[i].y + __shfl_xor_sync(0xffffffffu, vg[i].y, 16);
        }
        if (WITH_PHI) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                gox[i] = fmaf(ph[i], ux, gox[i]); goy[i] = fmaf(ph[i], uy, goy[i]); goz[i] = fmaf(ph[i], uz, goz[i]);
            }
        }
        if (half == 0 && cur_eid >= 0) {
            float *dst = a.E + (size_t(cur_eid) * 2) * nfa + a.ebase + 4 * l16;
            // write-once, read-once data: streaming stores keep positions and neighbour rows in L2
            if (WITH_PHI) __stcs(reinterpret_cast<float4 *>(dst), make_float4(ph[0], ph[1], ph[2], ph[3]));
            __stcs(reinterpret_cast<float4 *>(dst + nfa), make_float4(ga[0], ga[1], ga[2], ga[3]));
        }
    };

    int buf = 0;
    unsigned phase = 0;
    for (;;) {
        ws_mbar_wait(&full[buf], phase);
        const int cnt = hdr[buf].cnt, flags = hdr[buf].flags;
        const unsigned long long firsts = hdr[buf].firsts;
        const int64_t row = hdr[buf].row;
        const float *recs = recs2 + size_t(buf) * EBATCH * EREC_STRIDE;
        if (cnt > 0) {
            // walk the batch edge by edge: `rest` holds the edge openings after record t (everything here is warp-uniform)
            unsigned long long rest = firsts & ~1ull;
            bool opens = (firsts & 1ull) != 0;
            int t = 0;
            while (t < cnt) {
                if (opens) {  // record t opens the next edge: retire the previous one
                    if (has_edge) retire();
                    has_edge = true;
                    cur_eid = __float_as_int(recs[t * EREC_STRIDE + 3]);
                    const float4 tl = *reinterpret_cast<const float4 *>(recs + t * EREC_STRIDE + WS_TAIL);
                    ux = tl.x; uy = tl.y; uz = tl.z;
#pragma unroll
                    for (int i = 0; i < 4; ++i) { phi[i] = make_float2(0.f, 0.f); vg[i].y = 0.f; }
                }
                const int t_end = rest ? __ffsll(static_cast<long long>(rest)) - 1 : cnt;
                const float *r = recs + (t + half) * EREC_STRIDE + class_off;
#pragma unroll WS_UNROLL
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
                            q2[i] = q2[i - 1] * rr.z;
                        }
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
                            const float t2 = fmaf(zeta[i], c2.x, c2.y);
                            phi[i] = __ffma2_rn(q01[i], t01, phi[i]);
                            phi[i].x = fmaf(q2[i], t2, phi[i].x);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) vg[i] = __ffma2_rn(q01[i], make_float2(c2.z, c2.w), vg[i]);
                }
                t = t_end;
                rest &= rest - 1ull;
                opens = true;
            }
        }
        if (flags & WS_FLAG_ROW_END) {
            if (has_edge) retire();
            has_edge = false;
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
#pragma unroll
            for (int i = 0; i < 4; ++i) { phi[i] = make_float2(0.f, 0.f); vg[i] = phi[i]; gox[i] = goy[i] = goz[i] = 0.f; }
        }
        __syncwarp();
        if (lane == 0) ws_mbar_arrive(&empty[buf]);  // every lane has read what it needs from this slot
        if (flags & WS_FLAG_TERMINATE) break;
        buf ^= 1;
        if (buf == 0) phase ^= 1u;
    }
}

template <bool WITH_PHI, bool LADDER>
__global__ void __launch_bounds__(2 * WS_PAIRS * 32, 2) angular_edge4_ws_kernel(EdgeArgs a) {
    if (a.flag != nullptr && (*a.flag & NNMD_FLAG_OVERFLOW)) return;  // incomplete neighbour list (capacity-bound build)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int pair = warp & (WS_PAIRS - 1);
    const bool is_producer = warp >= WS_PAIRS;
    unsigned char *base = smem_raw + ws_pair_bytes(a.cap) * pair;
    float *recs2 = reinterpret_cast<float *>(base);
    float *nbase = recs2 + 2 * EBATCH * EREC_STRIDE;
    NbView<float> nb;
    nb.bind(nbase, a.cap);
    int *queue = reinterpret_cast<int *>(nbase + size_t(NB_FIELDS) * a.cap);
    WsBatchHdr *hdr = reinterpret_cast<WsBatchHdr *>(queue + 256);
    uint64_t *bars = reinterpret_cast<uint64_t *>(hdr + 2);  // full[0], full[1], empty[0], empty[1]
    if (threadIdx.x < WS_PAIRS) {
        uint64_t *b = reinterpret_cast<uint64_t *>(smem_raw + ws_pair_bytes(a.cap) * threadIdx.x + size_t(2) * EBATCH * EREC_STRIDE * sizeof(float) +
                                                   size_t(NB_FIELDS) * a.cap * sizeof(float) + 256 * sizeof(int) + 2 * sizeof(WsBatchHdr));
        for (int k = 0; k < 4; ++k) ws_mbar_init(b + k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (is_producer) ws_producer<WITH_PHI, LADDER>(a, recs2, nb, queue, hdr, bars, bars + 2);
    else ws_consumer<WITH_PHI, LADDER>(a, recs2, hdr, bars, bars + 2);
}
