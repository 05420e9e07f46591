// nlist.cu -- K0 (wrap-only PBC) and K1 (cell-list neighbour list) for sm_100a.
//
// Replaces the reference's dense O(N^2) distance matrix and pair mask
//   nnmd/features/symm_funcs.py:53-82   map_positions_pbc   (wrap only, no minimum image: SURVEY Q1)
//   nnmd/features/symm_funcs.py:85-104  calculate_distances
//   nnmd/features/symm_funcs.py:18-31   _pairwise_filter    0 < d < rc
// by an O(N) cell list.  Because the reference never adds periodic images, the wrapped atoms of a
// frame form a finite cluster: the grid is laid over the bounding box of the wrapped positions and
// no cell wraps around.  Output is a CSR list with ascending rows, so the pair set is reproducible
// bit for bit and every later reduction has a fixed summation order.
//
// Data layout in HBM
//   posw   (n_total,4) working dtype        wrapped positions, 16 B (fp32) / 32 B (fp64) per atom
//   ws     header | cell_start | cursor | cell_of | order | spos (cell-sorted copy of posw) | counts | scan scratch
//   offsets int64 (n_rows+1), idx int32 (total pairs, global atom indices)
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace nnmd {

// Every kernel of the build takes `go`: a device flag (or NULL = always run).  *go == 0 turns the whole build into
// no-ops, which is how a neighbour list is REUSED inside a fixed launch sequence (CUDA-graph replay of an MD step: the
// list is rebuilt only when some atom has moved more than half the skin, decided on the device by skin_check_kernel).
#define NNMD_GO_GUARD(go) do { if ((go) != nullptr && *(go) == 0) return; } while (0)

// ------------------------------------------------------------------ workspace layout (host+device)
struct GridHdr {
    long long bbox[6];  // order-preserving encoding of double: min xyz, max xyz
    double lo[3];
    double invw[3];
    int nc[3];
    int ncell;  // cells per frame actually used (<= cap)
};

struct WsLayout {
    int64_t cap;          // cell slots reserved per frame
    int64_t ncells_total; // n_frames * cap
    int64_t off_cell_start, off_cursor, off_cell_of, off_order, off_spos, off_counts, off_ecounts, off_scan, total;
};

static inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

static WsLayout ws_layout(int64_t n_frames, int64_t n_atoms, int64_t n_centres) {
    WsLayout L;
    int64_t cap = n_atoms;
    if (cap < 1) cap = 1;
    if (cap > (int64_t(1) << 22)) cap = int64_t(1) << 22;
    L.cap = cap;
    L.ncells_total = n_frames * cap;
    const int64_t n_total = n_frames * n_atoms;
    const int64_t n_rows = n_frames * n_centres;
    int64_t o = align256(sizeof(GridHdr));
    L.off_cell_start = o; o = align256(o + 4 * (L.ncells_total + 1));
    L.off_cursor = o;     o = align256(o + 4 * (L.ncells_total + 1));
    L.off_cell_of = o;    o = align256(o + 4 * n_total);
    L.off_order = o;      o = align256(o + 4 * n_total);
    L.off_spos = o;       o = align256(o + 32 * n_total);  // sized for double4
    L.off_counts = o;     o = align256(o + 4 * (n_rows + 1));
    L.off_ecounts = o;    o = align256(o + 4 * (n_rows + 1));
    const int64_t longest = (L.ncells_total > n_rows ? L.ncells_total : n_rows) + 1;
    L.off_scan = o;       o = align256(o + 8 * (longest / 2048 + 2));
    L.total = o;
    return L;
}

// ------------------------------------------------------------------ K0 wrap
template <typename T>
__global__ void wrap_kernel(const T *__restrict__ pos, int64_t n, T c00, T c01, T c02, T c10, T c11, T c12, T c20, T c21,
                            T c22, T i00, T i01, T i02, T i10, T i11, T i12, T i20, T i21, T i22, T p0, T p1, T p2,
                            int has_cell, T *__restrict__ posw) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    T x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    if (has_cell) {
        // frac = pos @ inv ; frac -= floor(frac) * pbc ; pos = frac @ cell   (symm_funcs.py:74-82)
        T f0 = x * i00 + y * i10 + z * i20;
        T f1 = x * i01 + y * i11 + z * i21;
        T f2 = x * i02 + y * i12 + z * i22;
        f0 -= M<T>::floor_(f0) * p0;
        f1 -= M<T>::floor_(f1) * p1;
        f2 -= M<T>::floor_(f2) * p2;
        x = f0 * c00 + f1 * c10 + f2 * c20;
        y = f0 * c01 + f1 * c11 + f2 * c21;
        z = f0 * c02 + f1 * c12 + f2 * c22;
    }
    T *o = posw + 4 * i;
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4 *>(o) = make_float4(x, y, z, 0.f);
    } else {
        *reinterpret_cast<double2 *>(o) = make_double2(x, y);
        *reinterpret_cast<double2 *>(o + 2) = make_double2(z, 0.0);
    }
}

// ------------------------------------------------------------------ bounding box
__device__ __forceinline__ long long enc_double(double x) {
    long long b = __double_as_longlong(x);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double dec_double(long long k) {
    return __longlong_as_double(k >= 0 ? k : (k ^ 0x7fffffffffffffffLL));
}

__global__ void hdr_init_kernel(GridHdr *h, const int32_t *go) {
    NNMD_GO_GUARD(go);
    if (threadIdx.x < 3) {
        h->bbox[threadIdx.x] = enc_double(1e300);
        h->bbox[3 + threadIdx.x] = enc_double(-1e300);
    }
}

template <typename T>
__global__ void bbox_kernel(const T *__restrict__ posw, int64_t n, GridHdr *h, const int32_t *go) {
    NNMD_GO_GUARD(go);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        T x, y, z;
        load_pos<T>(posw, i, x, y, z);
        lo[0] = fmin(lo[0], double(x)); hi[0] = fmax(hi[0], double(x));
        lo[1] = fmin(lo[1], double(y)); hi[1] = fmax(hi[1], double(y));
        lo[2] = fmin(lo[2], double(z)); hi[2] = fmax(hi[2], double(z));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&h->bbox[k], enc_double(lo[k]));
            atomicMax(&h->bbox[3 + k], enc_double(hi[k]));
        }
    }
}

__global__ void grid_setup_kernel(GridHdr *h, double rc, long long cap, const int32_t *go) {
    NNMD_GO_GUARD(go);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long nc[3];
    double ext[3];
    for (int k = 0; k < 3; ++k) {
        double lo = dec_double(h->bbox[k]), hi = dec_double(h->bbox[3 + k]);
        if (!(hi >= lo)) { lo = 0.0; hi = 0.0; }
        ext[k] = hi - lo;
        h->lo[k] = lo;
        // cell width >= rc*(1+1e-3): rounding in the binning can never separate a pair by two cells
        double m = floor(ext[k] / (rc * 1.001));
        if (!(m >= 1.0)) m = 1.0;
        if (m > 1048576.0) m = 1048576.0;
        nc[k] = (long long)m;
    }
    long long prod = nc[0] * nc[1] * nc[2];
    if (prod > cap) {
        const double s = cbrt(double(cap) / double(prod));
        for (int k = 0; k < 3; ++k) {
            nc[k] = (long long)floor(double(nc[k]) * s);
            if (nc[k] < 1) nc[k] = 1;
        }
        prod = nc[0] * nc[1] * nc[2];
        while (prod > cap) {
            int big = 0;
            if (nc[1] > nc[big]) big = 1;
            if (nc[2] > nc[big]) big = 2;
            nc[big] -= 1;
            prod = nc[0] * nc[1] * nc[2];
        }
    }
    for (int k = 0; k < 3; ++k) {
        h->nc[k] = int(nc[k]);
        h->invw[k] = ext[k] > 0.0 ? double(nc[k]) / ext[k] : 0.0;
    }
    h->ncell = int(prod);
}

template <typename T>
__device__ __forceinline__ void cell_coords(const GridHdr *h, T x, T y, T z, int &cx, int &cy, int &cz) {
    cx = int((double(x) - h->lo[0]) * h->invw[0]);
    cy = int((double(y) - h->lo[1]) * h->invw[1]);
    cz = int((double(z) - h->lo[2]) * h->invw[2]);
    cx = min(max(cx, 0), h->nc[0] - 1);
    cy = min(max(cy, 0), h->nc[1] - 1);
    cz = min(max(cz, 0), h->nc[2] - 1);
}

template <typename T>
__global__ void bin_count_kernel(const T *__restrict__ posw, int64_t n_total, int64_t n_atoms, int64_t cap,
                                 const GridHdr *__restrict__ h, int *__restrict__ cell_of, int *__restrict__ cell_cnt,
                                 const int32_t *go) {
    NNMD_GO_GUARD(go);
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n_total) return;
    T x, y, z;
    load_pos<T>(posw, i, x, y, z);
    int cx, cy, cz;
    cell_coords<T>(h, x, y, z, cx, cy, cz);
    const int c = (cx * h->nc[1] + cy) * h->nc[2] + cz;
    const int64_t frame = i / n_atoms;
    cell_of[i] = c;
    atomicAdd(&cell_cnt[frame * cap + c], 1);
}

template <typename T>
__global__ void bin_fill_kernel(const T *__restrict__ posw, int64_t n_total, int64_t n_atoms, int64_t cap,
                                const int *__restrict__ cell_of, const int *__restrict__ cell_start,
                                int *__restrict__ cursor, int *__restrict__ order, T *__restrict__ spos, const int32_t *go) {
    NNMD_GO_GUARD(go);
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n_total) return;
    const int64_t frame = i / n_atoms;
    const int64_t c = frame * cap + cell_of[i];
    const int slot = cell_start[c] + atomicAdd(&cursor[c], 1);
    order[slot] = int(i);
    T x, y, z;
    load_pos<T>(posw, i, x, y, z);
    T *o = spos + 4 * int64_t(slot);
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4 *>(o) = make_float4(x, y, z, 0.f);
    } else {
        *reinterpret_cast<double2 *>(o) = make_double2(x, y);
        *reinterpret_cast<double2 *>(o + 2) = make_double2(z, 0.0);
    }
}

// ------------------------------------------------------------------ exclusive scan (int32 -> OutT), 2048 items / block
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long *total, long long *sh) {
    // sh: 32 entries.  returns the exclusive prefix of v over the block; *total = block sum.
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < (blockDim.x >> 5) ? sh[lane] : 0;
        long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        sh[lane] = wi - w;  // exclusive prefix of warp sums
        if (lane == 31) sh[32] = wi;
    }
    __syncthreads();
    const long long res = sh[warp] + incl - v;
    *total = sh[32];
    __syncthreads();
    return res;
}

__global__ void scan_block_sums_kernel(const int *__restrict__ in, int64_t n, long long *__restrict__ block_sums, const int32_t *go) {
    NNMD_GO_GUARD(go);
    __shared__ long long sh[33];
    const int64_t base = int64_t(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    long long v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) v += in[base + k];
    long long total;
    block_exclusive_scan(v, &total, sh);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place; writes the grand total to block_sums[nb]
__global__ void scan_top_kernel(long long *__restrict__ block_sums, int64_t nb, const int32_t *go) {
    NNMD_GO_GUARD(go);
    __shared__ long long sh[33];
    long long carry = 0;
    for (int64_t s = 0; s < nb; s += blockDim.x) {
        const int64_t i = s + threadIdx.x;
        const long long v = i < nb ? block_sums[i] : 0;
        long long total;
        const long long ex = block_exclusive_scan(v, &total, sh);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) block_sums[nb] = carry;
}

template <typename OutT>
__global__ void scan_apply_kernel(const int *__restrict__ in, int64_t n, const long long *__restrict__ block_sums,
                                  OutT *__restrict__ out, int64_t nb, const int32_t *go) {
    NNMD_GO_GUARD(go);
    __shared__ long long sh[33];
    const int64_t base = int64_t(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int item[SCAN_ITEMS];
    long long v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        item[k] = base + k < n ? in[base + k] : 0;
        v += item[k];
    }
    long long total;
    long long run = block_exclusive_scan(v, &total, sh) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = OutT(run);
        run += item[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = OutT(block_sums[nb]);
}

template <typename OutT>
static int exclusive_scan(const int *in, int64_t n, OutT *out, long long *scratch, cudaStream_t st, const int32_t *go = nullptr) {
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 0) {
        NNMD_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(OutT), st));
        return NNMD_OK;
    }
    scan_block_sums_kernel<<<unsigned(nb), SCAN_THREADS, 0, st>>>(in, n, scratch, go);
    NNMD_LAUNCH_CHECK("scan_block_sums_kernel");
    scan_top_kernel<<<1, SCAN_THREADS, 0, st>>>(scratch, nb, go);
    NNMD_LAUNCH_CHECK("scan_top_kernel");
    scan_apply_kernel<OutT><<<unsigned(nb), SCAN_THREADS, 0, st>>>(in, n, scratch, out, nb, go);
    NNMD_LAUNCH_CHECK("scan_apply_kernel");
    return NNMD_OK;
}

// ------------------------------------------------------------------ K1 count / fill  (one warp per centre atom)
// The candidate test is ONE device function so that count and fill can never disagree.
template <typename T>
__device__ __forceinline__ bool is_pair(T xi, T yi, T zi, T xj, T yj, T zj, T rc) {
    const T dx = xj - xi, dy = yj - yi, dz = zj - zi;
    // plain products and sums (no contraction): the same expression in both passes
    T d2;
    if constexpr (sizeof(T) == 4) d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    else d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const T d = M<T>::sqrt_(d2);
    return (d < rc) && (d > T(0));  // symm_funcs.py:29
}

// Rank sort of a staged row (unique keys, so the ranks are a permutation): rows of up to 128 entries keep their up to
// four keys per lane in registers and read every key of the row ONCE; longer rows take one pass per 32 keys.
__device__ __forceinline__ void rank_sort_row(const int *__restrict__ buf, int n, int32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    if (n <= 128) {
        int v[4], r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = lane + 32 * i < n ? buf[lane + 32 * i] : 0x7fffffff; r[i] = 0; }
        for (int t = 0; t < n; ++t) {
            const int b = buf[t];
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] += (b < v[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (lane + 32 * i < n) out[r[i]] = v[i];
    } else {
        for (int e = lane; e < n; e += 32) {
            const int v = buf[e];
            int r = 0;
            for (int t = 0; t < n; ++t) r += (buf[t] < v);
            out[r] = v;
        }
    }
}

// Bitonic sort of a staged row of up to 32 * NR unique keys held NR per lane (key i lives in lane i & 31, register i >> 5):
// exchanges across lanes are shuffles, exchanges across registers stay in the lane.  ~430 warp instructions for 128 keys
// against ~700 for the rank sort above (measured on the 1M-atom cell: profiles/README.md, round 2).
template <int NR>
__device__ __forceinline__ void bitonic_sort_row(const int *__restrict__ buf, int n, int32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    int v[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = lane + 32 * r < n ? buf[lane + 32 * r] : 0x7fffffff;
#pragma unroll
    for (int k = 2; k <= 32 * NR; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {  // partner register r ^ (j / 32), same lane
                const int dr = j >> 5;
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    if ((r & dr) == 0) {
                        const bool asc = (((lane + 32 * r) & k) == 0);
                        const int lo = min(v[r], v[r | dr]), hi = max(v[r], v[r | dr]);
                        v[r] = asc ? lo : hi;
                        v[r | dr] = asc ? hi : lo;
                    }
                }
            } else {        // partner lane ^ j, same register
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const int p = __shfl_xor_sync(0xffffffffu, v[r], j);
                    const bool asc = (((lane + 32 * r) & k) == 0);
                    const bool keep_min = asc == ((lane & j) == 0);
                    v[r] = keep_min ? min(v[r], p) : max(v[r], p);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r)
        if (lane + 32 * r < n) out[lane + 32 * r] = v[r];
}

// sorted copy of a staged row: bitonic network for rows that fit 128 keys, rank sort beyond
__device__ __forceinline__ void sort_row(const int *__restrict__ buf, int n, int32_t *__restrict__ out) {
    if (n <= 32) bitonic_sort_row<1>(buf, n, out);
    else if (n <= 64) bitonic_sort_row<2>(buf, n, out);
    else if (n <= 128) bitonic_sort_row<4>(buf, n, out);
    else rank_sort_row(buf, n, out);
}

template <typename T, bool FILL>
__global__ void __launch_bounds__(256) nlist_rows_kernel(const T *__restrict__ posw, const T *__restrict__ spos,
                                                         int64_t n_rows, int64_t n_atoms, int64_t n_centres,
                                                         int64_t cap, const GridHdr *__restrict__ h,
                                                         const int *__restrict__ cell_of,
                                                         const int *__restrict__ cell_start,
                                                         const int *__restrict__ order, T rc, int *__restrict__ counts,
                                                         int *__restrict__ ecounts,
                                                         const int64_t *__restrict__ offsets, int max_row,
                                                         int32_t *__restrict__ idx,
                                                         unsigned long long *__restrict__ stats, int64_t idx_cap,
                                                         const int32_t *go) {
    NNMD_GO_GUARD(go);
    extern __shared__ int sh_rows[];  // FILL: warps * max_row candidate indices
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int ncx = h->nc[0], ncy = h->nc[1], ncz = h->nc[2];
    int *buf = FILL ? sh_rows + size_t(warp) * max_row : nullptr;

    for (int64_t row = int64_t(blockIdx.x) * wpb + warp; row < n_rows; row += int64_t(gridDim.x) * wpb) {
        const int64_t frame = row / n_centres;
        const int64_t gi = frame * n_atoms + (row - frame * n_centres);
        T xi, yi, zi;
        load_pos<T>(posw, gi, xi, yi, zi);
        const int c = cell_of[gi];
        const int cz = c % ncz, cy = (c / ncz) % ncy, cx = c / (ncz * ncy);
        const int z0 = max(cz - 1, 0), z1 = min(cz + 1, ncz - 1);
        const int64_t fbase = frame * cap;
        int n_found = 0, n_gt = 0;  // n_gt: neighbours with a larger global index ("owned edges" of this centre)
        for (int ix = max(cx - 1, 0); ix <= min(cx + 1, ncx - 1); ++ix) {
            for (int iy = max(cy - 1, 0); iy <= min(cy + 1, ncy - 1); ++iy) {
                // the z-run of cells is contiguous in the cell-sorted arrays
                const int64_t cb = fbase + (int64_t(ix) * ncy + iy) * ncz;
                const int s0 = cell_start[cb + z0], s1 = cell_start[cb + z1 + 1];
                for (int sb = s0; sb < s1; sb += 32) {  // warp-uniform trip count (ballot inside)
                    const int s = sb + lane;
                    bool ok = false;
                    if (s < s1) {
                        T xj, yj, zj;
                        load_pos<T>(spos, s, xj, yj, zj);
                        ok = is_pair<T>(xi, yi, zi, xj, yj, zj, rc);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (FILL) {
                        if (ok) {
                            const int p = n_found + __popc(m & ((1u << lane) - 1u));
                            if (p < max_row) buf[p] = order[s];
                        }
                    } else {
                        n_gt += __popc(__ballot_sync(0xffffffffu, ok && int64_t(order[s]) > gi));
                    }
                    n_found += __popc(m);
                }
            }
        }
        if (!FILL) {
            if (lane == 0) {
                counts[row] = n_found;
                ecounts[row] = n_gt;
                atomicMax(&stats[1], (unsigned long long)n_found);
            }
        } else {
            __syncwarp();
            const int n = min(n_found, max_row);
            const int64_t o = offsets[row];
            // rank sort: indices inside a row are unique (no periodic images), so ranks are a permutation
            if (o + n <= idx_cap) sort_row(buf, n, idx + o);  // beyond the capacity: the overflow flag is raised elsewhere
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ K1, cell-tiled form
// One CTA per occupied cell: the candidates of its 9 z-runs (27 cells) are staged ONCE in shared memory and shared by
// every atom of the cell, instead of every atom streaming ~600 candidate positions through L2 on its own
// (that form, nlist_rows_kernel above, is L2-bound and remains the path for cells whose candidate set does not fit).
#ifndef NNMD_CELL_MCAND
#define NNMD_CELL_MCAND 768   // staged candidates per cell (20 B each): 1536 -> 768 lets twice the blocks share an SM (K1 2.22 -> 1.89 ms per 1M atoms);
                               // cells with more candidates take the unstaged loop
#endif
#ifndef NNMD_CELL_WARPS
#define NNMD_CELL_WARPS 4
#endif
constexpr int CELL_MCAND = NNMD_CELL_MCAND;
constexpr int CELL_WARPS = NNMD_CELL_WARPS;

template <typename T> __device__ __forceinline__ void load_pos_smem(const T *p, int i, T &x, T &y, T &z);
template <> __device__ __forceinline__ void load_pos_smem<float>(const float *p, int i, float &x, float &y, float &z) {
    const float4 v = *(reinterpret_cast<const float4 *>(p) + i);
    x = v.x; y = v.y; z = v.z;
}
template <> __device__ __forceinline__ void load_pos_smem<double>(const double *p, int i, double &x, double &y, double &z) {
    const double2 a = *(reinterpret_cast<const double2 *>(p) + 2 * i), b = *(reinterpret_cast<const double2 *>(p) + 2 * i + 1);
    x = a.x; y = a.y; z = b.x;
}

// PASS 0: count (counts, ecounts, stats).  PASS 1: fill the CSR rows at offsets[row] (sorted).  PASS 2: both at once for
// the capacity-bound build: the sorted row goes to a padded scratch row (idx + row * max_row), the count pass disappears.
template <typename T, int PASS>
__global__ void __launch_bounds__(CELL_WARPS * 32) nlist_cell_kernel(const T *__restrict__ spos, int64_t n_frames, int64_t n_atoms,
                                                                     int64_t n_centres, int64_t cap, const GridHdr *__restrict__ h,
                                                                     const int *__restrict__ cell_start, const int *__restrict__ order,
                                                                     T rc, int *__restrict__ counts, int *__restrict__ ecounts,
                                                                     const int64_t *__restrict__ offsets, int max_row,
                                                                     int32_t *__restrict__ idx, unsigned long long *__restrict__ stats,
                                                                     int64_t idx_cap, const int32_t *go) {
    NNMD_GO_GUARD(go);
    constexpr bool FILL = PASS != 0, COUNT = PASS != 1;
    extern __shared__ __align__(16) unsigned char cell_smem[];
    T *cpos = reinterpret_cast<T *>(cell_smem);                                  // CELL_MCAND x 4
    int *cidx = reinterpret_cast<int *>(cpos + 4 * CELL_MCAND);                  // CELL_MCAND global atom indices
    int *rowbuf = cidx + CELL_MCAND;                                             // FILL: CELL_WARPS x max_row
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int ncx = h->nc[0], ncy = h->nc[1], ncz = h->nc[2];
    const int64_t ncell = h->ncell;
    int *buf = FILL ? rowbuf + size_t(warp) * max_row : nullptr;
    for (int64_t cid = blockIdx.x; cid < n_frames * ncell; cid += gridDim.x) {
        const int64_t frame = cid / ncell;
        const int c = int(cid - frame * ncell);
        const int64_t fbase = frame * cap;
        const int a0 = cell_start[fbase + c], a1 = cell_start[fbase + c + 1];
        if (a0 == a1) continue;  // empty cell (block-uniform)
        const int cz = c % ncz, cy = (c / ncz) % ncy, cx = c / (ncz * ncy);
        const int z0 = max(cz - 1, 0), z1 = min(cz + 1, ncz - 1);
        // the 9 runs and their total
        int rs0[9], rlen[9], nrun = 0, M = 0;
        for (int ix = max(cx - 1, 0); ix <= min(cx + 1, ncx - 1); ++ix)
            for (int iy = max(cy - 1, 0); iy <= min(cy + 1, ncy - 1); ++iy) {
                const int64_t cb = fbase + (int64_t(ix) * ncy + iy) * ncz;
                const int s0 = cell_start[cb + z0], s1 = cell_start[cb + z1 + 1];
                rs0[nrun] = s0; rlen[nrun] = s1 - s0; M += s1 - s0; ++nrun;
            }
        const bool staged = M <= CELL_MCAND;
        if (staged) {
            int off = 0;
            for (int r = 0; r < nrun; ++r) {
                for (int k = threadIdx.x; k < rlen[r]; k += blockDim.x) {
                    T x, y, z;
                    load_pos<T>(spos, rs0[r] + k, x, y, z);
                    T *o = cpos + 4 * (off + k);
                    o[0] = x; o[1] = y; o[2] = z; o[3] = T(0);
                    cidx[off + k] = order[rs0[r] + k];
                }
                off += rlen[r];
            }
        }
        __syncthreads();
        for (int s = a0 + warp; s < a1; s += CELL_WARPS) {
            const int64_t gi = order[s];
            const int64_t local = gi - frame * n_atoms;
            if (local >= n_centres) continue;  // halo atoms are neighbours, never centres (warp-uniform)
            const int64_t row = frame * n_centres + local;
            T xi, yi, zi;
            load_pos<T>(spos, s, xi, yi, zi);
            int n_found = 0, n_gt = 0;
            if (staged) {
                for (int jb = 0; jb < M; jb += 32) {
                    const int j = jb + lane;
                    bool ok = false;
                    int gj = 0;
                    if (j < M) {
                        T xj, yj, zj;
                        load_pos_smem<T>(cpos, j, xj, yj, zj);
                        ok = is_pair<T>(xi, yi, zi, xj, yj, zj, rc);
                        gj = cidx[j];
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (FILL) {
                        if (ok) {
                            const int p = n_found + __popc(m & ((1u << lane) - 1u));
                            if (p < max_row) buf[p] = gj;
                        }
                    }
                    if (COUNT) n_gt += __popc(__ballot_sync(0xffffffffu, ok && int64_t(gj) > gi));
                    n_found += __popc(m);
                }
            } else {
                for (int r = 0; r < nrun; ++r) {
                    const int s0 = rs0[r], s1 = rs0[r] + rlen[r];
                    for (int sb = s0; sb < s1; sb += 32) {
                        const int q = sb + lane;
                        bool ok = false;
                        int gj = 0;
                        if (q < s1) {
                            T xj, yj, zj;
                            load_pos<T>(spos, q, xj, yj, zj);
                            ok = is_pair<T>(xi, yi, zi, xj, yj, zj, rc);
                            gj = order[q];
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, ok);
                        if (FILL) {
                            if (ok) {
                                const int p = n_found + __popc(m & ((1u << lane) - 1u));
                                if (p < max_row) buf[p] = gj;
                            }
                        }
                        if (COUNT) n_gt += __popc(__ballot_sync(0xffffffffu, ok && int64_t(gj) > gi));
                        n_found += __popc(m);
                    }
                }
            }
            if (COUNT) {
                if (lane == 0) {
                    counts[row] = n_found;
                    ecounts[row] = n_gt;
                    atomicMax(&stats[1], (unsigned long long)n_found);
                }
            }
            if (FILL) {
                __syncwarp();
                const int n = min(n_found, max_row);
                const int64_t o = PASS == 2 ? row * int64_t(max_row) : offsets[row];
                if (o + n <= idx_cap) sort_row(buf, n, idx + o);  // unique indices: ascending row, fixed summation order downstream
                __syncwarp();
            }
        }
        __syncthreads();
    }
}

template <typename T, int PASS>
static int launch_cell_kernel(const T *spos, int64_t n_frames, int64_t n_atoms, int64_t n_centres, int64_t cap, const GridHdr *h,
                              const int *cell_start, const int *order, T rc, int *counts, int *ecounts, const int64_t *offsets,
                              int max_row, int32_t *idx, unsigned long long *stats, cudaStream_t st,
                              int64_t idx_cap = INT64_MAX, const int32_t *go = nullptr) {
    const size_t smem = size_t(CELL_MCAND) * (4 * sizeof(T) + sizeof(int)) + (PASS != 0 ? size_t(CELL_WARPS) * max_row * sizeof(int) : 0);
    if (smem > 220 * 1024) return set_err(NNMD_ERR_CAPACITY, "neighbour row of %d entries exceeds shared memory", max_row);
    auto kern = nlist_cell_kernel<T, PASS>;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 1;
    NNMD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CELL_WARPS * 32, smem));
    if (occ < 1) occ = 1;
#ifndef NNMD_CELL_WAVES
#define NNMD_CELL_WAVES 32  // grid = resident blocks x this (cells cost differently: 2: 2.04, 4: 1.89, 8: 1.83, 32: 1.79 ms per 1M atoms)
#endif
    int64_t blocks = std::min<int64_t>(n_frames * std::max<int64_t>(cap, 1), int64_t(sm_count()) * occ * NNMD_CELL_WAVES);
    if (blocks < 1) blocks = 1;
    kern<<<unsigned(blocks), CELL_WARPS * 32, smem, st>>>(spos, n_frames, n_atoms, n_centres, cap, h, cell_start, order, rc, counts,
                                                         ecounts, offsets, max_row, idx, stats, idx_cap, go);
    NNMD_LAUNCH_CHECK("nlist_cell_kernel");
    return NNMD_OK;
}

}  // namespace nnmd

using namespace nnmd;

extern "C" int64_t nnmd_nlist_workspace_bytes(int64_t n_frames, int64_t n_atoms) {
    if (n_frames < 0 || n_atoms < 0) return -1;
    return ws_layout(n_frames, n_atoms, n_atoms).total;
}

extern "C" int nnmd_wrap_positions(int dtype, const void *pos, int64_t n_total, const double *cell,
                                   const double *inv_cell, const double *pbc, int has_cell, void *posw, void *stream) {
    if (n_total < 0 || (n_total > 0 && (!pos || !posw))) return set_err(NNMD_ERR_INVALID, "wrap: null pointer");
    if (has_cell && (!cell || !inv_cell || !pbc)) return set_err(NNMD_ERR_INVALID, "wrap: cell data missing");
    if (n_total == 0) return NNMD_OK;
    double c[9] = {0}, iv[9] = {0}, p[3] = {0};
    if (has_cell) {
        memcpy(c, cell, sizeof(c));
        memcpy(iv, inv_cell, sizeof(iv));
        memcpy(p, pbc, sizeof(p));
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = unsigned((n_total + 255) / 256);
    ProfScope prof(PROF_WRAP, st);
    if (dtype == NNMD_F32) {
        wrap_kernel<float><<<blocks, 256, 0, st>>>((const float *)pos, n_total, (float)c[0], (float)c[1], (float)c[2],
                                                   (float)c[3], (float)c[4], (float)c[5], (float)c[6], (float)c[7],
                                                   (float)c[8], (float)iv[0], (float)iv[1], (float)iv[2], (float)iv[3],
                                                   (float)iv[4], (float)iv[5], (float)iv[6], (float)iv[7], (float)iv[8],
                                                   (float)p[0], (float)p[1], (float)p[2], has_cell, (float *)posw);
    } else if (dtype == NNMD_F64) {
        wrap_kernel<double><<<blocks, 256, 0, st>>>((const double *)pos, n_total, c[0], c[1], c[2], c[3], c[4], c[5],
                                                    c[6], c[7], c[8], iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6],
                                                    iv[7], iv[8], p[0], p[1], p[2], has_cell, (double *)posw);
    } else {
        return set_err(NNMD_ERR_INVALID, "wrap: bad dtype %d", dtype);
    }
    NNMD_LAUNCH_CHECK("wrap_kernel");
    return NNMD_OK;
}

static bool use_cell_kernel() {
    const char *env = tuning_env("NNMD_NLIST_ROWS");  // tuning knob: 1 = per-atom traversal (the first form of K1)
    return !(env && env[0] == '1');
}

// small guarded helpers of the capacity-bound build
__global__ void clear_i64_kernel(int64_t *p, int n, const int32_t *go) {
    NNMD_GO_GUARD(go);
    if (threadIdx.x < n) p[threadIdx.x] = 0;
}
__global__ void clear_i32_kernel(int *p, int64_t n, const int32_t *go) {
    NNMD_GO_GUARD(go);
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) p[i] = 0;
}
// stats = {total pairs, longest row, owned edges}: totals come from the scans' last elements
__global__ void stats_finish_kernel(int64_t *stats, const int64_t *offsets, const int64_t *edge_offsets, int64_t n_rows,
                                    int64_t idx_cap, int64_t max_row_cap, int64_t edge_cap, uint32_t *flag, int need_out,
                                    const int32_t *go) {
    NNMD_GO_GUARD(go);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    stats[0] = offsets[n_rows];
    if (edge_offsets) stats[2] = edge_offsets[n_rows];
    if (flag && (stats[0] > idx_cap || stats[1] > max_row_cap || (edge_offsets && stats[2] > edge_cap))) {
        // the demand of the FIRST build that did not fit is what the caller regrows to (stats[3..5]); the steps after it run
        // on garbage and must not overwrite it
        const unsigned old = atomicOr(flag, NNMD_FLAG_OVERFLOW);
        if (!(old & NNMD_FLAG_OVERFLOW) && need_out) { stats[3] = stats[0]; stats[4] = stats[1]; stats[5] = stats[2]; }
    }
}

template <typename T>
static int nlist_count_impl(const T *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres, double rc, char *ws,
                            int64_t *offsets, int64_t *edge_offsets, int64_t *stats, cudaStream_t st,
                            int64_t idx_cap = INT64_MAX, int64_t max_row_cap = INT64_MAX, int64_t edge_cap = INT64_MAX,
                            uint32_t *flag = nullptr, const int32_t *go = nullptr) {
    const WsLayout L = ws_layout(n_frames, n_atoms, n_centres);
    const int64_t n_total = n_frames * n_atoms, n_rows = n_frames * n_centres;
    GridHdr *h = (GridHdr *)ws;
    int *cell_start = (int *)(ws + L.off_cell_start);
    int *cursor = (int *)(ws + L.off_cursor);
    int *cell_of = (int *)(ws + L.off_cell_of);
    int *order = (int *)(ws + L.off_order);
    T *spos = (T *)(ws + L.off_spos);
    int *counts = (int *)(ws + L.off_counts);
    int *ecounts = (int *)(ws + L.off_ecounts);
    long long *scratch = (long long *)(ws + L.off_scan);

    clear_i64_kernel<<<1, 32, 0, st>>>(stats, 3, go);
    NNMD_LAUNCH_CHECK("clear_i64_kernel");
    if (n_rows == 0 || n_total == 0) {
        NNMD_CUDA_CHECK(cudaMemsetAsync(offsets, 0, (n_rows + 1) * sizeof(int64_t), st));
        if (edge_offsets) NNMD_CUDA_CHECK(cudaMemsetAsync(edge_offsets, 0, (n_rows + 1) * sizeof(int64_t), st));
        return NNMD_OK;
    }
    const int sms = sm_count();
    ProfScope prof(PROF_NLIST, st);
    hdr_init_kernel<<<1, 32, 0, st>>>(h, go);
    NNMD_LAUNCH_CHECK("hdr_init_kernel");
    {
        int64_t blocks = (n_total + 255) / 256;
        if (blocks > int64_t(sms) * 8) blocks = int64_t(sms) * 8;
        bbox_kernel<T><<<unsigned(blocks), 256, 0, st>>>(posw, n_total, h, go);
        NNMD_LAUNCH_CHECK("bbox_kernel");
    }
    grid_setup_kernel<<<1, 32, 0, st>>>(h, rc, (long long)L.cap, go);
    NNMD_LAUNCH_CHECK("grid_setup_kernel");
    // cursor doubles as the per-cell histogram, then is cleared to serve as the fill cursor
    const unsigned cblocks = unsigned(std::min<int64_t>((L.ncells_total + 1 + 255) / 256, int64_t(sms) * 8));
    clear_i32_kernel<<<cblocks, 256, 0, st>>>(cursor, L.ncells_total + 1, go);
    NNMD_LAUNCH_CHECK("clear_i32_kernel");
    const unsigned ablocks = unsigned((n_total + 255) / 256);
    bin_count_kernel<T><<<ablocks, 256, 0, st>>>(posw, n_total, n_atoms, L.cap, h, cell_of, cursor, go);
    NNMD_LAUNCH_CHECK("bin_count_kernel");
    int rcode = exclusive_scan<int>(cursor, L.ncells_total, cell_start, scratch, st, go);
    if (rcode) return rcode;
    clear_i32_kernel<<<cblocks, 256, 0, st>>>(cursor, L.ncells_total + 1, go);
    NNMD_LAUNCH_CHECK("clear_i32_kernel");
    bin_fill_kernel<T><<<ablocks, 256, 0, st>>>(posw, n_total, n_atoms, L.cap, cell_of, cell_start, cursor, order, spos, go);
    NNMD_LAUNCH_CHECK("bin_fill_kernel");
    if (use_cell_kernel()) {
        rcode = launch_cell_kernel<T, 0>(spos, n_frames, n_atoms, n_centres, L.cap, h, cell_start, order, T(rc), counts,
                                             ecounts, nullptr, 0, nullptr, (unsigned long long *)stats, st, INT64_MAX, go);
        if (rcode) return rcode;
    } else {
        const int wpb = 8;
        int64_t blocks = (n_rows + wpb - 1) / wpb;
        if (blocks > int64_t(sms) * 16) blocks = int64_t(sms) * 16;
        nlist_rows_kernel<T, false><<<unsigned(blocks), wpb * 32, 0, st>>>(
            posw, spos, n_rows, n_atoms, n_centres, L.cap, h, cell_of, cell_start, order, T(rc), counts, ecounts, nullptr, 0,
            nullptr, (unsigned long long *)stats, INT64_MAX, go);
        NNMD_LAUNCH_CHECK("nlist_rows_kernel<count>");
    }
    rcode = exclusive_scan<int64_t>(counts, n_rows, offsets, scratch, st, go);
    if (rcode) return rcode;
    if (edge_offsets) {
        rcode = exclusive_scan<int64_t>(ecounts, n_rows, edge_offsets, scratch, st, go);
        if (rcode) return rcode;
    }
    // stats[0] = total = offsets[n_rows], stats[2] = owned edges; raises NNMD_FLAG_OVERFLOW against the caller's capacities
    stats_finish_kernel<<<1, 32, 0, st>>>(stats, offsets, edge_offsets, n_rows, idx_cap, max_row_cap, edge_cap, flag, flag ? 1 : 0, go);
    NNMD_LAUNCH_CHECK("stats_finish_kernel");
    return NNMD_OK;
}

// padded scratch rows (row * stride) -> CSR rows at offsets[row]; one warp per row, coalesced
__global__ void compact_rows_kernel(const int32_t *__restrict__ rows, int64_t n_rows, int stride, const int64_t *__restrict__ offsets,
                                    int32_t *__restrict__ idx, int64_t idx_cap, const int32_t *go) {
    NNMD_GO_GUARD(go);
    const int lane = lane_id(), wpb = blockDim.x >> 5;
    for (int64_t row = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); row < n_rows; row += int64_t(gridDim.x) * wpb) {
        const int64_t o0 = offsets[row], o1 = offsets[row + 1];
        if (o1 > idx_cap || o1 - o0 > stride) continue;  // overflow: flagged by stats_finish_kernel
        const int32_t *src = rows + row * int64_t(stride);
        for (int64_t t = lane; t < o1 - o0; t += 32) idx[o0 + t] = src[t];
    }
}

// capacity-bound build in ONE pass over the candidates (cell-tiled kernel only): sorted rows into padded scratch, counts,
// scans, compaction.  Saves the separate count pass of the two-step build (1.3 of 4.0 ms on the 1M-atom cell).
template <typename T>
static int nlist_single_pass_impl(const T *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres, double rc, char *ws,
                                  int64_t *offsets, int64_t *edge_offsets, int32_t *idx, int64_t idx_cap, int64_t max_row_cap,
                                  int64_t edge_cap, int32_t *row_scratch, int64_t *stats, uint32_t *flag, const int32_t *go,
                                  cudaStream_t st) {
    const WsLayout L = ws_layout(n_frames, n_atoms, n_centres);
    const int64_t n_total = n_frames * n_atoms, n_rows = n_frames * n_centres;
    GridHdr *h = (GridHdr *)ws;
    int *cell_start = (int *)(ws + L.off_cell_start);
    int *cursor = (int *)(ws + L.off_cursor);
    int *cell_of = (int *)(ws + L.off_cell_of);
    int *order = (int *)(ws + L.off_order);
    T *spos = (T *)(ws + L.off_spos);
    int *counts = (int *)(ws + L.off_counts);
    int *ecounts = (int *)(ws + L.off_ecounts);
    long long *scratch = (long long *)(ws + L.off_scan);
    const int sms = sm_count();
    clear_i64_kernel<<<1, 32, 0, st>>>(stats, 3, go);
    NNMD_LAUNCH_CHECK("clear_i64_kernel");
    ProfScope prof(PROF_NLIST, st);
    hdr_init_kernel<<<1, 32, 0, st>>>(h, go);
    NNMD_LAUNCH_CHECK("hdr_init_kernel");
    {
        int64_t blocks = (n_total + 255) / 256;
        if (blocks > int64_t(sms) * 8) blocks = int64_t(sms) * 8;
        bbox_kernel<T><<<unsigned(blocks), 256, 0, st>>>(posw, n_total, h, go);
        NNMD_LAUNCH_CHECK("bbox_kernel");
    }
    grid_setup_kernel<<<1, 32, 0, st>>>(h, rc, (long long)L.cap, go);
    NNMD_LAUNCH_CHECK("grid_setup_kernel");
    const unsigned cblocks = unsigned(std::min<int64_t>((L.ncells_total + 1 + 255) / 256, int64_t(sms) * 8));
    clear_i32_kernel<<<cblocks, 256, 0, st>>>(cursor, L.ncells_total + 1, go);
    NNMD_LAUNCH_CHECK("clear_i32_kernel");
    const unsigned ablocks = unsigned((n_total + 255) / 256);
    bin_count_kernel<T><<<ablocks, 256, 0, st>>>(posw, n_total, n_atoms, L.cap, h, cell_of, cursor, go);
    NNMD_LAUNCH_CHECK("bin_count_kernel");
    int rcode = exclusive_scan<int>(cursor, L.ncells_total, cell_start, scratch, st, go);
    if (rcode) return rcode;
    clear_i32_kernel<<<cblocks, 256, 0, st>>>(cursor, L.ncells_total + 1, go);
    NNMD_LAUNCH_CHECK("clear_i32_kernel");
    bin_fill_kernel<T><<<ablocks, 256, 0, st>>>(posw, n_total, n_atoms, L.cap, cell_of, cell_start, cursor, order, spos, go);
    NNMD_LAUNCH_CHECK("bin_fill_kernel");
    rcode = launch_cell_kernel<T, 2>(spos, n_frames, n_atoms, n_centres, L.cap, h, cell_start, order, T(rc), counts, ecounts, nullptr,
                                     int(max_row_cap), row_scratch, (unsigned long long *)stats, st, n_rows * max_row_cap, go);
    if (rcode) return rcode;
    rcode = exclusive_scan<int64_t>(counts, n_rows, offsets, scratch, st, go);
    if (rcode) return rcode;
    if (edge_offsets) {
        rcode = exclusive_scan<int64_t>(ecounts, n_rows, edge_offsets, scratch, st, go);
        if (rcode) return rcode;
    }
    stats_finish_kernel<<<1, 32, 0, st>>>(stats, offsets, edge_offsets, n_rows, idx_cap, max_row_cap, edge_cap, flag, flag ? 1 : 0, go);
    NNMD_LAUNCH_CHECK("stats_finish_kernel");
    {
        const int wpb = 8;
        const int64_t blocks = std::min<int64_t>((n_rows + wpb - 1) / wpb, int64_t(sms) * 16);
        compact_rows_kernel<<<unsigned(blocks), wpb * 32, 0, st>>>(row_scratch, n_rows, int(max_row_cap), offsets, idx, idx_cap, go);
        NNMD_LAUNCH_CHECK("compact_rows_kernel");
    }
    return NNMD_OK;
}

template <typename T>
static int nlist_fill_impl(const T *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres, double rc,
                           const char *ws, const int64_t *offsets, int64_t max_row, int32_t *idx, cudaStream_t st,
                           int64_t idx_cap = INT64_MAX, const int32_t *go = nullptr) {
    const WsLayout L = ws_layout(n_frames, n_atoms, n_centres);
    const int64_t n_rows = n_frames * n_centres;
    if (n_rows == 0 || max_row == 0) return NNMD_OK;
    const GridHdr *h = (const GridHdr *)ws;
    const int *cell_start = (const int *)(ws + L.off_cell_start);
    const int *cell_of = (const int *)(ws + L.off_cell_of);
    const int *order = (const int *)(ws + L.off_order);
    const T *spos = (const T *)(ws + L.off_spos);
    if (use_cell_kernel()) {
        ProfScope prof(PROF_NLIST, st);
        return launch_cell_kernel<T, 1>(spos, n_frames, n_atoms, n_centres, L.cap, h, cell_start, order, T(rc), nullptr, nullptr,
                                           offsets, int(max_row), idx, nullptr, st, idx_cap, go);
    }
    int wpb = 8;
    while (wpb > 1 && size_t(wpb) * max_row * 4 > 160 * 1024) wpb >>= 1;
    const size_t smem = size_t(wpb) * max_row * 4;
    if (smem > 220 * 1024) return set_err(NNMD_ERR_CAPACITY, "neighbour row of %lld entries exceeds shared memory", (long long)max_row);
    auto kern = nlist_rows_kernel<T, true>;
    NNMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const int sms = sm_count();
    int64_t blocks = (n_rows + wpb - 1) / wpb;
    if (blocks > int64_t(sms) * 16) blocks = int64_t(sms) * 16;
    ProfScope prof(PROF_NLIST, st);
    kern<<<unsigned(blocks), wpb * 32, smem, st>>>(posw, spos, n_rows, n_atoms, n_centres, L.cap, h, cell_of, cell_start,
                                                   order, T(rc), nullptr, nullptr, offsets, int(max_row), idx, nullptr, idx_cap, go);
    NNMD_LAUNCH_CHECK("nlist_rows_kernel<fill>");
    return NNMD_OK;
}

// ------------------------------------------------------------------ Verlet-skin reuse decision (device side)
// acc: one uint32 holding the bit pattern of the largest squared displacement (non-negative floats order like uints)
template <typename T>
__global__ void skin_max_kernel(const T *__restrict__ posw, const T *__restrict__ ref, int64_t n, unsigned *__restrict__ acc) {
    float worst = 0.f;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        T x, y, z, rx, ry, rz;
        load_pos<T>(posw, i, x, y, z);
        load_pos<T>(ref, i, rx, ry, rz);
        const double dx = double(x) - double(rx), dy = double(y) - double(ry), dz = double(z) - double(rz);
        const float d2 = float(dx * dx + dy * dy + dz * dz);
        worst = (d2 == d2) ? fmaxf(worst, d2) : 3.0e38f;  // NaN (uninitialised reference) forces a rebuild
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if (lane_id() == 0) atomicMax(acc, __float_as_uint(worst));
}
__global__ void skin_decide_kernel(unsigned *acc, float thr2, int32_t *go, int64_t *n_builds) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int g = __uint_as_float(*acc) > thr2 ? 1 : 0;
    *go = g;
    *acc = 0u;
    if (n_builds && g) *n_builds += 1;
}
template <typename T>
__global__ void skin_copy_kernel(const T *__restrict__ posw, T *__restrict__ ref, int64_t n4, const int32_t *go) {
    NNMD_GO_GUARD(go);
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) ref[i] = posw[i];
}

static int nlist_args_ok(int dtype, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres, double rc,
                         const void *ws) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "nlist: bad dtype %d", dtype);
    if (n_frames < 0 || n_atoms < 0 || n_centres < 0 || n_centres > n_atoms)
        return set_err(NNMD_ERR_INVALID, "nlist: bad sizes frames=%lld atoms=%lld centres=%lld", (long long)n_frames,
                       (long long)n_atoms, (long long)n_centres);
    if (n_frames * n_atoms >= (int64_t(1) << 31)) return set_err(NNMD_ERR_INVALID, "nlist: more than 2^31 atoms");
    if (!(rc > 0.0)) return set_err(NNMD_ERR_INVALID, "nlist: cutoff must be positive");
    if (n_frames * n_atoms > 0 && (!posw || !ws)) return set_err(NNMD_ERR_INVALID, "nlist: null pointer");
    return NNMD_OK;
}

extern "C" int nnmd_nlist_count(int dtype, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres,
                                double rc, void *ws, int64_t *offsets, int64_t *edge_offsets, int64_t *stats, void *stream) {
    int r = nlist_args_ok(dtype, posw, n_frames, n_atoms, n_centres, rc, ws);
    if (r) return r;
    if (!offsets || !stats) return set_err(NNMD_ERR_INVALID, "nlist_count: null output");
    if (dtype == NNMD_F32)
        return nlist_count_impl<float>((const float *)posw, n_frames, n_atoms, n_centres, rc, (char *)ws, offsets, edge_offsets, stats, (cudaStream_t)stream);
    return nlist_count_impl<double>((const double *)posw, n_frames, n_atoms, n_centres, rc, (char *)ws, offsets, edge_offsets, stats, (cudaStream_t)stream);
}

extern "C" int nnmd_nlist_fill(int dtype, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres,
                               double rc, const void *ws, const int64_t *offsets, int64_t max_row, int32_t *idx,
                               void *stream) {
    int r = nlist_args_ok(dtype, posw, n_frames, n_atoms, n_centres, rc, ws);
    if (r) return r;
    if (max_row < 0) return set_err(NNMD_ERR_INVALID, "nlist_fill: negative max_row");
    if (max_row > 0 && (!offsets || !idx)) return set_err(NNMD_ERR_INVALID, "nlist_fill: null pointer");
    if (dtype == NNMD_F32)
        return nlist_fill_impl<float>((const float *)posw, n_frames, n_atoms, n_centres, rc, (const char *)ws, offsets, max_row, idx, (cudaStream_t)stream);
    return nlist_fill_impl<double>((const double *)posw, n_frames, n_atoms, n_centres, rc, (const char *)ws, offsets, max_row, idx, (cudaStream_t)stream);
}

extern "C" int nnmd_nlist_build(int dtype, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres, double rc,
                                void *ws, int64_t *offsets, int64_t *edge_offsets, int32_t *idx, int64_t idx_cap,
                                int64_t max_row_cap, int64_t edge_cap, void *row_scratch, int64_t *stats, uint32_t *flag,
                                const int32_t *go, void *stream) {
    int r = nlist_args_ok(dtype, posw, n_frames, n_atoms, n_centres, rc, ws);
    if (r) return r;
    if (!offsets || !stats || !flag || (idx_cap > 0 && !idx)) return set_err(NNMD_ERR_INVALID, "nlist_build: null pointer");
    if (idx_cap < 0 || max_row_cap < 1 || edge_cap < 0) return set_err(NNMD_ERR_INVALID, "nlist_build: bad capacities");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cell_smem = size_t(CELL_MCAND) * ((dtype == NNMD_F32 ? 16 : 32) + sizeof(int)) + size_t(CELL_WARPS) * max_row_cap * sizeof(int);
    if (row_scratch && use_cell_kernel() && n_frames * n_centres > 0 && n_frames * n_atoms > 0 && cell_smem <= 220 * 1024) {
        if (dtype == NNMD_F32)
            return nlist_single_pass_impl<float>((const float *)posw, n_frames, n_atoms, n_centres, rc, (char *)ws, offsets, edge_offsets, idx,
                                                 idx_cap, max_row_cap, edge_cap, (int32_t *)row_scratch, stats, flag, go, st);
        return nlist_single_pass_impl<double>((const double *)posw, n_frames, n_atoms, n_centres, rc, (char *)ws, offsets, edge_offsets, idx,
                                              idx_cap, max_row_cap, edge_cap, (int32_t *)row_scratch, stats, flag, go, st);
    }
    if (dtype == NNMD_F32) {
        r = nlist_count_impl<float>((const float *)posw, n_frames, n_atoms, n_centres, rc, (char *)ws, offsets, edge_offsets, stats, st,
                                    idx_cap, max_row_cap, edge_cap, flag, go);
        if (r) return r;
        return nlist_fill_impl<float>((const float *)posw, n_frames, n_atoms, n_centres, rc, (const char *)ws, offsets, max_row_cap, idx, st, idx_cap, go);
    }
    r = nlist_count_impl<double>((const double *)posw, n_frames, n_atoms, n_centres, rc, (char *)ws, offsets, edge_offsets, stats, st,
                                 idx_cap, max_row_cap, edge_cap, flag, go);
    if (r) return r;
    return nlist_fill_impl<double>((const double *)posw, n_frames, n_atoms, n_centres, rc, (const char *)ws, offsets, max_row_cap, idx, st, idx_cap, go);
}

extern "C" int64_t nnmd_nlist_row_scratch_bytes(int64_t n_rows, int64_t max_row_cap) {
    if (n_rows < 0 || max_row_cap < 0) return -1;
    return n_rows * max_row_cap * int64_t(sizeof(int32_t));
}

extern "C" int nnmd_skin_check(int dtype, const void *posw, void *pos_ref, int64_t n_total, double skin, uint32_t *acc,
                               int32_t *go, int64_t *n_builds, void *stream) {
    if (dtype != NNMD_F32 && dtype != NNMD_F64) return set_err(NNMD_ERR_INVALID, "skin_check: bad dtype %d", dtype);
    if (n_total < 0 || !(skin >= 0.0)) return set_err(NNMD_ERR_INVALID, "skin_check: bad arguments");
    if (!acc || !go || (n_total > 0 && (!posw || !pos_ref))) return set_err(NNMD_ERR_INVALID, "skin_check: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = unsigned(std::max<int64_t>(1, std::min<int64_t>((n_total + 255) / 256, int64_t(sm_count()) * 8)));
    const float thr2 = float(0.25 * skin * skin);  // an atom that moved more than skin / 2
    if (dtype == NNMD_F32) skin_max_kernel<float><<<blocks, 256, 0, st>>>((const float *)posw, (const float *)pos_ref, n_total, acc);
    else skin_max_kernel<double><<<blocks, 256, 0, st>>>((const double *)posw, (const double *)pos_ref, n_total, acc);
    NNMD_LAUNCH_CHECK("skin_max_kernel");
    skin_decide_kernel<<<1, 32, 0, st>>>(acc, thr2, go, n_builds);
    NNMD_LAUNCH_CHECK("skin_decide_kernel");
    if (dtype == NNMD_F32) skin_copy_kernel<float><<<blocks, 256, 0, st>>>((const float *)posw, (float *)pos_ref, n_total * 4, go);
    else skin_copy_kernel<double><<<blocks, 256, 0, st>>>((const double *)posw, (double *)pos_ref, n_total * 4, go);
    NNMD_LAUNCH_CHECK("skin_copy_kernel");
    return NNMD_OK;
}
