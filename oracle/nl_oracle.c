/*
 * nl_oracle.c -- O(N) neighbour-list CPU oracle (O3) for the nnmd descriptor path.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/build.py with gcc (+OpenMP) into
 * oracle/_build/libnl_oracle.so; loaded only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  The product
 * (nnmd_b200/) never links, imports or calls it.
 *
 * It restates in double precision, with a cell list and analytic gradients, what
 * the reference (ded-otshelnik/nnmd) computes with dense tensors + autograd:
 *
 *   pair set        0 < d_ij < rc                         nnmd/features/symm_funcs.py:18-31
 *   triplet set     d_ij, d_ik, d_jk all in (0, rc)       symm_funcs.py:34-50
 *   fc(d)           0.5*(cos(pi d/rc)+1)                  symm_funcs.py:107-120
 *   G1 / G2         sum_j [exp(-eta (d-rs)^2)] fc(d)      symm_funcs.py:123-162
 *   G4 / G5         2^(1-zeta) sum_{j,k ordered} P(c) exp(..) fc fc fc
 *                   P(c) = (1+lambda|c|)^zeta if |c|>1e-3 else 0,
 *                   c forced to 0 when 2 d_ij d_ik < 1e-3 symm_funcs.py:165-277
 *   dg[a,f,:]       d(sum_i G_i,f)/d r_a                  nnmd/features/calculate_sf.py:85-90
 *
 * Positions handed in are ALREADY wrapped (the Python wrapper applies the
 * reference's wrap arithmetic, symm_funcs.py:53-82).  The power is evaluated
 * exactly in double (oracle "O2/O3" semantics; the reference's complex64
 * round-off, SURVEY.md Q5, is not reproduced).
 *
 * Parity pinning: validated against oracle/dense_oracle.py and against outputs
 * of the reference itself (the .npz fixtures under tests/golden) in tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KIND_G1 1
#define KIND_G2 2
#define KIND_G4 4
#define KIND_G5 5

typedef struct {
    int n;
    int cap;
    int *idx;
    double *ux, *uy, *uz, *d;
} nbr_t;

static void nbr_reserve(nbr_t *nb, int cap) {
    if (cap <= nb->cap) return;
    nb->cap = cap;
    nb->idx = (int *)realloc(nb->idx, sizeof(int) * cap);
    nb->ux = (double *)realloc(nb->ux, sizeof(double) * cap);
    nb->uy = (double *)realloc(nb->uy, sizeof(double) * cap);
    nb->uz = (double *)realloc(nb->uz, sizeof(double) * cap);
    nb->d = (double *)realloc(nb->d, sizeof(double) * cap);
}

static int cmp_nbr(const void *a, const void *b) {
    return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b);
}

static inline double fcut(double d, double rc) { return 0.5 * (cos(M_PI * d / rc) + 1.0); }
static inline double dfcut(double d, double rc) { return -0.5 * M_PI / rc * sin(M_PI * d / rc); }

static inline double cos_law(double a, double b, double opp) {
    /* a, b adjacent sides, opp the opposite one; reference thresholds (symm_funcs.py:193-197) */
    double den = 2.0 * a * b;
    if (!(fabs(den) >= 1e-3)) return 0.0;
    if (den < 1e-8) den = 1e-8;
    return (a * a + b * b - opp * opp) / den;
}

/*
 * One frame.  pos: n*3 wrapped positions.  kinds[nf], prm[nf*4] = (rc, p1, p2, p3)
 * with the reference's positional binding: G1 (rc), G2 (rc, eta, rs), G4/G5 (rc, eta, zeta, lambda).
 * G: n*nf (un-normalised), dG: n*nf*3.
 * nbr_count (optional, n ints): number of neighbours inside rc_max.
 * pairs (optional): if non-NULL, receives (i,j) for every pair with 0<d<rc_max, row-sorted;
 *                   pair_cap entries of 2 ints; *n_pairs gets the total (even if > cap).
 */
static void frame_eval(const double *pos, int n, const int *kinds, const double *prm, int nf,
                       double *G, double *dG, int *nbr_count,
                       int *pairs, int64_t pair_cap, int64_t *n_pairs, int nthreads) {
    double rcmax = 0.0;
    for (int f = 0; f < nf; ++f)
        if (prm[4 * f] > rcmax) rcmax = prm[4 * f];

    /* ---- group the angular functions by (kind, rc, eta): they share exp/fc work ---- */
    int *grp_of = (int *)malloc(sizeof(int) * (nf > 0 ? nf : 1));
    int *grp_kind = (int *)malloc(sizeof(int) * (nf > 0 ? nf : 1));
    double *grp_rc = (double *)malloc(sizeof(double) * (nf > 0 ? nf : 1));
    double *grp_eta = (double *)malloc(sizeof(double) * (nf > 0 ? nf : 1));
    double *coef_f = (double *)malloc(sizeof(double) * (nf > 0 ? nf : 1));
    int ngrp = 0;
    for (int f = 0; f < nf; ++f) {
        grp_of[f] = -1;
        coef_f[f] = 0.0;
        if (kinds[f] != KIND_G4 && kinds[f] != KIND_G5) continue;
        coef_f[f] = 2.0 * pow(2.0, 1.0 - prm[4 * f + 2]); /* x2: ordered (j,k) and (k,j) */
        int g;
        for (g = 0; g < ngrp; ++g)
            if (grp_kind[g] == kinds[f] && grp_rc[g] == prm[4 * f] && grp_eta[g] == prm[4 * f + 1]) break;
        if (g == ngrp) { grp_kind[g] = kinds[f]; grp_rc[g] = prm[4 * f]; grp_eta[g] = prm[4 * f + 1]; ngrp++; }
        grp_of[f] = g;
    }

    /* ---- cell list on the bounding box ---- */
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            double v = pos[3 * i + k];
            if (v < lo[k]) lo[k] = v;
            if (v > hi[k]) hi[k] = v;
        }
    int nc[3];
    double inv_w[3];
    for (int k = 0; k < 3; ++k) {
        double ext = hi[k] - lo[k];
        int m = (int)floor(ext / rcmax);
        if (m < 1) m = 1;
        if (m > 1024) m = 1024;
        nc[k] = m;
        inv_w[k] = (ext > 0) ? m / ext : 0.0;
    }
    int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    int *cell_of = (int *)malloc(sizeof(int) * n);
    int *start = (int *)calloc(ncell + 1, sizeof(int));
    int *order = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; ++i) {
        int c[3];
        for (int k = 0; k < 3; ++k) {
            int v = (int)((pos[3 * i + k] - lo[k]) * inv_w[k]);
            if (v >= nc[k]) v = nc[k] - 1;
            if (v < 0) v = 0;
            c[k] = v;
        }
        cell_of[i] = (c[0] * nc[1] + c[1]) * nc[2] + c[2];
        start[cell_of[i] + 1]++;
    }
    for (int64_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    int *fill = (int *)malloc(sizeof(int) * ncell);
    memcpy(fill, start, sizeof(int) * ncell);
    for (int i = 0; i < n; ++i) order[fill[cell_of[i]]++] = i; /* ascending i inside a cell */
    free(fill);

    int64_t *row_off = NULL;
    if (pairs || n_pairs) row_off = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    int *cnt_tmp = (int *)calloc(n, sizeof(int));

    for (int pass = 0; pass < ((pairs || n_pairs) ? 2 : 1); ++pass) {
        /* pass 0: everything (+ counts); pass 1: write pairs at their offsets */
        if (pass == 1) {
            for (int i = 0; i < n; ++i) row_off[i + 1] = row_off[i] + cnt_tmp[i];
            if (n_pairs) *n_pairs = row_off[n];
            if (!pairs) break;
        }
#pragma omp parallel num_threads(nthreads)
        {
            nbr_t nb = {0, 0, NULL, NULL, NULL, NULL, NULL};
            double *fcj = NULL, *dfcj = NULL, *psi = NULL, *dpsi = NULL;
            int aux_cap = 0;
            double *accf = (double *)malloc(sizeof(double) * 4 * (nf > 0 ? nf : 1));
#pragma omp for schedule(dynamic, 16)
            for (int a = 0; a < n; ++a) {
                /* ---- gather neighbours of a inside rcmax ---- */
                nb.n = 0;
                int ca = cell_of[a];
                int cz = ca % nc[2], cy = (ca / nc[2]) % nc[1], cx = ca / (nc[2] * nc[1]);
                double ax = pos[3 * a], ay = pos[3 * a + 1], az = pos[3 * a + 2];
                for (int ix = cx - 1; ix <= cx + 1; ++ix) {
                    if (ix < 0 || ix >= nc[0]) continue;
                    for (int iy = cy - 1; iy <= cy + 1; ++iy) {
                        if (iy < 0 || iy >= nc[1]) continue;
                        for (int iz = cz - 1; iz <= cz + 1; ++iz) {
                            if (iz < 0 || iz >= nc[2]) continue;
                            int c = (ix * nc[1] + iy) * nc[2] + iz;
                            for (int s = start[c]; s < start[c + 1]; ++s) {
                                int j = order[s];
                                double ux = pos[3 * j] - ax, uy = pos[3 * j + 1] - ay, uz = pos[3 * j + 2] - az;
                                double d = sqrt(ux * ux + uy * uy + uz * uz);
                                if (d < rcmax && d > 0.0) {
                                    nbr_reserve(&nb, nb.n + 1 > nb.cap ? 2 * (nb.n + 8) : nb.cap);
                                    nb.idx[nb.n] = j;
                                    nb.n++;
                                }
                            }
                        }
                    }
                }
                qsort(nb.idx, nb.n, sizeof(int), cmp_nbr);
                for (int t = 0; t < nb.n; ++t) {
                    int j = nb.idx[t];
                    nb.ux[t] = pos[3 * j] - ax;
                    nb.uy[t] = pos[3 * j + 1] - ay;
                    nb.uz[t] = pos[3 * j + 2] - az;
                    nb.d[t] = sqrt(nb.ux[t] * nb.ux[t] + nb.uy[t] * nb.uy[t] + nb.uz[t] * nb.uz[t]);
                }
                if (pass == 1) {
                    int64_t o = row_off[a];
                    for (int t = 0; t < nb.n; ++t)
                        if (o + t < pair_cap) { pairs[2 * (o + t)] = a; pairs[2 * (o + t) + 1] = nb.idx[t]; }
                    continue;
                }
                cnt_tmp[a] = nb.n;
                if (nbr_count) nbr_count[a] = nb.n;
                if (!G) continue;
                double *Ga = G + (size_t)a * nf;
                double *dGa = dG + (size_t)a * nf * 3;
                for (int f = 0; f < nf * 4; ++f) accf[f] = 0.0;
                /* ---- radial functions ---- */
                for (int f = 0; f < nf; ++f) {
                    int kind = kinds[f];
                    if (kind != KIND_G1 && kind != KIND_G2) continue;
                    double rc = prm[4 * f];
                    double eta = (kind == KIND_G2) ? prm[4 * f + 1] : 0.0;
                    double rs = (kind == KIND_G2) ? prm[4 * f + 2] : 0.0;
                    double acc = 0, gx = 0, gy = 0, gz = 0;
                    for (int t = 0; t < nb.n; ++t) {
                        double d = nb.d[t];
                        if (!(d < rc)) continue;
                        double fc = fcut(d, rc), dfc = dfcut(d, rc);
                        double e = (kind == KIND_G2) ? exp(-eta * (d - rs) * (d - rs)) : 1.0;
                        double phi = e * fc;
                        double dphi = e * (dfc - 2.0 * eta * (d - rs) * fc);
                        acc += phi;
                        /* own-atom gradient plus the identical term inside every neighbour's G_j */
                        double s = -2.0 * dphi / d;
                        gx += s * nb.ux[t];
                        gy += s * nb.uy[t];
                        gz += s * nb.uz[t];
                    }
                    accf[4 * f] = acc; accf[4 * f + 1] = gx; accf[4 * f + 2] = gy; accf[4 * f + 3] = gz;
                }
                /* ---- angular functions, grouped by (kind, rc, eta) ---- */
                if (ngrp > 0) {
                    if (nb.n * ngrp > aux_cap) {
                        aux_cap = 2 * nb.n * ngrp;
                        fcj = (double *)realloc(fcj, sizeof(double) * aux_cap);
                        dfcj = (double *)realloc(dfcj, sizeof(double) * aux_cap);
                        psi = (double *)realloc(psi, sizeof(double) * aux_cap);
                        dpsi = (double *)realloc(dpsi, sizeof(double) * aux_cap);
                    }
                    for (int g = 0; g < ngrp; ++g) {
                        double rc = grp_rc[g], eta = grp_eta[g];
                        for (int t = 0; t < nb.n; ++t) {
                            double d = nb.d[t];
                            int o = g * nb.n + t;
                            fcj[o] = fcut(d, rc);
                            dfcj[o] = dfcut(d, rc);
                            double e = exp(-eta * d * d);
                            psi[o] = e * fcj[o];
                            dpsi[o] = e * (dfcj[o] - 2.0 * eta * d * fcj[o]);
                        }
                    }
                    for (int tj = 0; tj < nb.n; ++tj) {
                        double x = nb.d[tj];
                        for (int tk = tj + 1; tk < nb.n; ++tk) {
                            double y = nb.d[tk];
                            double wx = nb.ux[tj] - nb.ux[tk], wy = nb.uy[tj] - nb.uy[tk], wz = nb.uz[tj] - nb.uz[tk];
                            double z = sqrt(wx * wx + wy * wy + wz * wz);
                            if (!(z < rcmax && z > 0.0)) continue;
                            /* geometry shared by every function */
                            double cc[3], dcx[3] = {0, 0, 0}, dcy[3] = {0, 0, 0}, lbp[3], lbm[3];
                            cc[0] = cos_law(x, y, z); /* centre a */
                            cc[1] = cos_law(x, z, y); /* centre j: adjacent x (j-a), z (j-k) */
                            cc[2] = cos_law(y, z, x); /* centre k: adjacent y (k-a), z (k-j) */
                            if (fabs(2 * x * y) >= 1e-3) { dcx[0] = 1.0 / y - cc[0] / x; dcy[0] = 1.0 / x - cc[0] / y; }
                            if (fabs(2 * x * z) >= 1e-3) { dcx[1] = 1.0 / z - cc[1] / x; dcy[1] = -y / (x * z); }
                            if (fabs(2 * y * z) >= 1e-3) { dcy[2] = 1.0 / z - cc[2] / y; dcx[2] = -x / (y * z); }
                            for (int X = 0; X < 3; ++X) {
                                double ac = fabs(cc[X]);
                                lbp[X] = log(1.0 + ac);
                                lbm[X] = (1.0 - ac > 0.0) ? log(1.0 - ac) : -INFINITY;
                            }
                            for (int g = 0; g < ngrp; ++g) {
                                double rc = grp_rc[g], eta = grp_eta[g];
                                if (!(x < rc && y < rc && z < rc)) continue;
                                int oj = g * nb.n + tj, ok = g * nb.n + tk;
                                double fz = fcut(z, rc);
                                double psz = exp(-eta * z * z) * fz;
                                double E[3], Ex[3], Ey[3];
                                if (grp_kind[g] == KIND_G4) {
                                    E[0] = psi[oj] * psi[ok] * psz; Ex[0] = dpsi[oj] * psi[ok] * psz; Ey[0] = psi[oj] * dpsi[ok] * psz;
                                    E[1] = E[2] = E[0]; Ex[1] = Ex[2] = Ex[0]; Ey[1] = Ey[2] = Ey[0];
                                } else {
                                    E[0] = psi[oj] * psi[ok] * fz;  Ex[0] = dpsi[oj] * psi[ok] * fz;  Ey[0] = psi[oj] * dpsi[ok] * fz;
                                    E[1] = psi[oj] * psz * fcj[ok]; Ex[1] = dpsi[oj] * psz * fcj[ok]; Ey[1] = psi[oj] * psz * dfcj[ok];
                                    E[2] = psi[ok] * psz * fcj[oj]; Ex[2] = psi[ok] * psz * dfcj[oj]; Ey[2] = dpsi[ok] * psz * fcj[oj];
                                }
                                for (int f = 0; f < nf; ++f) {
                                    if (grp_of[f] != g) continue;
                                    double zeta = prm[4 * f + 2], lam = prm[4 * f + 3];
                                    double coef = coef_f[f];
                                    double sa = 0, sb = 0;
                                    for (int X = 0; X < 3; ++X) {
                                        double c = cc[X], ac = fabs(c);
                                        if (!(ac > 1e-3)) continue;
                                        double base = 1.0 + lam * ac, q;
                                        if (base < 0.0) base = 0.0;
                                        if (zeta == 1.0) q = 1.0;
                                        else if (lam == 1.0) q = exp((zeta - 1.0) * lbp[X]);
                                        else if (lam == -1.0) q = exp((zeta - 1.0) * lbm[X]);
                                        else q = pow(base, zeta - 1.0);
                                        double P = q * base;
                                        double dP = zeta * lam * (c > 0 ? 1.0 : -1.0) * q;
                                        if (X == 0) accf[4 * f] += coef * P * E[0];
                                        sa += dP * E[X] * dcx[X] + P * Ex[X];
                                        sb += dP * E[X] * dcy[X] + P * Ey[X];
                                    }
                                    /* d x / d r_a = -u_j / x ; d y / d r_a = -u_k / y */
                                    double sx = -coef * sa / x, sy = -coef * sb / y;
                                    accf[4 * f + 1] += sx * nb.ux[tj] + sy * nb.ux[tk];
                                    accf[4 * f + 2] += sx * nb.uy[tj] + sy * nb.uy[tk];
                                    accf[4 * f + 3] += sx * nb.uz[tj] + sy * nb.uz[tk];
                                }
                            }
                        }
                    }
                }
                for (int f = 0; f < nf; ++f) {
                    Ga[f] = accf[4 * f];
                    dGa[3 * f] = accf[4 * f + 1];
                    dGa[3 * f + 1] = accf[4 * f + 2];
                    dGa[3 * f + 2] = accf[4 * f + 3];
                }
            }
            free(accf);
            free(nb.idx); free(nb.ux); free(nb.uy); free(nb.uz); free(nb.d);
            free(fcj); free(dfcj); free(psi); free(dpsi);
        }
    }
    free(cnt_tmp);
    free(grp_of); free(grp_kind); free(grp_rc); free(grp_eta); free(coef_f);
    free(row_off);
    free(cell_of);
    free(start);
    free(order);
}

/* Exported entry points (ctypes). */

int nl_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* descriptors for B frames of n atoms each; pos (B,n,3) wrapped; G (B,n,nf), dG (B,n,nf,3) */
int nl_oracle_descriptors(const double *pos, int B, int n, const int *kinds, const double *prm, int nf,
                          double *G, double *dG, int *nbr_count, int nthreads) {
    if (nthreads <= 0) nthreads = nl_oracle_max_threads();
    for (int b = 0; b < B; ++b)
        frame_eval(pos + (size_t)b * n * 3, n, kinds, prm, nf, G + (size_t)b * n * nf,
                   dG + (size_t)b * n * nf * 3, nbr_count ? nbr_count + (size_t)b * n : NULL,
                   NULL, 0, NULL, nthreads);
    return 0;
}

/* sorted pair set of ONE frame inside rc: returns the total number of pairs; fills up to cap */
int64_t nl_oracle_pairs(const double *pos, int n, double rc, int *pairs, int64_t cap, int nthreads) {
    if (nthreads <= 0) nthreads = nl_oracle_max_threads();
    int kind = KIND_G1;
    double prm[4] = {rc, 0, 0, 0};
    int64_t total = 0;
    frame_eval(pos, n, &kind, prm, 1, NULL, NULL, NULL, pairs, cap, &total, nthreads);
    return total;
}

/*
 * Distance of every atom from the reference's angular discontinuity.  The angular factor is switched off for
 * |cos| <= 1e-3 (symm_funcs.py:210-211): a triangle whose cosine sits within rounding of that threshold is decided
 * differently by two arithmetics (fp32 vs fp64), and the term that flips is of full size for lambda = -1.  For every atom
 * a, margin[a] = min over the closed triangles that contain a, over their three vertex cosines, of | |c| - 1e-3 |:
 * exactly the cosines that enter G_a and dg[a].  Parity tests between precisions compare atoms whose margin exceeds
 * the fp32 rounding of a cosine (like the pair-set tests keep away from d = rc).  Brute-force neighbour search: meant
 * for sub-blocks of ~1e4 atoms.  pos: n*3 (already wrapped).
 */
int nl_oracle_threshold_margin(const double *pos, int n, double rc, double *margin, int nthreads) {
    if (nthreads <= 0) nthreads = nl_oracle_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        int cap = 256, m;
        double *u = (double *)malloc(sizeof(double) * 4 * cap);
#pragma omp for schedule(dynamic, 16)
        for (int a = 0; a < n; ++a) {
            m = 0;
            for (int j = 0; j < n; ++j) {
                double ux = pos[3 * j] - pos[3 * a], uy = pos[3 * j + 1] - pos[3 * a + 1], uz = pos[3 * j + 2] - pos[3 * a + 2];
                if (fabs(ux) >= rc || fabs(uy) >= rc || fabs(uz) >= rc) continue;
                double d = sqrt(ux * ux + uy * uy + uz * uz);
                if (!(d < rc && d > 0.0)) continue;
                if (m == cap) { cap *= 2; u = (double *)realloc(u, sizeof(double) * 4 * cap); }
                u[4 * m] = ux; u[4 * m + 1] = uy; u[4 * m + 2] = uz; u[4 * m + 3] = d;
                ++m;
            }
            double best = 1e300;
            for (int tj = 0; tj < m; ++tj)
                for (int tk = tj + 1; tk < m; ++tk) {
                    double wx = u[4 * tj] - u[4 * tk], wy = u[4 * tj + 1] - u[4 * tk + 1], wz = u[4 * tj + 2] - u[4 * tk + 2];
                    double z = sqrt(wx * wx + wy * wy + wz * wz), x = u[4 * tj + 3], y = u[4 * tk + 3];
                    if (!(z < rc && z > 0.0)) continue;
                    double c[3] = {cos_law(x, y, z), cos_law(x, z, y), cos_law(y, z, x)};
                    for (int X = 0; X < 3; ++X) {
                        double g = fabs(fabs(c[X]) - 1e-3);
                        if (g < best) best = g;
                    }
                }
            margin[a] = best;
        }
        free(u);
    }
    return 0;
}
