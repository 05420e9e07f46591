/*
 * nnmd_b200.h -- C ABI of the B200-native descriptor -> energy -> force path.
 *
 * The reference (ded-otshelnik/nnmd) has no FFI of its own: its boundary is the set of
 * Python signatures in nnmd/features/__init__.py:13-23 and nnmd/nn/__init__.py:1-6.  The
 * entry points below are what a binding for that path attaches to; each one names the
 * reference code it replaces (paths relative to the reference checkout).  The Python host
 * (nnmd_b200/) reaches them through ctypes with raw device pointers (torch owns all the
 * device memory and the stream); INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every pointer marked "dev" is a CUDA device pointer, "host" a host pointer
 *   - `dtype` is NNMD_F32 or NNMD_F64: the working precision of positions, descriptors,
 *     gradients, energies and forces (the reference's `dtype` of the cartesians tensor)
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); every call
 *     only enqueues work, nothing here synchronises except where stated
 *   - return value: NNMD_OK or an NNMD_ERR_* code; nnmd_last_error() gives the text
 *   - there is no CPU implementation behind any of these: without a CUDA device the
 *     calls fail with NNMD_ERR_CUDA
 */
#ifndef NNMD_B200_H
#define NNMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNMD_F32 0
#define NNMD_F64 1

#define NNMD_OK 0
#define NNMD_ERR_INVALID 1  /* bad argument (text in nnmd_last_error) */
#define NNMD_ERR_CUDA 2     /* CUDA runtime error */
#define NNMD_ERR_CAPACITY 3 /* a neighbour row does not fit the shared-memory budget */

/* bits of the device-side status word `flag` (uint32, dev) */
#define NNMD_FLAG_NAN_G 1u  /* 0/0 or non-finite value in the normalised descriptors  (calculate_sf.py:106-107) */
#define NNMD_FLAG_NAN_DG 2u /* non-finite value in the descriptor gradients           (calculate_sf.py:108-109) */
#define NNMD_FLAG_OVERFLOW 4u /* a capacity-bound neighbour build (nnmd_nlist_build) did not fit the caller's buffers: the
                                 list is incomplete and every K2..K4 kernel given this flag word returns without writing */

/* symmetry-function kinds = nnmd/features/symm_funcs.py:6-15 (SymmetryFunction enum values) */
#define NNMD_G1 1
#define NNMD_G2 2
#define NNMD_G4 4
#define NNMD_G5 5

int nnmd_abi_version(void);
const char *nnmd_last_error(void);
/* number of SMs of the current device (grid sizing), or -1 */
int nnmd_device_sm_count(void);

/* Process-wide kernel-selection switches (explicit API calls, never the environment).
 *   "edge_form" 0|1  form of the edge-centric angular kernel: 1 (default) moment form -- per record and function three
 *                    ladder multiplications and six accumulations (the derivative of the radial product is taken out of the
 *                    sum over third atoms); 0 the derivative-record form of round 1 (eleven operations), kept for A/B
 *                    measurements.  Same results to fp32 round-off (profiles/README.md). */
int nnmd_set_option(const char *name, int value);

/* Instrumentation for bench.py: number of kernel launches issued by this library so far, and optional
 * CUDA-event timing per kernel class on the launching stream (classes: 0 nlist, 1 radial, 2 angular,
 * 3 normalize, 4 contract, 5 mlp, 6 wrap, 7 edge gather).  nnmd_profile_read synchronises the device and resets. */
#define NNMD_PROF_NCLASS 8
long long nnmd_launch_count(void);
/* NVTX ranges around every kernel class ("nnmd:K1 neighbour list", "nnmd:K3 angular", ...) for nsys / ncu --nvtx; off by default */
int nnmd_nvtx_enable(int on);
int nnmd_profile_enable(int on);
int nnmd_profile_read(double *ms, long long *count);

/* ---------------------------------------------------------------------------------------
 * K0  wrap-only periodic boundaries            replaces symm_funcs.py:53-82 map_positions_pbc
 *   pos    dev  (n_total,3) working dtype
 *   cell, inv_cell, pbc  host double[9], double[9], double[3]; values are ROUNDED to the working
 *          dtype inside (the host computes inv_cell in the working dtype like torch.linalg.inv does)
 *   has_cell  0 when det(cell)==0: positions pass through unchanged (symm_funcs.py:71-72)
 *   posw   dev  (n_total,4) working dtype, xyz + zero pad: the layout every later kernel reads
 * ------------------------------------------------------------------------------------- */
int nnmd_wrap_positions(int dtype, const void *pos, int64_t n_total, const double *cell, const double *inv_cell,
                        const double *pbc, int has_cell, void *posw, void *stream);

/* ---------------------------------------------------------------------------------------
 * K1  cell-list neighbour list     replaces symm_funcs.py:85-104 calculate_distances +
 *                                  symm_funcs.py:18-31 _pairwise_filter (dense (B,N,N) mask)
 * The list is CSR over centre atoms: row c = frame*n_centres + i holds the GLOBAL indices
 * frame*n_atoms + j of every j with 0 < |p_j - p_i| < rc, ascending.  Centres are the first
 * n_centres atoms of each frame (n_centres < n_atoms only for spatial shards that carry halo atoms).
 *
 *   step 1  nnmd_nlist_workspace_bytes -> size of `ws` (dev scratch, owned by the caller)
 *   step 2  nnmd_nlist_count : bins atoms, counts neighbours, writes offsets (int64, n_rows+1),
 *           edge_offsets (int64, n_rows+1, optional: prefix sums of the per-row number of neighbours with a LARGER
 *           global index -- every unordered pair has exactly one such "owner" row; used by the edge-centric
 *           angular kernel) and stats = {total pairs, longest row, total owned edges} (int64[3], dev)
 *   step 3  caller reads stats (the one host sync of the build), allocates idx (int32, total)
 *   step 4  nnmd_nlist_fill  : writes the sorted rows
 * ------------------------------------------------------------------------------------- */
int64_t nnmd_nlist_workspace_bytes(int64_t n_frames, int64_t n_atoms);
int nnmd_nlist_count(int dtype, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres, double rc,
                     void *ws, int64_t *offsets, int64_t *edge_offsets, int64_t *stats, void *stream);
int nnmd_nlist_fill(int dtype, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres, double rc,
                    const void *ws, const int64_t *offsets, int64_t max_row, int32_t *idx, void *stream);

/* K1 without a host synchronisation: count + fill in one call into buffers of FIXED capacity (what a previous build
 * needed, plus margin): idx int32[idx_cap], rows of at most max_row_cap entries, at most edge_cap owned edges.  When the
 * structure outgrows a capacity the build sets NNMD_FLAG_OVERFLOW in `flag` (dev uint32) instead of writing out of bounds;
 * the caller reads `flag` / `stats` whenever it next synchronises (e.g. with the NaN flag at the end of a step) and
 * rebuilds with larger buffers.  This is what lets an MD step run with no host round trip and be replayed from a CUDA graph.
 *   go   dev int32 or NULL.  *go == 0 turns every kernel of the build into a no-op: the previous list is REUSED.
 * stats here is int64[6]: {total pairs, longest row, owned edges} of the latest build, then the same three of the FIRST build
 * that raised NNMD_FLAG_OVERFLOW since the flag was last cleared (what to regrow to).
 * nnmd_skin_check decides it on the device (Verlet skin): the list is built with cutoff rc + skin and stays valid until some
 * atom has moved more than skin / 2 from the positions of the last build (pos_ref, (n_total,4), updated when *go becomes 1;
 * fill it with NaN to force the first build).  acc: dev uint32 scratch, zero it once.  n_builds: dev int64 counter or NULL.
 * The K2..K4 kernels re-test every listed neighbour against the functions' own cutoff, so a skin list gives the SAME
 * descriptors as an exact one (replaces the O(N^2) Python loop of nnmd/md/verlet.py:97-218, which has no list at all).
 *   row_scratch  dev, nnmd_nlist_row_scratch_bytes(n_rows, max_row_cap) bytes, or NULL.  With it the candidates are visited
 *                ONCE: sorted rows go to padded scratch rows, then counts are scanned and the rows compacted into idx (the
 *                separate count pass of the two-step build disappears). */
int64_t nnmd_nlist_row_scratch_bytes(int64_t n_rows, int64_t max_row_cap);
int nnmd_nlist_build(int dtype, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres, double rc, void *ws,
                     int64_t *offsets, int64_t *edge_offsets, int32_t *idx, int64_t idx_cap, int64_t max_row_cap,
                     int64_t edge_cap, void *row_scratch, int64_t *stats, uint32_t *flag, const int32_t *go, void *stream);
int nnmd_skin_check(int dtype, const void *posw, void *pos_ref, int64_t n_total, double skin, uint32_t *acc, int32_t *go,
                    int64_t *n_builds, void *stream);

/* ---------------------------------------------------------------------------------------
 * Symmetry-function plan: the parameter table of one species
 *   replaces the per-call `symm_funcs_data` walk of calculate_sf.py:64-81
 *   kinds  host int32[nf]  (NNMD_G1|G2|G4|G5; anything else -> NNMD_ERR_INVALID,
 *                           "Unknown symmetry function number: k" like calculate_sf.py:81)
 *   params host double[nf*4], the reference's POSITIONAL binding:
 *          G1 (rc) ; G2 (rc, eta, rs) ; G4/G5 (rc, eta, zeta, lambda)
 * ------------------------------------------------------------------------------------- */
typedef struct nnmd_sf_plan nnmd_sf_plan;
int nnmd_sf_plan_create(int dtype, int nf, const int32_t *kinds, const double *params, nnmd_sf_plan **plan);
/* Multi-species plan (EXTENSION: the reference evaluates every species in isolation, nnmd/nn/dataset.py:59-67, and its
 * training step raises for two species -- SURVEY Q9, 8f-3; there is no reference oracle, the tests use finite differences
 * and a dense restatement).  Every atom carries a species index in posw.w (nnmd_set_species), all atoms of all species are
 * neighbours of each other, and every function f weights its terms by species:
 *     radial   G_i,f = sum_j  wr_f[s_j] phi_f(d_ij)
 *     angular  G_i,f = 2^(1-zeta) sum_{j != k} pw_f[s_j][s_k] P(c_ijk) exp(..) fc fc fc          pw_f symmetric
 *   weights host double[nf * n_species * n_species]: block f holds pw_f row-major; radial functions use its first row as
 *   wr_f.  One-hot weights give the element-resolved functions of Behler's multi-species scheme (per-pair-type tables),
 *   species-dependent scalars give weighted symmetry functions.  dG keeps the reference's meaning:
 *   d(sum over ALL atoms i of G_i,f) / d r_a. */
int nnmd_sf_plan_create_species(int dtype, int nf, const int32_t *kinds, const double *params, int n_species,
                                const double *weights, nnmd_sf_plan **plan);
/* posw[:, 3] = species[i % n_atoms] for n_frames frames of n_atoms atoms (species: dev int32[n_atoms], values < n_species) */
int nnmd_set_species(int dtype, void *posw, const int32_t *species, int64_t n_frames, int64_t n_atoms, void *stream);
void nnmd_sf_plan_destroy(nnmd_sf_plan *plan);
double nnmd_sf_plan_rc_max(const nnmd_sf_plan *plan);

/* ---------------------------------------------------------------------------------------
 * K2 + K3 + K4a  descriptors and their summed gradients
 *   replaces symm_funcs.py:107-277 (f_cutoff, g1/g2/g4/g5_function) and the per-function
 *   torch.autograd.grad of calculate_sf.py:85-90
 *   G   dev (n_rows, nf)    UN-normalised symmetry functions, column order = plan order
 *   dG  dev (n_rows, nf, 3) d(sum_i G_i,f)/d r_a  (SURVEY Q6), or NULL to skip
 *   flag dev uint32, OR-ed with NNMD_FLAG_NAN_DG; when it carries NNMD_FLAG_OVERFLOW on entry the kernels do nothing
 * ------------------------------------------------------------------------------------- */
int nnmd_sf_eval(const nnmd_sf_plan *plan, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres,
                 const int64_t *offsets, const int32_t *idx, int64_t max_row, const int64_t *edge_offsets,
                 int64_t n_edges, void *edge_ws, int64_t edge_ws_bytes, void *G, void *dG, uint32_t *flag, void *stream);

/* Edge scratch of the edge-centric angular kernel (fp32 plans whose angular groups fill a warp with 2 functions per
 * lane): bytes needed for n_edges owned edges, 0 when the plan does not use it.  Pass edge_ws = NULL to
 * nnmd_sf_eval / nnmd_sf_force to force the vertex-centric kernel (no scratch, about twice the arithmetic). */
int64_t nnmd_sf_edge_scratch_bytes(const nnmd_sf_plan *plan, int64_t n_edges);

/* K5  L2 normalisation over the feature axis + NaN guard   replaces calculate_sf.py:100-109
 *   ghat may alias G */
int nnmd_sf_normalize(int dtype, const void *G, int64_t n_rows, int nf, void *ghat, uint32_t *flag, void *stream);

/* ---------------------------------------------------------------------------------------
 * K4b  fused force: F[a,:] (+)= -sum_f w[a,f] * d(sum_i G_i,f)/d r_a without materialising dG
 *   replaces the chain autograd.grad (calculate_sf.py:85-90) -> einsum (bpnn.py:353, :538-544)
 *   w      dev (n_rows, nf)  dE/dg_hat of the centre atom
 *   force  dev (n_rows, 3)   overwritten
 *   virial dev (n_frames, 9) or NULL: W = - sum_a (p_a - centroid) (x) F_a of the forces just computed (extension, the
 *          reference has no virial: SURVEY Q8; same as calling nnmd_virial afterwards with origin = NULL)
 * ------------------------------------------------------------------------------------- */
int nnmd_sf_force(const nnmd_sf_plan *plan, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres,
                  const int64_t *offsets, const int32_t *idx, int64_t max_row, const int64_t *edge_offsets,
                  int64_t n_edges, void *edge_ws, int64_t edge_ws_bytes, const void *w, void *force, void *virial,
                  void *stream);

/* same, with a status word: does nothing when *flag carries NNMD_FLAG_OVERFLOW (sync-free MD steps) */
int nnmd_sf_force_flag(const nnmd_sf_plan *plan, const void *posw, int64_t n_frames, int64_t n_atoms, int64_t n_centres,
                       const int64_t *offsets, const int32_t *idx, int64_t max_row, const int64_t *edge_offsets,
                       int64_t n_edges, void *edge_ws, int64_t edge_ws_bytes, const void *w, void *force, void *virial,
                       uint32_t *flag, void *stream);

/* Virial of each frame from per-atom forces, W = - sum_a (p_a - o) (x) F_a, (n_frames, 9) row-major.
 *   The reference has no virial and never adds periodic images (Q1): every frame is a finite cluster, for which
 *   this equals the pair-decomposed virial whenever sum_a F_a = 0.  origin: host double[3], or NULL = centroid of the
 *   frame's centre atoms.  ws: dev scratch of nnmd_virial_workspace_bytes(n_frames).  Works on any (n_rows,3) forces:
 *   the fused ones of nnmd_sf_force or the contracted ones of nnmd_force_contract. */
int64_t nnmd_virial_workspace_bytes(int64_t n_frames);
int nnmd_virial(int dtype, const void *posw, const void *force, int64_t n_frames, int64_t n_atoms, int64_t n_centres,
                const double *origin, void *ws, void *virial, void *stream);

/* K4c  force contraction against a materialised dG      replaces bpnn.py:353  einsum("ijk,ijkl->ijl")
 *   force[a,:] = -sum_f dE[a,f] * dG[a,f,:] */
int nnmd_force_contract(int dtype, const void *dE, const void *dG, int64_t n_rows, int nf, void *force, void *stream);

/* ---------------------------------------------------------------------------------------
 * K6  atomic MLP: energy per atom and dE/dx      replaces atomic_nn.py:43-48 AtomicNN.forward and
 *                                                the autograd.grad of bpnn.py:347-349 / :537-538
 *   n_layers linear layers; sizes host int32[n_layers+1] = (in, h1, ..., 1)
 *   weights / biases: host arrays of n_layers DEVICE pointers; W_l is (sizes[l+1], sizes[l]) row-major
 *   (torch.nn.Linear layout, state_dict keys model.{l}.weight|bias); sigmoid after every layer but the last
 *   x dev (n_rows, sizes[0]); e dev (n_rows); dEdx dev (n_rows, sizes[0]) or NULL
 *   precision: NNMD_MLP_EXACT = fp32/fp64 FMA; NNMD_MLP_TF32 = tensor cores (dtype F32 only)
 * ------------------------------------------------------------------------------------- */
#define NNMD_MLP_EXACT 0
#define NNMD_MLP_TF32 1
int nnmd_mlp_energy_grad(int dtype, int precision, int n_layers, const int32_t *sizes, const void *const *weights,
                         const void *const *biases, const void *x, int64_t n_rows, void *e, void *dEdx,
                         void *stream);

/* K5 + K6 + K4c in ONE pass (fp32, tensor cores with 3xTF32, two hidden layers <= 96-64-64): reads the UN-normalised G and
 * dG once, normalises per atom (calculate_sf.py:100-101; NNMD_FLAG_NAN_G on 0/0), runs the network forward and backward
 * in registers and contracts dE/dg_hat with dG into the force (bpnn.py:353) -- g_hat, dE/dg_hat never travel through HBM.
 *   ghat dev (n_rows, F) or NULL: normalised descriptors out (may alias G);  dEdx dev (n_rows, F) or NULL
 *   returns NNMD_ERR_CAPACITY when the shape has no instantiation: use nnmd_sf_normalize + nnmd_mlp_energy_grad + nnmd_force_contract */
int nnmd_energy_force_fused(int n_layers, const int32_t *sizes, const void *const *weights, const void *const *biases,
                            const void *G, const void *dG, int64_t n_rows, void *ghat, void *e, void *dEdx, void *force,
                            uint32_t *flag, void *stream);

/* ---------------------------------------------------------------------------------------
 * K7  training backward of the atomic MLP: gradients of a loss L(e, w = dE/dx) w.r.t. weights and biases, i.e. the
 *     double backward torch.autograd performs for bpnn.py:342-366 (energy term through e, force term through dE/dg).
 *   ge dev (n_rows)            dL/de per atom
 *   u  dev (n_rows, sizes[0])  dL/d(dE/dx) per atom, or NULL when the loss has no force term
 *   grad_weights / grad_biases: host arrays of n_layers DEVICE pointers, same shapes as the parameters;
 *   the kernel ADDS into them (atomicAdd): zero them first.  Activations are recomputed, nothing is stored.
 * K4c' backward of the force contraction: u[a,f] = - sum_c gf[a,c] dG[a,f,c]   (gf = dL/dF, (n_rows,3))
 * ------------------------------------------------------------------------------------- */
int nnmd_mlp_train_backward(int dtype, int n_layers, const int32_t *sizes, const void *const *weights,
                            const void *const *biases, const void *x, int64_t n_rows, const void *ge, const void *u,
                            void *const *grad_weights, void *const *grad_biases, void *stream);
int nnmd_force_contract_backward(int dtype, const void *gf, const void *dG, int64_t n_rows, int nf, void *u, void *stream);

/* ---------------------------------------------------------------------------------------
 * K8  training loss and its gradient in one launch.  Replaces bpnn.py:196-215 (`_loss`: e_coeff * MSELoss(e.sum(1), E)
 *     + f_coeff / 3 * MSELoss(f, F), criterion fixed to MSELoss by bpnn.py:164) and the part of loss.backward()
 *     (bpnn.py:366) that runs from the scalar down to the per-atom energies and forces.
 *   n_species blocks of atoms (species are concatenated on the atom axis), n_atoms host int32[n_species];
 *   e / f / ge / gf: host arrays of n_species DEVICE pointers;  e[s], ge[s] (n_frames, n_atoms[s]);
 *   f[s], gf[s] (n_frames, n_atoms[s], 3);  E dev (n_frames);  F dev (n_frames, sum n_atoms, 3), species in order;
 *   ge = dL/de, gf = dL/df;  out dev [3] = (energy term, force term, total), coefficients included.
 *   ws: nnmd_loss_workspace_bytes(n_frames) device bytes, ZEROED once before the first call (reusable afterwards).
 *   Deterministic: per-frame partial sums are reduced in a fixed order.
 * ------------------------------------------------------------------------------------- */
size_t nnmd_loss_workspace_bytes(int64_t n_frames);
int nnmd_loss_mse(int dtype, int n_species, const int32_t *n_atoms, const void *const *e, const void *const *f,
                  const void *E, const void *F, int64_t n_frames, double e_coeff, double f_coeff, void *const *ge,
                  void *const *gf, void *ws, void *out, void *stream);

/* ---------------------------------------------------------------------------------------
 * Batch assembly for the training step (replaces the DataLoader collation of bpnn.py:234-236 on the cached
 * tensors of dataset.py:33-123): dst[k][i, :] = src[k][sel[i], :] for n_arrays row-major arrays in one launch.
 *   src / dst: host arrays of DEVICE pointers; row_bytes host int64[n_arrays] (multiples of 4); sel dev int64[n_sel]
 * ------------------------------------------------------------------------------------- */
int nnmd_gather_rows(int n_arrays, const void *const *src, void *const *dst, const int64_t *row_bytes,
                     const int64_t *sel, int64_t n_sel, void *stream);
/* The same gather driven from the device: rows perm[*cursor .. *cursor + n_sel), then *cursor += n_sel.  `perm` is an
 * epoch's frame order, `cursor` a device int64; a captured training step that starts with this call needs no per-batch
 * arguments (replaces the DataLoader iteration of nnmd/nn/bpnn.py:321-329 without a host round trip per batch). */
int nnmd_gather_rows_cursor(int n_arrays, const void *const *src, void *const *dst, const int64_t *row_bytes,
                            const int64_t *perm, int64_t *cursor, int64_t n_sel, void *stream);

/* ---------------------------------------------------------------------------------------
 * Peer windows: the two collectives of the path as ordinary kernels over NVLink peer memory (csrc/peer.cu).
 * One process per GPU; every rank creates a window of the same size, the ranks exchange the CUDA-IPC handles
 * (nnmd_peer_handle_bytes() bytes each; nnmd_b200/dist.py does it through torch.distributed), and connect.
 * The kernels synchronise through step numbers stored in the windows, so they can be captured into the CUDA graph of
 * the step they belong to; every rank must issue the same sequence of calls.  A peer that never arrives raises the
 * window's status word after ~2 s instead of hanging the device.
 *   nnmd_peer_allreduce_weighted_f32  replaces the NCCL all-reduce of the flat MLP gradient buffer of the data-parallel
 *       form of the reference's optimizer step (nnmd/nn/bpnn.py:366-368): buf[0..n-1) = this rank's batch-mean gradient,
 *       buf[n-1] = its frame count b_r;  afterwards buf[i] = sum_r b_r buf_r[i] / sum_r b_r, buf[n-1] = sum_r b_r,
 *       bit-identical on all ranks.
 *   nnmd_peer_halo_exchange  one halo round of the x-slab decomposition (nnmd_b200/dist.py:SlabShard): rows of `width`
 *       4-byte words; own[send_idx[k]] goes to row dst_slot[k] of the halo of rank dst_peer[k];
 *       local = [ own (n_own rows) | halo (n_halo rows) ].
 */
typedef struct nnmd_peer nnmd_peer;
int nnmd_peer_handle_bytes(void);
int nnmd_peer_create(int rank, int world, int64_t cap_bytes, nnmd_peer **out);
int nnmd_peer_export(nnmd_peer *p, void *handle_out);
int nnmd_peer_connect(nnmd_peer *p, const void *handles);
void nnmd_peer_destroy(nnmd_peer *p);
int64_t nnmd_peer_capacity(const nnmd_peer *p);
int nnmd_peer_status(nnmd_peer *p, int *status);
int nnmd_peer_allreduce_weighted_f32(nnmd_peer *p, float *buf, int64_t n, void *stream);
/* the same with the rank's weight b_r as an argument (it is what gets summed into buf[n-1]) and, optionally, the loss
 * accounting of the step: loss_acc[0..3) += loss_in[0..3) * b_r / sum_r b_r (both NULL to skip) */
int nnmd_peer_allreduce_step_f32(nnmd_peer *p, float *buf, int64_t n, float weight, const float *loss_in, float *loss_acc,
                                 void *stream);
/* plain sum of buf[0..n) over the ranks, in place (the total energy of a spatially sharded structure) */
int nnmd_peer_allreduce_sum_f32(nnmd_peer *p, float *buf, int64_t n, void *stream);
int nnmd_peer_halo_exchange(nnmd_peer *p, const void *own, int64_t n_own, const int64_t *send_idx, const int32_t *dst_peer,
                            const int64_t *dst_slot, int64_t n_send, int64_t n_halo, int width, void *local, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NNMD_B200_H */
