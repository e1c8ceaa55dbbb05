/*
 * m3gnet_b200.h — C ABI of the B200 (sm_100a) energy+forces hot path of M3GNet.
 *
 * The reference (lan496/torch-m3gnet) is 100 % Python: it has no FFI/plugin layer to mirror, so this
 * header *defines* the boundary a maintainer would bind (ctypes stub in INTEGRATION.md).  Each entry
 * point names the reference code it replaces (paths relative to /root/reference/src/torch_m3gnet/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; buffers are caller-owned
 *     (torch caching allocator); no entry point allocates, frees or synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = success, otherwise a negative M3G_ERR_* code; m3g_last_error() gives the
 *     message of the last failure on the calling thread;
 *   - floating tensors are float32 row-major; index tensors are int32 unless noted (the int64
 *     tensors of the reference API are narrowed once per batch by m3g_narrow_i64);
 *   - edges are grouped by source atom (ascending), `edge_ptr` (N+1) is that CSR;
 *     `in_ptr`/`in_perm` is the CSR of edge ids grouped by destination atom (ascending edge id
 *     inside a row); triplets are a CSR per bond: `tri_ptr` (E+1) rows = first bond e1,
 *     `tri_e2` (T) columns = second bond; `trt_ptr`/`trt_e1` is its transpose (rows = e2);
 *   - sizes: N atoms, E directed edges, T triplets, B structures, F feature width (node == edge),
 *     R = n_max radial functions, L = l_max, D = L*R.
 *   - accumulation order (stated, deterministic): every segmented sum is evaluated by one warp
 *     (or sub-warp group of G lanes): lane g adds elements g, g+G, g+2G, ... of the segment in
 *     ascending order, then lanes are combined by a butterfly (xor 16,8,4,2,1).  No float atomics
 *     on the default path.
 */
#ifndef M3GNET_B200_H
#define M3GNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3G_OK 0
#define M3G_ERR_INVALID (-1) /* bad argument (null pointer, unsupported size) */
#define M3G_ERR_CUDA (-2)    /* a CUDA runtime call / launch failed */
#define M3G_ERR_UNSUPPORTED (-3)

#define M3G_MAX_F 128 /* feature width supported by the kernels */
#define M3G_MAX_L 9   /* l_max supported by the kernels = the reference's range (nn/interaction.py:250-253) */
#define M3G_MAX_R 10  /* n_max supported by the three-body kernels = the reference's range */
#define M3G_MAX_RADIAL 10 /* n_max supported by the radial (edge) basis */

const char* m3g_last_error(void);
int m3g_abi_version(void);
/* Device properties the host side sizes grids with: out[0]=SM count, out[1]=cc major, out[2]=cc minor */
int m3g_device_info(int* out_host);

/* ---------------------------------------------------------------------------------------------
 * Graph preparation (once per batch; replaces nothing in the reference — it is the canonical
 * form the reference's `from_structure` output already has, data/material_graph.py:182-187,239-248)
 * ------------------------------------------------------------------------------------------- */
int m3g_narrow_i64(const int64_t* in, int32_t* out, int64_t n, void* stream);
/* flags[0] = 1 if keys is non-decreasing, flags[1] = 1 if every key is in [0, n_rows) */
int m3g_check_sorted(const int32_t* keys, int64_t n, int64_t n_rows, int32_t* flags, void* stream);
/* row_ptr[r] = first position whose key >= r (keys sorted); row_ptr has n_rows+1 entries */
int m3g_csr_from_sorted(const int32_t* keys, int64_t n, int64_t n_rows, int32_t* row_ptr, void* stream);
/* Stable counting sort of positions 0..n-1 by key: row_ptr (n_rows+1), perm (n) with ascending
 * positions inside each row.  `work` must hold n_rows+1 + m3g_scan_work_elems(n_rows) int32. */
int m3g_csr_by_key(const int32_t* keys, int64_t n, int64_t n_rows, int32_t* row_ptr, int32_t* perm,
                   int32_t* work, void* stream);
/* out[p] = vals[perm[p]] */
int m3g_gather_i32(const int32_t* vals, const int32_t* perm, int64_t n, int32_t* out, void* stream);
/* every row of (row_ptr, cols) is sorted ascending in place (rows are short: one neighbour shell) */
int m3g_sort_rows(const int32_t* row_ptr, int64_t n_rows, int32_t* cols, void* stream);
/* flags[0] = 1 iff (ptr_a, col_a) == transpose(ptr_a, col_a), i.e. (e1,e2) present <=> (e2,e1) present */
int m3g_csr_is_symmetric(const int32_t* row_ptr, const int32_t* cols, int64_t n_rows, int32_t* flags,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Graph construction on the GPU (replaces data/material_graph.py:168-193 [pymatgen
 * Structure.get_all_neighbors] and :196-254 [compute_threebody]).  Batched over B structures.
 *   lattice (B,3,3) f64 rows = lattice vectors, cart (N,3) f64, atom_ptr (B+1) int32.
 * Two-pass: count → caller allocates → fill.  Accept rule d^2 < r^2 + 1e-8 in float64, zero-length
 * self pairs dropped; images are relative to the unwrapped input coordinates; edges of atom i are
 * ordered by (j, s0, s1, s2) ascending.
 * ------------------------------------------------------------------------------------------- */
/* Optional cell list: bins (B,3) int32 = bins per lattice axis of each structure (0,0,0 = sweep the whole structure;
 * otherwise every axis needs >= 3 bins whose perpendicular width is >= cutoff), bin_base (B+1) = offset of each
 * structure's bins.  m3g_nbr_bin_count writes atom_bin (N) and adds to bin_count (zeroed by the caller); after an
 * exclusive scan into bin_ptr, m3g_nbr_bin_fill lists the atoms of every bin (bin_cursor zeroed by the caller).
 * The bins only restrict the candidate atoms: accept test, edge set and edge order are unchanged.  Pass
 * bins = NULL to m3g_nbr_count / m3g_nbr_fill for the plain sweep. */
int m3g_nbr_bin_count(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                      double cutoff, const int32_t* bins, const int32_t* bin_base, int32_t* atom_bin,
                      int32_t* bin_count, void* stream);
int m3g_nbr_bin_fill(const int32_t* atom_bin, const int32_t* bin_ptr, int64_t N, int32_t* bin_cursor,
                     int32_t* bin_atoms, void* stream);
int m3g_nbr_count(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                  double cutoff, const int32_t* bins, const int32_t* bin_base, const int32_t* bin_ptr,
                  const int32_t* bin_atoms, int32_t* edge_count /* (N) */, void* stream);
int m3g_exclusive_scan_i32(const int32_t* in, int32_t* out /* n+1 */, int64_t n, int32_t* work, void* stream);
int64_t m3g_scan_work_elems(int64_t n);
int m3g_nbr_fill(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                 double cutoff, double threebody_cutoff, const int32_t* bins, const int32_t* bin_base,
                 const int32_t* bin_ptr, const int32_t* bin_atoms, const int32_t* edge_ptr /* (N+1) */, int64_t E,
                 int64_t* edge_index /* (2,E) */, int32_t* edge_shift /* (E,3) */, float* edge_dist /* (E) */,
                 int32_t* member /* (E) 1 if float32(d) <= float32(r3) */, void* stream);

/* Verlet (skin) list for MD / relaxation loops (the reference rebuilds the pymatgen neighbour
 * list for every structure it sees, data/material_graph.py:168-193).  The candidate list is an ordinary neighbour
 * list built once with cutoff + skin (m3g_nbr_count / m3g_nbr_fill): cand_ptr (N+1), cand_j (C) int32, cand_shift
 * (C,3).  m3g_verlet_displacement writes max_i |cart_i - ref_i|^2 (f32, rounded up; +inf for NaN input) so the host
 * rebuilds the candidates once it exceeds (skin/2)^2.  While it does not, m3g_verlet_count / m3g_verlet_fill give
 * exactly the output of m3g_nbr_count / m3g_nbr_fill on the new coordinates (same float64 accept test on the
 * candidates, same (j, s0, s1, s2) order).  The lattice must be the one the candidates were built with. */
int m3g_verlet_displacement(const double* cart, const double* ref_cart, int64_t N, float* max_d2 /* (1) */,
                            void* stream);
int m3g_verlet_count(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                     double cutoff, const int32_t* cand_ptr, const int32_t* cand_j, const int32_t* cand_shift,
                     int32_t* edge_count /* (N) */, void* stream);
int m3g_verlet_fill(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                    double cutoff, double threebody_cutoff, const int32_t* cand_ptr, const int32_t* cand_j,
                    const int32_t* cand_shift, const int32_t* edge_ptr, int64_t E, int64_t* edge_index /* (2,E) */,
                    int32_t* edge_shift /* (E,3) */, float* edge_dist /* (E) */, int32_t* member /* (E) */,
                    void* stream);
/* per-atom member degree n3 -> num_triplet_i (N) int64 = n3(n3-1), num_triplet_ij (E) int32,
 * tri_count (E) = triplets whose first bond is e (n3-1 for member edges else 0), and member_list (E):
 * the member edges of atom i compacted (ascending) at positions edge_ptr[i]..
 * Optional (NULL to skip): used_count (N) = bonds of atom i that head a triplet (n3 if n3 >= 2 else 0) and
 * stats int64[3], zeroed by the caller: += T, max n3, += sum of used_count — one read-back sizes tri_e2,
 * the per-atom kernels' capacity and the member-bond list */
int m3g_triplet_count(const int32_t* edge_ptr, const int32_t* member, int64_t N, int64_t E,
                      int64_t* num_triplet_i, int32_t* num_triplet_ij, int32_t* tri_count, int32_t* member_list,
                      int32_t* used_count, int64_t* stats, void* stream);
/* tri_ptr = exclusive scan of tri_count.  Fills tri_e2 (T) int32 CSR columns and, if triplet_index != NULL,
 * the reference's (2,T) int64 list in the reference's order (atom, e1, e2).  Optional: member_edges = ascending
 * list of the bonds that head a triplet, written at used_ptr (N+1) = exclusive scan of used_count */
int m3g_triplet_fill(const int32_t* edge_ptr, const int32_t* tri_ptr, const int32_t* tri_count,
                     const int32_t* member_list, int64_t N, int64_t T, int32_t* tri_e2, int64_t* triplet_index,
                     const int32_t* used_ptr, int32_t* member_edges, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Geometry + bases
 * ------------------------------------------------------------------------------------------- */
/* nn/scale.py:24-29 — out = in / length_scale (true division) */
int m3g_scale_fwd(const float* in, float* out, int64_t n, float length_scale, void* stream);
/* nn/invariant.py:44-59 — vec4[e] = (pos[dst]+shift·lattice[batch[src]]-pos[src], |v|), dist[e] = |v| */
int m3g_geometry_fwd(const float* pos, const float* lattice, const int32_t* batch, const int32_t* src,
                     const int32_t* dst, const int32_t* shift, int64_t E, float* vec4, float* dist, void* stream);
/* nn/invariant.py:33-40 — cos[t] = clamp(v1·v2/(r1 r2), -1, 1) in the caller's triplet order (int64 list) */
int m3g_angles_fwd(const float* vec4, const int64_t* tri_index /* (2,T) */, int64_t T, float* cos_out, void* stream);
/* adjoint of m3g_angles_fwd: adds into g_vec4 (E,4) with float atomics (rare path: only when a caller
 * differentiates through `triplet_angles`) */
int m3g_angles_bwd(const float* vec4, const int64_t* tri_index, const float* g_cos, int64_t T, float* g_vec4,
                   void* stream);
/* adjoint of m3g_geometry_fwd w.r.t. pos: g_e = g_vec4.xyz + (g_vec4.w + g_dist)·v/r;
 * g_pos[i] = scale * (sum_{e in in(i)} g_e - sum_{e in out(i)} g_e).  g_dist may be NULL. */
int m3g_geometry_bwd(const float* vec4, const float* g_vec4, const float* g_dist, const int32_t* edge_ptr,
                     const int32_t* in_ptr, const int32_t* in_perm, int64_t N, float scale, float* g_pos,
                     void* stream);
/* nn/featurizer.py:81-100 — h (E,R).  consts: [k_0..k_R | coeff_0..coeff_{R-1} | a_0..a_{R-1} | b_0..b_{R-1}]
 * with k_m = (m+1)·pi/rc, a_m = sqrt(e_m/d_{m-1}), b_m = sqrt(d_m), all produced on the host by the
 * reference's float32 op sequence. */
int m3g_radial_fwd(const float* dist, const float* consts, int64_t E, int R, float* h, void* stream);
int m3g_radial_bwd(const float* dist, const float* consts, const float* g_h, int64_t E, int R, float* g_dist,
                   void* stream);
/* nn/atom_ref.py:25-29 and nn/featurizer.py:33-38 (one-hot · Linear == column gather of W (F,num_types)) */
int m3g_atomref_fwd(const float* table, const int32_t* types, int64_t N, float* out, void* stream);
int m3g_embed_fwd(const float* weight /* (F,num_types) */, const int32_t* types, int64_t N, int F, int num_types,
                  float* x, void* stream);
/* nn/featurizer.py:128-132 — e0 = SiLU(h · Wt), Wt (R,F) = weight^T */
int m3g_edge_adjust_fwd(const float* h, const float* Wt, int64_t E, int R, int F, float* e0, void* stream);
int m3g_edge_adjust_bwd(const float* h, const float* Wt, const float* g_e0, int64_t E, int R, int F, float* g_h,
                        void* stream);

/* nn/interaction.py:284-350 (SphericalBessel fwd + its custom derivative), :353-382 (LegendreCosPolynomial
 * fwd / bwd with the reference's grad_output-per-level rule), :389-400 (cutoff_function) — elementwise,
 * kept for the operator API (`spherical_bessel`, `legendre_cos`, `cutoff_function`).  dout may be NULL. */
int m3g_sph_bessel(const float* x, int order, int64_t n, float* out, float* dout, void* stream);
int m3g_legendre(const float* x, int order, int64_t n, float* out, void* stream);
int m3g_legendre_bwd(const float* x, const float* go, int order, int64_t n, float* gx, void* stream);
int m3g_cutoff(const float* r, float rc, int64_t n, float* out, float* dout, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ThreeBodyInteration (nn/interaction.py:187-223) — D = L*R, d = l*R + n
 * tb_consts: [zeros_ln (D) | factors_ln (D) | rc | r3]  (2*D + 2 floats; chi = j_l(z r / rc) / factors)
 * ------------------------------------------------------------------------------------------- */
/* sig (N,D) = sigmoid(x · Ws^T + bs), Ws (D,F) = linear_sigmoid1.weight */
int m3g_tb_sigma_fwd(const float* x, const float* Ws, const float* bs, int64_t N, int F, int D, float* sig,
                     void* stream);
/* per edge: bas[e][d] = chi_d(r_e)·fc(r_e)·sig[dst[e]][d]   (chi uses rc, fc uses r3; quirk Q5) */
int m3g_tb_edge_basis_fwd(const float* vec4, const int32_t* dst, const float* sig, const float* tb_consts,
                          int64_t E, int L, int R, const int32_t* edge_list, int64_t n_list, float* bas, void* stream);
/* red[e1][d] = fc(r_e1)·sum_{e2 in tri(e1)} Y_l(cos(e1,e2))·bas[e2][d];
 * e_out = e_in + SiLU(red·WdT) * sigmoid(red·WgT), WdT/WgT (D,F) */
int m3g_tb_reduce_fwd(const float* vec4, const float* bas, const int32_t* tri_ptr, const int32_t* tri_e2,
                      const float* tb_consts, const float* WdT, const float* WgT, const float* e_in, int64_t E,
                      int L, int R, int F, int group, float* red, float* e_out, void* stream);
/* g_red (E,D) from g_e (E,F): adjoint of the bias-free 1-layer GatedMLP (forward recomputed from red); rows
 * of bonds without triplets (tri_ptr[e+1] == tri_ptr[e]) are zero */
int m3g_tb_gate_bwd(const float* red, const float* g_e, const float* WdT, const float* WgT, const int32_t* tri_ptr,
                    int64_t E, int D, int F, float* g_red, void* stream);
/* gather-form adjoint of the triplet sum (uses tri CSR for "e as first bond" and its transpose for
 * "e as second bond"; they may alias when the list is symmetric).  Outputs: g_vec4 (E,4) (xyz from
 * cos, w from cos and fc(r_e1)), g_bas (E,D).  Legendre backward follows the reference (quirk Q3). */
int m3g_tb_reduce_bwd(const float* vec4, const float* bas, const float* g_red, const int32_t* tri_ptr,
                      const int32_t* tri_e2, const int32_t* trt_ptr, const int32_t* trt_e1, const float* tb_consts,
                      int64_t E, int L, int R, int group, float* g_vec4, float* g_bas, void* stream);
/* adjoint of m3g_tb_edge_basis_fwd: g_vec4[e].w += d/dr terms; g_sig_e (E,D) per-edge sigma gradient.
 * edge_list (optional, both directions): only the n_list listed bonds (the member bonds of the triplet list) are
 * evaluated — bas rows of other bonds are never read by the reduce kernels; for the adjoint the caller zero-fills
 * g_sig_e, whose unlisted rows stay zero. */
int m3g_tb_edge_basis_bwd(const float* vec4, const int32_t* dst, const float* sig, const float* g_bas,
                          const float* tb_consts, int64_t E, int L, int R, const int32_t* edge_list, int64_t n_list,
                          float* g_vec4, float* g_sig_e, void* stream);
/* g_x (N,F) = base (N,F or NULL) + (sum_{e in in(k)} g_sig_e[e]) * sig(1-sig) · Ws, Ws (D,F) */
int m3g_tb_sigma_bwd(const float* g_sig_e, const int32_t* in_ptr, const int32_t* in_perm, const float* sig,
                     const float* Ws, const float* base, int64_t N, int F, int D, float* g_x, void* stream);

/* Canonical triplet layout (what compute_threebody, data/material_graph.py:239-248, emits): for every atom the
 * rows of its member bonds (bonds with a non-empty triplet row) list exactly all other member bonds of that atom,
 * ascending.  m3g_tri_dense_check: flags[0] = 1 if the CSR has that layout, flags[1] = max members per atom.
 * m3g_tb_atom_fwd / _bwd (csrc/threebody_atom.cu; replaces nn/interaction.py:187-223 and its autograd; l_max = n_max =
 * 3, F = 64, members per atom <=
 * m3g_tb_atom_capacity()) evaluate the three-body op with one warp per centre atom out of shared memory: no
 * triplet index list is read.  fwd writes red for member bonds only and e_out for all bonds; bwd writes g_vec4 for
 * all bonds (zeros for non-members) and g_bas for member bonds only (pair it with the member edge_list of
 * m3g_tb_edge_basis_bwd).  The gradient w.r.t. e_in is g_e itself. */
int m3g_tri_dense_check(const int32_t* src, const int32_t* edge_ptr, const int32_t* tri_ptr, const int32_t* tri_e2,
                        int64_t E, int32_t* flags, void* stream);
int m3g_tb_atom_capacity(void);
int m3g_tb_atom_fwd(const float* vec4, const float* bas, const int32_t* edge_ptr, const int32_t* tri_ptr, float r3,
                    const float* WdT, const float* WgT, const float* e_in, int64_t N, int n_sm, float* red,
                    float* e_out, void* stream);
int m3g_tb_atom_bwd(const float* vec4, const float* bas, const float* red, const float* g_e, const int32_t* edge_ptr,
                    const int32_t* tri_ptr, float r3, const float* WdT, const float* WgT, int64_t N, int n_sm,
                    float* g_vec4, float* g_bas, void* stream);

/* O(n3)-per-atom moment form of the same op (csrc/threebody_moment.cu; replaces nn/interaction.py:187-223, :353-382 and
 * their autograd for l_max = n_max = 3, F = 64 and the canonical triplet layout, members per atom <=
 * m3g_tb_mom_capacity()).  m3g_tb_radial (once per step; nn/interaction.py:226-281, :389-400): G[e][d] = chi_d(r_e) fc(r_e)
 * and dG/dr for the listed member bonds.  m3g_tb_mom_fwd: red (member bonds) and e_out (all bonds) with bas = G *
 * sig[dst] formed in-kernel.  m3g_tb_mom_bwd: g_vec4 (E,4) complete (cos, fc' and radial chain; zeros for
 * non-members; accumulate != 0 adds to the existing rows: the sum over the model's blocks) and g_sig_e (E,9) (zeros for
 * non-members, written by the accumulate == 0 call: with accumulate != 0 both buffers must come from such a call on the
 * same bond list), ready for m3g_tb_sigma_bwd.  The Legendre adjoint follows
 * the reference's quirk (Q3) through second-order moments.  max_members: upper bound of member bonds per atom. */
int m3g_tb_radial(const float* vec4, const float* tb_consts, int64_t E, int L, int R, const int32_t* edge_list,
                  int64_t n_list, float* G, float* dG, void* stream);
int m3g_tb_mom_capacity(void);
/* m3g_tb_sigma_fwd / _bwd specialised for F = 64, D = 9 (weights in registers, one shared butterfly for the nine sums,
 * warps striding over atoms; n_sm sizes the persistent grid) */
int m3g_tb_sigma64_fwd(const float* x, const float* Ws, const float* bs, int64_t N, int n_sm, float* sig, void* stream);
int m3g_tb_sigma64_bwd(const float* g_sig_e, const int32_t* in_ptr, const int32_t* in_perm, const float* sig,
                       const float* Ws, const float* base, int64_t N, int n_sm, float* g_x, void* stream);
int m3g_tb_mom_fwd(const float* vec4, const float* G, const float* sig, const int32_t* dst, const int32_t* edge_ptr,
                   const int32_t* tri_ptr, float r3, const float* WdT, const float* WgT, const float* e_in, int64_t N,
                   int max_members, int n_sm, float* red, float* e_out, void* stream);
/* the forward in two launches (default): m3g_tb_mom_red = the per-atom moment part only (red for member bonds), then
 * m3g_tb_edge_update = e_out = e_in + SiLU(red WdT) * sigmoid(red WgT) streamed over all bond rows (members by
 * tri_ptr), independent of the atom structure and bound by the 512 B per bond it moves */
int m3g_tb_mom_red(const float* vec4, const float* G, const float* sig, const int32_t* dst, const int32_t* edge_ptr,
                   const int32_t* tri_ptr, float r3, int64_t N, int max_members, int n_sm, float* red, void* stream);
int m3g_tb_edge_update(const float* red, const int32_t* tri_ptr, const float* WdT, const float* WgT, const float* e_in,
                       int64_t E, int n_sm, float* e_out, void* stream);
/* the same with e_in = SiLU(h WaT) (the EdgeAdjustor's output, nn/featurizer.py:84-96; h (E,3), WaT (3,64)) formed
 * in-kernel: the first block of the whole-step executor neither writes nor re-reads e0 (bit-identical to
 * m3g_edge_adjust_fwd followed by m3g_tb_edge_update) */
int m3g_tb_edge_update_h(const float* red, const int32_t* tri_ptr, const float* WdT, const float* WgT, const float* h,
                         const float* WaT, int64_t E, int n_sm, float* e_out, void* stream);
int m3g_tb_mom_bwd(const float* vec4, const float* G, const float* dG, const float* sig, const int32_t* dst,
                   const float* red, const float* g_e, const int32_t* edge_ptr, const int32_t* tri_ptr, float r3,
                   const float* WdT, const float* WgT, int64_t N, int max_members, int n_sm, int accumulate,
                   float* g_vec4, float* g_sig_e, void* stream);
/* the backward in two launches (default): m3g_tb_mlp_adj = adjoint of the 9 -> 64 gated MLP (nn/interaction.py:219-220,
 * nn/core.py:61-62) over the packed member-bond list, q[e] = dL/dred[e] (a lane owns a row: no cross-lane sums; q may
 * alias red); then m3g_tb_mom_bwd_q = the per-atom moment part of m3g_tb_mom_bwd reading q (no weights, no edge rows) */
int m3g_tb_mlp_adj(const float* red, const float* g_e, const int32_t* member_edges, int64_t n_members, const float* WdT,
                   const float* WgT, int n_sm, float* q, void* stream);
int m3g_tb_mom_bwd_q(const float* vec4, const float* G, const float* dG, const float* sig, const int32_t* dst,
                     const float* q, const int32_t* edge_ptr, const int32_t* tri_ptr, float r3, int64_t N,
                     int max_members, int n_sm, int accumulate, float* g_vec4, float* g_sig_e, void* stream);

/* Specialised variants for the default model shape l_max = n_max = 3, F = 64 (compile-time loops, gated-MLP
 * weights in shared memory, vector row I/O, persistent grid of 8 x n_sm blocks).  Same results and buffers as
 * the generic entry points above; r3 = three-body cutoff.  m3g_tb_reduce_bwd_sym requires a symmetric triplet
 * list ((e1,e2) present <=> (e2,e1) present — always true for graphs from from_structure). */
int m3g_tb_reduce_fwd_fast(const float* vec4, const float* bas, const int32_t* tri_ptr, const int32_t* tri_e2,
                           float r3, const float* WdT, const float* WgT, const float* e_in, int64_t E, int group,
                           int n_sm, float* red, float* e_out, void* stream);
int m3g_tb_gate_bwd_fast(const float* red, const float* g_e, const float* WdT, const float* WgT,
                         const int32_t* tri_ptr, int64_t E, int n_sm, float* g_red, void* stream);
int m3g_tb_reduce_bwd_sym(const float* vec4, const float* bas, const float* g_red, const int32_t* tri_ptr,
                          const int32_t* tri_e2, float r3, int64_t E, int group, int n_sm, float* g_vec4, float* g_bas,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * Generic row-wise linear: out (n,M) = in (n,K) · Wt (K,M) + bias (M or NULL).  Used for the per-atom
 * first-layer projections of M3GNetConv (nn/conv.py:92-97 concat split: [x_i,x_j,e]·W = x_i·W_i + x_j·W_j + e·W_e)
 * ------------------------------------------------------------------------------------------- */
int m3g_linear_fwd(const float* in, const float* Wt, const float* bias, int64_t n, int K, int M, float* out,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Elementwise pieces of a GatedMLP called on its own (nn/core.py:45-62): activation kind 0 = SiLU, 1 = sigmoid;
 * m3g_act_bwd: out = g * act'(in); m3g_mul: out = a * b.  (csrc/elementwise.cu)
 * ------------------------------------------------------------------------------------------- */
int m3g_act_fwd(const float* in, int64_t n, int kind, float* out, void* stream);
int m3g_act_bwd(const float* in, const float* g, int64_t n, int kind, float* out, void* stream);
int m3g_mul(const float* a, const float* b, int64_t n, float* out, void* stream);
/* out = a + b (out may alias a or b) */
int m3g_add(const float* a, const float* b, int64_t n, float* out, void* stream);
/* out[i] = sum_k slices[k * n + i], k ascending */
int m3g_sum_slices(const float* slices, int K, int64_t n, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * M3GNetConv (nn/conv.py:63-97, nn/core.py:6-62).  One "gated MLP on edges":
 *   z1 = P[src][po .. po+2F) + P[dst][po+2F .. po+4F) + e · W1eT            (2F: dense | gate)
 *   a1 = SiLU(z1); z2d = a1d·W2dT + b2d; z2g = a1g·W2gT + b2g
 *   out = SiLU(z2d) * sigmoid(z2g) * (h · WhT)
 * P (N, ldp) holds the per-atom projections (bias of layer 1 folded into the src part).
 * mode 0 (edge update, conv.py:68-75): y = e + out, written to y (E,F)
 * mode 1 (node update, conv.py:77-89): y = out (messages); m3g_segment_sum_add then adds them per source atom
 * ------------------------------------------------------------------------------------------- */
int m3g_conv_mlp_fwd(const float* P, int ldp, int po, const int32_t* src, const int32_t* dst, const float* e,
                     const float* h, const float* W1eT, const float* W2dT, const float* b2d, const float* W2gT,
                     const float* b2g, const float* WhT, int64_t E, int F, int R, int mode, float* y, void* stream);
/* out (N,F) = base (N,F) + sum_{e in [edge_ptr[i], edge_ptr[i+1])} msg[e] */
int m3g_segment_sum_add(const float* base, const float* msg, const int32_t* edge_ptr, int64_t N, int F,
                        float* out, void* stream);
/* the same sum from the partial rows of m3g_conv_tc_fwd(mode = 2): the tensor-core forward reduces the messages of every
 * 32-row block per source atom in its epilogue and writes one partial row per (block, atom) at the block's first row of
 * that atom (the other rows of `part` are never written nor read); out[i] = base[i] + sum_k part[max(b_i, 32 k)] */
int m3g_segment_sum_parts(const float* base, const float* part, const int32_t* edge_ptr, int64_t N, int F,
                          float* out, void* stream);
/* adjoint of one gated MLP on edges (activations recomputed).  Upstream gradient of `out`:
 *   mode 0: g_u[e] rows of g_up (E,F);   mode 1: g_u[e] = g_up[src[e]] rows of g_up (N,F).
 * Outputs: g_e (E,F) = g_e_base (may be NULL = 0) + d out/d e adjoint; g_z1 (E,2F); g_h (E,R) accumulated (+=)
 * W1e (2F,F) is the e-part of layer 1 in (out,in) layout; W2d/W2g (F,F) in (out,in) layout; Wh (F,R). */
int m3g_conv_mlp_bwd(const float* P, int ldp, int po, const int32_t* src, const int32_t* dst, const float* e,
                     const float* h, const float* W1eT, const float* W2dT, const float* b2d, const float* W2gT,
                     const float* b2g, const float* WhT, const float* W1e, const float* W2d, const float* W2g,
                     const float* Wh, const float* g_up, const float* g_e_base, int64_t E, int F, int R, int mode,
                     float* g_e, float* g_z1, float* g_h, void* stream);
/* g_P (N, ldp)[po..po+2F) = sum_{out(i)} g_z1, [po+2F..po+4F) = sum_{in(i)} g_z1 */
int m3g_conv_gather_gz(const float* g_z1, const int32_t* edge_ptr, const int32_t* in_ptr, const int32_t* in_perm,
                       int64_t N, int F, int ldp, int po, float* g_P, void* stream);
/* out (n,K) = base (n,K or NULL) + g (n,M) · W (M,K)   — adjoint of m3g_linear_fwd w.r.t. its input */
int m3g_linear_bwd_input(const float* g, const float* W, const float* base, int64_t n, int K, int M, float* out,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core path of the same gated MLP for F = 64 (csrc/conv_tc.cu): tcgen05.mma kind::tf32 with the 3xTF32
 * split (passes = 3; passes = 1 is plain TF32), accumulators in TMEM, weight images resident in shared memory.
 * m3g_tc_pack_b writes the hi / lo SWIZZLE_128B operand images of a (rows x cols) row-major weight matrix
 * (rows % 8 == 0, cols % 32 == 0; each image rows*cols floats).  wimg = [W1e hi | W1e lo | W2d hi | W2d lo |
 * W2g hi | W2g lo] with W1e (128 x 64) = rows [dense | gate] of the e-part of layer 1, W2d/W2g (64 x 64), all in
 * the reference's (out,in) orientation.  n_sm = number of SMs (persistent grid: two warp groups per CTA ping-pong
 * two 128-edge tiles; R <= 4).
 * ------------------------------------------------------------------------------------------- */
int m3g_tc_pack_b(const float* W, int rows, int cols, float* img_hi, float* img_lo, void* stream);
/* debug: per-phase cycle sums of the backward kernel (HOST buffer of 16 int64; all zero unless the library was built
 * with -DM3G_TC_TIMING); synchronises the stream */
int m3g_debug_tc_timing(int64_t* out16, int reset, void* stream);
/* debug: cycles to issue / to complete n_mma back-to-back tcgen05.mma kind::tf32 (M = 128, K = 8) of width N with A
 * from shared (a_tmem = 0) or tensor memory (1); cycles2 = DEVICE buffer of two int64 */
int m3g_debug_mma_rate(int N, int a_tmem, int n_mma, int64_t* cycles2, void* stream);
/* same for tcgen05.mma.cta_group::2 (M = 256 over a two-CTA cluster, A and B from shared memory): cycles2[0] = issue
 * loop, cycles2[1] = until the multicast commit arrives in the leader CTA. */
int m3g_debug_mma_rate2(int N, int n_mma, int64_t* cycles2, void* stream);
/* dev tool (tools/pipe_rate.py): clock cycles of `iters` x 32 instructions of one kind per thread (8 independent
 * chains) with `threads` threads per SM — kinds in csrc/debug_rate.cu */
int m3g_debug_pipe_rate(int kind, int threads, int iters, int64_t* cycles, void* stream);
/* UMMA plumbing self test on one 128-row tile: out (128 x rows) = A (128 x cols) · W^T, W given as its image;
 * a_tmem = 1 feeds A from tensor memory (tcgen05.st + the [a_tmem] operand form) instead of shared memory */
int m3g_tc_selftest(const float* A, const float* img_hi, const float* img_lo, int rows, int cols, int passes,
                    int a_tmem, float* out, void* stream);
/* mode 0 / 1 as m3g_conv_mlp_fwd; mode 2 = mode 1 with the messages summed per source atom inside the epilogue: y then
 * holds one partial row per (32-row block, source atom) at the block's first row of that atom, for
 * m3g_segment_sum_parts (the other rows of y are not written) */
int m3g_conv_tc_fwd(const float* P, int ldp, int po, const int32_t* src, const int32_t* dst, const float* e,
                    const float* h, const float* wimg, const float* b2d, const float* b2g, const float* WhT, int64_t E,
                    int R, int mode, int passes, int n_sm, float* y, float* save, void* stream);
/* (nn/conv.py:63-89 with nn/core.py:30-62 for m3g_conv_tc_fwd; the two functions below replace its autograd)
 * save (optional): activations for m3g_conv_tc_bwd_saved — m3g_conv_tc_save_floats(E) floats (1 KB per edge:
 * SiLU'(z1) and the layer-2 pre-activations, in a tile-private fragment-major layout).  With them the backward needs
 * neither the forward weights nor P / e: output adjoint -> two 64x64 adjoint GEMM pairs -> g_e, g_z1, g_h (same outputs
 * and conventions as m3g_conv_tc_bwd; src is only read for mode 1, where g_up is indexed by source atom).
 * gh_store != 0: the g_h rows are stored instead of added to the rows already there (a caller that gives every launch
 * its own (E,R) slice and sums them with m3g_sum_slices: no read-modify-write inside the kernel). */
int64_t m3g_conv_tc_save_floats(int64_t E);
int m3g_conv_tc_bwd_saved(const int32_t* src, const float* h, const float* wimgT, const float* WhT, const float* save,
                          const float* g_up, const float* g_e_base, int64_t E, int R, int mode, int passes, int n_sm,
                          float* g_e, float* g_z1, float* g_h, int gh_store, void* stream);

/* adjoint of m3g_conv_tc_fwd (same outputs as m3g_conv_mlp_bwd; forward recomputed on the tensor cores).
 * wimgT = [W2d^T hi|lo, W2g^T hi|lo, W1e_dense^T hi|lo, W1e_gate^T hi|lo]: four 64x64 image pairs from
 * m3g_tc_pack_b of the transposed matrices (W1e_dense = rows 0..63 of W1e, W1e_gate = rows 64..127). R <= 3. */
int m3g_conv_tc_bwd(const float* P, int ldp, int po, const int32_t* src, const int32_t* dst, const float* e,
                    const float* h, const float* wimg, const float* wimgT, const float* b2d, const float* b2g,
                    const float* WhT, const float* g_up, const float* g_e_base, int64_t E, int R, int mode, int passes,
                    int n_sm, float* g_e, float* g_z1, float* g_h, void* stream);

/* ---------------------------------------------------------------------------------------------
 * AtomWiseReadout (nn/readout.py:39-58) + virial (nn/gradient.py:39-62)
 * weights: 3-layer gated MLP F→F→F→1; WT = (in,out) layout, W = (out,in) layout
 * ------------------------------------------------------------------------------------------- */
int m3g_readout_fwd(const float* x, const float* W0dT, const float* b0d, const float* W1dT, const float* b1d,
                    const float* w2d, const float* b2d, const float* W0gT, const float* b0g, const float* W1gT,
                    const float* b1g, const float* w2g, const float* b2g, const float* elemental, float scale,
                    int64_t N, int F, float* atomic, void* stream);
/* scaled_total[b] = sum_{i in [atom_ptr[b], atom_ptr[b+1])} atomic[i]; total = scale * scaled_total */
int m3g_structure_sum(const float* atomic, const int32_t* atom_ptr, int64_t B, float scale, float* scaled_total,
                      float* total, void* stream);
/* g_eps[i] = g_atomic[i] + g_scaled_total[b] + scale*g_total[b] (each may be NULL); g_x (N,F) */
int m3g_readout_bwd(const float* x, const float* W0dT, const float* b0d, const float* W1dT, const float* b1d,
                    const float* w2d, const float* b2d, const float* W0gT, const float* b0g, const float* W1gT,
                    const float* b1g, const float* w2g, const float* b2g, const float* W0d, const float* W1d,
                    const float* W0g, const float* W1g, const float* g_atomic, const float* g_scaled_total,
                    const float* g_total, const int32_t* batch, float scale, int64_t N, int F, float* g_x,
                    void* stream);
/* m3g_readout_fwd and m3g_readout_bwd in ONE launch for a caller that knows the upstream gradient beforehand (the
 * whole-step executor: g_total is an input): writes atomic (N) and g_x (N,F); same arithmetic as the two calls */
int m3g_readout_fwd_bwd(const float* x, const float* W0dT, const float* b0d, const float* W1dT, const float* b1d,
                        const float* w2d, const float* b2d, const float* W0gT, const float* b0g, const float* W1gT,
                        const float* b1g, const float* w2g, const float* b2g, const float* W0d, const float* W1d,
                        const float* W0g, const float* W1g, const float* elemental, const float* g_total,
                        const int32_t* batch, float scale, int64_t N, int F, float* atomic, float* g_x, void* stream);
/* forces = -g_pos; stresses (B,6) = Voigt(sum_i pos_i (x) F_i)/|det lattice| */
int m3g_forces_virial(const float* pos, const float* g_pos, const float* lattice, const int32_t* atom_ptr,
                      int64_t N, int64_t B, float* forces, float* stresses, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Domain decomposition helpers (new; nothing in the reference — nn/gradient.py:26 TODO)
 * ------------------------------------------------------------------------------------------- */
/* out[r] = in[idx[r]] rows of width W floats (halo pack) ; in[idx[r]] += add[r] (reverse halo unpack) */
int m3g_rows_gather(const float* in, const int32_t* idx, int64_t n, int W, float* out, void* stream);
int m3g_rows_scatter_add(const float* add, const int32_t* idx, int64_t n, int W, float* inout, void* stream);
/* halo push over NVLink / NVSwitch peer memory: row idx[r] (or r when idx == NULL) of `in` (rows of W floats; 16-byte
 * aligned rows when W % 4 == 0) is stored at the absolute device address dst_addr[r], typically a slot of a PEER
 * GPU's landing buffer (peer-mapped symmetric memory).  Pack + transfer in one kernel, no collective call; the caller
 * orders it against the consumer with a device-side barrier (domain.py::DomainStep, exchange = "p2p"). */
int m3g_rows_put(const float* in, const int32_t* idx, const int64_t* dst_addr, int64_t n, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-step executor (csrc/step.cu): the launch sequence of  Gradient(Sequential[ScaleLength, AtomRef,
 * DistanceAndAngle, AtomFeaturizer, EdgeFeaturizer, EdgeAdjustor, (ThreeBodyInteration, M3GNetConv) x n_blocks,
 * AtomWiseReadout])  (model/build.py:37-83, nn/gradient.py:25-64) — forward kernels, then the hand-written adjoint
 * kernels in reverse order, then forces + virial — issued from C with one call instead of ~70 Python round trips and
 * an autograd pass.  Default model shape only (F = 64, l_max = n_max = 3, radial degree 3, canonical triplet layout);
 * everything else stays on the per-operator entry points above.  All buffers are caller-owned device memory; the
 * struct itself lives in HOST memory.  The sequence is cut into phases so that a caller can interleave its own work
 * (the halo exchanges of a domain-decomposed cell, csrc/geometry.cu m3g_rows_*):
 *
 *   M3G_PHASE_PROLOGUE              scale, atom reference, geometry (+ cos), embedding, radial basis, e0, G / dG
 *   M3G_PHASE_TB(b), M3G_PHASE_CONV(b)      b = 0 .. n_blocks-1   (x of block b+1 is blocks[b].x_out)
 *   M3G_PHASE_READOUT(n)            energies; adjoint of the readout -> g_x
 *   M3G_PHASE_CONV_BWD(n,b), M3G_PHASE_TB_BWD(n,b)   b = n_blocks-1 .. 0   (the adjoint w.r.t. blocks[b].x_in is
 *                                   left in g_x[cur]; block 0 skips the dead gradient w.r.t. the embedding)
 *   M3G_PHASE_EPILOGUE(n)           e0 / radial / geometry adjoints -> g_pos
 *   M3G_PHASE_FORCES(n)             forces = -g_pos, virial stress
 * ------------------------------------------------------------------------------------------- */
#define M3G_STEP_MAX_BLOCKS 8
#define M3G_PHASE_PROLOGUE 0
#define M3G_PHASE_TB(b) (1 + 2 * (b))
#define M3G_PHASE_CONV(b) (2 + 2 * (b))
#define M3G_PHASE_READOUT(n) (1 + 2 * (n))
#define M3G_PHASE_CONV_BWD(n, b) (2 + 2 * (n) + 2 * ((n)-1 - (b)))
#define M3G_PHASE_TB_BWD(n, b) (3 + 2 * (n) + 2 * ((n)-1 - (b)))
#define M3G_PHASE_EPILOGUE(n) (2 + 4 * (n))
#define M3G_PHASE_FORCES(n) (3 + 4 * (n))

typedef struct M3GStepBlock {
  /* ThreeBodyInteration weights (nn/interaction.py:138-185): Ws (9,64), bs (9), WdT / WgT (9,64), tb_consts (20) */
  const float *Ws, *bs, *WdT, *WgT, *tb_consts;
  /* M3GNetConv weights (nn/conv.py:25-61): per-atom projection WpT (64,512) / bp (512) / Wp (512,64); per gated MLP
   * the tcgen05 weight images (m3g_tc_pack_b), second-layer biases and the (3,64) radial weights */
  const float *WpT, *bp, *Wp;
  const float *e_wimg, *e_wimgT, *e_b2d, *e_b2g, *e_WhT;
  const float *n_wimg, *n_wimgT, *n_b2d, *n_b2g, *n_WhT;
  /* activations: G / dG (E,9) radial tables (blocks with equal constants share them; radial_owner = 1 computes),
   * sig (N,9), red (E,9), x_in (N,64), e_in (E,64), e_tb (E,64), e_out (E,64), x_out (N,64), saved activations of
   * the two gated MLPs (ceil(E/128)*128*256 floats each) */
  float *G, *dG;
  int radial_owner;
  float *sig, *red;
  const float *x_in, *e_in;
  float *e_tb, *e_out, *x_out, *save_e, *save_n;
} M3GStepBlock;

typedef struct M3GStepDesc {
  int64_t N, E, T, B;
  int n_blocks, n_sm, passes, max_members;
  int64_t n_members; /* member bonds (bonds inside the three-body cutoff) */
  float length_scale, energy_scale, r3;
  int num_types;
  /* plan (int32 CSR views, see "Conventions") */
  const int32_t *batch, *src, *dst, *shift, *types, *edge_ptr, *in_ptr, *in_perm, *tri_ptr, *atom_ptr, *member_edges;
  const int64_t* tri_index; /* (2,T) caller-order list for the triplet_angles output, or NULL */
  /* inputs */
  const float *pos, *lattice;
  /* layer weights outside the blocks */
  const float *embed_W, *atomref_table, *radial_consts, *adjust_Wt;
  const float *ro_W0dT, *ro_b0d, *ro_W1dT, *ro_b1d, *ro_w2d, *ro_b2d, *ro_W0gT, *ro_b0g, *ro_W1gT, *ro_b1g, *ro_w2g,
      *ro_b2g, *ro_W0d, *ro_W1d, *ro_W0g, *ro_W1g;
  const float* g_total; /* (B) upstream gradient of the structure energies (ones; owned / ghost weights for a domain) */
  /* outputs / step-level activations */
  float *scaled_pos, *scaled_lattice, *elemental, *vec4, *dist, *cos_t, *x0, *h, *e0;
  float *atomic, *scaled_total, *total, *forces, *stresses;
  /* scratch */
  float *P, *msg;                 /* (N,512), (E,64) */
  float *g_x[2];                  /* (N,64) ping-pong; g_x[cur] holds the running adjoint */
  float *g_e[2], *ge2;            /* (E,64) */
  float *gz_edge, *gz_node, *gP;  /* (E,128) x 2, (N,512) */
  float *g_h, *g_hs, *g_sig_e, *g_vec4, *g_dist, *g_pos; /* (E,3), (2 n_blocks + 1, E, 3) per-launch slices, (E,9), (E,4), (E), (N,3) */
  int cur_x, cur_e;               /* state carried between phases (updated by m3g_step_run) */
  int have_g_e;                   /* 0 until the first conv adjoint has produced g_e */
  int msg_reduce;                 /* 1: node MLP sums its messages per atom in-kernel (m3g_conv_tc_fwd mode 2) */
  int tb_split;                   /* 1: three-body forward as m3g_tb_mom_red + m3g_tb_edge_update */
  int tb_bwd_split;               /* 1: three-body backward as m3g_tb_mlp_adj (q overwrites red) + m3g_tb_mom_bwd_q */
  int fuse_e0;                    /* 1 (needs tb_split): e0 is formed inside block 0's m3g_tb_edge_update_h, never stored */
  M3GStepBlock blocks[M3G_STEP_MAX_BLOCKS];
} M3GStepDesc;

/* run phases first .. last (inclusive) on `stream`; returns the first non-zero status */
int m3g_step_run(M3GStepDesc* desc, int first_phase, int last_phase, void* stream);
/* sizeof(M3GStepDesc) as compiled (binding self-check) */
int64_t m3g_step_desc_size(void);

#ifdef __cplusplus
}
#endif
#endif /* M3GNET_B200_H */
