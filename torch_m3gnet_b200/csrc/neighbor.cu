// PBC radius-graph neighbour list (replaces pymatgen's Structure.get_all_neighbors boundary,
// data/material_graph.py:168-193) and triplet enumeration (compute_threebody, :196-254), batched over
// structures.  Float64 accept test with contraction disabled, bit-identical to oracle/m3gnet_oracle.py:
//   shift = (s0*a0 + s1*a1) + s2*a2 ; v = (cart[j] + shift) - cart[i] ; d2 = (vx*vx + vy*vy) + vz*vz
//   accept  d2 < r*r + 1e-8  and not (i == j and sqrt(d2) <= 1e-8)
// One warp per centre atom; lanes sweep the candidate atoms j of the same structure in ascending order,
// each lane walks its image range lexicographically, and a warp prefix sum keeps the emitted edges ordered
// by (j, s0, s1, s2).  Candidate atoms come either from the whole structure (small cells) or from the
// 27-bin neighbourhood of a cell list (structures with >= 3 bins of width >= r along every axis).
#include "common.cuh"

namespace m3g {

struct Cell {
  double a[9];    // lattice rows
  double inv[9];  // inverse (columns b_k: frac = cart · inv)
  double reach[3];
};

__device__ __forceinline__ void load_cell(const double* __restrict__ lattice, int b, double cutoff, Cell& c) {
  const double* Lm = lattice + (int64_t)b * 9;
#pragma unroll
  for (int k = 0; k < 9; ++k) c.a[k] = Lm[k];
  const double *a0 = c.a, *a1 = c.a + 3, *a2 = c.a + 6;
  double c0[3] = {a1[1] * a2[2] - a1[2] * a2[1], a1[2] * a2[0] - a1[0] * a2[2], a1[0] * a2[1] - a1[1] * a2[0]};
  double c1[3] = {a2[1] * a0[2] - a2[2] * a0[1], a2[2] * a0[0] - a2[0] * a0[2], a2[0] * a0[1] - a2[1] * a0[0]};
  double c2[3] = {a0[1] * a1[2] - a0[2] * a1[1], a0[2] * a1[0] - a0[0] * a1[2], a0[0] * a1[1] - a0[1] * a1[0]};
  double det = a0[0] * c0[0] + a0[1] * c0[1] + a0[2] * c0[2];
  // inv[r][k] = (c_k)[r] / det  so that frac_k = sum_r cart_r inv[r][k]
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c.inv[r * 3 + 0] = c0[r] / det;
    c.inv[r * 3 + 1] = c1[r] / det;
    c.inv[r * 3 + 2] = c2[r] / det;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double bk2 = c.inv[0 + k] * c.inv[0 + k] + c.inv[3 + k] * c.inv[3 + k] + c.inv[6 + k] * c.inv[6 + k];
    // |s_k + df_k| <= (r + margin) * |b_k|   (1/|b_k| is the spacing of lattice planes along axis k)
    c.reach[k] = (cutoff + 1e-6) * sqrt(bk2) + 1e-9;
  }
}

__device__ __forceinline__ void frac_of(const Cell& c, const double* p, double* f) {
#pragma unroll
  for (int k = 0; k < 3; ++k) f[k] = p[0] * c.inv[0 + k] + p[1] * c.inv[3 + k] + p[2] * c.inv[6 + k];
}

__device__ __forceinline__ double dist2_exact(const Cell& c, const double* pi, const double* pj, int s0, int s1,
                                              int s2) {
  double d2 = 0.0;
  double v[3];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    double sh = __dadd_rn(__dadd_rn(__dmul_rn((double)s0, c.a[0 + x]), __dmul_rn((double)s1, c.a[3 + x])),
                          __dmul_rn((double)s2, c.a[6 + x]));
    v[x] = __dsub_rn(__dadd_rn(pj[x], sh), pi[x]);
  }
  d2 = __dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2]));
  return d2;
}

__device__ __forceinline__ int find_structure(const int32_t* __restrict__ atom_ptr, int B, int i) {
  int lo = 0, hi = B;  // largest b with atom_ptr[b] <= i
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (atom_ptr[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- cell list (structures whose cell holds >= 3 bins of perpendicular width >= cutoff along every axis) ----
// Bins only restrict the CANDIDATE atoms j; the accept test and the emission order are those of the brute-force
// sweep, so the edge set and its order are identical.  bins (B,3): bin counts per axis, 0 = no cell list for that
// structure; bin_base (B+1): offset of a structure's bins in the global bin arrays.
constexpr int NBR_WARPS = 4;
constexpr int CAND_MAX = 512;

__device__ __forceinline__ int bin_coord(double f, int nb) {
  double w = f - floor(f);
  int k = (int)(w * (double)nb);
  return k >= nb ? nb - 1 : (k < 0 ? 0 : k);
}

__global__ void bin_assign_kernel(const double* __restrict__ lattice, const double* __restrict__ cart,
                                  const int32_t* __restrict__ atom_ptr, int B, int64_t N, double cutoff,
                                  const int32_t* __restrict__ bins, const int32_t* __restrict__ bin_base,
                                  int32_t* __restrict__ atom_bin, int32_t* __restrict__ bin_count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int b = find_structure(atom_ptr, B, (int)i);
  const int nb0 = bins[b * 3], nb1 = bins[b * 3 + 1], nb2 = bins[b * 3 + 2];
  if (nb0 <= 0) { atom_bin[i] = -1; return; }
  Cell c;
  load_cell(lattice, b, cutoff, c);
  double p[3] = {cart[i * 3 + 0], cart[i * 3 + 1], cart[i * 3 + 2]};
  double f[3];
  frac_of(c, p, f);
  int g = bin_base[b] + (bin_coord(f[0], nb0) * nb1 + bin_coord(f[1], nb1)) * nb2 + bin_coord(f[2], nb2);
  atom_bin[i] = g;
  atomicAdd(&bin_count[g], 1);
}

__global__ void bin_fill_kernel(const int32_t* __restrict__ atom_bin, const int32_t* __restrict__ bin_ptr, int64_t N,
                                int32_t* __restrict__ bin_cursor, int32_t* __restrict__ bin_atoms) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int g = atom_bin[i];
  if (g < 0) return;
  bin_atoms[bin_ptr[g] + atomicAdd(&bin_cursor[g], 1)] = (int)i;  // order inside a bin is irrelevant (sorted later)
}

// ascending bitonic sort of s[0..P) by one warp (P a power of two); j is a power of two: shifts, no divisions
__device__ __forceinline__ void warp_sort(int* s, int P, int lane) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1, lj = 31 - __clz(k >> 1); j > 0; j >>= 1, --lj) {
      for (int t = lane; t < (P >> 1); t += 32) {
        int i = ((t >> lj) << (lj + 1)) + (t & (j - 1));
        int x = i + j;
        bool up = (i & k) == 0;
        int a = s[i], bv = s[x];
        if ((a > bv) == up) { s[i] = bv; s[x] = a; }
      }
      __syncwarp();
    }
  }
}

// FILL = false: edge_count[i]; FILL = true: ordered emission at edge_ptr[i]
template <bool FILL>
__global__ void __launch_bounds__(32 * NBR_WARPS)
nbr_kernel(const double* __restrict__ lattice, const double* __restrict__ cart, const int32_t* __restrict__ atom_ptr,
           int B, int64_t N, double cutoff, float r3_f32, const int32_t* __restrict__ bins,
           const int32_t* __restrict__ bin_base, const int32_t* __restrict__ bin_ptr,
           const int32_t* __restrict__ bin_atoms, const int32_t* __restrict__ edge_ptr,
           int32_t* __restrict__ edge_count, int64_t* __restrict__ edge_index, int64_t E,
           int32_t* __restrict__ edge_shift, float* __restrict__ edge_dist, int32_t* __restrict__ member) {
  __shared__ int cand_s[NBR_WARPS][CAND_MAX];
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= N) return;
  int* cand = cand_s[threadIdx.x >> 5];
  int b = find_structure(atom_ptr, B, (int)i);
  Cell c;
  load_cell(lattice, b, cutoff, c);
  const double r2 = __dadd_rn(__dmul_rn(cutoff, cutoff), 1e-8);
  double pi[3] = {cart[i * 3 + 0], cart[i * 3 + 1], cart[i * 3 + 2]};
  double fi[3];
  frac_of(c, pi, fi);
  int a_beg = atom_ptr[b], a_end = atom_ptr[b + 1];
  // candidate atoms: the whole structure, or the sorted content of the 27 bins around atom i
  bool use_cells = false;
  int n_c = a_end - a_beg;
  if (bins != nullptr && bins[b * 3] > 0) {
    const int nb[3] = {bins[b * 3], bins[b * 3 + 1], bins[b * 3 + 2]};
    const int bc[3] = {bin_coord(fi[0], nb[0]), bin_coord(fi[1], nb[1]), bin_coord(fi[2], nb[2])};
    // lanes 0..26 look up one bin each; a warp prefix sum places the bins' atoms back to back
    int p0 = 0, cnt = 0;
    if (lane < 27) {
      int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
      int gx = (bc[0] + dx + nb[0]) % nb[0], gy = (bc[1] + dy + nb[1]) % nb[1], gz = (bc[2] + dz + nb[2]) % nb[2];
      int g = bin_base[b] + (gx * nb[1] + gy) * nb[2] + gz;
      p0 = bin_ptr[g];
      cnt = bin_ptr[g + 1] - p0;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += y;
    }
    const int n = __shfl_sync(FULL, incl, 31);
    const bool overflow = n > CAND_MAX;
    if (!overflow) {
      for (int q = 0; q < 27; ++q) {
        const int off = __shfl_sync(FULL, incl - cnt, q), c = __shfl_sync(FULL, cnt, q), b0 = __shfl_sync(FULL, p0, q);
        for (int p = lane; p < c; p += 32) cand[off + p] = bin_atoms[b0 + p];
      }
    }
    if (!overflow) {
      __syncwarp();
      n_c = n;
      if (FILL) {
        // Only the emission order needs ascending candidates (counting does not), and only candidates with at least
        // one accepted image are emitted: keep those (compacted in place, reads of a chunk precede its writes), then
        // sort the few that remain instead of the whole 27-bin neighbourhood.
        int kept = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
          const bool have = (j0 + lane) < n;
          const int j = have ? cand[j0 + lane] : 0;
          bool any = false;
          if (have) {
            const double pj[3] = {cart[(int64_t)j * 3 + 0], cart[(int64_t)j * 3 + 1], cart[(int64_t)j * 3 + 2]};
            double fj[3];
            frac_of(c, pj, fj);
            int lo[3], hi[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              double df = fj[k] - fi[k];
              lo[k] = (int)ceil(-c.reach[k] - df);
              hi[k] = (int)floor(c.reach[k] - df);
            }
            for (int s0 = lo[0]; s0 <= hi[0] && !any; ++s0)
              for (int s1 = lo[1]; s1 <= hi[1] && !any; ++s1)
                for (int s2 = lo[2]; s2 <= hi[2] && !any; ++s2) {
                  double d2 = dist2_exact(c, pi, pj, s0, s1, s2);
                  bool ok = d2 < r2;
                  if (ok && j == (int)i && __dsqrt_rn(d2) <= 1e-8) ok = false;
                  any = ok;
                }
          }
          const unsigned bal = __ballot_sync(FULL, any);
          __syncwarp();
          if (any) cand[kept + __popc(bal & ((1u << lane) - 1u))] = j;
          kept += __popc(bal);
        }
        int Pk = 32;
        while (Pk < kept) Pk <<= 1;
        for (int t = kept + lane; t < Pk; t += 32) cand[t] = 0x7fffffff;
        __syncwarp();
        warp_sort(cand, Pk, lane);
        n_c = kept;
      }
      use_cells = true;
    }
  }
  int64_t base = FILL ? edge_ptr[i] : 0;
  int total = 0;
  for (int j0 = 0; j0 < n_c; j0 += 32) {
    bool have = (j0 + lane) < n_c;
    int j = have ? (use_cells ? cand[j0 + lane] : a_beg + j0 + lane) : a_end;
    double pj[3] = {0, 0, 0};
    int lo[3] = {0, 0, 0}, hi[3] = {-1, -1, -1};
    if (have) {
      pj[0] = cart[(int64_t)j * 3 + 0]; pj[1] = cart[(int64_t)j * 3 + 1]; pj[2] = cart[(int64_t)j * 3 + 2];
      double fj[3];
      frac_of(c, pj, fj);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double df = fj[k] - fi[k];
        lo[k] = (int)ceil(-c.reach[k] - df);
        hi[k] = (int)floor(c.reach[k] - df);
      }
    }
    // pass 1: count this lane's accepted images
    int cnt = 0;
    for (int s0 = lo[0]; s0 <= hi[0]; ++s0)
      for (int s1 = lo[1]; s1 <= hi[1]; ++s1)
        for (int s2 = lo[2]; s2 <= hi[2]; ++s2) {
          double d2 = dist2_exact(c, pi, pj, s0, s1, s2);
          bool ok = d2 < r2;
          if (ok && j == (int)i && __dsqrt_rn(d2) <= 1e-8) ok = false;
          cnt += ok ? 1 : 0;
        }
    // exclusive prefix over lanes
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += y;
    }
    int chunk_total = __shfl_sync(FULL, incl, 31);
    if (FILL && cnt > 0) {
      int64_t w = base + total + (incl - cnt);
      for (int s0 = lo[0]; s0 <= hi[0]; ++s0)
        for (int s1 = lo[1]; s1 <= hi[1]; ++s1)
          for (int s2 = lo[2]; s2 <= hi[2]; ++s2) {
            double d2 = dist2_exact(c, pi, pj, s0, s1, s2);
            bool ok = d2 < r2;
            double d = __dsqrt_rn(d2);
            if (ok && j == (int)i && d <= 1e-8) ok = false;
            if (ok) {
              edge_index[w] = i;
              edge_index[E + w] = j;
              edge_shift[w * 3 + 0] = s0;
              edge_shift[w * 3 + 1] = s1;
              edge_shift[w * 3 + 2] = s2;
              float df = __double2float_rn(d);
              edge_dist[w] = df;
              member[w] = (df <= r3_f32) ? 1 : 0;
              ++w;
            }
          }
    }
    total += chunk_total;
  }
  if (!FILL && lane == 0) edge_count[i] = total;
}

// ---- Verlet (skin) list: the bonds of a step out of a candidate list built once with cutoff + skin ----
// The candidate list is an ordinary neighbour list (same kernels, larger radius): per atom ordered by (j, s0, s1, s2),
// images relative to the unwrapped coordinates.  While no atom has moved further than skin/2 from the coordinates the
// candidates were built on (and the lattice is unchanged), every pair inside the cutoff is among the candidates, so
// filtering them with the builder's own float64 accept test gives the builder's edge set in the builder's order.

// max over atoms of |cart - ref|^2, as the bit pattern of a non-negative float (atomicMax on int keeps the order)
__global__ void verlet_displacement_kernel(const double* __restrict__ cart, const double* __restrict__ ref, int64_t N,
                                           int32_t* __restrict__ max_d2_bits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float d2 = 0.f;
  if (i < N) {
    double dx = cart[i * 3] - ref[i * 3], dy = cart[i * 3 + 1] - ref[i * 3 + 1], dz = cart[i * 3 + 2] - ref[i * 3 + 2];
    d2 = __double2float_ru(dx * dx + dy * dy + dz * dz);  // rounded up: the trigger may only fire early
    if (!(d2 >= 0.f)) d2 = __int_as_float(0x7f800000);    // NaN coordinates force a rebuild (which then fails loudly)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(FULL, d2, o));
  if ((threadIdx.x & 31) == 0 && d2 > 0.f) atomicMax(max_d2_bits, __float_as_int(d2));
}

template <bool FILL>
__global__ void __launch_bounds__(256)
verlet_kernel(const double* __restrict__ lattice, const double* __restrict__ cart, const int32_t* __restrict__ atom_ptr,
              int B, int64_t N, double cutoff, float r3_f32, const int32_t* __restrict__ cand_ptr,
              const int32_t* __restrict__ cand_j, const int32_t* __restrict__ cand_shift,
              const int32_t* __restrict__ edge_ptr, int32_t* __restrict__ edge_count,
              int64_t* __restrict__ edge_index, int64_t E, int32_t* __restrict__ edge_shift,
              float* __restrict__ edge_dist, int32_t* __restrict__ member) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= N) return;
  int b = find_structure(atom_ptr, B, (int)i);
  Cell c;
  const double* Lm = lattice + (int64_t)b * 9;
#pragma unroll
  for (int k = 0; k < 9; ++k) c.a[k] = Lm[k];
  const double r2 = __dadd_rn(__dmul_rn(cutoff, cutoff), 1e-8);
  const double pi[3] = {cart[i * 3 + 0], cart[i * 3 + 1], cart[i * 3 + 2]};
  const int c_beg = cand_ptr[i], c_end = cand_ptr[i + 1];
  int64_t w = FILL ? edge_ptr[i] : 0;
  int total = 0;
  for (int p0 = c_beg; p0 < c_end; p0 += 32) {
    int p = p0 + lane;
    bool ok = false;
    int j = 0, s0 = 0, s1 = 0, s2 = 0;
    double d = 0.0;
    if (p < c_end) {
      j = cand_j[p];
      s0 = cand_shift[(int64_t)p * 3]; s1 = cand_shift[(int64_t)p * 3 + 1]; s2 = cand_shift[(int64_t)p * 3 + 2];
      const double pj[3] = {cart[(int64_t)j * 3 + 0], cart[(int64_t)j * 3 + 1], cart[(int64_t)j * 3 + 2]};
      double d2 = dist2_exact(c, pi, pj, s0, s1, s2);
      ok = d2 < r2;
      d = __dsqrt_rn(d2);
      if (ok && j == (int)i && d <= 1e-8) ok = false;
    }
    unsigned mask = __ballot_sync(FULL, ok);
    if (FILL && ok) {
      int64_t q = w + __popc(mask & ((1u << lane) - 1));
      edge_index[q] = i;
      edge_index[E + q] = j;
      edge_shift[q * 3 + 0] = s0;
      edge_shift[q * 3 + 1] = s1;
      edge_shift[q * 3 + 2] = s2;
      float df = __double2float_rn(d);
      edge_dist[q] = df;
      member[q] = (df <= r3_f32) ? 1 : 0;
    }
    w += __popc(mask);
    total += __popc(mask);
  }
  if (!FILL && lane == 0) edge_count[i] = total;
}

// one warp per atom: member degree, per-edge triplet counts, compacted member list (ascending edge id)
// used_count / stats (optional): bonds of atom i that head at least one triplet (n3 if n3 >= 2 else 0), and the totals
// the host needs to size the next buffers — stats[0] += n3 (n3 - 1) (= T), stats[1] = max n3, stats[2] += used_count —
// so that ONE read-back replaces three (integer atomics: order-independent)
__global__ void triplet_count_kernel(const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ member,
                                     int64_t N, int64_t* __restrict__ num_triplet_i,
                                     int32_t* __restrict__ num_triplet_ij, int32_t* __restrict__ tri_count,
                                     int32_t* __restrict__ member_list, int32_t* __restrict__ used_count,
                                     unsigned long long* __restrict__ stats) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= N) return;
  int b = edge_ptr[i], en = edge_ptr[i + 1];
  int n3 = 0;
  for (int e0 = b; e0 < en; e0 += 32) {
    int e = e0 + lane;
    bool m = (e < en) && member[e] != 0;
    unsigned mask = __ballot_sync(FULL, m);
    if (m) member_list[b + n3 + __popc(mask & ((1u << lane) - 1))] = e;
    n3 += __popc(mask);
  }
  for (int e = b + lane; e < en; e += 32) {
    int v = member[e] ? (n3 - 1) : 0;
    num_triplet_ij[e] = v;
    tri_count[e] = v;
  }
  if (lane == 0) {
    num_triplet_i[i] = (int64_t)n3 * (n3 - 1);
    if (used_count) used_count[i] = (n3 >= 2) ? n3 : 0;
    if (stats && n3 >= 2) {
      atomicAdd(stats, (unsigned long long)n3 * (unsigned long long)(n3 - 1));
      atomicMax(stats + 1, (unsigned long long)n3);
      atomicAdd(stats + 2, (unsigned long long)n3);
    }
  }
}

// one warp per atom: the n3(n3-1) ordered pairs of member edges, ordered by (e1, e2) — the exact order of
// the reference's triple loop (data/material_graph.py:239-248)
__global__ void triplet_fill_kernel(const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ tri_ptr,
                                    const int32_t* __restrict__ tri_count, const int32_t* __restrict__ member_list,
                                    int64_t N, int64_t T, int32_t* __restrict__ tri_e2,
                                    int64_t* __restrict__ triplet_index, const int32_t* __restrict__ used_ptr,
                                    int32_t* __restrict__ member_edges) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= N) return;
  int b = edge_ptr[i], en = edge_ptr[i + 1];
  if (en <= b) return;
  // n3 - 1 is stored on every member edge; find the first member through the compacted list
  int first = member_list[b];
  // member_list[b] is only valid if the atom has at least one member edge: check through tri_count sum
  int n3m1 = -1;
  for (int e = b + lane; e < en; e += 32) {
    int v = tri_count[e];
    if (v > 0) n3m1 = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n3m1 = max(n3m1, __shfl_xor_sync(FULL, n3m1, o));
  if (n3m1 <= 0) return;  // 0 or 1 member edges: no triplets
  int n3 = n3m1 + 1;
  // the global list of bonds that head a triplet (ascending: atoms ascending, members ascending per atom)
  if (member_edges)
    for (int k = lane; k < n3; k += 32) member_edges[used_ptr[i] + k] = member_list[b + k];
  int64_t base = tri_ptr[first];
  int total = n3 * n3m1;
  for (int t = lane; t < total; t += 32) {
    int a = t / n3m1;
    int bb = t - a * n3m1;
    int bidx = bb + (bb >= a ? 1 : 0);
    int e1 = member_list[b + a], e2 = member_list[b + bidx];
    tri_e2[base + t] = e2;
    if (triplet_index) {
      triplet_index[base + t] = e1;
      triplet_index[T + base + t] = e2;
    }
  }
}

}  // namespace m3g

using namespace m3g;

extern "C" {

int m3g_nbr_bin_count(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                      double cutoff, const int32_t* bins, const int32_t* bin_base, int32_t* atom_bin,
                      int32_t* bin_count, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(lattice && cart && atom_ptr && bins && bin_base && atom_bin && bin_count && B > 0,
              "m3g_nbr_bin_count: bad argument");
  bin_assign_kernel<<<blocks_for(N, 256), 256, 0, as_stream(stream)>>>(lattice, cart, atom_ptr, (int)B, N, cutoff, bins,
                                                                       bin_base, atom_bin, bin_count);
  M3G_LAUNCH_CHECK("m3g_nbr_bin_count");
  return M3G_OK;
}

int m3g_nbr_bin_fill(const int32_t* atom_bin, const int32_t* bin_ptr, int64_t N, int32_t* bin_cursor,
                     int32_t* bin_atoms, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(atom_bin && bin_ptr && bin_cursor && bin_atoms, "m3g_nbr_bin_fill: null pointer");
  bin_fill_kernel<<<blocks_for(N, 256), 256, 0, as_stream(stream)>>>(atom_bin, bin_ptr, N, bin_cursor, bin_atoms);
  M3G_LAUNCH_CHECK("m3g_nbr_bin_fill");
  return M3G_OK;
}

int m3g_nbr_count(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                  double cutoff, const int32_t* bins, const int32_t* bin_base, const int32_t* bin_ptr,
                  const int32_t* bin_atoms, int32_t* edge_count, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(lattice && cart && atom_ptr && edge_count && B > 0, "m3g_nbr_count: bad argument");
  M3G_REQUIRE(!bins || (bin_base && bin_ptr && bin_atoms), "m3g_nbr_count: incomplete cell list");
  nbr_kernel<false><<<blocks_for(N * 32, 32 * NBR_WARPS), 32 * NBR_WARPS, 0, as_stream(stream)>>>(
      lattice, cart, atom_ptr, (int)B, N, cutoff, 0.0f, bins, bin_base, bin_ptr, bin_atoms, nullptr, edge_count,
      nullptr, 0, nullptr, nullptr, nullptr);
  M3G_LAUNCH_CHECK("m3g_nbr_count");
  return M3G_OK;
}

int m3g_nbr_fill(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                 double cutoff, double threebody_cutoff, const int32_t* bins, const int32_t* bin_base,
                 const int32_t* bin_ptr, const int32_t* bin_atoms, const int32_t* edge_ptr, int64_t E,
                 int64_t* edge_index, int32_t* edge_shift, float* edge_dist, int32_t* member, void* stream) {
  if (N == 0 || E == 0) return M3G_OK;
  M3G_REQUIRE(lattice && cart && atom_ptr && edge_ptr && edge_index && edge_shift && edge_dist && member && B > 0,
              "m3g_nbr_fill: bad argument");
  M3G_REQUIRE(!bins || (bin_base && bin_ptr && bin_atoms), "m3g_nbr_fill: incomplete cell list");
  nbr_kernel<true><<<blocks_for(N * 32, 32 * NBR_WARPS), 32 * NBR_WARPS, 0, as_stream(stream)>>>(
      lattice, cart, atom_ptr, (int)B, N, cutoff, (float)threebody_cutoff, bins, bin_base, bin_ptr, bin_atoms, edge_ptr,
      nullptr, edge_index, E, edge_shift, edge_dist, member);
  M3G_LAUNCH_CHECK("m3g_nbr_fill");
  return M3G_OK;
}

int m3g_verlet_displacement(const double* cart, const double* ref_cart, int64_t N, float* max_d2, void* stream) {
  M3G_REQUIRE(max_d2, "m3g_verlet_displacement: null pointer");
  if (cudaMemsetAsync(max_d2, 0, sizeof(float), as_stream(stream)) != cudaSuccess) {
    m3g::set_error("m3g_verlet_displacement: memset failed");
    return M3G_ERR_CUDA;
  }
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(cart && ref_cart, "m3g_verlet_displacement: null pointer");
  verlet_displacement_kernel<<<blocks_for(N, 256), 256, 0, as_stream(stream)>>>(cart, ref_cart, N,
                                                                                reinterpret_cast<int32_t*>(max_d2));
  M3G_LAUNCH_CHECK("m3g_verlet_displacement");
  return M3G_OK;
}

int m3g_verlet_count(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                     double cutoff, const int32_t* cand_ptr, const int32_t* cand_j, const int32_t* cand_shift,
                     int32_t* edge_count, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(lattice && cart && atom_ptr && cand_ptr && cand_j && cand_shift && edge_count && B > 0,
              "m3g_verlet_count: bad argument");
  verlet_kernel<false><<<blocks_for(N * 32, 256), 256, 0, as_stream(stream)>>>(
      lattice, cart, atom_ptr, (int)B, N, cutoff, 0.0f, cand_ptr, cand_j, cand_shift, nullptr, edge_count, nullptr, 0,
      nullptr, nullptr, nullptr);
  M3G_LAUNCH_CHECK("m3g_verlet_count");
  return M3G_OK;
}

int m3g_verlet_fill(const double* lattice, const double* cart, const int32_t* atom_ptr, int64_t B, int64_t N,
                    double cutoff, double threebody_cutoff, const int32_t* cand_ptr, const int32_t* cand_j,
                    const int32_t* cand_shift, const int32_t* edge_ptr, int64_t E, int64_t* edge_index,
                    int32_t* edge_shift, float* edge_dist, int32_t* member, void* stream) {
  if (N == 0 || E == 0) return M3G_OK;
  M3G_REQUIRE(lattice && cart && atom_ptr && cand_ptr && cand_j && cand_shift && edge_ptr && edge_index &&
                  edge_shift && edge_dist && member && B > 0,
              "m3g_verlet_fill: bad argument");
  verlet_kernel<true><<<blocks_for(N * 32, 256), 256, 0, as_stream(stream)>>>(
      lattice, cart, atom_ptr, (int)B, N, cutoff, (float)threebody_cutoff, cand_ptr, cand_j, cand_shift, edge_ptr,
      nullptr, edge_index, E, edge_shift, edge_dist, member);
  M3G_LAUNCH_CHECK("m3g_verlet_fill");
  return M3G_OK;
}

int m3g_triplet_count(const int32_t* edge_ptr, const int32_t* member, int64_t N, int64_t E, int64_t* num_triplet_i,
                      int32_t* num_triplet_ij, int32_t* tri_count, int32_t* member_list, int32_t* used_count,
                      int64_t* stats, void* stream) {
  (void)E;
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(edge_ptr && member && num_triplet_i && num_triplet_ij && tri_count && member_list,
              "m3g_triplet_count: null pointer");
  triplet_count_kernel<<<blocks_for(N * 32, 256), 256, 0, as_stream(stream)>>>(
      edge_ptr, member, N, num_triplet_i, num_triplet_ij, tri_count, member_list, used_count,
      reinterpret_cast<unsigned long long*>(stats));
  M3G_LAUNCH_CHECK("m3g_triplet_count");
  return M3G_OK;
}

int m3g_triplet_fill(const int32_t* edge_ptr, const int32_t* tri_ptr, const int32_t* tri_count,
                     const int32_t* member_list, int64_t N, int64_t T, int32_t* tri_e2, int64_t* triplet_index,
                     const int32_t* used_ptr, int32_t* member_edges, void* stream) {
  if (N == 0 || T == 0) return M3G_OK;
  M3G_REQUIRE(edge_ptr && tri_ptr && tri_count && member_list && tri_e2, "m3g_triplet_fill: null pointer");
  M3G_REQUIRE(!member_edges || used_ptr, "m3g_triplet_fill: member_edges needs used_ptr");
  triplet_fill_kernel<<<blocks_for(N * 32, 256), 256, 0, as_stream(stream)>>>(
      edge_ptr, tri_ptr, tri_count, member_list, N, T, tri_e2, triplet_index, used_ptr, member_edges);
  M3G_LAUNCH_CHECK("m3g_triplet_fill");
  return M3G_OK;
}

}  // extern "C"
