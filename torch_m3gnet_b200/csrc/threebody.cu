// ThreeBodyInteration (nn/interaction.py:187-223) forward and hand-written adjoint.
//
// Data flow per block (D = L*R, d = l*R + n):
//   sig  (N,D) = sigmoid(x Ws^T + bs)                                     [tb_sigma_fwd]
//   bas  (E,D) = chi_d(r_e) * fc(r_e) * sig[dst(e)][d]                    [tb_edge_basis_fwd]
//   red  (E,D) = fc(r_e1) * sum_{e2 in tri(e1)} Y_l(cos(e1,e2)) bas[e2]   [tb_reduce_fwd, fused with]
//   e_out      = e_in + SiLU(red WdT) * sigmoid(red WgT)                  [the 1-layer gated MLP]
// The triplet sum is a segmented reduction over the per-bond CSR row: a sub-warp group of G lanes owns one
// bond, lane g takes columns g, g+G, ... in ascending order, partial sums are combined by a butterfly.
// All per-triplet quantities are recomputed from per-edge data (vec4 = (v, r), bas), so HBM traffic is
// 4 B/triplet (the column index) + per-edge rows that stay L1/L2 resident inside one atom's shell.
#include "common.cuh"

namespace m3g {

__device__ __constant__ float kYpref[9] = {0.28209479177387814f, 0.48860251190291992f, 0.63078313050504009f, 0.7463526651802308f, 0.84628437532163447f, 0.9356025796273888f, 1.0171072362820548f, 1.0925484305920792f, 1.1631066229203195f};

// j_l(x), l < L, by the reference's upward recurrence incl. its small-x branches (quirk Q4);
// jl[l] and derivative dj[l] (nn/interaction.py:288-348)
template <int LC>
__device__ __forceinline__ void sph_bessel_all(float x, int l_top, float* jl, float* dj) {
  const float EPS = 1e-8f;
  bool big = x > EPS;
  float s, c;
  sincosf(x, &s, &c);
  float sx = s / x;
  jl[0] = big ? sx : 1.0f;
  dj[0] = big ? -((sx - c) / x) : 0.0f;
  if (LC > 1 && l_top >= 1) {
    jl[1] = big ? (sx - c) / x : x / 3.0f;
    dj[1] = big ? (jl[0] - 2.0f / x * jl[1]) : (1.0f / 3.0f);
    float coeff = 3.0f;
#pragma unroll
    for (int n = 1; n < LC - 1; ++n) {
      if (n < l_top) {
        coeff *= (float)(2 * n + 3);
        jl[n + 1] = big ? ((float)(2 * n + 1) / x * jl[n] - jl[n - 1]) : x / coeff;
        dj[n + 1] = big ? (jl[n] - (float)(n + 2) / x * jl[n + 1]) : 0.0f;
      }
    }
  }
}

// chi_d(r) = j_l(z_d r / rc) / factors_d and d chi_d / dr
template <int LC, int RC>
__device__ __forceinline__ void chi_eval(float r, const float* __restrict__ consts, int L, int R, float* chi,
                                         float* dchi) {
  const int D = L * R;
  const float* z = consts;
  const float* fac = consts + D;
  const float rc = consts[2 * D];
#pragma unroll
  for (int l = 0; l < LC; ++l) {
#pragma unroll
    for (int n = 0; n < RC; ++n) {
      if (l < L && n < R) {
        int d = l * R + n;
        float x = __fdiv_rn(__fmul_rn(z[d], r), rc);
        float jl[LC], dj[LC];
        sph_bessel_all<LC>(x, l, jl, dj);
        float jv = jl[0], dv = dj[0];
#pragma unroll
        for (int q = 1; q < LC; ++q)
          if (q == l) { jv = jl[q]; dv = dj[q]; }
        chi[l * RC + n] = jv / fac[d];
        if (dchi) dchi[l * RC + n] = dv * (z[d] / rc) / fac[d];
      }
    }
  }
}

template <int LC>
__device__ __forceinline__ void legendre_all(float x, float* P) {
  P[0] = 1.0f;
  if (LC > 1) P[1] = x;
#pragma unroll
  for (int n = 1; n < LC - 1; ++n) P[n + 1] = ((float)(2 * n + 1) * x * P[n] - (float)n * P[n - 1]) / (float)(n + 1);
}

// reference LegendreCosPolynomial.backward (nn/interaction.py:373-382, quirk Q3): for order l with upstream go,
// g = 0; for n = 1..l: g = (n P_{n-1} + x g) * go
template <int LC>
__device__ __forceinline__ float legendre_bwd_sum(float x, const float* P, const float* go, int L) {
  float total = 0.0f;
#pragma unroll
  for (int l = 1; l < LC; ++l) {
    if (l < L) {
      float g = 0.0f;
#pragma unroll
      for (int n = 1; n <= l; ++n) g = ((float)n * P[n - 1] + x * g) * go[l];
      total += g;
    }
  }
  return total;
}

// ------------------------------------------------------------------------------------------------
__global__ void tb_sigma_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Ws,
                                    const float* __restrict__ bs, int64_t N, int F, int D, float* __restrict__ sig) {
  int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (k >= N) return;
  for (int d = 0; d < D; ++d) {
    float acc = 0.0f;
    for (int f = lane; f < F; f += 32) acc += x[k * F + f] * Ws[d * F + f];
    acc = warp_sum(acc);
    if (lane == 0) sig[k * D + d] = sigmoid_acc(acc + bs[d]);
  }
}

template <int LC, int RC>
__global__ void tb_edge_basis_fwd_kernel(const float4* __restrict__ vec4, const int32_t* __restrict__ dst,
                                         const float* __restrict__ sig, const float* __restrict__ consts, int64_t E,
                                         int L, int R, const int32_t* __restrict__ edge_list,
                                         float* __restrict__ bas) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;  // E = number of work items: all bonds, or the listed (member) bonds only
  if (edge_list) e = edge_list[e];
  const int D = L * R;
  float r = vec4[e].w;
  float r3 = consts[2 * D + 1];
  float c = cutoff_poly(r, r3);
  float chi[LC * RC];
  if (c != 0.0f) chi_eval<LC, RC>(r, consts, L, R, chi, nullptr);
  const float* sg = sig + (int64_t)dst[e] * D;
#pragma unroll
  for (int l = 0; l < LC; ++l)
#pragma unroll
    for (int n = 0; n < RC; ++n)
      if (l < L && n < R) {
        int d = l * R + n;
        bas[e * D + d] = (c != 0.0f) ? chi[l * RC + n] * c * sg[d] : 0.0f;
      }
}

template <int LC, int RC, int G>
__global__ void tb_reduce_fwd_kernel(const float4* __restrict__ vec4, const float* __restrict__ bas,
                                     const int32_t* __restrict__ tri_ptr, const int32_t* __restrict__ tri_e2,
                                     const float* __restrict__ consts, const float* __restrict__ WdT,
                                     const float* __restrict__ WgT, const float* __restrict__ e_in, int64_t E, int L,
                                     int R, int F, float* __restrict__ red, float* __restrict__ e_out) {
  const int D = L * R;
  int64_t e1 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  int gl = threadIdx.x % G;
  bool valid = e1 < E;
  float acc[LC * RC];
#pragma unroll
  for (int d = 0; d < LC * RC; ++d) acc[d] = 0.0f;
  float4 v1 = make_float4(0.f, 0.f, 0.f, 1.f);
  int beg = 0, end = 0;
  if (valid) {
    v1 = vec4[e1];
    beg = tri_ptr[e1];
    end = tri_ptr[e1 + 1];
  }
  for (int p = beg + gl; p < end; p += G) {
    int e2 = tri_e2[p];
    float4 v2 = vec4[e2];
    float dot = __fadd_rn(__fadd_rn(__fmul_rn(v1.x, v2.x), __fmul_rn(v1.y, v2.y)), __fmul_rn(v1.z, v2.z));
    float cs = __fdiv_rn(dot, __fmul_rn(v1.w, v2.w));
    cs = fminf(fmaxf(cs, -1.0f), 1.0f);
    float P[LC];
    legendre_all<LC>(cs, P);
    const float* b2 = bas + (int64_t)e2 * D;
#pragma unroll
    for (int l = 0; l < LC; ++l) {
      if (l < L) {
        float y = kYpref[l] * P[l];
#pragma unroll
        for (int n = 0; n < RC; ++n)
          if (n < R) acc[l * RC + n] += y * b2[l * R + n];
      }
    }
  }
  float c1 = valid ? cutoff_poly(v1.w, consts[2 * D + 1]) : 0.0f;
#pragma unroll
  for (int d = 0; d < LC * RC; ++d) acc[d] = c1 * group_sum<G>(acc[d]);
  if (!valid) return;
  if (gl == 0) {
#pragma unroll
    for (int l = 0; l < LC; ++l)
#pragma unroll
      for (int n = 0; n < RC; ++n)
        if (l < L && n < R) red[e1 * D + l * R + n] = acc[l * RC + n];
  }
  bool any = end > beg;
  for (int f = gl; f < F; f += G) {
    float upd = 0.0f;
    if (any) {
      float u = 0.0f, g = 0.0f;
#pragma unroll
      for (int l = 0; l < LC; ++l)
#pragma unroll
        for (int n = 0; n < RC; ++n)
          if (l < L && n < R) {
            int d = l * R + n;
            u += acc[l * RC + n] * WdT[d * F + f];
            g += acc[l * RC + n] * WgT[d * F + f];
          }
      upd = silu_acc(u) * sigmoid_acc(g);
    }
    e_out[e1 * F + f] = e_in[e1 * F + f] + upd;
  }
}

// adjoint of the bias-free 1-layer gated MLP: 8 lanes per bond (lane g owns features g, g+8, ...), weights staged
// in shared memory; bonds that are not the first bond of any triplet are skipped (their g_red row is never read
// as "first bond" data; it is zero-filled so that stale memory cannot leak).
constexpr int GATE_G = 8;
template <int DM>
__global__ void tb_gate_bwd_kernel(const float* __restrict__ red, const float* __restrict__ g_e,
                                   const float* __restrict__ WdT, const float* __restrict__ WgT,
                                   const int32_t* __restrict__ tri_ptr, int64_t E, int D, int F,
                                   float* __restrict__ g_red) {
  extern __shared__ float w_s[];  // [2][D][F]
  for (int i = threadIdx.x; i < D * F; i += blockDim.x) {
    w_s[i] = WdT[i];
    w_s[D * F + i] = WgT[i];
  }
  __syncthreads();
  int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GATE_G;
  int gl = threadIdx.x % GATE_G;
  bool valid = e < E;
  bool work = valid && (tri_ptr[e + 1] > tri_ptr[e]);
  float rd[DM], acc[DM];
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    rd[d] = (work && d < D) ? red[e * D + d] : 0.0f;
    acc[d] = 0.0f;
  }
  if (work) {
    for (int f = gl; f < F; f += GATE_G) {
      float u = 0.0f, g = 0.0f;
#pragma unroll
      for (int d = 0; d < DM; ++d)
        if (d < D) {
          u += rd[d] * w_s[d * F + f];
          g += rd[d] * w_s[D * F + d * F + f];
        }
      float ge = g_e[e * F + f];
      float sg = sigmoid_acc(g);
      float du = ge * sg * silu_grad(u);
      float dg = ge * silu_acc(u) * sg * (1.0f - sg);
#pragma unroll
      for (int d = 0; d < DM; ++d)
        if (d < D) acc[d] += du * w_s[d * F + f] + dg * w_s[D * F + d * F + f];
    }
  }
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    if (d < D) {
      float s = group_sum<GATE_G>(acc[d]);
      if (valid && gl == (d % GATE_G)) g_red[e * D + d] = s;
    }
  }
}

template <int LC, int RC, int G>
__global__ void tb_reduce_bwd_kernel(const float4* __restrict__ vec4, const float* __restrict__ bas,
                                     const float* __restrict__ g_red, const int32_t* __restrict__ tri_ptr,
                                     const int32_t* __restrict__ tri_e2, const int32_t* __restrict__ trt_ptr,
                                     const int32_t* __restrict__ trt_e1, const float* __restrict__ consts, int64_t E,
                                     int L, int R, float4* __restrict__ g_vec4, float* __restrict__ g_bas) {
  const int D = L * R;
  int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  int gl = threadIdx.x % G;
  bool valid = e < E;
  const float r3 = consts[2 * D + 1];
  float4 ve = make_float4(0.f, 0.f, 0.f, 1.f);
  float d1[LC * RC], be[LC * RC], gB[LC * RC];
#pragma unroll
  for (int d = 0; d < LC * RC; ++d) { d1[d] = 0.f; be[d] = 0.f; gB[d] = 0.f; }
  int a_beg = 0, a_end = 0, b_beg = 0, b_end = 0;
  if (valid) {
    ve = vec4[e];
    a_beg = tri_ptr[e]; a_end = tri_ptr[e + 1];
    b_beg = trt_ptr[e]; b_end = trt_ptr[e + 1];
#pragma unroll
    for (int l = 0; l < LC; ++l)
#pragma unroll
      for (int n = 0; n < RC; ++n)
        if (l < L && n < R) {
          d1[l * RC + n] = g_red[e * D + l * R + n];
          be[l * RC + n] = bas[e * D + l * R + n];
        }
  }
  float ce = cutoff_poly(ve.w, r3);
  float gx = 0.f, gy = 0.f, gz = 0.f, gr = 0.f, gc = 0.f;
  // pass A: e is the first bond of (e, p)
  for (int p = a_beg + gl; p < a_end; p += G) {
    int ep = tri_e2[p];
    float4 vp = vec4[ep];
    float inv = 1.0f / (ve.w * vp.w);
    float craw = (ve.x * vp.x + ve.y * vp.y + ve.z * vp.z) * inv;
    bool inside = (craw >= -1.0f) && (craw <= 1.0f);
    float cs = fminf(fmaxf(craw, -1.0f), 1.0f);
    float P[LC], go[LC];
    legendre_all<LC>(cs, P);
    const float* bp = bas + (int64_t)ep * D;
#pragma unroll
    for (int l = 0; l < LC; ++l) {
      go[l] = 0.0f;
      if (l < L) {
        float s = 0.0f;
#pragma unroll
        for (int n = 0; n < RC; ++n)
          if (n < R) s += bp[l * R + n] * d1[l * RC + n];
        gc += kYpref[l] * P[l] * s;
        go[l] = kYpref[l] * ce * s;
      }
    }
    float gcos = legendre_bwd_sum<LC>(cs, P, go, L);
    if (inside) {
      float w = gcos * inv;
      gx += w * vp.x; gy += w * vp.y; gz += w * vp.z;
      gr -= gcos * craw / ve.w;
    }
  }
  // pass B: e is the second bond of (q, e)
  for (int p = b_beg + gl; p < b_end; p += G) {
    int eq = trt_e1[p];
    float4 vq = vec4[eq];
    float cq = cutoff_poly(vq.w, r3);
    float inv = 1.0f / (ve.w * vq.w);
    float craw = (ve.x * vq.x + ve.y * vq.y + ve.z * vq.z) * inv;
    bool inside = (craw >= -1.0f) && (craw <= 1.0f);
    float cs = fminf(fmaxf(craw, -1.0f), 1.0f);
    float P[LC], go[LC];
    legendre_all<LC>(cs, P);
    const float* dq = g_red + (int64_t)eq * D;
#pragma unroll
    for (int l = 0; l < LC; ++l) {
      go[l] = 0.0f;
      if (l < L) {
        float y = kYpref[l] * P[l] * cq;
        float t = 0.0f;
#pragma unroll
        for (int n = 0; n < RC; ++n)
          if (n < R) {
            float dv = dq[l * R + n];
            gB[l * RC + n] += y * dv;
            t += be[l * RC + n] * dv;
          }
        go[l] = kYpref[l] * cq * t;
      }
    }
    float gcos = legendre_bwd_sum<LC>(cs, P, go, L);
    if (inside) {
      float w = gcos * inv;
      gx += w * vq.x; gy += w * vq.y; gz += w * vq.z;
      gr -= gcos * craw / ve.w;
    }
  }
  gx = group_sum<G>(gx); gy = group_sum<G>(gy); gz = group_sum<G>(gz);
  gr = group_sum<G>(gr); gc = group_sum<G>(gc);
#pragma unroll
  for (int d = 0; d < LC * RC; ++d) gB[d] = group_sum<G>(gB[d]);
  if (!valid || gl != 0) return;
  gr += gc * cutoff_poly_grad(ve.w, r3);
  g_vec4[e] = make_float4(gx, gy, gz, gr);
#pragma unroll
  for (int l = 0; l < LC; ++l)
#pragma unroll
    for (int n = 0; n < RC; ++n)
      if (l < L && n < R) g_bas[e * D + l * R + n] = gB[l * RC + n];
}

template <int LC, int RC>
__global__ void tb_edge_basis_bwd_kernel(const float4* __restrict__ vec4, const int32_t* __restrict__ dst,
                                         const float* __restrict__ sig, const float* __restrict__ g_bas,
                                         const float* __restrict__ consts, int64_t E, int L, int R,
                                         const int32_t* __restrict__ edge_list, float* __restrict__ g_vec4,
                                         float* __restrict__ g_sig_e) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  if (edge_list) e = edge_list[e];
  const int D = L * R;
  float r = vec4[e].w;
  float r3 = consts[2 * D + 1];
  float c = cutoff_poly(r, r3);
  if (c == 0.0f) {
    for (int d = 0; d < D; ++d) g_sig_e[e * D + d] = 0.0f;
    return;
  }
  float dc = cutoff_poly_grad(r, r3);
  float chi[LC * RC], dchi[LC * RC];
  chi_eval<LC, RC>(r, consts, L, R, chi, dchi);
  const float* sg = sig + (int64_t)dst[e] * D;
  float gr = 0.0f;
#pragma unroll
  for (int l = 0; l < LC; ++l)
#pragma unroll
    for (int n = 0; n < RC; ++n)
      if (l < L && n < R) {
        int d = l * R + n;
        float gb = g_bas[e * D + d];
        g_sig_e[e * D + d] = gb * chi[l * RC + n] * c;
        gr += gb * sg[d] * (dchi[l * RC + n] * c + chi[l * RC + n] * dc);
      }
  g_vec4[e * 4 + 3] += gr;
}

// block-invariant radial part of the three-body basis (once per step): G[e][d] = chi_d(r_e) fc(r_e) and dG/dr, for the
// listed (member) bonds.  bas = G * sigma[dst] is then formed per block inside the per-atom moment kernels.
template <int LC, int RC>
__global__ void tb_radial_kernel(const float4* __restrict__ vec4, const float* __restrict__ consts, int64_t n_work,
                                 int L, int R, const int32_t* __restrict__ edge_list, float* __restrict__ G,
                                 float* __restrict__ dG) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_work) return;
  if (edge_list) e = edge_list[e];
  const int D = L * R;
  const float r = vec4[e].w;
  const float r3 = consts[2 * D + 1];
  const float c = cutoff_poly(r, r3);
  const float dc = cutoff_poly_grad(r, r3);
  float chi[LC * RC], dchi[LC * RC];
  if (c != 0.0f) chi_eval<LC, RC>(r, consts, L, R, chi, dchi);
#pragma unroll
  for (int l = 0; l < LC; ++l)
#pragma unroll
    for (int n = 0; n < RC; ++n)
      if (l < L && n < R) {
        const int d = l * R + n;
        G[e * D + d] = (c != 0.0f) ? chi[l * RC + n] * c : 0.0f;
        dG[e * D + d] = (c != 0.0f) ? (dchi[l * RC + n] * c + chi[l * RC + n] * dc) : 0.0f;
      }
}

// sin / cos for the bounded arguments of the radial basis (0 <= x < 64; here x = z_ln r / r_c <= 12.4): three-term
// Cody-Waite reduction by pi/2 and the usual degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4], without the
// large-argument path of sincosf.  Checked on the CPU against float64 over [0, 16]: <= 1.43 ulp for both.
__device__ __forceinline__ void sincos_bounded(float x, float* sn, float* cs) {
  const float j = rintf(x * 0.636619747f);
  float a = fmaf(j, -1.57079601e+00f, x);
  a = fmaf(j, -3.13916473e-07f, a);
  a = fmaf(j, -5.39030253e-15f, a);
  const float s = a * a;
  float r = 2.86567956e-6f;
  r = fmaf(r, s, -1.98559923e-4f);
  r = fmaf(r, s, 8.33338592e-3f);
  r = fmaf(r, s, -1.66666672e-1f);
  const float sv = fmaf(r, a * s, a);
  float c = 2.44677067e-5f;
  c = fmaf(c, s, -1.38877297e-3f);
  c = fmaf(c, s, 4.16666567e-2f);
  c = fmaf(c, s, -5.00000000e-1f);
  const float cv = fmaf(c, s, 1.0f);
  const int q = (int)j;
  const float s0 = (q & 1) ? cv : sv, c0 = (q & 1) ? sv : cv;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
}

// the same for l_max = n_max = 3 (the shape the moment kernels serve): one reciprocal per argument instead of a
// division per recurrence step (results within 2 ulp of the generic kernel; x <= 1e-8 keeps the reference's branch)
// per-block table in shared memory: z (9), z / r_c (9), 1 / factor (9), r_c, 1 / r_c, r3 — the IEEE divisions of uniform
// data are done once per block by nine threads instead of nine times per bond; x = (z r) / r_c keeps the reference's
// rounding through one Newton refinement of the quotient (q = z r; x0 = q inv; x = x0 + (q - x0 r_c) inv).
// Stores: when the 32 bonds of a warp are consecutive rows (always for a dense member list), the 2 x 9 values per bond
// go through a shared tile and leave as coalesced 128-byte lines instead of 18 scalar stores with a 36-byte stride.
__global__ void __launch_bounds__(128) tb_radial33_kernel(const float4* __restrict__ vec4, const float* __restrict__ consts,
                                                          int64_t n_work, const int32_t* __restrict__ edge_list,
                                                          float* __restrict__ G, float* __restrict__ dG) {
  __shared__ float tab[32];
  __shared__ float tile[4][2][9 * 32 + 8];
  if (threadIdx.x < 9) {
    const float z = consts[threadIdx.x], rc = consts[18];
    tab[threadIdx.x] = z;
    tab[9 + threadIdx.x] = __fdiv_rn(z, rc);
    tab[18 + threadIdx.x] = __frcp_rn(consts[9 + threadIdx.x]);
  } else if (threadIdx.x == 9) {
    tab[27] = consts[18];
    tab[28] = __frcp_rn(consts[18]);
    tab[29] = consts[19];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n_work;
  int64_t e = live ? i : n_work - 1;
  if (edge_list) e = edge_list[e];
  const float r = vec4[e].w;
  const float rc = tab[27], inv_rc = tab[28], r3 = tab[29];
  const float c = cutoff_poly(r, r3);
  const float dc = cutoff_poly_grad(r, r3);
  float g[9], dg[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) {
    const int l = d / 3;
    const float q = __fmul_rn(tab[d], r);
    const float x0 = q * inv_rc;
    const float x = fmaf(fmaf(-x0, rc, q), inv_rc, x0);
    float j = 1.0f, dj = 0.0f;
    if (x > 1e-8f) {
      float sn, cs;
      if (x < 64.0f) sincos_bounded(x, &sn, &cs);
      else sincosf(x, &sn, &cs);
      const float ix = __fdividef(1.0f, x);
      const float j0 = sn * ix;
      const float j1 = (j0 - cs) * ix;
      if (l == 0) { j = j0; dj = -j1; }
      else if (l == 1) { j = j1; dj = j0 - 2.0f * ix * j1; }
      else { const float j2 = 3.0f * ix * j1 - j0; j = j2; dj = j1 - 3.0f * ix * j2; }
    } else {
      float jl[3], djl[3];
      sph_bessel_all<3>(x, l, jl, djl);
      j = jl[0]; dj = djl[0];
      if (l == 1) { j = jl[1]; dj = djl[1]; }
      if (l == 2) { j = jl[2]; dj = djl[2]; }
    }
    const float ifac = tab[18 + d];
    const float chi = j * ifac, dchi = dj * tab[9 + d] * ifac;
    g[d] = (c != 0.0f) ? chi * c : 0.0f;
    dg[d] = (c != 0.0f) ? dchi * c + chi * dc : 0.0f;
  }
  // consecutive rows in this warp?
  const int64_t e_first = __shfl_sync(FULL, e, 0);
  const bool dense = __all_sync(FULL, live && e == e_first + lane);
  if (dense) {
    float* tg = tile[warp][0];
    float* td = tile[warp][1];
#pragma unroll
    for (int d = 0; d < 9; ++d) {  // bank (9 lane + d) mod 32: conflict-free
      tg[lane * 9 + d] = g[d];
      td[lane * 9 + d] = dg[d];
    }
    __syncwarp();
    float* Go = G + e_first * 9;
    float* Do = dG + e_first * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      Go[32 * k + lane] = tg[32 * k + lane];
      Do[32 * k + lane] = td[32 * k + lane];
    }
  } else if (live) {
#pragma unroll
    for (int d = 0; d < 9; ++d) {
      G[e * 9 + d] = g[d];
      dG[e * 9 + d] = dg[d];
    }
  }
}

template <int DM>
__global__ void tb_sigma_bwd_kernel(const float* __restrict__ g_sig_e, const int32_t* __restrict__ in_ptr,
                                    const int32_t* __restrict__ in_perm, const float* __restrict__ sig,
                                    const float* __restrict__ Ws, const float* __restrict__ base, int64_t N, int F,
                                    int D, float* __restrict__ g_x) {
  int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (k >= N) return;
  float acc[DM];
#pragma unroll
  for (int d = 0; d < DM; ++d) acc[d] = 0.0f;
  for (int p = in_ptr[k] + lane; p < in_ptr[k + 1]; p += 32) {
    const float* row = g_sig_e + (int64_t)in_perm[p] * D;
#pragma unroll
    for (int d = 0; d < DM; ++d)
      if (d < D) acc[d] += row[d];
  }
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    if (d < D) {
      float s = warp_sum(acc[d]);
      float sg = sig[k * D + d];
      acc[d] = s * sg * (1.0f - sg);
    }
  }
  for (int f = lane; f < F; f += 32) {
    float v = 0.0f;
#pragma unroll
    for (int d = 0; d < DM; ++d)
      if (d < D) v += acc[d] * Ws[d * F + f];
    g_x[k * F + f] = base ? base[k * F + f] + v : v;
  }
}

// ================================================================================================
// Specialised kernels for the default model shape (l_max = n_max = 3, F = 64): compile-time loops, gated-MLP
// weights in shared memory, every lane owns F/G contiguous features (vector loads / stores of the edge rows),
// persistent blocks (grid-stride over bond groups) so that the weight staging is amortised.
// ================================================================================================
constexpr int FD = 9;   // D = 3 x 3
constexpr int FF = 64;  // feature width

__device__ __forceinline__ float silu_f(float z) { return __fdividef(z, 1.0f + __expf(-z)); }
__device__ __forceinline__ float sigmoid_f(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }

__device__ __forceinline__ void legendre3(float x, float& y0, float& y1, float& y2) {
  y0 = kYpref[0];
  y1 = kYpref[1] * x;
  y2 = kYpref[2] * ((3.0f * x * x - 1.0f) * 0.5f);
}

template <int G>
__global__ void __launch_bounds__(256) tb_reduce_fwd_fast_kernel(
    const float4* __restrict__ vec4, const float* __restrict__ bas, const int32_t* __restrict__ tri_ptr,
    const int32_t* __restrict__ tri_e2, float r3, const float* __restrict__ WdT, const float* __restrict__ WgT,
    const float* __restrict__ e_in, int64_t E, float* __restrict__ red, float* __restrict__ e_out) {
  constexpr int FPL = FF / G;
  __shared__ __align__(16) float wd_s[FD * FF];
  __shared__ __align__(16) float wg_s[FD * FF];
  for (int i = threadIdx.x; i < FD * FF; i += blockDim.x) {
    wd_s[i] = WdT[i];
    wg_s[i] = WgT[i];
  }
  __syncthreads();
  const int gpb = 256 / G;
  const int gl = threadIdx.x % G;
  const int f0 = gl * FPL;
  for (int64_t base = (int64_t)blockIdx.x * gpb; base < E; base += (int64_t)gridDim.x * gpb) {
    int64_t e1 = base + threadIdx.x / G;
    bool valid = e1 < E;
    float acc[FD];
#pragma unroll
    for (int d = 0; d < FD; ++d) acc[d] = 0.0f;
    float4 v1 = make_float4(0.f, 0.f, 0.f, 1.f);
    int beg = 0, end = 0;
    if (valid) {
      v1 = vec4[e1];
      beg = tri_ptr[e1];
      end = tri_ptr[e1 + 1];
    }
    for (int p = beg + gl; p < end; p += G) {
      int e2 = tri_e2[p];
      float4 v2 = vec4[e2];
      float dot = __fadd_rn(__fadd_rn(__fmul_rn(v1.x, v2.x), __fmul_rn(v1.y, v2.y)), __fmul_rn(v1.z, v2.z));
      float cs = fminf(fmaxf(__fdiv_rn(dot, __fmul_rn(v1.w, v2.w)), -1.0f), 1.0f);
      float y0, y1, y2;
      legendre3(cs, y0, y1, y2);
      const float* b2 = bas + (int64_t)e2 * FD;
      acc[0] += y0 * b2[0]; acc[1] += y0 * b2[1]; acc[2] += y0 * b2[2];
      acc[3] += y1 * b2[3]; acc[4] += y1 * b2[4]; acc[5] += y1 * b2[5];
      acc[6] += y2 * b2[6]; acc[7] += y2 * b2[7]; acc[8] += y2 * b2[8];
    }
    float c1 = valid ? cutoff_poly(v1.w, r3) : 0.0f;
#pragma unroll
    for (int d = 0; d < FD; ++d) acc[d] = c1 * group_sum<G>(acc[d]);
    if (!valid) continue;
    if (gl == 0) {
#pragma unroll
      for (int d = 0; d < FD; ++d) red[e1 * FD + d] = acc[d];
    }
    float row[FPL];
#pragma unroll
    for (int j = 0; j < FPL; j += 2) {
      float2 t = *reinterpret_cast<const float2*>(e_in + e1 * FF + f0 + j);
      row[j] = t.x;
      row[j + 1] = t.y;
    }
    if (end > beg) {
      float u[FPL], g[FPL];
#pragma unroll
      for (int j = 0; j < FPL; ++j) { u[j] = 0.0f; g[j] = 0.0f; }
#pragma unroll
      for (int d = 0; d < FD; ++d) {
#pragma unroll
        for (int j = 0; j < FPL; ++j) {
          u[j] += acc[d] * wd_s[d * FF + f0 + j];
          g[j] += acc[d] * wg_s[d * FF + f0 + j];
        }
      }
#pragma unroll
      for (int j = 0; j < FPL; ++j) row[j] += silu_f(u[j]) * sigmoid_f(g[j]);
    }
#pragma unroll
    for (int j = 0; j < FPL; j += 2)
      *reinterpret_cast<float2*>(e_out + e1 * FF + f0 + j) = make_float2(row[j], row[j + 1]);
  }
}

template <int G>
__global__ void __launch_bounds__(256) tb_gate_bwd_fast_kernel(
    const float* __restrict__ red, const float* __restrict__ g_e, const float* __restrict__ WdT,
    const float* __restrict__ WgT, const int32_t* __restrict__ tri_ptr, int64_t E, float* __restrict__ g_red) {
  constexpr int FPL = FF / G;
  __shared__ __align__(16) float wd_s[FD * FF];
  __shared__ __align__(16) float wg_s[FD * FF];
  for (int i = threadIdx.x; i < FD * FF; i += blockDim.x) {
    wd_s[i] = WdT[i];
    wg_s[i] = WgT[i];
  }
  __syncthreads();
  const int gpb = 256 / G;
  const int gl = threadIdx.x % G;
  const int f0 = gl * FPL;
  for (int64_t base = (int64_t)blockIdx.x * gpb; base < E; base += (int64_t)gridDim.x * gpb) {
    int64_t e = base + threadIdx.x / G;
    bool valid = e < E;
    bool work = valid && (tri_ptr[e + 1] > tri_ptr[e]);
    float acc[FD];
#pragma unroll
    for (int d = 0; d < FD; ++d) acc[d] = 0.0f;
    if (work) {
      float rd[FD];
#pragma unroll
      for (int d = 0; d < FD; ++d) rd[d] = red[e * FD + d];
      float u[FPL], g[FPL];
#pragma unroll
      for (int j = 0; j < FPL; ++j) { u[j] = 0.0f; g[j] = 0.0f; }
#pragma unroll
      for (int d = 0; d < FD; ++d)
#pragma unroll
        for (int j = 0; j < FPL; ++j) {
          u[j] += rd[d] * wd_s[d * FF + f0 + j];
          g[j] += rd[d] * wg_s[d * FF + f0 + j];
        }
#pragma unroll
      for (int j = 0; j < FPL; j += 2) {
        float2 ge2 = *reinterpret_cast<const float2*>(g_e + e * FF + f0 + j);
        float gev[2] = {ge2.x, ge2.y};
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          float sg = sigmoid_f(g[j + t]);
          float su = sigmoid_f(u[j + t]);
          float du = gev[t] * sg * su * (1.0f + u[j + t] * (1.0f - su));
          float dg = gev[t] * (u[j + t] * su) * sg * (1.0f - sg);
#pragma unroll
          for (int d = 0; d < FD; ++d)
            acc[d] += du * wd_s[d * FF + f0 + j + t] + dg * wg_s[d * FF + f0 + j + t];
        }
      }
    }
#pragma unroll
    for (int d = 0; d < FD; ++d) acc[d] = group_sum<G>(acc[d]);
    if (valid && gl == 0) {
#pragma unroll
      for (int d = 0; d < FD; ++d) g_red[e * FD + d] = acc[d];
    }
  }
}

// symmetric triplet lists ((e,p) present <=> (p,e) present): one sweep over the partners serves both roles of e
template <int G>
__global__ void __launch_bounds__(256) tb_reduce_bwd_sym_kernel(
    const float4* __restrict__ vec4, const float* __restrict__ bas, const float* __restrict__ g_red,
    const int32_t* __restrict__ tri_ptr, const int32_t* __restrict__ tri_e2, float r3, int64_t E,
    float4* __restrict__ g_vec4, float* __restrict__ g_bas) {
  const int gpb = 256 / G;
  const int gl = threadIdx.x % G;
  for (int64_t base = (int64_t)blockIdx.x * gpb; base < E; base += (int64_t)gridDim.x * gpb) {
    int64_t e = base + threadIdx.x / G;
    bool valid = e < E;
    float4 ve = make_float4(0.f, 0.f, 0.f, 1.f);
    float d1[FD], be[FD], gB[FD];
#pragma unroll
    for (int d = 0; d < FD; ++d) { d1[d] = 0.f; be[d] = 0.f; gB[d] = 0.f; }
    int beg = 0, end = 0;
    if (valid) {
      ve = vec4[e];
      beg = tri_ptr[e];
      end = tri_ptr[e + 1];
      if (end > beg) {
#pragma unroll
        for (int d = 0; d < FD; ++d) {
          d1[d] = g_red[e * FD + d];
          be[d] = bas[e * FD + d];
        }
      }
    }
    const float ce = cutoff_poly(ve.w, r3);
    float gx = 0.f, gy = 0.f, gz = 0.f, gr = 0.f, gc = 0.f;
    for (int p = beg + gl; p < end; p += G) {
      int ep = tri_e2[p];
      float4 vp = vec4[ep];
      float cp = cutoff_poly(vp.w, r3);
      float inv = 1.0f / (ve.w * vp.w);
      float craw = (ve.x * vp.x + ve.y * vp.y + ve.z * vp.z) * inv;
      bool inside = (craw >= -1.0f) && (craw <= 1.0f);
      float cs = fminf(fmaxf(craw, -1.0f), 1.0f);
      float y0, y1, y2;
      legendre3(cs, y0, y1, y2);
      const float* bp = bas + (int64_t)ep * FD;
      const float* dq = g_red + (int64_t)ep * FD;
      float b[FD], q[FD];
#pragma unroll
      for (int d = 0; d < FD; ++d) { b[d] = bp[d]; q[d] = dq[d]; }
      // e as first bond of (e, p)
      float s0 = b[0] * d1[0] + b[1] * d1[1] + b[2] * d1[2];
      float s1 = b[3] * d1[3] + b[4] * d1[4] + b[5] * d1[5];
      float s2 = b[6] * d1[6] + b[7] * d1[7] + b[8] * d1[8];
      gc += y0 * s0 + y1 * s1 + y2 * s2;
      // e as second bond of (p, e)
      float w0 = y0 * cp, w1 = y1 * cp, w2 = y2 * cp;
      gB[0] += w0 * q[0]; gB[1] += w0 * q[1]; gB[2] += w0 * q[2];
      gB[3] += w1 * q[3]; gB[4] += w1 * q[4]; gB[5] += w1 * q[5];
      gB[6] += w2 * q[6]; gB[7] += w2 * q[7]; gB[8] += w2 * q[8];
      float t1 = be[3] * q[3] + be[4] * q[4] + be[5] * q[5];
      float t2 = be[6] * q[6] + be[7] * q[7] + be[8] * q[8];
      // Legendre backward with the reference's per-level grad_output (quirk Q3): l=1: go; l=2: go*(2x + x*go)
      float goA1 = kYpref[1] * ce * s1, goA2 = kYpref[2] * ce * s2;
      float goB1 = kYpref[1] * cp * t1, goB2 = kYpref[2] * cp * t2;
      float gcos = goA1 + goA2 * (2.0f * cs + cs * goA2) + goB1 + goB2 * (2.0f * cs + cs * goB2);
      if (inside) {
        float w = gcos * inv;
        gx += w * vp.x; gy += w * vp.y; gz += w * vp.z;
        gr -= gcos * craw / ve.w;
      }
    }
    gx = group_sum<G>(gx); gy = group_sum<G>(gy); gz = group_sum<G>(gz);
    gr = group_sum<G>(gr); gc = group_sum<G>(gc);
#pragma unroll
    for (int d = 0; d < FD; ++d) gB[d] = group_sum<G>(gB[d]);
    if (!valid || gl != 0) continue;
    gr += gc * cutoff_poly_grad(ve.w, r3);
    g_vec4[e] = make_float4(gx, gy, gz, gr);
#pragma unroll
    for (int d = 0; d < FD; ++d) g_bas[e * FD + d] = gB[d];
  }
}

}  // namespace m3g

using namespace m3g;

#define M3G_CHECK_LR(name)                                                                              \
  M3G_REQUIRE(L >= 1 && L <= M3G_MAX_L && R >= 1 && R <= M3G_MAX_R,                                     \
              name ": l_max=%d n_max=%d outside the supported range [1,%d]x[1,%d]", L, R, M3G_MAX_L, M3G_MAX_R)

#define M3G_DISPATCH_LR(KERNEL, ...)             \
  do {                                           \
    if (L <= 3 && R <= 3) { KERNEL(3, 3, __VA_ARGS__); } \
    else if (L <= 4 && R <= 4) { KERNEL(4, 4, __VA_ARGS__); } \
    else { KERNEL(M3G_MAX_L, M3G_MAX_R, __VA_ARGS__); } \
  } while (0)

extern "C" {

int m3g_tb_sigma_fwd(const float* x, const float* Ws, const float* bs, int64_t N, int F, int D, float* sig,
                     void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(x && Ws && bs && sig, "m3g_tb_sigma_fwd: null pointer");
  tb_sigma_fwd_kernel<<<blocks_for(N * 32, 256), 256, 0, as_stream(stream)>>>(x, Ws, bs, N, F, D, sig);
  M3G_LAUNCH_CHECK("m3g_tb_sigma_fwd");
  return M3G_OK;
}

int m3g_tb_edge_basis_fwd(const float* vec4, const int32_t* dst, const float* sig, const float* tb_consts, int64_t E,
                          int L, int R, const int32_t* edge_list, int64_t n_list, float* bas, void* stream) {
  const int64_t n_work = edge_list ? n_list : E;
  if (n_work == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && dst && sig && tb_consts && bas, "m3g_tb_edge_basis_fwd: null pointer");
  M3G_CHECK_LR("m3g_tb_edge_basis_fwd");
#define K_(LC, RC, ...) \
  tb_edge_basis_fwd_kernel<LC, RC><<<blocks_for(n_work, 128), 128, 0, as_stream(stream)>>>(__VA_ARGS__)
  M3G_DISPATCH_LR(K_, (const float4*)vec4, dst, sig, tb_consts, n_work, L, R, edge_list, bas);
#undef K_
  M3G_LAUNCH_CHECK("m3g_tb_edge_basis_fwd");
  return M3G_OK;
}

int m3g_tb_reduce_fwd(const float* vec4, const float* bas, const int32_t* tri_ptr, const int32_t* tri_e2,
                      const float* tb_consts, const float* WdT, const float* WgT, const float* e_in, int64_t E, int L,
                      int R, int F, int group, float* red, float* e_out, void* stream) {
  if (E == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && bas && tri_ptr && tb_consts && WdT && WgT && e_in && red && e_out,
              "m3g_tb_reduce_fwd: null pointer");
  M3G_CHECK_LR("m3g_tb_reduce_fwd");
  M3G_REQUIRE(group == 8 || group == 16 || group == 32, "m3g_tb_reduce_fwd: group must be 8, 16 or 32");
#define K_(LC, RC, ...)                                                                                              \
  do {                                                                                                               \
    if (group == 8) tb_reduce_fwd_kernel<LC, RC, 8><<<blocks_for(E * 8, 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__);        \
    else if (group == 16) tb_reduce_fwd_kernel<LC, RC, 16><<<blocks_for(E * 16, 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__); \
    else tb_reduce_fwd_kernel<LC, RC, 32><<<blocks_for(E * 32, 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__);                 \
  } while (0)
  M3G_DISPATCH_LR(K_, (const float4*)vec4, bas, tri_ptr, tri_e2, tb_consts, WdT, WgT, e_in, E, L, R, F, red, e_out);
#undef K_
  M3G_LAUNCH_CHECK("m3g_tb_reduce_fwd");
  return M3G_OK;
}

int m3g_tb_gate_bwd(const float* red, const float* g_e, const float* WdT, const float* WgT, const int32_t* tri_ptr,
                    int64_t E, int D, int F, float* g_red, void* stream) {
  if (E == 0) return M3G_OK;
  M3G_REQUIRE(red && g_e && WdT && WgT && tri_ptr && g_red, "m3g_tb_gate_bwd: null pointer");
  M3G_REQUIRE(D >= 1 && D <= M3G_MAX_L * M3G_MAX_R, "m3g_tb_gate_bwd: D=%d unsupported", D);
  size_t smem = (size_t)2 * D * F * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(tb_gate_bwd_kernel<M3G_MAX_L * M3G_MAX_R>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_error("m3g_tb_gate_bwd: %zu bytes of shared memory for D=%d, F=%d: %s", smem, D, F, cudaGetErrorString(err));
      return M3G_ERR_INVALID;
    }
  }
  if (D <= 16)
    tb_gate_bwd_kernel<16><<<blocks_for(E * GATE_G, 256), 256, smem, as_stream(stream)>>>(red, g_e, WdT, WgT, tri_ptr,
                                                                                          E, D, F, g_red);
  else
    tb_gate_bwd_kernel<M3G_MAX_L * M3G_MAX_R><<<blocks_for(E * GATE_G, 256), 256, smem, as_stream(stream)>>>(
        red, g_e, WdT, WgT, tri_ptr, E, D, F, g_red);
  M3G_LAUNCH_CHECK("m3g_tb_gate_bwd");
  return M3G_OK;
}

int m3g_tb_reduce_bwd(const float* vec4, const float* bas, const float* g_red, const int32_t* tri_ptr,
                      const int32_t* tri_e2, const int32_t* trt_ptr, const int32_t* trt_e1, const float* tb_consts,
                      int64_t E, int L, int R, int group, float* g_vec4, float* g_bas, void* stream) {
  if (E == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && bas && g_red && tri_ptr && trt_ptr && tb_consts && g_vec4 && g_bas,
              "m3g_tb_reduce_bwd: null pointer");
  M3G_CHECK_LR("m3g_tb_reduce_bwd");
  M3G_REQUIRE(group == 8 || group == 16 || group == 32, "m3g_tb_reduce_bwd: group must be 8, 16 or 32");
#define K_(LC, RC, ...)                                                                                              \
  do {                                                                                                               \
    if (group == 8) tb_reduce_bwd_kernel<LC, RC, 8><<<blocks_for(E * 8, 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__);        \
    else if (group == 16) tb_reduce_bwd_kernel<LC, RC, 16><<<blocks_for(E * 16, 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__); \
    else tb_reduce_bwd_kernel<LC, RC, 32><<<blocks_for(E * 32, 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__);                 \
  } while (0)
  M3G_DISPATCH_LR(K_, (const float4*)vec4, bas, g_red, tri_ptr, tri_e2, trt_ptr, trt_e1, tb_consts, E, L, R,
                  (float4*)g_vec4, g_bas);
#undef K_
  M3G_LAUNCH_CHECK("m3g_tb_reduce_bwd");
  return M3G_OK;
}

int m3g_tb_edge_basis_bwd(const float* vec4, const int32_t* dst, const float* sig, const float* g_bas,
                          const float* tb_consts, int64_t E, int L, int R, const int32_t* edge_list, int64_t n_list,
                          float* g_vec4, float* g_sig_e, void* stream) {
  const int64_t n_work = edge_list ? n_list : E;
  if (n_work == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && dst && sig && g_bas && tb_consts && g_vec4 && g_sig_e, "m3g_tb_edge_basis_bwd: null pointer");
  M3G_CHECK_LR("m3g_tb_edge_basis_bwd");
#define K_(LC, RC, ...) \
  tb_edge_basis_bwd_kernel<LC, RC><<<blocks_for(n_work, 128), 128, 0, as_stream(stream)>>>(__VA_ARGS__)
  M3G_DISPATCH_LR(K_, (const float4*)vec4, dst, sig, g_bas, tb_consts, n_work, L, R, edge_list, g_vec4, g_sig_e);
#undef K_
  M3G_LAUNCH_CHECK("m3g_tb_edge_basis_bwd");
  return M3G_OK;
}

int m3g_tb_radial(const float* vec4, const float* tb_consts, int64_t E, int L, int R, const int32_t* edge_list,
                  int64_t n_list, float* G, float* dG, void* stream) {
  const int64_t n_work = edge_list ? n_list : E;
  if (n_work == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && tb_consts && G && dG, "m3g_tb_radial: null pointer");
  M3G_CHECK_LR("m3g_tb_radial");
  if (L == 3 && R == 3) {
    tb_radial33_kernel<<<blocks_for(n_work, 128), 128, 0, as_stream(stream)>>>((const float4*)vec4, tb_consts, n_work,
                                                                              edge_list, G, dG);
  } else {
#define K_(LC, RC, ...) tb_radial_kernel<LC, RC><<<blocks_for(n_work, 128), 128, 0, as_stream(stream)>>>(__VA_ARGS__)
    M3G_DISPATCH_LR(K_, (const float4*)vec4, tb_consts, n_work, L, R, edge_list, G, dG);
#undef K_
  }
  M3G_LAUNCH_CHECK("m3g_tb_radial");
  return M3G_OK;
}

int m3g_tb_sigma_bwd(const float* g_sig_e, const int32_t* in_ptr, const int32_t* in_perm, const float* sig,
                     const float* Ws, const float* base, int64_t N, int F, int D, float* g_x, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(g_sig_e && in_ptr && in_perm && sig && Ws && g_x, "m3g_tb_sigma_bwd: null pointer");
  M3G_REQUIRE(D >= 1 && D <= M3G_MAX_L * M3G_MAX_R, "m3g_tb_sigma_bwd: D=%d unsupported", D);
  if (D <= 16)
    tb_sigma_bwd_kernel<16><<<blocks_for(N * 32, 256), 256, 0, as_stream(stream)>>>(g_sig_e, in_ptr, in_perm, sig, Ws,
                                                                                    base, N, F, D, g_x);
  else
    tb_sigma_bwd_kernel<M3G_MAX_L * M3G_MAX_R><<<blocks_for(N * 32, 256), 256, 0, as_stream(stream)>>>(
        g_sig_e, in_ptr, in_perm, sig, Ws, base, N, F, D, g_x);
  M3G_LAUNCH_CHECK("m3g_tb_sigma_bwd");
  return M3G_OK;
}

#define M3G_GROUP_SWITCH(KERNEL, GRID_EXPR, ...)                                                      \
  do {                                                                                               \
    if (group == 8) KERNEL<8><<<GRID_EXPR(8), 256, 0, as_stream(stream)>>>(__VA_ARGS__);              \
    else if (group == 16) KERNEL<16><<<GRID_EXPR(16), 256, 0, as_stream(stream)>>>(__VA_ARGS__);      \
    else KERNEL<32><<<GRID_EXPR(32), 256, 0, as_stream(stream)>>>(__VA_ARGS__);                       \
  } while (0)

static inline unsigned persistent_grid(int64_t E, int group, int n_sm) {
  int64_t need = (E * group + 255) / 256;
  int64_t cap = (int64_t)n_sm * 8;
  return (unsigned)((need < cap) ? (need < 1 ? 1 : need) : cap);
}
#define M3G_PGRID(G_) persistent_grid(E, G_, n_sm)

int m3g_tb_reduce_fwd_fast(const float* vec4, const float* bas, const int32_t* tri_ptr, const int32_t* tri_e2,
                           float r3, const float* WdT, const float* WgT, const float* e_in, int64_t E, int group,
                           int n_sm, float* red, float* e_out, void* stream) {
  if (E == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && bas && tri_ptr && WdT && WgT && e_in && red && e_out, "m3g_tb_reduce_fwd_fast: null pointer");
  M3G_REQUIRE(group == 8 || group == 16 || group == 32, "m3g_tb_reduce_fwd_fast: group must be 8, 16 or 32");
  M3G_GROUP_SWITCH(tb_reduce_fwd_fast_kernel, M3G_PGRID, (const float4*)vec4, bas, tri_ptr, tri_e2, r3, WdT, WgT, e_in,
                   E, red, e_out);
  M3G_LAUNCH_CHECK("m3g_tb_reduce_fwd_fast");
  return M3G_OK;
}

int m3g_tb_gate_bwd_fast(const float* red, const float* g_e, const float* WdT, const float* WgT,
                         const int32_t* tri_ptr, int64_t E, int n_sm, float* g_red, void* stream) {
  if (E == 0) return M3G_OK;
  M3G_REQUIRE(red && g_e && WdT && WgT && tri_ptr && g_red, "m3g_tb_gate_bwd_fast: null pointer");
  tb_gate_bwd_fast_kernel<8><<<persistent_grid(E, 8, n_sm), 256, 0, as_stream(stream)>>>(red, g_e, WdT, WgT, tri_ptr,
                                                                                         E, g_red);
  M3G_LAUNCH_CHECK("m3g_tb_gate_bwd_fast");
  return M3G_OK;
}

int m3g_tb_reduce_bwd_sym(const float* vec4, const float* bas, const float* g_red, const int32_t* tri_ptr,
                          const int32_t* tri_e2, float r3, int64_t E, int group, int n_sm, float* g_vec4, float* g_bas,
                          void* stream) {
  if (E == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && bas && g_red && tri_ptr && g_vec4 && g_bas, "m3g_tb_reduce_bwd_sym: null pointer");
  M3G_REQUIRE(group == 8 || group == 16 || group == 32, "m3g_tb_reduce_bwd_sym: group must be 8, 16 or 32");
  M3G_GROUP_SWITCH(tb_reduce_bwd_sym_kernel, M3G_PGRID, (const float4*)vec4, bas, g_red, tri_ptr, tri_e2, r3, E,
                   (float4*)g_vec4, g_bas);
  M3G_LAUNCH_CHECK("m3g_tb_reduce_bwd_sym");
  return M3G_OK;
}

}  // extern "C"
