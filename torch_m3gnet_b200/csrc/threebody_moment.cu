// ThreeBodyInteration (nn/interaction.py:187-223) in O(n3) per centre atom: the pair sum over partner bonds factorises
// into per-atom MOMENTS of the unit bond vectors, because for l <= 2 the angular factor Y_l(u_j . u_k) is a polynomial
// in u_k (nn/interaction.py:195-200, :353-371):
//
//   red_j[0,n] = c_j Y0 ( S0[n]                      - b_j[0,n] )         S0[n] = sum_k b_k[0,n]
//   red_j[1,n] = c_j Y1 ( u_j . V1[n]                - b_j[1,n] )         V1[n] = sum_k b_k[1,n] u_k
//   red_j[2,n] = c_j Y2 ( 3/2 u_j^T M2[n] u_j - 1/2 tr M2[n] - b_j[2,n] ) M2[n] = sum_k b_k[2,n] u_k u_k^T
//
// (b_k = chi_ln(r_k) fc(r_k) sigma[dst k]; the subtracted term is the excluded pair k = j, where cos = 1).  The adjoint
// has the same structure with the moments of a_j = c_j dL/dred_j (roles of first / second bond swapped), and the
// reference's Legendre backward (quirk Q3: d/dx P_2 -> go (2x + x go), QUADRATIC in the upstream gradient) needs the
// second-order moments Q[n][m] = sum_k b_k[2,n] b_k[2,m] u_k u_k^T (and the same of a).  The algebra is restated on the
// CPU in oracle/threebody_moments.py and checked there against the reference's autograd (float64: 1e-16).
//
// One warp per centre atom; member bonds are staged in shared memory once (O(n3) reads), moments are accumulated with
// one lane per moment component (members in ascending order: the stated accumulation order), then every lane evaluates
// one member bond.  No atomics, no triplet index list, no O(n3^2) work.  The radial part G = chi fc (and dG/dr) is
// block-invariant and comes from m3g_tb_radial (once per step, csrc/threebody.cu).
#include <cstdlib>

#include "common.cuh"

namespace m3g {

namespace {

constexpr int MD = 9;    // D = 3 x 3
constexpr int MF = 64;   // feature width
constexpr int MW = 4;    // warps (= atoms in flight) per CTA
constexpr int ES = 36;   // entry stride in floats (36 j mod 32 = 4 j: conflict-free 128-bit row-per-lane access)
// entry layout (floats): 0 ux uy uz 1 | 4 xx yy zz xy | 8 xz yz c r | 12 b0[3] eidx | 16 b1[3] - | 20 b2[3] - |
//                        24 q0[3] c | 28 q1[3] c | 32 q2[3] c        (q slots: red during the MLP adjoint, then q = dL/dred;
//                                                                     a = c q is formed on the fly: c itself may round
//                                                                     to zero next to the cutoff while fc' does not)
constexpr int O_U = 0, O_O6 = 4, O_C = 10, O_R = 11, O_B = 12, O_E = 15, O_A = 24;
// moment layout per side (floats), every group 16-byte aligned so that the evaluation reads it with broadcast LDS.128
// (ncu: the scalar reads of ~110 moment words per bond and role made the kernels shared-memory-wavefront bound):
//   S0 [0,3) | V1 [4,16) n*4+a | M2 [16,40) n*8+s | Q [40,112) (n*3+m)*8+s
constexpr int M_S0 = 0, M_V1 = 4, M_M2 = 16, M_Q = 40, M_SIDE = 112;
constexpr int V1S = 4, M2S = 8, QS = 8;  // strides of V1[n], M2[n], Q[n][m]

constexpr float kY0 = 0.28209479177387814f, kY1 = 0.4886025119029199f, kY2 = 0.6307831305050401f;

__device__ __forceinline__ float silu_m(float z) { return __fdividef(z, 1.0f + __expf(-z)); }
__device__ __forceinline__ float sigmoid_m(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }
__device__ __forceinline__ float gated_m(float u, float g) {
  return __fdividef(u, (1.0f + __expf(-u)) * (1.0f + __expf(-g)));
}

// packed fp32 pair arithmetic (FFMA2 / FMUL2 on sm_100): two feature columns per lane
__device__ __forceinline__ float2 fma2s(float s, float2 b, float2 c) {
  unsigned long long ra, rb = *reinterpret_cast<unsigned long long*>(&b), rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ra) : "f"(s));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

// per-lane role in the moment accumulation: coefficient vector (b_l / a_l), component of it, shape word, output slots
struct LaneRole {
  int vec;    // float offset of the coefficient vector inside the b block (0, 4, 8 for l = 0, 1, 2)
  int nsel;   // which component of the vector multiplies the shape
  int shape;  // float offset of the shape word (u_a, 1 or u_a u_b)
  int lin;    // output slot of the linear moment, -1 for idle lanes
  int quad;   // output slot of Q[nsel][0][s] (stride 6 per m), -1 if the lane has none
};
__device__ __forceinline__ LaneRole lane_role(int lane) {
  LaneRole r;
  if (lane < 18) {  // M2[n][s] and Q[n][.][s]
    const int n = lane / 6, s = lane - 6 * n;
    r.vec = 8; r.nsel = n; r.shape = O_O6 + s; r.lin = M_M2 + n * M2S + s; r.quad = M_Q + n * 3 * QS + s;
  } else if (lane < 27) {  // V1[n][a]
    const int n = (lane - 18) / 3, a = (lane - 18) - 3 * n;
    r.vec = 4; r.nsel = n; r.shape = O_U + a; r.lin = M_V1 + n * V1S + a; r.quad = -1;
  } else if (lane < 30) {  // S0[n]
    r.vec = 0; r.nsel = lane - 27; r.shape = O_U + 3; r.lin = M_S0 + (lane - 27); r.quad = -1;
  } else {
    r.vec = 0; r.nsel = 0; r.shape = O_U + 3; r.lin = -1; r.quad = -1;
  }
  return r;
}

// accumulate the moments of one side (coefficients at float offset `base` of every entry): members ascending.
// SCALED: the coefficient vector is (q0, q1, q2, c) and stands for a = c q
template <bool WITH_Q, bool SCALED>
__device__ __forceinline__ void accumulate_side(const float* ent, int n3, int base, const LaneRole& role,
                                                float* __restrict__ mom) {
  float lin = 0.0f, q0 = 0.0f, q1 = 0.0f, q2 = 0.0f;
  const int voff = base + role.vec;
#pragma unroll 2
  for (int k = 0; k < n3; ++k) {
    const float* e = ent + k * ES;
    float4 cv = *reinterpret_cast<const float4*>(e + voff);
    if (SCALED) { cv.x *= cv.w; cv.y *= cv.w; cv.z *= cv.w; }
    const float sh = e[role.shape];
    const float cn = (role.nsel == 0) ? cv.x : ((role.nsel == 1) ? cv.y : cv.z);
    const float p = cn * sh;
    lin += p;
    if (WITH_Q) { q0 = fmaf(p, cv.x, q0); q1 = fmaf(p, cv.y, q1); q2 = fmaf(p, cv.z, q2); }
  }
  if (role.lin >= 0) mom[role.lin] = lin;
  if (WITH_Q && role.quad >= 0) { mom[role.quad] = q0; mom[role.quad + QS] = q1; mom[role.quad + 2 * QS] = q2; }
}

// u^T M u and M u for a symmetric matrix given as (xx yy zz xy xz yz); o6 = products of u
__device__ __forceinline__ float quad6(const float* M, const float* o6) {
  return M[0] * o6[0] + M[1] * o6[1] + M[2] * o6[2] + 2.0f * (M[3] * o6[3] + M[4] * o6[4] + M[5] * o6[5]);
}
__device__ __forceinline__ void matvec6(const float* M, const float* u, float* out) {
  out[0] = M[0] * u[0] + M[3] * u[1] + M[4] * u[2];
  out[1] = M[3] * u[0] + M[1] * u[1] + M[5] * u[2];
  out[2] = M[4] * u[0] + M[5] * u[1] + M[2] * u[2];
}

// the linear moments of one side in registers (broadcast LDS.128: 10 loads instead of 30)
struct Mom {
  float S0[3], V1[9], M2[18];
};
__device__ __forceinline__ Mom load_moments(const float* mom) {
  Mom m;
  const float4* p = reinterpret_cast<const float4*>(mom);
  const float4 s = p[0];
  m.S0[0] = s.x; m.S0[1] = s.y; m.S0[2] = s.z;
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    const float4 v = p[1 + n];
    m.V1[3 * n] = v.x; m.V1[3 * n + 1] = v.y; m.V1[3 * n + 2] = v.z;
    const float4 a = p[4 + 2 * n], b = p[5 + 2 * n];
    m.M2[6 * n] = a.x; m.M2[6 * n + 1] = a.y; m.M2[6 * n + 2] = a.z; m.M2[6 * n + 3] = a.w;
    m.M2[6 * n + 4] = b.x; m.M2[6 * n + 5] = b.y;
  }
  return m;
}

// acc_j[l,n] of the header comment (without the c_j factor) from one side's moments and the bond's own coefficients
__device__ __forceinline__ void eval_rows(const Mom& mom, const float* u, const float* o6, const float* own,
                                          float* acc) {
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    acc[n] = kY0 * (mom.S0[n] - own[n]);
    const float* v = mom.V1 + 3 * n;
    acc[3 + n] = kY1 * ((v[0] * u[0] + v[1] * u[1] + v[2] * u[2]) - own[3 + n]);
    const float* m2 = mom.M2 + 6 * n;
    const float tr = m2[0] + m2[1] + m2[2];
    acc[6 + n] = kY2 * ((1.5f * quad6(m2, o6) - 0.5f * tr) - own[6 + n]);
  }
}

// d cos terms of one role (see oracle/threebody_moments.py::backward_atom): pa / pb = coefficient 9-vectors of the
// bond in this role / the other role, mom = moments of the partners' coefficients (momq: the same side in shared
// memory, for the second-order block Q).  Adds to G (vector) and X (scalar).
__device__ __forceinline__ void role_terms(const Mom& mom, const float* momq, const float* u, const float* o6,
                                           const float* pa, const float* pb, float* G, float& X) {
  // l = 1 (linear in the upstream gradient)
  const float d11 = pa[3] * pb[3] + pa[4] * pb[4] + pa[5] * pb[5];
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    const float* v = mom.V1 + 3 * n;
    g0 = fmaf(pa[3 + n], v[0], g0); g1 = fmaf(pa[3 + n], v[1], g1); g2 = fmaf(pa[3 + n], v[2], g2);
  }
  float x = g0 * u[0] + g1 * u[1] + g2 * u[2];
  G[0] += kY1 * (g0 - d11 * u[0]); G[1] += kY1 * (g1 - d11 * u[1]); G[2] += kY1 * (g2 - d11 * u[2]);
  X += kY1 * (x - d11);
  // l = 2, term 2 x go
  const float d22 = pa[6] * pb[6] + pa[7] * pb[7] + pa[8] * pb[8];
  float W[6];
#pragma unroll
  for (int s = 0; s < 6; ++s)
    W[s] = pa[6] * mom.M2[s] + pa[7] * mom.M2[6 + s] + pa[8] * mom.M2[12 + s];
  float wu[3];
  matvec6(W, u, wu);
  x = wu[0] * u[0] + wu[1] * u[1] + wu[2] * u[2];
  G[0] += 2.0f * kY2 * (wu[0] - d22 * u[0]); G[1] += 2.0f * kY2 * (wu[1] - d22 * u[1]);
  G[2] += 2.0f * kY2 * (wu[2] - d22 * u[2]);
  X += 2.0f * kY2 * (x - d22);
  // l = 2, quirk term x go^2: sum_{n,m} pa_n pa_m Q[n][m]
#pragma unroll
  for (int s = 0; s < 6; ++s) W[s] = 0.0f;
  const float4* q4 = reinterpret_cast<const float4*>(momq + M_Q);
#pragma unroll
  for (int n = 0; n < 3; ++n)
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const float pp = pa[6 + n] * pa[6 + m];
      const float4 qa = q4[2 * (n * 3 + m)], qb = q4[2 * (n * 3 + m) + 1];
      W[0] = fmaf(pp, qa.x, W[0]); W[1] = fmaf(pp, qa.y, W[1]); W[2] = fmaf(pp, qa.z, W[2]);
      W[3] = fmaf(pp, qa.w, W[3]); W[4] = fmaf(pp, qb.x, W[4]); W[5] = fmaf(pp, qb.y, W[5]);
    }
  matvec6(W, u, wu);
  x = wu[0] * u[0] + wu[1] * u[1] + wu[2] * u[2];
  const float dd = d22 * d22;
  G[0] += kY2 * kY2 * (wu[0] - dd * u[0]); G[1] += kY2 * kY2 * (wu[1] - dd * u[1]);
  G[2] += kY2 * kY2 * (wu[2] - dd * u[2]);
  X += kY2 * kY2 * (x - dd);
  (void)o6;
}

// member bonds of one atom -> shared-memory entries.  b = G sigma[dst]; WITH_RED also stages the saved reduced features
template <bool WITH_RED>
__device__ __forceinline__ int stage(float* ent, int cap, int beg, int end, const float4* __restrict__ vec4,
                                     const float* __restrict__ G, const float* __restrict__ sig,
                                     const int32_t* __restrict__ dst, const float* __restrict__ red,
                                     const int32_t* __restrict__ tri_ptr, float r3, int lane) {
  int n3 = 0;
  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    bool m = false;
    if (e < end) m = __ldg(tri_ptr + e + 1) > __ldg(tri_ptr + e);
    const unsigned bal = __ballot_sync(FULL, m);
    const int pos = n3 + __popc(bal & ((1u << lane) - 1u));
    if (m && pos < cap) {
      const float4 v = __ldg(vec4 + e);
      const float* g = G + (int64_t)e * MD;
      const float* sg = sig + (int64_t)__ldg(dst + e) * MD;
      float b[MD];
#pragma unroll
      for (int d = 0; d < MD; ++d) b[d] = __ldg(g + d) * __ldg(sg + d);
      const float ir = 1.0f / v.w;
      const float ux = v.x * ir, uy = v.y * ir, uz = v.z * ir;
      float4* o = reinterpret_cast<float4*>(ent + pos * ES);
      o[0] = make_float4(ux, uy, uz, 1.0f);
      o[1] = make_float4(ux * ux, uy * uy, uz * uz, ux * uy);
      o[2] = make_float4(ux * uz, uy * uz, cutoff_poly(v.w, r3), v.w);
      o[3] = make_float4(b[0], b[1], b[2], __int_as_float(e));
      o[4] = make_float4(b[3], b[4], b[5], 0.0f);
      o[5] = make_float4(b[6], b[7], b[8], 0.0f);
      if (WITH_RED) {
        const float* q = red + (int64_t)e * MD;
        float qq[MD];
#pragma unroll
        for (int d = 0; d < MD; ++d) qq[d] = __ldg(q + d);
        const float cw = cutoff_poly(v.w, r3);
        o[6] = make_float4(qq[0], qq[1], qq[2], cw);
        o[7] = make_float4(qq[3], qq[4], qq[5], cw);
        o[8] = make_float4(qq[6], qq[7], qq[8], cw);
      }
    }
    n3 += __popc(bal);
  }
  __syncwarp();
  return n3;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// forward: red (member bonds) and e_out = e_in + SiLU(red WdT) * sigmoid(red WgT) (all bonds)
__global__ void __launch_bounds__(32 * MW, 4) tb_mom_fwd_kernel(
    const float4* __restrict__ vec4, const float* __restrict__ G, const float* __restrict__ sig,
    const int32_t* __restrict__ dst, const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ tri_ptr,
    float r3, const float* __restrict__ WdT, const float* __restrict__ WgT, const float* __restrict__ e_in, int64_t N,
    int cap, float* __restrict__ red, float* __restrict__ e_out) {
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ent = smem_f + (size_t)warp * (cap * ES + M_SIDE);
  float* mom = ent + cap * ES;
  const LaneRole role = lane_role(lane);
  // this lane's two feature columns of the 9 -> 64 gated MLP
  float2 wd[MD], wg[MD];
#pragma unroll
  for (int d = 0; d < MD; ++d) {
    wd[d] = __ldg(reinterpret_cast<const float2*>(WdT + d * MF) + lane);
    wg[d] = __ldg(reinterpret_cast<const float2*>(WgT + d * MF) + lane);
  }
  for (int64_t atom = (int64_t)blockIdx.x * MW + warp; atom < N; atom += (int64_t)gridDim.x * MW) {
    const int beg = __ldg(edge_ptr + atom), end = __ldg(edge_ptr + atom + 1);
    constexpr int RB = 8;
    float2 ra[RB], rb[RB];
    auto load_rows = [&](float2* row, int e0) {
#pragma unroll
      for (int i = 0; i < RB; ++i)
        if (e0 + i < end) row[i] = __ldg(reinterpret_cast<const float2*>(e_in + (int64_t)(e0 + i) * MF) + lane);
    };
    load_rows(ra, beg);
    load_rows(rb, beg + RB);
    const int n3 = stage<false>(ent, cap, beg, end, vec4, G, sig, dst, nullptr, tri_ptr, r3, lane);
    accumulate_side<false, false>(ent, n3, O_B, role, mom);
    __syncwarp();
    const Mom mb = load_moments(mom);
    for (int j = lane; j < n3; j += 32) {
      float* en = ent + j * ES;
      const float4 u4 = *reinterpret_cast<const float4*>(en + O_U);
      const float4 oa = *reinterpret_cast<const float4*>(en + O_O6), ob = *reinterpret_cast<const float4*>(en + 8);
      const float4 b0 = *reinterpret_cast<const float4*>(en + O_B), b1 = *reinterpret_cast<const float4*>(en + O_B + 4),
                   b2 = *reinterpret_cast<const float4*>(en + O_B + 8);
      const float u[3] = {u4.x, u4.y, u4.z};
      const float o6[6] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y};
      const float own[MD] = {b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, b2.x, b2.y, b2.z};
      float acc[MD];
      eval_rows(mb, u, o6, own, acc);
      const float c = ob.z;
      const int e1 = __float_as_int(b0.w);
      float* ro = red + (int64_t)e1 * MD;
#pragma unroll
      for (int d = 0; d < MD; ++d) { acc[d] *= c; ro[d] = acc[d]; }
      // the b slots now hold the reduced features for the MLP phase
      *reinterpret_cast<float4*>(en + O_B) = make_float4(acc[0], acc[1], acc[2], b0.w);
      *reinterpret_cast<float4*>(en + O_B + 4) = make_float4(acc[3], acc[4], acc[5], 0.0f);
      *reinterpret_cast<float4*>(en + O_B + 8) = make_float4(acc[6], acc[7], acc[8], 0.0f);
    }
    __syncwarp();
    // ---- edge rows, lane owns features 2*lane, 2*lane+1; two register buffers of RB rows ----
    int pos = 0;
    auto process_rows = [&](float2* row, int e0) {
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int e = e0 + i;
        if (e >= end) break;
        if (pos < n3 && __float_as_int(ent[pos * ES + O_E]) == e) {
          const float* en = ent + pos * ES + O_B;
          const float4 q0 = *reinterpret_cast<const float4*>(en), q1 = *reinterpret_cast<const float4*>(en + 4),
                       q2 = *reinterpret_cast<const float4*>(en + 8);
          const float rd[MD] = {q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z};
          float2 u2 = make_float2(0.f, 0.f), g2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int d = 0; d < MD; ++d) {
            u2 = fma2s(rd[d], wd[d], u2);
            g2 = fma2s(rd[d], wg[d], g2);
          }
          row[i].x += silu_m(u2.x) * sigmoid_m(g2.x);
          row[i].y += silu_m(u2.y) * sigmoid_m(g2.y);
          ++pos;
        }
        reinterpret_cast<float2*>(e_out + (int64_t)e * MF)[lane] = row[i];
      }
    };
    for (int e0 = beg; e0 < end; e0 += 2 * RB) {
      process_rows(ra, e0);
      load_rows(ra, e0 + 2 * RB);
      process_rows(rb, e0 + RB);
      load_rows(rb, e0 + 3 * RB);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// The forward split in two (default): the per-atom part only produces red (no weights, no edge rows: few registers,
// many warps per SM to hide the staging latency) ...
__global__ void __launch_bounds__(32 * MW, 8) tb_mom_red_kernel(
    const float4* __restrict__ vec4, const float* __restrict__ G, const float* __restrict__ sig,
    const int32_t* __restrict__ dst, const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ tri_ptr,
    float r3, int64_t N, int cap, float* __restrict__ red) {
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ent = smem_f + (size_t)warp * (cap * ES + M_SIDE);
  float* mom = ent + cap * ES;
  const LaneRole role = lane_role(lane);
  for (int64_t atom = (int64_t)blockIdx.x * MW + warp; atom < N; atom += (int64_t)gridDim.x * MW) {
    const int beg = __ldg(edge_ptr + atom), end = __ldg(edge_ptr + atom + 1);
    const int n3 = stage<false>(ent, cap, beg, end, vec4, G, sig, dst, nullptr, tri_ptr, r3, lane);
    accumulate_side<false, false>(ent, n3, O_B, role, mom);
    __syncwarp();
    const Mom mb = load_moments(mom);
    for (int j = lane; j < n3; j += 32) {
      const float* en = ent + j * ES;
      const float4 u4 = *reinterpret_cast<const float4*>(en + O_U);
      const float4 oa = *reinterpret_cast<const float4*>(en + O_O6), ob = *reinterpret_cast<const float4*>(en + 8);
      const float4 b0 = *reinterpret_cast<const float4*>(en + O_B), b1 = *reinterpret_cast<const float4*>(en + O_B + 4),
                   b2 = *reinterpret_cast<const float4*>(en + O_B + 8);
      const float u[3] = {u4.x, u4.y, u4.z};
      const float o6[6] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y};
      const float own[MD] = {b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, b2.x, b2.y, b2.z};
      float acc[MD];
      eval_rows(mb, u, o6, own, acc);
      float* ro = red + (int64_t)__float_as_int(b0.w) * MD;
#pragma unroll
      for (int d = 0; d < MD; ++d) ro[d] = ob.z * acc[d];
    }
    __syncwarp();
  }
}

// ... and the edge update e_out = e_in + SiLU(red WdT) * sigmoid(red WgT) streams ALL bond rows, independent of the
// atom structure: a warp takes 32 consecutive rows at a time (member flags by ballot, the chunk's red rows staged in
// shared memory for broadcast reads, 2 x 8 rows of e in flight per warp), lane = two feature columns (FFMA2).  This half
// is bound by the 512 B per bond it has to move.
// FROM_H (first block of the model, whole-step executor): the incoming rows are the EdgeAdjustor's output
// e0 = SiLU(h Wa^T) (nn/featurizer.py:84-96), formed here from the 12-byte h row (same operation order as
// edge_adjust_fwd4_kernel: bit-identical) instead of being written by one kernel and read back by this one.
constexpr int UW = 8;  // warps per CTA
template <bool FROM_H>
__global__ void __launch_bounds__(32 * UW, 2) tb_edge_update_kernel(
    const float* __restrict__ red, const int32_t* __restrict__ tri_ptr, const float* __restrict__ WdT,
    const float* __restrict__ WgT, const float* __restrict__ e_in, const float* __restrict__ h,
    const float* __restrict__ WaT, int64_t E, float* __restrict__ e_out) {
  __shared__ __align__(16) float red_s[UW][32][12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 wa[3];
  if (FROM_H) {
#pragma unroll
    for (int m = 0; m < 3; ++m) wa[m] = __ldg(reinterpret_cast<const float2*>(WaT + m * MF) + lane);
  }
  float2 wd[MD], wg[MD];
#pragma unroll
  for (int d = 0; d < MD; ++d) {
    wd[d] = __ldg(reinterpret_cast<const float2*>(WdT + d * MF) + lane);
    wg[d] = __ldg(reinterpret_cast<const float2*>(WgT + d * MF) + lane);
  }
  const int64_t n_chunks = (E + 31) >> 5;
  const int64_t stride = (int64_t)gridDim.x * UW;
  constexpr int RB = 8;
  for (int64_t chunk = (int64_t)blockIdx.x * UW + warp; chunk < n_chunks; chunk += stride) {
    const int64_t e0 = chunk << 5;
    const int rows = (int)min((int64_t)32, E - e0);
    float2 ra[RB], rb[RB];
    float hv[3] = {0.f, 0.f, 0.f};
    if (FROM_H && lane < rows) {
#pragma unroll
      for (int m = 0; m < 3; ++m) hv[m] = __ldg(h + (e0 + lane) * 3 + m);
    }
    auto load_rows = [&](float2* row, int r0) {
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        if (FROM_H) {
          float2 z = make_float2(0.f, 0.f);
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            const float hm = __shfl_sync(FULL, hv[m], r0 + i);
            z.x = fmaf(hm, wa[m].x, z.x);
            z.y = fmaf(hm, wa[m].y, z.y);
          }
          row[i] = make_float2(z.x * sigmoid_m(z.x), z.y * sigmoid_m(z.y));
        } else if (r0 + i < rows) {
          row[i] = __ldg(reinterpret_cast<const float2*>(e_in + (e0 + r0 + i) * MF) + lane);
        }
      }
    };
    load_rows(ra, 0);
    load_rows(rb, RB);
    bool member = false;
    if (lane < rows) member = __ldg(tri_ptr + e0 + lane + 1) > __ldg(tri_ptr + e0 + lane);
    const unsigned mask = __ballot_sync(FULL, member);
    if (member) {
      const float* q = red + (e0 + lane) * MD;
      float qq[MD];
#pragma unroll
      for (int d = 0; d < MD; ++d) qq[d] = __ldg(q + d);
      float4* o = reinterpret_cast<float4*>(&red_s[warp][lane][0]);
      o[0] = make_float4(qq[0], qq[1], qq[2], qq[3]);
      o[1] = make_float4(qq[4], qq[5], qq[6], qq[7]);
      o[2] = make_float4(qq[8], 0.f, 0.f, 0.f);
    }
    __syncwarp();
    auto process_rows = [&](float2* row, int r0) {
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int r = r0 + i;
        if (r >= rows) break;
        if (mask & (1u << r)) {
          const float4* rs = reinterpret_cast<const float4*>(&red_s[warp][r][0]);
          const float4 q0 = rs[0], q1 = rs[1], q2 = rs[2];
          const float rd[MD] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
          float2 u2 = make_float2(0.f, 0.f), g2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int d = 0; d < MD; ++d) {
            u2 = fma2s(rd[d], wd[d], u2);
            g2 = fma2s(rd[d], wg[d], g2);
          }
          // SiLU(u) sigmoid(g) = u / ((1 + e^-u)(1 + e^-g)): one reciprocal for both sigmoids (3 MUFU results per
          // element instead of 4; the kernel sits at ~50 % of the MUFU pipe at C5)
          row[i].x += gated_m(u2.x, g2.x);
          row[i].y += gated_m(u2.y, g2.y);
        }
        reinterpret_cast<float2*>(e_out + (e0 + r) * MF)[lane] = row[i];
      }
    };
    process_rows(ra, 0);
    load_rows(ra, 2 * RB);
    process_rows(rb, RB);
    load_rows(rb, 3 * RB);
    process_rows(ra, 2 * RB);
    process_rows(rb, 3 * RB);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// backward: g_vec4 (E,4) = d/d(v, r) (zeros for non-member bonds; includes the fc' and the radial-basis chain through
// dG) and g_sig_e (E,9) = per-bond gradient of sigma[dst] (zeros for non-member bonds)
// WITH_MLP = false (default, m3g_tb_mom_bwd_q): `red` already holds q = dL/dred (m3g_tb_mlp_adj, below) and the kernel
// carries no weights: fewer registers, more warps per SM for the staging latency.
template <bool WITH_MLP>
__global__ void __launch_bounds__(32 * MW, WITH_MLP ? 4 : 5) tb_mom_bwd_kernel(
    const float4* __restrict__ vec4, const float* __restrict__ G, const float* __restrict__ dG,
    const float* __restrict__ sig, const int32_t* __restrict__ dst, const float* __restrict__ red,
    const float* __restrict__ g_e, const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ tri_ptr, float r3,
    const float* __restrict__ WdT, const float* __restrict__ WgT, int64_t N, int cap, int accumulate,
    float4* __restrict__ g_vec4, float* __restrict__ g_sig_e) {
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ent = smem_f + (size_t)warp * (cap * ES + 2 * M_SIDE);
  float* mom = ent + cap * ES;
  const LaneRole role = lane_role(lane);
  float2 wd[MD], wg[MD];
  if (WITH_MLP) {
#pragma unroll
    for (int d = 0; d < MD; ++d) {
      wd[d] = __ldg(reinterpret_cast<const float2*>(WdT + d * MF) + lane);
      wg[d] = __ldg(reinterpret_cast<const float2*>(WgT + d * MF) + lane);
    }
  }
  // reduce9 target of this lane: component 5 b4 + 3 b3 + 2 b2 + b1 (bits of the lane id), written by even lanes
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
  const int dsel = 5 * b4 + 3 * b3 + 2 * b2 + b1;
  const bool writer = !(lane & 1) && !(b3 & b2) && !(b2 & b1) && dsel < MD;
  const int wslot = O_A + (dsel / 3) * 4 + (dsel % 3);
  for (int64_t atom = (int64_t)blockIdx.x * MW + warp; atom < N; atom += (int64_t)gridDim.x * MW) {
    const int beg = __ldg(edge_ptr + atom), end = __ldg(edge_ptr + atom + 1);
    // non-member bonds carry no three-body term: their rows are zeroed by the first call of a step (accumulate == 0);
    // later calls (the other blocks of the model) write member rows only, so the zeros stay
    if (!accumulate)
      for (int e = beg + lane; e < end; e += 32)
        if (!(__ldg(tri_ptr + e + 1) > __ldg(tri_ptr + e))) {
          g_vec4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
          float* gs = g_sig_e + (int64_t)e * MD;
#pragma unroll
          for (int d = 0; d < MD; ++d) gs[d] = 0.0f;
        }
    const int n3 = stage<true>(ent, cap, beg, end, vec4, G, sig, dst, red, tri_ptr, r3, lane);
    // ---- gated-MLP adjoint: q slots hold red -> become q = dL/dred ; lane owns features 2*lane, 2*lane+1 ----
    constexpr int GB = 4;
    float2 ga[GB], gb2[GB];
    auto load_g = [&](float2* ge, int p0) {
#pragma unroll
      for (int i = 0; i < GB; ++i)
        if (p0 + i < n3)
          ge[i] = __ldg(reinterpret_cast<const float2*>(g_e + (int64_t)__float_as_int(ent[(p0 + i) * ES + O_E]) * MF) +
                        lane);
    };
    auto process_g = [&](const float2* ge, int p0) {
#pragma unroll
      for (int i = 0; i < GB; ++i) {
        const int pos = p0 + i;
        if (pos >= n3) break;
        float* en = ent + pos * ES;
        const float4 q0 = *reinterpret_cast<const float4*>(en + O_A), q1 = *reinterpret_cast<const float4*>(en + O_A + 4),
                     q2 = *reinterpret_cast<const float4*>(en + O_A + 8);
        const float rd[MD] = {q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z};
        float2 u2 = make_float2(0.f, 0.f), g2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int d = 0; d < MD; ++d) {
          u2 = fma2s(rd[d], wd[d], u2);
          g2 = fma2s(rd[d], wg[d], g2);
        }
        const float sg0 = sigmoid_m(g2.x), sg1 = sigmoid_m(g2.y), su0 = sigmoid_m(u2.x), su1 = sigmoid_m(u2.y);
        float2 du, dg;
        du.x = ge[i].x * sg0 * su0 * (1.0f + u2.x * (1.0f - su0));
        du.y = ge[i].y * sg1 * su1 * (1.0f + u2.y * (1.0f - su1));
        dg.x = ge[i].x * (u2.x * su0) * sg0 * (1.0f - sg0);
        dg.y = ge[i].y * (u2.y * su1) * sg1 * (1.0f - sg1);
        float part[MD];
#pragma unroll
        for (int d = 0; d < MD; ++d) {
          const float2 t = fma2(dg, wg[d], mul2(du, wd[d]));
          part[d] = t.x + t.y;
        }
        // 9 sums over the 32 lanes with a value-halving butterfly (12 shuffles; fixed tree -> fixed order)
        const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
        float a5[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float keep = h4 ? part[5 + k] : part[k], send = h4 ? part[k] : part[5 + k];
          a5[k] = keep + __shfl_xor_sync(FULL, send, 16);
        }
        a5[4] = (h4 ? 0.0f : part[4]) + __shfl_xor_sync(FULL, h4 ? part[4] : 0.0f, 16);
        float c3[3];
        c3[0] = (h3 ? a5[3] : a5[0]) + __shfl_xor_sync(FULL, h3 ? a5[0] : a5[3], 8);
        c3[1] = (h3 ? a5[4] : a5[1]) + __shfl_xor_sync(FULL, h3 ? a5[1] : a5[4], 8);
        c3[2] = (h3 ? 0.0f : a5[2]) + __shfl_xor_sync(FULL, h3 ? a5[2] : 0.0f, 8);
        float d2[2];
        d2[0] = (h2 ? c3[2] : c3[0]) + __shfl_xor_sync(FULL, h2 ? c3[0] : c3[2], 4);
        d2[1] = (h2 ? 0.0f : c3[1]) + __shfl_xor_sync(FULL, h2 ? c3[1] : 0.0f, 4);
        float tot = (h1 ? d2[1] : d2[0]) + __shfl_xor_sync(FULL, h1 ? d2[0] : d2[1], 2);
        tot += __shfl_xor_sync(FULL, tot, 1);
        __syncwarp();
        if (writer) en[wslot] = tot;
      }
    };
    if (WITH_MLP) {
      load_g(ga, 0);
      load_g(gb2, GB);
      for (int p0 = 0; p0 < n3; p0 += 2 * GB) {
        process_g(ga, p0);
        load_g(ga, p0 + 2 * GB);
        process_g(gb2, p0 + GB);
        load_g(gb2, p0 + 3 * GB);
      }
      __syncwarp();
    }
    // ---- moments of b and of a (members ascending) ----
    accumulate_side<true, false>(ent, n3, O_B, role, mom);
    accumulate_side<true, true>(ent, n3, O_A, role, mom + M_SIDE);
    __syncwarp();
    // ---- one lane per member bond: both roles ----
    for (int j = lane; j < n3; j += 32) {
      const float* en = ent + j * ES;
      const float4 u4 = *reinterpret_cast<const float4*>(en + O_U);
      const float4 oa = *reinterpret_cast<const float4*>(en + O_O6), ob = *reinterpret_cast<const float4*>(en + 8);
      const float4 b0 = *reinterpret_cast<const float4*>(en + O_B), b1 = *reinterpret_cast<const float4*>(en + O_B + 4),
                   b2v = *reinterpret_cast<const float4*>(en + O_B + 8);
      const float4 a0 = *reinterpret_cast<const float4*>(en + O_A), a1 = *reinterpret_cast<const float4*>(en + O_A + 4),
                   a2 = *reinterpret_cast<const float4*>(en + O_A + 8);
      const float u[3] = {u4.x, u4.y, u4.z};
      const float o6[6] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y};
      const float bj[MD] = {b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, b2v.x, b2v.y, b2v.z};
      const float c = ob.z, r = ob.w;
      const float qj[MD] = {a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, a2.x, a2.y, a2.z};
      float aj[MD];
#pragma unroll
      for (int d = 0; d < MD; ++d) aj[d] = c * qj[d];
      const int e = __float_as_int(b0.w);
      float gB[MD];
      float gcq = 0.0f;
      float Gv[3] = {0.f, 0.f, 0.f}, X = 0.0f;
      {  // moments of b: forward inner sums (for d/dc_j) ; j as first bond
        const Mom mb = load_moments(mom);
        float acc[MD];
        eval_rows(mb, u, o6, bj, acc);
#pragma unroll
        for (int d = 0; d < MD; ++d) gcq = fmaf(qj[d], acc[d], gcq);
        role_terms(mb, mom, u, o6, aj, bj, Gv, X);
      }
      {  // moments of a: dL/db_j ; j as second bond
        const Mom ma = load_moments(mom + M_SIDE);
        eval_rows(ma, u, o6, aj, gB);
        role_terms(ma, mom + M_SIDE, u, o6, bj, aj, Gv, X);
      }
      // chain to sigma[dst] and to r through b = G sigma
      const float* gr_ = G + (int64_t)e * MD;
      const float* dg_ = dG + (int64_t)e * MD;
      const float* sg_ = sig + (int64_t)__ldg(dst + e) * MD;
      float* gs = g_sig_e + (int64_t)e * MD;
      float grad_r = 0.0f;
#pragma unroll
      for (int d = 0; d < MD; ++d) {
        gs[d] = gB[d] * __ldg(gr_ + d);
        grad_r = fmaf(gB[d] * __ldg(sg_ + d), __ldg(dg_ + d), grad_r);
      }
      const float ir = 1.0f / r;
      float4 gout = make_float4(Gv[0] * ir, Gv[1] * ir, Gv[2] * ir, (grad_r + gcq * cutoff_poly_grad(r, r3)) - X * ir);
      if (accumulate) {  // sum over the blocks of the model: this bond's row is owned by this lane
        const float4 old = g_vec4[e];
        gout.x += old.x; gout.y += old.y; gout.z += old.z; gout.w += old.w;
      }
      g_vec4[e] = gout;
    }
    __syncwarp();
  }
}


// ------------------------------------------------------------------------------------------------
// The backward split in two (default), like the forward.  First half: the adjoint of the 9 -> 64 gated MLP over the
// PACKED list of member bonds, q = dL/dred (E,9).  In the fused kernel above a lane owns two feature columns, so every
// row pays a 12-shuffle butterfly for its nine sums (~130 issued instructions per row).  Here a LANE OWNS A ROW: its g_e
// row comes out of a coalesced shared-memory staging tile, the weights are warp-uniform broadcast reads and the nine sums
// stay in registers (no shuffles, ~70 instructions per row).  q may alias red (a lane reads its row before it writes it).
// Measured alternatives (C5, 2.2 M member rows; tools/pipe_rate.py for the instruction rates): scalar FFMA instead of
// packed FFMA2 0.373 vs 0.340 ms (FFMA2 is slower per FMA on B200, 2.2-3.0 vs 1.0-1.3 cycles per warp instruction for
// two vs one FMA per lane, but halves the issue slots); weights as a kernel parameter read through uniform registers
// from the constant bank (no shared-memory wavefronts, 80 registers) 0.336 ms; cp.async double buffering of the staging
// tile (kept) 0.340 -> 0.304 ms.
constexpr int AW = 8;    // warps per CTA

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed math (FFMA2 / FMUL2, two feature columns per instruction) and the g_e rows streamed through a DOUBLE-BUFFERED
// staging tile with cp.async: a warp computes features [0,32) of its 32 rows out of one half-tile while
// the other half (features [32,64), then the next chunk's first half) is in flight, so the global-load latency that
// took 30 % of the stall samples of the single-buffer version (ncu) is hidden.
// Accumulation order (stated): features ascending in two interleaved partial sums (even / odd columns), added at the end.
constexpr int GSH = 36;  // half-tile row stride in floats (32 features + 4: 16-byte aligned, conflict-free LDS.128)

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(32 * AW, 2) tb_mlp_adj_kernel(
    const float* red, const float* __restrict__ g_e, const int32_t* __restrict__ members, int64_t n_members,
    const float* __restrict__ WdT, const float* __restrict__ WgT, float* q) {
  extern __shared__ __align__(16) float smem_f[];
  // weight table [column pair p][d]: (wd[d][2p], wd[d][2p+1], wg[d][2p], wg[d][2p+1])
  float4* wtab = reinterpret_cast<float4*>(smem_f);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf0 = smem_f + 32 * MD * 4 + (size_t)warp * (2 * 32 * GSH);
  float* buf1 = buf0 + 32 * GSH;
  for (int i = threadIdx.x; i < 32 * MD; i += 32 * AW) {
    const int p = i / MD, d = i - p * MD;
    const float2 a = __ldg(reinterpret_cast<const float2*>(WdT + d * MF) + p);
    const float2 b = __ldg(reinterpret_cast<const float2*>(WgT + d * MF) + p);
    wtab[i] = make_float4(a.x, a.y, b.x, b.y);
  }
  __syncthreads();
  const int64_t n_chunks = (n_members + 31) >> 5;
  const int64_t stride = (int64_t)gridDim.x * AW;
  const int rsub = lane >> 3, c8 = lane & 7;  // copy role: row 4 i + rsub, 16-byte piece c8 of the 128-byte half row
  const float2 one2 = make_float2(1.0f, 1.0f), mone2 = make_float2(-1.0f, -1.0f);
  const float2 nl2e = make_float2(-1.4426950408889634f, -1.4426950408889634f);
  auto load_e = [&](int64_t chunk) -> int {
    const int64_t m = (chunk << 5) + lane;
    return (chunk < n_chunks && m < n_members) ? __ldg(members + m) : -1;
  };
  auto issue_copy = [&](float* buf, int e_lane, int half) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int er = __shfl_sync(FULL, e_lane, 4 * i + rsub);
      if (er >= 0) cp_async16(buf + (4 * i + rsub) * GSH + 4 * c8, g_e + (int64_t)er * MF + 32 * half + 4 * c8);
    }
    cp_async_commit();
  };
  int64_t chunk = (int64_t)blockIdx.x * AW + warp;
  int my_e = load_e(chunk);
  issue_copy(buf0, my_e, 0);
  for (; chunk < n_chunks; chunk += stride) {
    const int next_e = load_e(chunk + stride);
    issue_copy(buf1, my_e, 1);
    float2 rd2[MD];
#pragma unroll
    for (int d = 0; d < MD; ++d) {
      const float r = (my_e >= 0) ? red[(int64_t)my_e * MD + d] : 0.0f;
      rd2[d] = make_float2(r, r);
    }
    float2 acc[MD];
#pragma unroll
    for (int d = 0; d < MD; ++d) acc[d] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      cp_async_wait<1>();  // everything but the newest group has landed: this half-tile
      __syncwarp();
      const float* grow = (half ? buf1 : buf0) + lane * GSH;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const float4 g4 = *reinterpret_cast<const float4*>(grow + 4 * c);
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          const float2 ge = hp ? make_float2(g4.z, g4.w) : make_float2(g4.x, g4.y);
          const float4* wt = wtab + (16 * half + 2 * c + hp) * MD;
          float4 w[MD];
          float2 zd = make_float2(0.f, 0.f), zg = make_float2(0.f, 0.f);
#pragma unroll
          for (int d = 0; d < MD; ++d) {
            w[d] = wt[d];
            zd = fma2(rd2[d], make_float2(w[d].x, w[d].y), zd);
            zg = fma2(rd2[d], make_float2(w[d].z, w[d].w), zg);
          }
          // su = sigmoid(zd), sg = sigmoid(zg) ; du = ge sg su (1 + zd (1 - su)) ; dg = ge (zd su) sg (1 - sg)
          const float2 nd = mul2(zd, nl2e), ng = mul2(zg, nl2e);
          const float2 dd = add2(make_float2(ex2_fast(nd.x), ex2_fast(nd.y)), one2);
          const float2 dq = add2(make_float2(ex2_fast(ng.x), ex2_fast(ng.y)), one2);
          const float2 su = make_float2(rcp_fast(dd.x), rcp_fast(dd.y));
          const float2 sg = make_float2(rcp_fast(dq.x), rcp_fast(dq.y));
          const float2 a = mul2(ge, mul2(su, sg));
          const float2 du = mul2(a, fma2(zd, fma2(su, mone2, one2), one2));
          const float2 dg = mul2(mul2(a, zd), fma2(sg, mone2, one2));
#pragma unroll
          for (int d = 0; d < MD; ++d) {
            acc[d] = fma2(du, make_float2(w[d].x, w[d].y), acc[d]);
            acc[d] = fma2(dg, make_float2(w[d].z, w[d].w), acc[d]);
          }
        }
      }
      __syncwarp();  // every lane is done with this half-tile
      if (half == 0) issue_copy(buf0, next_e, 0);  // next chunk's first half (an empty group past the end)
    }
    if (my_e >= 0) {
      float* qo = q + (int64_t)my_e * MD;
#pragma unroll
      for (int d = 0; d < MD; ++d) qo[d] = acc[d].x + acc[d].y;
    }
    my_e = next_e;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// sigma = sigmoid(x Ws^T + bs) for F = 64, D = 9 (nn/interaction.py:204-206) and its adjoint.  The generic kernels
// (threebody.cu) walk the 9 outputs one by one with a full warp reduction each and re-read Ws from global memory; here a
// lane keeps its two columns of all nine weight rows in registers, the nine sums share one value-halving butterfly, and
// a warp strides over atoms.
__device__ __forceinline__ float reduce9_lanes(const float* v, int lane) {
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
  float a[5];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h4 ? v[5 + i] : v[i], send = h4 ? v[i] : v[5 + i];
    a[i] = keep + __shfl_xor_sync(FULL, send, 16);
  }
  a[4] = (h4 ? 0.0f : v[4]) + __shfl_xor_sync(FULL, h4 ? v[4] : 0.0f, 16);
  float c[3];
  c[0] = (h3 ? a[3] : a[0]) + __shfl_xor_sync(FULL, h3 ? a[0] : a[3], 8);
  c[1] = (h3 ? a[4] : a[1]) + __shfl_xor_sync(FULL, h3 ? a[1] : a[4], 8);
  c[2] = (h3 ? 0.0f : a[2]) + __shfl_xor_sync(FULL, h3 ? a[2] : 0.0f, 8);
  float d[2];
  d[0] = (h2 ? c[2] : c[0]) + __shfl_xor_sync(FULL, h2 ? c[0] : c[2], 4);
  d[1] = (h2 ? 0.0f : c[1]) + __shfl_xor_sync(FULL, h2 ? c[1] : 0.0f, 4);
  const float e = (h1 ? d[1] : d[0]) + __shfl_xor_sync(FULL, h1 ? d[0] : d[1], 2);
  return e + __shfl_xor_sync(FULL, e, 1);  // lane l: total of component 5 b4 + 3 b3 + 2 b2 + b1 (bits of l)
}

__global__ void __launch_bounds__(256) tb_sigma64_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Ws,
                                                             const float* __restrict__ bs, int64_t N,
                                                             float* __restrict__ sig) {
  const int lane = threadIdx.x & 31;
  float2 w[MD];
#pragma unroll
  for (int d = 0; d < MD; ++d) w[d] = __ldg(reinterpret_cast<const float2*>(Ws + d * MF) + lane);
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
  const int dsel = 5 * b4 + 3 * b3 + 2 * b2 + b1;
  const bool writer = !(lane & 1) && !(b3 & b2) && !(b2 & b1) && dsel < MD;
  const float bias = writer ? __ldg(bs + dsel) : 0.0f;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < N; k += warps) {
    const float2 xv = __ldg(reinterpret_cast<const float2*>(x + k * MF) + lane);
    float part[MD];
#pragma unroll
    for (int d = 0; d < MD; ++d) part[d] = fmaf(xv.y, w[d].y, xv.x * w[d].x);
    const float tot = reduce9_lanes(part, lane);
    if (writer) sig[k * MD + dsel] = sigmoid_acc(tot + bias);
  }
}

// g_x (N,64) = base + (sum_{e in in(k)} g_sig_e[e]) * sig (1 - sig) . Ws
__global__ void __launch_bounds__(256) tb_sigma64_bwd_kernel(const float* __restrict__ g_sig_e,
                                                             const int32_t* __restrict__ in_ptr,
                                                             const int32_t* __restrict__ in_perm,
                                                             const float* __restrict__ sig, const float* __restrict__ Ws,
                                                             const float* __restrict__ base, int64_t N,
                                                             float* __restrict__ g_x) {
  const int lane = threadIdx.x & 31;
  float2 w[MD];
#pragma unroll
  for (int d = 0; d < MD; ++d) w[d] = __ldg(reinterpret_cast<const float2*>(Ws + d * MF) + lane);
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < N; k += warps) {
    const int pb = __ldg(in_ptr + k), pe = __ldg(in_ptr + k + 1);
    float acc[MD];
#pragma unroll
    for (int d = 0; d < MD; ++d) acc[d] = 0.0f;
    for (int p = pb + lane; p < pe; p += 32) {  // incoming bonds: ascending edge id per lane, then the fixed tree
      const float* row = g_sig_e + (int64_t)__ldg(in_perm + p) * MD;
#pragma unroll
      for (int d = 0; d < MD; ++d) acc[d] += __ldg(row + d);
    }
    const float tot = reduce9_lanes(acc, lane);
    // component d sits in lane 16 b4 + 8 b3 + 4 b2 + 2 b1 with d = 5 b4 + 3 b3 + 2 b2 + b1
    float t[MD];
    const int src_lane[MD] = {0, 2, 4, 8, 10, 16, 18, 20, 24};
#pragma unroll
    for (int d = 0; d < MD; ++d) {
      const float s = __shfl_sync(FULL, tot, src_lane[d]);
      const float sg = __ldg(sig + k * MD + d);
      t[d] = s * sg * (1.0f - sg);
    }
    float2 v = base ? __ldg(reinterpret_cast<const float2*>(base + k * MF) + lane) : make_float2(0.f, 0.f);
#pragma unroll
    for (int d = 0; d < MD; ++d) { v.x = fmaf(t[d], w[d].x, v.x); v.y = fmaf(t[d], w[d].y, v.y); }
    reinterpret_cast<float2*>(g_x + k * MF)[lane] = v;
  }
}

}  // namespace m3g

using namespace m3g;

static inline int pick_cap(int max_members) { return max_members <= 32 ? 32 : (max_members <= 64 ? 64 : 128); }

template <typename Kernel>
static inline int mom_launch_shape(Kernel kernel, size_t smem, int64_t N, int n_sm, bool max_shared, unsigned* grid) {
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    set_error("three-body moment kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    return M3G_ERR_CUDA;
  }
  // backward: all of the unified L1 / shared memory as shared memory (4 CTAs per SM; the entries hold the reused data);
  // the forward streams the edge rows and keeps the default split (measured: 0.416 vs 0.445 ms at C5)
  if (max_shared)
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * MW, smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const int64_t need = (N + MW - 1) / MW, capb = (int64_t)n_sm * per_sm;
  *grid = (unsigned)((need < capb) ? (need < 1 ? 1 : need) : capb);
  return M3G_OK;
}

extern "C" {

int m3g_tb_sigma64_fwd(const float* x, const float* Ws, const float* bs, int64_t N, int n_sm, float* sig, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(x && Ws && bs && sig, "m3g_tb_sigma64_fwd: null pointer");
  const int64_t need = (N + 7) / 8, cap = (int64_t)n_sm * 8;
  tb_sigma64_fwd_kernel<<<(unsigned)(need < cap ? need : cap), 256, 0, as_stream(stream)>>>(x, Ws, bs, N, sig);
  M3G_LAUNCH_CHECK("m3g_tb_sigma64_fwd");
  return M3G_OK;
}

int m3g_tb_sigma64_bwd(const float* g_sig_e, const int32_t* in_ptr, const int32_t* in_perm, const float* sig,
                       const float* Ws, const float* base, int64_t N, int n_sm, float* g_x, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(g_sig_e && in_ptr && in_perm && sig && Ws && g_x, "m3g_tb_sigma64_bwd: null pointer");
  const int64_t need = (N + 7) / 8, cap = (int64_t)n_sm * 8;
  tb_sigma64_bwd_kernel<<<(unsigned)(need < cap ? need : cap), 256, 0, as_stream(stream)>>>(g_sig_e, in_ptr, in_perm, sig,
                                                                                          Ws, base, N, g_x);
  M3G_LAUNCH_CHECK("m3g_tb_sigma64_bwd");
  return M3G_OK;
}

int m3g_tb_mom_capacity(void) { return 128; }

int m3g_tb_mom_fwd(const float* vec4, const float* G, const float* sig, const int32_t* dst, const int32_t* edge_ptr,
                   const int32_t* tri_ptr, float r3, const float* WdT, const float* WgT, const float* e_in, int64_t N,
                   int max_members, int n_sm, float* red, float* e_out, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && G && sig && dst && edge_ptr && tri_ptr && WdT && WgT && e_in && red && e_out,
              "m3g_tb_mom_fwd: null pointer");
  M3G_REQUIRE(max_members >= 0 && max_members <= 128, "m3g_tb_mom_fwd: %d member bonds per atom exceed the capacity 128",
              max_members);
  const int cap = pick_cap(max_members);
  const size_t smem = (size_t)MW * (cap * ES + M_SIDE) * sizeof(float);
  unsigned grid;
  int rc = mom_launch_shape(tb_mom_fwd_kernel, smem, N, n_sm, false, &grid);
  if (rc != M3G_OK) return rc;
  tb_mom_fwd_kernel<<<grid, 32 * MW, smem, as_stream(stream)>>>((const float4*)vec4, G, sig, dst, edge_ptr, tri_ptr, r3,
                                                               WdT, WgT, e_in, N, cap, red, e_out);
  M3G_LAUNCH_CHECK("m3g_tb_mom_fwd");
  return M3G_OK;
}

int m3g_tb_mom_red(const float* vec4, const float* G, const float* sig, const int32_t* dst, const int32_t* edge_ptr,
                   const int32_t* tri_ptr, float r3, int64_t N, int max_members, int n_sm, float* red, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && G && sig && dst && edge_ptr && tri_ptr && red, "m3g_tb_mom_red: null pointer");
  M3G_REQUIRE(max_members >= 0 && max_members <= 128, "m3g_tb_mom_red: %d member bonds per atom exceed the capacity 128",
              max_members);
  const int cap = pick_cap(max_members);
  const size_t smem = (size_t)MW * (cap * ES + M_SIDE) * sizeof(float);
  unsigned grid;
  int rc = mom_launch_shape(tb_mom_red_kernel, smem, N, n_sm, true, &grid);
  if (rc != M3G_OK) return rc;
  tb_mom_red_kernel<<<grid, 32 * MW, smem, as_stream(stream)>>>((const float4*)vec4, G, sig, dst, edge_ptr, tri_ptr, r3, N,
                                                               cap, red);
  M3G_LAUNCH_CHECK("m3g_tb_mom_red");
  return M3G_OK;
}

int m3g_tb_edge_update(const float* red, const int32_t* tri_ptr, const float* WdT, const float* WgT, const float* e_in,
                       int64_t E, int n_sm, float* e_out, void* stream) {
  if (E == 0) return M3G_OK;
  M3G_REQUIRE(red && tri_ptr && WdT && WgT && e_in && e_out, "m3g_tb_edge_update: null pointer");
  const int64_t need = ((E + 31) / 32 + UW - 1) / UW, capb = (int64_t)n_sm * 2;
  tb_edge_update_kernel<false><<<(unsigned)(need < capb ? need : capb), 32 * UW, 0, as_stream(stream)>>>(
      red, tri_ptr, WdT, WgT, e_in, nullptr, nullptr, E, e_out);
  M3G_LAUNCH_CHECK("m3g_tb_edge_update");
  return M3G_OK;
}

int m3g_tb_edge_update_h(const float* red, const int32_t* tri_ptr, const float* WdT, const float* WgT, const float* h,
                         const float* WaT, int64_t E, int n_sm, float* e_out, void* stream) {
  if (E == 0) return M3G_OK;
  M3G_REQUIRE(red && tri_ptr && WdT && WgT && h && WaT && e_out, "m3g_tb_edge_update_h: null pointer");
  const int64_t need = ((E + 31) / 32 + UW - 1) / UW, capb = (int64_t)n_sm * 2;
  tb_edge_update_kernel<true><<<(unsigned)(need < capb ? need : capb), 32 * UW, 0, as_stream(stream)>>>(
      red, tri_ptr, WdT, WgT, nullptr, h, WaT, E, e_out);
  M3G_LAUNCH_CHECK("m3g_tb_edge_update_h");
  return M3G_OK;
}

int m3g_tb_mlp_adj(const float* red, const float* g_e, const int32_t* member_edges, int64_t n_members, const float* WdT,
                   const float* WgT, int n_sm, float* q, void* stream) {
  if (n_members == 0) return M3G_OK;
  M3G_REQUIRE(red && g_e && member_edges && WdT && WgT && q, "m3g_tb_mlp_adj: null pointer");
  const size_t smem = (size_t)(32 * MD * 4 + AW * 2 * 32 * GSH) * sizeof(float);
  cudaError_t err = cudaFuncSetAttribute(tb_mlp_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    set_error("m3g_tb_mlp_adj: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    return M3G_ERR_CUDA;
  }
  const int64_t need = ((n_members + 31) / 32 + AW - 1) / AW, capb = (int64_t)n_sm * 2;
  tb_mlp_adj_kernel<<<(unsigned)(need < capb ? need : capb), 32 * AW, smem, as_stream(stream)>>>(
      red, g_e, member_edges, n_members, WdT, WgT, q);
  M3G_LAUNCH_CHECK("m3g_tb_mlp_adj");
  return M3G_OK;
}

int m3g_tb_mom_bwd_q(const float* vec4, const float* G, const float* dG, const float* sig, const int32_t* dst,
                     const float* q, const int32_t* edge_ptr, const int32_t* tri_ptr, float r3, int64_t N,
                     int max_members, int n_sm, int accumulate, float* g_vec4, float* g_sig_e, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && G && dG && sig && dst && q && edge_ptr && tri_ptr && g_vec4 && g_sig_e,
              "m3g_tb_mom_bwd_q: null pointer");
  M3G_REQUIRE(max_members >= 0 && max_members <= 128,
              "m3g_tb_mom_bwd_q: %d member bonds per atom exceed the capacity 128", max_members);
  const int cap = pick_cap(max_members);
  const size_t smem = (size_t)MW * (cap * ES + 2 * M_SIDE) * sizeof(float);
  unsigned grid;
  int rc = mom_launch_shape(tb_mom_bwd_kernel<false>, smem, N, n_sm, true, &grid);
  if (rc != M3G_OK) return rc;
  tb_mom_bwd_kernel<false><<<grid, 32 * MW, smem, as_stream(stream)>>>(
      (const float4*)vec4, G, dG, sig, dst, q, nullptr, edge_ptr, tri_ptr, r3, nullptr, nullptr, N, cap, accumulate,
      (float4*)g_vec4, g_sig_e);
  M3G_LAUNCH_CHECK("m3g_tb_mom_bwd_q");
  return M3G_OK;
}

int m3g_tb_mom_bwd(const float* vec4, const float* G, const float* dG, const float* sig, const int32_t* dst,
                   const float* red, const float* g_e, const int32_t* edge_ptr, const int32_t* tri_ptr, float r3,
                   const float* WdT, const float* WgT, int64_t N, int max_members, int n_sm, int accumulate,
                   float* g_vec4, float* g_sig_e, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(vec4 && G && dG && sig && dst && red && g_e && edge_ptr && tri_ptr && WdT && WgT && g_vec4 && g_sig_e,
              "m3g_tb_mom_bwd: null pointer");
  M3G_REQUIRE(max_members >= 0 && max_members <= 128, "m3g_tb_mom_bwd: %d member bonds per atom exceed the capacity 128",
              max_members);
  const int cap = pick_cap(max_members);
  const size_t smem = (size_t)MW * (cap * ES + 2 * M_SIDE) * sizeof(float);
  unsigned grid;
  int rc = mom_launch_shape(tb_mom_bwd_kernel<true>, smem, N, n_sm, true, &grid);
  if (rc != M3G_OK) return rc;
  tb_mom_bwd_kernel<true><<<grid, 32 * MW, smem, as_stream(stream)>>>((const float4*)vec4, G, dG, sig, dst, red, g_e, edge_ptr,
                                                               tri_ptr, r3, WdT, WgT, N, cap, accumulate, (float4*)g_vec4, g_sig_e);
  M3G_LAUNCH_CHECK("m3g_tb_mom_bwd");
  return M3G_OK;
}

}  // extern "C"
