// Stand-alone elementwise basis functions kept importable by the reference API
// (nn/interaction.py:284-400: spherical_bessel, legendre_cos, cutoff_function).  The model kernels in
// threebody.cu evaluate the same formulas inline; these entry points exist for the operator API and tests.
#include "common.cuh"

namespace m3g {

constexpr int MAX_ORDER = 9;

// j_order(x) and the reference's custom derivative (incl. the small-x branches, quirk Q4)
__global__ void sph_bessel_kernel(const float* __restrict__ x, int order, int64_t n, float* __restrict__ out,
                                  float* __restrict__ dout) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float EPS = 1e-8f;
  float v = x[i];
  bool big = v > EPS;
  float s, c;
  sincosf(v, &s, &c);
  float prev = big ? s / v : 1.0f;  // j_0
  float cur = prev, below = prev;
  if (order >= 1) {
    cur = big ? (s / v - c) / v : v / 3.0f;  // j_1
    below = prev;
    float coeff = 3.0f;
    for (int m = 1; m < order; ++m) {
      coeff *= (float)(2 * m + 3);
      float nxt = big ? ((float)(2 * m + 1) / v * cur - below) : v / coeff;
      below = cur;
      cur = nxt;
    }
  }
  out[i] = cur;
  if (dout) {
    float d;
    if (order == 0) d = big ? -((s / v - c) / v) : 0.0f;
    else if (order == 1) d = big ? (below - 2.0f / v * cur) : (1.0f / 3.0f);
    else d = big ? (below - (float)(order + 1) / v * cur) : 0.0f;
    dout[i] = d;
  }
}

__global__ void legendre_kernel(const float* __restrict__ x, int order, int64_t n, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  float below = 1.0f, cur = (order >= 1) ? v : 1.0f;
  for (int m = 1; m < order; ++m) {
    float nxt = ((float)(2 * m + 1) * v * cur - (float)m * below) / (float)(m + 1);
    below = cur;
    cur = nxt;
  }
  out[i] = cur;
}

// reference backward (quirk Q3): g = 0; for m = 1..order: g = (m P_{m-1} + x g) * go
__global__ void legendre_bwd_kernel(const float* __restrict__ x, const float* __restrict__ go, int order, int64_t n,
                                    float* __restrict__ gx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i], g0 = go[i];
  float P[MAX_ORDER + 1];
  P[0] = 1.0f;
  P[1] = v;
#pragma unroll
  for (int m = 1; m < MAX_ORDER; ++m) P[m + 1] = ((float)(2 * m + 1) * v * P[m] - (float)m * P[m - 1]) / (float)(m + 1);
  float g = 0.0f;
#pragma unroll
  for (int m = 1; m <= MAX_ORDER; ++m)
    if (m <= order) g = ((float)m * P[m - 1] + v * g) * g0;
  gx[i] = g;
}

__global__ void cutoff_kernel(const float* __restrict__ r, float rc, int64_t n, float* __restrict__ out,
                              float* __restrict__ dout) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = cutoff_poly(r[i], rc);
  if (dout) dout[i] = cutoff_poly_grad(r[i], rc);
}

}  // namespace m3g

using namespace m3g;

extern "C" {

int m3g_sph_bessel(const float* x, int order, int64_t n, float* out, float* dout, void* stream) {
  if (n == 0) return M3G_OK;
  M3G_REQUIRE(x && out, "m3g_sph_bessel: null pointer");
  M3G_REQUIRE(order >= 0 && order <= MAX_ORDER, "m3g_sph_bessel: order %d outside [0,%d]", order, MAX_ORDER);
  sph_bessel_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(x, order, n, out, dout);
  M3G_LAUNCH_CHECK("m3g_sph_bessel");
  return M3G_OK;
}

int m3g_legendre(const float* x, int order, int64_t n, float* out, void* stream) {
  if (n == 0) return M3G_OK;
  M3G_REQUIRE(x && out, "m3g_legendre: null pointer");
  M3G_REQUIRE(order >= 0 && order <= MAX_ORDER, "m3g_legendre: order %d outside [0,%d]", order, MAX_ORDER);
  legendre_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(x, order, n, out);
  M3G_LAUNCH_CHECK("m3g_legendre");
  return M3G_OK;
}

int m3g_legendre_bwd(const float* x, const float* go, int order, int64_t n, float* gx, void* stream) {
  if (n == 0) return M3G_OK;
  M3G_REQUIRE(x && go && gx, "m3g_legendre_bwd: null pointer");
  M3G_REQUIRE(order >= 0 && order <= MAX_ORDER, "m3g_legendre_bwd: order %d outside [0,%d]", order, MAX_ORDER);
  legendre_bwd_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(x, go, order, n, gx);
  M3G_LAUNCH_CHECK("m3g_legendre_bwd");
  return M3G_OK;
}

int m3g_cutoff(const float* r, float rc, int64_t n, float* out, float* dout, void* stream) {
  if (n == 0) return M3G_OK;
  M3G_REQUIRE(r && out, "m3g_cutoff: null pointer");
  cutoff_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(r, rc, n, out, dout);
  M3G_LAUNCH_CHECK("m3g_cutoff");
  return M3G_OK;
}

}  // extern "C"
