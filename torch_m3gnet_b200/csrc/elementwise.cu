// Elementwise kernels behind the module-level GatedMLP.forward (reference nn/core.py:61-62: dense(x) * gate(x), with
// torch.nn.SiLU / torch.nn.Sigmoid between the Linear layers, nn/core.py:45-59).  The fused layer kernels (conv_tc.cu,
// threebody_moment.cu, readout.cu) own the hot path; these serve a GatedMLP that is called on its own.
#include "common.cuh"

namespace m3g {

// kind 0: SiLU, 1: sigmoid (accurate expf: the values may feed energies directly)
__global__ void act_fwd_kernel(const float* __restrict__ in, int64_t n, int kind, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float z = in[i];
  out[i] = kind == 0 ? silu_acc(z) : sigmoid_acc(z);
}

// g_in = g_out * act'(in)
__global__ void act_bwd_kernel(const float* __restrict__ in, const float* __restrict__ g, int64_t n, int kind,
                               float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float z = in[i];
  float d;
  if (kind == 0) {
    d = silu_grad(z);
  } else {
    const float s = sigmoid_acc(z);
    d = s * (1.0f - s);
  }
  out[i] = g[i] * d;
}

__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}

__global__ void add_kernel(const float* a, const float* b, int64_t n, float* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// out[i] = sum_k slices[k * n + i], k ascending (the stated order of the per-launch g_h slices of the executor)
__global__ void sum_slices_kernel(const float* __restrict__ slices, int K, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = slices[i];
  for (int k = 1; k < K; ++k) acc += slices[(int64_t)k * n + i];
  out[i] = acc;
}

}  // namespace m3g

using namespace m3g;

extern "C" {

int m3g_act_fwd(const float* in, int64_t n, int kind, float* out, void* stream) {
  if (n == 0) return M3G_OK;
  M3G_REQUIRE(in && out && (kind == 0 || kind == 1), "m3g_act_fwd: bad argument");
  act_fwd_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(in, n, kind, out);
  M3G_LAUNCH_CHECK("m3g_act_fwd");
  return M3G_OK;
}

int m3g_act_bwd(const float* in, const float* g, int64_t n, int kind, float* out, void* stream) {
  if (n == 0) return M3G_OK;
  M3G_REQUIRE(in && g && out && (kind == 0 || kind == 1), "m3g_act_bwd: bad argument");
  act_bwd_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(in, g, n, kind, out);
  M3G_LAUNCH_CHECK("m3g_act_bwd");
  return M3G_OK;
}

int m3g_mul(const float* a, const float* b, int64_t n, float* out, void* stream) {
  if (n == 0) return M3G_OK;
  M3G_REQUIRE(a && b && out, "m3g_mul: null pointer");
  mul_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(a, b, n, out);
  M3G_LAUNCH_CHECK("m3g_mul");
  return M3G_OK;
}

int m3g_add(const float* a, const float* b, int64_t n, float* out, void* stream) {
  if (n == 0) return M3G_OK;
  M3G_REQUIRE(a && b && out, "m3g_add: null pointer");
  add_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(a, b, n, out);
  M3G_LAUNCH_CHECK("m3g_add");
  return M3G_OK;
}

int m3g_sum_slices(const float* slices, int K, int64_t n, float* out, void* stream) {
  if (n == 0) return M3G_OK;
  M3G_REQUIRE(slices && out && K >= 1, "m3g_sum_slices: bad argument");
  sum_slices_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(slices, K, n, out);
  M3G_LAUNCH_CHECK("m3g_sum_slices");
  return M3G_OK;
}

}  // extern "C"
