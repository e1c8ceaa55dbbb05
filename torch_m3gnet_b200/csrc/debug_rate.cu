// Dev tool (m3g_debug_pipe_rate): issue interval of the instruction kinds that pace the three-body MLP kernels, per
// SM sub-partition, as a function of the resident warps.  One CTA per SM; every thread runs ITER rounds of 8 independent
// dependency chains of one instruction kind; the kernel reports clock64 cycles of warp 0 of CTA 0.
//   kind 0: fma.rn.f32, three distinct register operands           kind 1: fma.rn.f32, multiplicands shared by the chains
//   kind 2: fma.rn.f32 with a kernel-parameter (constant bank) multiplicand
//   kind 3: fma.rn.f32x2, distinct operands                         kind 4: fma.rn.f32x2, multiplicands shared
//   kind 5: ex2.approx.ftz.f32 (MUFU)                               kind 6: mma.sync m16n8k8 tf32
//   kind 7: fma.rn.f32 with an immediate multiplicand
#include "common.cuh"

namespace m3g {

struct RateW { float w[64]; };

template <int KIND>
__global__ void __launch_bounds__(1024, 1) pipe_rate_kernel(int iters, const __grid_constant__ RateW cw, float seed,
                                                            float* sink, long long* out) {
  float a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + 0.001f * (threadIdx.x + i);
    b[i] = 1.0f - seed * 1e-6f * (i + 1 + (threadIdx.x & 3));
    c[i] = 0.5f * i;
  }
  unsigned long long a2[8], b2[8], c2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a2[i]) : "f"(a[i]), "f"(b[i]));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(b2[i]) : "f"(b[i]), "f"(b[i]));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(c2[i]) : "f"(c[i]), "f"(a[i]));
  }
  uint32_t ua[4], ub[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) ua[i] = __float_as_uint(a[i]) & 0xffffe000u;
  ub[0] = __float_as_uint(b[0]) & 0xffffe000u;
  ub[1] = __float_as_uint(b[1]) & 0xffffe000u;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(c[i]) : "f"(a[i]), "f"(b[i]));
        if (KIND == 1) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(c[i]) : "f"(a[0]), "f"(b[0]));
        if (KIND == 2) c[i] = fmaf(c[i], cw.w[8 * r + i], a[i]);
        if (KIND == 3) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c2[i]) : "l"(a2[i]), "l"(b2[i]));
        if (KIND == 4) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c2[i]) : "l"(a2[0]), "l"(b2[0]));
        if (KIND == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c[i]));
        if (KIND == 6)
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                       : "r"(ua[0]), "r"(ua[1]), "r"(ua[2]), "r"(ua[3]), "r"(ub[0]), "r"(ub[1]));
        if (KIND == 7) asm volatile("fma.rn.f32 %0, %0, 0f3F7FFFEF, %1;" : "+f"(c[i]) : "f"(a[i]));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(c2[i]));
    s += c[i] + lo + hi + acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  }
  if (s == 123.456f) sink[threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

}  // namespace m3g

using namespace m3g;

extern "C" int m3g_debug_pipe_rate(int kind, int threads, int iters, int64_t* cycles, void* stream) {
  M3G_REQUIRE(cycles && threads >= 32 && threads <= 1024 && threads % 32 == 0 && kind >= 0 && kind <= 7,
              "m3g_debug_pipe_rate: bad arguments");
  RateW cw;
  for (int i = 0; i < 64; ++i) cw.w[i] = 1.0f - 1e-6f * (i + 1);
  float* sink = nullptr;
  if (cudaMalloc(&sink, 1024 * sizeof(float)) != cudaSuccess) return M3G_ERR_CUDA;
  int dev = 0, n_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = as_stream(stream);
  long long* out = reinterpret_cast<long long*>(cycles);
#define RUN_(K) pipe_rate_kernel<K><<<n_sm, threads, 0, st>>>(iters, cw, 0.25f, sink, out)
  switch (kind) {
    case 0: RUN_(0); break;
    case 1: RUN_(1); break;
    case 2: RUN_(2); break;
    case 3: RUN_(3); break;
    case 4: RUN_(4); break;
    case 5: RUN_(5); break;
    case 6: RUN_(6); break;
    default: RUN_(7); break;
  }
#undef RUN_
  cudaError_t err = cudaStreamSynchronize(st);
  cudaFree(sink);
  if (err != cudaSuccess) {
    set_error("m3g_debug_pipe_rate: %s", cudaGetErrorString(err));
    return M3G_ERR_CUDA;
  }
  return M3G_OK;
}
