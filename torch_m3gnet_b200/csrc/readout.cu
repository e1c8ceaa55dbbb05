// AtomWiseReadout (nn/readout.py:39-58): 3-layer gated MLP F→F→F→1 per atom, per-structure energy sums,
// and the hand-written adjoint w.r.t. the node features.
#include <cstdlib>

#include "common.cuh"

namespace m3g {

// Two shapes of the same kernel:
//  * SW = false: 4 warps x 4 atoms, weights read through L2 for every k (any F <= M3G_MAX_F);
//  * SW = true : persistent CTAs of 8 warps x 8 atoms with all weight matrices staged once in shared memory
//                (F % 4 == 0 and 4 F^2 (fwd) / 8 F^2 (bwd) floats + scratch within 227 KB, i.e. F <= 64).
// The per-atom accumulation order (k ascending) is the same in both.
constexpr int RO_EPW = 4;
constexpr int RO_WARPS = 4;
constexpr int RO_EPW_SW = 8;
constexpr int RO_WARPS_SW = 8;

template <int NJ, int EPW, bool VEC>
__device__ __forceinline__ void ro_matvec(const float* xs, int xs_stride, int K, const float* __restrict__ Wt, int ldw,
                                          int ncols, int lane, float (&acc)[EPW][NJ]) {
  if (VEC) {  // K % 4 == 0, xs rows 16-byte aligned: four k per step, activations as float4 broadcasts
    for (int k = 0; k < K; k += 4) {
      float w[4][NJ];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          int col = lane + 32 * j;
          w[kk][j] = (col < ncols) ? Wt[(k + kk) * ldw + col] : 0.0f;
        }
#pragma unroll
      for (int q = 0; q < EPW; ++q) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + q * xs_stride + k);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          acc[q][j] += xv.x * w[0][j];
          acc[q][j] += xv.y * w[1][j];
          acc[q][j] += xv.z * w[2][j];
          acc[q][j] += xv.w * w[3][j];
        }
      }
    }
    return;
  }
  for (int k = 0; k < K; ++k) {
    float w[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int col = lane + 32 * j;
      w[j] = (col < ncols) ? Wt[(int64_t)k * ldw + col] : 0.0f;
    }
#pragma unroll
    for (int q = 0; q < EPW; ++q) {
      float xv = xs[q * xs_stride + k];
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[q][j] += xv * w[j];
    }
  }
}

struct ReadoutW {
  const float *W0dT, *b0d, *W1dT, *b1d, *w2d, *b2d;
  const float *W0gT, *b0g, *W1gT, *b1g, *w2g, *b2g;
  const float *W0d, *W1d, *W0g, *W1g;  // (out,in) layouts, backward only
};

// BWD = false: writes atomic[i] = elemental[i]/scale + eps_i
// BWD = true : writes g_x (and, when atomic_out is given, the forward result as well: the whole-step executor knows the
//              upstream gradient before the readout runs, so forward and adjoint are ONE launch there)
template <int NJ, bool BWD, int EPW, bool SW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
readout_kernel(const float* __restrict__ x, ReadoutW w, const float* __restrict__ elemental, float scale,
               const float* __restrict__ g_atomic, const float* __restrict__ g_scaled_total,
               const float* __restrict__ g_total, const int32_t* __restrict__ batch, int64_t N, int F,
               float* __restrict__ out, float* __restrict__ atomic_out) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RO_EPW = EPW;  // shadows the namespace constant inside the kernel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int F2 = 2 * F;
  if (SW) {  // weights -> shared memory, once per CTA
    float* ws = smem + n_warps * (RO_EPW * 5 * F);
    const int FF = F * F;
    const float* srcs[8] = {w.W0dT, w.W0gT, w.W1dT, w.W1gT, w.W0d, w.W0g, w.W1d, w.W1g};
    const int n_mat = BWD ? 8 : 4;
    for (int m = 0; m < n_mat; ++m) {
      const float4* g4 = reinterpret_cast<const float4*>(srcs[m]);
      float4* s4 = reinterpret_cast<float4*>(ws + m * FF);
      for (int t = threadIdx.x; t < FF / 4; t += blockDim.x) s4[t] = g4[t];
    }
    w.W0dT = ws; w.W0gT = ws + FF; w.W1dT = ws + 2 * FF; w.W1gT = ws + 3 * FF;
    if (BWD) { w.W0d = ws + 4 * FF; w.W0g = ws + 5 * FF; w.W1d = ws + 6 * FF; w.W1g = ws + 7 * FF; }
    __syncthreads();
  }
  float* xs = smem + warp * (RO_EPW * 5 * F);  // [EPW][F]
  float* z0_s = xs + RO_EPW * F;               // [EPW][2F]
  float* a0_s = z0_s + RO_EPW * F2;            // [EPW][2F]
  for (int64_t grp = (int64_t)blockIdx.x * n_warps + warp; grp * RO_EPW < N; grp += (int64_t)gridDim.x * n_warps) {
  const int64_t i0 = grp * RO_EPW;
  __syncwarp();  // the previous group's reads of the scratch rows are complete
  int64_t iq[RO_EPW];
#pragma unroll
  for (int q = 0; q < RO_EPW; ++q) iq[q] = min(i0 + q, N - 1);
  for (int q = 0; q < RO_EPW; ++q)
    for (int k = lane; k < F; k += 32) xs[q * F + k] = x[iq[q] * F + k];
  __syncwarp();
  // layer 0
  {
    float zd[RO_EPW][NJ], zg[RO_EPW][NJ];
#pragma unroll
    for (int q = 0; q < RO_EPW; ++q)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int col = lane + 32 * j;
        zd[q][j] = (col < F) ? w.b0d[col] : 0.0f;
        zg[q][j] = (col < F) ? w.b0g[col] : 0.0f;
      }
    ro_matvec<NJ, RO_EPW, SW>(xs, F, F, w.W0dT, F, F, lane, zd);
    ro_matvec<NJ, RO_EPW, SW>(xs, F, F, w.W0gT, F, F, lane, zg);
#pragma unroll
    for (int q = 0; q < RO_EPW; ++q)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int col = lane + 32 * j;
        if (col < F) {
          z0_s[q * F2 + col] = zd[q][j];
          z0_s[q * F2 + F + col] = zg[q][j];
          a0_s[q * F2 + col] = silu_acc(zd[q][j]);
          a0_s[q * F2 + F + col] = silu_acc(zg[q][j]);
        }
      }
  }
  __syncwarp();
  // layer 1
  float zd1[RO_EPW][NJ], zg1[RO_EPW][NJ];
#pragma unroll
  for (int q = 0; q < RO_EPW; ++q)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int col = lane + 32 * j;
      zd1[q][j] = (col < F) ? w.b1d[col] : 0.0f;
      zg1[q][j] = (col < F) ? w.b1g[col] : 0.0f;
    }
  ro_matvec<NJ, RO_EPW, SW>(a0_s, F2, F, w.W1dT, F, F, lane, zd1);
  ro_matvec<NJ, RO_EPW, SW>(a0_s + F, F2, F, w.W1gT, F, F, lane, zg1);
  // layer 2 (1 output): dense branch linear, gate branch sigmoid
  float dout[RO_EPW], gout[RO_EPW];
#pragma unroll
  for (int q = 0; q < RO_EPW; ++q) {
    float pd = 0.0f, pg = 0.0f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int col = lane + 32 * j;
      if (col < F) {
        pd += silu_acc(zd1[q][j]) * w.w2d[col];
        pg += silu_acc(zg1[q][j]) * w.w2g[col];
      }
    }
    dout[q] = warp_sum(pd) + w.b2d[0];
    gout[q] = sigmoid_acc(warp_sum(pg) + w.b2g[0]);
  }
  if (!BWD) {
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < RO_EPW; ++q)
        if (i0 + q < N) out[iq[q]] = elemental[iq[q]] / scale + dout[q] * gout[q];
    }
    continue;
  }
  if (atomic_out && lane == 0) {
#pragma unroll
    for (int q = 0; q < RO_EPW; ++q)
      if (i0 + q < N) atomic_out[iq[q]] = elemental[iq[q]] / scale + dout[q] * gout[q];
  }
  __syncwarp();  // a0_s is about to be overwritten with dz1
#pragma unroll
  for (int q = 0; q < RO_EPW; ++q) {
    float ge = 0.0f;
    if (g_atomic) ge += g_atomic[iq[q]];
    int b = batch[iq[q]];
    if (g_scaled_total) ge += g_scaled_total[b];
    if (g_total) ge += scale * g_total[b];
    float g_d = ge * gout[q];
    float g_gpre = ge * dout[q] * gout[q] * (1.0f - gout[q]);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int col = lane + 32 * j;
      if (col < F) {
        a0_s[q * F2 + col] = g_d * w.w2d[col] * silu_grad(zd1[q][j]);
        a0_s[q * F2 + F + col] = g_gpre * w.w2g[col] * silu_grad(zg1[q][j]);
      }
    }
  }
  __syncwarp();
  {
    float dd[RO_EPW][NJ], dg[RO_EPW][NJ];
#pragma unroll
    for (int q = 0; q < RO_EPW; ++q)
#pragma unroll
      for (int j = 0; j < NJ; ++j) { dd[q][j] = 0.0f; dg[q][j] = 0.0f; }
    ro_matvec<NJ, RO_EPW, SW>(a0_s, F2, F, w.W1d, F, F, lane, dd);
    ro_matvec<NJ, RO_EPW, SW>(a0_s + F, F2, F, w.W1g, F, F, lane, dg);
#pragma unroll
    for (int q = 0; q < RO_EPW; ++q)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int col = lane + 32 * j;
        if (col < F) {
          z0_s[q * F2 + col] = dd[q][j] * silu_grad(z0_s[q * F2 + col]);
          z0_s[q * F2 + F + col] = dg[q][j] * silu_grad(z0_s[q * F2 + F + col]);
        }
      }
  }
  __syncwarp();
  {
    float gx[RO_EPW][NJ];
#pragma unroll
    for (int q = 0; q < RO_EPW; ++q)
#pragma unroll
      for (int j = 0; j < NJ; ++j) gx[q][j] = 0.0f;
    ro_matvec<NJ, RO_EPW, SW>(z0_s, F2, F, w.W0d, F, F, lane, gx);
    ro_matvec<NJ, RO_EPW, SW>(z0_s + F, F2, F, w.W0g, F, F, lane, gx);
#pragma unroll
    for (int q = 0; q < RO_EPW; ++q)
      if (i0 + q < N) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          int col = lane + 32 * j;
          if (col < F) out[iq[q] * F + col] = gx[q][j];
        }
      }
  }
  }  // group loop
}

__global__ void structure_sum_kernel(const float* __restrict__ atomic, const int32_t* __restrict__ atom_ptr,
                                     int64_t B, float scale, float* __restrict__ scaled_total,
                                     float* __restrict__ total) {
  int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  float acc = 0.0f;
  for (int i = atom_ptr[b] + lane; i < atom_ptr[b + 1]; i += 32) acc += atomic[i];
  acc = warp_sum(acc);
  if (lane == 0) {
    scaled_total[b] = acc;
    total[b] = scale * acc;
  }
}

// persistent shape of the shared-weight variant: 8 warps x 8 atoms, or 16 warps x 4 atoms (M3G_RO_WARPS=16: more warps
// to hide the shared-memory latency of the matvecs at 2x the weight reads per atom)
static int ro_warps_sw() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("M3G_RO_WARPS");
    v = (e && atoi(e) == 8) ? 8 : 16;
  }
  return v;
}

template <bool BWD>
static int launch_readout(const float* x, const ReadoutW& w, const float* elemental, float scale,
                          const float* g_atomic, const float* g_st, const float* g_t, const int32_t* batch, int64_t N,
                          int F, float* out, float* atomic_out, cudaStream_t st) {
  const int nj = (F + 31) / 32;
  const int sw_warps = ro_warps_sw(), sw_epw = sw_warps == 16 ? 4 : 8;
  const size_t smem_sw = ((size_t)sw_warps * sw_epw * 5 * F + (size_t)(BWD ? 8 : 4) * F * F) * sizeof(float);
  const bool sw = (F % 4 == 0) && smem_sw <= 227 * 1024 && nj <= 2;
  if (sw) {
    static int n_sm = 0;
    if (n_sm == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    unsigned grid = blocks_for(N, sw_warps * sw_epw);
    if (grid > (unsigned)n_sm) grid = (unsigned)n_sm;
#define LAUNCH_SW_(NJ, EPW_, WARPS_)                                                                                \
  do {                                                                                                              \
    cudaFuncSetAttribute(readout_kernel<NJ, BWD, EPW_, true, WARPS_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                         (int)smem_sw);                                                                             \
    readout_kernel<NJ, BWD, EPW_, true, WARPS_><<<grid, WARPS_ * 32, smem_sw, st>>>(                                \
        x, w, elemental, scale, g_atomic, g_st, g_t, batch, N, F, out, atomic_out);                                 \
  } while (0)
    if (sw_warps == 16) {
      if (nj == 1) LAUNCH_SW_(1, 4, 16); else LAUNCH_SW_(2, 4, 16);
    } else {
      if (nj == 1) LAUNCH_SW_(1, 8, 8); else LAUNCH_SW_(2, 8, 8);
    }
#undef LAUNCH_SW_
    return 0;
  }
  size_t smem = (size_t)RO_WARPS * RO_EPW * 5 * F * sizeof(float);
  unsigned grid = blocks_for(N, RO_WARPS * RO_EPW);
#define LAUNCH_(NJ)                                                                                        \
  do {                                                                                                     \
    if (smem > 48 * 1024)                                                                                  \
      cudaFuncSetAttribute(readout_kernel<NJ, BWD, RO_EPW, false, RO_WARPS>,                               \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
    readout_kernel<NJ, BWD, RO_EPW, false, RO_WARPS><<<grid, RO_WARPS * 32, smem, st>>>(                   \
        x, w, elemental, scale, g_atomic, g_st, g_t, batch, N, F, out, atomic_out);                        \
  } while (0)
  if (nj == 1) LAUNCH_(1); else if (nj == 2) LAUNCH_(2); else if (nj == 3) LAUNCH_(3); else LAUNCH_(4);
#undef LAUNCH_
  return 0;
}

}  // namespace m3g

using namespace m3g;

extern "C" {

int m3g_readout_fwd(const float* x, const float* W0dT, const float* b0d, const float* W1dT, const float* b1d,
                    const float* w2d, const float* b2d, const float* W0gT, const float* b0g, const float* W1gT,
                    const float* b1g, const float* w2g, const float* b2g, const float* elemental, float scale,
                    int64_t N, int F, float* atomic, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(x && W0dT && b0d && W1dT && b1d && w2d && b2d && W0gT && b0g && W1gT && b1g && w2g && b2g &&
                  elemental && atomic,
              "m3g_readout_fwd: null pointer");
  M3G_REQUIRE(F >= 1 && F <= M3G_MAX_F, "m3g_readout_fwd: F=%d outside [1,%d]", F, M3G_MAX_F);
  ReadoutW w{W0dT, b0d, W1dT, b1d, w2d, b2d, W0gT, b0g, W1gT, b1g, w2g, b2g, nullptr, nullptr, nullptr, nullptr};
  launch_readout<false>(x, w, elemental, scale, nullptr, nullptr, nullptr, nullptr, N, F, atomic, nullptr,
                        as_stream(stream));
  M3G_LAUNCH_CHECK("m3g_readout_fwd");
  return M3G_OK;
}

int m3g_structure_sum(const float* atomic, const int32_t* atom_ptr, int64_t B, float scale, float* scaled_total,
                      float* total, void* stream) {
  if (B == 0) return M3G_OK;
  M3G_REQUIRE(atomic && atom_ptr && scaled_total && total, "m3g_structure_sum: null pointer");
  structure_sum_kernel<<<blocks_for(B * 32, 128), 128, 0, as_stream(stream)>>>(atomic, atom_ptr, B, scale,
                                                                               scaled_total, total);
  M3G_LAUNCH_CHECK("m3g_structure_sum");
  return M3G_OK;
}

int m3g_readout_bwd(const float* x, const float* W0dT, const float* b0d, const float* W1dT, const float* b1d,
                    const float* w2d, const float* b2d, const float* W0gT, const float* b0g, const float* W1gT,
                    const float* b1g, const float* w2g, const float* b2g, const float* W0d, const float* W1d,
                    const float* W0g, const float* W1g, const float* g_atomic, const float* g_scaled_total,
                    const float* g_total, const int32_t* batch, float scale, int64_t N, int F, float* g_x,
                    void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(x && W0dT && b0d && W1dT && b1d && w2d && b2d && W0gT && b0g && W1gT && b1g && w2g && b2g && W0d &&
                  W1d && W0g && W1g && batch && g_x,
              "m3g_readout_bwd: null pointer");
  M3G_REQUIRE(F >= 1 && F <= M3G_MAX_F, "m3g_readout_bwd: F=%d outside [1,%d]", F, M3G_MAX_F);
  ReadoutW w{W0dT, b0d, W1dT, b1d, w2d, b2d, W0gT, b0g, W1gT, b1g, w2g, b2g, W0d, W1d, W0g, W1g};
  launch_readout<true>(x, w, nullptr, scale, g_atomic, g_scaled_total, g_total, batch, N, F, g_x, nullptr,
                       as_stream(stream));
  M3G_LAUNCH_CHECK("m3g_readout_bwd");
  return M3G_OK;
}

int m3g_readout_fwd_bwd(const float* x, const float* W0dT, const float* b0d, const float* W1dT, const float* b1d,
                        const float* w2d, const float* b2d, const float* W0gT, const float* b0g, const float* W1gT,
                        const float* b1g, const float* w2g, const float* b2g, const float* W0d, const float* W1d,
                        const float* W0g, const float* W1g, const float* elemental, const float* g_total,
                        const int32_t* batch, float scale, int64_t N, int F, float* atomic, float* g_x, void* stream) {
  if (N == 0) return M3G_OK;
  M3G_REQUIRE(x && W0dT && b0d && W1dT && b1d && w2d && b2d && W0gT && b0g && W1gT && b1g && w2g && b2g && W0d &&
                  W1d && W0g && W1g && elemental && g_total && batch && atomic && g_x,
              "m3g_readout_fwd_bwd: null pointer");
  M3G_REQUIRE(F >= 1 && F <= M3G_MAX_F, "m3g_readout_fwd_bwd: F=%d outside [1,%d]", F, M3G_MAX_F);
  ReadoutW w{W0dT, b0d, W1dT, b1d, w2d, b2d, W0gT, b0g, W1gT, b1g, w2g, b2g, W0d, W1d, W0g, W1g};
  launch_readout<true>(x, w, elemental, scale, nullptr, nullptr, g_total, batch, N, F, g_x, atomic, as_stream(stream));
  M3G_LAUNCH_CHECK("m3g_readout_fwd_bwd");
  return M3G_OK;
}

}  // extern "C"
