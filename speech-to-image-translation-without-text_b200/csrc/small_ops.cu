// Small fp32 operators of the path whose cost is launch latency or weight bandwidth, not FLOPs:
//   CA_NET fc + GLU + reparameterisation   (model.py:172-200)
//   INIT_STAGE_G fc 228 -> 32768           (model.py:216-219)    [M = batch <= 64 rows]
//   D logits: conv k4 s4 512 -> 1 + bias + sigmoid == one 8192-long dot product per sample (model.py:414-422)
//   fused Adam(beta1=.5) + EMA             (trainer.py:236-252, 571-572)
// Warp-shuffle reductions, coalesced weight reads; fp32 master weights are read directly (no bf16 copy).
#include "../../include/sg2b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sg2 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }

// ------------------------------------------------------------------------------------------ linear (tiny M)
// out[m][n] = sum_k x(m,k) * w[n][k] (+ bias[n]);  x(m,k) = k < K1 ? x1[m][k] : x2[m][k-K1]  (cat(c_code, z))
// Weight-bandwidth bound (M = batch rows <= 64): the activations are staged in shared memory (K chunks of 256), each
// warp walks `npw` output features, lanes stride K (coalesced weight rows read ONCE), 32 batch rows accumulate in
// registers and one butterfly transpose-sum per feature leaves row m's result in lane m.
constexpr int kLinKC = 256;
constexpr int kLinF = 2;     // features per warp step in linear_fwd (they share the shared-memory reads of x)
__device__ __forceinline__ float lin_x(const float* x1, int K1, const float* x2, int K2, int m, int k) {
  return k < K1 ? x1[(long long)m * K1 + k] : x2[(long long)m * K2 + (k - K1)];
}
template <bool OUT_BF16>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x1, int K1, const float* __restrict__ x2,
                                                         int K2, const float* __restrict__ w,
                                                         const float* __restrict__ bias, void* __restrict__ out, int M,
                                                         int N, int npw) {
  // each warp walks `npw` groups of kLinF features; the kLinF features of a group share every shared-memory read of x
  __shared__ float xs[32][kLinKC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = K1 + K2;
  const int nchunks = (K + kLinKC - 1) / kLinKC;
  const int nbase = (blockIdx.x * 8 + warp) * npw * kLinF;
  for (int m0 = 0; m0 < M; m0 += 32) {
    for (int i = 0; i < npw; ++i) {
      const int n = nbase + i * kLinF;
      float acc[kLinF][32];
#pragma unroll
      for (int f = 0; f < kLinF; ++f)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[f][j] = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        if (nchunks > 1 || i == 0) {  // a single chunk stays staged for all features of the block
          __syncthreads();
          for (int e = threadIdx.x; e < 32 * kLinKC; e += 256) {
            const int m = m0 + e / kLinKC, k = c * kLinKC + e % kLinKC;
            xs[e / kLinKC][e % kLinKC] = (m < M && k < K) ? lin_x(x1, K1, x2, K2, m, k) : 0.f;
          }
          __syncthreads();
        }
        const int kend = min(kLinKC, K - c * kLinKC);
        for (int kk = lane; kk < kend; kk += 32) {
          float wv[kLinF];
#pragma unroll
          for (int f = 0; f < kLinF; ++f) wv[f] = n + f < N ? w[(long long)(n + f) * K + c * kLinKC + kk] : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float xv = xs[j][kk];
#pragma unroll
            for (int f = 0; f < kLinF; ++f) acc[f][j] = fmaf(wv[f], xv, acc[f][j]);
          }
        }
      }
#pragma unroll
      for (int f = 0; f < kLinF; ++f) {
        const float s = warp_transpose_sum(acc[f], lane);  // lane m: sum over k of row m0 + m
        const int m = m0 + lane;
        if (n + f < N && m < M) {
          const float v = s + (bias ? bias[n + f] : 0.f);
          if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(out)[(long long)m * N + n + f] = __float2bfloat16_rn(v);
          else reinterpret_cast<float*>(out)[(long long)m * N + n + f] = v;
        }
      }
    }
  }
}

// Register-tiled form for K <= 1024 (both linears of the path: 1024 -> 512 and 228 -> 32768): the kernel above reads
// shared memory 32 times per weight element (one batch row each), which caps it near 27 us for the 30 MB weight matrix
// and leaves it latency bound at ~70 us. Here a warp owns 8 batch rows (m-group = warp & 3) and keeps their activations
// for its lane's k values in registers (8 k x 8 rows); four features are in flight per step (32 coalesced weight loads
// per lane), the 8 row sums of a feature are reduced across the lanes with a halving butterfly (9 shuffles). The whole
// activation matrix [32][K] sits in shared memory once per block.
__device__ __forceinline__ float warp_reduce8(float (&v)[8], int lane) {
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];   // every lane: the sum of row ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)
}
constexpr int kRtF = 4;   // features in flight per warp step
// RPW = batch rows per warp: 8 (M up to 32) or 6 (M <= 24: all four m-groups carry rows, and x (8 x 6) + weights (4 x 8) +
// accumulators (4 x 6) fit the 128 registers of two resident blocks per SM — the 8-row form needs 147, i.e. one block per
// SM and two waves for the 296-block grid of the 30 MB fc).
template <bool OUT_BF16, int RPW>
__global__ void __launch_bounds__(256, 2) linear_fwd_rt_kernel(const float* __restrict__ x1, int K1, const float* __restrict__ x2,
                                                               int K2, const float* __restrict__ w,
                                                               const float* __restrict__ bias, void* __restrict__ out, int M,
                                                               int N, int Kp) {
  extern __shared__ __align__(16) float xs[];   // [4 * RPW][Kp]
  constexpr int kRows = 4 * RPW;
  const int K = K1 + K2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mg = warp & 3, fs = warp >> 2;
  const int nchunks = (Kp + 255) / 256;
  const int ngroups = (N + kRtF - 1) / kRtF;
  const int row = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const bool vec = ((K1 | K2) & 3) == 0 && ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2)) & 15) == 0;
  for (int m0 = 0; m0 < M; m0 += kRows) {
    __syncthreads();
    if (vec) {
      // 16-byte loads, every load of the thread in flight at once (the scalar loop was a chain of ~128 dependent round trips)
      const int kq = Kp >> 2;
      for (int e = threadIdx.x; e < kRows * kq; e += 256) {
        const int r = e / kq, k = (e - r * kq) << 2, m = m0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M && k < K)
          v = k < K1 ? __ldg(reinterpret_cast<const float4*>(x1 + (long long)m * K1 + k))
                     : __ldg(reinterpret_cast<const float4*>(x2 + (long long)m * K2 + (k - K1)));
        *reinterpret_cast<float4*>(xs + r * Kp + k) = v;
      }
    } else {
      for (int e = threadIdx.x; e < kRows * Kp; e += 256) {
        const int m = m0 + e / Kp, k = e % Kp;
        xs[e] = (m < M && k < K) ? lin_x(x1, K1, x2, K2, m, k) : 0.f;
      }
    }
    __syncthreads();
    float xr[8][RPW];
    auto load_x = [&](int ch) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = ch * 256 + j * 32 + lane;
#pragma unroll
        for (int r = 0; r < RPW; ++r) xr[j][r] = k < Kp ? xs[(mg * RPW + r) * Kp + k] : 0.f;
      }
    };
    if (nchunks == 1) load_x(0);
    for (int g = blockIdx.x * 2 + fs; g < ngroups; g += gridDim.x * 2) {
      const int n = g * kRtF;
      float acc[kRtF][8];
#pragma unroll
      for (int f = 0; f < kRtF; ++f)
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[f][r] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        if (nchunks > 1) load_x(ch);
        float wv[kRtF][8];
#pragma unroll
        for (int f = 0; f < kRtF; ++f)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = ch * 256 + j * 32 + lane;
            wv[f][j] = (n + f < N && k < K) ? __ldg(w + (long long)(n + f) * K + k) : 0.f;
          }
#pragma unroll
        for (int f = 0; f < kRtF; ++f)
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int r = 0; r < RPW; ++r) acc[f][r] = fmaf(wv[f][j], xr[j][r], acc[f][r]);
      }
#pragma unroll
      for (int f = 0; f < kRtF; ++f) {
        const float s = warp_reduce8(acc[f], lane);
        const int m = m0 + mg * RPW + row;
        if ((lane & 3) == 0 && row < RPW && n + f < N && m < M) {
          const float v = s + (bias ? bias[n + f] : 0.f);
          if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(out)[(long long)m * N + n + f] = __float2bfloat16_rn(v);
          else reinterpret_cast<float*>(out)[(long long)m * N + n + f] = v;
        }
      }
    }
  }
}

// dw[n][k] (+)= sum_m dy[m][n] * x(m,k);  dbias[n] (+)= sum_m dy[m][n].   M <= 32 per launch.
// block = 256 consecutive k (grid.y = k slabs) x kLinNB features (grid.x): thread k keeps its activation column in
// registers, the dy columns of the block sit in shared memory (broadcast reads), dw rows are written coalesced.
constexpr int kLinNB = 64;
template <bool DY_BF16>
__global__ void __launch_bounds__(256) linear_bwd_w_kernel(const void* __restrict__ dy, const float* __restrict__ x1,
                                                           int K1, const float* __restrict__ x2, int K2,
                                                           float* __restrict__ dw, float* __restrict__ dbias, int M, int N,
                                                           int accumulate) {
  __shared__ float ds[kLinNB][32];  // [n][m]: one feature's batch column is one 128-byte row
  const int K = K1 + K2;
  const int n0 = blockIdx.x * kLinNB;
  const int k = blockIdx.y * 256 + threadIdx.x;
  for (int e = threadIdx.x; e < 32 * kLinNB; e += 256) {
    const int m = e / kLinNB, j = e % kLinNB, n = n0 + j;
    float d = 0.f;
    if (n < N && m < M)
      d = DY_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(dy)[(long long)m * N + n])
                  : reinterpret_cast<const float*>(dy)[(long long)m * N + n];
    ds[j][m] = d;
  }
  __syncthreads();
  if (dbias && blockIdx.y == 0 && threadIdx.x < kLinNB && n0 + threadIdx.x < N) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += ds[threadIdx.x][m];
    if (accumulate) dbias[n0 + threadIdx.x] += s; else dbias[n0 + threadIdx.x] = s;
  }
  if (k >= K) return;
  float xr[32];
#pragma unroll
  for (int m = 0; m < 32; ++m) xr[m] = m < M ? lin_x(x1, K1, x2, K2, m, k) : 0.f;
  const int nend = min(kLinNB, N - n0);
  for (int j = 0; j < nend; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int m4 = 0; m4 < 8; ++m4) {   // broadcast 128-bit reads of the feature's batch column
      const float4 d4 = reinterpret_cast<const float4*>(ds[j])[m4];
      acc = fmaf(d4.x, xr[4 * m4], acc);
      acc = fmaf(d4.y, xr[4 * m4 + 1], acc);
      acc = fmaf(d4.z, xr[4 * m4 + 2], acc);
      acc = fmaf(d4.w, xr[4 * m4 + 3], acc);
    }
    float* o = dw + (long long)(n0 + j) * K + k;
    if (accumulate) *o += acc; else *o = acc;
  }
}

// dx[m][k] = sum_n dy[m][n] * w[n][k] for k < Kout (only the leading Kout inputs need a gradient).   M <= 32 per launch.
// Pass 1 (grid.x = feature slabs of kLinNB, thread = k): the slab's dy sits in shared memory, all batch rows accumulate
// in registers while the weight rows are read once; the slab's partial [M][Kout] goes to scratch. Pass 2 sums the slabs
// (deterministic, no atomics).
template <bool DY_BF16>
__global__ void __launch_bounds__(256) linear_bwd_x_kernel(const void* __restrict__ dy, const float* __restrict__ w,
                                                           float* __restrict__ part, int M, int N, int K, int Kout) {
  __shared__ float ds[kLinNB][32];  // [n][m]
  const int n0 = blockIdx.x * kLinNB;
  for (int e = threadIdx.x; e < 32 * kLinNB; e += blockDim.x) {
    const int m = e / kLinNB, j = e % kLinNB, n = n0 + j;
    float d = 0.f;
    if (n < N && m < M)
      d = DY_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(dy)[(long long)m * N + n])
                  : reinterpret_cast<const float*>(dy)[(long long)m * N + n];
    ds[j][m] = d;
  }
  __syncthreads();
  const int nend = min(kLinNB, N - n0);
  for (int k = threadIdx.x; k < Kout; k += blockDim.x) {
    float acc[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) acc[m] = 0.f;
    for (int j0 = 0; j0 < nend; j0 += 8) {      // 8 weight rows in flight per thread
      float wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = j0 + u < nend ? w[(long long)(n0 + j0 + u) * K + k] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (j0 + u < nend) {
#pragma unroll
          for (int m4 = 0; m4 < 8; ++m4) {
            const float4 d4 = reinterpret_cast<const float4*>(ds[j0 + u])[m4];
            acc[4 * m4] = fmaf(d4.x, wv[u], acc[4 * m4]);
            acc[4 * m4 + 1] = fmaf(d4.y, wv[u], acc[4 * m4 + 1]);
            acc[4 * m4 + 2] = fmaf(d4.z, wv[u], acc[4 * m4 + 2]);
            acc[4 * m4 + 3] = fmaf(d4.w, wv[u], acc[4 * m4 + 3]);
          }
        }
      }
    }
    float* o = part + ((long long)blockIdx.x * 32) * Kout + k;
#pragma unroll
    for (int m = 0; m < 32; ++m)
      if (m < M) o[(long long)m * Kout] = acc[m];
  }
}
// grid (ceil(Kout / 32), M), block (32, 8): the 8 thread rows split the slabs, then one shared-memory reduction
__global__ void linear_bwd_x_reduce_kernel(const float* __restrict__ part, float* __restrict__ dx, int nslabs, int M,
                                           int Kout) {
  const int k = blockIdx.x * 32 + threadIdx.x, m = blockIdx.y, g = threadIdx.y;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (k < Kout) {
    int b = g;
    for (; b + 24 < nslabs; b += 32) {
      s0 += part[((long long)(b + 0) * 32 + m) * Kout + k];
      s1 += part[((long long)(b + 8) * 32 + m) * Kout + k];
      s2 += part[((long long)(b + 16) * 32 + m) * Kout + k];
      s3 += part[((long long)(b + 24) * 32 + m) * Kout + k];
    }
    for (; b < nslabs; b += 8) s0 += part[((long long)b * 32 + m) * Kout + k];
  }
  __shared__ float sh[8][32];
  sh[g][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (g == 0 && k < Kout) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += sh[j][threadIdx.x];
    dx[(long long)m * Kout + k] = t;
  }
}

// ------------------------------------------------------------------------------------------ CA_NET tail
// fc [B][4E] -> GLU -> [B][2E] = (mu | logvar);  c = eps * exp(0.5 * logvar) + mu
__global__ void ca_glu_reparam_fwd_kernel(const float* __restrict__ fc, const float* __restrict__ eps,
                                          float* __restrict__ mu, float* __restrict__ logvar, float* __restrict__ c,
                                          int B, int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * E) return;
  const int b = i / E, e = i % E;
  const float* r = fc + (long long)b * 4 * E;
  const float m = r[e] * sigm(r[2 * E + e]);
  const float lv = r[E + e] * sigm(r[3 * E + e]);
  mu[i] = m;
  logvar[i] = lv;
  c[i] = eps[i] * __expf(0.5f * lv) + m;
}
// dfc from (dmu, dlogvar, dc): dmu_t = dmu + dc ; dlv_t = dlogvar + dc * eps * 0.5 * exp(0.5 logvar); then GLU'.
__global__ void ca_glu_reparam_bwd_kernel(const float* __restrict__ fc, const float* __restrict__ eps,
                                          const float* __restrict__ dmu, const float* __restrict__ dlogvar,
                                          const float* __restrict__ dc, float* __restrict__ dfc, int B, int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * E) return;
  const int b = i / E, e = i % E;
  const float* r = fc + (long long)b * 4 * E;
  float* o = dfc + (long long)b * 4 * E;
  const float s_m = sigm(r[2 * E + e]), s_l = sigm(r[3 * E + e]);
  const float lv = r[E + e] * s_l;
  const float g_c = dc ? dc[i] : 0.f;
  const float g_m = (dmu ? dmu[i] : 0.f) + g_c;
  const float g_l = (dlogvar ? dlogvar[i] : 0.f) + g_c * eps[i] * 0.5f * __expf(0.5f * lv);
  o[e] = g_m * s_m;
  o[2 * E + e] = g_m * r[e] * s_m * (1.f - s_m);
  o[E + e] = g_l * s_l;
  o[3 * E + e] = g_l * r[E + e] * s_l * (1.f - s_l);
}

// ------------------------------------------------------------------------------------------ fc feature permute
// INIT_STAGE_G: GLU output [B][C*HW] in NCHW feature order (f = c*HW + p) <-> NHWC bf16 [B][HW][C]
__global__ void chw_to_hwc_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B,
                                       int C, int HW, int to_hwc) {
  const long long total = (long long)B * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes the HWC side (coalesced on c)
    const int c = (int)(i % C);
    const int p = (int)((i / C) % HW);
    const long long b = i / ((long long)C * HW);
    const long long j = (b * C + c) * HW + p;
    if (to_hwc) out[i] = in[j]; else out[j] = in[i];
  }
}

// ------------------------------------------------------------------------------------------ D logits
// prob[b] = sigmoid(bias + sum_{p,c} x[b][p][c] * w[c][p]),  x NHWC bf16 [B][HW=16][C], w fp32 OIHW [1][C][4][4]
// Each thread moves 8-channel (16-byte) vectors of x; w[c][p] of those channels comes through L1 (32 KB in all).
__device__ __forceinline__ void lg_unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__global__ void logits_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ bias, float* __restrict__ prob, int HW, int C) {
  const int b = blockIdx.x;
  const int n = HW * C, nv = n >> 3;
  const uint4* xv = reinterpret_cast<const uint4*>(x + (long long)b * n);
  float acc = 0.f;
  for (int v = threadIdx.x; v < nv; v += blockDim.x) {
    const int i = v << 3, p = i / C, c = i - p * C;
    float f[8];
    lg_unpack8(xv[v], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(f[j], w[(c + j) * HW + p], acc);
  }
  __shared__ float sh[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) prob[b] = sigm(v + bias[0]);
  }
}
// dpre[b] = dprob[b] * p (1-p);  dx[b][p][c] (=|+=) dpre[b] * w[c][p];  dw_parts[chunk][c][p] = sum_{b in chunk} dpre[b] x[b][p][c];  dbias += sum dpre
// grid (vectors / 256, sample chunks of kLgChunk): a thread owns one 8-channel vector position for its chunk of samples.
constexpr int kLgChunk = 8;
__global__ void logits_bwd_kernel(const float* __restrict__ dprob, const float* __restrict__ prob,
                                  const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                  __nv_bfloat16* __restrict__ dx, int dx_accumulate, float* __restrict__ dw_parts,
                                  float* __restrict__ dbias, int B, int HW, int C) {
  const int n = HW * C, nv = n >> 3;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int b0 = blockIdx.y * kLgChunk, b1 = min(B, b0 + kLgChunk);
  if (dbias && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) { const float pr = prob[b]; s += dprob[b] * pr * (1.f - pr); }
    dbias[0] += s;
  }
  if (v >= nv) return;
  const int i = v << 3, p = i / C, c = i - p * C;
  float wv[8], gw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { wv[j] = w[(c + j) * HW + p]; gw[j] = 0.f; }
  for (int b = b0; b < b1; ++b) {
    const float pr = prob[b];
    const float dpre = dprob[b] * pr * (1.f - pr);
    const long long o = ((long long)b * n >> 3) + v;
    float f[8];
    lg_unpack8(reinterpret_cast<const uint4*>(x)[o], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) gw[j] = fmaf(dpre, f[j], gw[j]);
    if (dx) {
      float d[8];
      if (dx_accumulate) lg_unpack8(reinterpret_cast<const uint4*>(dx)[o], d);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = dx_accumulate ? fmaf(dpre, wv[j], d[j]) : dpre * wv[j];
      uint4 q;
      q.x = pack_bf16x2(d[0], d[1]); q.y = pack_bf16x2(d[2], d[3]);
      q.z = pack_bf16x2(d[4], d[5]); q.w = pack_bf16x2(d[6], d[7]);
      reinterpret_cast<uint4*>(dx)[o] = q;
    }
  }
  if (dw_parts) {   // this sample chunk's partial, summed over the chunks in order by sg2_reduce_slabs (no atomics)
    float* dst = dw_parts + (long long)blockIdx.y * n;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[(c + j) * HW + p] = gw[j];
  }
}

// ------------------------------------------------------------------------------------------ jointConv c_code folding
// NEXT_STAGE_G.jointConv (model.py:274-279) convolves cat(c_code broadcast over the image, h). The c_code channels are
// constant over the image, so their part of the 3x3 convolution is a per-sample bias that only depends on which taps
// fall inside the image: 9 border classes (top / interior / bottom row) x (left / interior / right column).
//   Wc[b][tap][o]   = sum_e W[o][e][tap] * c[b][e]
//   bias9[b][r][o]  = sum over the taps valid in border class r of Wc[b][tap][o]
// Backward needs S[b][tap][o] = sum of dy[b][y][x][o] over the pixels where tap is valid, then
//   dW[o][e][tap] = sum_b c[b][e] * S[b][tap][o],   dc[b][e] += sum_{o,tap} W[o][e][tap] * S[b][tap][o].
// W is addressed as w[o * so + e * se + tap * st] (OIHW or OHWI master; e runs over the first E input channels).
__device__ __forceinline__ bool joint_tap_valid(int tap, int r) {
  const int ky = tap / 3, kx = tap % 3, ry = r / 3, rx = r % 3;
  return !(ry == 0 && ky == 0) && !(ry == 2 && ky == 2) && !(rx == 0 && kx == 0) && !(rx == 2 && kx == 2);
}
// grid (Cout, B), 288 threads: warp = tap
__global__ void joint_bias_kernel(const float* __restrict__ c, const float* __restrict__ w, long long so, long long se,
                                  long long st, float* __restrict__ bias9, int E, int Cout) {
  const int o = blockIdx.x, b = blockIdx.y, tap = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ float wc[9];
  float acc = 0.f;
  for (int e = lane; e < E; e += 32) acc = fmaf(w[o * so + e * se + tap * st], c[(long long)b * E + e], acc);
  acc = warp_sum(acc);
  if (lane == 0) wc[tap] = acc;
  __syncthreads();
  if (threadIdx.x < 9) {
    const int r = threadIdx.x;
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) v += joint_tap_valid(t, r) ? wc[t] : 0.f;
    bias9[((long long)b * 9 + r) * Cout + o] = v;
  }
}
// R[b][r][o] += sums of dy over the pixels of border class r. grid (H, B): one image row per block (one row class);
// 256 threads = 8-channel vectors x pixel slots.
__global__ void __launch_bounds__(256)
joint_region_sums_kernel(const uint4* __restrict__ dy, double* __restrict__ R, int H, int W, int Cout) {
  const int y = blockIdx.x, b = blockIdx.y;
  const int vc = Cout >> 3;                       // vectors per pixel
  const int v = threadIdx.x % vc, slot = threadIdx.x / vc, nslot = blockDim.x / vc;
  const int ry = y == 0 ? 0 : (y == H - 1 ? 2 : 1);
  const uint4* row = dy + ((long long)b * H + y) * W * vc;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  double* Rb = R + ((long long)b * 9 + ry * 3) * Cout + v * 8;   // fp64 atomics: the sum does not depend on block order
  // interior columns 1 .. W-2 (branch-free, loads pipelined); the two border columns go straight to their classes
#pragma unroll 4
  for (int x = 1 + slot; x < W - 1; x += nslot) {
    float f[8];
    lg_unpack8(row[(long long)x * vc + v], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
  if (slot < 2) {
    float f[8];
    lg_unpack8(row[(long long)(slot == 0 ? 0 : W - 1) * vc + v], f);
    double* dst = Rb + (slot == 0 ? 0 : 2) * Cout;
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dst + j, (double)f[j]);
  }
  __shared__ float sh[256][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (slot == 0) {
    for (int k = 1; k < nslot; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += sh[k * vc + v][j];
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(Rb + Cout + j, (double)acc[j]);
  }
}
// S[b][tap][o] = sum over the border classes where tap is valid of R[b][r][o]
__global__ void joint_tap_sums_kernel(const double* __restrict__ R, float* __restrict__ S, int B, int Cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 9 * Cout) return;
  const int o = i % Cout, tap = (i / Cout) % 9, b = i / (9 * Cout);
  double v = 0.0;
#pragma unroll
  for (int r = 0; r < 9; ++r) v += joint_tap_valid(tap, r) ? R[((long long)b * 9 + r) * Cout + o] : 0.0;
  S[i] = (float)v;
}
// dW[o][e][tap] (=|+=) sum_b c[b][e] * S[b][tap][o].  grid (Cout, 9), threads over e.
__global__ void joint_dw_kernel(const float* __restrict__ S, const float* __restrict__ c, float* __restrict__ dw,
                                long long so, long long se, long long st, int accumulate, int B, int E, int Cout) {
  const int o = blockIdx.x, tap = blockIdx.y;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(c[(long long)b * E + e], S[((long long)b * 9 + tap) * Cout + o], acc);
    float* d = dw + o * so + e * se + tap * st;
    *d = accumulate ? *d + acc : acc;
  }
}
// dc[b][e] += sum_{o,tap} W[o][e][tap] * S[b][tap][o].  grid (B), 1024 threads = 8 K-slices x up to 128 channels e:
// slice k walks the (o, tap) pairs k, k+8, ... (fixed order), the 8 slice sums are combined in slice order through
// shared memory — deterministic, no atomics. S[b] is staged in shared memory (9 * Cout floats).
constexpr int kJointSlices = 8;
__global__ void __launch_bounds__(1024)
joint_dc_kernel(const float* __restrict__ S, const float* __restrict__ w, long long so, long long se, long long st,
                float* __restrict__ dc, int E, int Cout) {
  extern __shared__ float jsm[];            // [9 * Cout] S of this sample, then [kJointSlices][ept] partials
  const int b = blockIdx.x;
  const int K = 9 * Cout;
  for (int i = threadIdx.x; i < K; i += blockDim.x) jsm[i] = S[(long long)b * K + i];   // index = tap * Cout + o
  __syncthreads();
  const int ept = blockDim.x / kJointSlices;   // channels handled per pass
  const int el = threadIdx.x % ept, ks = threadIdx.x / ept;
  float* part = jsm + K;
  for (int e0 = 0; e0 < E; e0 += ept) {
    const int e = e0 + el;
    float acc = 0.f;
    if (e < E) {
#pragma unroll 4
      for (int k = ks; k < K; k += kJointSlices) {
        const int tap = k / Cout, o = k - tap * Cout;
        acc = fmaf(w[o * so + e * se + tap * st], jsm[k], acc);
      }
    }
    part[ks * ept + el] = acc;
    __syncthreads();
    if (ks == 0 && e < E) {
      float t = part[el];
#pragma unroll
      for (int k2 = 1; k2 < kJointSlices; ++k2) t += part[k2 * ept + el];
      dc[(long long)b * E + e] += t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ Adam (+ EMA)
// torch.optim.Adam semantics (no amsgrad, no weight decay): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).  Optional EMA: avg = 0.999 avg + 0.001 p  (trainer.py:571-572).
// step counter lives on the device so that a captured CUDA graph replays the right bias corrections
__global__ void adam_tick_kernel(int* __restrict__ step, float* __restrict__ bc, float b1, float b2) {
  const int t = ++(*step);
  bc[0] = 1.f - powf(b1, (float)t);
  bc[1] = sqrtf(1.f - powf(b2, (float)t));
}
__device__ __forceinline__ float adam_one(float pi, float gi, float& mi, float& vi, float b1, float b2, float eps,
                                          float step, float inv_bc2_sqrt) {
  mi = b1 * mi + (1.f - b1) * gi;
  vi = b2 * vi + (1.f - b2) * gi * gi;
  const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
  return pi - step * (mi / denom);
}
// One Adam step (+ EMA of the parameters, + bf16 mirror) over a flat slice. The slice may start at any element of
// the flat buckets: a scalar head brings every array to a 16-byte boundary (they all share the slice offset), the
// body moves float4 vectors, a scalar tail finishes.
__global__ void __launch_bounds__(256)
adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                float* __restrict__ avg, long long n, float lr, float b1, float b2, float eps,
                const float* __restrict__ bc, float ema_decay, __nv_bfloat16* __restrict__ p_bf16, int head) {
  const float step = lr / bc[0], inv_bc2_sqrt = 1.f / bc[1];
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  auto scalar = [&](long long i) {
    float mi = m[i], vi = v[i];
    const float pi = adam_one(p[i], g[i], mi, vi, b1, b2, eps, step, inv_bc2_sqrt);
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pi);   // bf16 mirror = the conv operand pack of OHWI-stored weights
    if (avg) avg[i] = ema_decay * avg[i] + (1.f - ema_decay) * pi;
  };
  if (head < 0) {                                      // arrays not mutually aligned: all scalar
    for (long long i = tid; i < n; i += nthreads) scalar(i);
    return;
  }
  const long long h = head < n ? head : n;
  const long long n4 = (n - h) >> 2;
  if (tid < h) scalar(tid);
  const long long tail0 = h + (n4 << 2);
  if (tid < n - tail0) scalar(tail0 + tid);
  float4* p4 = reinterpret_cast<float4*>(p + h);
  const float4* g4 = reinterpret_cast<const float4*>(g + h);
  float4* m4 = reinterpret_cast<float4*>(m + h);
  float4* v4 = reinterpret_cast<float4*>(v + h);
  float4* a4 = avg ? reinterpret_cast<float4*>(avg + h) : nullptr;
  uint2* q4 = p_bf16 ? reinterpret_cast<uint2*>(p_bf16 + h) : nullptr;
  for (long long i = tid; i < n4; i += nthreads) {
    const float4 gi = g4[i], pi = p4[i];
    float4 mi = m4[i], vi = v4[i], po;
    po.x = adam_one(pi.x, gi.x, mi.x, vi.x, b1, b2, eps, step, inv_bc2_sqrt);
    po.y = adam_one(pi.y, gi.y, mi.y, vi.y, b1, b2, eps, step, inv_bc2_sqrt);
    po.z = adam_one(pi.z, gi.z, mi.z, vi.z, b1, b2, eps, step, inv_bc2_sqrt);
    po.w = adam_one(pi.w, gi.w, mi.w, vi.w, b1, b2, eps, step, inv_bc2_sqrt);
    m4[i] = mi; v4[i] = vi; p4[i] = po;
    if (q4) q4[i] = make_uint2(pack_bf16x2(po.x, po.y), pack_bf16x2(po.z, po.w));
    if (a4) {
      float4 ai = a4[i];
      const float w = 1.f - ema_decay;
      ai.x = ema_decay * ai.x + w * po.x; ai.y = ema_decay * ai.y + w * po.y;
      ai.z = ema_decay * ai.z + w * po.z; ai.w = ema_decay * ai.w + w * po.w;
      a4[i] = ai;
    }
  }
}

// ------------------------------------------------------------------------------------------ losses
// nn.BCELoss (trainer.py:499; mean reduction, log clamped at -100) over `nvec` probability vectors of length B.
// probs / dprobs are contiguous [nvec][B].
// loss[0] += sum_v weight[v] * BCE(prob[v], target[v]);  dprob[v][b] = weight[v] * (p - t) / max(p (1-p), 1e-12) / B.
__global__ void gan_bce_kernel(const float* __restrict__ probs, const float* __restrict__ targets,
                               const float* __restrict__ weights, int nvec, int B, float* __restrict__ loss,
                               float* __restrict__ dprobs) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < nvec * B; i += blockDim.x) {
    const int v = i / B;
    const float p = probs[i], t = targets[v], w = weights[v];
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(log1pf(-p), -100.f);
    acc += -w * (t * lp + (1.f - t) * l1p) / (float)B;
    if (dprobs) dprobs[i] = w * (p - t) / fmaxf(p * (1.f - p), 1e-12f) / (float)B;
  }
  __shared__ float sh[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] += v;
  }
}

// KL_loss (trainer.py:54-58) * coeff: loss[0] += coeff * -0.5 * mean(1 + lv - mu^2 - exp(lv)); grads (=).
__global__ void kl_loss_kernel(const float* __restrict__ mu, const float* __restrict__ lv, int n, float coeff,
                               float* __restrict__ loss, float* __restrict__ dmu, float* __restrict__ dlv) {
  float acc = 0.f;
  const float inv = 1.f / (float)n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float m = mu[i], l = lv[i], e = __expf(l);
    acc += 1.f + l - m * m - e;
    if (dmu) dmu[i] = coeff * m * inv;
    if (dlv) dlv[i] = coeff * (-0.5f) * (1.f - e) * inv;
  }
  __shared__ float sh[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] += coeff * (-0.5f) * v * inv;
  }
}

// class_aware_loss (trainer.py:298-311): S = X X^T; loss = max(0, mean(S) - mean(S[same class, off-diagonal])) / F.
__global__ void cal_scores_kernel(const float* __restrict__ x, int B, int F, float* __restrict__ S) {
  // fp64 accumulation: the loss is a difference of two means of these scores (cancellation amplifies their rounding)
  const int i = blockIdx.x / B, j = blockIdx.x % B;
  double acc = 0.0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) acc += (double)x[(long long)i * F + f] * (double)x[(long long)j * F + f];
  __shared__ double sh[128];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) S[blockIdx.x] = (float)sh[0];
}
// one block: loss and the symmetric coefficient matrix G = dL/dS + (dL/dS)^T   (B <= 128)
__global__ void cal_coeff_kernel(const float* __restrict__ S, const int* __restrict__ labels, int B, int F,
                                 float* __restrict__ loss, float* __restrict__ G) {
  __shared__ float s_all, s_pair;
  __shared__ int n_pair;
  __shared__ double sh_a[256], sh_p[256];
  __shared__ int sh_n[256];
  double a = 0.0, pr = 0.0;
  int np = 0;
  for (int i = threadIdx.x; i < B * B; i += blockDim.x) {
    const int r = i / B, c = i % B;
    const double v = (double)S[i];
    a += v;
    if (r != c && labels[r] == labels[c]) { pr += v; ++np; }
  }
  // ordered block reduction (fixed tree): no atomics, reproducible; fp64 because the loss is a difference of means
  sh_a[threadIdx.x] = a; sh_p[threadIdx.x] = pr; sh_n[threadIdx.x] = np;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      sh_a[threadIdx.x] += sh_a[threadIdx.x + s]; sh_p[threadIdx.x] += sh_p[threadIdx.x + s]; sh_n[threadIdx.x] += sh_n[threadIdx.x + s];
    }
    __syncthreads();
  }
  __shared__ float l_sh;
  if (threadIdx.x == 0) {
    n_pair = sh_n[0];
    s_all = (float)sh_a[0]; s_pair = (float)sh_p[0];
    l_sh = n_pair > 0 ? (float)(sh_a[0] / (double)(B * B) - sh_p[0] / (double)n_pair) : 0.f;
  }
  __syncthreads();
  const float l = l_sh;
  const bool active = (n_pair > 0) && (l > 0.f);
  if (threadIdx.x == 0 && active) loss[0] += l / (float)F;
  for (int i = threadIdx.x; i < B * B; i += blockDim.x) {
    const int r = i / B, c = i % B;
    float d = 0.f;
    if (active) {
      d = 1.f / (float)(B * B);
      if (r != c && labels[r] == labels[c]) d -= 1.f / (float)n_pair;
      d *= 2.f / (float)F;   // the pair mask is symmetric, so dL/dS + its transpose = 2 dL/dS
    }
    G[i] = d;
  }
}
__global__ void cal_dx_kernel(const float* __restrict__ x, const float* __restrict__ G, int B, int F,
                              float* __restrict__ dx) {
  const long long total = (long long)B * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F), r = (int)(i / F);
    float acc = 0.f;
    for (int j = 0; j < B; ++j) acc += G[r * B + j] * x[(long long)j * F + f];
    dx[i] = acc;
  }
}

static inline unsigned grid1d(long long n, int threads = 256) {
  long long g = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sg2

using namespace sg2;

extern "C" {

int sg2_linear_fwd(const float* x1, int K1, const float* x2, int K2, const float* w, const float* bias, void* out,
                   int out_bf16, int M, int N, void* stream) {
  if (M < 1 || M > 4096) SG2_FAIL(SG2_EINVAL, "linear_fwd: M=%d", M);
  if (K1 + K2 <= 1024) {
    const int Kp = (K1 + K2 + 31) / 32 * 32;
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
      cudaFuncSetAttribute(linear_fwd_rt_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4);
      cudaFuncSetAttribute(linear_fwd_rt_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4);
      cudaFuncSetAttribute(linear_fwd_rt_kernel<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4);
      cudaFuncSetAttribute(linear_fwd_rt_kernel<false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4);
      attr_done[dev] = true;
    }
    const int ngroups = (N + kRtF - 1) / kRtF;
    int blocks = (ngroups + 1) / 2;                 // two feature slots per block
    if (blocks > 148 * 2) blocks = 148 * 2;
    const bool r6 = M <= 24;                        // four m-groups of 6 rows (batch 24) instead of 8
    const size_t smem = (size_t)(r6 ? 24 : 32) * Kp * sizeof(float);
    auto s_ = (cudaStream_t)stream;
    if (out_bf16) {
      if (r6) linear_fwd_rt_kernel<true, 6><<<blocks, 256, smem, s_>>>(x1, K1, x2, K2, w, bias, out, M, N, Kp);
      else linear_fwd_rt_kernel<true, 8><<<blocks, 256, smem, s_>>>(x1, K1, x2, K2, w, bias, out, M, N, Kp);
    } else {
      if (r6) linear_fwd_rt_kernel<false, 6><<<blocks, 256, smem, s_>>>(x1, K1, x2, K2, w, bias, out, M, N, Kp);
      else linear_fwd_rt_kernel<false, 8><<<blocks, 256, smem, s_>>>(x1, K1, x2, K2, w, bias, out, M, N, Kp);
    }
    SG2_LAUNCH_OK("linear_fwd_rt");
  }
  // feature groups per warp: enough to amortise the activation staging, few enough to fill the SMs (8 warps per block)
  int npw = N / (kLinF * 8 * 148 * 4);
  npw = npw < 1 ? 1 : (npw > 8 ? 8 : npw);
  dim3 grid((N + kLinF * 8 * npw - 1) / (kLinF * 8 * npw)), block(256);
  if (out_bf16)
    linear_fwd_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(x1, K1, x2, K2, w, bias, out, M, N, npw);
  else
    linear_fwd_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(x1, K1, x2, K2, w, bias, out, M, N, npw);
  SG2_LAUNCH_OK("linear_fwd");
}

int sg2_linear_bwd_w(const void* dy, int dy_bf16, const float* x1, int K1, const float* x2, int K2, float* dw,
                     float* dbias, int M, int N, int accumulate, void* stream) {
  const int K = K1 + K2;
  dim3 grid((N + kLinNB - 1) / kLinNB, (K + 255) / 256);
  for (int m0 = 0; m0 < M; m0 += 32) {  // batch rows in slabs of 32 (the kernel keeps one activation column in registers)
    const int mc = M - m0 < 32 ? M - m0 : 32;
    const int acc = (accumulate || m0 > 0) ? 1 : 0;
    const float* x1o = x1 + (size_t)m0 * K1;
    const float* x2o = x2 ? x2 + (size_t)m0 * K2 : nullptr;
    if (dy_bf16)
      linear_bwd_w_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)m0 * N, x1o, K1, x2o, K2, dw, dbias, mc, N, acc);
    else
      linear_bwd_w_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(dy) + (size_t)m0 * N,
                                                                         x1o, K1, x2o, K2, dw, dbias, mc, N, acc);
  }
  SG2_LAUNCH_OK("linear_bwd_w");
}

int sg2_linear_bwd_x_scratch_floats(int N, int Kout) { return ((N + kLinNB - 1) / kLinNB) * 32 * Kout; }

int sg2_linear_bwd_x(const void* dy, int dy_bf16, const float* w, float* dx, float* scratch, int M, int N, int K,
                     int Kout, void* stream) {
  if (Kout > K) SG2_FAIL(SG2_EINVAL, "linear_bwd_x: Kout=%d", Kout);
  if (!scratch) SG2_FAIL(SG2_EINVAL, "linear_bwd_x: scratch of sg2_linear_bwd_x_scratch_floats(N, Kout) floats required");
  int threads = ((Kout + 31) / 32) * 32;
  threads = threads > 256 ? 256 : threads;
  const int nslabs = (N + kLinNB - 1) / kLinNB;
  for (int m0 = 0; m0 < M; m0 += 32) {
    const int mc = M - m0 < 32 ? M - m0 : 32;
    if (dy_bf16)
      linear_bwd_x_kernel<true><<<nslabs, threads, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)m0 * N, w, scratch, mc, N, K, Kout);
    else
      linear_bwd_x_kernel<false><<<nslabs, threads, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const float*>(dy) + (size_t)m0 * N, w, scratch, mc, N, K, Kout);
    linear_bwd_x_reduce_kernel<<<dim3((Kout + 31) / 32, mc), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        scratch, dx + (size_t)m0 * Kout, nslabs, mc, Kout);
  }
  SG2_LAUNCH_OK("linear_bwd_x");
}

int sg2_ca_glu_reparam_fwd(const float* fc, const float* eps, float* mu, float* logvar, float* c, int B, int E,
                           void* stream) {
  ca_glu_reparam_fwd_kernel<<<(B * E + 255) / 256, 256, 0, (cudaStream_t)stream>>>(fc, eps, mu, logvar, c, B, E);
  SG2_LAUNCH_OK("ca_glu_reparam_fwd");
}

int sg2_ca_glu_reparam_bwd(const float* fc, const float* eps, const float* dmu, const float* dlogvar, const float* dc,
                           float* dfc, int B, int E, void* stream) {
  ca_glu_reparam_bwd_kernel<<<(B * E + 255) / 256, 256, 0, (cudaStream_t)stream>>>(fc, eps, dmu, dlogvar, dc, dfc, B,
                                                                                  E);
  SG2_LAUNCH_OK("ca_glu_reparam_bwd");
}

int sg2_chw_hwc_bf16(const void* in, void* out, int B, int C, int HW, int to_hwc, void* stream) {
  chw_to_hwc_bf16_kernel<<<grid1d((long long)B * C * HW), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in, (__nv_bfloat16*)out, B, C, HW, to_hwc);
  SG2_LAUNCH_OK("chw_hwc_bf16");
}

int sg2_logits_fwd(const void* x, const float* w, const float* bias, float* prob, int B, int HW, int C,
                   void* stream) {
  if (C % 8) SG2_FAIL(SG2_EINVAL, "logits_fwd: C=%d not a multiple of 8", C);
  logits_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, w, bias, prob, HW, C);
  SG2_LAUNCH_OK("logits_fwd");
}

int sg2_logits_bwd_scratch_floats(int B, int HW, int C) { return ((B + kLgChunk - 1) / kLgChunk) * HW * C; }

int sg2_logits_bwd(const float* dprob, const float* prob, const void* x, const float* w, void* dx, int dx_accumulate,
                   float* dw, float* dbias, float* dw_scratch, int B, int HW, int C, void* stream) {
  if (C % 8) SG2_FAIL(SG2_EINVAL, "logits_bwd: C=%d not a multiple of 8", C);
  if (dw && !dw_scratch) SG2_FAIL(SG2_EINVAL, "logits_bwd: dw needs sg2_logits_bwd_scratch_floats() floats of scratch");
  const int nv = HW * C / 8;
  const int chunks = (B + kLgChunk - 1) / kLgChunk;
  logits_bwd_kernel<<<dim3((nv + 127) / 128, chunks), 128, 0, (cudaStream_t)stream>>>(
      dprob, prob, (const __nv_bfloat16*)x, w, (__nv_bfloat16*)dx, dx_accumulate, dw ? dw_scratch : nullptr, dbias, B, HW, C);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) SG2_FAIL((int)e, "logits_bwd launch: %s", cudaGetErrorString(e));
  // dw += the chunk partials in chunk order
  if (dw) return sg2_reduce_slabs(dw_scratch, chunks, (long long)HW * C, (long long)HW * C, dw, 1, stream);
  return 0;
}

int sg2_gan_bce(const float* probs, const float* targets, const float* weights, int nvec, int B, float* loss,
                float* dprobs, void* stream) {
  gan_bce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(probs, targets, weights, nvec, B, loss, dprobs);
  SG2_LAUNCH_OK("gan_bce");
}

int sg2_kl_loss(const float* mu, const float* logvar, int n, float coeff, float* loss, float* dmu, float* dlogvar,
                void* stream) {
  kl_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(mu, logvar, n, coeff, loss, dmu, dlogvar);
  SG2_LAUNCH_OK("kl_loss");
}

int sg2_cal_loss(const float* x, const int* labels, int B, int F, float* ws /* 2*B*B floats */, float* loss,
                 float* dx, void* stream) {
  if (B > 1024) SG2_FAIL(SG2_EINVAL, "cal_loss: B=%d", B);
  cudaStream_t st = (cudaStream_t)stream;
  cal_scores_kernel<<<B * B, 128, 0, st>>>(x, B, F, ws);
  cal_coeff_kernel<<<1, 256, 0, st>>>(ws, labels, B, F, loss, ws + (size_t)B * B);
  if (dx) cal_dx_kernel<<<grid1d((long long)B * F), 256, 0, st>>>(x, ws + (size_t)B * B, B, F, dx);
  SG2_LAUNCH_OK("cal_loss");
}

int sg2_joint_bias(const float* c, const float* w, long long so, long long se, long long st, float* bias9, int B, int E,
                   int Cout, void* stream) {
  joint_bias_kernel<<<dim3(Cout, B), 288, 0, (cudaStream_t)stream>>>(c, w, so, se, st, bias9, E, Cout);
  SG2_LAUNCH_OK("joint_bias");
}

int sg2_joint_tap_sums(const void* dy, double* R, float* S, int B, int H, int W, int Cout, void* stream) {
  if (Cout % 8 || Cout > 1024) SG2_FAIL(SG2_EINVAL, "joint_tap_sums: Cout=%d", Cout);
  if (H < 2 || W < 2) SG2_FAIL(SG2_EINVAL, "joint_tap_sums: %dx%d image", H, W);
  cudaStream_t st = (cudaStream_t)stream;
  const int vc = Cout / 8;
  const int threads = vc >= 256 ? vc : (256 / vc) * vc;
  joint_region_sums_kernel<<<dim3(H, B), threads, 0, st>>>((const uint4*)dy, R, H, W, Cout);
  const int n = B * 9 * Cout;
  joint_tap_sums_kernel<<<(n + 255) / 256, 256, 0, st>>>(R, S, B, Cout);
  SG2_LAUNCH_OK("joint_tap_sums");
}

int sg2_joint_c_bwd(const float* S, const float* c, const float* w, long long so, long long se, long long st, float* dc,
                    float* dw, int dw_accumulate, int B, int E, int Cout, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (dw) joint_dw_kernel<<<dim3(Cout, 9), 128, 0, s>>>(S, c, dw, so, se, st, dw_accumulate, B, E, Cout);
  if (dc) {
    const int ept = E >= 128 ? 128 : (E >= 64 ? 64 : 32);
    const size_t smem = (size_t)(9 * Cout + kJointSlices * ept) * sizeof(float);
    if (smem > 48 * 1024) SG2_FAIL(SG2_EINVAL, "joint_c_bwd: Cout=%d too large", Cout);
    joint_dc_kernel<<<B, kJointSlices * ept, smem, s>>>(S, w, so, se, st, dc, E, Cout);
  }
  SG2_LAUNCH_OK("joint_c_bwd");
}

int sg2_adam_tick(int* step, float* bc, float beta1, float beta2, void* stream) {
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, bc, beta1, beta2);
  SG2_LAUNCH_OK("adam_tick");
}

int sg2_adam_ema(float* p, const float* g, float* m, float* v, float* avg, long long n, float lr, float beta1,
                 float beta2, float eps, const float* bc, float ema_decay, void* p_bf16, void* stream) {
  // elements to the next 16-byte boundary; every array must agree (they are slices of identically laid out buckets)
  auto mis = [](const void* q, int elem) { return (int)(((uintptr_t)q / elem) & 3); };
  int head = (4 - mis(p, 4)) & 3;
  const bool aligned = ((uintptr_t)p % 4 == 0) && mis(g, 4) == mis(p, 4) && mis(m, 4) == mis(p, 4) &&
                       mis(v, 4) == mis(p, 4) && (!avg || mis(avg, 4) == mis(p, 4)) &&
                       (!p_bf16 || ((uintptr_t)p_bf16 % 2 == 0 && mis(p_bf16, 2) == mis(p, 4)));
  if (!aligned) head = -1;
  adam_ema_kernel<<<grid1d((n + 3) / 4 + 8), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, avg, n, lr, beta1, beta2, eps, bc, ema_decay, (__nv_bfloat16*)p_bf16, head);
  SG2_LAUNCH_OK("adam_ema");
}

}  // extern "C"
