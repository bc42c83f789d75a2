// fp32-accurate mode ("precise"): the north star's second precision (relative error <= 1e-4 against the fp32 reference).
//
// The convolutions stay on the bf16 tcgen05 kernels: an fp32 operand is split into three bf16 planes
// x = hi + mid + lo (8 + 8 + 8 mantissa bits), and a product of two such operands is the sum of the six cross terms
// (hi,hi) (hi,mid) (mid,hi) (hi,lo) (lo,hi) (mid,mid), each an ordinary bf16 x bf16 -> fp32 implicit GEMM accumulated
// into the same fp32 output; the dropped terms are O(2^-24) of the product. TF32 (10 mantissa bits) cannot meet 1e-4
// against an fp32 reference; the 3-way split can, on the same tensor-core path (no CUDA-core conv, no library call).
//
// Everything between the convolutions — BatchNorm statistics / apply / backward with GLU, LeakyReLU(0.2) or residual,
// the c_code concat, heads, stems, layout changes, D logits — runs here in plain fp32 on NHWC fp32 tensors, with fp64
// accumulation for the cross-row sums. These kernels are written for exactness, not for bandwidth; the bf16 kernels in
// elementwise.cu are the fast path.
// Reference constructs: model.py:112-169 (GLU, upBlock, Block3x3_relu, ResBlock), 287-298 (heads), 358-445 (D blocks).
#include <cstdio>

#include "../../include/sg2b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sg2 {

static inline unsigned pgrid(long long n, int threads = 256) {
  long long g = (n + threads - 1) / threads;
  const long long cap = 148LL * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
#define P_LOOP(i, n) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__device__ __forceinline__ float p_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------------------------------ operand split
// out[part][i]: part 0 = bf16(x), 1 = bf16(x - hi), 2 = bf16(x - hi - mid)
__global__ void split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n) {
  P_LOOP(i, n) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    out[i] = h;
    out[n + i] = m;
    out[2 * n + i] = __float2bfloat16_rn(r2);
  }
}

// ------------------------------------------------------------------------------------------ BatchNorm statistics
// stats[g][2][C] (fp64, +=): per-channel sum / sum of squares of rows [g*P, (g+1)*P). grid (C/32, row blocks, groups),
// block (32, 8).
__global__ void bn_stats_f32_kernel(const float* __restrict__ x, long long P, int C, double* __restrict__ stats) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  x += (long long)blockIdx.z * P * C;
  stats += (long long)blockIdx.z * 2 * C;
  double s = 0.0, q = 0.0;
  if (c < C)
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < P; r += (long long)gridDim.y * 8) {
      const double v = (double)x[r * C + c];
      s += v;
      q += v * v;
    }
  __shared__ double sh[2][8][32];
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int k = 1; k < 8; ++k) { s += sh[0][k][threadIdx.x]; q += sh[1][k][threadIdx.x]; }
    atomicAdd(&stats[c], s);
    atomicAdd(&stats[C + c], q);
  }
}

// mean / rstd of every (group, channel) from the sums; running statistics (momentum, unbiased variance) updated once
// per group in group order, num_batches_tracked += groups — nn.BatchNorm train-mode semantics.
__global__ void bn_finalize_f32_kernel(const double* __restrict__ stats, long long P, int C, int groups, float eps,
                                       float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                                       float* __restrict__ rmean, float* __restrict__ rvar, long long* __restrict__ nbt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += groups;
  if (c >= C) return;
  float rm = rmean ? rmean[c] : 0.f, rv = rvar ? rvar[c] : 0.f;
  for (int g = 0; g < groups; ++g) {
    const double m = stats[(long long)g * 2 * C + c] / (double)P;
    double var = stats[(long long)g * 2 * C + C + c] / (double)P - m * m;
    if (var < 0) var = 0;
    mean[(long long)g * C + c] = (float)m;
    rstd[(long long)g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
    const double unb = P > 1 ? var * ((double)P / (double)(P - 1)) : var;
    rm = (1.f - momentum) * rm + momentum * (float)m;
    rv = (1.f - momentum) * rv + momentum * (float)unb;
  }
  if (rmean) { rmean[c] = rm; rvar[c] = rv; }
}

// out = act(bn(x)) (+ residual); act 0 none, 1 GLU (out has C/2 channels), 2 LeakyReLU(0.2). has_bn = 0: no BatchNorm.
__global__ void bn_act_fwd_f32_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                      const float* __restrict__ rstd, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ residual,
                                      float* __restrict__ out, long long P, int C, int groups, int act, int has_bn) {
  const int Co = act == 1 ? C / 2 : C;
  const long long total = P * groups * Co;
  P_LOOP(i, total) {
    const int c = (int)(i % Co);
    const long long r = i / Co;
    const int g = (int)(r / P);
    const float* xr = x + r * C;
    auto bn = [&](int ch) {
      const float v = xr[ch];
      if (!has_bn) return v;
      return (v - mean[(long long)g * C + ch]) * rstd[(long long)g * C + ch] * gamma[ch] + beta[ch];
    };
    float o;
    if (act == 1) o = bn(c) * p_sigmoid(bn(c + Co));
    else if (act == 2) { const float z = bn(c); o = z > 0.f ? z : 0.2f * z; }
    else { o = bn(c); if (residual) o += residual[i]; }
    out[i] = o;
  }
}

// dz of input channel ch of row r (the gradient w.r.t. the BatchNorm OUTPUT of that channel)
__device__ __forceinline__ float bn_dz_f32(const float* __restrict__ xr, const float* __restrict__ dr, int ch, int C,
                                           int act, int has_bn, const float* mean, const float* rstd,
                                           const float* gamma, const float* beta, long long gC) {
  auto bn = [&](int k) {
    const float v = xr[k];
    if (!has_bn) return v;
    return (v - mean[gC + k]) * rstd[gC + k] * gamma[k] + beta[k];
  };
  if (act == 1) {
    const int Co = C / 2;
    if (ch < Co) return dr[ch] * p_sigmoid(bn(ch + Co));
    const float s = p_sigmoid(bn(ch));
    return dr[ch - Co] * bn(ch - Co) * s * (1.f - s);
  }
  if (act == 2) return bn(ch) > 0.f ? dr[ch] : 0.2f * dr[ch];
  return dr[ch];
}

// sums[g][2][C] (fp64, +=): sum dz, sum dz * xhat. grid (C/32, row blocks, groups), block (32, 8).
__global__ void bn_bwd_reduce_f32_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, long long P,
                                         int C, int act, double* __restrict__ sums) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int g = blockIdx.z;
  const int Co = act == 1 ? C / 2 : C;
  x += (long long)g * P * C;
  dout += (long long)g * P * Co;
  sums += (long long)g * 2 * C;
  const long long gC = (long long)g * C;
  double s = 0.0, t = 0.0;
  if (c < C)
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < P; r += (long long)gridDim.y * 8) {
      const float dz = bn_dz_f32(x + r * C, dout + r * Co, c, C, act, 1, mean, rstd, gamma, beta, gC);
      const float xh = (x[r * C + c] - mean[gC + c]) * rstd[gC + c];
      s += (double)dz;
      t += (double)dz * (double)xh;
    }
  __shared__ double sh[2][8][32];
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int k = 1; k < 8; ++k) { s += sh[0][k][threadIdx.x]; t += sh[1][k][threadIdx.x]; }
    atomicAdd(&sums[c], s);
    atomicAdd(&sums[C + c], t);
  }
}

__global__ void bn_bwd_params_f32_kernel(const double* __restrict__ sums, int C, int groups, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double db = 0.0, dg = 0.0;
  for (int g = 0; g < groups; ++g) { db += sums[(long long)g * 2 * C + c]; dg += sums[(long long)g * 2 * C + C + c]; }
  if (accumulate) { dgamma[c] += (float)dg; dbeta[c] += (float)db; } else { dgamma[c] = (float)dg; dbeta[c] = (float)db; }
}

// dx = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat))   (has_bn = 0: dx = dz)
__global__ void bn_bwd_apply_f32_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const double* __restrict__ sums, long long P, int C, int groups, int act,
                                        int has_bn, float* __restrict__ dx) {
  const int Co = act == 1 ? C / 2 : C;
  const long long total = P * groups * C;
  P_LOOP(i, total) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int g = (int)(r / P);
    const long long gC = (long long)g * C;
    const float dz = bn_dz_f32(x + r * C, dout + r * Co, c, C, act, has_bn, mean, rstd, gamma, beta, gC);
    if (!has_bn) { dx[i] = dz; continue; }
    const float m_dz = (float)(sums[(long long)g * 2 * C + c] / (double)P);
    const float m_dzx = (float)(sums[(long long)g * 2 * C + C + c] / (double)P);
    const float xh = (x[i] - mean[gC + c]) * rstd[gC + c];
    dx[i] = gamma[c] * rstd[gC + c] * (dz - m_dz - xh * m_dzx);
  }
}

// mode 0: out = a + b;  1: out = (a > 0 ? b : 0.2 b)  (LeakyReLU backward: a = the activation output / input, b = dout)
__global__ void ew_f32_kernel(int mode, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                              long long n) {
  P_LOOP(i, n) out[i] = mode == 0 ? a[i] + b[i] : (a[i] > 0.f ? b[i] : 0.2f * b[i]);
}

// bias9 / activation applied to an fp32 conv output in place: y[b][y][x][c] += bias9[b][class][c]; act 2: LeakyReLU
__global__ void conv_post_f32_kernel(float* __restrict__ y, const float* __restrict__ bias9, int act, int B, int H, int W,
                                     int C) {
  const long long total = (long long)B * H * W * C;
  P_LOOP(i, total) {
    float v = y[i];
    if (bias9) {
      const int c = (int)(i % C);
      const int xx = (int)((i / C) % W), yy = (int)((i / ((long long)C * W)) % H), b = (int)(i / ((long long)C * W * H));
      const int ry = yy == 0 ? 0 : (yy == H - 1 ? 2 : 1), rx = xx == 0 ? 0 : (xx == W - 1 ? 2 : 1);
      v += bias9[((long long)b * 9 + ry * 3 + rx) * C + c];
    }
    if (act == 2) v = v > 0.f ? v : 0.2f * v;
    y[i] = v;
  }
}

// ------------------------------------------------------------------------------------------ c_code concat
__global__ void concat_c_f32_kernel(const float* __restrict__ c, const float* __restrict__ h, float* __restrict__ out,
                                    int B, int HW, int E, int Ch) {
  const int Ct = E + Ch;
  const long long total = (long long)B * HW * Ct;
  P_LOOP(i, total) {
    const int k = (int)(i % Ct);
    const long long pix = i / Ct;
    out[i] = k < E ? c[(pix / HW) * E + k] : h[pix * Ch + (k - E)];
  }
}
// dh = dcat[..., E:]
__global__ void concat_c_bwd_dh_f32_kernel(const float* __restrict__ dcat, float* __restrict__ dh, long long npix, int E,
                                           int Ch) {
  const long long total = npix * Ch;
  P_LOOP(i, total) dh[i] = dcat[(i / Ch) * (E + Ch) + E + (i % Ch)];
}
// dc[b][e] += sum over the sample's pixels (fp64 accumulation, one thread per (b, e): deterministic)
__global__ void concat_c_bwd_dc_f32_kernel(const float* __restrict__ dcat, float* __restrict__ dc, int B, int HW, int E,
                                           int Ch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * E) return;
  const int b = i / E, e = i % E;
  double s = 0.0;
  for (int p = 0; p < HW; ++p) s += (double)dcat[((long long)b * HW + p) * (E + Ch) + e];
  dc[i] += (float)s;
}

// ------------------------------------------------------------------------------------------ heads / stems / layouts
__global__ void head_tanh_fwd_f32_kernel(const float* __restrict__ y, float* __restrict__ img, int B, int HW, int CP) {
  const long long total = (long long)B * 3 * HW;
  P_LOOP(i, total) {
    const int p = (int)(i % HW), c = (int)((i / HW) % 3);
    const long long b = i / (3LL * HW);
    img[i] = tanhf(y[(b * HW + p) * CP + c]);
  }
}
__global__ void head_tanh_bwd_f32_kernel(const float* __restrict__ dimg, const float* __restrict__ img,
                                         float* __restrict__ dy, int B, int HW, int CP) {
  const long long total = (long long)B * HW * CP;
  P_LOOP(i, total) {
    const int c = (int)(i % CP);
    const long long pix = i / CP;
    float v = 0.f;
    if (c < 3) {
      const long long o = ((pix / HW) * 3 + c) * HW + (pix % HW);
      const float t = img[o];
      v = dimg[o] * (1.f - t * t);
    }
    dy[i] = v;
  }
}
// im2col rows [B*(S/2)^2][64] of the 4x4 s2 p1 stem, k = (kh*4+kw)*3 + c (48 used)
__global__ void stem_im2col_f32_kernel(const float* __restrict__ img, float* __restrict__ col, int B, int S) {
  const int So = S / 2;
  const long long total = (long long)B * So * So * 64;
  P_LOOP(i, total) {
    const int k = (int)(i & 63);
    const long long pix = i >> 6;
    const int ox = (int)(pix % So), oy = (int)((pix / So) % So);
    const long long b = pix / ((long long)So * So);
    float v = 0.f;
    if (k < 48) {
      const int c = k % 3, t = k / 3, kh = t >> 2, kw = t & 3;
      const int y = 2 * oy + kh - 1, x = 2 * ox + kw - 1;
      if (y >= 0 && y < S && x >= 0 && x < S) v = img[((b * 3 + c) * S + y) * S + x];
    }
    col[i] = v;
  }
}
__global__ void stem_col2im_f32_kernel(const float* __restrict__ dcol, float* __restrict__ dimg, int B, int S) {
  const int So = S / 2;
  const long long total = (long long)B * 3 * S * S;
  P_LOOP(i, total) {
    const int x = (int)(i % S), y = (int)((i / S) % S), c = (int)((i / ((long long)S * S)) % 3);
    const long long b = i / (3LL * S * S);
    float acc = 0.f;
    for (int kh = 0; kh < 4; ++kh) {
      const int ty = y + 1 - kh;
      if (ty < 0 || (ty & 1) || (ty >> 1) >= So) continue;
      for (int kw = 0; kw < 4; ++kw) {
        const int tx = x + 1 - kw;
        if (tx < 0 || (tx & 1) || (tx >> 1) >= So) continue;
        acc += dcol[((b * So + (ty >> 1)) * So + (tx >> 1)) * 64 + (kh * 4 + kw) * 3 + c];
      }
    }
    dimg[i] = acc;
  }
}
// [B][HW][C] <-> [B][C][HW], fp32 both sides (to_chw = 1: NHWC -> NCHW)
__global__ void hwc_chw_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int HW, int C, int to_chw) {
  const long long total = (long long)B * HW * C;
  P_LOOP(i, total) {
    const int c = (int)(i % C), p = (int)((i / C) % HW);
    const long long b = i / ((long long)C * HW);
    const long long j = (b * C + c) * HW + p;
    if (to_chw) out[j] = in[i]; else out[i] = in[j];
  }
}

// ------------------------------------------------------------------------------------------ D logits (fp32 x)
__global__ void logits_fwd_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ bias, float* __restrict__ prob, int HW, int C) {
  const int b = blockIdx.x, n = HW * C;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int p = i / C, c = i - p * C;
    acc += (double)x[(long long)b * n + i] * (double)w[c * HW + p];
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) prob[b] = 1.f / (1.f + expf(-((float)sh[0] + bias[0])));
}
// thread per (p, c): dx[b] (=|+=) dpre[b] w;  dw += sum_b dpre[b] x[b];  dbias += sum dpre
__global__ void logits_bwd_f32_kernel(const float* __restrict__ dprob, const float* __restrict__ prob,
                                      const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ dx,
                                      int dx_accumulate, float* __restrict__ dw, float* __restrict__ dbias, int B, int HW,
                                      int C) {
  const int n = HW * C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (dbias && i == 0) {
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += (double)dprob[b] * prob[b] * (1.f - prob[b]);
    dbias[0] += (float)s;
  }
  if (i >= n) return;
  const int p = i / C, c = i - p * C;
  const float wv = w[c * HW + p];
  double gw = 0.0;
  for (int b = 0; b < B; ++b) {
    const float dpre = dprob[b] * prob[b] * (1.f - prob[b]);
    gw += (double)dpre * (double)x[(long long)b * n + i];
    if (dx) {
      float* d = dx + (long long)b * n + i;
      *d = dx_accumulate ? *d + dpre * wv : dpre * wv;
    }
  }
  if (dw) dw[c * HW + p] += (float)gw;
}

// S[b][tap][o] for the folded jointConv is not used in precise mode (the concat is materialised).

}  // namespace sg2

using namespace sg2;
#define PF_FAIL SG2_FAIL

extern "C" {

int sg2_split3(const float* x, void* out, long long n, void* stream) {
  split3_kernel<<<pgrid(n), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out, n);
  SG2_LAUNCH_OK("split3");
}

int sg2_bn_stats_f32(const float* x, long long P, int C, int groups, double* stats, void* stream) {
  if (groups < 1 || P % groups) PF_FAIL(SG2_EINVAL, "bn_stats_f32: %lld rows in %d groups", P, groups);
  P /= groups;
  long long rb = (P + 63) / 64;
  if (rb > 256) rb = 256;
  bn_stats_f32_kernel<<<dim3((C + 31) / 32, (unsigned)rb, groups), dim3(32, 8), 0, (cudaStream_t)stream>>>(x, P, C, stats);
  SG2_LAUNCH_OK("bn_stats_f32");
}

int sg2_bn_act_fwd_f32(const float* x, const double* stats, float* mean, float* rstd, const float* gamma,
                       const float* beta, const float* residual, float* out, long long P, int C, int groups, int act,
                       float eps, float momentum, float* running_mean, float* running_var,
                       long long* num_batches_tracked, void* stream) {
  if (groups < 1 || P % groups) PF_FAIL(SG2_EINVAL, "bn_act_fwd_f32: %lld rows in %d groups", P, groups);
  if (act == 1 && (C & 1)) PF_FAIL(SG2_EINVAL, "bn_act_fwd_f32: GLU needs an even channel count");
  P /= groups;
  cudaStream_t st = (cudaStream_t)stream;
  const int has_bn = mean != nullptr;
  if (stats) {
    if (!mean) PF_FAIL(SG2_EINVAL, "bn_act_fwd_f32: stats given without mean/rstd outputs");
    bn_finalize_f32_kernel<<<(C + 127) / 128, 128, 0, st>>>(stats, P, C, groups, eps, momentum, mean, rstd, running_mean,
                                                            running_var, num_batches_tracked);
  }
  const int Co = act == 1 ? C / 2 : C;
  bn_act_fwd_f32_kernel<<<pgrid(P * groups * Co), 256, 0, st>>>(x, mean, rstd, gamma, beta, residual, out, P, C, groups,
                                                                act, has_bn);
  SG2_LAUNCH_OK("bn_act_fwd_f32");
}

int sg2_bn_act_bwd_f32(const float* x, const float* dout, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, double* sums, float* dx, float* dgamma, float* dbeta, int accumulate,
                       long long P, int C, int groups, int act, void* stream) {
  if (groups < 1 || P % groups) PF_FAIL(SG2_EINVAL, "bn_act_bwd_f32: %lld rows in %d groups", P, groups);
  P /= groups;
  cudaStream_t st = (cudaStream_t)stream;
  const int has_bn = mean != nullptr;
  if (has_bn) {
    long long rb = (P + 63) / 64;
    if (rb > 256) rb = 256;
    bn_bwd_reduce_f32_kernel<<<dim3((C + 31) / 32, (unsigned)rb, groups), dim3(32, 8), 0, st>>>(x, dout, mean, rstd, gamma,
                                                                                             beta, P, C, act, sums);
    if (dgamma) bn_bwd_params_f32_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, groups, dgamma, dbeta, accumulate);
  }
  bn_bwd_apply_f32_kernel<<<pgrid(P * groups * C), 256, 0, st>>>(x, dout, mean, rstd, gamma, beta, sums, P, C, groups, act,
                                                                 has_bn, dx);
  SG2_LAUNCH_OK("bn_act_bwd_f32");
}

int sg2_ew_f32(int mode, const float* a, const float* b, float* out, long long n, void* stream) {
  if (mode != 0 && mode != 1) PF_FAIL(SG2_EINVAL, "ew_f32: mode %d", mode);
  ew_f32_kernel<<<pgrid(n), 256, 0, (cudaStream_t)stream>>>(mode, a, b, out, n);
  SG2_LAUNCH_OK("ew_f32");
}

int sg2_conv_post_f32(float* y, const float* bias9, int act, int B, int H, int W, int C, void* stream) {
  conv_post_f32_kernel<<<pgrid((long long)B * H * W * C), 256, 0, (cudaStream_t)stream>>>(y, bias9, act, B, H, W, C);
  SG2_LAUNCH_OK("conv_post_f32");
}

int sg2_concat_c_f32(const float* c, const float* h, float* out, int B, int HW, int E, int Ch, void* stream) {
  concat_c_f32_kernel<<<pgrid((long long)B * HW * (E + Ch)), 256, 0, (cudaStream_t)stream>>>(c, h, out, B, HW, E, Ch);
  SG2_LAUNCH_OK("concat_c_f32");
}

int sg2_concat_c_bwd_f32(const float* dcat, float* dh, float* dc, int B, int HW, int E, int Ch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dh) concat_c_bwd_dh_f32_kernel<<<pgrid((long long)B * HW * Ch), 256, 0, st>>>(dcat, dh, (long long)B * HW, E, Ch);
  if (dc) concat_c_bwd_dc_f32_kernel<<<(B * E + 127) / 128, 128, 0, st>>>(dcat, dc, B, HW, E, Ch);
  SG2_LAUNCH_OK("concat_c_bwd_f32");
}

int sg2_head_tanh_fwd_f32(const float* y, float* img, int B, int HW, int CP, void* stream) {
  head_tanh_fwd_f32_kernel<<<pgrid((long long)B * 3 * HW), 256, 0, (cudaStream_t)stream>>>(y, img, B, HW, CP);
  SG2_LAUNCH_OK("head_tanh_fwd_f32");
}

int sg2_head_tanh_bwd_f32(const float* dimg, const float* img, float* dy, int B, int HW, int CP, void* stream) {
  head_tanh_bwd_f32_kernel<<<pgrid((long long)B * HW * CP), 256, 0, (cudaStream_t)stream>>>(dimg, img, dy, B, HW, CP);
  SG2_LAUNCH_OK("head_tanh_bwd_f32");
}

int sg2_stem_im2col_f32(const float* img, float* col, int B, int S, void* stream) {
  if (S % 2) PF_FAIL(SG2_EINVAL, "stem_im2col_f32: odd image size");
  stem_im2col_f32_kernel<<<pgrid((long long)B * (S / 2) * (S / 2) * 64), 256, 0, (cudaStream_t)stream>>>(img, col, B, S);
  SG2_LAUNCH_OK("stem_im2col_f32");
}

int sg2_stem_col2im_f32(const float* dcol, float* dimg, int B, int S, void* stream) {
  stem_col2im_f32_kernel<<<pgrid((long long)B * 3 * S * S), 256, 0, (cudaStream_t)stream>>>(dcol, dimg, B, S);
  SG2_LAUNCH_OK("stem_col2im_f32");
}

int sg2_hwc_chw_f32(const float* in, float* out, int B, int HW, int C, int to_chw, void* stream) {
  hwc_chw_f32_kernel<<<pgrid((long long)B * HW * C), 256, 0, (cudaStream_t)stream>>>(in, out, B, HW, C, to_chw);
  SG2_LAUNCH_OK("hwc_chw_f32");
}

int sg2_logits_fwd_f32(const float* x, const float* w, const float* bias, float* prob, int B, int HW, int C, void* stream) {
  logits_fwd_f32_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, w, bias, prob, HW, C);
  SG2_LAUNCH_OK("logits_fwd_f32");
}

int sg2_logits_bwd_f32(const float* dprob, const float* prob, const float* x, const float* w, float* dx,
                       int dx_accumulate, float* dw, float* dbias, int B, int HW, int C, void* stream) {
  logits_bwd_f32_kernel<<<(HW * C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dprob, prob, x, w, dx, dx_accumulate, dw,
                                                                                dbias, B, HW, C);
  SG2_LAUNCH_OK("logits_bwd_f32");
}

}  // extern "C"
