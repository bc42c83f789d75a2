// Bandwidth-bound kernels around the convolutions: weight packing, BatchNorm statistics / apply fused with
// GLU / LeakyReLU / residual (forward and backward), c_code concat, image-head tanh, layout casts.
// All activations are NHWC bf16 viewed as [P pixels][C channels]; every thread moves 128-bit vectors
// (8 channels) and consecutive threads touch consecutive 16-byte segments (fully coalesced).
// Reference constructs: nn.BatchNorm2d/1d + GLU (model.py:112-122,133-150), LeakyReLU (model.py:358-376),
// torch.cat of the broadcast c_code (model.py:274-277,431-434), nn.Tanh heads (model.py:291-294).
#include <cstdio>

#include "../../include/sg2b200.h"
#include "common.cuh"
#include "ptx.cuh"

#define EW_FAIL SG2_FAIL

namespace sg2 {

static inline int launch_ok(const char* what) { SG2_LAUNCH_OK(what); }

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// Column-vector geometry shared by the [P][C] kernels: a block covers `cpb` 8-channel columns x `rpb` pixel rows.
struct Geo {
  int vc, cpb, rpb;
  dim3 grid, block;
};
static Geo make_geo(long long P, int C, int max_row_blocks, int width = 8, int min_iters = 8) {
  Geo g;
  g.vc = C / width;
  g.cpb = g.vc < 256 ? g.vc : 256;
  // round cpb down to a power of two that divides vc? vc is a multiple of 4 for every layer here; keep generic:
  while (g.vc % g.cpb) --g.cpb;
  g.rpb = 256 / g.cpb;
  if (g.rpb < 1) g.rpb = 1;
  const int col_blocks = g.vc / g.cpb;
  long long rb = (P + (long long)g.rpb * min_iters - 1) / ((long long)g.rpb * min_iters);   // >= min_iters row iterations per block
  long long cap = max_row_blocks / col_blocks;
  if (cap < 1) cap = 1;
  if (rb > cap) rb = cap;
  if (rb < 1) rb = 1;
  g.grid = dim3(col_blocks, (unsigned)rb);
  g.block = dim3(g.cpb * g.rpb);
  return g;
}

// ============================================================================================ weights
// kinds as in sg2b200.h; SG2_STEM4x4 = 4: conv4x4 s2 on 3 channels packed as a K=48(->64) GEMM.
__device__ __forceinline__ int up_lo(int p, int a) { return (p == 0) ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2); }
__device__ __forceinline__ int up_hi(int p, int a) { return (p == 0) ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2); }
__device__ __forceinline__ int s2_kh(int p, int a) { return (p == 0) ? (a == 0 ? 1 : 3) : (a == 0 ? 2 : 0); }

// Offsets of (co, slot, ci) in the fprop pack / dgrad pack of each kind.
__device__ __forceinline__ void pack_offsets(int kind, int co, int slot, int ci, int CoP, int CiP, long long& fo,
                                             long long& to) {
  if (kind == SG2_CONV3x3) {
    fo = ((long long)co * 9 + slot) * CiP + ci;
    to = ((long long)ci * 9 + slot) * CoP + co;
  } else if (kind == SG2_GEMM) {
    fo = (long long)co * CiP + ci;
    to = (long long)ci * CoP + co;
  } else if (kind == SG2_CONV4x4S2) {
    const int kh = slot >> 2, kw = slot & 3;
    fo = ((long long)co * 16 + slot) * CiP + ci;
    // dgrad pack [g=(py,px)][ci][a*2+b][co] with kh = s2_kh(py,a)
    const int py = (kh == 1 || kh == 3) ? 0 : 1, a = (kh == 1 || kh == 2) ? 0 : 1;
    const int px = (kw == 1 || kw == 3) ? 0 : 1, b = (kw == 1 || kw == 2) ? 0 : 1;
    to = ((((long long)(py * 2 + px) * CiP + ci) * 4) + a * 2 + b) * CoP + co;
  } else {  // SG2_UPCONV3x3: slot = (py*2+px)*4 + a*2+b
    const int g = slot >> 2;
    fo = ((((long long)g * CoP + co) * 4) + (slot & 3)) * CiP + ci;
    to = ((long long)ci * 16 + slot) * CoP + co;
  }
}

// fp32 OIHW -> bf16 packs. One block = 16 output channels x 32 input channels x all taps, staged in shared memory:
// the OIHW read is a contiguous run per output channel, the two packs are written in 64-byte / 32-byte runs.
constexpr int kPackCo = 16, kPackCi = 32;
__device__ __forceinline__ void store8(__nv_bfloat16* dst, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(dst) = pack8(v);
}
__device__ __forceinline__ void store8(float* dst, const float (&v)[8]) {   // fp32 packs (precise mode: split into 3 bf16 planes later)
  reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <typename T>
__global__ void pack_weights_kernel(int kind, const float* __restrict__ w, T* __restrict__ wpk,
                                    T* __restrict__ wpkT, int Cout, int Cin, int CoP, int CiP,
                                    int src_ohwi) {
  __shared__ float tile[kPackCo][kPackCi][17];
  const int kk = (kind == SG2_CONV3x3 || kind == SG2_UPCONV3x3) ? 9 : ((kind == SG2_GEMM) ? 1 : 16);   // source taps
  const int slots = (kind == SG2_CONV3x3) ? 9 : ((kind == SG2_GEMM) ? 1 : 16);                         // packed slots
  const int ci_tiles = (CiP + kPackCi - 1) / kPackCi, co_tiles = (CoP + kPackCo - 1) / kPackCo;
  for (int blk = blockIdx.x; blk < ci_tiles * co_tiles; blk += gridDim.x) {
    const int co0 = (blk / ci_tiles) * kPackCo, ci0 = (blk % ci_tiles) * kPackCi;
    if (src_ohwi) {   // master stored [Cout][kh][kw][Cin]: ci is the contiguous axis
      for (int e = threadIdx.x; e < kPackCo * kPackCi * kk; e += blockDim.x) {
        const int cl = e % kPackCi, t = (e / kPackCi) % kk, ol = e / (kk * kPackCi);
        const int co = co0 + ol, ci = ci0 + cl;
        tile[ol][cl][t] = (co < Cout && ci < Cin) ? w[((long long)co * kk + t) * Cin + ci] : 0.f;
      }
    } else {
      for (int e = threadIdx.x; e < kPackCo * kPackCi * kk; e += blockDim.x) {
        const int t = e % kk, cl = (e / kk) % kPackCi, ol = e / (kk * kPackCi);
        const int co = co0 + ol, ci = ci0 + cl;
        tile[ol][cl][t] = (co < Cout && ci < Cin) ? w[((long long)co * Cin + ci) * kk + t] : 0.f;
      }
    }
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
      T* dst = pass == 0 ? wpk : wpkT;
      if (!dst) continue;
      // each thread produces 8 consecutive elements of the destination's contiguous axis -> one 128-bit store
      const int inner8 = (pass == 0 ? kPackCi : kPackCo) / 8;
      const int outer = pass == 0 ? kPackCo : kPackCi;
      for (int e = threadIdx.x; e < outer * slots * inner8; e += blockDim.x) {
        const int i8 = e % inner8, slot = (e / inner8) % slots, o = e / (inner8 * slots);
        const int ol0 = pass == 0 ? o : i8 * 8, cl0 = pass == 0 ? i8 * 8 : o;
        const int co = co0 + ol0, ci = ci0 + cl0;
        if (co >= CoP || ci >= CiP) continue;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int ol = pass == 0 ? ol0 : ol0 + j, cl = pass == 0 ? cl0 + j : cl0;
          if (kind == SG2_UPCONV3x3) {
            const int g = slot >> 2, a = (slot >> 1) & 1, b = slot & 1, py = g >> 1, px = g & 1;
            float acc = 0.f;
            for (int kh = up_lo(py, a); kh <= up_hi(py, a); ++kh)
              for (int kw = up_lo(px, b); kw <= up_hi(px, b); ++kw) acc += tile[ol][cl][kh * 3 + kw];
            v[j] = acc;
          } else {
            v[j] = tile[ol][cl][slot];
          }
        }
        long long fo, to;
        pack_offsets(kind, co, slot, ci, CoP, CiP, fo, to);
        store8(dst + (pass == 0 ? fo : to), v);
      }
    }
    __syncthreads();
  }
}

// dgrad operand from the bf16 fprop operand (CONV3x3 / CONV4x4S2 / GEMM, no channel padding): per tap a [Cout][Cin] ->
// [Cin][Cout] transpose, 64 x 64 tiles through shared memory so that reads run along ci and writes along co.
__global__ void __launch_bounds__(256) pack_transpose_kernel(int kind, const __nv_bfloat16* __restrict__ wpk,
                                                             __nv_bfloat16* __restrict__ wpkT, int Cout, int Cin) {
  __shared__ __nv_bfloat16 tile[64][72];
  const int slots = (kind == SG2_CONV3x3) ? 9 : ((kind == SG2_GEMM) ? 1 : 16);
  const int ci_tiles = (Cin + 63) / 64, co_tiles = (Cout + 63) / 64;
  const long long nblk = (long long)ci_tiles * co_tiles * slots;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int slot = (int)(blk % slots);
    const int ct = (int)((blk / slots) % ci_tiles), ot = (int)(blk / ((long long)slots * ci_tiles));
    const int co0 = ot * 64, ci0 = ct * 64;
    for (int e = threadIdx.x; e < 64 * 8; e += 256) {       // 64 rows (co) x 8 vectors of 8 ci
      const int r = e >> 3, v8 = e & 7;
      const int co = co0 + r, ci = ci0 + v8 * 8;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (co < Cout && ci < Cin) val = *reinterpret_cast<const uint4*>(wpk + ((long long)co * slots + slot) * Cin + ci);
      *reinterpret_cast<uint4*>(&tile[r][v8 * 8]) = val;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * 8; e += 256) {       // 64 rows (ci) x 8 vectors of 8 co
      const int r = e >> 3, v8 = e & 7;
      const int ci = ci0 + r, co = co0 + v8 * 8;
      if (ci < Cin && co < Cout) {
        __nv_bfloat16 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = tile[v8 * 8 + j][r];
        long long fo, to;
        pack_offsets(kind, co, slot, ci, Cout, Cin, fo, to);
        *reinterpret_cast<uint4*>(wpkT + to) = *reinterpret_cast<const uint4*>(v);
      }
    }
    __syncthreads();
  }
}

// 3-channel stem (tiny): GEMM over im2col rows, k = (kh*4+kw)*3 + c, padded to CiP.
__device__ __forceinline__ void cvt_store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void cvt_store(float* p, float v) { *p = v; }
template <typename T>
__global__ void pack_stem_kernel(const float* __restrict__ w, T* __restrict__ wpk,
                                 T* __restrict__ wpkT, int Cout, int CoP, int CiP, int src_ohwi) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CoP * CiP; i += gridDim.x * blockDim.x) {
    const int ci = i % CiP, co = i / CiP;
    float v = 0.f;
    if (co < Cout && ci < 48) {
      const int c = ci % 3, t = ci / 3;
      v = src_ohwi ? w[(long long)co * 48 + ci] : w[((long long)co * 3 + c) * 16 + t];
    }
    if (wpk) cvt_store(&wpk[(long long)co * CiP + ci], v);
    if (wpkT) cvt_store(&wpkT[(long long)ci * CoP + co], v);
  }
}

// dwpk [CoP][jobs][CiP] -> grad OIHW [Cout][Cin][kk].  One block = one output channel x 32 input channels:
// reads run along ci (contiguous in dwpk), writes along (ci, tap) (contiguous in OIHW); the transpose goes
// through shared memory so both sides are coalesced.
__global__ void unpack_wgrad_kernel(int kind, const float* __restrict__ dwpk, float* __restrict__ grad, int Cout,
                                    int Cin, int CoP, int CiP, int accumulate, int dst_ohwi) {
  __shared__ float tile[16][33];
  const int kk = (kind == SG2_CONV3x3 || kind == SG2_UPCONV3x3) ? 9 : ((kind == SG2_GEMM) ? 1 : 16);
  const int jobs = (kind == SG2_CONV3x3) ? 9 : ((kind == SG2_GEMM || kind == SG2_STEM4x4) ? 1 : 16);
  if (kind == SG2_STEM4x4) {  // grad[co][c][t] <- dwpk[co][0][t*3 + c]   (tiny: 64 x 48)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Cout * 48; i += gridDim.x * blockDim.x) {
      const int t = i % 16, c = (i / 16) % 3, co = i / 48;
      const float v = dst_ohwi ? dwpk[(long long)co * CiP + i % 48] : dwpk[(long long)co * CiP + t * 3 + c];
      if (accumulate) grad[i] += v; else grad[i] = v;
    }
    return;
  }
  const int ci_blocks = (Cin + 31) / 32;
  for (int blk = blockIdx.x; blk < Cout * ci_blocks; blk += gridDim.x) {
    const int co = blk / ci_blocks, ci0 = (blk % ci_blocks) * 32;
    const float* row = dwpk + (long long)co * jobs * CiP;
    // load: thread (j = tid / 32, lane) reads job j, channel ci0 + lane
    for (int j = threadIdx.x >> 5; j < jobs; j += blockDim.x >> 5) {
      const int ci = ci0 + (threadIdx.x & 31);
      tile[j][threadIdx.x & 31] = ci < Cin ? row[(long long)j * CiP + ci] : 0.f;
    }
    __syncthreads();
    const int nci = min(32, Cin - ci0);
    for (int e = threadIdx.x; e < nci * kk; e += blockDim.x) {
      const int l = dst_ohwi ? e % nci : e / kk, t = dst_ohwi ? e / nci : e % kk;
      float v = 0.f;
      if (kind == SG2_UPCONV3x3) {
        const int kh = t / 3, kw = t % 3;
        for (int py = 0; py < 2; ++py)
          for (int a = 0; a < 2; ++a) {
            if (kh < up_lo(py, a) || kh > up_hi(py, a)) continue;
            for (int px = 0; px < 2; ++px)
              for (int b = 0; b < 2; ++b) {
                if (kw < up_lo(px, b) || kw > up_hi(px, b)) continue;
                v += tile[(py * 2 + px) * 4 + a * 2 + b][l];
              }
          }
      } else {
        v = tile[t][l];
      }
      const long long o = dst_ohwi ? ((long long)co * kk + t) * Cin + ci0 + l : ((long long)co * Cin + ci0) * kk + e;
      if (accumulate) grad[o] += v; else grad[o] = v;
    }
    __syncthreads();
  }
}

// ---- fused-upsample layers stored [Cout][3x3][Cin] (the fused trainer's buckets): streaming pack / unpack.
// These run on the tail of the generator's backward pass (wgrad -> unpack -> Adam -> pack of the first, largest
// layers), so they are written for bandwidth: every global access is a 16-byte vector in a fully coalesced row.
// up_lo / up_hi give the 3x3 taps that collapse onto 2x2 tap (a, b) of output parity (py, px).
__device__ __forceinline__ void upconv_presum(const float (&t)[9], float (&o)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int py = g >> 1, px = g & 1;
        float acc = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            if (kh >= up_lo(py, a) && kh <= up_hi(py, a) && kw >= up_lo(px, b) && kw <= up_hi(px, b)) acc += t[kh * 3 + kw];
        o[g * 4 + a * 2 + b] = acc;
      }
}
// master [Cout][9][Cin] fp32 -> wpk [4][Cout][4][Cin] bf16 and wpkT [Cin][16][Cout] bf16. Tile = 64 co x 64 ci per block.
__global__ void __launch_bounds__(256) pack_upconv_ohwi_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wpk,
                                                               __nv_bfloat16* __restrict__ wpkT, int Cout, int Cin) {
  __shared__ __nv_bfloat16 tile[4][64][72];   // [slot within group][ci][co (+pad)]
  const int ci_tiles = Cin / 64;
  const int co0 = (blockIdx.x / ci_tiles) * 64, ci0 = (blockIdx.x % ci_tiles) * 64;
  for (int g = 0; g < 4; ++g) {
    for (int e = threadIdx.x; e < 64 * 8; e += 256) {    // (co, 8-channel vector)
      const int v8 = e & 7, col = e >> 3;
      const int co = co0 + col, ci = ci0 + v8 * 8;
      float sl[4][8];
#pragma unroll
      for (int j8 = 0; j8 < 2; ++j8) {
        float4 t4[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) t4[t] = *reinterpret_cast<const float4*>(w + ((long long)co * 9 + t) * Cin + ci + j8 * 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t9[9], o16[16];
#pragma unroll
          for (int t = 0; t < 9; ++t) t9[t] = c == 0 ? t4[t].x : (c == 1 ? t4[t].y : (c == 2 ? t4[t].z : t4[t].w));
          upconv_presum(t9, o16);
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) sl[s4][j8 * 4 + c] = o16[g * 4 + s4];
        }
      }
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const uint4 pk = pack8(sl[s4]);
        if (wpk) *reinterpret_cast<uint4*>(wpk + ((((long long)g * Cout + co) * 4) + s4) * Cin + ci) = pk;
        const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&pk);
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[s4][v8 * 8 + j][col] = pb[j];
      }
    }
    __syncthreads();
    if (wpkT) {
      for (int e = threadIdx.x; e < 4 * 64 * 8; e += 256) {   // (slot, ci, 8-co vector): 128-byte rows of wpkT
        const int v8 = e & 7, cl = (e >> 3) & 63, s4 = e >> 9;
        *reinterpret_cast<uint4*>(wpkT + ((long long)(ci0 + cl) * 16 + g * 4 + s4) * Cout + co0 + v8 * 8) =
            *reinterpret_cast<const uint4*>(&tile[s4][cl][v8 * 8]);
      }
    }
    __syncthreads();
  }
}
// dwpk [Cout][16 jobs][Cin] fp32 -> grad [Cout][9][Cin] fp32 (=|+=): tap (kh, kw) sums the jobs it was folded into
__global__ void __launch_bounds__(256) unpack_upconv_ohwi_kernel(const float4* __restrict__ dwpk, float4* __restrict__ grad,
                                                                 int Cout, int Cin4, int accumulate) {
  const long long total = (long long)Cout * Cin4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long co = i / Cin4;
    const int c4 = (int)(i - co * Cin4);
    float4 j16[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) j16[j] = dwpk[(co * 16 + j) * Cin4 + c4];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            if (kh < up_lo(py, a) || kh > up_hi(py, a)) continue;
#pragma unroll
            for (int px = 0; px < 2; ++px)
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                if (kw < up_lo(px, b) || kw > up_hi(px, b)) continue;
                const float4 t = j16[(py * 2 + px) * 4 + a * 2 + b];
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
              }
          }
        float4* o = grad + (co * 9 + kh * 3 + kw) * Cin4 + c4;
        if (accumulate) { const float4 d = *o; v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w; }
        *o = v;
      }
  }
}

// ============================================================================================ BN statistics
// Per-channel sum / sum of squares of a [P][C] tensor into a ZEROED fp64 workspace stats[2][C]. Each block forms its
// partial sums in a fixed order in fp32; blocks combine with fp64 atomics (order-independent beyond 2^-53).
// IN_F32: the input is `nsplit` fp32 split-K slabs (slab s at xin + s * slab elements) that are summed in slab order,
// optionally combined with an epilogue operand (epi_mode 1: += src, 2: *= LeakyReLU'(src)), rounded to bf16 (written
// to `y`); the statistics (optional) are those of the rounded values (what BatchNorm-apply reads back).
template <bool IN_F32>
__global__ void bn_stats_kernel(const void* __restrict__ xin, uint4* __restrict__ y, long long P, int vc, int cpb,
                                int rpb, double* __restrict__ stats, int C, int nsplit, long long slab,
                                const uint4* __restrict__ epi_src, int epi_mode) {
  // blockIdx.z = statistics group: rows [z * P, (z + 1) * P) of the tensor, sums into stats[z][2][C]
  if (IN_F32) xin = reinterpret_cast<const float*>(xin) + (long long)blockIdx.z * P * C;
  else xin = reinterpret_cast<const uint4*>(xin) + (long long)blockIdx.z * P * vc;
  if (y) y += (long long)blockIdx.z * P * vc;
  if (epi_src) epi_src += (long long)blockIdx.z * P * vc;
  if (stats) stats += (long long)blockIdx.z * 2 * C;
  const int col = blockIdx.x * cpb + threadIdx.x % cpb;
  const int rl = threadIdx.x / cpb;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const long long stride = (long long)gridDim.y * rpb;
  for (long long r = (long long)blockIdx.y * rpb + rl; r < P; r += stride) {
    float f[8];
    if (IN_F32) {
      const float4* x4 = reinterpret_cast<const float4*>(xin) + (r * vc + col) * 2;
      float4 lo = x4[0], hi = x4[1];
      // slabs are added in slab order (fixed summation order); four slabs' loads are in flight at a time
      for (int sp = 1; sp < nsplit; sp += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int s2 = sp + u < nsplit ? sp + u : 0;
          const float4* xs = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xin) + (long long)s2 * slab) +
                             (r * vc + col) * 2;
          a[u] = xs[0];
          b[u] = xs[1];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (sp + u < nsplit) {
            lo.x += a[u].x; lo.y += a[u].y; lo.z += a[u].z; lo.w += a[u].w;
            hi.x += b[u].x; hi.y += b[u].y; hi.z += b[u].z; hi.w += b[u].w;
          }
        }
      }
      float t[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      if (epi_mode) {
        float e[8];
        unpack8(epi_src[r * vc + col], e);
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = epi_mode == 1 ? t[j] + e[j] : (e[j] > 0.f ? t[j] : 0.2f * t[j]);
      }
      const uint4 pk = pack8(t);
      y[r * vc + col] = pk;
      unpack8(pk, f);
    } else {
      unpack8(reinterpret_cast<const uint4*>(xin)[r * vc + col], f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] += f[j] * f[j]; }
  }
  if (!stats) return;
  __shared__ float sh[2][256][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sh[0][threadIdx.x][j] = s[j]; sh[1][threadIdx.x][j] = q[j]; }
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < rpb; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += sh[0][threadIdx.x + k * cpb][j]; q[j] += sh[1][threadIdx.x + k * cpb][j]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&stats[col * 8 + j], (double)s[j]);
      atomicAdd(&stats[C + col * 8 + j], (double)q[j]);
    }
  }
}

// Many slabs (the tile wgrad kernel writes one per CTA lane, up to 148): block = 32 float4 columns x 8 slab lanes; lane l
// adds slabs l, l+8, ... in order, the 8 lane sums are combined in lane order through shared memory — a fixed summation
// tree (reproducible) with 8 x the loads in flight of the serial loop below.
__global__ void __launch_bounds__(256) reduce_slabs_wide_kernel(const float4* __restrict__ parts, int nslabs, long long n4,
                                                                long long slab4, float4* __restrict__ dst, int accumulate) {
  __shared__ float4 sh[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long c0 = (long long)blockIdx.x * 32; c0 < n4; c0 += (long long)gridDim.x * 32) {
    const long long col = c0 + tx;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < n4) {
      for (int sp = ty; sp < nslabs; sp += 32) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = parts[(long long)(sp + 8 * u < nslabs ? sp + 8 * u : ty) * slab4 + col];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (sp + 8 * u < nslabs) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
      }
    }
    sh[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && col < n4) {
#pragma unroll
      for (int l = 1; l < 8; ++l) { const float4 b = sh[l][tx]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
      if (accumulate) { const float4 d = dst[col]; a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w; }
      dst[col] = a;
    }
    __syncthreads();
  }
}

// dst[i] (=|+=) sum over the slabs, in slab order (deterministic counterpart of the red.global.add split reductions)
__global__ void reduce_slabs_kernel(const float4* __restrict__ parts, int nslabs, long long n4, long long slab4,
                                    float4* __restrict__ dst, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 a = parts[i];
    for (int sp = 1; sp < nslabs; ++sp) {
      const float4 b = parts[(long long)sp * slab4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (accumulate) {
      const float4 d = dst[i];
      a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
    }
    dst[i] = a;
  }
}

// batch mean / rstd of channel c from the accumulated sums (biased variance, like nn.BatchNorm in train mode)
__device__ __forceinline__ void stat_mean_rstd(const double* __restrict__ stats, int C, int c, double invP, float eps,
                                               float& m, float& r, float& var_out) {
  const double mm = stats[c] * invP;
  double var = stats[C + c] * invP - mm * mm;
  if (var < 0) var = 0;
  m = (float)mm;
  r = (float)(1.0 / sqrt(var + (double)eps));
  var_out = (float)var;
}

__global__ void bn_eval_prepare_kernel(const float* rmean, const float* rvar, float eps, float* mean, float* rstd,
                                       int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { mean[c] = rmean[c]; rstd[c] = rsqrtf(rvar[c] + eps); }
}

// ============================================================================================ BN apply (+act)
enum { ACT_NONE = 0, ACT_GLU = 1, ACT_LRELU = 2 };

// Row streaming for the [P][C] tensors. A thread owns one 4-channel column (8-byte vectors) of a few rows, so the
// per-channel constants of the fused BN + activation math cost 16-40 registers and three 256-thread blocks fit an SM.
// Large tensors take the STAGED path: row tiles of ~20 KB are pulled through a shared-memory ring with cp.async.bulk
// (the TMA engine; one elected thread issues, mbarrier completion), so ~180 KB per SM are in flight regardless of
// register use, and the math reads its operands from shared memory. body(xs, ds, ri, rg): xs/ds = row-major tiles,
// ri = row inside them, rg = row of the whole tensor (for the outputs).
constexpr int kBnW = 4;                       // channels per thread
constexpr float kNegLog2e = -1.4426950408889634f;
__device__ __forceinline__ void unpack4(const uint2& v, float (&f)[4]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
}
__device__ __forceinline__ uint2 pack4(const float (&f)[4]) {
  return make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
}
// 1 / (1 + 2^t): the sigmoid of x when t = -x * log2(e) (the scale is folded into the per-channel constants)
// t is clamped so that e = 2^t stays finite: the backward pass forms 1 - s as e * s, and a gate pre-activation far
// below zero (a few outlier pixels of a sharply peaked channel reach |z| ~ sqrt(pixels) after BatchNorm) gave
// e = inf, s = 0, e * s = NaN — one NaN channel, then every generator gradient (found at step 189 of a run on
// three rotating batches; tests/test_gpu_kernels.py::test_glu_backward_extreme_gate). 2^-100 is zero at bf16 output scale.
__device__ __forceinline__ float sigmoid_from_t(float t, float& e) {
  e = exp2f(fminf(t, 100.f));
  return __fdividef(1.f, 1.f + e);
}

constexpr int kStMaxStages = 4;
template <bool WITH_D, typename F>
__device__ __forceinline__ void staged_rows(const uint2* __restrict__ x, const uint2* __restrict__ d, long long P,
                                            int vc_in, int vc_out, int R, int stages, int rl, int rpb, F&& body) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  __shared__ uint64_t st_full[kStMaxStages];
  const int xv = R * vc_in, dv = WITH_D ? R * vc_out : 0;   // uint2 vectors per stage
  const long long tiles = (P + R - 1) / R;
  const int lane_id = blockIdx.y, nl = gridDim.y;
  const long long my_tiles = lane_id < tiles ? (tiles - lane_id + nl - 1) / nl : 0;
  uint2* ring = reinterpret_cast<uint2*>(st_smem);
  auto issue = [&](long long i) {
    const long long t = lane_id + i * nl;
    const int s = (int)(i % stages);
    const long long r0 = t * R;
    const int rows = (int)((P - r0) < R ? (P - r0) : R);
    uint2* sx = ring + (size_t)s * (xv + dv);
    mbar_expect_tx(&st_full[s], (uint32_t)rows * (uint32_t)(vc_in + (WITH_D ? vc_out : 0)) * 8u);
    bulk_g2s(sx, x + r0 * vc_in, (uint32_t)rows * vc_in * 8u, &st_full[s]);
    if (WITH_D) bulk_g2s(sx + xv, d + r0 * vc_out, (uint32_t)rows * vc_out * 8u, &st_full[s]);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&st_full[s], 1);
    fence_barrier_init();
    for (long long i = 0; i < my_tiles && i < stages; ++i) issue(i);
  }
  __syncthreads();
  int s = 0;
  uint32_t phase = 0;
  for (long long i = 0; i < my_tiles; ++i) {
    mbar_wait(&st_full[s], phase);
    const long long r0 = (lane_id + i * nl) * (long long)R;
    const int rows = (int)((P - r0) < R ? (P - r0) : R);
    const uint2* sx = ring + (size_t)s * (xv + dv);
    for (int r = rl; r < rows; r += rpb) body(sx, sx + xv, (long long)r, r0 + r);
    __syncthreads();
    if (threadIdx.x == 0 && i + stages < my_tiles) issue(i + stages);
    if (++s == stages) { s = 0; phase ^= 1u; }
  }
}

// out = act(bn(x)) (+ residual).  GLU: out[:, c] = bn(x)[:, c] * sigmoid(bn(x)[:, c + C/2]), out has C/2 channels.
template <int ACT, bool STAGED>
__global__ void __launch_bounds__(256, 3)
bn_act_fwd_kernel(const uint2* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const uint2* __restrict__ residual,
                  uint2* __restrict__ out, long long P, int vc_in, int vc_out, int cpb, int rpb, int has_bn,
                  const double* __restrict__ stats, float eps, float momentum, float* __restrict__ mean_out,
                  float* __restrict__ rstd_out, float* __restrict__ rmean, float* __restrict__ rvar,
                  long long* __restrict__ nbt, int tile_rows, int stages) {
  // stats != NULL: train mode — mean/rstd are derived here from the sums the conv epilogue accumulated; the first
  // row-block also saves them for backward and updates the running statistics (momentum, unbiased variance).
  const int col = blockIdx.x * cpb + threadIdx.x % cpb;  // output vector column
  const int rl = threadIdx.x / cpb;
  const int C = vc_in * kBnW;
  // blockIdx.z = statistics group (rows [z * P, (z + 1) * P)): each group of the batch is normalised on its own
  // statistics, exactly like separate nn.BatchNorm calls on the group's samples (train_Dnet's real / wrong / fake passes).
  const int gz = blockIdx.z, ngroups = gridDim.z;
  x += (long long)gz * P * vc_in;
  out += (long long)gz * P * vc_out;
  if (residual) residual += (long long)gz * P * vc_out;
  const double* stats_all = stats;
  if (stats) stats += (long long)gz * 2 * C;
  if (mean) { mean += (long long)gz * C; rstd += (long long)gz * C; }
  if (mean_out) { mean_out += (long long)gz * C; rstd_out += (long long)gz * C; }
  const bool writer = stats && blockIdx.y == 0 && rl == 0;
  const double invP = 1.0 / (double)P;
  if (writer && nbt && gz == 0 && blockIdx.x == 0 && threadIdx.x == 0) *nbt += ngroups;
  // running statistics: updated once per group, in group order, by group 0's writer threads
  auto update_running = [&](int c) {
    if (!rmean || gz != 0) return;
    float rm = rmean[c], rv = rvar[c];
    for (int g = 0; g < ngroups; ++g) {
      float m, r, var;
      stat_mean_rstd(stats_all + (long long)g * 2 * C, C, c, invP, eps, m, r, var);
      const float unb = P > 1 ? var * (float)((double)P / (double)(P - 1)) : var;
      rm = (1.f - momentum) * rm + momentum * m;
      rv = (1.f - momentum) * rv + momentum * unb;
    }
    rmean[c] = rm;
    rvar[c] = rv;
  };
  // per-channel scale / shift of channel c; `t_scale` folds the sigmoid's -log2(e) into the gate half
  auto scale_shift = [&](int c, float t_scale, float& sc, float& sh) {
    if (!has_bn) { sc = t_scale; sh = 0.f; return; }
    float m, r, var = 0.f;
    if (stats) stat_mean_rstd(stats, C, c, invP, eps, m, r, var); else { m = mean[c]; r = rstd[c]; }
    if (writer) {
      mean_out[c] = m; rstd_out[c] = r;
      update_running(c);
    }
    const float s = gamma[c] * r;
    sc = s * t_scale;
    sh = (beta[c] - m * s) * t_scale;
  };
  float sc0[kBnW], sh0[kBnW], sc1[kBnW], sh1[kBnW];
#pragma unroll
  for (int j = 0; j < kBnW; ++j) {
    float a, b;
    scale_shift(col * kBnW + j, 1.f, a, b);
    sc0[j] = a; sh0[j] = b;
    if (ACT == ACT_GLU) {
      scale_shift(col * kBnW + j + vc_out * kBnW, kNegLog2e, a, b);
      sc1[j] = a; sh1[j] = b;
    }
  }
  auto body = [&](const uint2* xs, const uint2* /*ds*/, long long ri, long long r) {
    float a[kBnW], o[kBnW];
    unpack4(xs[ri * vc_in + col], a);
    if (ACT == ACT_GLU) {
      float g[kBnW];
      unpack4(xs[ri * vc_in + col + vc_out], g);
#pragma unroll
      for (int j = 0; j < kBnW; ++j) {
        float e;
        o[j] = fmaf(a[j], sc0[j], sh0[j]) * sigmoid_from_t(fmaf(g[j], sc1[j], sh1[j]), e);
      }
    } else if (ACT == ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < kBnW; ++j) { const float z = fmaf(a[j], sc0[j], sh0[j]); o[j] = z > 0.f ? z : 0.2f * z; }
    } else {
#pragma unroll
      for (int j = 0; j < kBnW; ++j) o[j] = fmaf(a[j], sc0[j], sh0[j]);
      if (residual) {
        float rr[kBnW];
        unpack4(residual[r * vc_out + col], rr);
#pragma unroll
        for (int j = 0; j < kBnW; ++j) o[j] += rr[j];
      }
    }
    out[r * vc_out + col] = pack4(o);
  };
  if (STAGED) {
    staged_rows<false>(x, nullptr, P, vc_in, vc_out, tile_rows, stages, rl, rpb, body);
  } else {
    for (long long r = (long long)blockIdx.y * rpb + rl; r < P; r += (long long)gridDim.y * rpb) body(x, nullptr, r, r);
  }
}

// dz = d(loss)/d(bn output) of one vector column (GLU: dz0 of the value half, dz1 of the gate half).
// sc0/sh0 = BN scale / shift of the value half; sc1/sh1 = those of the gate half times -log2(e).
template <int ACT>
__device__ __forceinline__ void bn_act_dz(const float (&a)[kBnW], const float (&g)[kBnW], const float (&d)[kBnW],
                                          const float (&sc0)[kBnW], const float (&sh0)[kBnW],
                                          const float (&sc1)[kBnW], const float (&sh1)[kBnW], float (&dz0)[kBnW],
                                          float (&dz1)[kBnW]) {
#pragma unroll
  for (int j = 0; j < kBnW; ++j) {
    if (ACT == ACT_GLU) {
      const float za = fmaf(a[j], sc0[j], sh0[j]);
      float e;
      const float s = sigmoid_from_t(fmaf(g[j], sc1[j], sh1[j]), e);
      dz0[j] = d[j] * s;
      dz1[j] = dz0[j] * za * (e * s);          // 1 - s = e * s
    } else if (ACT == ACT_LRELU) {
      const float z = fmaf(a[j], sc0[j], sh0[j]);
      dz0[j] = z > 0.f ? d[j] : 0.2f * d[j];
    } else {
      dz0[j] = d[j];
    }
  }
}

// pass 1: sums[0][c] = sum dz, sums[1][c] = sum dz * xhat   (per input channel c, fp64 atomics).
// The loop accumulates S = sum dz and T = sum dz * x; sum dz * xhat = rstd * (T - mean * S) is formed in fp64 per block.
template <int ACT, bool STAGED>
__global__ void __launch_bounds__(256, 3)
bn_act_bwd_reduce_kernel(const uint2* __restrict__ x, const uint2* __restrict__ dout, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                         const float* __restrict__ beta, long long P, int vc_in, int vc_out, int cpb, int rpb,
                         double* __restrict__ sums, int C, int tile_rows, int stages) {
  x += (long long)blockIdx.z * P * vc_in;       // blockIdx.z = statistics group (see bn_act_fwd_kernel)
  dout += (long long)blockIdx.z * P * vc_out;
  mean += (long long)blockIdx.z * C;
  rstd += (long long)blockIdx.z * C;
  sums += (long long)blockIdx.z * 2 * C;
  const int col = blockIdx.x * cpb + threadIdx.x % cpb;
  const int rl = threadIdx.x / cpb;
  float sc0[kBnW], sh0[kBnW], sc1[kBnW], sh1[kBnW];
#pragma unroll
  for (int j = 0; j < kBnW; ++j) {
    const int c = col * kBnW + j;
    sc0[j] = gamma[c] * rstd[c]; sh0[j] = beta[c] - mean[c] * sc0[j];
    if (ACT == ACT_GLU) {
      const int c2 = c + vc_out * kBnW;
      const float s = gamma[c2] * rstd[c2];
      sc1[j] = s * kNegLog2e; sh1[j] = (beta[c2] - mean[c2] * s) * kNegLog2e;
    } else { sc1[j] = sh1[j] = 0.f; }
  }
  float s0[kBnW], t0[kBnW], s1[kBnW], t1[kBnW];
#pragma unroll
  for (int j = 0; j < kBnW; ++j) s0[j] = t0[j] = s1[j] = t1[j] = 0.f;
  auto body = [&](const uint2* xs, const uint2* ds, long long ri, long long /*r*/) {
    float a[kBnW], g[kBnW], d[kBnW], dz0[kBnW], dz1[kBnW];
    unpack4(xs[ri * vc_in + col], a);
    if (ACT == ACT_GLU) unpack4(xs[ri * vc_in + col + vc_out], g);
    unpack4(ds[ri * vc_out + col], d);
    bn_act_dz<ACT>(a, g, d, sc0, sh0, sc1, sh1, dz0, dz1);
#pragma unroll
    for (int j = 0; j < kBnW; ++j) {
      s0[j] += dz0[j]; t0[j] = fmaf(dz0[j], a[j], t0[j]);
      if (ACT == ACT_GLU) { s1[j] += dz1[j]; t1[j] = fmaf(dz1[j], g[j], t1[j]); }
    }
  };
  if (STAGED) {
    staged_rows<true>(x, dout, P, vc_in, vc_out, tile_rows, stages, rl, rpb, body);
  } else {
    for (long long r = (long long)blockIdx.y * rpb + rl; r < P; r += (long long)gridDim.y * rpb) body(x, dout, r, r);
  }
  // block reduction over the rpb row slots (the staged ring is idle by now: reuse it as the scratch)
  extern __shared__ __align__(128) uint8_t st_smem[];
  __shared__ float sh_static[STAGED ? 1 : 4 * 256 * kBnW];
  float (*sh)[256][kBnW] = reinterpret_cast<float (*)[256][kBnW]>(STAGED ? reinterpret_cast<float*>(st_smem) : sh_static);
#pragma unroll
  for (int j = 0; j < kBnW; ++j) {
    sh[0][threadIdx.x][j] = s0[j]; sh[1][threadIdx.x][j] = t0[j];
    if (ACT == ACT_GLU) { sh[2][threadIdx.x][j] = s1[j]; sh[3][threadIdx.x][j] = t1[j]; }
  }
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < rpb; ++k)
#pragma unroll
      for (int j = 0; j < kBnW; ++j) {
        s0[j] += sh[0][threadIdx.x + k * cpb][j]; t0[j] += sh[1][threadIdx.x + k * cpb][j];
        if (ACT == ACT_GLU) { s1[j] += sh[2][threadIdx.x + k * cpb][j]; t1[j] += sh[3][threadIdx.x + k * cpb][j]; }
      }
#pragma unroll
    for (int j = 0; j < kBnW; ++j) {
      const int c = col * kBnW + j;
      atomicAdd(&sums[c], (double)s0[j]);
      atomicAdd(&sums[C + c], (double)rstd[c] * ((double)t0[j] - (double)mean[c] * (double)s0[j]));
      if (ACT == ACT_GLU) {
        const int c2 = c + vc_out * kBnW;
        atomicAdd(&sums[c2], (double)s1[j]);
        atomicAdd(&sums[C + c2], (double)rstd[c2] * ((double)t1[j] - (double)mean[c2] * (double)s1[j]));
      }
    }
  }
}

// pass 2: dx = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat)) = sc * dz + nA * x + nB  with per-channel
// nA = -sc * rstd * mean(dz * xhat), nB = -sc * mean(dz) - nA * mean.
template <int ACT, bool STAGED>
__global__ void __launch_bounds__(256, 3)
bn_act_bwd_apply_kernel(const uint2* __restrict__ x, const uint2* __restrict__ dout, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const double* __restrict__ sums, long long P, int vc_in,
                        int vc_out, int cpb, int rpb, uint2* __restrict__ dx, int C, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, int accumulate, int tile_rows, int stages) {
  const int gz = blockIdx.z, ngroups = gridDim.z;   // statistics group (see bn_act_fwd_kernel)
  const double* sums_all = sums;
  x += (long long)gz * P * vc_in;
  dout += (long long)gz * P * vc_out;
  dx += (long long)gz * P * vc_in;
  mean += (long long)gz * C;
  rstd += (long long)gz * C;
  sums += (long long)gz * 2 * C;
  const int col = blockIdx.x * cpb + threadIdx.x % cpb;
  const int rl = threadIdx.x / cpb;
  const float invP = 1.f / (float)P;
  if (dgamma && gz == 0 && blockIdx.y == 0 && rl == 0) {   // dgamma = sum dz * xhat, dbeta = sum dz (over all groups)
    auto total = [&](int idx) {
      double t = 0.0;
#pragma unroll 1
      for (int g = 0; g < ngroups; ++g) t += sums_all[(long long)g * 2 * C + idx];
      return (float)t;
    };
#pragma unroll 1
    for (int j = 0; j < (ACT == ACT_GLU ? 2 * kBnW : kBnW); ++j) {
      const int c = col * kBnW + (j % kBnW) + (j >= kBnW ? vc_out * kBnW : 0);
      const float db = total(c), dg = total(C + c);
      if (accumulate) { dgamma[c] += dg; dbeta[c] += db; } else { dgamma[c] = dg; dbeta[c] = db; }
    }
  }
  float sc0[kBnW], sh0[kBnW], nA0[kBnW], nB0[kBnW], sc1[kBnW], sh1[kBnW], g1[kBnW], nA1[kBnW], nB1[kBnW];
#pragma unroll
  for (int j = 0; j < kBnW; ++j) {
    const int c = col * kBnW + j;
    const float m = mean[c], r = rstd[c];
    sc0[j] = gamma[c] * r; sh0[j] = beta[c] - m * sc0[j];
    nA0[j] = -sc0[j] * r * ((float)sums[C + c] * invP);
    nB0[j] = -sc0[j] * ((float)sums[c] * invP) - nA0[j] * m;
    if (ACT == ACT_GLU) {
      const int c2 = c + vc_out * kBnW;
      const float m2 = mean[c2], r2 = rstd[c2];
      g1[j] = gamma[c2] * r2;
      sc1[j] = g1[j] * kNegLog2e; sh1[j] = (beta[c2] - m2 * g1[j]) * kNegLog2e;
      nA1[j] = -g1[j] * r2 * ((float)sums[C + c2] * invP);
      nB1[j] = -g1[j] * ((float)sums[c2] * invP) - nA1[j] * m2;
    } else { sc1[j] = sh1[j] = g1[j] = nA1[j] = nB1[j] = 0.f; }
  }
  auto body = [&](const uint2* xs, const uint2* ds, long long ri, long long r) {
    float a[kBnW], g[kBnW], d[kBnW], dz0[kBnW], dz1[kBnW], o[kBnW];
    unpack4(xs[ri * vc_in + col], a);
    if (ACT == ACT_GLU) unpack4(xs[ri * vc_in + col + vc_out], g);
    unpack4(ds[ri * vc_out + col], d);
    bn_act_dz<ACT>(a, g, d, sc0, sh0, sc1, sh1, dz0, dz1);
#pragma unroll
    for (int j = 0; j < kBnW; ++j) o[j] = fmaf(sc0[j], dz0[j], fmaf(nA0[j], a[j], nB0[j]));
    dx[r * vc_in + col] = pack4(o);
    if (ACT == ACT_GLU) {
#pragma unroll
      for (int j = 0; j < kBnW; ++j) o[j] = fmaf(g1[j], dz1[j], fmaf(nA1[j], g[j], nB1[j]));
      dx[r * vc_in + col + vc_out] = pack4(o);
    }
  };
  if (STAGED) {
    staged_rows<true>(x, dout, P, vc_in, vc_out, tile_rows, stages, rl, rpb, body);
  } else {
    for (long long r = (long long)blockIdx.y * rpb + rl; r < P; r += (long long)gridDim.y * rpb) body(x, dout, r, r);
  }
}

// LeakyReLU without BN (D stem): backward is dx = dout * (x > 0 ? 1 : 0.2)
__global__ void lrelu_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dout, uint4* __restrict__ dx,
                                 long long nvec) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float a[8], d[8];
    unpack8(x[i], a);
    unpack8(dout[i], d);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = a[j] > 0.f ? d[j] : 0.2f * d[j];
    dx[i] = pack8(d);
  }
}

__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                                long long nvec) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(a[i], x);
    unpack8(b[i], y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    out[i] = pack8(x);
  }
}

__global__ void bf16_to_f32_kernel(const uint2* __restrict__ in, float4* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const uint2 v = in[i];
    out[i] = make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
  }
}
__global__ void f32_to_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    out[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ============================================================================================ c_code concat
// out[b, y, x, :E] = c[b, :] ; out[b, y, x, E:] = h[b, y, x, :]     (model.py:274-277, 431-434)
__global__ void concat_c_kernel(const float* __restrict__ c, const uint4* __restrict__ h, uint4* __restrict__ out,
                                int B, int HW, int E, int Ch) {
  const int ve = E / 8, vh = Ch / 8, vo = ve + vh;
  const long long total = (long long)B * HW * vo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vo);
    const long long pix = i / vo;
    if (v < ve) {
      const int b = (int)(pix / HW);
      const float* cp = c + (long long)b * E + v * 8;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = cp[j];
      out[i] = pack8(f);
    } else {
      out[i] = h[pix * vh + (v - ve)];
    }
  }
}
// dh = dcat[..., E:] ; dc[b, e] += sum_{pixels} dcat[b, pixel, e]
// grid (chunks, B). Each block sweeps a chunk of one sample's pixels: thread (pl, v) owns vector column v of the
// c part (ve = E/8 columns) for pixel lanes pl; block-reduce over pixel lanes in smem, one atomic per channel.
__global__ void concat_c_bwd_kernel(const uint4* __restrict__ dcat, uint4* __restrict__ dh, float* __restrict__ dc,
                                    int B, int HW, int E, int Ch) {
  const int ve = E / 8, vh = Ch / 8, vo = ve + vh;
  const int b = blockIdx.y;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  const long long base = (long long)b * HW;
  // ---- dc reduction
  const int lanes = blockDim.x / ve;           // pixel lanes (blockDim.x is a multiple of ve)
  const int v = threadIdx.x % ve, pl = threadIdx.x / ve;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (pl < lanes) {
    for (int p = p0 + pl; p < p1; p += lanes) {
      float f[8];
      unpack8(dcat[(base + p) * vo + v], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
  __shared__ float sh[256][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (pl == 0) {
    for (int k = 1; k < lanes; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += sh[threadIdx.x + k * ve][j];
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&dc[(long long)b * E + v * 8 + j], acc[j]);
  }
  // ---- dh copy
  if (dh) {
    const long long n = (long long)(p1 - p0) * vh;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const long long p = p0 + i / vh;
      const int c = (int)(i % vh);
      dh[(base + p) * vh + c] = dcat[(base + p) * vo + ve + c];
    }
  }
}

// ============================================================================================ image heads / stems
// y: [P][CP] bf16 conv output (first 3 channels valid) -> img fp32 NCHW = tanh(y)      (model.py:291-294)
__global__ void head_tanh_fwd_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ img, int B, int HW,
                                     int CP) {
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int p = (int)(i % HW);
    const uint2 v = *reinterpret_cast<const uint2*>(y + i * CP);
    img[((long long)b * 3 + 0) * HW + p] = tanhf(bf16_lo(v.x));
    img[((long long)b * 3 + 1) * HW + p] = tanhf(bf16_hi(v.x));
    img[((long long)b * 3 + 2) * HW + p] = tanhf(bf16_lo(v.y));
  }
}
// dy[P][CP] bf16 = dimg * (1 - img^2) in channels 0..2, zero elsewhere
__global__ void head_tanh_bwd_kernel(const float* __restrict__ dimg, const float* __restrict__ img,
                                     __nv_bfloat16* __restrict__ dy, int B, int HW, int CP) {
  const int vpp = CP / 8;
  const long long total = (long long)B * HW * vpp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpp);
    const long long pix = i / vpp;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (v == 0) {
      const int b = (int)(pix / HW);
      const int p = (int)(pix % HW);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const long long o = ((long long)b * 3 + c) * HW + p;
        const float t = img[o];
        f[c] = dimg[o] * (1.f - t * t);
      }
    }
    reinterpret_cast<uint4*>(dy)[i] = pack8(f);
  }
}

// D stem: conv4x4 s2 p1 on a 3-channel fp32 NCHW image == GEMM over im2col rows [P][64] (k = (kh*4+kw)*3 + c, 48 used)
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ img, uint4* __restrict__ col, int B,
                                                          int S) {
  // one block = one output row (b, oy): the 4 input rows x 3 channels it reads are staged in shared memory with
  // coalesced loads (+1 zero column each side = the conv padding). The col row is assembled in shared memory as bf16
  // pairs — item = (pair q of k, ox) with q uniform over a warp, so the k -> (c, kh, kw) decode is one broadcast table
  // read instead of eight run-time divisions per vector — and leaves with coalesced 16-byte stores.
  extern __shared__ float rows[];   // [3][4][S + 2] floats | [So][33] words | [48] offsets
  const int So = S / 2, SP = S + 2;
  uint32_t* outw = reinterpret_cast<uint32_t*>(rows + 12 * SP);
  int* off = reinterpret_cast<int*>(outw + So * 33);
  if (threadIdx.x < 48) {
    const int k = threadIdx.x, c = k % 3, t = k / 3;
    off[k] = (c * 4 + (t >> 2)) * SP + (t & 3);
  }
  for (int blk = blockIdx.x; blk < B * So; blk += gridDim.x) {
    const int b = blk / So, oy = blk % So;
    __syncthreads();
    for (int e = threadIdx.x; e < 12 * SP; e += blockDim.x) {
      const int cr = e / SP, xx = e - cr * SP, rr = cr & 3, c = cr >> 2;
      const int y = 2 * oy + rr - 1, x = xx - 1;
      rows[e] = (y >= 0 && y < S && x >= 0 && x < S) ? __ldg(img + (((long long)b * 3 + c) * S + y) * S + x) : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 24 * So; e += blockDim.x) {
      const int q = e / So, ox = e - q * So;
      const float v0 = rows[off[2 * q] + 2 * ox], v1 = rows[off[2 * q + 1] + 2 * ox];
      outw[ox * 33 + q] = pack_bf16x2(v0, v1);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < So * 8; e += blockDim.x) {
      const int v = e & 7, ox = e >> 3;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (v < 6) {
        const uint32_t* w = outw + ox * 33 + 4 * v;
        o = make_uint4(w[0], w[1], w[2], w[3]);
      }
      col[((long long)blk * So + ox) * 8 + v] = o;
    }
  }
}

// dimg fp32 NCHW (=, not +=) from dcol [P][64] bf16 (gather form: no atomics).
// Image rows y = 2t+1 and 2t+2 read exactly the dcol rows oy = t and t+1 (kh = y + 1 - 2 oy in 0..3), so one block per
// (b, t), t = -1 .. So-1, stages those two dcol rows (48 used columns of each im2col row, 16-byte coalesced loads, rows
// padded to 33 words: conflict-free stores) and writes the two image rows of the three channels with coalesced stores.
// The first form gathered four scattered 2-byte values per pixel straight from global memory: 122 us for D256's 50 MB.
__global__ void __launch_bounds__(256) stem_col2im_kernel(const __nv_bfloat16* __restrict__ dcol, float* __restrict__ dimg,
                                                          int B, int S) {
  extern __shared__ uint32_t crow[];   // [2][So][33] words (bf16 pairs)
  const int So = S / 2;
  const int nt = So + 1;
  for (int blk = blockIdx.x; blk < B * nt; blk += gridDim.x) {
    const int b = blk / nt, t = blk % nt - 1;
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * So * 6; e += blockDim.x) {   // 6 x 16 bytes = the 48 used columns
      const int j = e % 6, r = e / 6, oi = r / So, ox = r - oi * So, oy = t + oi;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (oy >= 0 && oy < So) v = __ldg(reinterpret_cast<const uint4*>(dcol + (((long long)b * So + oy) * So + ox) * 64) + j);
      uint32_t* d = crow + r * 33 + 4 * j;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * 3 * S; e += blockDim.x) {
      const int x = e % S, c = (e / S) % 3, yi = e / (3 * S);
      const int y = 2 * t + 1 + yi;
      if (y < 0 || y >= S) continue;
      float acc = 0.f;
#pragma unroll
      for (int oi = 1; oi >= 0; --oi) {               // kh ascending, like the reference's accumulation order
        const int kh = y + 1 - 2 * (t + oi);          // oi = 1: yi (0 | 1);  oi = 0: 2 + yi (2 | 3)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int kw = ((x + 1) & 1) + 2 * kk;      // x + 1 - kw even
          const int ox = (x + 1 - kw) >> 1;
          if (ox < 0 || ox >= So) continue;
          const int k = (kh * 4 + kw) * 3 + c;
          const uint32_t w = crow[(oi * So + ox) * 33 + (k >> 1)];
          acc += (k & 1) ? bf16_hi(w) : bf16_lo(w);
        }
      }
      dimg[(((long long)b * 3 + c) * S + y) * S + x] = acc;
    }
  }
}

// NHWC bf16 [B][HW][C] <-> NCHW fp32 [B][C][HW]   (x_immediate = x_code.reshape(B,-1), model.py:427-428)
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int B,
                                             int HW, int C) {
  const long long total = (long long)B * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int p = (int)((i / C) % HW);
    const int b = (int)(i / ((long long)C * HW));
    out[((long long)b * C + c) * HW + p] = __bfloat162float(in[i]);
  }
}
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B,
                                             int HW, int C) {
  const long long total = (long long)B * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int p = (int)((i / C) % HW);
    const int b = (int)(i / ((long long)C * HW));
    out[i] = __float2bfloat16_rn(in[((long long)b * C + c) * HW + p]);
  }
}

static inline unsigned grid1d(long long n, int threads = 256) {
  long long g = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sg2

using namespace sg2;

// Staged (cp.async.bulk) geometry for a [P][C] tensor: one block spans all columns; row tiles of ~tile_kb KB
// (x rows + dout rows), `stages` of them in flight per block, `blocks` blocks per SM.
struct StagedGeo {
  bool ok;
  int tile_rows, stages;
  unsigned lanes;
  size_t smem;
};
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static StagedGeo make_staged(const Geo& g, long long P, int groups, int vc_in, int vc_out, bool with_d) {
  StagedGeo sg{false, 0, 0, 0, 0};
  static const int on = env_int("SG2_BN_STAGED", 1);
  static const int env_tile_kb = env_int("SG2_BN_TILE_KB", 0);
  static const int env_stages = env_int("SG2_BN_STAGES", 0);
  static const int env_blocks = env_int("SG2_BN_BLOCKS", 0);
  const long long row_bytes = (long long)(vc_in + (with_d ? vc_out : 0)) * 8;
  const long long x_bytes = P * groups * vc_in * 8;
  if (!on || g.grid.x != 1 || x_bytes < (4LL << 20)) return sg;   // small tensors are latency bound anyway
  // measured (tools/bench_bn.py): two blocks per SM with two 44 KB stages for the big tensors; the backward pair on
  // tensors under 32 MB does better with one block per SM (half as many fp64 atomics at the end of the reduce pass)
  const int blocks = env_blocks ? env_blocks : ((with_d && x_bytes < (32LL << 20)) ? 1 : 2);
  const int tile_kb = env_tile_kb ? env_tile_kb : (blocks >= 2 ? 44 : 60);
  const int stages = env_stages ? env_stages : (blocks >= 2 ? 2 : 3);
  int k = (int)((long long)tile_kb * 1024 / (g.rpb * row_bytes));
  if (k < 1) k = 1;
  sg.tile_rows = g.rpb * k;
  sg.stages = stages < 2 ? 2 : (stages > kStMaxStages ? kStMaxStages : stages);
  const long long tiles = (P + sg.tile_rows - 1) / sg.tile_rows;
  long long want = 148LL * blocks / groups;       // blockIdx.z = groups multiplies the grid
  if (want < 1) want = 1;
  sg.lanes = (unsigned)(tiles < want ? tiles : want);
  sg.smem = (size_t)sg.stages * sg.tile_rows * row_bytes + 128;
  if (sg.smem < 17 * 1024) sg.smem = 17 * 1024;   // the backward reduce reuses the ring as its 16 KB block scratch
  sg.ok = sg.smem <= (size_t)(blocks >= 2 ? 220 / blocks : 200) * 1024;
  return sg;
}
template <typename K>
static void staged_attr(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
// the attribute is per device: remember it per device ordinal
static bool attr_needed(bool (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

template <typename T>
static int pack_weights_t(int kind, const float* w, T* wpk, T* wpkT, int Cout, int Cin, int CoP, int CiP, int src_ohwi,
                          void* stream) {
  if (kind < 0 || kind > 4 || CoP < Cout || CiP < Cin) EW_FAIL(SG2_EINVAL, "pack_weights: bad arguments");
  if (kind == SG2_STEM4x4 && Cin != 48) EW_FAIL(SG2_EINVAL, "pack_weights: stem expects Cin == 48 (3 x 4 x 4)");
  if ((CoP % 8) || (CiP % 8)) EW_FAIL(SG2_EINVAL, "pack_weights: padded extents %% 8");
  if (kind == SG2_STEM4x4) {
    pack_stem_kernel<T><<<grid1d((long long)CoP * CiP), 256, 0, (cudaStream_t)stream>>>(w, wpk, wpkT, Cout, CoP, CiP,
                                                                                      src_ohwi);
    return launch_ok("pack_stem");
  }
  long long nblk = (long long)((CoP + kPackCo - 1) / kPackCo) * ((CiP + kPackCi - 1) / kPackCi);
  if (nblk > 148 * 16) nblk = 148 * 16;
  pack_weights_kernel<T><<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(kind, w, wpk, wpkT, Cout, Cin, CoP, CiP,
                                                                           src_ohwi);
  return launch_ok("pack_weights");
}

extern "C" {

int sg2_pack_weights(int kind, const float* w, void* wpk, void* wpkT, int Cout, int Cin, int CoP, int CiP,
                     int src_ohwi, void* stream) {
  if (kind == SG2_UPCONV3x3 && src_ohwi && CoP == Cout && CiP == Cin && Cout % 64 == 0 && Cin % 64 == 0 &&
      !(reinterpret_cast<uintptr_t>(w) & 15)) {
    pack_upconv_ohwi_kernel<<<(Cout / 64) * (Cin / 64), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wpk,
                                                                                      (__nv_bfloat16*)wpkT, Cout, Cin);
    return launch_ok("pack_upconv_ohwi");
  }
  return pack_weights_t<__nv_bfloat16>(kind, w, (__nv_bfloat16*)wpk, (__nv_bfloat16*)wpkT, Cout, Cin, CoP, CiP, src_ohwi,
                                       stream);
}

int sg2_pack_weights_f32(int kind, const float* w, float* wpk, float* wpkT, int Cout, int Cin, int CoP, int CiP,
                         int src_ohwi, void* stream) {
  return pack_weights_t<float>(kind, w, wpk, wpkT, Cout, Cin, CoP, CiP, src_ohwi, stream);
}

int sg2_pack_transpose(int kind, const void* wpk, void* wpkT, int Cout, int Cin, void* stream) {
  if (kind != SG2_CONV3x3 && kind != SG2_CONV4x4S2 && kind != SG2_GEMM) EW_FAIL(SG2_EINVAL, "pack_transpose: kind %d", kind);
  if ((Cout % 8) || (Cin % 8)) EW_FAIL(SG2_EINVAL, "pack_transpose: channels %% 8");
  const int slots = (kind == SG2_CONV3x3) ? 9 : ((kind == SG2_GEMM) ? 1 : 16);
  long long nblk = (long long)((Cout + 63) / 64) * ((Cin + 63) / 64) * slots;
  if (nblk > 148 * 16) nblk = 148 * 16;
  pack_transpose_kernel<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(kind, (const __nv_bfloat16*)wpk,
                                                                          (__nv_bfloat16*)wpkT, Cout, Cin);
  return launch_ok("pack_transpose");
}

int sg2_unpack_wgrad(int kind, const float* dwpk, float* grad, int Cout, int Cin, int CoP, int CiP, int accumulate,
                     int dst_ohwi, void* stream) {
  if (kind < 0 || kind > 4) EW_FAIL(SG2_EINVAL, "unpack_wgrad: bad kind");
  if (kind == SG2_UPCONV3x3 && dst_ohwi && CoP == Cout && CiP == Cin && Cin % 4 == 0 &&
      !((reinterpret_cast<uintptr_t>(dwpk) | reinterpret_cast<uintptr_t>(grad)) & 15)) {
    unpack_upconv_ohwi_kernel<<<grid1d((long long)Cout * (Cin / 4)), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)dwpk, (float4*)grad, Cout, Cin / 4, accumulate);
    return launch_ok("unpack_upconv_ohwi");
  }
  const int kk = (kind == SG2_CONV3x3 || kind == SG2_UPCONV3x3) ? 9 : ((kind == SG2_GEMM || kind == SG2_STEM4x4) ? 1 : 16);
  (void)kk;
  long long nblk = (long long)Cout * ((Cin + 31) / 32);
  if (nblk > 148 * 32) nblk = 148 * 32;
  unpack_wgrad_kernel<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(
      kind, dwpk, grad, Cout, Cin, CoP, CiP, accumulate, dst_ohwi);
  return launch_ok("unpack_wgrad");
}

int sg2_bn_stats(const void* x, long long P, int C, int groups, double* stats, void* stream) {
  if (C % 8) EW_FAIL(SG2_EINVAL, "bn_stats: C %% 8");
  if (groups < 1 || P % groups) EW_FAIL(SG2_EINVAL, "bn_stats: %lld rows in %d groups", P, groups);
  P /= groups;
  Geo g = make_geo(P, C, 148 * 4);
  g.grid.z = groups;
  bn_stats_kernel<false><<<g.grid, g.block, 0, (cudaStream_t)stream>>>(x, nullptr, P, g.vc, g.cpb, g.rpb, stats, C, 1,
                                                                       0, nullptr, 0);
  return launch_ok("bn_stats");
}

int sg2_f32_to_bf16_stats(const float* x, void* y, long long P, int C, int groups, double* stats, void* stream) {
  return sg2_splitk_finish(x, 1, 0, y, P, C, groups, stats, nullptr, 0, stream);
}

int sg2_splitk_finish(const float* parts, int nsplit, long long slab, void* y, long long P, int C, int groups,
                      double* stats, const void* epi_src, int epi_mode, void* stream) {
  if (C % 8) EW_FAIL(SG2_EINVAL, "splitk_finish: C %% 8");
  if (groups < 1 || P % groups) EW_FAIL(SG2_EINVAL, "splitk_finish: %lld rows in %d groups", P, groups);
  if (nsplit < 1 || (nsplit > 1 && slab < P * C)) EW_FAIL(SG2_EINVAL, "splitk_finish: %d slabs of %lld elements", nsplit, slab);
  if (epi_mode != 0 && epi_mode != SG2_EPI_ADD && epi_mode != SG2_EPI_LRELU_MASK) EW_FAIL(SG2_EINVAL, "splitk_finish: epilogue mode %d", epi_mode);
  if (epi_mode && !epi_src) EW_FAIL(SG2_EINVAL, "splitk_finish: epilogue operand missing");
  P /= groups;
  // split-K outputs are small (the layers that cannot fill the GPU): one or two rows per thread, many blocks
  static const int finish_iters = env_int("SG2_FINISH_ITERS", 4);
  Geo g = make_geo(P, C, 148 * 16, 8, nsplit > 1 ? finish_iters : 4);
  g.grid.z = groups;
  bn_stats_kernel<true><<<g.grid, g.block, 0, (cudaStream_t)stream>>>(parts, (uint4*)y, P, g.vc, g.cpb, g.rpb, stats, C,
                                                                      nsplit, slab, (const uint4*)epi_src, epi_mode);
  return launch_ok("splitk_finish");
}

int sg2_reduce_slabs(const float* parts, int nslabs, long long n, long long slab, float* dst, int accumulate,
                     void* stream) {
  if (nslabs < 1 || (n % 4) || (slab % 4)) EW_FAIL(SG2_EINVAL, "reduce_slabs: %d slabs, n %lld, slab %lld", nslabs, n, slab);
  if ((reinterpret_cast<uintptr_t>(parts) | reinterpret_cast<uintptr_t>(dst)) & 15) EW_FAIL(SG2_EINVAL, "reduce_slabs: 16-byte alignment");
  if (nslabs > 4) {
    long long blocks = (n / 4 + 31) / 32;
    if (blocks > 148 * 16) blocks = 148 * 16;
    reduce_slabs_wide_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)parts, nslabs, n / 4,
                                                                                 slab / 4, (float4*)dst, accumulate);
  } else {
    reduce_slabs_kernel<<<grid1d(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)parts, nslabs, n / 4, slab / 4,
                                                                         (float4*)dst, accumulate);
  }
  return launch_ok("reduce_slabs");
}

int sg2_bn_eval_prepare(const float* running_mean, const float* running_var, float eps, float* mean, float* rstd,
                        int C, void* stream) {
  bn_eval_prepare_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(running_mean, running_var, eps, mean,
                                                                            rstd, C);
  return launch_ok("bn_eval_prepare");
}

int sg2_bn_act_fwd(const void* x, const double* stats, float* mean, float* rstd, const float* gamma, const float* beta,
                   const void* residual, void* out, long long P, int C, int groups, int act, float eps, float momentum,
                   float* running_mean, float* running_var, long long* num_batches_tracked, void* stream) {
  const int Cout = act == ACT_GLU ? C / 2 : C;
  if (Cout % 8) EW_FAIL(SG2_EINVAL, "bn_act_fwd: channels %% 8");
  if (groups < 1 || P % groups) EW_FAIL(SG2_EINVAL, "bn_act_fwd: %lld rows in %d groups", P, groups);
  P /= groups;
  Geo g = make_geo(P, Cout, 148 * 6, kBnW);
  g.grid.z = groups;
  const int has_bn = (mean != nullptr);
  if (stats && !mean) EW_FAIL(SG2_EINVAL, "bn_act_fwd: stats given without mean/rstd outputs");
  cudaStream_t st = (cudaStream_t)stream;
  const StagedGeo sg = make_staged(g, P, groups, C / kBnW, Cout / kBnW, false);
  static bool attr[64] = {};
  if (attr_needed(attr)) {
    staged_attr(bn_act_fwd_kernel<ACT_GLU, true>);
    staged_attr(bn_act_fwd_kernel<ACT_LRELU, true>);
    staged_attr(bn_act_fwd_kernel<ACT_NONE, true>);
  }
#define ARGS (const uint2*)x, mean, rstd, gamma, beta, (const uint2*)residual, (uint2*)out, P, C / kBnW, Cout / kBnW, g.cpb, g.rpb, has_bn, stats, eps, momentum, mean, rstd, running_mean, running_var, num_batches_tracked, sg.tile_rows, sg.stages
  if (sg.ok) {
    dim3 grid(1, sg.lanes, groups);
    if (act == ACT_GLU) bn_act_fwd_kernel<ACT_GLU, true><<<grid, g.block, sg.smem, st>>>(ARGS);
    else if (act == ACT_LRELU) bn_act_fwd_kernel<ACT_LRELU, true><<<grid, g.block, sg.smem, st>>>(ARGS);
    else bn_act_fwd_kernel<ACT_NONE, true><<<grid, g.block, sg.smem, st>>>(ARGS);
  } else {
    if (act == ACT_GLU) bn_act_fwd_kernel<ACT_GLU, false><<<g.grid, g.block, 0, st>>>(ARGS);
    else if (act == ACT_LRELU) bn_act_fwd_kernel<ACT_LRELU, false><<<g.grid, g.block, 0, st>>>(ARGS);
    else bn_act_fwd_kernel<ACT_NONE, false><<<g.grid, g.block, 0, st>>>(ARGS);
  }
#undef ARGS
  return launch_ok("bn_act_fwd");
}

int sg2_bn_act_bwd(const void* x, const void* dout, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, double* sums, void* dx, float* dgamma, float* dbeta, int accumulate,
                   long long P, int C, int groups, int act, void* stream) {
  const int Cout = act == ACT_GLU ? C / 2 : C;
  if (Cout % 8) EW_FAIL(SG2_EINVAL, "bn_act_bwd: channels %% 8");
  if (groups < 1 || P % groups) EW_FAIL(SG2_EINVAL, "bn_act_bwd: %lld rows in %d groups", P, groups);
  P /= groups;
  Geo g = make_geo(P, Cout, 148 * 6, kBnW);
  g.grid.z = groups;
  cudaStream_t st = (cudaStream_t)stream;
  const StagedGeo sg = make_staged(g, P, groups, C / kBnW, Cout / kBnW, true);
  static bool attr[64] = {};
  if (attr_needed(attr)) {
    staged_attr(bn_act_bwd_reduce_kernel<ACT_GLU, true>);
    staged_attr(bn_act_bwd_reduce_kernel<ACT_LRELU, true>);
    staged_attr(bn_act_bwd_reduce_kernel<ACT_NONE, true>);
    staged_attr(bn_act_bwd_apply_kernel<ACT_GLU, true>);
    staged_attr(bn_act_bwd_apply_kernel<ACT_LRELU, true>);
    staged_attr(bn_act_bwd_apply_kernel<ACT_NONE, true>);
  }
#define RARGS (const uint2*)x, (const uint2*)dout, mean, rstd, gamma, beta, P, C / kBnW, Cout / kBnW, g.cpb, g.rpb, sums, C, sg.tile_rows, sg.stages
#define AARGS (const uint2*)x, (const uint2*)dout, mean, rstd, gamma, beta, sums, P, C / kBnW, Cout / kBnW, g.cpb, g.rpb, (uint2*)dx, C, dgamma, dbeta, accumulate, sg.tile_rows, sg.stages
#define SG2_BWD(ACT_)                                                                       \
  if (sg.ok) {                                                                              \
    dim3 grid(1, sg.lanes, groups);                                                         \
    bn_act_bwd_reduce_kernel<ACT_, true><<<grid, g.block, sg.smem, st>>>(RARGS);            \
    bn_act_bwd_apply_kernel<ACT_, true><<<grid, g.block, sg.smem, st>>>(AARGS);             \
  } else {                                                                                  \
    bn_act_bwd_reduce_kernel<ACT_, false><<<g.grid, g.block, 0, st>>>(RARGS);               \
    bn_act_bwd_apply_kernel<ACT_, false><<<g.grid, g.block, 0, st>>>(AARGS);                \
  }
  if (act == ACT_GLU) { SG2_BWD(ACT_GLU) }
  else if (act == ACT_LRELU) { SG2_BWD(ACT_LRELU) }
  else { SG2_BWD(ACT_NONE) }
#undef SG2_BWD
#undef RARGS
#undef AARGS
  return launch_ok("bn_act_bwd");
}

int sg2_lrelu_bwd(const void* x, const void* dout, void* dx, long long n, void* stream) {
  if (n % 8) EW_FAIL(SG2_EINVAL, "lrelu_bwd: n %% 8");
  lrelu_bwd_kernel<<<grid1d(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)dout, (uint4*)dx,
                                                                    n / 8);
  return launch_ok("lrelu_bwd");
}

int sg2_add_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
  if (n % 8) EW_FAIL(SG2_EINVAL, "add_bf16: n %% 8");
  add_bf16_kernel<<<grid1d(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)a, (const uint4*)b, (uint4*)out,
                                                                   n / 8);
  return launch_ok("add_bf16");
}

int sg2_f32_to_bf16(const float* in, void* out, long long n, void* stream) {
  if (n % 4) EW_FAIL(SG2_EINVAL, "f32_to_bf16: n %% 4");
  f32_to_bf16_kernel<<<grid1d(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)in, (uint2*)out, n / 4);
  return launch_ok("f32_to_bf16");
}

int sg2_bf16_to_f32(const void* in, float* out, long long n, void* stream) {
  if (n % 4) EW_FAIL(SG2_EINVAL, "bf16_to_f32: n %% 4");
  bf16_to_f32_kernel<<<grid1d(n / 4), 256, 0, (cudaStream_t)stream>>>((const uint2*)in, (float4*)out, n / 4);
  return launch_ok("bf16_to_f32");
}

int sg2_concat_c(const float* c, const void* h, void* out, int B, int HW, int E, int Ch, void* stream) {
  if (E % 8 || Ch % 8) EW_FAIL(SG2_EINVAL, "concat_c: channels %% 8");
  const long long total = (long long)B * HW * ((E + Ch) / 8);
  concat_c_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(c, (const uint4*)h, (uint4*)out, B, HW, E, Ch);
  return launch_ok("concat_c");
}

int sg2_concat_c_bwd(const void* dcat, void* dh, float* dc, int B, int HW, int E, int Ch, void* stream) {
  if (E % 8 || Ch % 8) EW_FAIL(SG2_EINVAL, "concat_c_bwd: channels %% 8");
  const int ve = E / 8;
  if (ve > 256) EW_FAIL(SG2_EINVAL, "concat_c_bwd: E too large");
  unsigned gx = (unsigned)((HW + 255) / 256);   // >= 256 pixels per block
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  const int threads = (256 / ve) * ve;
  concat_c_bwd_kernel<<<dim3(gx, B), threads, 0, (cudaStream_t)stream>>>((const uint4*)dcat, (uint4*)dh, dc, B, HW, E, Ch);
  return launch_ok("concat_c_bwd");
}

int sg2_head_tanh_fwd(const void* y, float* img, int B, int HW, int CP, void* stream) {
  head_tanh_fwd_kernel<<<grid1d((long long)B * HW), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)y, img, B,
                                                                                   HW, CP);
  return launch_ok("head_tanh_fwd");
}

int sg2_head_tanh_bwd(const float* dimg, const float* img, void* dy, int B, int HW, int CP, void* stream) {
  if (CP % 8) EW_FAIL(SG2_EINVAL, "head_tanh_bwd: CP %% 8");
  head_tanh_bwd_kernel<<<grid1d((long long)B * HW * (CP / 8)), 256, 0, (cudaStream_t)stream>>>(
      dimg, img, (__nv_bfloat16*)dy, B, HW, CP);
  return launch_ok("head_tanh_bwd");
}

int sg2_stem_im2col(const float* img, void* col, int B, int S, void* stream) {
  if (S % 2) EW_FAIL(SG2_EINVAL, "stem_im2col: odd image size");
  if ((12 * (S + 2) + (S / 2) * 33 + 48) * sizeof(float) > 48 * 1024)
    EW_FAIL(SG2_EINVAL, "stem_im2col: image size %d needs more than 48 KB of shared memory (the path's scales are 64 / 128 / 256)", S);
  long long nblk = (long long)B * (S / 2);
  if (nblk > 148 * 16) nblk = 148 * 16;
  stem_im2col_kernel<<<(unsigned)nblk, 256, (12 * (S + 2) + (S / 2) * 33 + 48) * sizeof(float), (cudaStream_t)stream>>>(img, (uint4*)col, B, S);
  return launch_ok("stem_im2col");
}

int sg2_stem_col2im(const void* dcol, float* dimg, int B, int S, void* stream) {
  if (S % 2) EW_FAIL(SG2_EINVAL, "stem_col2im: odd image size");
  if ((size_t)S * 33 * sizeof(uint32_t) > 48 * 1024)
    EW_FAIL(SG2_EINVAL, "stem_col2im: image size %d needs more than 48 KB of shared memory (the path's scales are 64 / 128 / 256)", S);
  const int So = S / 2;
  long long nblk = (long long)B * (So + 1);
  if (nblk > 148 * 16) nblk = 148 * 16;
  stem_col2im_kernel<<<(unsigned)nblk, 256, (size_t)2 * So * 33 * sizeof(uint32_t), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dcol, dimg, B, S);
  return launch_ok("stem_col2im");
}

int sg2_nhwc_to_nchw_f32(const void* in, float* out, int B, int HW, int C, void* stream) {
  nhwc_bf16_to_nchw_f32_kernel<<<grid1d((long long)B * HW * C), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in, out, B, HW, C);
  return launch_ok("nhwc_to_nchw_f32");
}

int sg2_nchw_f32_to_nhwc(const float* in, void* out, int B, int HW, int C, void* stream) {
  nchw_f32_to_nhwc_bf16_kernel<<<grid1d((long long)B * HW * C), 256, 0, (cudaStream_t)stream>>>(
      in, (__nv_bfloat16*)out, B, HW, C);
  return launch_ok("nchw_f32_to_nhwc");
}

}  // extern "C"
