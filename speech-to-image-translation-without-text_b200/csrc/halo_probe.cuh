// EXPERIMENT: conv3x3 fprop where the activation tile is loaded ONCE per channel chunk as a halo box
// {BK channels, PITCH, 18, 1} and the nine taps are nine UMMA descriptors that start at shifted rows of that box.
// Purpose: find out whether a K-major swizzled UMMA descriptor may start at an address that is not aligned to the
// swizzle atom (row offset inside the atom), and with an SBO that is not a multiple of the atom size.
#pragma once
#include "igemm.cuh"

namespace sg2 {

struct HaloProbeParams {
  CUtensorMap tmA, tmB;
  int kchunks;   // Cin / BK
  int pitch;     // halo box width in pixels (>= 10)
  int bo_mode;   // 0: base_offset = 0, 1: base_offset = (start >> 7) & 7
  int tiles_x, tiles_y;
  int W, H, B, N;
  void* out;
  int stages;
};

constexpr int kHaloTW = 8, kHaloTH = 16;

template <int BN, int BK>
__global__ void __launch_bounds__(kNumThreads) halo_probe_kernel(const __grid_constant__ HaloProbeParams p) {
  constexpr int kRowB = BK * 2;
  constexpr int kBBytes = BN * kRowB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const int a_bytes = ((p.pitch * (kHaloTH + 2) * kRowB) + 1023) & ~1023;
  uint8_t* sA = smem;                       // 2 halo buffers
  uint8_t* sB = smem + 2 * a_bytes;         // S weight stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + size_t(S) * kBBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + 2;
  uint64_t* b_full = bars + 4;
  uint64_t* b_empty = b_full + S;
  uint64_t* tmem_full = b_empty + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int b = t / p.tiles_y;
  const int x0 = tx * kHaloTW, y0 = ty * kHaloTH;
  const int n0 = blockIdx.y * BN;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int ch = 0; ch < p.kchunks; ++ch) {
        const int as = ch & 1;
        mbar_wait(&a_empty[as], ((ch >> 1) & 1) ^ 1);
        mbar_expect_tx(&a_full[as], p.pitch * (kHaloTH + 2) * kRowB);
        tma_load_4d(&p.tmA, &a_full[as], sA + as * a_bytes, ch * BK, x0 - 1, y0 - 1, b);
        for (int tap = 0; tap < 9; ++tap, ++it) {
          const int s = it % S;
          mbar_wait(&b_empty[s], ((it / S) & 1) ^ 1);
          mbar_expect_tx(&b_full[s], kBBytes);
          tma_load_2d(&p.tmB, &b_full[s], sB + size_t(s) * kBBytes, (tap * p.kchunks + ch) * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, 0, 0);
      constexpr uint32_t swc = swizzle_code(kRowB);
      int it = 0;
      for (int ch = 0; ch < p.kchunks; ++ch) {
        const int as = ch & 1;
        mbar_wait(&a_full[as], (ch >> 1) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + as * a_bytes);
        for (int tap = 0; tap < 9; ++tap, ++it) {
          const int s = it % S;
          mbar_wait(&b_full[s], (it / S) & 1);
          tc_fence_after();
          const int dy = tap / 3, dx = tap % 3;   // offsets inside the halo box (origin = (y0-1, x0-1))
          const uint32_t a_start = a_base + uint32_t((dy * p.pitch + dx) * kRowB);
          uint64_t adesc = make_smem_desc(a_start, 16, uint32_t(p.pitch * kRowB), swc);
          if (p.bo_mode == 1) adesc |= uint64_t((a_start >> 7) & 7u) << 49;
          const uint64_t bdesc = make_smem_desc(smem_u32(sB + size_t(s) * kBBytes), 16, 8 * kRowB, swc);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(tmem_base, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(&b_empty[s]);
        }
        umma_commit(&a_empty[as]);
      }
      umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int xi = row % kHaloTW, yi = row / kHaloTW;
    const int x = x0 + xi, y = y0 + yi;
    const bool valid = (x < p.W) && (y < p.H);
    const long long off = (((long long)b * p.H + y) * p.W + x) * p.N + n0;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + c0, v);
      tmem_ld_wait();
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
          dst[j] = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace sg2
