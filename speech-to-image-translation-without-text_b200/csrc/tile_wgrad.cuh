// Tile-resident weight-gradient kernel on tcgen05 + TMA, sm_100a.
//
//   dw[co, job(group, tap), ci] += sum over pixels p of  dy[p, co] * x[p + shift(tap), ci]
//
// Both operands are MN-major UMMA operands (channels contiguous, pixels = the GEMM K axis). The gather form in
// igemm.cuh runs one CTA per tap, so every tap re-loads the dy tile and a shifted copy of the x tile. Here a CTA loads
// ONE dy tile (8 x 16 pixels) and ONE x halo box per pixel tile and issues every tap as a descriptor that starts at a
// shifted row of the halo box (see tile_conv.cuh for the addressing argument); tap t accumulates into its own TMEM
// column block [t * BN, (t+1) * BN), which stays resident over all pixel tiles of the CTA. One epilogue per CTA adds
// the 128 x (taps * BN) fp32 tile into dw with red.global.add.v4.f32.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = epilogue.
#pragma once
#include "tile_conv.cuh"

namespace sg2 {

struct TileWgradParams {
  CUtensorMap tmA[4];  // dy sources:   boxes {CWA channels, 8, 16, 1}
  CUtensorMap tmB[4];  // x sources:    boxes {CWB channels, pitch, ph, 1}
  uint32_t tap_off16[4][16];  // [group][tap] window start inside the x halo box (16-byte descriptor units)
  int tap_job[4][16];         // [group][tap] job index in dw[Cout][njobs][Cin]
  int a_src[4], b_src[4];     // [group] dy / x source
  int org_y[4], org_x[4];     // [x source] halo box origin relative to the tile origin
  int ngroups, ntaps;         // taps per group
  int taps_cta, tap_sets;     // taps handled by one CTA; tap_sets = ceil(ntaps / taps_cta)
  int njobs;
  int cwa, cwb;               // channels per swizzle chunk (64/32/16) of dy / x
  int a_chunks, b_chunks;     // chunks actually loaded per tile (Cout tile / cwa, BN / cwb)
  int a_chunk_bytes, b_chunk_bytes;  // SMEM bytes of one chunk box (1024-aligned)
  int a_box_bytes, b_box_bytes;      // bytes one chunk box transfers
  int pitch, ph;
  int bn;                     // Cin tile = UMMA N
  int m_tiles, n_tiles;
  int tiles_x, tiles_y, B;
  uint32_t magic_img, magic_x;
  int lanes;                  // CTAs sharing one (group, tap set, m tile, n tile): they split the pixel tiles
  int Cout, Cin;
  float* dw;
  int stages;
  float* partials;  // deterministic mode: CTA lane l STORES its 128 x (taps x BN) tile into slab l (partials + l * slab,
  long long slab;   // laid out like dw); the caller sums the `lanes` slabs in order. NULL: red.global.add into dw
  int merge3;  // 1: the three taps of a filter row run as ONE MMA of N = 3 * bn: the x operand's chunk stride (LBO) is one
               //    pixel, so chunk j is the window shifted by j pixels; needs bn == cwb and taps ordered (row, col)
};

__global__ void __launch_bounds__(kNumThreads, 1) tile_wgrad_kernel(const __grid_constant__ TileWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  // M = 128 rows of dy channels: the descriptor always spans 128 / cwa chunks; chunks past a_chunks are never loaded
  // (their accumulator rows are never stored), but the SMEM behind them must exist.
  const int a_stage_bytes = (kBlockM / p.cwa) * p.a_chunk_bytes;
  const int b_stage_bytes = p.b_chunks * p.b_chunk_bytes;
  const int stage_bytes = a_stage_bytes + b_stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(S) * stage_bytes);
  uint64_t* empty = full + S;
  uint64_t* tmem_full = empty + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work assignment: blockIdx.x = ((((group * tap_sets + tset) * m_tiles + mt) * n_tiles + nt) * lanes + my_lane)
  int bid = blockIdx.x;
  const int my_lane = bid % p.lanes;
  bid /= p.lanes;
  const int nt = bid % p.n_tiles;
  bid /= p.n_tiles;
  const int mt = bid % p.m_tiles;
  bid /= p.m_tiles;
  const int tset = bid % p.tap_sets;
  const int g = bid / p.tap_sets;
  const int m0 = mt * kBlockM, n0 = nt * p.bn;
  const int tap0 = tset * p.taps_cta;
  const int ntap = min(p.taps_cta, p.ntaps - tap0);
  const int tiles_img = p.tiles_x * p.tiles_y;
  const int pix_tiles = tiles_img * p.B;
  const int my_tiles = my_lane < pix_tiles ? (pix_tiles - my_lane + p.lanes - 1) / p.lanes : 0;
  const int a_chunks = min(p.a_chunks, (p.Cout - m0 + p.cwa - 1) / p.cwa);
  const int need_cols = p.taps_cta * p.bn;
  const int tmem_cols = need_cols <= 32 ? 32 : (need_cols <= 64 ? 64 : (need_cols <= 128 ? 128 : (need_cols <= 256 ? 256 : 512)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[p.a_src[g]]);
    tma_prefetch_desc(&p.tmB[p.b_src[g]]);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================= TMA producer
    const int asrc = p.a_src[g], bsrc = p.b_src[g];
    const uint32_t tx_bytes = a_chunks * p.a_box_bytes + p.b_chunks * p.b_box_bytes;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one()) {
        const int t = my_lane + i * p.lanes;
        int b, ty, tx;
        b = tiles_img == 1 ? t : (int)__umulhi((uint32_t)t, p.magic_img);
        const int r = t - b * tiles_img;
        ty = p.tiles_x == 1 ? r : (int)__umulhi((uint32_t)r, p.magic_x);
        tx = r - ty * p.tiles_x;
        const int x0 = tx * kTileW, y0 = ty * kTileH;
        mbar_expect_tx(&full[s], tx_bytes);
        uint8_t* sa = smem + size_t(s) * stage_bytes;
        uint8_t* sb = sa + a_stage_bytes;
        for (int c = 0; c < a_chunks; ++c)
          tma_load_4d(&p.tmA[asrc], &full[s], sa + c * p.a_chunk_bytes, m0 + c * p.cwa, x0, y0, b);
        for (int c = 0; c < p.b_chunks; ++c)
          tma_load_4d(&p.tmB[bsrc], &full[s], sb + c * p.b_chunk_bytes, n0 + c * p.cwb, x0 + p.org_x[bsrc], y0 + p.org_y[bsrc],
                      b);
      }
      __syncwarp();
      if (++s == S) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    const bool merge = p.merge3 != 0;
    const uint32_t idesc = make_idesc_bf16(kBlockM, merge ? 3 * p.bn : p.bn, 1, 1);
    const uint32_t row_a = uint32_t(p.cwa * 2), row_b = uint32_t(p.cwb * 2);
    // MN-major: LBO = bytes between consecutive channel chunks, SBO = bytes between 8-pixel groups (= one tile row)
    const uint64_t adesc0 = make_smem_desc(smem_u32(smem), uint32_t(p.a_chunk_bytes), 8 * row_a, swizzle_code(int(row_a)));
    const uint64_t bdesc0 = make_smem_desc(smem_u32(smem) + uint32_t(a_stage_bytes), merge ? row_b : uint32_t(p.b_chunk_bytes),
                                           uint32_t(p.pitch) * row_b, swizzle_code(int(row_b)));
    const uint32_t stage16 = uint32_t(stage_bytes) >> 4;
    const uint32_t ka16 = (16u * row_a) >> 4;                        // 16 pixels = two dense tile rows of dy
    const uint32_t kb16 = (2u * uint32_t(p.pitch) * row_b) >> 4;     // two halo rows of x
    const uint32_t bn = uint32_t(p.bn);
    const int tstep = merge ? 3 : 1;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_st = adesc0 + uint64_t(uint32_t(s) * stage16);
        const uint64_t b_st = bdesc0 + uint64_t(uint32_t(s) * stage16);
        const uint32_t acc0 = i > 0 ? 1u : 0u;
        for (int t = 0; t < ntap; t += tstep) {
          const uint64_t b_tap = b_st + uint64_t(p.tap_off16[g][tap0 + t]);
          const uint32_t d = tmem_base + uint32_t(t) * bn;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16(d, a_st + uint64_t(uint32_t(k) * ka16), b_tap + uint64_t(uint32_t(k) * kb16), idesc, k > 0 ? 1u : acc0);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
      if (++s == S) {
        s = 0;
        ph ^= 1;
      }
    }
    if (elect_one()) umma_commit(tmem_full);
    __syncwarp();
  } else {
    // ================================================================= epilogue (once per CTA)
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool slabs = p.partials != nullptr;
    const bool valid = m < p.Cout && (my_tiles > 0 || slabs);
    float* const base = slabs ? p.partials + (long long)my_lane * p.slab : p.dw;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
    for (int t = 0; t < ntap; ++t) {
      float* rowp = base + ((long long)m * p.njobs + p.tap_job[g][tap0 + t]) * p.Cin + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        uint32_t v[32];
        if (p.bn - c0 >= 32) {
          tmem_ld_32x32(taddr + uint32_t(t * p.bn + c0), v);
        } else {
          uint32_t w[16];
          tmem_ld_32x16(taddr + uint32_t(t * p.bn + c0), w);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = w[j];
#pragma unroll
          for (int j = 16; j < 32; ++j) v[j] = 0u;
        }
        tmem_ld_wait();
        if (valid && slabs) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c0 + j < p.bn) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              if (my_tiles <= 0) o = make_float4(0.f, 0.f, 0.f, 0.f);
              *reinterpret_cast<float4*>(rowp + c0 + j) = o;
            }
          }
        } else if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c0 + j < p.bn) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(rowp + c0 + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                           : "memory");
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace sg2
