// Tile-resident implicit-GEMM convolution (fprop / dgrad) on tcgen05 + TMA, sm_100a.
//
// The gather kernel in igemm.cuh re-loads the 128-pixel activation tile once per filter tap, so a 3x3 layer moves
// 9x its input through L2->SMEM; at ~42 B/cycle/SM of TMA feed that caps a 128x128 tile at about a third of the
// tensor peak. Here the activation tile is loaded ONCE per (channel chunk, source) as a halo box
// {BK channels, PITCH, PH, 1 image} (TMA; zero fill outside the image = the conv padding) and every tap is a UMMA
// shared-memory descriptor that STARTS AT A SHIFTED ROW of that box: an 8-pixel-wide output tile makes each 8-row
// core-matrix group one tile row, so SBO = PITCH * row bytes walks down the rows and the tap offset
// (oy * PITCH + ox) * row bytes moves the window. The hardware applies the 128/64/32 B swizzle on absolute
// shared-memory address bits (validated by tools/probe_halo.py), so TMA's write pattern and the shifted descriptor's
// read pattern agree without any base-offset correction.
//
// Work unit = MT (1 or 2) pixel tiles of 8 x 16 pixels: both tiles multiply against the same weight stage (two TMEM
// accumulators), which halves the weight bytes streamed per MMA. Small layers keep their whole weight slice resident
// in SMEM (b_resident). Persistent CTAs: each CTA owns one (output group, BN-channel slice) and strides over the
// units; accumulators are double buffered in TMEM (2 x MT x BN columns) so the epilogue of unit i overlaps the MMAs
// of unit i+1; BatchNorm batch statistics are accumulated in shared memory across all units of the CTA.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer + TMEM owner, warps 2-9 = epilogue
// (two warps per TMEM lane quarter, splitting the 32-column blocks between them).
// Producer and issuer loops run warp-convergent with one elected lane issuing and read per-tap constants through
// uniform loads from the kernel parameters: a divergent single-thread issue loop costs ~146 cycles per tcgen05.mma
// (tools/probe_mma.py) against a tensor floor of N/2 cycles.
#pragma once
#include "igemm.cuh"

namespace sg2 {

constexpr int kTileW = 8, kTileH = 16;  // output tile: 8 x 16 pixels of one image = 128 GEMM rows

struct TileParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  uint32_t tap_off16[4][16];  // [group][tap] window start inside its source's halo box, in 16-byte descriptor units
  int tap_kblk[4][16];        // [group][tap] weight K-block of the tap: B column = (kblk * kchunks + ch) * BK
  int src_begin[5];           // taps [src_begin[s], src_begin[s+1]) read source s (tap lists are sorted by source)
  int org_y[4], org_x[4];     // box origin of each source relative to the tile origin (<= 0)
  int nsrc, ntaps, ngroups, kchunks;
  int pitch, ph;              // halo box extent in pixels (all sources)
  int a_box_bytes;            // bytes one halo box occupies in SMEM (1024-aligned)
  int box_bytes;              // bytes one halo box transfers
  int mt;                     // pixel tiles per work unit (1 or 2)
  int tiles_x, tiles_y, B;
  int n_tiles;                // N / BN
  int lanes;                  // CTAs per (group, n-tile) combination; gridDim.x = ngroups * n_tiles * lanes
  int Wo, Ho, N;
  long long out_off[4], sb, sy, sx;
  void* out;
  double* stats;   // optional [groups][2][N] fp64 (+=): per-channel sum / sum of squares of the bf16-rounded outputs.
                   // Per-CTA partial sums are formed in a fixed order in fp32; the cross-CTA accumulation is an fp64
                   // atomic, whose result does not depend on the arrival order beyond 2^-53 (run-to-run reproducible)
  int stats_bg;    // images per statistics group (0: the whole batch is one group)
  int stages;      // weight ring depth (ring mode), in stages of `tps` taps
  int tps;         // taps per weight stage (divides the taps of every source)
  uint32_t stage_start_mask, stage_end_mask;  // bit i: tap i of a step opens / closes a weight stage (ring mode)
  int b_resident;  // 1: all ntaps*kchunks weight tiles of this CTA's (group, n-tile) stay in SMEM for the whole kernel
  int na;          // halo stages (2..4): the producer runs na-1 (chunk, source) steps ahead of the MMA issuer
  int act;         // epilogue activation: 0 none, 2 LeakyReLU(0.2) (layers without BatchNorm)
  // per-sample, per-border-region bias [B][9][N] added to the accumulators before statistics / store: the
  // contribution of input channels that are constant over the image (the broadcast c_code of a jointConv)
  const float* bias9;
  // epilogue operand with the layout of the output tensor (bf16): epi_mode 1 adds it (residual gradient), 2 scales the
  // result by LeakyReLU'(src) = (src > 0 ? 1 : 0.2) (backward of a LeakyReLU whose output is src)
  const void* epi_src;
  int epi_mode;
  uint32_t magic_img, magic_x;  // ceil(2^32 / tiles_img), ceil(2^32 / tiles_x): division-free tile decoding
  long long* dbg_out;  // diagnostics: per-CTA wait-cycle counters (SG2_TILE_DBG & 64)
  int dbg;
};

constexpr int kTileThreads = 320;

__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long& acc_cycles, bool on) {
  if (!on) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc_cycles += clock64() - t0;
}

template <int BN, int BK>
struct TileCfg {
  static constexpr int kRowB = BK * 2;
  static constexpr int kBBytes = BN * kRowB;
  static constexpr int kSN = BN < 32 ? 32 : BN;
  // Column block one epilogue warp handles per step. 32 in general; BN = 32 is split into two 16-column blocks so that
  // BOTH warps of a TMEM lane quarter work (with one 32-column block the second warp of each quarter idles and the
  // epilogue, not the MMAs, bounds the 32-channel layers of generator stage 3: tools/tile_waits.py, 1.4k vs 0.9k cycles).
  static constexpr int kCB = BN == 32 ? 16 : 32;
  static constexpr int kChunks = (BN + kCB - 1) / kCB;  // column blocks of the accumulator
  static constexpr int kCPW = (kChunks + 1) / 2;        // blocks per epilogue warp (two warps share a lane quarter)
  // BatchNorm sums live in registers across all units of the CTA only while that costs 64 registers per thread
  // (one 32-column block per warp). Two blocks per warp (BN = 128) would need 128 accumulator registers on top of the
  // 32-wide TMEM read: ptxas spills them (664 bytes of spill loads per thread) and the epilogue runs ~5x slower than
  // the MMAs it is supposed to hide behind (tools/tile_waits.py: 7.1k cycles per 128 x 128 tile against 2.3k). Wider
  // tiles therefore reduce each 32 x 32 block across the warp's rows with one butterfly and keep per-quarter
  // partial sums in shared memory.
  static constexpr bool kRegStats = kCPW <= 1;
  __host__ __device__ static int tmem_cols(int mt) {
    const int need = 2 * mt * BN;
    return need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
  }
};

// tile index -> (image, tile row, tile column) without integer division (exact while t * d < 2^32)
__device__ __forceinline__ void tile_decode(const TileParams& p, int t, int tiles_img, int& b, int& ty, int& tx) {
  b = tiles_img == 1 ? t : (int)__umulhi((uint32_t)t, p.magic_img);  // (2^32 / 1 does not fit the 32-bit magic)
  const int r = t - b * tiles_img;
  ty = p.tiles_x == 1 ? r : (int)__umulhi((uint32_t)r, p.magic_x);
  tx = r - ty * p.tiles_x;
}

struct TileSmem {
  uint8_t *sA, *sB;
  uint64_t *a_full, *a_empty, *t_full, *t_empty, *b_full, *b_empty;
};

// MMA issuer body (whole warp convergent, elected lane issues). NT = taps per (chunk, source) step, fully unrolled with
// the window offsets in registers; MT2 = two pixel tiles per unit (second accumulator at +BN columns).
template <int BN, int BK, int NT, bool MT2>
__device__ __forceinline__ void tile_mma_loop(const TileParams& p, const TileSmem& sm, uint32_t tmem_base, int g, int my_units,
                                              int a_stage_bytes, int stage_bytes) {
  using Cfg = TileCfg<BN, BK>;
  constexpr int kRowB = Cfg::kRowB;
  constexpr int MT = MT2 ? 2 : 1;
  constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, 0, 0);
  constexpr uint32_t swc = swizzle_code(kRowB);
  constexpr uint32_t kB16 = uint32_t(Cfg::kBBytes) >> 4;
  const uint64_t adesc0 = make_smem_desc(smem_u32(sm.sA), 16, uint32_t(p.pitch * kRowB), swc);
  const uint64_t bdesc0 = make_smem_desc(smem_u32(sm.sB), 16, 8 * kRowB, swc);
  const uint32_t a_stage16 = uint32_t(a_stage_bytes) >> 4, a_box16 = uint32_t(p.a_box_bytes) >> 4;
  const uint32_t stage16 = uint32_t(stage_bytes) >> 4;
  const int nsrc = p.nsrc, kchunks = p.kchunks, S = p.stages, NA = p.na;
  const bool resident = p.b_resident != 0;
  const uint32_t start_mask = p.stage_start_mask, end_mask = p.stage_end_mask;
  const int lane = threadIdx.x & 31;
  uint32_t off[NT];
  if (nsrc == 1) {
#pragma unroll
    for (int i = 0; i < NT; ++i) off[i] = p.tap_off16[g][i];
  }
  if (resident && my_units > 0) {
    mbar_wait(&sm.b_full[0], 0);
    tc_fence_after();
  }
  const bool tm_on = (p.dbg & 64) != 0;
  long long w_te = 0, w_af = 0, w_bf = 0;
  const long long tstart = clock64();
  int as_c = 0, bs = 0, acc = 0;
  uint32_t aph = 0, bph = 0, tph = 0;
  for (int ul = 0; ul < my_units; ++ul) {
    mbar_wait_t(&sm.t_empty[acc], tph ^ 1, w_te, tm_on);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + uint32_t(acc * MT * BN);
    auto issue_tap = [&](uint64_t adesc, uint64_t bdesc, uint32_t accum) {
#pragma unroll
      for (int k = 0; k < BK / 16; ++k)
        umma_f16(d_tmem, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (k > 0) ? 1u : accum);
      if (MT2) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_f16(d_tmem + uint32_t(BN), adesc + uint64_t(a_box16) + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc,
                   (k > 0) ? 1u : accum);
      }
    };
    uint32_t b_res16 = 0;  // resident mode: running offset of this step's first weight tile
    for (int ch = 0; ch < kchunks; ++ch) {
      for (int src = 0; src < nsrc; ++src) {
        mbar_wait_t(&sm.a_full[as_c], aph, w_af, tm_on);
        tc_fence_after();
        const uint64_t a_stage = adesc0 + uint64_t(uint32_t(as_c) * a_stage16);
        if (nsrc > 1) {
#pragma unroll
          for (int i = 0; i < NT; ++i) off[i] = p.tap_off16[g][src * NT + i];
        }
        const uint32_t accum0 = (ch == 0 && src == 0) ? 0u : 1u;
        if (resident) {
          if (elect_one()) {
            const uint64_t b_step = bdesc0 + uint64_t(b_res16);
#pragma unroll
            for (int i = 0; i < NT; ++i)
              issue_tap(a_stage + uint64_t(off[i]), b_step + uint64_t(uint32_t(i) * kB16), i == 0 ? accum0 : 1u);
            umma_commit(&sm.a_empty[as_c]);
          }
          __syncwarp();
          b_res16 += uint32_t(NT) * kB16;
        } else {
          uint32_t u16 = 0;
#pragma unroll
          for (int i = 0; i < NT; ++i) {
            if ((start_mask >> i) & 1u) {
              mbar_wait_t(&sm.b_full[bs], bph, w_bf, tm_on);
              tc_fence_after();
              u16 = 0;
            }
            const bool closes = ((end_mask >> i) & 1u) != 0;
            if (elect_one()) {
              issue_tap(a_stage + uint64_t(off[i]), bdesc0 + uint64_t(uint32_t(bs) * stage16 + u16), i == 0 ? accum0 : 1u);
              if (closes) umma_commit(&sm.b_empty[bs]);
              if (i == NT - 1) umma_commit(&sm.a_empty[as_c]);
            }
            __syncwarp();
            u16 += kB16;
            if (closes) {
              if (++bs == S) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
        }
        if (++as_c == NA) {
          as_c = 0;
          aph ^= 1;
        }
      }
    }
    if (elect_one()) umma_commit(&sm.t_full[acc]);
    __syncwarp();
    acc ^= 1;
    if (acc == 0) tph ^= 1;
  }
  if (tm_on && lane == 0) {
    p.dbg_out[blockIdx.x * 16 + 4] = clock64() - tstart;
    p.dbg_out[blockIdx.x * 16 + 5] = w_te;
    p.dbg_out[blockIdx.x * 16 + 6] = w_af;
    p.dbg_out[blockIdx.x * 16 + 7] = w_bf;
  }
}

template <int BN, int BK, int NT>
__global__ void __launch_bounds__(kTileThreads, 1) tile_conv_kernel(const __grid_constant__ TileParams p) {
  using Cfg = TileCfg<BN, BK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const int NA = p.na;
  const int MT = p.mt;
  const int a_stage_bytes = MT * p.a_box_bytes;
  const int stage_bytes = p.tps * Cfg::kBBytes;
  const int b_total = p.b_resident ? p.ntaps * p.kchunks * Cfg::kBBytes : S * stage_bytes;
  TileSmem sm;
  sm.sA = smem;
  sm.sB = smem + NA * a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm.sB + b_total);
  sm.a_full = bars;        // [4]
  sm.a_empty = bars + 4;   // [4]
  sm.t_full = bars + 8;    // [2]
  sm.t_empty = bars + 10;  // [2]
  sm.b_full = bars + 12;   // [16]
  sm.b_empty = bars + 28;  // [16]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 44);
  // wide tiles (BN > 128): per-TMEM-lane-quarter partial sums, each slot written by exactly one warp (no atomics)
  __shared__ float s_stats[Cfg::kRegStats ? 2 : 4 * 2 * Cfg::kSN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- static work assignment: this CTA's (group, n-tile) is fixed, it strides over the units of MT pixel tiles
  const int combos = p.ngroups * p.n_tiles;
  const int combo = blockIdx.x % combos;
  const int my_lane = blockIdx.x / combos;
  const int g = combo % p.ngroups;
  const int n0 = (combo / p.ngroups) * BN;
  const int tiles_img = p.tiles_x * p.tiles_y;
  const int pix_tiles = tiles_img * p.B;
  const int units = (pix_tiles + MT - 1) / MT;
  const int my_units = my_lane < units ? (units - my_lane + p.lanes - 1) / p.lanes : 0;
  const int tmem_cols = Cfg::tmem_cols(MT);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA[0]);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&sm.a_full[s], 1);
      mbar_init(&sm.a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.t_full[s], 1);
      mbar_init(&sm.t_empty[s], 8);  // one arrive per epilogue warp
    }
    for (int s = 0; s < 16; ++s) {
      mbar_init(&sm.b_full[s], 1);
      mbar_init(&sm.b_empty[s], 1);
    }
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
    const int steps_per_unit = p.kchunks * p.nsrc;
    const int total = my_units * steps_per_unit;  // (chunk, source) steps this CTA walks through
    if (p.b_resident && total > 0) {
      if (elect_one()) {
        mbar_expect_tx(&sm.b_full[0], p.ntaps * p.kchunks * Cfg::kBBytes);
        for (int ch = 0; ch < p.kchunks; ++ch)
          for (int tap = 0; tap < p.ntaps; ++tap)
            tma_load_2d(&p.tmB, &sm.b_full[0], sm.sB + size_t(ch * p.ntaps + tap) * Cfg::kBBytes,
                        (p.tap_kblk[g][tap] * p.kchunks + ch) * BK, g * p.N + n0);
      }
      __syncwarp();
    }
    const bool tm_on = (p.dbg & 64) != 0;
    long long w_ae = 0, w_be = 0;
    const long long tstart = clock64();
    int ja = 0, as_p = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    int a_ul = 0, a_ch = 0, a_src = 0;  // (local unit, chunk, source) of halo step `ja`
    auto issue_a = [&]() {
      mbar_wait_t(&sm.a_empty[as_p], aph ^ 1, w_ae, tm_on);
      if (elect_one()) {
        mbar_expect_tx(&sm.a_full[as_p], MT * p.box_bytes);
        const int u = my_lane + a_ul * p.lanes;
        for (int m = 0; m < MT; ++m) {
          int t = u * MT + m;
          t = t < pix_tiles ? t : pix_tiles - 1;  // odd tail: the second tile repeats the last one (never stored)
          int b, ty, tx;
          tile_decode(p, t, tiles_img, b, ty, tx);
          tma_load_4d(&p.tmA[a_src], &sm.a_full[as_p], sm.sA + as_p * a_stage_bytes + m * p.a_box_bytes, a_ch * BK,
                      tx * kTileW + p.org_x[a_src], ty * kTileH + p.org_y[a_src], b);
        }
      }
      __syncwarp();
      ++ja;
      if (++a_src == p.nsrc) {
        a_src = 0;
        if (++a_ch == p.kchunks) {
          a_ch = 0;
          ++a_ul;
        }
      }
      if (++as_p == NA) {
        as_p = 0;
        aph ^= 1;
      }
    };
    while (ja < total && ja < NA - 1) issue_a();
    int ch = 0, src = 0;
    for (int j = 0; j < total; ++j) {
      if (!p.b_resident) {
        const int t_end = p.src_begin[src + 1];
        for (int tap = p.src_begin[src]; tap < t_end; tap += p.tps) {
          mbar_wait_t(&sm.b_empty[bs], bph ^ 1, w_be, tm_on);
          if (elect_one()) {
            mbar_expect_tx(&sm.b_full[bs], stage_bytes);
            for (int u = 0; u < p.tps; ++u)
              tma_load_2d(&p.tmB, &sm.b_full[bs], sm.sB + size_t(bs) * stage_bytes + u * Cfg::kBBytes,
                          (p.tap_kblk[g][tap + u] * p.kchunks + ch) * BK, g * p.N + n0);
          }
          __syncwarp();
          if (++bs == S) {
            bs = 0;
            bph ^= 1;
          }
        }
      }
      if (++src == p.nsrc) {
        src = 0;
        if (++ch == p.kchunks) ch = 0;
      }
      if (ja < total) issue_a();
    }
    if (tm_on && lane == 0) {
      p.dbg_out[blockIdx.x * 16 + 0] = clock64() - tstart;
      p.dbg_out[blockIdx.x * 16 + 1] = w_ae;
      p.dbg_out[blockIdx.x * 16 + 2] = w_be;
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (MT == 2)
      tile_mma_loop<BN, BK, NT, true>(p, sm, tmem_base, g, my_units, a_stage_bytes, stage_bytes);
    else
      tile_mma_loop<BN, BK, NT, false>(p, sm, tmem_base, g, my_units, a_stage_bytes, stage_bytes);
  } else {
    // ================================================================= epilogue: warps 2..9; TMEM lane quarter = warp % 4,
    // the two warps of a quarter take alternate 32-column blocks
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int xi = row % kTileW, yi = row / kTileW;
    const int et = threadIdx.x - 64;
    const bool do_stats = p.stats != nullptr;
    float acc_s[Cfg::kRegStats ? Cfg::kCPW : 1][32], acc_q[Cfg::kRegStats ? Cfg::kCPW : 1][32];
    if constexpr (Cfg::kRegStats) {
#pragma unroll
      for (int ci = 0; ci < Cfg::kCPW; ++ci)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc_s[ci][j] = acc_q[ci][j] = 0.f;
    } else {
      if (do_stats) {
        for (int i = et; i < 4 * 2 * Cfg::kSN; i += 256) s_stats[i] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    // the running statistics belong to one statistics group (sub-batch) at a time; tiles arrive in image order
    int cur_grp = -1;
    auto flush_stats = [&](int grp) {
      if constexpr (!Cfg::kRegStats) {
        // every epilogue warp reaches this point for the same tile (the group only depends on the tile's image)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        double* st = p.stats + (long long)grp * 2 * p.N;
        for (int i = et; i < BN; i += 256) {
          float cs = 0.f, cq = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            cs += s_stats[qq * 2 * Cfg::kSN + i];
            cq += s_stats[qq * 2 * Cfg::kSN + Cfg::kSN + i];
          }
          atomicAdd(&st[n0 + i], (double)cs);
          atomicAdd(&st[p.N + n0 + i], (double)cq);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = et; i < 4 * 2 * Cfg::kSN; i += 256) s_stats[i] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if constexpr (Cfg::kRegStats) {
        double* st = p.stats + (long long)grp * 2 * p.N;
#pragma unroll
        for (int ci = 0; ci < Cfg::kCPW; ++ci) {
          const int c0 = (half + 2 * ci) * Cfg::kCB;
          if (c0 < BN) {
            const float cs = warp_transpose_sum(acc_s[ci], lane);
            const float cq = warp_transpose_sum(acc_q[ci], lane);
            if (lane < Cfg::kCB && c0 + lane < BN) {
              atomicAdd(&st[n0 + c0 + lane], (double)cs);
              atomicAdd(&st[p.N + n0 + c0 + lane], (double)cq);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) acc_s[ci][j] = acc_q[ci][j] = 0.f;
          }
        }
      }
    };
    const bool tm_on = (p.dbg & 64) != 0;
    long long w_tf = 0;
    const long long tstart = clock64();
    int acc = 0;
    uint32_t tph = 0;
    for (int ul = 0; ul < my_units; ++ul) {
      const int u = my_lane + ul * p.lanes;
      mbar_wait_t(&sm.t_full[acc], tph, w_tf, tm_on);
      tc_fence_after();
      for (int m = 0; m < MT; ++m) {
        const int t = u * MT + m;
        const bool tile_ok = t < pix_tiles;
        int b, ty, tx;
        tile_decode(p, tile_ok ? t : pix_tiles - 1, tiles_img, b, ty, tx);
        const int y = ty * kTileH + yi, x = tx * kTileW + xi;
        const bool valid = tile_ok && (x < p.Wo) && (y < p.Ho);
        if (do_stats && tile_ok) {
          const int grp = p.stats_bg > 0 ? b / p.stats_bg : 0;
          if (grp != cur_grp) {
            if (cur_grp >= 0) flush_stats(cur_grp);
            cur_grp = grp;
          }
        }
        const long long row_off = p.out_off[g] + (long long)b * p.sb + (long long)y * p.sy + (long long)x * p.sx + n0;
        __nv_bfloat16* dst_row = reinterpret_cast<__nv_bfloat16*>(p.out) + row_off;
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t((acc * MT + m) * BN);
#pragma unroll
        for (int ci = 0; ci < Cfg::kCPW; ++ci) {
          constexpr int kCB = Cfg::kCB;
          const int c0 = (half + 2 * ci) * kCB;
          if (c0 < BN) {
            uint32_t v[32];
            if (kCB == 32 && BN - c0 >= 32) {
              tmem_ld_32x32(taddr + c0, v);
            } else {  // BN = 16 / 80: the last column block is 16 wide; BN = 32: two 16-column blocks
              uint32_t w[16];
              tmem_ld_32x16(taddr + c0, w);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = w[j];
#pragma unroll
              for (int j = 16; j < 32; ++j) v[j] = 0u;
            }
            tmem_ld_wait();
            if (p.bias9 != nullptr && valid) {
              const int ry = y == 0 ? 0 : (y == p.Ho - 1 ? 2 : 1), rx = x == 0 ? 0 : (x == p.Wo - 1 ? 2 : 1);
              const float4* bp =
                  reinterpret_cast<const float4*>(p.bias9 + ((long long)(b * 9 + ry * 3 + rx) * p.N + n0 + c0));
#pragma unroll
              for (int j = 0; j < kCB / 4; ++j) {
                if (c0 + 4 * j < BN) {
                  const float4 t = __ldg(bp + j);
                  v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + t.x);
                  v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + t.y);
                  v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + t.z);
                  v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + t.w);
                }
              }
            }
            if (p.epi_mode != 0 && valid) {
              const uint4* sp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.epi_src) + row_off + c0);
#pragma unroll
              for (int j = 0; j < kCB / 8; ++j) {
                if (c0 + 8 * j < BN) {
                  const uint4 sv = __ldg(sp + j);
                  const uint32_t w4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float lo = bf16_lo(w4[k]), hi = bf16_hi(w4[k]);
                    float a0 = __uint_as_float(v[8 * j + 2 * k]), a1 = __uint_as_float(v[8 * j + 2 * k + 1]);
                    if (p.epi_mode == 1) { a0 += lo; a1 += hi; }
                    else { a0 = lo > 0.f ? a0 : 0.2f * a0; a1 = hi > 0.f ? a1 : 0.2f * a1; }
                    v[8 * j + 2 * k] = __float_as_uint(a0);
                    v[8 * j + 2 * k + 1] = __float_as_uint(a1);
                  }
                }
              }
            }
            if (p.act == 2) {
#pragma unroll
              for (int j = 0; j < kCB; ++j) {
                const float f = __uint_as_float(v[j]);
                v[j] = __float_as_uint(f > 0.f ? f : 0.2f * f);
              }
            }
            if (do_stats) {
              // statistics of what BatchNorm will read back: the bf16-rounded outputs of the valid rows
              if constexpr (Cfg::kRegStats) {
                if (valid) {
#pragma unroll
                  for (int j = 0; j < kCB; ++j) {
                    const float rr = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[j])));
                    acc_s[ci][j] += rr;
                    acc_q[ci][j] = fmaf(rr, rr, acc_q[ci][j]);
                  }
                }
              } else {
                float a[32], qq[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float rr = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[j])));
                  a[j] = (valid && (c0 + j < BN)) ? rr : 0.f;
                  qq[j] = a[j] * a[j];
                }
                const float cs = warp_transpose_sum(a, lane);
                const float cq = warp_transpose_sum(qq, lane);
                if (c0 + lane < BN) {   // this warp is the only writer of quarter q's slots of its column blocks
                  s_stats[q * 2 * Cfg::kSN + c0 + lane] += cs;
                  s_stats[q * 2 * Cfg::kSN + Cfg::kSN + c0 + lane] += cq;
                }
              }
            }
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(dst_row + c0);
              const int nvec = (BN - c0 >= kCB) ? kCB / 8 : (BN - c0) / 8;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (j < nvec) {
                  uint4 o;
                  o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                  o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                  o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                  o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                  dst[j] = o;
                }
              }
            }
          }
        }
      }
      // this warp is done reading accumulator buffer `acc`: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.t_empty[acc]);
      acc ^= 1;
      if (acc == 0) tph ^= 1;
    }
    if (tm_on && et == 0) {
      p.dbg_out[blockIdx.x * 16 + 8] = clock64() - tstart;
      p.dbg_out[blockIdx.x * 16 + 9] = w_tf;
      p.dbg_out[blockIdx.x * 16 + 10] = my_units;
    }
    if (do_stats && cur_grp >= 0) flush_stats(cur_grp);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace sg2
