// Gather-form implicit GEMM on CTA PAIRS (tcgen05 cta_group::2), sm_100a.
//
// The main loop of igemm_fprop_kernel<256, 64> is bound by the bytes TMA can bring into one SM (~42 B/cycle): a
// 128 x 256 tile needs 16 KB of activations + 32 KB of weights per 64-deep K block against 512 tensor cycles, 96 B/cycle.
// Here two CTAs with adjacent pixel tiles (cluster ranks 2i, 2i+1: one TPC) execute ONE MMA of M = 256: each CTA stages
// its own 128 pixel rows and HALF of the weight tile (128 of the 256 output channels), 32 KB per K block and SM, and
// receives the 128 x 256 accumulator of its own rows in its own TMEM. Only the even CTA issues tcgen05.mma; the TMA loads
// of both CTAs complete on the leader's `full` barrier; the MMA commit is multicast to the `empty` barriers of both.
//
// Split-K: the cluster is (2, 1, splitk); after the main loop every CTA parks its fp32 partial tile in its own shared
// memory and CTA (pair rank r, split s) reduces column units [s*U/S, (s+1)*U/S) of ITS pixel tile over the shared memory of
// the CTAs (r, 0..S-1) in split order (deterministic), applies the epilogue (dgrad operand, LeakyReLU, bf16 rounding,
// BatchNorm statistics) and stores bf16 — the epilogue of igemm_fprop_cluster_kernel. splitk = 1 takes the same path.
#pragma once
#include "igemm.cuh"

namespace sg2 {

template <int BN, int BK>
struct PairCfg {
  static constexpr int kSw = BK * 2;
  static constexpr int kABytes = kBlockM * BK * 2;
  static constexpr int kBHalfBytes = (BN / 2) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBHalfBytes;
  static constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : (BN <= 256 ? 256 : 512)));
  static constexpr int kPartRow = BN + kPartPad;
  static constexpr size_t kPartBytes = size_t(kBlockM) * kPartRow * sizeof(float);
  static constexpr int kMaxStages = 6;
  static constexpr int kDirectStages = 3;   // 96 KB ring: two CTAs per SM
  static size_t smem_bytes(int stages) {
    const size_t ring = size_t(stages) * kStageBytes;
    return (ring > kPartBytes ? ring : kPartBytes) + 1024 + 512 + 8192 + 256;   // align slack, barriers, statistics partials
  }
  static size_t smem_bytes_direct(int stages) { return size_t(stages) * kStageBytes + 1024 + 512; }
};

// kDirect (splitk == 1): the epilogue warps take their rows straight from TMEM to global memory like igemm_fprop_kernel and
// the operand ring is the only large shared-memory user (3 stages = 96 KB: two CTAs of different pairs per SM, one's
// epilogue overlaps the other's main loop). Otherwise the cluster reduction below (one CTA per SM).
template <int BN, int BK, bool kDirect>
__global__ void __launch_bounds__(kNumThreads, kDirect ? 2 : 1) igemm_fprop_pair_kernel(const __grid_constant__ FpropParams p) {
  using Cfg = PairCfg<BN, BK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const size_t ring = size_t(S) * Cfg::kStageBytes;
  const size_t data_bytes = kDirect ? ring : (ring > Cfg::kPartBytes ? ring : Cfg::kPartBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + data_bytes);
  uint64_t* empty = full + S;
  uint64_t* tmem_full = empty + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_red = reinterpret_cast<float*>(smem + data_bytes + 512);
  float* part = reinterpret_cast<float*>(smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t pr = crank & 1u;          // rank inside the CTA pair
  const uint32_t leader = crank & ~1u;     // cluster rank of the pair's MMA-issuing CTA
  int t = blockIdx.x;                      // pixel tile of this CTA; a tile past the last decodes to an image >= B:
  const int tx = t % p.tiles_x;            // zero-filled boxes, nothing stored (its half of the weights is still loaded)
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int tb = t / p.tiles_y;
  const int x0 = tx * p.tw, y0 = ty * p.th, b0 = tb * p.nb;
  const int n0 = blockIdx.y * BN;
  const int g = blockIdx.z / p.splitk;
  const int split = blockIdx.z % p.splitk;
  const int KB = p.ntaps * p.kchunks;
  const int kb_begin = (int)((long long)KB * split / p.splitk);
  const int kb_end = (int)((long long)KB * (split + 1) / p.splitk);
  const int nkb = kb_end - kb_begin;
  if (threadIdx.x == 0 && (pr != (blockIdx.x & 1u) || (crank >> 1) != (uint32_t)split)) {
    printf("sg2b200: pair kernel: cluster rank %u does not match block (%d,%d,%d)\n", crank, blockIdx.x, blockIdx.y, blockIdx.z);
    __trap();
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmA[0]);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  // both CTAs of the pair are running before either issues the two-SM TMEM allocation: the CTAs of a cluster are
  // co-scheduled but, next to other streams' kernels, start at different times
  cluster_sync_all();
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();   // the peer's barriers are initialised before any remote complete_tx / multicast commit reaches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint16_t pair_mask = (uint16_t)(3u << leader);

  if (warp == 0) {
    const uint32_t full_leader = dsmem_addr(smem_u32(&full[0]), leader);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it) {
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one()) {
        if (pr == 0) mbar_expect_tx(&full[s], 2 * Cfg::kStageBytes);   // the bytes of both CTAs land on the leader's barrier
        const int kb = kb_begin + it;
        const int tap = kb / p.kchunks;
        const int ch = kb - tap * p.kchunks;
        const TapF tp = p.taps[g][tap];
        uint8_t* sa = smem + size_t(s) * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        const uint32_t bar = full_leader + uint32_t(s) * 8u;
        tma_load_4d_pair(&p.tmA[tp.map], bar, sa, ch * BK, x0 + tp.dx, y0 + tp.dy, b0);
        tma_load_2d_pair(&p.tmB, bar, sb, kb * BK, n0 + g * p.N + int(pr) * (BN / 2));
      }
      __syncwarp();
      if (++s == S) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1 && pr == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BN, 0, 0);
    constexpr uint32_t swc = swizzle_code(Cfg::kSw);
    const uint64_t adesc0 = make_smem_desc(smem_u32(smem), 16, 8 * Cfg::kSw, swc);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(smem) + Cfg::kABytes, 16, 8 * Cfg::kSw, swc);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = uint64_t((uint32_t(s) * uint32_t(Cfg::kStageBytes)) >> 4);
        const uint64_t adesc = adesc0 + so, bdesc = bdesc0 + so;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_f16_pair(tmem_base, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit_pair(&empty[s], pair_mask);
      }
      __syncwarp();
      if (++s == S) {
        s = 0;
        ph ^= 1;
      }
    }
    if (elect_one()) umma_commit_pair(tmem_full, pair_mask);
    __syncwarp();
  }
  if constexpr (kDirect) {
    __shared__ float s_stats[4 * 2 * BN];   // [TMEM lane quarter][2][BN]: one writer warp per slot
    if (warp >= 2) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int xi = row % p.tw, yi = (row / p.tw) % p.th, bi = row / (p.tw * p.th);
      const int x = x0 + xi, y = y0 + yi, b = b0 + bi;
      const bool valid = (x < p.Wo) && (y < p.Ho) && (b < p.B);
      const bool do_stats = p.stats != nullptr;
      const int et = threadIdx.x - 64;
      const long long off = p.out_off[g] + (long long)b * p.sb + (long long)y * p.sy + (long long)x * p.sx + n0;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        if (nkb <= 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (p.epi_mode != 0 && valid) {
          const uint4* sp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.epi_src) + off + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
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
        if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float f = __uint_as_float(v[j]);
            v[j] = __float_as_uint(f > 0.f ? f : 0.2f * f);
          }
        }
        if (do_stats) {
          float a[32], qq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float r = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[j])));
            a[j] = valid ? r : 0.f;
            qq[j] = a[j] * a[j];
          }
          const float cs = warp_transpose_sum(a, lane);
          const float cq = warp_transpose_sum(qq, lane);
          s_stats[q * 2 * BN + c0 + lane] = cs;
          s_stats[q * 2 * BN + BN + c0 + lane] = cq;
        }
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
      if (do_stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (b0 < p.B) {
          double* st = p.stats + (p.stats_bg > 0 ? (long long)(b0 / p.stats_bg) * 2 * p.N : 0);
          for (int i = et; i < BN; i += 128) {
            float cs = 0.f, cq = 0.f;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              cs += s_stats[qq * 2 * BN + i];
              cq += s_stats[qq * 2 * BN + BN + i];
            }
            atomicAdd(&st[n0 + i], (double)cs);
            atomicAdd(&st[p.N + n0 + i], (double)cq);
          }
        }
      }
      tc_fence_before();
    }
    cluster_sync_all();   // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
    if (warp == 1) {
      tc_fence_after();
      tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    }
    return;
  }
  if (warp >= 2) {
    mbar_wait(tmem_full, 0);   // every MMA of the pair has completed: both operand rings are idle from here on
    tc_fence_after();
    // ---- phase 1: my partial tile TMEM -> my shared memory (over the idle operand ring)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
    float* prow = part + (size_t)row * Cfg::kPartRow;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                               __uint_as_float(v[j + 3]));
        if (nkb <= 0) o = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(prow + c0 + j) = o;
      }
    }
    tc_fence_before();
  }
  cluster_sync_all();   // all partial tiles of the cluster are in shared memory

  if (warp >= 2) {
    // ---- phase 2: this CTA owns column units [u0, u1) of 8 channels of its pixel tile; thread = (row slot, unit)
    constexpr int U = BN / 8;
    const int Sx = p.splitk;
    const int u0 = split * U / Sx, u1 = (split + 1) * U / Sx;
    const int nu = u1 - u0;
    const int et = threadIdx.x - 64;
    const bool do_stats = p.stats != nullptr;
    float cs[8], cq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = cq[j] = 0.f;
    int slots = 0, slot = 0, uu = 0;
    if (nu > 0) {
      slots = 128 / nu;
      if (slots < 1) slots = 1;
      slot = et / nu;
      uu = et - slot * nu;
    }
    // nu can exceed 128 threads' worth only if BN / 8 > 128 (never); with nu <= 32 every unit has >= 4 row slots
    const bool active = nu > 0 && slot < slots;
    if (active) {
      const int col = (u0 + uu) * 8;
      const uint32_t my_off = smem_u32(part) + uint32_t(col) * 4u;
      for (int r = slot; r < kBlockM; r += slots) {
        const int xi = r % p.tw, yi = (r / p.tw) % p.th, bi = r / (p.tw * p.th);
        const int x = x0 + xi, y = y0 + yi, b = b0 + bi;
        if (!((x < p.Wo) && (y < p.Ho) && (b < p.B))) continue;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 0.f;
        const uint32_t roff = my_off + uint32_t(r) * uint32_t(Cfg::kPartRow * 4);
        for (int s2 = 0; s2 < Sx; ++s2) {
          const uint32_t ra = dsmem_addr(roff, pr + 2u * (uint32_t)s2);
          const float4 lo = dsmem_ld_f4(ra), hi = dsmem_ld_f4(ra + 16);
          a[0] += lo.x; a[1] += lo.y; a[2] += lo.z; a[3] += lo.w;
          a[4] += hi.x; a[5] += hi.y; a[6] += hi.z; a[7] += hi.w;
        }
        const long long off = p.out_off[g] + (long long)b * p.sb + (long long)y * p.sy + (long long)x * p.sx + n0 + col;
        if (p.epi_mode != 0) {
          const uint4 sv = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.epi_src) + off));
          const uint32_t w4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float elo = bf16_lo(w4[k]), ehi = bf16_hi(w4[k]);
            if (p.epi_mode == 1) { a[2 * k] += elo; a[2 * k + 1] += ehi; }
            else { a[2 * k] = elo > 0.f ? a[2 * k] : 0.2f * a[2 * k]; a[2 * k + 1] = ehi > 0.f ? a[2 * k + 1] : 0.2f * a[2 * k + 1]; }
          }
        }
        if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = a[j] > 0.f ? a[j] : 0.2f * a[j];
        }
        uint4 o;
        o.x = pack_bf16x2(a[0], a[1]); o.y = pack_bf16x2(a[2], a[3]);
        o.z = pack_bf16x2(a[4], a[5]); o.w = pack_bf16x2(a[6], a[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = o;
        if (do_stats) {
          const uint32_t w4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float f0 = bf16_lo(w4[k]), f1 = bf16_hi(w4[k]);
            cs[2 * k] += f0; cq[2 * k] = fmaf(f0, f0, cq[2 * k]);
            cs[2 * k + 1] += f1; cq[2 * k + 1] = fmaf(f1, f1, cq[2 * k + 1]);
          }
        }
      }
    }
    if (do_stats) {
      // ordered combine of the row slots: s_red[et][16]; one thread per (unit, statistic) sums the slots in order
      float* mine = s_red + et * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { mine[j] = cs[j]; mine[8 + j] = cq[j]; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int task = et; task < nu * 16; task += 128) {
        const int u2 = task / 16, k = task % 16;
        float tsum = 0.f;
        for (int sl = 0; sl < slots; ++sl) tsum += s_red[(sl * nu + u2) * 16 + k];
        double* st = p.stats + (p.stats_bg > 0 ? (long long)(b0 / p.stats_bg) * 2 * p.N : 0);
        const int c = n0 + (u0 + u2) * 8 + (k & 7);
        if (b0 < p.B) atomicAdd(&st[(k < 8 ? 0 : p.N) + c], (double)tsum);
      }
    }
  }
  cluster_sync_all();   // nobody leaves (and frees its shared memory / TMEM) while a peer still reads it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace sg2

namespace sg2 {

// ------------------------------------------------------------------------------------------------ wgrad on CTA pairs
// igemm_wgrad_kernel<256, 64, 64> with M = 256: the pair covers 256 output channels x 256 input channels of one tap; each
// CTA stages the dy columns of ITS 128 output channels and HALF of the activation tile (128 of the 256 input channels):
// 32 KB per 64-pixel K block and SM instead of 48 KB. blockIdx.x = n_tile * m_tiles + m_tile (m_tiles even), so cluster
// ranks 0 / 1 are adjacent output-channel tiles. Needs Cout % 256 == 0 and Cin % 256 == 0.
struct WgradPairCfg {
  static constexpr int kCW = 64;
  static constexpr int kBN = 256;
  static constexpr int kChunkBytes = kWgradBKP * kCW * 2;   // 8 KB: 64 pixels x 64 channels
  static constexpr int kABytes = 2 * kChunkBytes;            // 128 output channels
  static constexpr int kBHalfBytes = 2 * kChunkBytes;        // 128 of the 256 input channels
  static constexpr int kStageBytes = kABytes + kBHalfBytes;
  static constexpr int kStages = 3;                          // 96 KB: two CTAs per SM
  static size_t smem_bytes(int stages) { return size_t(stages) * kStageBytes + 1024 + 256; }
};

__global__ void __launch_bounds__(kNumThreads, 2) igemm_wgrad_pair_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = WgradPairCfg;
  constexpr int BN = Cfg::kBN, CW = Cfg::kCW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(S) * Cfg::kStageBytes);
  uint64_t* empty = full + S;
  uint64_t* tmem_full = empty + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t pr = crank & 1u;
  const uint32_t leader = crank & ~1u;
  const int m_tiles = p.Cout / kBlockM;
  const int m0 = (blockIdx.x % m_tiles) * kBlockM;
  const int n0 = (blockIdx.x / m_tiles) * BN;
  const int job = blockIdx.y;
  const int split = blockIdx.z;
  const int PT = p.tiles_x * p.tiles_y * p.tiles_b;  // pixel tiles = K blocks
  const int kb_begin = (int)((long long)PT * split / p.splitk);
  const int kb_end = (int)((long long)PT * (split + 1) / p.splitk);
  const int nkb = kb_end - kb_begin;
  if (threadIdx.x == 0 && pr != (blockIdx.x & 1u)) {
    printf("sg2b200: wgrad pair kernel: cluster rank %u does not match block %d\n", crank, blockIdx.x);
    __trap();
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  cluster_sync_all();   // see igemm_fprop_pair_kernel: the peer is running before the two-SM allocation is issued
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, BN);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint16_t pair_mask = (uint16_t)(3u << leader);

  if (warp == 0) {
    const JobW jb = p.jobs[job];
    const uint32_t full_leader = dsmem_addr(smem_u32(&full[0]), leader);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it) {
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one()) {
        if (pr == 0) mbar_expect_tx(&full[s], 2 * Cfg::kStageBytes);
        int t = kb_begin + it;
        const int tx = t % p.tiles_x;
        t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int tb = t / p.tiles_y;
        const int x0 = tx * p.tw, y0 = ty * p.th, b0 = tb * p.nb;
        uint8_t* sa = smem + size_t(s) * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        const uint32_t bar = full_leader + uint32_t(s) * 8u;
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_4d_pair(&p.tmA[jb.amap], bar, sa + c * Cfg::kChunkBytes, m0 + c * CW, x0, y0, b0);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_4d_pair(&p.tmB[jb.bmap], bar, sb + c * Cfg::kChunkBytes, n0 + int(pr) * (BN / 2) + c * CW, x0 + jb.dx,
                           y0 + jb.dy, b0);
      }
      __syncwarp();
      if (++s == S) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1 && pr == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BN, 1, 1);
    constexpr uint32_t sw = swizzle_code(CW * 2);
    // MN-major: LBO = bytes between consecutive channel chunks, SBO = bytes between 8-pixel groups.
    const uint64_t adesc0 = make_smem_desc(smem_u32(smem), Cfg::kChunkBytes, 8 * CW * 2, sw);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(smem) + Cfg::kABytes, Cfg::kChunkBytes, 8 * CW * 2, sw);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = uint64_t((uint32_t(s) * uint32_t(Cfg::kStageBytes)) >> 4);
        const uint64_t adesc = adesc0 + so, bdesc = bdesc0 + so;
#pragma unroll
        for (int k = 0; k < kWgradBKP / 16; ++k) {
          const uint64_t ko = uint64_t((k * 16 * CW * 2) >> 4);
          umma_f16_pair(tmem_base, adesc + ko, bdesc + ko, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit_pair(&empty[s], pair_mask);
      }
      __syncwarp();
      if (++s == S) {
        s = 0;
        ph ^= 1;
      }
    }
    if (elect_one()) umma_commit_pair(tmem_full, pair_mask);
    __syncwarp();
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool slabs = p.partials != nullptr;
    float* rowp = (slabs ? p.partials + (long long)split * p.slab : p.dw) + ((long long)m * p.njobs + job) * p.Cin + n0;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + c0, v);
      tmem_ld_wait();
      if (slabs) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                 __uint_as_float(v[j + 3]));
          if (nkb <= 0) o = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(rowp + c0 + j) = o;
        }
      } else if (nkb > 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(rowp + c0 + j),
                       "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                       "f"(__uint_as_float(v[j + 3]))
                       : "memory");
        }
      }
    }
    tc_fence_before();
  }
  cluster_sync_all();   // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, BN);
  }
}

}  // namespace sg2
