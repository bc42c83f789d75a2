// Diagnostics (tools/probe_mma.py): tcgen05.mma issue/latency microbenchmark. Not on the product path.
// One CTA per SM; one thread issues `iters` MMAs (M=128, N, K=16, bf16) whose operands already sit in shared memory
// and measures cycles until the commit arrives. Variants: descriptor start aligned to the swizzle atom or shifted by
// rows (halo-window addressing), SBO 8 rows or 10 rows, accumulators rotated over `nacc` TMEM tiles.
#include "../../include/sg2b200_probes.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sg2 {

__global__ void __launch_bounds__(128) mma_probe_kernel(long long* out, int N, int row_bytes, int shift_rows, int pitch,
                                                        int nacc, int iters, int kpi, int mn_major, int style) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, mn_major, mn_major);
    const uint32_t swc = swizzle_code(row_bytes);
    const uint32_t a0 = smem_u32(smem) + uint32_t(shift_rows * row_bytes);
    const uint32_t b0 = smem_u32(smem + 64 * 1024);
    const uint64_t adesc = mn_major ? make_smem_desc(a0, 64 * row_bytes, uint32_t(pitch * row_bytes), swc)
                                    : make_smem_desc(a0, 16, uint32_t(pitch * row_bytes), swc);
    const uint64_t bdesc = mn_major ? make_smem_desc(b0, 64 * row_bytes, 8 * row_bytes, swc)
                                    : make_smem_desc(b0, 16, 8 * row_bytes, swc);
    for (int rep = 0; rep < 2; ++rep) {
      long long t0 = 0, t1 = 0;
      if (style == 0) {            // one divergent thread, runtime loops (what the conv kernels did)
        if (lane == 0) {
          t0 = clock64();
          for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem_base + uint32_t((i % nacc) * N);
            for (int k = 0; k < kpi; ++k) {
              const uint64_t adv = mn_major ? uint64_t((k * 16 * row_bytes) >> 4) : uint64_t(k * 2);
              umma_f16(d, adesc + adv, bdesc + adv, idesc, 1u);
            }
          }
          umma_commit(&bar);
          t1 = clock64();
        }
      } else if (style == 1) {     // one divergent thread, constant operands, unrolled x8
        if (lane == 0) {
          t0 = clock64();
          for (int i = 0; i < iters * kpi; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) umma_f16(tmem_base, adesc, bdesc, idesc, 1u);
          }
          umma_commit(&bar);
          t1 = clock64();
        }
      } else if (style == 2) {     // convergent warp, elected lane issues; descriptors advance by immediates, unrolled x4
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, 1u);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        t1 = clock64();
      } else {                     // convergent warp, per-iteration descriptor from a small table (tap windows), unrolled x4
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
          const uint32_t tap = uint32_t(i % 9);
          const uint64_t ad = adesc + uint64_t(((tap / 3) * pitch + tap % 3) * (row_bytes >> 4));
          const uint32_t d = tmem_base + uint32_t((i & 1) * N);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d, ad + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, 1u);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        t1 = clock64();
      }
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (blockIdx.x == 0 && rep == 1 && lane == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sg2

extern "C" int sg2_probe_mma(long long* out, int N, int row_bytes, int shift_rows, int pitch, int nacc, int iters,
                             int kpi, int mn_major, int style, int nblocks, void* stream) {
  using namespace sg2;
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_probe_kernel<<<nblocks, 128, smem, (cudaStream_t)stream>>>(out, N, row_bytes, shift_rows, pitch, nacc, iters, kpi,
                                                                  mn_major, style);
  SG2_LAUNCH_OK("mma_probe");
}

// ---------------------------------------------------------------------------------------------------------------
// Diagnostics: TMA streaming bandwidth. Each CTA walks over its share of an NHWC bf16 tensor with 4-D boxes
// {C, bw, bh, 1} through a ring of `depth` shared-memory buffers; nothing consumes the data. out[0] = cycles.
namespace sg2 {
__global__ void __launch_bounds__(64) tma_probe_kernel(const __grid_constant__ CUtensorMap tm, long long* out, int box_bytes,
                                                       int depth, int tiles_x, int tiles_y, int B, int bw, int bh) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int stride = (box_bytes + 1023) & ~1023;
  const int total = tiles_x * tiles_y * B;
  if (threadIdx.x < 32) {
    const long long t0 = clock64();
    int issued = 0, done = 0;
    for (int t = blockIdx.x; t < total || done < issued; t += gridDim.x) {
      if (t < total) {
        if (issued - done == depth) {
          mbar_wait(&full[done % depth], (done / depth) & 1);
          ++done;
        }
        if (elect_one()) {
          const int s = issued % depth;
          const int b = t / (tiles_x * tiles_y), r = t % (tiles_x * tiles_y);
          mbar_expect_tx(&full[s], box_bytes);
          tma_load_4d(&tm, &full[s], smem + s * stride, 0, (r % tiles_x) * bw, (r / tiles_x) * bh, b);
        }
        __syncwarp();
        ++issued;
      } else {
        mbar_wait(&full[done % depth], (done / depth) & 1);
        ++done;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = clock64() - t0;
  }
}
}  // namespace sg2

#include <cuda.h>
extern "C" int sg2_probe_tma(const void* x, long long* out, int B, int H, int W, int C, int bw, int bh, int depth,
                             int nblocks, int swizzle_bytes, void* stream) {
  using namespace sg2;
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp)
    SG2_FAIL(SG2_EDRIVER, "no cuTensorMapEncodeTiled");
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                               : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                      : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                                                             : CU_TENSOR_MAP_SWIZZLE_NONE));
  CUresult r = reinterpret_cast<EncodeTiledFn>(fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims,
                                                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) SG2_FAIL(SG2_EDRIVER, "encode failed %d", (int)r);
  const int box_bytes = C * 2 * bw * bh;
  const int smem = depth * ((box_bytes + 1023) & ~1023) + 1024;
  cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  tma_probe_kernel<<<nblocks, 64, smem, (cudaStream_t)stream>>>(tm, out, box_bytes, depth, W / bw, H / bh, B, bw, bh);
  SG2_LAUNCH_OK("tma_probe");
}
