/*
 * Diagnostics and hardware probes (tools/probe_*.py, tools/tile_waits.py) — NOT part of the product library.
 * Built only by `python -m sg2b200.build --probes` into libsg2b200_probes.so (all product sources compiled with
 * -DSG2_BUILD_PROBES plus csrc/probe.cu); libsg2b200.so neither contains nor exports these symbols.
 */
#ifndef SG2B200_PROBES_H
#define SG2B200_PROBES_H
#include "sg2b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* EXPERIMENT (not part of the product path): conv3x3 fprop with one halo tile per channel chunk and shifted UMMA
 * descriptors per tap; used by tools/probe_halo.py to validate the addressing scheme on hardware. */
int sg2_probe_halo_fprop(const void* x, const void* wpk, void* y, int B, int H, int W, int Cin, int Cout, int pitch,
                         int bo_mode, void* stream);

/* DIAGNOSTICS: tcgen05.mma issue/latency microbenchmark (tools/probe_mma.py); out[0] = issue cycles, out[1] = cycles
 * until the commit arrives, for `iters` x `kpi` MMAs of shape 128 x N x 16. */
int sg2_probe_mma(long long* out, int N, int row_bytes, int shift_rows, int pitch, int nacc, int iters, int kpi,
                  int mn_major, int style, int nblocks, void* stream);

/* DIAGNOSTICS: wait-cycle counters of the last tile-conv launch run with SG2_TILE_DBG & 64 (16 int64 per CTA). */
int sg2_tile_dbg_read(long long* host_out, int n);
/* DIAGNOSTICS: TMA streaming bandwidth of {C, bw, bh, 1} boxes over an NHWC bf16 tensor (tools/probe_tma.py). */
int sg2_probe_tma(const void* x, long long* out, int B, int H, int W, int C, int bw, int bh, int depth, int nblocks,
                  int swizzle_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
