// tcgen05 implicit-GEMM Conv1d for the UNet (sm_100a).
//
// Replaces every F.conv1d of Unet1D.forward (reference srcs/modules/unet.py:80,61,65,201,204,232,307,369)
// on channels-last h16 activations with fp32 accumulation in TMEM.
//
//   D[m, n] (TMEM, fp32; lane = output channel m, column = position n)
//     = sum over groups / 64-channel chunks / taps of
//       A = W[m0:m0+128, kofs_tap + 64c : +64]            (h16, K-major, TMA 2D, SWIZZLE_128B)
//       B = X[b, l0+shift+row_off_tap : +N, ch0+64c : +64] (h16, K-major, TMA 3D, SWIZZLE_128B, OOB rows -> 0 = zero padding)
//
// The activation tile of a chunk is loaded ONCE with its halo rows; every tap reads it through a shared-memory
// descriptor whose start address is advanced by row_off*128 B (the 128B-swizzle XOR is a function of the absolute
// shared-memory address, so a row-shifted start stays consistent with what TMA wrote).
//
// Persistent, warp-specialised CTAs (one per SM):
//   warp 0    TMA producer (two rings: weights 16 KB slots, activations)
//   warp 1    TMEM allocator + MMA issuer (tcgen05.mma cta_group::1 kind::f16, M=128, N<=256, K=16)
//   warps 2-9 epilogue: two warpgroups (each covers the four TMEM lane quadrants) that take alternate store chunks of a tile:
//             tcgen05.ld -> +bias -> GroupNorm partial sums -> h16 -> smem staging -> TMA store
// Two TMEM accumulator stages, so the epilogue of tile i overlaps the main loop of tile i+1.
// Short clips (L < 128) are packed several per tile (one TMA box per clip, per-clip zero padding kept).
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int kThreads = 320;
constexpr int kEpiGroups = 2;     // epilogue warpgroups; GroupNorm partials carry one slot set per group
constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;  // 16 KB
constexpr int kMaxStages = 6;
constexpr int kStageTaps = 3;    // weight tiles per pipeline stage
constexpr size_t kSmemLimit = 232448 - 1024;     // 227 KB minus the static barriers
constexpr size_t kSmemLimit2 = 112 * 1024;        // per CTA when two CTAs share an SM (228 KB per SM, 1 KB reserved per CTA)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp.  The compiler knows an elect.sync region is single-threaded and emits the uniform-datapath
// instructions (UTCHMMA / UTMALDG / UTMASTG) directly; under `lane == 0` it wraps each of them in an ELECT + BRA.U.ANY
// waterfall loop that costs ~50 cycles per instruction on the latency-exposed single issuing thread.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Parity wait with a watchdog: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins == 64) t0 = clock64();
    if (spins > 64 && (spins & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long& acc, bool on) {
  if (!on) { mbar_wait(bar, parity); return; }
  const long long t = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// half a weight tile to the same shared-memory offset of BOTH CTAs of the pair; each CTA's own barrier (same offset) gets the bytes
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(int wg) {      // constant ids: ptxas then reserves 4 named barriers per CTA, not all 16
  if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same, arriving on the barrier at this offset in both CTAs of the pair (a stage is refilled by multicast writes into both
// CTAs' shared memory, so it is free only when both CTAs' MMAs have read it)
__device__ __forceinline__ void tc_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 | LBO(=1, ignored for swizzled K-major)<<16 | SBO(=1024 B: 8 rows x 128 B)>>4 <<32 | version 1 <<46 | layout 2 <<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// asynchronous: the registers are valid only after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct TileCoord { int m0, b0, l0, pt; };
// Tile schedule.  Single CTAs: CTA x takes tiles x, x + grid, ... of mt-fastest order.  Pairs: pair q = blockIdx.x / 2 takes "pair
// tiles" q, q + npairs, ...; pair tile u = (mt = u % MT, j = u / MT) covers position tiles 2j (rank 0) and 2j + 1 (rank 1); a rank
// whose position tile does not exist (odd count) is a dummy: it still loads and multicasts its weight halves and releases stages.
// PAIR is a template parameter: the single-CTA instantiation carries none of the pair logic.
template <int PAIR> __device__ __forceinline__ int tsched_first() { return PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x; }
template <int PAIR> __device__ __forceinline__ int tsched_step() { return PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x; }
template <int PAIR> __device__ __forceinline__ int tsched_count(const TcConvParams& p) { return PAIR ? p.MT * ((p.n_ntiles + 1) >> 1) : p.MT * p.n_ntiles; }
template <int PAIR>
__device__ __forceinline__ TileCoord decode_tile(const TcConvParams& p, int t, uint32_t rank, bool* dummy) {
  TileCoord c;
  const int mt = t % p.MT;
  int nt = t / p.MT;
  *dummy = false;
  if (PAIR) {
    nt = 2 * nt + (int)rank;
    if (nt >= p.n_ntiles) { *dummy = true; nt = p.n_ntiles - 1; }
  }
  c.m0 = mt * TC_BM;
  if (p.NCLIP == 1) { c.pt = nt % p.n_ptiles; c.b0 = nt / p.n_ptiles; c.l0 = c.pt * p.NT; }
  else { c.pt = 0; c.b0 = nt * p.NCLIP; c.l0 = 0; }
  return c;
}

// MINB = 2: the same code compiled to <= 102 registers so that two CTAs of a small-footprint launch (<= 112 KB shared memory, <= 256
// TMEM columns) share an SM: one CTA's prologue / epilogue then overlaps the other's main loop.
// PROF = 1 (LADIFF_TC_PROF): per-role wait-cycle counters and the LADIFF_TC_DBG ablations; the production instantiation has neither.
template <int MINB, int PAIR, int PROF>
__global__ void __launch_bounds__(kThreads, MINB) tc_conv_kernel(const __grid_constant__ TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_off = (uint32_t)p.a_cap * A_BYTES;                 // stage = [a_cap weight tiles][activation tile]
  const uint32_t stage0 = smem_base + (uint32_t)p.S * (uint32_t)p.stage_bytes;   // epilogue staging
  uint32_t acc_stride = 32;
  while ((int)acc_stride < p.NMMA) acc_stride <<= 1;
  const uint32_t tmem_cols = 2 * acc_stride;
  const int total_tiles = tsched_count<PAIR>(p), t_first = tsched_first<PAIR>(), t_step = tsched_step<PAIR>();
  uint32_t rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool prof = PROF && p.prof != nullptr;
  const int dbg = PROF ? p.dbg : 0;
  const long long t_begin = prof ? clock64() : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], PAIR ? 2 : 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4 * kEpiGroups); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmX) : "memory");
    if (!p.direct) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmY) : "memory");
    if (p.split_m) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmY2) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR)            // the peer's barriers are initialised before any multicast write / remote arrive can reach them
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail
  pdl_wait();
  pdl_trigger();

  // The producer and the MMA issuer are single threads: every instruction of their loops is latency-exposed, so the
  // taps a tile uses are resolved into registers once per (tile, group) and the per-chunk loops touch no parameter memory.
  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer: one pipeline stage per (group, 64-channel chunk) = activation tile + its taps' weight tiles
      const uint32_t b_bytes = (uint32_t)(p.NCLIP * p.BOXROWS) * 128u;
      const uint32_t box_bytes = (uint32_t)p.BOXROWS * 128u;
      const int nclip = p.NCLIP, S = p.S;
      const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
      int st = 0, issued = 0; uint32_t ph = 0;
      long long w_empty = 0;
      constexpr bool pair = PAIR != 0;
      for (int t = t_first; t < total_tiles; t += t_step) {
        bool dummy;
        const TileCoord tc = decode_tile<PAIR>(p, t, rank, &dummy);
        for (int g = 0; g < p.ngrp; ++g) {
          const TcGroup& gr = p.grp[g];
          int n_a = 0, k0 = 0, k1 = 0, k2 = 0;
          for (int tp = 0; tp < gr.ntaps; ++tp) {
            const TcTap tap = gr.tap[tp];
            if (tc.m0 < tap.m_lo || tc.m0 >= tap.m_hi) continue;
            if (n_a == 0) k0 = tap.kofs; else if (n_a == 1) k1 = tap.kofs; else k2 = tap.kofs;
            ++n_a;
          }
          const int nchunk = gr.nchunk, ch0 = gr.ch0, row = tc.l0 + gr.shift;
          const uint32_t tx_bytes = (dummy ? 0u : b_bytes) + (uint32_t)n_a * A_BYTES;
          for (int c = 0; c < nchunk; ++c) {
            mbar_wait_t(&empty_bar[st], ph ^ 1, w_empty, prof);
            if ((dbg & 1) && issued >= S) { mbar_arrive(&full_bar[st]); if (++st == S) { st = 0; ph ^= 1; } continue; }
            ++issued;
            uint64_t* fb = &full_bar[st];
            mbar_expect_tx(fb, tx_bytes);
            const uint32_t a_dst = smem_base + (uint32_t)st * stage_bytes;
            const int kc = c * TC_BK;
            if (!dummy) {
              tma_load_3d(a_dst + b_off, &p.tmX, fb, ch0 + kc, row, tc.b0);
              for (int j = 1; j < nclip; ++j) tma_load_3d(a_dst + b_off + (uint32_t)j * box_bytes, &p.tmX, fb, ch0 + kc, row, tc.b0 + j);
            }
            if (!pair) {
              if (n_a > 0) tma_load_2d(a_dst, &p.tmW, fb, k0 + kc, tc.m0);
              if (n_a > 1) tma_load_2d(a_dst + A_BYTES, &p.tmW, fb, k1 + kc, tc.m0);
              if (n_a > 2) tma_load_2d(a_dst + 2 * A_BYTES, &p.tmW, fb, k2 + kc, tc.m0);
            } else {     // this CTA's half (rows 64 rank .. +63) of every weight tile goes to both CTAs of the pair
              const uint32_t hoff = rank * (A_BYTES / 2);
              const int mrow = tc.m0 + 64 * (int)rank;
              if (n_a > 0) tma_load_2d_mcast(a_dst + hoff, &p.tmWh, fb, k0 + kc, mrow, 3);
              if (n_a > 1) tma_load_2d_mcast(a_dst + A_BYTES + hoff, &p.tmWh, fb, k1 + kc, mrow, 3);
              if (n_a > 2) tma_load_2d_mcast(a_dst + 2 * A_BYTES + hoff, &p.tmWh, fb, k2 + kc, mrow, 3);
            }
            if (++st == S) { st = 0; ph ^= 1; }
          }
        }
      }
      if (prof) p.prof[blockIdx.x * 8 + 0] = (unsigned long long)w_empty;
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer.  Instruction descriptor (InstrDescriptor, mma_sm100_desc.hpp):
      // c_format F32 (1<<4) | a_format BF16 (1<<7) | b_format BF16 (1<<10) | K-major A,B | N>>3 <<17 | M>>4 <<24
      const uint32_t idesc = (1u << 4) | (TC_IDESC_AB_FMT << 7) | (TC_IDESC_AB_FMT << 10) | ((uint32_t)(p.NMMA >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const int S = p.S;
      const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
      const bool no_mma = (dbg & 4) != 0;
      int st = 0, tl = 0; uint32_t ph = 0;
      long long w_full = 0, w_tmem = 0;
      constexpr bool pair = PAIR != 0;
      for (int t = t_first; t < total_tiles; t += t_step, ++tl) {
        bool dummy;
        const TileCoord tc = decode_tile<PAIR>(p, t, rank, &dummy);
        const int acc = tl & 1;
        mbar_wait_t(&tmem_empty[acc], ((tl >> 1) & 1) ^ 1, w_tmem, prof);     // the epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        uint32_t accumulate = 0;
        for (int g = 0; g < p.ngrp; ++g) {
          const TcGroup& gr = p.grp[g];
          int n_a = 0;
          uint32_t r0 = 0, r1 = 0, r2 = 0;       // per tap: row offset in 16-byte descriptor units (128 B per row)
          for (int tp = 0; tp < gr.ntaps; ++tp) {
            const TcTap tap = gr.tap[tp];
            if (tc.m0 < tap.m_lo || tc.m0 >= tap.m_hi) continue;
            const uint32_t ro = (uint32_t)tap.row_off * 8u;
            if (n_a == 0) r0 = ro; else if (n_a == 1) r1 = ro; else r2 = ro;
            ++n_a;
          }
          const int nchunk = gr.nchunk;
          for (int c = 0; c < nchunk; ++c) {
            mbar_wait_t(&full_bar[st], ph, w_full, prof);
            tc_fence_after();
            const uint32_t a_base = smem_base + (uint32_t)st * stage_bytes;
            const uint64_t adesc = umma_desc(a_base);
            const uint64_t bdesc = umma_desc(a_base + b_off);
            if (!no_mma && !dummy) {
              if (n_a > 0) {
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) { umma_bf16(d_tmem, adesc + 2 * k, bdesc + r0 + 2 * k, idesc, accumulate); accumulate = 1; }
              }
              if (n_a > 1) {
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(d_tmem, adesc + (A_BYTES >> 4) + 2 * k, bdesc + r1 + 2 * k, idesc, 1u);
              }
              if (n_a > 2) {
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(d_tmem, adesc + 2 * (A_BYTES >> 4) + 2 * k, bdesc + r2 + 2 * k, idesc, 1u);
              }
            }
            if (pair) tc_commit_both(&empty_bar[st]); else tc_commit(&empty_bar[st]);     // frees the stage once these MMAs have read it
            if (++st == S) { st = 0; ph ^= 1; }
          }
        }
        tc_commit(&tmem_full[acc]);        // accumulator complete
      }
      if (prof) { p.prof[blockIdx.x * 8 + 1] = (unsigned long long)w_full; p.prof[blockIdx.x * 8 + 2] = (unsigned long long)w_tmem; }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: warp w owns TMEM lanes 32*(w&3) .. +31; warpgroup wg (warps 2-5 / 6-9) takes the store chunks
    // whose running index cc has (cc & 1) == wg, with its own pair of staging buffers, named barrier and TMA-store issuer
    const int q = warp & 3;
    const int wg = (warp - 2) >> 2;
    const int Cc = p.up_cout ? p.up_cout : (p.split_m ? p.split_m : p.Cout);
    // this warp's two staging buffers, [CR rows][32 channels] each; its elected lane issues the stores, bulk-group waits are warp-wide
    const uint32_t stage_w = stage0 + (uint32_t)((wg * 4 + q) * 2) * (uint32_t)p.CR * 64u;
    int tl = 0, cc = 0, kk = 0;
    // bias of the NEXT tile is requested one tile ahead (an exposed L2 round trip per tile otherwise)
    bool dmy0;
    float bias_next = (p.bias && t_first < total_tiles) ? __ldg(p.bias + decode_tile<PAIR>(p, t_first, rank, &dmy0).m0 + q * 32 + lane) : 0.f;
    long long w_acc = 0, w_e1 = 0, w_e2 = 0, w_e3 = 0;   // profiling: chunk entry (store drain + barrier), body, fence + barrier + store issue
    for (int t = t_first; t < total_tiles; t += t_step, ++tl) {
      bool dummy;
      const TileCoord tc = decode_tile<PAIR>(p, t, rank, &dummy);
      const int acc = tl & 1;
      const int ch = tc.m0 + q * 32 + lane;
      const float bias = bias_next;
      if (p.bias && t + t_step < total_tiles) bias_next = __ldg(p.bias + decode_tile<PAIR>(p, t + t_step, rank, &dmy0).m0 + q * 32 + lane);
      const bool second = p.split_m && tc.m0 >= p.split_m;
      if (p.res) {
        // direct epilogue with a residual: its tile ([clip region][row][128 channels] h16) is staged in shared memory by all
        // epilogue threads while the main loop of this tile is still running
        asm volatile("bar.sync 3, 256;" ::: "memory");          // the previous tile's residual has been consumed
        const int et = threadIdx.x - 64;                         // 0..255
        for (int i = et; i < p.NMMA * 16; i += 256) {
          const int rr = i >> 4, seg = i & 15, j = rr / p.NT, r = rr - j * p.NT;
          const int b = tc.b0 + j, l = tc.l0 + r;
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (b < p.B && l < p.Lout)
            v = __ldg(reinterpret_cast<const uint4*>(p.res + (long long)b * p.res_bstride + (long long)l * p.res_pitch + tc.m0) + seg);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage0 + (uint32_t)i * 16u), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
      }
      mbar_wait_t(&tmem_full[acc], (tl >> 1) & 1, w_acc, prof);
      tc_fence_after();
      const uint32_t tlane = tmem_base + (uint32_t)acc * acc_stride + ((uint32_t)(q * 32) << 16);
      for (int j = 0; j < (dummy ? 0 : p.NCLIP); ++j) {
        const int b = tc.b0 + j;
        int vr = p.Lout - tc.l0;                 // valid rows of this clip region
        vr = vr > p.NT ? p.NT : vr;
        if (b >= p.B) vr = 0;
        float s1 = 0.f, s2 = 0.f;
        for (int r0 = 0; r0 < p.NT; r0 += p.CR, ++cc) {
          if ((cc & 1) != wg || vr <= r0 || (dbg & 2)) continue;    // uniform over the warpgroup
          const int nrows = (p.NT - r0) < p.CR ? (p.NT - r0) : p.CR;
          uint32_t stg = 0;
          const long long te0 = prof ? clock64() : 0;
          if (!p.direct) {
            bulk_wait_read_1();                  // the store this warp issued two chunks ago has finished reading its buffer
            __syncwarp();
            stg = stage_w + (uint32_t)(kk & 1) * (uint32_t)p.CR * 64u + (uint32_t)lane * 2u;
          }
          const long long te1 = prof ? clock64() : 0;
          const uint32_t tcol = tlane + (uint32_t)(j * p.NT + r0);
          const bool no_ld = (dbg & 8) != 0, no_st = (dbg & 16) != 0;   // ablations (LADIFF_TC_DBG)
          auto consume = [&](const uint32_t (&r)[16], int c0) {
            if (no_st) {
#pragma unroll
              for (int i = 0; i < 16; ++i) s1 += __uint_as_float(r[i]);
            } else if (!p.direct) {
              if (r0 + c0 + 16 <= vr) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float v = __uint_as_float(r[i]) + bias;
                  s1 += v; s2 += v * v;
                  const unsigned short hv = h16_bits(f2h(v));
                  asm volatile("st.shared.u16 [%0], %1;" ::"r"(stg + (uint32_t)(c0 + i) * 64u), "h"(hv) : "memory");
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float v = __uint_as_float(r[i]) + bias;
                  if (r0 + c0 + i < vr) { s1 += v; s2 += v * v; }
                  const unsigned short hv = h16_bits(f2h(v));
                  asm volatile("st.shared.u16 [%0], %1;" ::"r"(stg + (uint32_t)(c0 + i) * 64u), "h"(hv) : "memory");
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int row = r0 + c0 + i;
                if (row < vr) {
                  const int l = tc.l0 + row;
                  float v = __uint_as_float(r[i]) + bias;
                  s1 += v; s2 += v * v;
                  if (p.res) {
                    unsigned short rv16;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(rv16) : "r"(stage0 + (uint32_t)(j * p.NT + row) * 256u + (uint32_t)(q * 32 + lane) * 2u));
                    v += h2f(h16_from_bits(rv16));
                  }
                  const long long o = (long long)b * p.out_bstride + (long long)l * p.out_pitch + ch;
                  if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = v;
                  else reinterpret_cast<h16*>(p.out)[o] = f2h(v);
                }
              }
            }
          };
          // both 16-column loads of a chunk are issued back to back and waited for ONCE: tcgen05.wait::ld costs a few hundred
          // cycles while the tensor pipe is busy, whatever is outstanding
          uint32_t ra[16], rb[16];
          const bool has_b = nrows > 16;
          if (!no_ld) { tmem_ld16_async(tcol, ra); if (has_b) tmem_ld16_async(tcol + 16u, rb); }
          tmem_ld_wait();
          consume(ra, 0);
          if (has_b) consume(rb, 16);
          const long long te2 = prof ? clock64() : 0;
          if (!p.direct) {
            fence_async_smem();
            __syncwarp();
            if (!no_st && elect_one()) {
              const uint32_t src = stage_w + (uint32_t)(kk & 1) * (uint32_t)p.CR * 64u;
              if (second) tma_store_4d(nrows == p.CR ? &p.tmY2 : &p.tmY2r, src, tc.m0 - p.split_m + q * 32, 0, tc.l0 + r0, b);
              else tma_store_4d(nrows == p.CR ? &p.tmY : &p.tmYr, src, tc.m0 % Cc + q * 32, tc.m0 / Cc, tc.l0 + r0, b);
              bulk_commit();
            }
            __syncwarp();
            ++kk;
          }
          if (prof) { const long long te3 = clock64(); w_e1 += te1 - te0; w_e2 += te2 - te1; w_e3 += te3 - te2; }
        }
        if (p.stats && vr > 0 && !second) {      // one partial per (clip, position tile, warpgroup); zero if this group had no chunk
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          if (lane == 0)
            p.stats[(((long long)b * p.n_ptiles + tc.pt) * kEpiGroups + wg) * p.stat_slots + (tc.m0 / 32 + q)] = make_float2(s1, s2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);   // 8 warps -> accumulator stage free for the MMA issuer
    }
    // the staging buffers must outlive the stores' shared-memory reads; their global writes are covered by grid completion
    // (and by griddepcontrol.wait in the dependent kernel), so the CTA does not wait for them
    bulk_wait_read_0();
    if (prof && threadIdx.x == 64) { p.prof[blockIdx.x * 8 + 3] = (unsigned long long)w_acc; p.prof[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - t_begin);
      p.prof[blockIdx.x * 8 + 5] = (unsigned long long)w_e1; p.prof[blockIdx.x * 8 + 6] = (unsigned long long)w_e2; p.prof[blockIdx.x * 8 + 7] = (unsigned long long)w_e3; }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR)            // neither CTA exits while the peer can still multicast into its shared memory or arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ================================================================================================ positions-on-M variant
// D[m = position (128 TMEM lanes), n = output channel (NCH <= 256 columns)]:
//   A = activation tile [128 + halo rows][64 ch] (one TMA box per 64-channel chunk; a tap is a row-shifted descriptor start)
//   B = weight tile [NCH rows][64 ch] per (tap, chunk), in its own ring
// An epilogue thread owns one position: it reads 16 consecutive channels per tcgen05.ld and stores them as 32 contiguous bytes —
// no shared-memory staging, no barriers, GroupNorm partials reduced per 32-channel slot with warp shuffles.  One activation load
// serves all <= 256 output channels of the tile (the channels-on-M kernel re-reads it per 128-channel tile).
constexpr int kThreadsT = 192;
constexpr int kActSlots = 3, kActBytes = 18432;        // 136 rows x 128 B, rounded to 1 KB
constexpr int kMaxWSlots = 8;
constexpr size_t kSmemLimitT = 232448 - 3072;         // 227 KB minus this kernel's static shared memory (barriers, bias tiles)

// ---- CTA-pair (cta_group::2) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of a CTA pair: the data lands in this CTA's shared memory, the bytes are accounted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {      // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

// CG = 1: one CTA per tile.  CG = 2: a CTA pair (cta_group::2) computes 256 positions x NCH channels; CTA r loads its own 128
// positions and rows [r*NCH/2, (r+1)*NCH/2) of every weight tile, so the weight bytes landing in each SM are halved; the leader
// (rank 0) issues the M = 256 MMAs, its commits free the ring slots of both CTAs.
template <int CG>
__global__ void __launch_bounds__(kThreadsT, 1) tc_conv_t_kernel(const __grid_constant__ TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t af[kActSlots], ae[kActSlots], wf[kMaxWSlots], we[kMaxWSlots], tf[2], te[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[2][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base + kActSlots * kActBytes;
  const uint32_t w_bytes = (uint32_t)(p.NCH / CG) * 128u;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int SW = p.S;                                     // weight-ring slots
  uint32_t acc_stride = 32;
  while ((int)acc_stride < p.NCH) acc_stride <<= 1;
  const uint32_t tmem_cols = 2 * acc_stride;
  const int n_pt = CG == 2 ? (p.Lout + 255) / 256 : p.n_ptiles;      // position tiles (pairs of 128-row tiles for CG = 2)
  const int total_tiles = p.n_chtiles * n_pt * p.B;
  const int tile0 = (int)blockIdx.x / CG, tile_step = (int)gridDim.x / CG;
  const bool prof = p.prof != nullptr;
  const long long t_begin = prof ? clock64() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kActSlots; ++s) { mbar_init(&af[s], 1); mbar_init(&ae[s], 1); }
    for (int s = 0; s < SW; ++s) { mbar_init(&wf[s], 1); mbar_init(&we[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tf[s], 1); mbar_init(&te[s], 4 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmWt) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmX) : "memory");
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();          // the peer's barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  auto tile_of = [&](int t, int& n0, int& l0, int& b) {
    const int ct = t % p.n_chtiles, r = t / p.n_chtiles;
    n0 = ct * p.NCH; l0 = (r % n_pt) * (128 * CG) + 128 * (int)rank; b = r / n_pt;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer
      int sa = 0, sw = 0; uint32_t pa = 0, pw = 0;
      long long w_empty = 0;
      for (int t = tile0; t < total_tiles; t += tile_step) {
        int n0, l0, b;
        tile_of(t, n0, l0, b);
        for (int g = 0; g < p.ngrp; ++g) {
          const TcGroup& gr = p.grp[g];
          int n_a = 0, kof[TC_MAX_TAPS];
          for (int tp = 0; tp < gr.ntaps; ++tp) {
            const TcTap tap = gr.tap[tp];
            if (n0 >= tap.m_lo && n0 < tap.m_hi) kof[n_a++] = tap.kofs;
          }
          if (n_a == 0) continue;
          for (int c = 0; c < gr.nchunk; ++c) {
            const int kc = c * TC_BK;
            mbar_wait_t(&ae[sa], pa ^ 1, w_empty, prof);
            if (CG == 1) {
              mbar_expect_tx(&af[sa], (uint32_t)p.BOXROWS * 128u);
              tma_load_3d(smem_base + (uint32_t)sa * kActBytes, &p.tmX, &af[sa], gr.ch0 + kc, l0 + gr.shift, b);
            } else {
              if (leader) mbar_expect_tx(&af[sa], 2u * (uint32_t)p.BOXROWS * 128u);
              tma_load_3d_pair(smem_base + (uint32_t)sa * kActBytes, &p.tmX, mapa_u32(smem_u32(&af[sa]), 0), gr.ch0 + kc, l0 + gr.shift, b);
            }
            if (++sa == kActSlots) { sa = 0; pa ^= 1; }
            for (int a = 0; a < n_a; ++a) {
              mbar_wait_t(&we[sw], pw ^ 1, w_empty, prof);
              if (CG == 1) {
                mbar_expect_tx(&wf[sw], w_bytes);
                tma_load_2d(w_base + (uint32_t)sw * w_bytes, &p.tmWt, &wf[sw], kof[a] + kc, n0);
              } else {
                if (leader) mbar_expect_tx(&wf[sw], 2u * w_bytes);
                tma_load_2d_pair(w_base + (uint32_t)sw * w_bytes, &p.tmWt, mapa_u32(smem_u32(&wf[sw]), 0), kof[a] + kc, n0 + (int)rank * (p.NCH / 2));
              }
              if (++sw == SW) { sw = 0; pw ^= 1; }
            }
          }
        }
      }
      if (prof) p.prof[blockIdx.x * 8 + 0] = (unsigned long long)w_empty;
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ---------------- MMA issuer: M = 128*CG positions, N = NCH channels, K = 16 per instruction
      const uint32_t idesc = (1u << 4) | (TC_IDESC_AB_FMT << 7) | (TC_IDESC_AB_FMT << 10) | ((uint32_t)(p.NCH >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
      int sa = 0, sw = 0, tl = 0; uint32_t pa = 0, pw = 0;
      long long w_full = 0, w_tmem = 0;
      for (int t = tile0; t < total_tiles; t += tile_step, ++tl) {
        int n0, l0, b;
        tile_of(t, n0, l0, b);
        const int acc = tl & 1;
        mbar_wait_t(&te[acc], ((tl >> 1) & 1) ^ 1, w_tmem, prof);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        uint32_t accumulate = 0;
        for (int g = 0; g < p.ngrp; ++g) {
          const TcGroup& gr = p.grp[g];
          int n_a = 0; uint32_t ro[TC_MAX_TAPS];
          for (int tp = 0; tp < gr.ntaps; ++tp) {
            const TcTap tap = gr.tap[tp];
            if (n0 >= tap.m_lo && n0 < tap.m_hi) ro[n_a++] = (uint32_t)tap.row_off * 8u;
          }
          if (n_a == 0) continue;
          for (int c = 0; c < gr.nchunk; ++c) {
            mbar_wait_t(&af[sa], pa, w_full, prof);
            tc_fence_after();
            const uint64_t adesc0 = umma_desc(smem_base + (uint32_t)sa * kActBytes);
            for (int a = 0; a < n_a; ++a) {
              mbar_wait_t(&wf[sw], pw, w_full, prof);
              tc_fence_after();
              const uint64_t adesc = adesc0 + ro[a];
              const uint64_t bdesc = umma_desc(w_base + (uint32_t)sw * w_bytes);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                if (CG == 1) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
                else umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
                accumulate = 1;
              }
              if (CG == 1) tc_commit(&we[sw]); else tc_commit_pair(&we[sw]);
              if (++sw == SW) { sw = 0; pw ^= 1; }
            }
            if (CG == 1) tc_commit(&ae[sa]); else tc_commit_pair(&ae[sa]);
            if (++sa == kActSlots) { sa = 0; pa ^= 1; }
          }
        }
        if (CG == 1) tc_commit(&tf[acc]); else tc_commit_pair(&tf[acc]);
      }
      if (prof) { p.prof[blockIdx.x * 8 + 1] = (unsigned long long)w_full; p.prof[blockIdx.x * 8 + 2] = (unsigned long long)w_tmem; }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: warp w owns TMEM lanes 32*(w&3)..+31 = positions l0 + 32*(w&3) + lane
    const int q = warp & 3, et = threadIdx.x - 64;
    int tl = 0;
    long long w_acc = 0;
    for (int t = tile0; t < total_tiles; t += tile_step, ++tl) {
      int n0, l0, b;
      tile_of(t, n0, l0, b);
      const int acc = tl & 1;
      float* sb = s_bias[tl & 1];
      for (int i = et; i < p.NCH; i += 128) sb[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int l = l0 + q * 32 + lane;
      const bool valid = l < p.Lout;
      const bool second = p.split_m && n0 >= p.split_m;
      char* obase;                                           // first output element of this thread's row, channel n0
      if (second) obase = (char*)p.out2v + ((long long)b * p.out2_bstride + (long long)l * p.out2_pitch + (n0 - p.split_m)) * 2;
      else if (p.up_cout) obase = (char*)p.out + ((long long)b * p.out_bstride + (long long)(2 * l + n0 / p.up_cout) * p.out_pitch + n0 % p.up_cout) * 2;
      else obase = (char*)p.out + ((long long)b * p.out_bstride + (long long)l * p.out_pitch + n0) * (p.out_f32 ? 4 : 2);
      const h16* rbase = p.res ? p.res + (long long)b * p.res_bstride + (long long)l * p.res_pitch + n0 : nullptr;
      const bool want_stats = p.stats && !second && l0 < p.Lout;    // (a pair's second tile can lie wholly past the clip)
      mbar_wait_t(&tf[acc], (tl >> 1) & 1, w_acc, prof);
      tc_fence_after();
      const uint32_t tlane = tmem_base + (uint32_t)acc * acc_stride + ((uint32_t)(q * 32) << 16);
      // 64 columns per tcgen05.wait::ld: the wait is the expensive part of a TMEM read while the tensor pipe is busy
      for (int c0 = 0; c0 < p.NCH; c0 += 64) {
        uint32_t r0[16], r1[16], r2[16], r3[16];
        tmem_ld16_async(tlane + (uint32_t)c0, r0);
        tmem_ld16_async(tlane + (uint32_t)(c0 + 16), r1);
        tmem_ld16_async(tlane + (uint32_t)(c0 + 32), r2);
        tmem_ld16_async(tlane + (uint32_t)(c0 + 48), r3);
        tmem_ld_wait();
        float a1 = 0.f, a2 = 0.f;
        auto consume = [&](const uint32_t (&r)[16], int cb) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + sb[cb + i];
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { a1 += v[i]; a2 += v[i] * v[i]; }
            if (rbase) {
              const uint4 q0 = __ldcg(reinterpret_cast<const uint4*>(rbase + cb)), q1 = __ldcg(reinterpret_cast<const uint4*>(rbase + cb + 8));
              const h162* h0 = reinterpret_cast<const h162*>(&q0);
              const h162* h1 = reinterpret_cast<const h162*>(&q1);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f0 = h22ff(h0[i]), f1 = h22ff(h1[i]);
                v[2 * i] += f0.x; v[2 * i + 1] += f0.y; v[8 + 2 * i] += f1.x; v[8 + 2 * i + 1] += f1.y;
              }
            }
            if (p.out_f32 && !second) {
              float4* o = reinterpret_cast<float4*>(obase + (size_t)cb * 4);
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
              uint4 o0, o1;
              h162* g0 = reinterpret_cast<h162*>(&o0);
              h162* g1 = reinterpret_cast<h162*>(&o1);
#pragma unroll
              for (int i = 0; i < 4; ++i) { g0[i] = ff2h2(v[2 * i], v[2 * i + 1]); g1[i] = ff2h2(v[8 + 2 * i], v[8 + 2 * i + 1]); }
              uint4* o = reinterpret_cast<uint4*>(obase + (size_t)cb * 2);
              o[0] = o0; o[1] = o1;
            }
          }
        };
        auto flush_slot = [&](int cb) {                      // one 32-channel GroupNorm slot finished: reduce over the warp's 32 positions
          if (want_stats) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              a1 += __shfl_xor_sync(0xffffffffu, a1, o);
              a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (lane == 0)
              p.stats[(((long long)b * p.n_ptiles + l0 / 128) * p.stat_parts + q) * p.stat_slots + ((n0 + cb) >> 5)] = make_float2(a1, a2);
          }
          a1 = 0.f; a2 = 0.f;
        };
        consume(r0, c0); consume(r1, c0 + 16); flush_slot(c0);
        consume(r2, c0 + 32); consume(r3, c0 + 48); flush_slot(c0 + 32);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CG == 1) mbar_arrive(&te[acc]); else mbar_arrive_cluster(mapa_u32(smem_u32(&te[acc]), 0)); }
    }
    if (prof && threadIdx.x == 64) { p.prof[blockIdx.x * 8 + 3] = (unsigned long long)w_acc; p.prof[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - t_begin); }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();          // neither CTA may free TMEM / exit while the pair's MMAs or remote arrives are in flight
  if (warp == 1) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// SIMT check kernel: identical operands, tiling and epilogue semantics, plain FMA loop.
// grid (n_ptiles, Cout/32, B), 128 threads: lane = channel, warp w handles rows w, w+4, ...
__global__ void __launch_bounds__(128) tc_conv_ref_kernel(TcConvParams p, TcRefView v) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pt = blockIdx.x, ch = blockIdx.y * 32 + lane, b = blockIdx.z;
  const int m0 = (ch / TC_BM) * TC_BM;
  const int l0 = pt * p.NT;
  const int Cc = p.up_cout ? p.up_cout : (p.split_m ? p.split_m : p.Cout);
  const float bias = p.bias ? p.bias[ch] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  for (int r = warp; r < p.NT; r += 4) {
    const int l = l0 + r;
    if (l >= p.Lout) break;
    float acc = 0.f;
    for (int g = 0; g < p.ngrp; ++g) {
      const TcGroup& gr = p.grp[g];
      for (int tp = 0; tp < gr.ntaps; ++tp) {
        const TcTap& tap = gr.tap[tp];
        const int row = l + gr.shift + tap.row_off;
        if (m0 < tap.m_lo || m0 >= tap.m_hi || row < 0 || row >= v.Lv) continue;
        const h16* xr = v.x + (long long)b * v.bstride + (long long)row * v.pitch + gr.ch0;
        const h16* wr = v.w + (long long)ch * v.Ktot + tap.kofs;
        for (int c = 0; c < gr.nchunk * TC_BK; ++c) acc += h2f(wr[c]) * h2f(xr[c]);
      }
    }
    float val = acc + bias;
    s1 += val; s2 += val * val;
    if (p.direct) {
      if (p.res) val += h2f(p.res[(long long)b * p.res_bstride + (long long)l * p.res_pitch + ch]);
      const long long o = (long long)b * p.out_bstride + (long long)l * p.out_pitch + ch;
      if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = val;
      else reinterpret_cast<h16*>(p.out)[o] = f2h(val);
    } else if (p.split_m && ch >= p.split_m) {
      v.out2[(long long)b * v.out2_bstride + (long long)l * v.out2_pitch + (ch - p.split_m)] = f2h(val);
    } else {
      const int phase = ch / Cc, cc = ch % Cc, nph = p.up_cout ? 2 : 1;
      v.out[(long long)b * v.out_bstride + (long long)(l * nph + phase) * v.out_pitch + cc] = f2h(val);
    }
  }
  if (p.stats && !(p.split_m && ch >= p.split_m)) {
    __shared__ float red[2][4][32];
    red[0][warp][lane] = s1; red[1][warp][lane] = s2;
    __syncthreads();
    if (warp == 0) {
      float a = red[0][0][lane] + red[0][1][lane] + red[0][2][lane] + red[0][3][lane];
      float c = red[1][0][lane] + red[1][1][lane] + red[1][2][lane] + red[1][3][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
      }
      if (lane == 0) {
        float2* sp = p.stats + (((long long)b * p.n_ptiles + pt) * p.stat_parts) * p.stat_slots + blockIdx.y;
        sp[0] = make_float2(a, c);
        for (int gi = 1; gi < p.stat_parts; ++gi) sp[(long long)gi * p.stat_slots] = make_float2(0.f, 0.f);
      }
    }
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

int make_tmap_x(CUtensorMap* tm, const h16* x, int B, int Lv, int Cv, int pitch, long long bstride, int boxrows) {
  auto enc = get_encode();
  LADIFF_REQUIRE(enc != nullptr, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)Cv, (cuuint64_t)Lv, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)bstride * 2};
  cuuint32_t box[3] = {TC_BK, (cuuint32_t)boxrows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(tm, TC_TMAP_DTYPE, 3, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LADIFF_REQUIRE(r == CUDA_SUCCESS, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled(X B%d L%d C%d pitch %d box %d) failed: %d", B, Lv, Cv,
                 pitch, boxrows, (int)r);
  return 0;
}

// output map {Cc, phases, rows, B}: element (c, ph, l, b) at out[b*bstride + (l*phases + ph)*pitch + c]
int make_tmap_y(CUtensorMap* tm, const h16* out, int Cc, int phases, int rows, int B, int pitch, long long bstride, int boxrows) {
  auto enc = get_encode();
  LADIFF_REQUIRE(enc != nullptr, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)Cc, (cuuint64_t)phases, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 2 * phases, (cuuint64_t)bstride * 2};
  cuuint32_t box[4] = {32, 1, (cuuint32_t)boxrows, 1};      // one epilogue warp's slice: 32 channels x boxrows rows
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, TC_TMAP_DTYPE, 4, (void*)out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LADIFF_REQUIRE(r == CUDA_SUCCESS, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled(Y C%d ph%d rows%d B%d pitch %d box %d) failed: %d", Cc,
                 phases, rows, B, pitch, boxrows, (int)r);
  return 0;
}

}  // namespace

int tc_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;   // every GPU of one box is the same part
}

static int make_tmap_wt(CUtensorMap* tm, const h16* w, int Cout, int Ktot, int nch) {
  auto enc = get_encode();
  LADIFF_REQUIRE(enc != nullptr, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {TC_BK, (cuuint32_t)nch};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(tm, TC_TMAP_DTYPE, 2, (void*)w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LADIFF_REQUIRE(r == CUDA_SUCCESS, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled(Wt %dx%d box %d) failed: %d", Cout, Ktot, nch, (int)r);
  return 0;
}

int tc_make_tmap_w(CUtensorMap* tm, const h16* w, int Cout, int Ktot) {
  auto enc = get_encode();
  LADIFF_REQUIRE(enc != nullptr, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {TC_BK, TC_BM};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(tm, TC_TMAP_DTYPE, 2, (void*)w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LADIFF_REQUIRE(r == CUDA_SUCCESS, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled(W %dx%d) failed: %d", Cout, Ktot, (int)r);
  return 0;
}

// Tile shape for clips of Lout rows: minimise rounds-over-the-SMs x per-tile cost; ties go to the wider tile (less operand traffic).
// Tile shape for clips of Lout rows.  want_nt/want_nclip > 0 force a shape (the plan builder's autotuner tries several);
// otherwise: minimise rounds-over-the-SMs x per-tile cost; ties go to the wider tile (less operand traffic).
static void pick_tiling(int Lout, int B, int MT, int halo2, int want_nt, int want_nclip, int* NT, int* NCLIP, int* n_ptiles) {
  const int nsm = tc_num_sms();
  long best = -1;
  const int maxn = halo2 ? 240 : 256;
  if (Lout + halo2 <= 120) {                      // several clips per tile: region = clip + halo rows, multiple of 16
    const int rp = cdiv(Lout + halo2, 16) * 16;
    for (int nc = 256 / rp; nc >= 1; --nc) {
      if (nc > B && nc > 1) continue;
      if (want_nclip > 0 && nc != want_nclip) continue;
      const long tiles = (long)MT * cdiv(B, nc);
      const long cost = (long)cdiv((int)tiles, nsm) * (nc * rp + 32);
      if (best < 0 || cost < best) { best = cost; *NT = rp; *NCLIP = nc; *n_ptiles = 1; }
    }
    if (best >= 0) return;
    *NT = rp; *NCLIP = 1; *n_ptiles = 1;
    return;
  }
  static const int env_nt = getenv("LADIFF_TC_NT") ? atoi(getenv("LADIFF_TC_NT")) : 0;   // experiment knob
  const int force_nt = want_nt > 0 ? want_nt : env_nt;
  if (force_nt >= 64 && force_nt <= maxn && force_nt % 16 == 0) {
    *NT = force_nt; *NCLIP = 1; *n_ptiles = cdiv(Lout, force_nt);
    return;
  }
  for (int nt = maxn; nt >= 64; nt -= 16) {
    const int np = cdiv(Lout, nt);
    const long tiles = (long)MT * B * np;
    const long cost = (long)cdiv((int)tiles, nsm) * (nt + 32);
    if (best < 0 || cost < best) { best = cost; *NT = nt; *NCLIP = 1; *n_ptiles = np; }
  }
}

int tc_conv_plan(const TcConvDesc& d, TcConvParams* pp, TcRefView* rv) {
  TcConvParams& p = *pp;
  memset(&p, 0, sizeof(p));
  LADIFF_REQUIRE(d.CoutV % TC_BM == 0 && d.Cin % TC_BK == 0 && d.B >= 1, LADIFF_ERR_ARG, "tc_conv: CoutV=%d Cin=%d B=%d", d.CoutV, d.Cin, d.B);
  const int nch = d.Cin / TC_BK;
  int Lout = d.Lin, Lv = d.Lin, Cv = d.Cin, pitch_v = d.x_pitch, halo2 = 0;
  // taps as (view-row shift, view channel offset, weight K offset, m range)
  struct T { int shift, ch0, kofs, m_lo, m_hi; } taps[8];
  int ntap = 0;
  if (d.kind == TC_KIND_DOWN) {         // Conv1d(k4, s2, p1) on the row-pair view: even rows = channels [0,C), odd rows = [pitch, pitch+C)
    LADIFF_REQUIRE(d.K == 4 && d.Lin % 2 == 0, LADIFF_ERR_ARG, "tc_conv: stride-2 conv needs k=4 and an even length (%d)", d.Lin);
    Lout = d.Lin / 2; Lv = d.Lin / 2; pitch_v = 2 * d.x_pitch; Cv = d.x_pitch + d.Cin;
    const int sh[4] = {-1, 0, 0, 1}, c0[4] = {d.x_pitch, 0, d.x_pitch, 0};
    for (int s = 0; s < 4; ++s) taps[ntap++] = T{sh[s], c0[s], s * d.Cin, 0, d.CoutV};
    halo2 = 2;
  } else if (d.kind == TC_KIND_UP) {    // nearest x2 + Conv1d(k3, p1): even phase rows [0, Cout), odd phase [Cout, 2Cout)
    LADIFF_REQUIRE(d.K == 3, LADIFF_ERR_ARG, "tc_conv: upsample conv needs k=3");
    taps[ntap++] = T{-1, 0, 0, 0, d.CoutV / 2};
    taps[ntap++] = T{0, 0, d.Cin, 0, d.CoutV};
    taps[ntap++] = T{1, 0, 2 * d.Cin, d.CoutV / 2, d.CoutV};
    halo2 = 2;
    LADIFF_REQUIRE((d.CoutV / 2) % TC_BM == 0, LADIFF_ERR_ARG, "tc_conv: upsample conv needs Cout %% 128 == 0");
  } else {
    LADIFF_REQUIRE(d.K >= 1 && d.K <= TC_MAX_TAPS && (d.K & 1), LADIFF_ERR_ARG, "tc_conv: K=%d", d.K);
    const int m_hi = d.split_m ? d.split_m : d.CoutV;
    for (int s = 0; s < d.K; ++s) taps[ntap++] = T{s - (d.K - 1) / 2, 0, s * d.Cin, 0, m_hi};
    halo2 = d.K - 1;
  }
  if (d.split_m) {                      // fused 1x1 conv of the same input on rows [split_m, CoutV)
    LADIFF_REQUIRE(d.kind == TC_KIND_PLAIN && d.split_m % TC_BM == 0 && d.split_m < d.CoutV && d.out2 && !d.out32 && !d.res && ntap < 8,
                   LADIFF_ERR_ARG, "tc_conv: bad second-output configuration");
    int at = ntap;                      // keep taps ordered by shift: insert after the last tap with shift <= 0
    for (int i = 0; i < ntap; ++i) if (taps[i].shift > 0) { at = i; break; }
    for (int i = ntap; i > at; --i) taps[i] = taps[i - 1];
    taps[at] = T{0, 0, 0, d.split_m, d.CoutV};
    ++ntap;
  }
  // groups: taps that read the same channel range share one shared-memory tile; a group is one pipeline stage per chunk, so it
  // holds at most kStageTaps weight tiles for any output-channel tile (k=7 becomes 3+3+1)
  p.ngrp = 0;
  int a_cap = 1, tile_halo = 0;
  for (int i = 0; i < ntap; ++i) {
    int g = -1;
    if (d.tap_share)
      for (int k = 0; k < p.ngrp; ++k) {
        if (p.grp[k].ch0 != taps[i].ch0) continue;
        int overlap = 0;
        for (int t2 = 0; t2 < p.grp[k].ntaps; ++t2)
          overlap += (p.grp[k].tap[t2].m_lo < taps[i].m_hi && taps[i].m_lo < p.grp[k].tap[t2].m_hi) ? 1 : 0;
        if (overlap < kStageTaps && p.grp[k].ntaps < TC_MAX_TAPS) g = k;
      }
    if (g < 0) {
      LADIFF_REQUIRE(p.ngrp < TC_MAX_GRP, LADIFF_ERR_ARG, "tc_conv: too many tap groups");
      g = p.ngrp++;
      p.grp[g].ch0 = taps[i].ch0; p.grp[g].shift = taps[i].shift; p.grp[g].nchunk = nch; p.grp[g].ntaps = 0;
    }
    TcGroup& gr = p.grp[g];
    TcTap& tp = gr.tap[gr.ntaps++];
    tp.row_off = taps[i].shift - gr.shift;      // taps are listed by increasing shift, so row_off >= 0
    tp.kofs = taps[i].kofs; tp.m_lo = taps[i].m_lo; tp.m_hi = taps[i].m_hi;
    LADIFF_REQUIRE(tp.row_off >= 0 && tp.row_off <= 7, LADIFF_ERR_ARG, "tc_conv: tap row offset %d", tp.row_off);
    tile_halo = tp.row_off > tile_halo ? tp.row_off : tile_halo;
  }
  for (int g = 0; g < p.ngrp; ++g)               // weight tiles a stage must hold = most taps any one m-tile uses
    for (int m0 = 0; m0 < d.CoutV; m0 += TC_BM) {
      int n = 0;
      for (int t2 = 0; t2 < p.grp[g].ntaps; ++t2) n += (m0 >= p.grp[g].tap[t2].m_lo && m0 < p.grp[g].tap[t2].m_hi) ? 1 : 0;
      a_cap = n > a_cap ? n : a_cap;
    }
  p.a_cap = a_cap;
  // The tiling is decided on the conv's own halo (K - 1 rows), not on the halo one shared-memory tile happens to need: the
  // tap-shared and the per-tap (check) variants of the same conv must cut the clips identically (same GroupNorm partials).
  if (halo2 < tile_halo) halo2 = tile_halo;
  p.MT = d.CoutV / TC_BM;
  p.B = d.B; p.Lout = Lout; p.Cout = d.CoutV; p.bias = d.bias; p.stats = d.stats;
  p.up_cout = d.kind == TC_KIND_UP ? d.CoutV / 2 : 0;
  p.split_m = d.split_m;
  p.stat_slots = (d.split_m ? d.split_m : d.CoutV) / 32;
  p.stat_parts = kEpiGroups;
  if (d.want_transposed) {
    // positions-on-M kernel: 128-row position tiles of one clip x NCH-channel tiles; NCH must divide every tap's channel range
    int nch = 0;
    for (int cand : {256, 192, 128}) {
      bool ok = d.CoutV % cand == 0 && (!d.split_m || d.split_m % cand == 0) && (!p.up_cout || p.up_cout % cand == 0);
      for (int g = 0; ok && g < p.ngrp; ++g)
        for (int t2 = 0; ok && t2 < p.grp[g].ntaps; ++t2)
          ok = p.grp[g].tap[t2].m_lo % cand == 0 && (p.grp[g].tap[t2].m_hi % cand == 0 || p.grp[g].tap[t2].m_hi == d.CoutV);
      if (ok) { nch = cand; break; }
    }
    LADIFF_REQUIRE(nch > 0 && tile_halo <= 8, LADIFF_ERR_ARG, "tc_conv(transposed): no channel tile fits");
    p.transposed = d.want_transposed == 2 ? 2 : 1; p.NCH = nch; p.n_chtiles = d.CoutV / nch; p.stat_parts = 4;
    LADIFF_REQUIRE(p.transposed == 1 || nch % 32 == 0, LADIFF_ERR_ARG, "tc_conv(pair): channel tile");
    p.NT = 128; p.NCLIP = 1; p.n_ptiles = cdiv(Lout, 128); p.NMMA = nch; p.n_ntiles = d.B * p.n_ptiles; p.BOXROWS = 136; p.CR = 32;
    const long wslots = ((long)kSmemLimitT - 2048 - (long)kActSlots * kActBytes) / ((long)(nch / p.transposed) * 128);
    p.S = wslots > kMaxWSlots ? kMaxWSlots : (int)wslots;
    LADIFF_REQUIRE(p.S >= 3, LADIFF_ERR_ARG, "tc_conv(transposed): weight ring too small");
    p.direct = (d.out32 != nullptr || d.res != nullptr) ? 1 : 0;
    p.tmW = *d.tmW;
    int rc = make_tmap_x(&p.tmX, d.x, d.B, Lv, Cv, pitch_v, d.x_bstride, p.BOXROWS);
    if (rc) return rc;
    rc = make_tmap_wt(&p.tmWt, d.w, d.CoutV, d.Ktot, nch / p.transposed);
    if (rc) return rc;
    if (d.out32) { p.out = d.out32; p.out_f32 = 1; p.out_pitch = d.CoutV; p.out_bstride = (long long)Lout * d.CoutV; }
    else { p.out = d.out; p.out_f32 = 0; p.out_pitch = d.out_pitch; p.out_bstride = d.out_bstride; }
    p.out2v = d.out2; p.out2_bstride = d.out2_bstride; p.out2_pitch = d.out2_pitch;
    p.res = d.res; p.res_bstride = d.res_bstride; p.res_pitch = d.res_pitch;
    LADIFF_REQUIRE(((uintptr_t)p.out % 16) == 0 && p.out_pitch % 8 == 0 && p.out_bstride % 8 == 0 &&
                       (!d.out2 || (((uintptr_t)d.out2 % 16) == 0 && d.out2_pitch % 8 == 0 && d.out2_bstride % 8 == 0)) &&
                       (!d.res || (((uintptr_t)d.res % 16) == 0 && d.res_pitch % 8 == 0 && d.res_bstride % 8 == 0)),
                   LADIFF_ERR_ARG, "tc_conv(transposed): output / residual views are not 16-byte aligned");
    if (rv) {
      rv->x = d.x; rv->bstride = d.x_bstride; rv->pitch = pitch_v; rv->Lv = Lv; rv->Cv = Cv; rv->w = d.w; rv->Ktot = d.Ktot;
      rv->out = d.out; rv->out_bstride = d.out_bstride; rv->out_pitch = d.out_pitch;
      rv->out2 = d.out2; rv->out2_bstride = d.out2_bstride; rv->out2_pitch = d.out2_pitch;
    }
    return 0;
  }
  pick_tiling(Lout, d.B, p.MT, halo2, d.want_nt, d.want_nclip, &p.NT, &p.NCLIP, &p.n_ptiles);
  p.NMMA = p.NT * p.NCLIP;
  p.n_ntiles = p.NCLIP == 1 ? d.B * p.n_ptiles : cdiv(d.B, p.NCLIP);
  p.BOXROWS = p.NCLIP == 1 ? p.NT + (tile_halo ? 8 : 0) : p.NT;
  LADIFF_REQUIRE(p.NMMA % 16 == 0 && p.NMMA >= 16 && p.NMMA <= 256 && p.BOXROWS <= 256, LADIFF_ERR_ARG, "tc_conv: tile N=%d box=%d", p.NMMA,
                 p.BOXROWS);
  LADIFF_REQUIRE(p.NCLIP == 1 || Lout + tile_halo <= p.NT, LADIFF_ERR_ARG, "tc_conv: clip region too small");
  p.direct = (d.out32 != nullptr || d.res != nullptr) ? 1 : 0;
  p.b_slot_bytes = (int)align_up((size_t)p.NCLIP * p.BOXROWS * 128, 1024);
  p.stage_bytes = p.a_cap * (int)A_BYTES + p.b_slot_bytes;
  // shared-memory budget: S pipeline stages + 2 epilogue staging chunks per epilogue warpgroup + 1 KB alignment + 1 KB tap over-read.
  // Smaller staging chunks (32 rows) are used when they buy another pipeline stage.
  // LADIFF_TC_SMEM_KB caps a conv CTA's shared memory: what it leaves free lets the small kernels of ANOTHER stream (a second batch
  // in flight) become resident next to a running conv CTA instead of waiting for the conv to drain
  static const long smem_cap = getenv("LADIFF_TC_SMEM_KB") ? atol(getenv("LADIFF_TC_SMEM_KB")) * 1024 : (long)kSmemLimit;
  auto stages_for = [&](int cr) {
    const long rest = (long)(d.want_two_per_sm ? kSmemLimit2 : (smem_cap < (long)kSmemLimit ? smem_cap : (long)kSmemLimit)) - 2048 - (p.direct ? (d.res ? (long)p.NMMA * 256 : 0) : (long)2 * kEpiGroups * cr * 256);
    const int s2 = (int)(rest / p.stage_bytes);
    return s2 > kMaxStages ? kMaxStages : s2;
  };
  p.CR = p.NT < 32 ? p.NT : 32;
  if (p.NT > 16 && stages_for(16) > stages_for(p.CR)) p.CR = 16;
  p.S = stages_for(p.CR);
  p.minb = 1;
  if (d.want_two_per_sm) {
    LADIFF_REQUIRE(p.NMMA <= 128 && p.S >= 2, LADIFF_ERR_ARG, "tc_conv: shape does not fit two CTAs per SM");
    p.minb = 2;
  }
  LADIFF_REQUIRE(p.S >= 2, LADIFF_ERR_ARG, "tc_conv: tile N=%d with %d taps per stage does not fit two pipeline stages", p.NMMA, p.a_cap);
  p.tmW = *d.tmW;
  int rc = make_tmap_x(&p.tmX, d.x, d.B, Lv, Cv, pitch_v, d.x_bstride, p.BOXROWS);
  if (rc) return rc;
  if (d.want_pair) {
    LADIFF_REQUIRE(!d.want_two_per_sm && p.n_ntiles >= 2, LADIFF_ERR_ARG, "tc_conv: pair mode needs one CTA per SM and >= 2 position tiles");
    p.pair = 1;
    rc = make_tmap_wt(&p.tmWh, d.w, d.CoutV, d.Ktot, 64);
    if (rc) return rc;
  }
  if (p.direct) {
    if (d.out32) { p.out = d.out32; p.out_f32 = 1; p.out_pitch = d.CoutV; p.out_bstride = (long long)Lout * d.CoutV; }
    else { p.out = d.out; p.out_f32 = 0; p.out_pitch = d.out_pitch; p.out_bstride = d.out_bstride; }
    p.res = d.res; p.res_bstride = d.res_bstride; p.res_pitch = d.res_pitch;
    LADIFF_REQUIRE(d.kind != TC_KIND_UP, LADIFF_ERR_ARG, "tc_conv: the upsample conv has no direct epilogue");
  } else {
    const int phases = p.up_cout ? 2 : 1, Cc = p.up_cout ? p.up_cout : (d.split_m ? d.split_m : d.CoutV);
    LADIFF_REQUIRE(((uintptr_t)d.out % 16) == 0 && d.out_pitch % 8 == 0 && d.out_bstride % 8 == 0, LADIFF_ERR_ARG,
                   "tc_conv: output view is not 16-byte aligned");
    rc = make_tmap_y(&p.tmY, d.out, Cc, phases, Lout, d.B, d.out_pitch, d.out_bstride, p.CR);
    if (rc) return rc;
    const int rem = p.NT % p.CR;
    rc = make_tmap_y(&p.tmYr, d.out, Cc, phases, Lout, d.B, d.out_pitch, d.out_bstride, rem ? rem : p.CR);
    if (rc) return rc;
    if (d.split_m) {
      LADIFF_REQUIRE(((uintptr_t)d.out2 % 16) == 0 && d.out2_pitch % 8 == 0 && d.out2_bstride % 8 == 0, LADIFF_ERR_ARG,
                     "tc_conv: second output view is not 16-byte aligned");
      rc = make_tmap_y(&p.tmY2, d.out2, d.CoutV - d.split_m, 1, Lout, d.B, d.out2_pitch, d.out2_bstride, p.CR);
      if (rc) return rc;
      rc = make_tmap_y(&p.tmY2r, d.out2, d.CoutV - d.split_m, 1, Lout, d.B, d.out2_pitch, d.out2_bstride, rem ? rem : p.CR);
      if (rc) return rc;
    }
  }
  if (rv) {
    rv->x = d.x; rv->bstride = d.x_bstride; rv->pitch = pitch_v; rv->Lv = Lv; rv->Cv = Cv; rv->w = d.w; rv->Ktot = d.Ktot;
    rv->out = d.out; rv->out_bstride = d.out_bstride; rv->out_pitch = d.out_pitch;
    rv->out2 = d.out2; rv->out2_bstride = d.out2_bstride; rv->out2_pitch = d.out2_pitch;
  }
  return 0;
}

static size_t tc_smem_bytes(const TcConvParams& p) {
  return (size_t)p.S * p.stage_bytes + (p.direct ? (p.res ? (size_t)p.NMMA * 256 : 0) : (size_t)2 * kEpiGroups * p.CR * 256) + 2048;
}

template <int CG>
static int tc_conv_t_launch_cg(const TcConvParams& p, cudaStream_t st) {
  const size_t smem = (size_t)kActSlots * kActBytes + (size_t)p.S * (p.NCH / CG) * 128 + 2048;
  LADIFF_REQUIRE(smem <= kSmemLimitT, LADIFF_ERR_ARG, "tc_conv(transposed): smem %zu too large", smem);
  static unsigned long long attr_set = 0;
  if (ladiff_first_on_device(&attr_set)) {
    LADIFF_CUDA_OK(cudaFuncSetAttribute(tc_conv_t_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimitT));
  }
  const int n_pt = CG == 2 ? (p.Lout + 255) / 256 : p.n_ptiles;
  const int tiles = p.n_chtiles * n_pt * p.B, nsm = tc_num_sms();
  const int slots = nsm / CG;
  const int grid = (tiles < slots ? tiles : slots) * CG;
  static const bool want_prof = getenv("LADIFF_TC_PROF") != nullptr;
  TcConvParams q = p;
  unsigned long long* dprof = nullptr;
  if (want_prof) {
    LADIFF_CUDA_OK(cudaMalloc((void**)&dprof, sizeof(unsigned long long) * 8 * grid));
    LADIFF_CUDA_OK(cudaMemset(dprof, 0, sizeof(unsigned long long) * 8 * grid));
    q.prof = dprof;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreadsT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (ladiff_pdl_enabled() && !want_prof) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  LADIFF_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_conv_t_kernel<CG>, q));
  if (!want_prof) return 0;
  LADIFF_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<unsigned long long> hp((size_t)8 * grid);
  LADIFF_CUDA_OK(cudaMemcpy(hp.data(), dprof, sizeof(unsigned long long) * hp.size(), cudaMemcpyDeviceToHost));
  cudaFree(dprof);
  double a[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < grid; i += CG) for (int k = 0; k < 5; ++k) a[k] += (double)hp[(size_t)i * 8 + k] / (grid / CG);
  fprintf(stderr, "[tc_prof_t] CG=%d Cout=%d NCH=%d Wslots=%d grid=%d tiles=%d | cycles/CTA(leader): total %.0f  producer-wait-empty %.0f  mma-wait-full %.0f  "
                  "mma-wait-tmem %.0f  epi-wait-acc %.0f\n", CG, p.Cout, p.NCH, p.S, grid, tiles, a[4], a[0], a[1], a[2], a[3]);
  return 0;
}
static int tc_conv_t_launch(const TcConvParams& p, cudaStream_t st) {
  return p.transposed == 2 ? tc_conv_t_launch_cg<2>(p, st) : tc_conv_t_launch_cg<1>(p, st);
}

template <int MINB>
static int tc_conv_launch_minb(const TcConvParams& p, cudaStream_t st) {
  const size_t smem = tc_smem_bytes(p);
  const size_t limit = MINB == 2 ? kSmemLimit2 : kSmemLimit;
  LADIFF_REQUIRE(smem <= limit, LADIFF_ERR_ARG, "tc_conv: smem %zu too large", smem);
  static unsigned long long attr_set = 0;     // opt in to the dynamic shared memory once per device
  if (ladiff_first_on_device(&attr_set)) {
    LADIFF_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel<MINB, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    LADIFF_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel<MINB, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    LADIFF_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel<MINB, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
  }
  const int tiles = p.pair ? p.MT * ((p.n_ntiles + 1) / 2) : p.MT * p.n_ntiles, slots = p.pair ? tc_num_sms() / 2 : tc_num_sms() * MINB;
  const int grid = (tiles < slots ? tiles : slots) * (p.pair ? 2 : 1);
  static const bool want_prof = getenv("LADIFF_TC_PROF") != nullptr;   // debug aid: per-role mbarrier wait cycles, printed per launch
  if (!want_prof && p.pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    int na = 1;
    if (ladiff_pdl_enabled()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    LADIFF_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_conv_kernel<MINB, 1, 0>, p));
    return 0;
  }
  if (!want_prof) {
    LADIFF_CUDA_OK(launch_pdl(tc_conv_kernel<MINB, 0, 0>, dim3(grid), dim3(kThreads), smem, st, p));
    return 0;
  }
  LADIFF_REQUIRE(!p.pair, LADIFF_ERR_ARG, "LADIFF_TC_PROF does not cover the pair mode");
  TcConvParams q = p;
  q.dbg = getenv("LADIFF_TC_DBG") ? atoi(getenv("LADIFF_TC_DBG")) : 0;   // 1: no TMA after ring fill, 2: no epilogue, 4: no MMA, 8: epilogue without tcgen05.ld, 16: epilogue without smem/TMA stores
  unsigned long long* dprof = nullptr;
  LADIFF_CUDA_OK(cudaMalloc((void**)&dprof, sizeof(unsigned long long) * 8 * grid));
  LADIFF_CUDA_OK(cudaMemset(dprof, 0, sizeof(unsigned long long) * 8 * grid));
  q.prof = dprof;
  tc_conv_kernel<MINB, 0, 1><<<grid, kThreads, smem, st>>>(q);
  LADIFF_CUDA_OK(cudaGetLastError());
  LADIFF_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<unsigned long long> hp((size_t)8 * grid);
  LADIFF_CUDA_OK(cudaMemcpy(hp.data(), dprof, sizeof(unsigned long long) * hp.size(), cudaMemcpyDeviceToHost));
  cudaFree(dprof);
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < grid; ++i) for (int k = 0; k < 8; ++k) a[k] += (double)hp[(size_t)i * 8 + k] / grid;
  fprintf(stderr, "[tc_prof] dbg=%d minb=%d Cout=%d N=%d(NT=%d x%d) S=%d a_cap=%d grid=%d tiles=%d | cycles/CTA: total %.0f  producer-wait-empty %.0f  mma-wait-full %.0f  "
                  "mma-wait-tmem %.0f  epi-wait-acc %.0f  epi(wg0): entry %.0f body %.0f exit %.0f\n", q.dbg, MINB, p.Cout, p.NMMA, p.NT, p.NCLIP, p.S, p.a_cap, grid, tiles, a[4], a[0], a[1], a[2], a[3], a[5], a[6], a[7]);
  return 0;
}

int tc_conv_launch(const TcConvParams& p, cudaStream_t st) {
  if (p.transposed) return tc_conv_t_launch(p, st);
  return p.minb == 2 ? tc_conv_launch_minb<2>(p, st) : tc_conv_launch_minb<1>(p, st);
}

int tc_conv_ref_launch(const TcConvParams& p, const TcRefView& v, cudaStream_t st) {
  dim3 grid(p.NCLIP == 1 ? p.n_ptiles : 1, p.Cout / 32, p.B);
  tc_conv_ref_kernel<<<grid, 128, 0, st>>>(p, v);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
