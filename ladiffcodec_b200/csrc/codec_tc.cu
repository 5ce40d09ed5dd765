// tcgen05 Conv1d for the SEANet codec with fp32-level accuracy (3xTF32).
//
// Replaces the fp32 FMA kernel (conv1d_f32_v2_kernel, codec_ops.cu) for the convs of the cond encoder, the cond upsamplers, the
// decoder and the LSTM input projections (reference srcs/modules/seanet.py:45-63,143-200,  conv.py:217-274, lstm.py:22-28).
// The codec feeds a vector quantiser (argmin over 1024 codewords, quantization/core_vq.py), so 10-bit operands are not an option;
// every fp32 value is split  v = hi + lo  (hi = rna_tf32(v), lo = rna_tf32(v - hi)) and  x*w ~= lo*hi + hi*lo + hi*hi  is three
// kind::tf32 MMAs into the same fp32 TMEM accumulator (dropped lo*lo term and the rounding of lo: <= 2^-22 relative, unbiased).
// (A first version on mma.sync (HMMA.1688.F32.TF32, operands fetched from shared memory with LDS) reached 55-60 TFLOP/s of TF32,
//  i.e. the speed of the FMA kernel once divided by three.  The instruction itself peaks at 278 TFLOP/s with register operands
//  (profiles/ubench/mma_sync_rate.cu), a quarter of tcgen05 kind::tf32's nominal rate - and tcgen05 reads both operands from
//  shared memory with no register traffic at all, which is what a conv whose operands must be transposed and split needs.)
//
// Formulation (as conv1d_f32_v2_kernel): a conv of stride S with K = KT*S taps is a stride-1 conv with KT taps over the S "phase
// channels" of every input channel: cv = ci*S + p,  xv[cv][u] = xpad[ci][u*S + p - padL];  transposed convs arrive here as
// K = 2 stride-1 convs over s*Cout virtual output channels with a phase-interleaved store (ConvF32Args.il_*).
//   D[co, t] (TMEM fp32: lane = output channel, column = position) += W[co, (k, cv)] * X[(k, cv), t]
//   A = weights, K-major [128 channels][32 cv] per (tap, plane), TMA SWIZZLE_128B from a [plane][tap][CoutV][CinV] copy made at load
//   B = activations, K-major [T_T + 8 positions][32 cv]: the fill warps read the NCL fp32 input (coalesced along time), apply ELU,
//       split, and store both planes in the SWIZZLE_128B layout (conflict-free); a tap is a row shift of the same tile (descriptor
//       start + tap*128 B: the swizzle is a function of the absolute shared-memory address)
// CTA: warp 0 TMA producer (weights), warp 1 TMEM allocation + MMA issue, warps 2-9 fill then epilogue
//      (tcgen05.ld -> shared [channel][T_T+1] -> coalesced, optionally phase-interleaved global stores with bias / residual).
// The tensor core adds into its accumulator with truncation (a bias that grows with the chain length): the small products have
// their own accumulator and, for the quantised encoder, the hi*hi products of consecutive K blocks rotate over three (see launcher).
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

#include "codec_ops.cuh"
#include "common.cuh"

namespace {

constexpr int CT_THREADS = 320, CT_M = 128, CT_N = 128, CT_KB = 32;       // tile: 128 channels x 128 positions, K block 32 cv
constexpr int CT_XROWS = CT_N + 8;                                         // + tap halo (KT - 1 <= 8)
constexpr uint32_t CT_XPLANE = CT_XROWS * 128, CT_XSTAGE = 2 * CT_XPLANE;  // 17408 B per plane
constexpr uint32_t CT_WPLANE = CT_M * 128, CT_WSTAGE = 2 * CT_WPLANE;      // 16384 B per plane
constexpr int CT_MAXS = 4;
constexpr int CT_EPW = CT_N + 1;

__device__ __forceinline__ uint32_t c_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool c_elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void c_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void c_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(c_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void c_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void c_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = c_smem_u32(bar);
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    if (spins == 64) t0 = clock64();
    if (spins > 64 && (spins & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();     // protocol bug -> launch error, not a hang
  }
}
__device__ __forceinline__ void c_tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tm), "r"(c_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void c_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c_smem_u32(bar)) : "memory");
}
// K-major SWIZZLE_128B matrix descriptor (SBO = 1024 B: 8 rows x 128 B), as in tc_conv.cu
__device__ __forceinline__ uint64_t c_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void c_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void c_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ float c_tf32_rna(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

struct CodecTcParams {
  ConvF32Args a;
  int S, KT, CinV, nkb;       // phases, taps, virtual input channels, K blocks of 32
  int XS, WS, NHI;            // activation / weight ring depths; hi*hi accumulator chains (the small terms have their own: NHI + 1 in all)
};

__global__ void __launch_bounds__(CT_THREADS) codec_tc_kernel(const __grid_constant__ CodecTcParams p, const __grid_constant__ CUtensorMap tmW) {
  extern __shared__ uint8_t ct_smem_raw[];
  __shared__ __align__(8) uint64_t x_full[CT_MAXS], x_empty[CT_MAXS], w_full[CT_MAXS], w_empty[CT_MAXS], acc_full;
  __shared__ uint32_t tmem_base_s;
  const ConvF32Args& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (c_smem_u32(ct_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = ct_smem_raw + (smem_base - c_smem_u32(ct_smem_raw));
  const uint32_t x_ring = smem_base, w_ring = smem_base + (uint32_t)p.XS * CT_XSTAGE;
  const int b = blockIdx.z, co0 = blockIdx.y * CT_M, t0 = blockIdx.x * CT_N;
  const uint32_t tmem_cols = p.NHI == 1 ? 256u : 512u;
  const int KT = p.KT, nkb = p.nkb;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.XS; ++i) { c_mbar_init(&x_full[i], 8); c_mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < p.WS; ++i) { c_mbar_init(&w_full[i], 1); c_mbar_init(&w_empty[i], 1); }
    c_mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(c_smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ---- weights: per (K block, tap) the hi and lo tiles [128 channels][32 cv] of this channel tile
    if (c_elect_one()) {
      int j = 0;
      for (int kb = 0; kb < nkb; ++kb)
        for (int k = 0; k < KT; ++k, ++j) {
          const int s = j % p.WS;
          c_mbar_wait(&w_empty[s], (((uint32_t)(j / p.WS)) & 1u) ^ 1u);
          c_mbar_expect_tx(&w_full[s], CT_WSTAGE);
          const uint32_t dst = w_ring + (uint32_t)s * CT_WSTAGE;
          c_tma_load_2d(dst, &tmW, &w_full[s], kb * CT_KB, k * a.CoutV + co0);
          c_tma_load_2d(dst + CT_WPLANE, &tmW, &w_full[s], kb * CT_KB, (KT + k) * a.CoutV + co0);
        }
    }
  } else if (warp == 1) {
    // ---- MMA issue: M = 128, N = 128, K = 8 per instruction; per (K block, tap): 4 K steps x (lo*hi, hi*lo, hi*hi)
    if (c_elect_one()) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CT_N >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);
      int j = 0;
      uint32_t acc_lo = 0u;
      for (int kb = 0; kb < nkb; ++kb) {
        const int xs = kb % p.XS;
        c_mbar_wait(&x_full[xs], ((uint32_t)(kb / p.XS)) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_hi = tmem_base + (uint32_t)((kb % p.NHI) * CT_N), d_lo = tmem_base + (uint32_t)(p.NHI * CT_N);
        uint32_t acc_hi = kb >= p.NHI ? 1u : 0u;
        const uint64_t bh = c_desc(x_ring + (uint32_t)xs * CT_XSTAGE), bl = c_desc(x_ring + (uint32_t)xs * CT_XSTAGE + CT_XPLANE);
        for (int k = 0; k < KT; ++k, ++j) {
          const int ws = j % p.WS;
          c_mbar_wait(&w_full[ws], ((uint32_t)(j / p.WS)) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t ah = c_desc(w_ring + (uint32_t)ws * CT_WSTAGE), al = c_desc(w_ring + (uint32_t)ws * CT_WSTAGE + CT_WPLANE);
          const uint64_t roff = (uint64_t)(k * 8);            // tap = row shift of the activation tile: k * 128 B >> 4
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {                    // 8 tf32 = 32 B per K step
            c_mma_tf32(d_lo, al + 2 * kk, bh + roff + 2 * kk, idesc, acc_lo);
            c_mma_tf32(d_lo, ah + 2 * kk, bl + roff + 2 * kk, idesc, 1u);
            c_mma_tf32(d_hi, ah + 2 * kk, bh + roff + 2 * kk, idesc, acc_hi);
            acc_lo = 1u; acc_hi = 1u;
          }
          c_commit(&w_empty[ws]);
        }
        c_commit(&x_empty[xs]);
      }
      c_commit(&acc_full);
    }
  } else {
    // ---- fill: thread = (channel cv0 + 4*fw + c_sub, rows u_sub + 8*i); a warp instruction covers 8 rows x 4 channels:
    //      32-byte runs in global memory (stride-1) and 32 distinct banks in the swizzled tile
    const int fw = warp - 2, tid = threadIdx.x - 64;
    const int u_sub = lane & 7, c_sub = lane >> 3, cl = 4 * fw + c_sub;
    const int XW = CT_N + KT - 1;
    const float* xb = a.x + (long long)b * a.Cin * a.Lin;
    const uint32_t soff = (uint32_t)u_sub * 128u + (uint32_t)((fw ^ u_sub) << 4) + (uint32_t)c_sub * 4u;   // rows u_sub + 8*i: same chunk XOR
    const int S = p.S, Lin = a.Lin, gstep = 8 * S;
    const bool act = a.act_in == 1, reflect = a.pad_reflect != 0;
    const bool last = u_sub < KT - 1;                          // row 128 + u_sub belongs to the tap halo
    // tiles whose whole input window lies inside the clip (all but the first / last of a clip) load without index checks
    const long long glo = (long long)t0 * S - a.padL;
    const bool interior = glo >= 0 && glo + (long long)(XW - 1) * S + (S - 1) < Lin;
    const int gbase = (t0 + u_sub) * S - a.padL;
    float v[17], vn[17];
    auto load_block = [&](int kb, float (&o)[17]) {
      const int cv = kb * CT_KB + cl;
      const bool cvalid = cv < p.CinV;
      const int ci = cv / S, ph = cv - ci * S;
      const float* xr = xb + (long long)ci * Lin;
      const int g = gbase + ph;
      if (interior) {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = cvalid ? __ldg(xr + g + i * gstep) : 0.f;
        o[16] = (cvalid && last) ? __ldg(xr + g + 16 * gstep) : 0.f;
      } else {
#pragma unroll
        for (int i = 0; i < 17; ++i) {
          int gi = g + i * gstep;
          if (reflect) gi = gi < 0 ? -gi : (gi >= Lin ? 2 * (Lin - 1) - gi : gi);
          const bool ok = cvalid && (i < 16 || last) && (unsigned)gi < (unsigned)Lin;
          o[i] = ok ? __ldg(xr + gi) : 0.f;
        }
      }
    };
    load_block(0, v);
    for (int kb = 0; kb < nkb; ++kb) {
      const int xs = kb % p.XS;
      if (kb + 1 < nkb) load_block(kb + 1, vn);               // in flight while this block is split and stored
      c_mbar_wait(&x_empty[xs], (((uint32_t)(kb / p.XS)) & 1u) ^ 1u);
      uint8_t* dst = smem_gen + (size_t)xs * CT_XSTAGE + soff;
      if (act) {
#pragma unroll
        for (int i = 0; i < 17; ++i) {
          float vv = v[i];
          vv = vv > 0.f ? vv : expm1f(vv);
          const float hi = c_tf32_rna(vv), lo = c_tf32_rna(vv - hi);
          if (i < 16 || last) {
            *reinterpret_cast<float*>(dst + i * 1024) = hi;
            *reinterpret_cast<float*>(dst + CT_XPLANE + i * 1024) = lo;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 17; ++i) {
          const float hi = c_tf32_rna(v[i]), lo = c_tf32_rna(v[i] - hi);
          if (i < 16 || last) {
            *reinterpret_cast<float*>(dst + i * 1024) = hi;
            *reinterpret_cast<float*>(dst + CT_XPLANE + i * 1024) = lo;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) c_mbar_arrive(&x_full[xs]);
#pragma unroll
      for (int i = 0; i < 17; ++i) v[i] = vn[i];
    }
    // ---- epilogue: TMEM -> shared [channel][CT_EPW] (sum of the accumulator chains) -> coalesced stores
    c_mbar_wait(&acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* es = reinterpret_cast<float*>(smem_gen);
    {
      const int q = warp & 3, half = fw >> 2;                 // TMEM lane quadrant of this warp; column half
      const int nhi = nkb < p.NHI ? nkb : p.NHI;
      float* er = es + (q * 32 + lane) * CT_EPW + half * 64;
#pragma unroll 1
      for (int c16 = 0; c16 < 4; ++c16) {
        uint32_t r[16];
        float s[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64 + c16 * 16);
        c_ld16(taddr, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) s[i] = __uint_as_float(r[i]);
        for (int ch = 1; ch <= nhi; ++ch) {                  // the other hi chains, then the chain of the small terms
          c_ld16(taddr + (uint32_t)((ch < nhi ? ch : p.NHI) * CT_N), r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) s[i] += __uint_as_float(r[i]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) er[c16 * 16 + i] = s[i];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int s_il = a.il_s ? a.il_s : 1;
    if (a.il_s == 0) {
      // thread = one position, every second channel: coalesced rows, no index arithmetic beyond a pointer step
      const int pos = tid & (CT_N - 1);
      if (t0 + pos < a.LoutV) {
        const int nchv = min(CT_M, a.CoutV - co0);
        const long long rstep = 2LL * a.LoutV;
        long long idx = ((long long)b * a.CoutV + co0 + (tid >> 7)) * a.LoutV + t0 + pos;
        const float* ep = es + (tid >> 7) * CT_EPW + pos;
#pragma unroll 4
        for (int crl = tid >> 7; crl < nchv; crl += 2, idx += rstep, ep += 2 * CT_EPW) {
          float o = *ep;
          if (a.bias) o += __ldg(a.bias + co0 + crl);
          if (a.res) o += __ldg(a.res + idx);
          a.y[idx] = o;
        }
      }
    } else {
      // virtual channel v = cr*s + ph (phase fastest): real channel cr, output position (t0 + pos)*s + ph - trim
      const int sh = 31 - __clz(s_il);
      const int ow = CT_N << sh, etotal = CT_M * CT_N;
      for (int e = tid; e < etotal; e += 256) {
        const int crl = e >> (7 + sh), op = e & (ow - 1);
        const int pos = op >> sh, ph = op & (s_il - 1);
        const int vch = co0 + crl * s_il + ph;
        if (vch >= a.CoutV || t0 + pos >= a.LoutV) continue;
        const int cr = vch >> sh;
        const long long opos = (long long)(t0 + pos) * s_il + ph - a.il_trim;
        if (opos >= 0 && opos < a.il_lout)
          a.y[((long long)b * a.il_cout + cr) * a.il_lout + opos] = es[(crl * s_il + ph) * CT_EPW + pos] + (a.bias ? __ldg(a.bias + cr) : 0.f);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// wt [(cv*KT + k)][CoutV] (conv_w_transpose_launch) -> planes [2][KT][CoutV][CinV] of rna_tf32 hi / lo
__global__ void codec_tc_split_kernel(const float* __restrict__ wt, float* __restrict__ out, int CinV, int KT, int CoutV) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)CinV * KT * CoutV;
  if (i >= total) return;
  const int cv = (int)(i % CinV);
  const long long r = i / CinV;
  const int col = (int)(r % CoutV), k = (int)(r / CoutV);
  const float v = wt[((long long)cv * KT + k) * CoutV + col];
  const float hi = c_tf32_rna(v);
  out[i] = hi;
  out[total + i] = c_tf32_rna(v - hi);
}

PFN_cuTensorMapEncodeTiled_v12000 ct_get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

}  // namespace

int codec_tc_prepare(const float* wt, int CinV, int KT, int CoutV, float* planes, CodecTcWeights* out, cudaStream_t st) {
  LADIFF_REQUIRE(CinV % 4 == 0, LADIFF_ERR_ARG, "codec_tc_prepare: CinV=%d", CinV);
  const long long total = (long long)CinV * KT * CoutV;
  codec_tc_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(wt, planes, CinV, KT, CoutV);
  LADIFF_CUDA_OK(cudaGetLastError());
  auto enc = ct_get_encode();
  LADIFF_REQUIRE(enc != nullptr, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)CinV, (cuuint64_t)2 * KT * CoutV};
  cuuint64_t strides[1] = {(cuuint64_t)CinV * 4};
  cuuint32_t box[2] = {CT_KB, CT_M};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&out->tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LADIFF_REQUIRE(r == CUDA_SUCCESS, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled(codec W %dx%dx%d) failed: %d", CoutV, KT, CinV, (int)r);
  out->CinV = CinV; out->KT = KT; out->CoutV = CoutV; out->valid = 1;
  return 0;
}

// 1: this conv has no tensor-core form (caller uses the FMA kernel); 0: launched
int codec_tc_launch(const ConvF32Args& a, int B, cudaStream_t st) {
  static const bool off = getenv("LADIFF_CODEC_SIMT") != nullptr;
  const CodecTcWeights* w = a.tcw;
  if (off || !w || !w->valid || a.K % a.stride != 0) return 1;
  const int S = a.stride, KT = a.K / S, CinV = a.Cin * S;
  if (w->CinV != CinV || w->KT != KT || w->CoutV != a.CoutV || KT > 8 || CinV < 32) return 1;
  if (a.il_s != 0 && !(a.il_s == 2 || a.il_s == 4 || a.il_s == 8)) return 1;
  CodecTcParams p;
  memset(&p, 0, sizeof(p));
  p.a = a; p.S = S; p.KT = KT; p.CinV = CinV; p.nkb = cdiv(CinV, CT_KB);
  // Accumulator chains.  The tensor core adds into its fp32 accumulator with truncation: a bias of ~1e-4 abs at the encoder output
  // with a single chain (measured; the FMA kernel: ~1e-5).  The two small products go to their own accumulator (their truncations
  // are 2^-11 smaller), which leaves one truncating add per K step on the main chain: bias / 3.  Convs flagged `tc_precise` (the
  // cond encoder: its output is quantised) with long K loops also rotate the hi*hi products over three accumulators (bias / 9;
  // 512 TMEM columns, so one CTA per SM).
  static const char* env_nhi = getenv("LADIFF_CODEC_TC_NHI");
  p.NHI = (a.tc_precise && p.nkb * KT >= 16) ? 3 : 1;
  if (env_nhi) p.NHI = atoi(env_nhi) == 3 ? 3 : 1;
  // NHI = 1: two CTAs per SM (<= 112 KB each, 256 TMEM columns): the fill of one overlaps the MMAs / prologue / epilogue of the other,
  // which measured better than one CTA with deeper rings on every codec layer (profiles/r2d/codec_tc_stage_sweep.txt); many taps per
  // K block want the second weight stage, few taps the second activation stage.
  if (p.NHI == 1) { p.XS = KT >= 3 ? 1 : 2; p.WS = KT >= 3 ? 2 : 1; }
  else { p.XS = 2; p.WS = 4; }
  static const char* env_cfg = getenv("LADIFF_CODEC_TC_CFG");      // "XS,WS" for experiments
  if (env_cfg) sscanf(env_cfg, "%d,%d", &p.XS, &p.WS);
  p.XS = p.XS < 1 ? 1 : (p.XS > CT_MAXS ? CT_MAXS : p.XS);
  p.WS = p.WS < 1 ? 1 : (p.WS > CT_MAXS ? CT_MAXS : p.WS);
  while ((size_t)p.XS * CT_XSTAGE + (size_t)p.WS * CT_WSTAGE > 216 * 1024) { if (p.WS > 1) --p.WS; else --p.XS; }
  size_t smem = (size_t)p.XS * CT_XSTAGE + (size_t)p.WS * CT_WSTAGE;
  const size_t epi = (size_t)CT_M * CT_EPW * sizeof(float);
  if (epi > smem) smem = epi;
  smem += 1024;
  static unsigned long long attr = 0;
  if (ladiff_first_on_device(&attr))
    LADIFF_CUDA_OK(cudaFuncSetAttribute(codec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  dim3 grid(cdiv(a.LoutV, CT_N), cdiv(a.CoutV, CT_M), B);
  codec_tc_kernel<<<grid, CT_THREADS, smem, st>>>(p, w->tm);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
