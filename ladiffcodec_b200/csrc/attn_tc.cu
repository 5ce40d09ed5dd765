// tcgen05 full attention for the UNet's bottleneck block (reference srcs/modules/unet.py:234-246: sim = q k^T, softmax over keys,
// out = attn v; 4 heads x 32) when the bottleneck is long (whole utterances: n = L/16 in the thousands).
//
// One CTA per (clip, head, 128-query tile), flash-attention style over 128-key tiles:
//   S[128 q, 128 k] = Q K^T      tcgen05.mma kind::f16, M = 128, N = 128, K = 32 (2 instructions), fp32 in TMEM
//   online softmax               thread = query row (tcgen05.ld 32x32b gives a thread its row), running max / sum in registers
//   O_tile[128 q, 32 d] = P V    tcgen05.mma M = 128, N = 32, K = 128 (8 instructions); P written by the row threads as the
//                                16-bit A operand, V^T staged as the B operand; the tile's result is folded into fp32 registers
// Operands live in shared memory in the K-major SWIZZLE_128B layout the conv kernel uses (128-byte rows, 16-byte chunk index XOR
// row%8); they are written by ordinary stores (rows of 64 bytes: half of every 128-byte row is unused), made visible to the tensor
// core with fence.proxy.async.  Warps 0-3: loads / softmax / epilogue (TMEM lane quadrant = warp), warp 4: TMEM allocation + MMA issue.
#include "unet_ops.cuh"

namespace {

__device__ __forceinline__ uint32_t a_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool a_elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void a_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void a_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = a_smem_u32(bar);
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
__device__ __forceinline__ void a_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a_smem_u32(bar)) : "memory");
}
// K-major SWIZZLE_128B matrix descriptor (SBO = 1024 B: 8 rows x 128 B), as in tc_conv.cu
__device__ __forceinline__ uint64_t a_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void a_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void a_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a [rows][128 B] SWIZZLE_128B tile (tile base 1024-byte aligned)
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

constexpr int FT_THREADS = 160;
constexpr uint32_t FT_Q = 0, FT_K = 16384, FT_VT = 32768, FT_P = 40960, FT_SMEM = 73728 + 1024;

__global__ void __launch_bounds__(FT_THREADS) fullattn_tc_kernel(ClView qkv, ClView out, int L) {
  extern __shared__ uint8_t fa_raw[];
  __shared__ __align__(8) uint64_t s_bar, o_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * 128;
  const uint32_t sbase = (a_smem_u32(fa_raw) + 1023u) & ~1023u;
  if (tid == 0) {
    a_mbar_init(&s_bar, 1); a_mbar_init(&o_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_base_s)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();
  const h16* base = qkv.p + (long long)b * qkv.bstride;
  const float scale = 0.17677669529663687f;          // q * 32^-0.5 (unet.py:238)
  // ---- Q tile: thread r owns query q0 + r; zero row beyond the clip
  if (warp < 4) {
    const int qi = q0 + tid;
    uint4 v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = make_uint4(0u, 0u, 0u, 0u);
    if (qi < L) {
      const h16* qr = base + (long long)qi * qkv.pitch + h * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 raw = __ldcg(reinterpret_cast<const uint4*>(qr + 8 * c));
        h162* hp = reinterpret_cast<h162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = h22ff(hp[i]); hp[i] = ff2h2(f.x * scale, f.y * scale); }
        v[c] = raw;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + FT_Q + sw128(tid, c)), "r"(v[c].x), "r"(v[c].y), "r"(v[c].z), "r"(v[c].w) : "memory");
  }
  // instruction descriptors: fp32 accumulate, 16-bit A/B (K-major), N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc_s = (1u << 4) | (TC_IDESC_AB_FMT << 7) | (TC_IDESC_AB_FMT << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_o = (1u << 4) | (TC_IDESC_AB_FMT << 7) | (TC_IDESC_AB_FMT << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  float m_run = -INFINITY, l_run = 0.f, o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
  const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t ph = 0;
  for (int k0 = 0; k0 < L; k0 += 128, ph ^= 1u) {
    if (warp < 4) {
      // ---- K tile (row = key, 32 d) and V^T tile (row = d, 128 keys = 2 atoms of 64)
      const int kj = k0 + tid;
      uint4 kv[4], vv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { kv[c] = make_uint4(0u, 0u, 0u, 0u); vv[c] = make_uint4(0u, 0u, 0u, 0u); }
      if (kj < L) {
        const h16* kr = base + (long long)kj * qkv.pitch + 128 + h * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) { kv[c] = __ldcg(reinterpret_cast<const uint4*>(kr + 8 * c)); vv[c] = __ldcg(reinterpret_cast<const uint4*>(kr + 128 + 8 * c)); }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + FT_K + sw128(tid, c)), "r"(kv[c].x), "r"(kv[c].y), "r"(kv[c].z), "r"(kv[c].w) : "memory");
      const uint32_t atom = (uint32_t)tid >> 6, kc = (uint32_t)tid & 63u;          // key column inside its 64-key atom
      const unsigned short* ve = reinterpret_cast<const unsigned short*>(vv);
#pragma unroll
      for (int d = 0; d < 32; ++d)
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(sbase + FT_VT + atom * 4096u + sw128((uint32_t)d, kc >> 3) + (kc & 7u) * 2u), "h"(ve[d]) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (a_elect_one()) {
        const uint64_t ad = a_desc(sbase + FT_Q), bd = a_desc(sbase + FT_K);
        a_mma(tmem, ad, bd, idesc_s, 0u);
        a_mma(tmem, ad + 2, bd + 2, idesc_s, 1u);
        a_commit(&s_bar);
      }
      __syncwarp();
    } else {
      a_mbar_wait(&s_bar, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- online softmax over this tile's 128 scores of the thread's query row
      const int nvalid = min(128, L - k0);
      float mx = m_run;
#pragma unroll
      for (int cb = 0; cb < 128; cb += 32) {
        uint32_t r0[16], r1[16];
        a_ld16(tlane + (uint32_t)cb, r0); a_ld16(tlane + (uint32_t)cb + 16u, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (cb + i < nvalid) mx = fmaxf(mx, __uint_as_float(r0[i]));
          if (cb + 16 + i < nvalid) mx = fmaxf(mx, __uint_as_float(r1[i]));
        }
      }
      const float corr = __expf(m_run - mx);          // exp(-inf) = 0 on the first tile
      float psum = 0.f;
#pragma unroll
      for (int cb = 0; cb < 128; cb += 32) {
        uint32_t r0[16], r1[16];
        a_ld16(tlane + (uint32_t)cb, r0); a_ld16(tlane + (uint32_t)cb + 16u, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float p[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          p[i] = cb + i < nvalid ? __expf(__uint_as_float(r0[i]) - mx) : 0.f;
          p[16 + i] = cb + 16 + i < nvalid ? __expf(__uint_as_float(r1[i]) - mx) : 0.f;
        }
        // the sum uses the values the tensor core will see (rounded to 16 bits), like a softmax computed in that precision would
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const h162 pr = ff2h2(p[8 * c + 2 * i], p[8 * c + 2 * i + 1]);
            const float2 back = h22ff(pr);
            psum += back.x + back.y;
            w[i] = *reinterpret_cast<const uint32_t*>(&pr);
          }
          const uint32_t chunk = (uint32_t)(cb >> 3) + (uint32_t)c;       // 16-byte chunk (8 keys) of the 128-key row
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                       ::"r"(sbase + FT_P + (chunk >> 3) * 16384u + sw128((uint32_t)tid, chunk & 7u)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        }
      }
      l_run = l_run * corr + psum;
      m_run = mx;
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] *= corr;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (a_elect_one()) {
#pragma unroll
        for (int at = 0; at < 2; ++at) {
          const uint64_t ad = a_desc(sbase + FT_P + (uint32_t)at * 16384u), bd = a_desc(sbase + FT_VT + (uint32_t)at * 4096u);
#pragma unroll
          for (int k = 0; k < 4; ++k) a_mma(tmem + 128u, ad + 2 * k, bd + 2 * k, idesc_o, (at | k) ? 1u : 0u);
        }
        a_commit(&o_bar);
      }
      __syncwarp();
    } else {
      a_mbar_wait(&o_bar, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r0[16], r1[16];
      a_ld16(tlane + 128u, r0); a_ld16(tlane + 144u, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) { o[i] += __uint_as_float(r0[i]); o[16 + i] += __uint_as_float(r1[i]); }
    }
    // the next iteration's stores into K / V^T / P happen after every row thread passed o_bar (all MMAs of this tile are complete)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp < 4) {
    const int qi = q0 + tid;
    if (qi < L) {
      const float inv = 1.f / l_run;
      h16* orow = out.p + (long long)b * out.bstride + (long long)qi * out.pitch + h * 32;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 w;
        h162* hp = reinterpret_cast<h162*>(&w);
#pragma unroll
        for (int i = 0; i < 4; ++i) hp[i] = ff2h2(o[8 * c + 2 * i] * inv, o[8 * c + 2 * i + 1] * inv);
        *reinterpret_cast<uint4*>(orow + 8 * c) = w;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// ================================================================================================ linear-attention tail
// Everything of Residual(PreNorm(LinearAttention)) (unet.py:208-222, 50-56) after the context matrix, in ONE kernel:
//   out[e, n] = sum_d ctx[h][d][e] softmax_d(q[h, :, n])[d]      (SIMT: thread = (head, position), ctx broadcast from shared memory)
//   y = W_out out + b                                            (tcgen05: D[128 positions, C] = A[128, 128] W^T, K = 128)
//   z = LayerNorm_c(y) g + x                                     (row-owning epilogue threads: statistics and the write from TMEM)
// replacing three launches (linattn_out_kernel, the 1x1 to_out conv, layernorm_cl_kernel) and the y / out round trips.
// CTA = 128 positions of one clip; 16 warps: warp w -> head w >> 2, rows 32 (w & 3) + lane; warps 0-3 also run the epilogue
// (TMEM lane quadrant = warp & 3); warp 16 allocates TMEM and issues the MMAs.  C output channels in chunks of 256 (one MMA N);
// C <= 512 stays resident in TMEM for the two LayerNorm sweeps, C = 1024 is computed twice (K is only 128).
constexpr int LT_THREADS = 544;
constexpr uint32_t LT_CTX = 0, LT_A = 16384, LT_W = 49152, LT_PAR = 114688, LT_SMEM = 114688 + 8192 + 1024;

__global__ void __launch_bounds__(LT_THREADS) linattn_tail_kernel(ClView qkv, const float* __restrict__ ctx, const h16* __restrict__ wout,
                                                                   const float* __restrict__ bias, const float* __restrict__ gain, ClView xres,
                                                                   ClView out, int L, int C) {
  extern __shared__ uint8_t lt_raw[];
  __shared__ __align__(8) uint64_t m_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int b = blockIdx.y, n0 = blockIdx.x * 128;
  const uint32_t sbase = (a_smem_u32(lt_raw) + 1023u) & ~1023u;
  uint8_t* sgen = lt_raw + (sbase - a_smem_u32(lt_raw));
  float* cs = reinterpret_cast<float*>(sgen + LT_CTX);              // [4][32][32]
  float* s_bias = reinterpret_cast<float*>(sgen + LT_PAR);           // [C]
  float* s_gain = s_bias + 1024;                                     // [C]
  if (tid == 0) {
    a_mbar_init(&m_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < C; i += LT_THREADS) { s_bias[i] = bias ? __ldg(bias + i) : 0.f; s_gain[i] = __ldg(gain + i); }   // parameters: before the wait
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();
  // ---- phase A: context to shared memory, out = ctx^T softmax(q) as the 16-bit A operand
  if (warp < 16) {
    const float4* src = reinterpret_cast<const float4*>(ctx + (long long)b * 4096);
    float4* dst = reinterpret_cast<float4*>(cs);
    for (int i = tid; i < 1024; i += 512) dst[i] = __ldcg(src + i);
  }
  const int h = warp >> 2, r = (warp & 3) * 32 + lane, n = n0 + r;
  uint4 qraw[4];
  if (warp < 16 && n < L) {
    const h16* qr = qkv.p + (long long)b * qkv.bstride + (long long)n * qkv.pitch + h * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) qraw[i] = __ldcg(reinterpret_cast<const uint4*>(qr + 8 * i));
  }
  __syncthreads();
  if (warp < 16) {
    uint4 w4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w4[i] = make_uint4(0u, 0u, 0u, 0u);
    if (n < L) {
      float q[32];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const h162* hp = reinterpret_cast<const h162*>(&qraw[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = h22ff(hp[j]); q[8 * i + 2 * j] = f.x; q[8 * i + 2 * j + 1] = f.y; }
      }
      float mx = q[0];
#pragma unroll
      for (int i = 1; i < 32; ++i) mx = fmaxf(mx, q[i]);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) { q[i] = __expf(q[i] - mx); sum += q[i]; }
      const float inv = 1.f / sum;
      float acc[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        const float pd = q[d] * inv;
        const float4* c4 = reinterpret_cast<const float4*>(cs + (h * 32 + d) * 32);
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 c = c4[e4];
          acc[4 * e4 + 0] += pd * c.x; acc[4 * e4 + 1] += pd * c.y;
          acc[4 * e4 + 2] += pd * c.z; acc[4 * e4 + 3] += pd * c.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        h162* hp = reinterpret_cast<h162*>(&w4[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) hp[j] = ff2h2(acc[8 * i + 2 * j], acc[8 * i + 2 * j + 1]);
      }
    }
    // channels h*32 .. +31 of row r: atom h >> 1 (64 channels each), 16-byte chunks (h & 1) * 4 + i
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                   ::"r"(sbase + LT_A + (uint32_t)(h >> 1) * 16384u + sw128((uint32_t)r, (uint32_t)((h & 1) * 4 + i))), "r"(w4[i].x), "r"(w4[i].y),
                     "r"(w4[i].z), "r"(w4[i].w) : "memory");
  }
  // ---- phase B: y = W_out out (+ b), LayerNorm statistics, normalise + residual, per half of <= 512 channels
  const uint32_t idesc = (1u << 4) | (TC_IDESC_AB_FMT << 7) | (TC_IDESC_AB_FMT << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
  const int nchunk = C / 256, nhalf = nchunk > 2 ? nchunk / 2 : 1, cph = nchunk / nhalf;     // chunks per half (1 or 2)
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float s1 = 0.f, s2 = 0.f, mean = 0.f, rstd = 0.f;
  uint32_t ph = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int hf = 0; hf < nhalf; ++hf) {
      if (pass == 0 || nhalf > 1) {
        for (int cc = 0; cc < cph; ++cc) {
          const int ch0 = (hf * cph + cc) * 256;
          if (warp < 16) {        // W chunk: rows ch0 .. +255, K = 128 as two 64-wide atoms
            for (int i = tid; i < 4096; i += 512) {
              const int row = i >> 4, c16 = i & 15;
              const uint4 v = __ldg(reinterpret_cast<const uint4*>(wout + (long long)(ch0 + row) * 128 + c16 * 8));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                           ::"r"(sbase + LT_W + (uint32_t)(c16 >> 3) * 32768u + sw128((uint32_t)row, (uint32_t)(c16 & 7))), "r"(v.x), "r"(v.y), "r"(v.z),
                             "r"(v.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncthreads();
          if (warp == 16) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (a_elect_one()) {
#pragma unroll
              for (int at = 0; at < 2; ++at) {
                const uint64_t ad = a_desc(sbase + LT_A + (uint32_t)at * 16384u), bd = a_desc(sbase + LT_W + (uint32_t)at * 32768u);
#pragma unroll
                for (int k = 0; k < 4; ++k) a_mma(tmem + (uint32_t)cc * 256u, ad + 2 * k, bd + 2 * k, idesc, (at | k) ? 1u : 0u);
              }
              a_commit(&m_bar);
            }
            __syncwarp();
          }
          a_mbar_wait(&m_bar, ph);      // every thread: the W buffer may be refilled, the accumulator read
          ph ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
      }
      if (warp < 4) {
        for (int cc = 0; cc < cph; ++cc) {
          const int ch0 = (hf * cph + cc) * 256;
          for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t r0[16], r1[16];
            a_ld16(tlane + (uint32_t)(cc * 256 + c0), r0); a_ld16(tlane + (uint32_t)(cc * 256 + c0 + 16), r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (pass == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float v0 = __uint_as_float(r0[i]) + s_bias[ch0 + c0 + i], v1 = __uint_as_float(r1[i]) + s_bias[ch0 + c0 + 16 + i];
                s1 += v0 + v1; s2 = fmaf(v0, v0, fmaf(v1, v1, s2));
              }
            } else if (n0 + (int)tid < L) {
              const long long rowoff = (long long)(n0 + (int)tid);
              const h16* xr = xres.p + (long long)b * xres.bstride + rowoff * xres.pitch + ch0 + c0;
              h16* orow = out.p + (long long)b * out.bstride + rowoff * out.pitch + ch0 + c0;
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                const uint4 xv = __ldcg(reinterpret_cast<const uint4*>(xr + 8 * g8));
                const h162* xh = reinterpret_cast<const h162*>(&xv);
                uint4 ov;
                h162* oh = reinterpret_cast<h162*>(&ov);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int i0 = 8 * g8 + 2 * j;
                  const float a0 = __uint_as_float(i0 < 16 ? r0[i0 & 15] : r1[i0 & 15]) + s_bias[ch0 + c0 + i0];
                  const float a1 = __uint_as_float(i0 + 1 < 16 ? r0[(i0 + 1) & 15] : r1[(i0 + 1) & 15]) + s_bias[ch0 + c0 + i0 + 1];
                  const float2 xf = h22ff(xh[j]);
                  oh[j] = ff2h2((a0 - mean) * rstd * s_gain[ch0 + c0 + i0] + xf.x, (a1 - mean) * rstd * s_gain[ch0 + c0 + i0 + 1] + xf.y);
                }
                *reinterpret_cast<uint4*>(orow + 8 * g8) = ov;
              }
            }
          }
        }
      }
      if (nhalf > 1) {          // the next half's MMAs overwrite the accumulator: all reads of this half are done
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
      }
    }
    if (pass == 0) {
      mean = s1 / (float)C;
      float var = s2 / (float)C - mean * mean;
      var = var < 0.f ? 0.f : var;
      rstd = rsqrtf(var + 1e-5f);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------ linear attention, second half, on tcgen05
// out[n, h*32 + e] = sum_d softmax_d(q[n, h, :])[d] ctx[h][d][e]   (unet.py:214-221) as D[128 positions, 32 e] = P[128, 32 d] C_h^T per head.
// The SIMT kernel (linattn_out_kernel) spends 1024 FMA + 256 LDS per (position, head); here the thread only does the softmax of its
// (position, head) row and stores it as the A operand.  Both operands are split  v = hi + lo  in the 16-bit type and the product is
// three MMAs (lo*hi + hi*lo + hi*hi, fp32 in TMEM): the result keeps the accuracy of the fp32 kernel (ctx and the probabilities are
// not rounded to 11 bits), at 6 K=16 instructions per head.
// CTA = 128 positions of one clip; warp w < 16: head w >> 2, rows 32 (w & 3) + lane (= its TMEM lane quadrant); warp 16: TMEM + MMA.
// Shared memory (SWIZZLE_128B K-major, two 128-byte atoms per row: heads 0-1, heads 2-3): P hi / lo [2][2][128 rows], C^T hi / lo [2][2][32 rows].
constexpr int LO_THREADS = 544;
constexpr uint32_t LO_A = 0, LO_B = 65536, LO_SMEM = 65536 + 16384 + 1024;

__global__ void __launch_bounds__(LO_THREADS) linattn_out_tc_kernel(ClView qkv, const float* __restrict__ ctx, ClView out, int L) {
  extern __shared__ uint8_t lo_raw[];
  __shared__ __align__(8) uint64_t m_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int b = blockIdx.y, n0 = blockIdx.x * 128;
  const uint32_t sbase = (a_smem_u32(lo_raw) + 1023u) & ~1023u;
  if (tid == 0) {
    a_mbar_init(&m_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_base_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();
  const int h = warp >> 2, r = (warp & 3) * 32 + lane, n = n0 + r;
  if (warp < 16) {
    uint4 qraw[4];
    if (n < L) {
      const h16* qr = qkv.p + (long long)b * qkv.bstride + (long long)n * qkv.pitch + h * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) qraw[i] = __ldcg(reinterpret_cast<const uint4*>(qr + 8 * i));
    }
    {
      // C_h^T: thread = (head, e, eight d): B operand row e, bytes (head & 1) * 64 + 2 d
      const int hh = tid >> 7, e = (tid >> 2) & 31, oct = tid & 3;
      const float* cp = ctx + (long long)b * 4096 + hh * 1024 + (oct * 8) * 32 + e;
      float c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) c[j] = __ldcg(cp + j * 32);
      uint4 hi4, lo4;
      h162* hp = reinterpret_cast<h162*>(&hi4);
      h162* lp = reinterpret_cast<h162*>(&lo4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const h162 hv = ff2h2(c[2 * j], c[2 * j + 1]);
        const float2 hf = h22ff(hv);
        hp[j] = hv;
        lp[j] = ff2h2(c[2 * j] - hf.x, c[2 * j + 1] - hf.y);
      }
      const uint32_t off = (uint32_t)(hh >> 1) * 4096u + sw128((uint32_t)e, (uint32_t)((hh & 1) * 4 + oct));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + LO_B + off), "r"(hi4.x), "r"(hi4.y), "r"(hi4.z), "r"(hi4.w) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + LO_B + 8192u + off), "r"(lo4.x), "r"(lo4.y), "r"(lo4.z), "r"(lo4.w) : "memory");
    }
    uint4 ph4[4], pl4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ph4[i] = make_uint4(0u, 0u, 0u, 0u); pl4[i] = make_uint4(0u, 0u, 0u, 0u); }
    if (n < L) {
      float q[32];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const h162* hp = reinterpret_cast<const h162*>(&qraw[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = h22ff(hp[j]); q[8 * i + 2 * j] = f.x; q[8 * i + 2 * j + 1] = f.y; }
      }
      float mx = q[0];
#pragma unroll
      for (int i = 1; i < 32; ++i) mx = fmaxf(mx, q[i]);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) { q[i] = __expf(q[i] - mx); sum += q[i]; }
      const float inv = 1.f / sum;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        h162* hp = reinterpret_cast<h162*>(&ph4[i]);
        h162* lp = reinterpret_cast<h162*>(&pl4[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p0 = q[8 * i + 2 * j] * inv, p1 = q[8 * i + 2 * j + 1] * inv;
          const h162 hv = ff2h2(p0, p1);
          const float2 hf = h22ff(hv);
          hp[j] = hv;
          lp[j] = ff2h2(p0 - hf.x, p1 - hf.y);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t off = (uint32_t)(h >> 1) * 16384u + sw128((uint32_t)r, (uint32_t)((h & 1) * 4 + i));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + LO_A + off), "r"(ph4[i].x), "r"(ph4[i].y), "r"(ph4[i].z), "r"(ph4[i].w) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + LO_A + 32768u + off), "r"(pl4[i].x), "r"(pl4[i].y), "r"(pl4[i].z), "r"(pl4[i].w) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (a_elect_one()) {
      const uint32_t idesc = (1u << 4) | (TC_IDESC_AB_FMT << 7) | (TC_IDESC_AB_FMT << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        const uint32_t ao = sbase + LO_A + (uint32_t)(hh >> 1) * 16384u + (uint32_t)(hh & 1) * 64u;
        const uint32_t bo = sbase + LO_B + (uint32_t)(hh >> 1) * 4096u + (uint32_t)(hh & 1) * 64u;
        const uint64_t ah = a_desc(ao), al = a_desc(ao + 32768u), bh = a_desc(bo), bl = a_desc(bo + 8192u);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          a_mma(tmem + (uint32_t)hh * 32u, al + 2 * k, bh + 2 * k, idesc, k ? 1u : 0u);
          a_mma(tmem + (uint32_t)hh * 32u, ah + 2 * k, bl + 2 * k, idesc, 1u);
          a_mma(tmem + (uint32_t)hh * 32u, ah + 2 * k, bh + 2 * k, idesc, 1u);
        }
      }
      a_commit(&m_bar);
    }
    __syncwarp();
  }
  a_mbar_wait(&m_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 16) {
    uint32_t r0[16], r1[16];
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)h * 32u;
    a_ld16(taddr, r0); a_ld16(taddr + 16u, r1);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (n < L) {
      h16* orow = out.p + (long long)b * out.bstride + (long long)n * out.pitch + h * 32;
      uint4 o[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        h162* hp0 = reinterpret_cast<h162*>(&o[i]);
        h162* hp1 = reinterpret_cast<h162*>(&o[2 + i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          hp0[j] = ff2h2(__uint_as_float(r0[8 * i + 2 * j]), __uint_as_float(r0[8 * i + 2 * j + 1]));
          hp1[j] = ff2h2(__uint_as_float(r1[8 * i + 2 * j]), __uint_as_float(r1[8 * i + 2 * j + 1]));
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(orow + 8 * i) = o[i];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

}  // namespace

int fullattn_tc_launch(ClView qkv, ClView out, int B, int L, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (ladiff_first_on_device(&attr)) {
    LADIFF_CUDA_OK(cudaFuncSetAttribute(fullattn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FT_SMEM));
    prefer_max_smem_carveout(fullattn_tc_kernel);
  }
  LADIFF_CUDA_OK(launch_pdl(fullattn_tc_kernel, dim3(cdiv(L, 128), 4, B), dim3(FT_THREADS), (size_t)FT_SMEM, st, qkv, out, L));
  return 0;
}

// wout: the to_out 1x1 conv's packed 16-bit weights [C][128]; bias [C] or null; gain [C] = to_out.1.g
int linattn_tail_launch(ClView qkv, const float* ctx, const h16* wout, const float* bias, const float* gain, ClView xres, ClView out, int B,
                        int L, int C, cudaStream_t st) {
  LADIFF_REQUIRE(C % 256 == 0 && C <= 1024 && qkv.pitch % 8 == 0 && xres.pitch % 8 == 0 && out.pitch % 8 == 0, LADIFF_ERR_ARG,
                 "linattn_tail: C=%d", C);
  static unsigned long long attr = 0;
  if (ladiff_first_on_device(&attr)) {
    LADIFF_CUDA_OK(cudaFuncSetAttribute(linattn_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM));
    prefer_max_smem_carveout(linattn_tail_kernel);
  }
  LADIFF_CUDA_OK(launch_pdl(linattn_tail_kernel, dim3(cdiv(L, 128), B), dim3(LT_THREADS), (size_t)LT_SMEM, st, qkv, ctx, wout, bias, gain, xres,
                            out, L, C));
  return 0;
}

int linattn_out_tc_launch(ClView qkv, const float* ctx, ClView out, int B, int L, cudaStream_t st) {
  LADIFF_REQUIRE(qkv.pitch % 8 == 0 && out.pitch % 8 == 0, LADIFF_ERR_ARG, "linattn_out_tc: pitch %d %d", qkv.pitch, out.pitch);
  static unsigned long long attr = 0;
  if (ladiff_first_on_device(&attr)) {
    LADIFF_CUDA_OK(cudaFuncSetAttribute(linattn_out_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LO_SMEM));
    prefer_max_smem_carveout(linattn_out_tc_kernel);
  }
  LADIFF_CUDA_OK(launch_pdl(linattn_out_tc_kernel, dim3(cdiv(L, 128), B), dim3(LO_THREADS), (size_t)LO_SMEM, st, qkv, ctx, out, L));
  return 0;
}
