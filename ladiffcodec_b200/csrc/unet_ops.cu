// Element-wise / normalisation / attention kernels of the UNet (channels-last h16, fp32 math).
#include <stdlib.h>

#include "unet_ops.cuh"

namespace {

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// SiLU(2h) = h + h*tanh(h): one MUFU (tanh.approx.f32, rel. error 2^-11 -> <= 2^-12 of the result, below the h16 rounding of
// the stored activation) instead of ex2 + rcp; callers fold the factor 1/2 into their affine transform
__device__ __forceinline__ float silu_from_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const h162* h = reinterpret_cast<const h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = h22ff(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  h162* h = reinterpret_cast<h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = ff2h2(f[2 * i], f[2 * i + 1]);
  return u;
}

// ------------------------------------------------------------------ GroupNorm apply
// grid (ceil(L / rows_per_cta), B), 256 threads; every thread owns ONE 8-channel chunk for the whole kernel, so its fused
// scale/shift (gamma * rstd * (film_scale + 1), ...) live in registers.  The kernel is latency-bound, so its dependent load
// chains are flattened: the parameters (gamma, beta, FiLM row of step t: written long before the previous kernel) are
// fetched BEFORE the programmatic-dependency wait, i.e. while the producing conv is still draining; after the wait the
// first batch of rows and the GroupNorm partials are requested together.  Warp g reduces the conv epilogue's partial sums
// of group g (fixed lane assignment + xor-shuffle tree: deterministic).
constexpr int GN_U = 4;   // rows per thread per batch (all loads of a batch in flight at once)

__global__ void __launch_bounds__(256) gn_apply_kernel(GnApplyArgs a, int rows_per_cta) {
  const int C = a.y.C, Cg = C / 8, spg = Cg / 32, nslots = C / 32, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cv = C / 8;                       // 8-channel chunks per row: 32, 64 or 128
  const int c0 = (threadIdx.x % cv) * 8, rsub = threadIdx.x / cv, rstep = 256 / cv;
  __shared__ float s_mean[8], s_rstd[8];
  __shared__ float2 ln_part[2][GN_U][8];      // fused LayerNorm: per-warp (sum, sumsq) of a row segment, double-buffered by batch
  float ga[8], be[8], fsc[8], fsh[8], lng[8];
  if (a.ln_g) {
    const float4 l0 = __ldg(reinterpret_cast<const float4*>(a.ln_g + c0)), l1 = __ldg(reinterpret_cast<const float4*>(a.ln_g + c0) + 1);
    lng[0] = l0.x; lng[1] = l0.y; lng[2] = l0.z; lng[3] = l0.w; lng[4] = l1.x; lng[5] = l1.y; lng[6] = l1.z; lng[7] = l1.w;
  }
  {
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma + c0);
    const float4* b4 = reinterpret_cast<const float4*>(a.beta + c0);
    const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
    if (a.film) {
      const float* film = a.film + (long long)a.t_dev[b] * a.film_stride + c0;
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(film)), s1 = __ldg(reinterpret_cast<const float4*>(film) + 1);
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(film + C)), h1 = __ldg(reinterpret_cast<const float4*>(film + C) + 1);
      fsc[0] = s0.x; fsc[1] = s0.y; fsc[2] = s0.z; fsc[3] = s0.w; fsc[4] = s1.x; fsc[5] = s1.y; fsc[6] = s1.z; fsc[7] = s1.w;
      fsh[0] = h0.x; fsh[1] = h0.y; fsh[2] = h0.z; fsh[3] = h0.w; fsh[4] = h1.x; fsh[5] = h1.y; fsh[6] = h1.z; fsh[7] = h1.w;
    }
  }
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, a.L);
  const h16* yb = a.y.p + (long long)b * a.y.bstride + c0;
  const h16* rb = a.res.p ? a.res.p + (long long)b * a.res.bstride + c0 : nullptr;
  h16* ob = a.out.p + (long long)b * a.out.bstride + c0;
  pdl_wait();
  pdl_trigger();
  uint4 u[GN_U], ur[GN_U];
  int row = r0 + rsub;
  auto load_batch = [&](int rw) {
#pragma unroll
    for (int k = 0; k < GN_U; ++k) {
      const int r = rw + k * rstep;
      if (r < r1) {
        u[k] = __ldcg(reinterpret_cast<const uint4*>(yb + (long long)r * a.y.pitch));
        if (rb) ur[k] = __ldcg(reinterpret_cast<const uint4*>(rb + (long long)r * a.res.pitch));
      }
    }
  };
  load_batch(row);
  {
    float s1 = 0.f, s2 = 0.f;
    const int n = a.n_ntiles * spg;
    const float2* base = a.stats + (long long)b * a.n_ntiles * nslots + warp * spg;
    for (int i = lane; i < n; i += 32) {
      const int pt = i / spg, s = i - pt * spg;
      const float2 v = __ldcg(base + pt * nslots + s);
      s1 += v.x; s2 += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      const float cnt = (float)Cg * (float)a.L;
      const float m = s1 / cnt;
      float var = s2 / cnt - m * m;
      var = var < 0.f ? 0.f : var;
      s_mean[warp] = m;
      s_rstd[warp] = rsqrtf(var + 1e-5f);
    }
  }
  __syncthreads();
  float sa[8], sb[8];
  {
    const float mean = s_mean[c0 / Cg], rstd = s_rstd[c0 / Cg];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g = ga[j] * rstd;
      float bb = be[j] - mean * g;
      if (a.film) {
        const float sc = fsc[j] + 1.f;
        g *= sc;
        bb = bb * sc + fsh[j];
      }
      sa[j] = 0.5f * g; sb[j] = 0.5f * bb;      // half-scale: silu_from_half
    }
  }
  if (!a.ln_g) {
    for (;;) {
#pragma unroll
      for (int k = 0; k < GN_U; ++k) {
        const int r = row + k * rstep;
        if (r >= r1) break;
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = silu_from_half(fmaf(f[j], sa[j], sb[j]));
        if (rb) {
          float rr[8];
          unpack8(ur[k], rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += rr[j];
        }
        if (a.do_tanh) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = tanhf(f[j]);
        }
        *reinterpret_cast<uint4*>(ob + (long long)r * a.out.pitch) = pack8(f);
      }
      row += GN_U * rstep;
      if (row >= r1) break;
      load_batch(row);
    }
    return;
  }
  // ---- with the attention pre-norm fused: the cv threads that share a row (cv/32 whole warps) reduce (sum, sumsq) of the
  // h16-rounded result, so LN sees exactly the tensor the unfused kernel would read back.  Uniform trip count per CTA.
  const int wpr = cv >> 5;                                 // warps per row
  h16* lb = a.ln_out.p + (long long)b * a.ln_out.bstride + c0;
  const int nbatch = (r1 - r0 + GN_U * rstep - 1) / (GN_U * rstep);
  for (int it = 0; it < nbatch; ++it) {
    uint4 o[GN_U];
    float2 st[GN_U];
#pragma unroll
    for (int k = 0; k < GN_U; ++k) {
      const int r = row + k * rstep;
      st[k] = make_float2(0.f, 0.f);
      o[k] = make_uint4(0u, 0u, 0u, 0u);
      if (r < r1) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = silu_from_half(fmaf(f[j], sa[j], sb[j]));
        if (rb) {
          float rr[8];
          unpack8(ur[k], rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += rr[j];
        }
        o[k] = pack8(f);
        *reinterpret_cast<uint4*>(ob + (long long)r * a.out.pitch) = o[k];
        unpack8(o[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { st[k].x += f[j]; st[k].y = fmaf(f[j], f[j], st[k].y); }
      }
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) {
        st[k].x += __shfl_xor_sync(0xffffffffu, st[k].x, sh);
        st[k].y += __shfl_xor_sync(0xffffffffu, st[k].y, sh);
      }
      if (wpr > 1 && lane == 0) ln_part[it & 1][k][warp] = st[k];
    }
    if (wpr > 1) __syncthreads();
    const int rnext = row + GN_U * rstep;
    if (it + 1 < nbatch) load_batch(rnext);                // next batch in flight during the normalisation
#pragma unroll
    for (int k = 0; k < GN_U; ++k) {
      const int r = row + k * rstep;
      if (r >= r1) continue;
      float2 t = st[k];
      if (wpr > 1) {
        const int w0 = (warp / wpr) * wpr;
        t = make_float2(0.f, 0.f);
        for (int w = w0; w < w0 + wpr; ++w) { const float2 v = ln_part[it & 1][k][w]; t.x += v.x; t.y += v.y; }
      }
      const float mean = t.x / (float)C;
      float var = t.y / (float)C - mean * mean;
      var = var < 0.f ? 0.f : var;
      const float rstd = rsqrtf(var + 1e-5f);
      float f[8];
      unpack8(o[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * lng[j];
      *reinterpret_cast<uint4*>(lb + (long long)r * a.ln_out.pitch) = pack8(f);
    }
    row = rnext;
  }
}

// ------------------------------------------------------------------ channel LayerNorm: each warp normalises LN_R rows at once
// (all of their loads are issued before the first reduction: the kernel is latency-bound, not bandwidth-bound)
template <int NVEC>
__global__ void __launch_bounds__(256) layernorm_cl_kernel(ClView x, const float* __restrict__ g, ClView res, ClView out, int L, int B) {
  constexpr int LN_R = NVEC >= 4 ? 1 : (NVEC >= 2 ? 2 : 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int C = NVEC * 256;
  float gg[NVEC][8];
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + (lane + 32 * i) * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + (lane + 32 * i) * 8 + 4));
    gg[i][0] = g0.x; gg[i][1] = g0.y; gg[i][2] = g0.z; gg[i][3] = g0.w;
    gg[i][4] = g1.x; gg[i][5] = g1.y; gg[i][6] = g1.z; gg[i][7] = g1.w;
  }
  pdl_wait();       // the gains above are parameters: fetched while the producing kernel drains
  pdl_trigger();
  const int gpc = (L + LN_R - 1) / LN_R, total = gpc * B;      // row groups per clip / in all
  for (int rg = blockIdx.x * 8 + warp; rg < total; rg += gridDim.x * 8) {
    const int b = rg / gpc, row0 = (rg - b * gpc) * LN_R;
    uint4 raw[LN_R][NVEC], rres[LN_R][NVEC];
#pragma unroll
    for (int r = 0; r < LN_R; ++r) {
      if (row0 + r >= L) break;
      const h16* xr = x.p + (long long)b * x.bstride + (long long)(row0 + r) * x.pitch;
#pragma unroll
      for (int i = 0; i < NVEC; ++i) raw[r][i] = __ldcg(reinterpret_cast<const uint4*>(xr + (lane + 32 * i) * 8));
      if (res.p) {
        const h16* rr = res.p + (long long)b * res.bstride + (long long)(row0 + r) * res.pitch;
#pragma unroll
        for (int i = 0; i < NVEC; ++i) rres[r][i] = __ldcg(reinterpret_cast<const uint4*>(rr + (lane + 32 * i) * 8));
      }
    }
#pragma unroll
    for (int r = 0; r < LN_R; ++r) {
      if (row0 + r >= L) break;
      float v[NVEC][8];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NVEC; ++i) {
        unpack8(raw[r][i], v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i][j];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / (float)C;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NVEC; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q += d * d; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q / (float)C + 1e-5f);
      h16* orow = out.p + (long long)b * out.bstride + (long long)(row0 + r) * out.pitch;
#pragma unroll
      for (int i = 0; i < NVEC; ++i) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (v[i][j] - mean) * rstd * gg[i][j];
        if (res.p) {
          float rr[8];
          unpack8(rres[r][i], rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += rr[j];
        }
        *reinterpret_cast<uint4*>(orow + (lane + 32 * i) * 8) = pack8(f);
      }
    }
  }
}

// ------------------------------------------------------------------ linear attention (unet.py:208-221)
// ctx[b][h][d][e] = 32^-1/2 * sum_n softmax_n(k)[d,n] v[e,n].
// grid (nsplit, 4, B), 256 threads: each CTA reduces a segment of <= LA_S positions held in shared memory (exact
// two-pass softmax inside the segment) and writes (max, sum, ctx) partials; the last CTA of a (clip, head) to arrive
// (ticket counter) merges the partials in segment order, so the result does not depend on CTA scheduling.
constexpr int LA_S = 160;
constexpr int LA_PART = 64 + 1024;   // floats per partial: m[32], z[32], ctx[32][32]

__global__ void __launch_bounds__(256) linattn_ctx_kernel(ClView qkv, float* __restrict__ ctx, float* part, int* counters, int L, int S,
                                                          int nsplit) {
  pdl_wait();
  pdl_trigger();
  const int sp = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __shared__ __align__(16) float ks[LA_S][32];
  __shared__ __align__(16) float vs[LA_S][32];
  __shared__ __align__(16) float red[8][32];
  __shared__ float m_s[32], z_s[32];
  __shared__ int is_last;
  const int n_lo = sp * S, cnt = min(L, n_lo + S) - n_lo;
  {
    // 8 threads per row: 4 x 16 B of k, 4 x 16 B of v; all (<= LA_S/32) row loads of a thread are issued before the first use
    const int c = tid & 7;
    const h16* base = qkv.p + (long long)b * qkv.bstride + (c < 4 ? 128 : 256) + h * 32 + (c & 3) * 8;
    float (*dst)[32] = c < 4 ? ks : vs;
    constexpr int NB = LA_S / 32;
    uint4 raw[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const int i = (tid >> 3) + 32 * k;
      if (i < cnt) raw[k] = __ldcg(reinterpret_cast<const uint4*>(base + (long long)(n_lo + i) * qkv.pitch));
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const int i = (tid >> 3) + 32 * k;
      if (i < cnt) {
        float f[8];
        unpack8(raw[k], f);
        float4* d4 = reinterpret_cast<float4*>(&dst[i][(c & 3) * 8]);
        d4[0] = make_float4(f[0], f[1], f[2], f[3]);
        d4[1] = make_float4(f[4], f[5], f[6], f[7]);
      }
    }
  }
  __syncthreads();
  float m = -INFINITY;
  for (int i = warp; i < cnt; i += 8) m = fmaxf(m, ks[i][lane]);
  red[warp][lane] = m;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w][lane]);
    m_s[lane] = m;
  }
  __syncthreads();
  m = m_s[lane];
  float z = 0.f;
  for (int i = warp; i < cnt; i += 8) {
    const float pe = __expf(ks[i][lane] - m);
    ks[i][lane] = pe;
    z += pe;
  }
  __syncthreads();          // all of red[] has been read by warp 0 and every p is in place
  red[warp][lane] = z;
  __syncthreads();
  if (warp == 0) {
    float zz = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) zz += red[w][lane];
    z_s[lane] = zz;
  }
  // ctx partial: 4 row groups x 64 threads, each thread a 4 (d) x 4 (e) register tile -> 2 LDS.128 per 16 FMA
  const int rg = tid >> 6, tt = tid & 63, d4 = (tt >> 3) * 4, e4 = (tt & 7) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int i = rg; i < cnt; i += 4) {
    const float4 pe = *reinterpret_cast<const float4*>(&ks[i][d4]);
    const float4 v = *reinterpret_cast<const float4*>(&vs[i][e4]);
    acc[0][0] += pe.x * v.x; acc[0][1] += pe.x * v.y; acc[0][2] += pe.x * v.z; acc[0][3] += pe.x * v.w;
    acc[1][0] += pe.y * v.x; acc[1][1] += pe.y * v.y; acc[1][2] += pe.y * v.z; acc[1][3] += pe.y * v.w;
    acc[2][0] += pe.z * v.x; acc[2][1] += pe.z * v.y; acc[2][2] += pe.z * v.z; acc[2][3] += pe.z * v.w;
    acc[3][0] += pe.w * v.x; acc[3][1] += pe.w * v.y; acc[3][2] += pe.w * v.z; acc[3][3] += pe.w * v.w;
  }
  __syncthreads();          // everyone is done with ks/vs: row groups 1..3 park their tiles in vs, group 0 sums them in order
  float* pool = &vs[0][0];  // [3][32][32]
  if (rg > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(pool + ((rg - 1) * 32 + d4 + i) * 32 + e4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;  // 32^-0.5 (q * scale, unet.py:216)
  if (rg == 0) {
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 o = *reinterpret_cast<const float4*>(pool + (g * 32 + d4 + i) * 32 + e4);
        acc[i][0] += o.x; acc[i][1] += o.y; acc[i][2] += o.z; acc[i][3] += o.w;
      }
    if (nsplit == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float inv = scale / z_s[d4 + i];
        *reinterpret_cast<float4*>(ctx + (((long long)b * 4 + h) * 32 + d4 + i) * 32 + e4) =
            make_float4(acc[i][0] * inv, acc[i][1] * inv, acc[i][2] * inv, acc[i][3] * inv);
      }
    } else {
      float* pp = part + (((long long)b * 4 + h) * nsplit + sp) * LA_PART;
      if (tt < 32) { pp[tt] = m_s[tt]; pp[32 + tt] = z_s[tt]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(pp + 64 + (d4 + i) * 32 + e4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
  if (nsplit == 1) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(&counters[b * 4 + h], 1) == nsplit - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // the last CTA of this (clip, head) merges the segment partials in segment order (independent of CTA scheduling)
  const int d = tid >> 3, ee = (tid & 7) * 4;
  const float* p0 = part + ((long long)b * 4 + h) * nsplit * LA_PART;
  float M = -INFINITY;
  for (int s2 = 0; s2 < nsplit; ++s2) M = fmaxf(M, __ldcg(p0 + (long long)s2 * LA_PART + d));
  float Z = 0.f;
  float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s2 = 0; s2 < nsplit; ++s2) {
    const float* ps = p0 + (long long)s2 * LA_PART;
    const float w = __expf(__ldcg(ps + d) - M);
    Z += w * __ldcg(ps + 32 + d);
    const float4 c = __ldcg(reinterpret_cast<const float4*>(ps + 64 + d * 32 + ee));
    a4.x += w * c.x; a4.y += w * c.y; a4.z += w * c.z; a4.w += w * c.w;
  }
  const float inv = scale / Z;
  *reinterpret_cast<float4*>(ctx + (((long long)b * 4 + h) * 32 + d) * 32 + ee) = make_float4(a4.x * inv, a4.y * inv, a4.z * inv, a4.w * inv);
  if (tid == 0) counters[b * 4 + h] = 0;     // ready for the next launch on this stream
}

// out[b][n][h*32+e] = sum_d ctx[h][d][e] softmax_d(q[n,h,:])[d]      grid (ceil(L/64), B), 256 threads
// warp w: head w&3, rows 32*(w>>2) + lane; one (row, head) per thread: q row in registers (requested before the ctx tile
// is staged, so both round trips overlap), ctx broadcast from smem
__global__ void __launch_bounds__(256, 3) linattn_out_kernel(ClView qkv, const float* __restrict__ ctx, ClView out, int L) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ __align__(16) float cs[4][32][32];
  const int h = warp & 3, n = blockIdx.x * 64 + (warp >> 2) * 32 + lane;
  pdl_wait();
  pdl_trigger();
  uint4 qraw[4];
  if (n < L) {
    const h16* qr = qkv.p + (long long)b * qkv.bstride + (long long)n * qkv.pitch + h * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) qraw[i] = __ldcg(reinterpret_cast<const uint4*>(qr + 8 * i));
  }
  {
    const float4* src = reinterpret_cast<const float4*>(ctx + (long long)b * 4096);
    float4* dst = reinterpret_cast<float4*>(&cs[0][0][0]);
    float4 t4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t4[i] = __ldcg(src + threadIdx.x + 256 * i);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[threadIdx.x + 256 * i] = t4[i];
  }
  __syncthreads();
  if (n >= L) return;
  float q[32];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float f[8];
    unpack8(qraw[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) q[8 * i + j] = f[j];
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
    const float4* c4 = reinterpret_cast<const float4*>(&cs[h][d][0]);
#pragma unroll
    for (int e4 = 0; e4 < 8; ++e4) {
      const float4 c = c4[e4];
      acc[4 * e4 + 0] += pd * c.x; acc[4 * e4 + 1] += pd * c.y;
      acc[4 * e4 + 2] += pd * c.z; acc[4 * e4 + 3] += pd * c.w;
    }
  }
  h16* orow = out.p + (long long)b * out.bstride + (long long)n * out.pitch + h * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = acc[8 * i + j];
    *reinterpret_cast<uint4*>(orow + 8 * i) = pack8(f);
  }
}

// ------------------------------------------------------------------ full attention (mid block), online softmax
// grid (ceil(n/32), 4, B), 256 threads: each warp owns 4 queries; keys/values staged in smem tiles of 128
__global__ void __launch_bounds__(256) fullattn_kernel(ClView qkv, ClView out, int L) {
  pdl_wait();
  pdl_trigger();
  const int h = blockIdx.y, b = blockIdx.z, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KT = 128;
  __shared__ float ks[KT][33];
  __shared__ float vs[KT][33];
  const h16* base = qkv.p + (long long)b * qkv.bstride;
  const float scale = 0.17677669529663687f;
  float qreg[4][32];
  float m[4], l[4], acc[4];
  int qi[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    qi[u] = blockIdx.x * 32 + warp * 4 + u;
    m[u] = -INFINITY; l[u] = 0.f; acc[u] = 0.f;
    float qv = 0.f;
    if (qi[u] < L) qv = h2f(base[(long long)qi[u] * qkv.pitch + h * 32 + lane]) * scale;
#pragma unroll
    for (int d = 0; d < 32; ++d) qreg[u][d] = __shfl_sync(0xffffffffu, qv, d);
  }
  for (int k0 = 0; k0 < L; k0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * 32; i += 256) {
      const int j = i >> 5, d = i & 31;
      float kk = 0.f, vv = 0.f;
      if (k0 + j < L) {
        kk = h2f(base[(long long)(k0 + j) * qkv.pitch + 128 + h * 32 + d]);
        vv = h2f(base[(long long)(k0 + j) * qkv.pitch + 256 + h * 32 + d]);
      }
      ks[j][d] = kk; vs[j][d] = vv;
    }
    __syncthreads();
    const int kmax = min(KT, L - k0);
    for (int kb = 0; kb < kmax; kb += 32) {
      const int j = kb + lane;
      const bool valid = j < kmax;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (qi[u] >= L) continue;     // warp-uniform
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) s += qreg[u][d] * ks[j][d];
        s = valid ? s : -INFINITY;
        float mx = s;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float mn = fmaxf(m[u], mx);
        const float p = valid ? expf(s - mn) : 0.f;
        const float corr = expf(m[u] - mn);
        float ps = p;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
        l[u] = l[u] * corr + ps;
        float a = acc[u] * corr;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) a += __shfl_sync(0xffffffffu, p, jj) * vs[kb + jj][lane];
        acc[u] = a;
        m[u] = mn;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (qi[u] < L)
      out.p[(long long)b * out.bstride + (long long)qi[u] * out.pitch + h * 32 + lane] = f2h(acc[u] / l[u]);
}

// ------------------------------------------------------------------ full attention, keys resident in shared memory
// grid (ceil(n/128), 4, B), 128 threads: one query per thread (registers), K and V of the (clip, head) staged once as fp32;
// exact two-pass softmax (scores recomputed in the second pass instead of stored).  Used when n <= FA2_MAX_KEYS.
constexpr int FA2_MAX_KEYS = 512;
__global__ void __launch_bounds__(128) fullattn2_kernel(ClView qkv, ClView out, int L) {
  extern __shared__ __align__(16) float fa_sm[];
  float* ks = fa_sm;                 // [L][32]
  float* vs = fa_sm + (size_t)L * 32;
  const int h = blockIdx.y, b = blockIdx.z;
  pdl_wait();
  pdl_trigger();
  const h16* base = qkv.p + (long long)b * qkv.bstride;
  for (int i = threadIdx.x; i < L * 8; i += 128) {          // 8 x 16 B per key row: k (4) then v (4)
    const int j = i >> 3, c = i & 7;
    float f[8];
    unpack8(__ldcg(reinterpret_cast<const uint4*>(base + (long long)j * qkv.pitch + (c < 4 ? 128 : 256) + h * 32 + (c & 3) * 8)), f);
    float4* d4 = reinterpret_cast<float4*>((c < 4 ? ks : vs) + (size_t)j * 32 + (c & 3) * 8);
    d4[0] = make_float4(f[0], f[1], f[2], f[3]);
    d4[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  const int qi = blockIdx.x * 128 + threadIdx.x;
  float q[32];
  if (qi < L) {
    const h16* qr = base + (long long)qi * qkv.pitch + h * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f[8];
      unpack8(__ldcg(reinterpret_cast<const uint4*>(qr + 8 * i)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) q[8 * i + j] = f[j] * 0.17677669529663687f;     // q * 32^-0.5 (unet.py:238)
    }
  }
  __syncthreads();
  if (qi >= L) return;
  auto score = [&](int j) {
    const float4* k4 = reinterpret_cast<const float4*>(ks + (size_t)j * 32);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float4 a = k4[i], c = k4[i + 1];
      s0 += q[4 * i] * a.x; s0 += q[4 * i + 1] * a.y; s0 += q[4 * i + 2] * a.z; s0 += q[4 * i + 3] * a.w;
      s1 += q[4 * i + 4] * c.x; s1 += q[4 * i + 5] * c.y; s1 += q[4 * i + 6] * c.z; s1 += q[4 * i + 7] * c.w;
    }
    return s0 + s1;
  };
  float m = -INFINITY;
  for (int j = 0; j < L; ++j) m = fmaxf(m, score(j));
  float l = 0.f, acc[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) acc[e] = 0.f;
  for (int j = 0; j < L; ++j) {
    const float pj = __expf(score(j) - m);
    l += pj;
    const float4* v4 = reinterpret_cast<const float4*>(vs + (size_t)j * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = v4[i];
      acc[4 * i] += pj * v.x; acc[4 * i + 1] += pj * v.y; acc[4 * i + 2] += pj * v.z; acc[4 * i + 3] += pj * v.w;
    }
  }
  const float inv = 1.f / l;
  h16* orow = out.p + (long long)b * out.bstride + (long long)qi * out.pitch + h * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = acc[8 * i + j] * inv;
    *reinterpret_cast<uint4*>(orow + 8 * i) = pack8(f);
  }
}

// ------------------------------------------------------------------ layout conversion
// NCL f32 -> channels-last h16 (optionally scaled per clip).  grid (ceil(L/32), C/32, B), block (32, 8)
__global__ void ncl_to_cl_kernel(const float* __restrict__ x, const float* __restrict__ inv_scale, ClView out, int C, int L) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
  const float sc = inv_scale ? inv_scale[b] : 1.f;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + threadIdx.x;
    t[i][threadIdx.x] = (l < L) ? x[((long long)b * C + c) * L + l] * sc : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int l = l0 + i;
    if (l < L) out.p[(long long)b * out.bstride + (long long)l * out.pitch + c0 + threadIdx.x] = f2h(t[threadIdx.x][i]);
  }
}

__global__ void cl_to_ncl_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int L) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int l = l0 + i;
    t[i][threadIdx.x] = (l < L) ? x[((long long)b * L + l) * C + c0 + threadIdx.x] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int l = l0 + threadIdx.x;
    if (l < L) y[((long long)b * C + c0 + i) * L + l] = t[threadIdx.x][i];
  }
}

__global__ void __launch_bounds__(1024) absmax_inv_kernel(const float* __restrict__ x, float* __restrict__ inv, long long n, float eps) {
  const float* p = x + (long long)blockIdx.x * n;
  float m = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(p[i]));
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) inv[blockIdx.x] = 1.f / (m + eps);
  }
}

// ------------------------------------------------------------------ DDPM posterior step
// Counter-based noise for the throughput mode: Philox4x32-10 (Salmon et al., SC'11) keyed by the seed, counter =
// (thread's first element index, step); one call yields the four normals (Box-Muller) of the thread's four elements.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = ((float)a + 1.f) * 2.3283064365386963e-10f;     // (0, 1]
  const float u2 = (float)b * 2.3283064365386963e-10f;              // [0, 1)
  const float r = sqrtf(-2.f * __logf(u1));
  float sn, cs;
  __sincosf(6.283185307179586f * u2, &sn, &cs);
  n0 = r * cs; n1 = r * sn;
}

// One sampler step for every element of x [B][C][L] (NCL fp32, in place) given eps [B][L][C] (channels-last fp32):
//   x0  = clamp(a x - b eps, -1, 1)                      predict_start_from_noise + clamp   ddpm_loss.py:175-179, 237-238
//   DDPM (mode 0): x <- (k0 x0 + k1 x) + ks z            q_posterior + p_sample             :199-206, 244-251
//   DDIM (mode 1): x <- (k0 x0 + k1 eps) + ks z          ddim_sample                        :296-300
//   mode 2:        x <- x0                               ddim_sample's last pair (time_next < 0)  :289-291
// Every product and sum is rounded separately (no FMA contraction): given the same eps the step reproduces the reference's
// fp32 tensor algebra bit for bit.  Also writes the new x as h16 into xin (channels-last view, the UNet's next input).
// In-kernel noise (noise == null, ks != 0): Philox4x32-10 keyed by the seed, counter = (global element index of the thread's first
// element [clip_offset + b], absolute timestep t_abs): independent of how a trajectory is split into calls or a job over ranks.
// grid (ceil(L/32), C/32, B), block (32, 8)
__global__ void ddpm_step_kernel(const float* __restrict__ eps, float* __restrict__ x, const float* __restrict__ noise,
                                 unsigned long long seed, int t_abs, unsigned long long clip_offset, StepCoef cf,
                                 ClView xin, int C, int L) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
  const bool want_z = cf.ks != 0.f && cf.mode != 2;
  float ev[4], xv[4], zv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.y + 8 * k, l = l0 + i;
    ev[k] = (l < L) ? __ldcg(eps + ((long long)b * L + l) * C + c0 + threadIdx.x) : 0.f;   // t[l][c]
  }
  const int lx = l0 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + threadIdx.y + 8 * k;
    const long long idx = ((long long)b * C + c) * L + lx;
    xv[k] = (lx < L) ? x[idx] : 0.f;
    if (want_z && noise && lx < L) zv[k] = __ldg(noise + idx);
  }
  if (want_z && !noise) {
    const unsigned long long first = ((clip_offset + (unsigned long long)b) * C + c0 + threadIdx.y) * (unsigned long long)L + lx;
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)first, (uint32_t)(first >> 32), (uint32_t)t_abs, 0x1ad1ffu),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    box_muller(rnd.x, rnd.y, zv[0], zv[1]);
    box_muller(rnd.z, rnd.w, zv[2], zv[3]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) t[threadIdx.y + 8 * k][threadIdx.x] = ev[k];
  __syncthreads();
  float xn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.y + 8 * k;
    const float e = t[threadIdx.x][i];
    float x0 = __fsub_rn(__fmul_rn(cf.a, xv[k]), __fmul_rn(cf.b, e));
    x0 = fminf(fmaxf(x0, -1.f), 1.f);
    if (cf.mode == 2) xn[k] = x0;
    else {
      const float m = __fadd_rn(__fmul_rn(cf.k0, x0), __fmul_rn(cf.k1, cf.mode == 1 ? e : xv[k]));
      xn[k] = want_z ? __fadd_rn(m, __fmul_rn(cf.ks, zv[k])) : (cf.mode == 1 ? __fadd_rn(m, 0.f) : m);
    }
    if (lx < L) x[((long long)b * C + c0 + i) * L + lx] = xn[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) t[threadIdx.x][threadIdx.y + 8 * k] = lx < L ? xn[k] : 0.f;   // t[l][c]
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.y + 8 * k, l = l0 + i;
    if (l < L && xin.p) xin.p[(long long)b * xin.bstride + (long long)l * xin.pitch + c0 + threadIdx.x] = f2h(t[i][threadIdx.x]);
  }
}

// N(0,1) fill (the initial torch.randn(shape) of p_sample_loop / ddim_sample when the caller passes none): same generator,
// counter word 3 = 0x1417 so the stream is disjoint from every step's.  One thread = 4 consecutive elements.
__global__ void randn_fill_kernel(float* __restrict__ x, long long n, unsigned long long seed, unsigned long long elem_offset, int uniform) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  const unsigned long long first = elem_offset + (unsigned long long)i4;
  const uint4 rnd = philox4x32_10(make_uint4((uint32_t)first, (uint32_t)(first >> 32), 0u, 0x1417u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float z[4];
  if (uniform) {
    z[0] = (float)rnd.x * 2.3283064365386963e-10f; z[1] = (float)rnd.y * 2.3283064365386963e-10f;
    z[2] = (float)rnd.z * 2.3283064365386963e-10f; z[3] = (float)rnd.w * 2.3283064365386963e-10f;
  } else {
    box_muller(rnd.x, rnd.y, z[0], z[1]);
    box_muller(rnd.z, rnd.w, z[2], z[3]);
  }
  for (int k = 0; k < 4 && i4 + k < n; ++k) x[i4 + k] = z[k];
}

// q_sample (ddpm_loss.py:387-393): out = sqrt_ac[t_b] x + sqrt_1m_ac[t_b] noise, per clip; products rounded separately
__global__ void q_sample_kernel(const float* __restrict__ x, const float* __restrict__ noise, const int* __restrict__ t_dev,
                                const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1m_ac, float* __restrict__ out, long long n) {
  const int b = blockIdx.y;
  const float a = sqrt_ac[t_dev[b]], s = sqrt_1m_ac[t_dev[b]];
  const long long base = (long long)b * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[base + i] = __fadd_rn(__fmul_rn(a, x[base + i]), __fmul_rn(s, noise[base + i]));
}

// x <- a x + b y  (infilling's (1 - lam) img + lam infill_img, ddpm_loss.py:360,364; scaling by a constant with b = 0)
__global__ void axpby_kernel(float* __restrict__ x, float a, const float* __restrict__ y, float b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float ax = __fmul_rn(a, x[i]);
    x[i] = y ? __fadd_rn(ax, __fmul_rn(b, y[i])) : ax;
  }
}

// p_losses tail (ddpm_loss.py:404-437, loss_type l1, objective pred_noise): per clip b
//   pred_x0 = a_t x_t - r_t eps (no clamp);  part[b] = mean |eps - noise| * p2_weight[t_b]
// eps channels-last [B][L][C], the rest NCL.  grid (B), 1024 threads; fixed-order tree reduction (deterministic).
__global__ void __launch_bounds__(1024) p_losses_kernel(const float* __restrict__ eps, const float* __restrict__ noise, const float* __restrict__ xt,
                                                        const int* __restrict__ t_dev, const float* __restrict__ recip, const float* __restrict__ recipm1,
                                                        const float* __restrict__ p2w, float* __restrict__ pred_x0, float* __restrict__ eps_ncl,
                                                        float* __restrict__ part, int C, int L) {
  const int b = blockIdx.x, ti = t_dev[b];
  const float a = recip[ti], r = recipm1[ti];
  const long long n = (long long)C * L;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const int c = (int)(i / L), l = (int)(i - (long long)c * L);
    const float e = __ldcg(eps + ((long long)b * L + l) * C + c);
    const long long o = (long long)b * n + i;
    acc += fabsf(e - noise[o]);
    if (pred_x0) pred_x0[o] = __fsub_rn(__fmul_rn(a, xt[o]), __fmul_rn(r, e));
    if (eps_ncl) eps_ncl[o] = e;
  }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) part[b] = acc / (float)n * p2w[ti];
  }
}
__global__ void mean_small_kernel(const float* __restrict__ part, float* __restrict__ out, int B) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < B; ++i) s += part[i];
    out[0] = s / (float)B;
  }
}

// ClippedSDR(MultiSrcNegSDR("sdsdr")) per clip (losses_fn.py:54-66; asteroid's published sd-sdr, EPS 1e-8):
//   zero-mean est/target; scaled = <est,tgt> tgt / (|tgt|^2 + EPS); noise = est - tgt;
//   out[b] = max(-10 log10(|scaled|^2 / (|noise|^2 + EPS) + EPS), clip)        grid (B), 1024 threads, two passes
__global__ void __launch_bounds__(1024) sdsdr_kernel(const float* __restrict__ est, const float* __restrict__ tgt, float* __restrict__ out, long long n,
                                                     float clip) {
  const float* e = est + (long long)blockIdx.x * n;
  const float* t = tgt + (long long)blockIdx.x * n;
  __shared__ double red[4][32];
  __shared__ double tot[4];
  auto reduce = [&](double (&v)[4], int cnt) {
    for (int k = 0; k < cnt; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      for (int k = 0; k < cnt; ++k) {
        double w = red[k][threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) tot[k] = w;
      }
    }
    __syncthreads();
  };
  double v[4] = {0, 0, 0, 0};
  for (long long i = threadIdx.x; i < n; i += 1024) { v[0] += e[i]; v[1] += t[i]; }
  reduce(v, 2);
  const float me = (float)(tot[0] / (double)n), mt = (float)(tot[1] / (double)n);
  __syncthreads();
  v[0] = v[1] = v[2] = 0;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const float a = e[i] - me, c = t[i] - mt;
    v[0] += (double)a * c; v[1] += (double)c * c; v[2] += (double)(a - c) * (a - c);
  }
  reduce(v, 3);
  if (threadIdx.x == 0) {
    const double dot = tot[0], en = tot[1] + 1e-8, nn = tot[2] + 1e-8;
    const double scaled2 = dot * dot * tot[1] / (en * en);
    const float neg = (float)(-10.0 * log10(scaled2 / nn + 1e-8));
    out[blockIdx.x] = neg < clip ? clip : neg;
  }
}

__global__ void fill_t_kernel(int* t_dev, int t, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) t_dev[i] = t;
}
__global__ void time_to_int_kernel(const long long* time, int* t_dev, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    long long v = time[i];
    v = v < 0 ? 0 : (v > 999 ? 999 : v);
    t_dev[i] = (int)v;
  }
}

// ------------------------------------------------------------------ load-time helpers
__device__ __forceinline__ float act_in_f(float v, int a) { return a == 1 ? v / (1.f + expf(-v)) : v; }
__device__ __forceinline__ float act_out_f(float v, int a) { return a == 1 ? 0.5f * v * (1.f + erff(v * 0.70710678118654752f)) : v; }

// 32x32 output tile per CTA, K in chunks of 32; block (32, 32)
__global__ void linear_f32_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
                                  float* __restrict__ y, int ldy, int M, int N, int K, int act_in, int act_out) {
  __shared__ float xs[32][33];
  __shared__ float ws[32][33];
  const int m = blockIdx.y * 32 + threadIdx.y, n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int xm = blockIdx.y * 32 + threadIdx.y, xk = k0 + threadIdx.x;
    xs[threadIdx.y][threadIdx.x] = (xm < M && xk < K) ? act_in_f(x[(long long)xm * ldx + xk], act_in) : 0.f;
    const int wn = blockIdx.x * 32 + threadIdx.y;
    ws[threadIdx.y][threadIdx.x] = (wn < N && xk < K) ? W[(long long)wn * K + xk] : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) acc += xs[threadIdx.y][k] * ws[threadIdx.x][k];
    __syncthreads();
  }
  if (m < M && n < N) y[(long long)m * ldy + n] = act_out_f(acc + (bias ? bias[n] : 0.f), act_out);
}

// SinusoidalPosEmb (unet.py:109-116): emb[t][i] = sin(t f_i), emb[t][half+i] = cos(t f_i), f_i = exp(-i ln(1e4)/(half-1))
__global__ void sinusoid_kernel(float* emb, int T, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * half) return;
  const int t = i / half, k = i % half;
  const float e = logf(10000.f) / (float)(half - 1);
  const float f = expf((float)k * -e);
  const float a = (float)t * f;
  emb[(long long)t * dim + k] = sinf(a);
  emb[(long long)t * dim + half + k] = cosf(a);
}

}  // namespace

int gn_apply_launch(const GnApplyArgs& a, int B, cudaStream_t st) {
  const int C = a.y.C;
  LADIFF_REQUIRE(C == 256 || C == 512 || C == 1024, LADIFF_ERR_ARG, "gn_apply: C=%d", C);
  const int rstep = 256 / (C / 8);
  static const int gn_cps = getenv("LADIFF_GN_CPS") ? atoi(getenv("LADIFF_GN_CPS")) : 2;   // CTAs per SM (experiment knob)
  int rows = cdiv(a.L * B, gn_cps * 148);                // one wave of ~gn_cps CTAs per SM, resident before the producer ends (PDL)
  rows = cdiv(rows < rstep ? rstep : rows, rstep) * rstep;
  dim3 grid(cdiv(a.L, rows), B);
  LADIFF_CARVEOUT_ONCE(gn_apply_kernel);
  LADIFF_CUDA_OK(launch_pdl(gn_apply_kernel, grid, dim3(256), 0, st, a, rows));
  return 0;
}

int layernorm_cl_launch(ClView x, const float* g, ClView res, ClView out, int B, int L, cudaStream_t st) {
  const int nvec = x.C / 256, ln_r = nvec >= 4 ? 1 : (nvec >= 2 ? 2 : 4);   // rows per warp, as in the kernel
  static const int ln_cps = getenv("LADIFF_LN_CPS") ? atoi(getenv("LADIFF_LN_CPS")) : 0;   // CTAs per SM (0: one row group per warp)
  int nblk = cdiv(cdiv(L, ln_r) * B, 8);
  if (ln_cps > 0 && nblk > ln_cps * 148) nblk = ln_cps * 148;
  dim3 grid(nblk);
  switch (x.C) {
    case 256: LADIFF_CARVEOUT_ONCE(layernorm_cl_kernel<1>); LADIFF_CUDA_OK(launch_pdl(layernorm_cl_kernel<1>, grid, dim3(256), 0, st, x, g, res, out, L, B)); break;
    case 512: LADIFF_CARVEOUT_ONCE(layernorm_cl_kernel<2>); LADIFF_CUDA_OK(launch_pdl(layernorm_cl_kernel<2>, grid, dim3(256), 0, st, x, g, res, out, L, B)); break;
    case 768: LADIFF_CARVEOUT_ONCE(layernorm_cl_kernel<3>); LADIFF_CUDA_OK(launch_pdl(layernorm_cl_kernel<3>, grid, dim3(256), 0, st, x, g, res, out, L, B)); break;
    case 1024: LADIFF_CARVEOUT_ONCE(layernorm_cl_kernel<4>); LADIFF_CUDA_OK(launch_pdl(layernorm_cl_kernel<4>, grid, dim3(256), 0, st, x, g, res, out, L, B)); break;
    default: LADIFF_REQUIRE(false, LADIFF_ERR_ARG, "layernorm_cl: unsupported C=%d", x.C);
  }
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

void linattn_split(int B, int L, int* S, int* nsplit) {
  static const int la_ctas = getenv("LADIFF_LA_CTAS") ? atoi(getenv("LADIFF_LA_CTAS")) : 600;   // CTAs aimed at over (split, head, clip)
  const int target = cdiv(la_ctas, 4 * B);
  int s = cdiv(L, target < 1 ? 1 : target);
  s = s < 32 ? 32 : (s > LA_S ? LA_S : s);
  *S = s;
  *nsplit = cdiv(L, s);
}
size_t linattn_part_floats(int B, int L) {
  int S, ns;
  linattn_split(B, L, &S, &ns);
  return (size_t)B * 4 * ns * LA_PART;
}

int linattn_launch(ClView qkv, float* ctx, float* part, int* counters, ClView out, int B, int L, cudaStream_t st) {
  int S, ns;
  linattn_split(B, L, &S, &ns);
  LADIFF_CARVEOUT_ONCE(linattn_ctx_kernel);
  LADIFF_CARVEOUT_ONCE(linattn_out_kernel);
  LADIFF_CUDA_OK(launch_pdl(linattn_ctx_kernel, dim3(ns, 4, B), dim3(256), 0, st, qkv, ctx, part, counters, L, S, ns));
  static const bool out_simt = getenv("LADIFF_LINATTN_OUT_SIMT") != nullptr;       // ablation: the fp32 FMA kernel
  if (!out_simt && qkv.pitch % 8 == 0 && out.pitch % 8 == 0) return linattn_out_tc_launch(qkv, ctx, out, B, L, st);
  LADIFF_CUDA_OK(launch_pdl(linattn_out_kernel, dim3(cdiv(L, 64), B), dim3(256), 0, st, qkv, (const float*)ctx, out, L));
  return 0;
}

int linattn_ctx_launch(ClView qkv, float* ctx, float* part, int* counters, int B, int L, cudaStream_t st) {
  int S, ns;
  linattn_split(B, L, &S, &ns);
  LADIFF_CARVEOUT_ONCE(linattn_ctx_kernel);
  LADIFF_CUDA_OK(launch_pdl(linattn_ctx_kernel, dim3(ns, 4, B), dim3(256), 0, st, qkv, ctx, part, counters, L, S, ns));
  return 0;
}

// impl: 1 tiled SIMT (online softmax), 2 SIMT with the keys resident in shared memory (L <= FA2_MAX_KEYS), 3 tcgen05
int fullattn_launch_impl(ClView qkv, ClView out, int B, int L, int impl, cudaStream_t st) {
  if (impl == 3) return fullattn_tc_launch(qkv, out, B, L, st);
  if (impl == 2) {
    LADIFF_REQUIRE(L <= FA2_MAX_KEYS, LADIFF_ERR_ARG, "fullattn2: L=%d > %d", L, FA2_MAX_KEYS);
    const size_t smem = (size_t)L * 64 * sizeof(float);
    LADIFF_CUDA_OK(cudaFuncSetAttribute(fullattn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA2_MAX_KEYS * 64 * (int)sizeof(float)));
    LADIFF_CUDA_OK(launch_pdl(fullattn2_kernel, dim3(cdiv(L, 128), 4, B), dim3(128), smem, st, qkv, out, L));
    return 0;
  }
  LADIFF_CUDA_OK(launch_pdl(fullattn_kernel, dim3(cdiv(L, 32), 4, B), dim3(256), 0, st, qkv, out, L));
  return 0;
}

int fullattn_launch(ClView qkv, ClView out, int B, int L, cudaStream_t st) {
  // LADIFF_ATTN_TC_MIN: bottleneck length from which QK^T / PV run on tcgen05 (default: beyond what the keys-in-smem SIMT kernel holds)
  static const int tc_min = getenv("LADIFF_ATTN_TC_MIN") ? atoi(getenv("LADIFF_ATTN_TC_MIN")) : FA2_MAX_KEYS + 1;
  if (L >= tc_min && qkv.pitch % 8 == 0 && out.pitch % 8 == 0) return fullattn_tc_launch(qkv, out, B, L, st);
  static const bool no_fa2 = getenv("LADIFF_NO_FA2") != nullptr;
  if (L <= FA2_MAX_KEYS && !no_fa2) {
    const size_t smem = (size_t)L * 64 * sizeof(float);
    static unsigned long long attr = 0;
    if (ladiff_first_on_device(&attr)) {
      LADIFF_CUDA_OK(cudaFuncSetAttribute(fullattn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA2_MAX_KEYS * 64 * (int)sizeof(float)));
      prefer_max_smem_carveout(fullattn2_kernel);
    }
    LADIFF_CUDA_OK(launch_pdl(fullattn2_kernel, dim3(cdiv(L, 128), 4, B), dim3(128), smem, st, qkv, out, L));
    return 0;
  }
  LADIFF_CARVEOUT_ONCE(fullattn_kernel);
  LADIFF_CUDA_OK(launch_pdl(fullattn_kernel, dim3(cdiv(L, 32), 4, B), dim3(256), 0, st, qkv, out, L));
  return 0;
}

int ncl_to_cl_launch(const float* x, const float* inv_scale, ClView out, int B, int C, int L, cudaStream_t st) {
  LADIFF_REQUIRE(C % 32 == 0, LADIFF_ERR_ARG, "ncl_to_cl: C=%d", C);
  ncl_to_cl_kernel<<<dim3(cdiv(L, 32), C / 32, B), dim3(32, 8), 0, st>>>(x, inv_scale, out, C, L);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int cl_to_ncl_f32_launch(const float* x, float* y, int B, int C, int L, cudaStream_t st) {
  LADIFF_REQUIRE(C % 32 == 0, LADIFF_ERR_ARG, "cl_to_ncl: C=%d", C);
  cl_to_ncl_f32_kernel<<<dim3(cdiv(L, 32), C / 32, B), dim3(32, 8), 0, st>>>(x, y, C, L);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int absmax_inv_launch(const float* x, float* inv, int B, long long n, float eps, cudaStream_t st) {
  absmax_inv_kernel<<<B, 1024, 0, st>>>(x, inv, n, eps);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int ddpm_step_launch(const float* eps, float* x, const float* noise, unsigned long long seed, int t_abs, unsigned long long clip_offset,
                     StepCoef cf, ClView xin, int B, int C, int L, cudaStream_t st) {
  LADIFF_REQUIRE(C % 32 == 0, LADIFF_ERR_ARG, "ddpm_step: C=%d", C);
  LADIFF_CARVEOUT_ONCE(ddpm_step_kernel);
  ddpm_step_kernel<<<dim3(cdiv(L, 32), C / 32, B), dim3(32, 8), 0, st>>>(eps, x, noise, seed, t_abs, clip_offset, cf, xin, C, L);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int randn_fill_launch(float* x, long long n, unsigned long long seed, unsigned long long elem_offset, int uniform, cudaStream_t st) {
  const long long thr = (n + 3) / 4;
  randn_fill_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, st>>>(x, n, seed, elem_offset, uniform);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int q_sample_launch(const float* x, const float* noise, const int* t_dev, const float* sqrt_ac, const float* sqrt_1m_ac, float* out, int B,
                    long long n, cudaStream_t st) {
  const int gx = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
  q_sample_kernel<<<dim3(gx, B), 256, 0, st>>>(x, noise, t_dev, sqrt_ac, sqrt_1m_ac, out, n);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int axpby_launch(float* x, float a, const float* y, float b, long long n, cudaStream_t st) {
  const long long blocks = (n + 255) / 256;
  axpby_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, st>>>(x, a, y, b, n);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int p_losses_launch(const float* eps, const float* noise, const float* xt, const int* t_dev, const float* recip, const float* recipm1,
                    const float* p2w, float* pred_x0, float* eps_ncl, float* part, float* loss, int B, int C, int L, cudaStream_t st) {
  p_losses_kernel<<<B, 1024, 0, st>>>(eps, noise, xt, t_dev, recip, recipm1, p2w, pred_x0, eps_ncl, part, C, L);
  mean_small_kernel<<<1, 32, 0, st>>>(part, loss, B);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int sdsdr_launch(const float* est, const float* tgt, float* out, int B, long long n, float clip, cudaStream_t st) {
  sdsdr_kernel<<<B, 1024, 0, st>>>(est, tgt, out, n, clip);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int fill_t_launch(int* t_dev, int t, int B, cudaStream_t st) {
  LADIFF_CARVEOUT_ONCE(fill_t_kernel);
  fill_t_kernel<<<cdiv(B, 128), 128, 0, st>>>(t_dev, t, B);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int time_to_int_launch(const long long* time, int* t_dev, int B, cudaStream_t st) {
  time_to_int_kernel<<<cdiv(B, 128), 128, 0, st>>>(time, t_dev, B);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int linear_f32_launch(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int M, int N, int K, int act_in,
                      int act_out, cudaStream_t st) {
  linear_f32_kernel<<<dim3(cdiv(N, 32), cdiv(M, 32)), dim3(32, 32), 0, st>>>(x, ldx, W, b, y, ldy, M, N, K, act_in, act_out);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int sinusoid_launch(float* emb, int T, int dim, cudaStream_t st) {
  sinusoid_kernel<<<cdiv(T * dim / 2, 256), 256, 0, st>>>(emb, T, dim);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
