// Element-wise / normalisation / attention kernels of the UNet (channels-last bf16, fp32 math).
#include <curand_kernel.h>

#include "unet_ops.cuh"

namespace {

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// ------------------------------------------------------------------ GroupNorm apply
__global__ void __launch_bounds__(256) gn_apply_kernel(GnApplyArgs a, int rows_per_cta) {
  extern __shared__ float sm[];
  const int C = a.y.C, Cg = C / 8, spg = Cg / 32, b = blockIdx.y;
  float* sa = sm;
  float* sb = sm + C;
  __shared__ float s_mean[8], s_rstd[8];
  if (threadIdx.x < 8) {
    const int g = threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    for (int nt = 0; nt < a.n_ntiles; ++nt) {
      const float2* row = a.stats + ((long long)b * a.n_ntiles + nt) * (C / 32) + g * spg;
      for (int s = 0; s < spg; ++s) { s1 += row[s].x; s2 += row[s].y; }
    }
    const float n = (float)Cg * (float)a.L;
    const float m = s1 / n;
    float var = s2 / n - m * m;
    var = var < 0.f ? 0.f : var;
    s_mean[g] = m;
    s_rstd[g] = rsqrtf(var + 1e-5f);
  }
  __syncthreads();
  const float* film = nullptr;
  if (a.film) film = a.film + (long long)a.t_dev[b] * a.film_stride;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / Cg;
    float ga = a.gamma[c] * s_rstd[g];
    float be = a.beta[c] - s_mean[g] * ga;
    if (film) {
      const float sc = film[c] + 1.f, sh = film[C + c];
      ga *= sc;
      be = be * sc + sh;
    }
    sa[c] = ga; sb[c] = be;
  }
  __syncthreads();
  const int cv = C / 8;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, a.L);
  const long long total = (long long)(r1 - r0) * cv;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const int row = r0 + (int)(i / cv), c0 = (int)(i % cv) * 8;
    const uint4 u = *reinterpret_cast<const uint4*>(a.y.p + (long long)b * a.y.bstride + (long long)row * a.y.pitch + c0);
    float f[8];
    unpack8(u, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j] * sa[c0 + j] + sb[c0 + j]);
    if (a.res.p) {
      const uint4 ur = *reinterpret_cast<const uint4*>(a.res.p + (long long)b * a.res.bstride + (long long)row * a.res.pitch + c0);
      float r[8];
      unpack8(ur, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    if (a.do_tanh) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = tanhf(f[j]);
    }
    *reinterpret_cast<uint4*>(a.out.p + (long long)b * a.out.bstride + (long long)row * a.out.pitch + c0) = pack8(f);
  }
}

// ------------------------------------------------------------------ channel LayerNorm: one warp per row
template <int NVEC>
__global__ void __launch_bounds__(256) layernorm_cl_kernel(ClView x, const float* __restrict__ g, ClView res, ClView out, int L) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp, b = blockIdx.y;
  if (row >= L) return;
  const int C = NVEC * 256;
  const bf16* xr = x.p + (long long)b * x.bstride + (long long)row * x.pitch;
  float v[NVEC][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    unpack8(*reinterpret_cast<const uint4*>(xr + (lane + 32 * i) * 8), v[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + 1e-5f);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int c0 = (lane + 32 * i) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (v[i][j] - mean) * rstd * g[c0 + j];
    if (res.p) {
      float r[8];
      unpack8(*reinterpret_cast<const uint4*>(res.p + (long long)b * res.bstride + (long long)row * res.pitch + c0), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    *reinterpret_cast<uint4*>(out.p + (long long)b * out.bstride + (long long)row * out.pitch + c0) = pack8(f);
  }
}

// ------------------------------------------------------------------ linear attention
// ctx[b][h][d][e] = 32^-1/2 * sum_n softmax_n(k)[d,n] v[e,n]      grid (4, B), 256 threads
__global__ void __launch_bounds__(256) linattn_ctx_kernel(ClView qkv, float* __restrict__ ctx, int L) {
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* base = qkv.p + (long long)b * qkv.bstride;
  const int kc = 128 + h * 32 + lane, vc = 256 + h * 32 + lane;
  __shared__ float red_m[8][32];
  __shared__ float red_z[8][32];
  __shared__ float tile_k[8][4][32];
  __shared__ __align__(16) float tile_v[8][4][32];
  __shared__ float ctx_s[32][33];
  // pass 1: max over positions per d
  float m = -INFINITY;
  for (int n = warp; n < L; n += 8) m = fmaxf(m, __bfloat162float(base[(long long)n * qkv.pitch + kc]));
  red_m[warp][lane] = m;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) m = fmaxf(m, red_m[w][lane]);
  // pass 2: exp-sum and context
  float z = 0.f, acc[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) acc[e] = 0.f;
  for (int n0 = warp * 4; n0 < L; n0 += 32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + j;
      float kk = 0.f, vv = 0.f;
      if (n < L) {
        kk = __bfloat162float(base[(long long)n * qkv.pitch + kc]);
        vv = __bfloat162float(base[(long long)n * qkv.pitch + vc]);
      }
      tile_k[warp][j][lane] = kk;
      tile_v[warp][j][lane] = vv;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n0 + j < L) {
        const float p = expf(tile_k[warp][j][lane] - m);
        z += p;
        const float4* v4 = reinterpret_cast<const float4*>(tile_v[warp][j]);
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 t = v4[e4];
          acc[4 * e4 + 0] += p * t.x; acc[4 * e4 + 1] += p * t.y;
          acc[4 * e4 + 2] += p * t.z; acc[4 * e4 + 3] += p * t.w;
        }
      }
    }
    __syncwarp();
  }
  red_z[warp][lane] = z;
  for (int w = 0; w < 8; ++w) {       // deterministic warp-ordered reduction
    if (warp == w) {
#pragma unroll
      for (int e = 0; e < 32; ++e) ctx_s[lane][e] = (w == 0 ? 0.f : ctx_s[lane][e]) + acc[e];
    }
    __syncthreads();
  }
  const float scale = 0.17677669529663687f;  // 32^-0.5 (q * scale, unet.py:216)
  for (int i = threadIdx.x; i < 1024; i += 256) {
    const int d = i >> 5, e = i & 31;
    float zz = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) zz += red_z[w][d];
    ctx[(((long long)b * 4 + h) * 32 + d) * 32 + e] = ctx_s[d][e] / zz * scale;
  }
}

// out[b][n][h*32+e] = sum_d ctx[d][e] softmax_d(q[:,n])[d]          grid (ceil(L/64), B), 256 threads
__global__ void __launch_bounds__(256) linattn_out_kernel(ClView qkv, const float* __restrict__ ctx, ClView out, int L) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ float cs[4][32][32];
  for (int i = threadIdx.x; i < 4096; i += 256) (&cs[0][0][0])[i] = ctx[(long long)b * 4096 + i];
  __syncthreads();
  const int n0 = blockIdx.x * 64;
  for (int task = warp; task < 256; task += 8) {
    const int n = n0 + (task >> 2), h = task & 3;
    if (n >= L) break;
    const float q = __bfloat162float(qkv.p[(long long)b * qkv.bstride + (long long)n * qkv.pitch + h * 32 + lane]);
    float mx = q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float ex = expf(q - mx);
    float sum = ex;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float qs = ex / sum;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) acc += cs[h][d][lane] * __shfl_sync(0xffffffffu, qs, d);
    out.p[(long long)b * out.bstride + (long long)n * out.pitch + h * 32 + lane] = __float2bfloat16(acc);
  }
}

// ------------------------------------------------------------------ full attention (mid block), online softmax
// grid (ceil(n/32), 4, B), 256 threads: each warp owns 4 queries; keys/values staged in smem tiles of 128
__global__ void __launch_bounds__(256) fullattn_kernel(ClView qkv, ClView out, int L) {
  const int h = blockIdx.y, b = blockIdx.z, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KT = 128;
  __shared__ float ks[KT][33];
  __shared__ float vs[KT][33];
  const bf16* base = qkv.p + (long long)b * qkv.bstride;
  const float scale = 0.17677669529663687f;
  float qreg[4][32];
  float m[4], l[4], acc[4];
  int qi[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    qi[u] = blockIdx.x * 32 + warp * 4 + u;
    m[u] = -INFINITY; l[u] = 0.f; acc[u] = 0.f;
    float qv = 0.f;
    if (qi[u] < L) qv = __bfloat162float(base[(long long)qi[u] * qkv.pitch + h * 32 + lane]) * scale;
#pragma unroll
    for (int d = 0; d < 32; ++d) qreg[u][d] = __shfl_sync(0xffffffffu, qv, d);
  }
  for (int k0 = 0; k0 < L; k0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * 32; i += 256) {
      const int j = i >> 5, d = i & 31;
      float kk = 0.f, vv = 0.f;
      if (k0 + j < L) {
        kk = __bfloat162float(base[(long long)(k0 + j) * qkv.pitch + 128 + h * 32 + d]);
        vv = __bfloat162float(base[(long long)(k0 + j) * qkv.pitch + 256 + h * 32 + d]);
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
      out.p[(long long)b * out.bstride + (long long)qi[u] * out.pitch + h * 32 + lane] = __float2bfloat16(acc[u] / l[u]);
}

// ------------------------------------------------------------------ layout conversion
// NCL f32 -> channels-last bf16 (optionally scaled per clip).  grid (ceil(L/32), C/32, B), block (32, 8)
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
    if (l < L) out.p[(long long)b * out.bstride + (long long)l * out.pitch + c0 + threadIdx.x] = __float2bfloat16(t[threadIdx.x][i]);
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
// grid (ceil(L/32), C/32, B), block (32, 8)
__global__ void ddpm_step_kernel(const float* __restrict__ eps, float* __restrict__ x, const float* __restrict__ noise,
                                 unsigned long long seed, int step_index, const int* __restrict__ t_dev, DdpmTables tb,
                                 ClView xin, int C, int L) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
  const int ti = t_dev[b];
  const float a = tb.sqrt_recip_ac[ti], bb = tb.sqrt_recipm1_ac[ti], c1 = tb.coef1[ti], c2 = tb.coef2[ti];
  const float sigma = ti > 0 ? expf(0.5f * tb.logvar[ti]) : 0.f;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int l = l0 + i;
    t[i][threadIdx.x] = (l < L) ? eps[((long long)b * L + l) * C + c0 + threadIdx.x] : 0.f;   // t[l][c]
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, l = l0 + threadIdx.x;
    float xn = 0.f;
    if (l < L) {
      const long long idx = ((long long)b * C + c) * L + l;
      const float xv = x[idx];
      float x0 = a * xv - bb * t[threadIdx.x][i];
      x0 = fminf(fmaxf(x0, -1.f), 1.f);
      float z = 0.f;
      if (ti > 0) {
        if (noise) z = noise[idx];
        else {
          curandStatePhilox4_32_10_t st;
          curand_init(seed, (unsigned long long)idx, (unsigned long long)step_index, &st);
          z = curand_normal(&st);
        }
      }
      xn = c1 * x0 + c2 * xv + sigma * z;
      x[idx] = xn;
    }
    t[threadIdx.x][i] = xn;  // t[l][c]: same thread read this element above, no cross-thread hazard
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int l = l0 + i;
    if (l < L) xin.p[(long long)b * xin.bstride + (long long)l * xin.pitch + c0 + threadIdx.x] = __float2bfloat16(t[i][threadIdx.x]);
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
  LADIFF_REQUIRE(C % 256 == 0 && C <= 4096, LADIFF_ERR_ARG, "gn_apply: C=%d", C);
  const int rows = 32;
  dim3 grid(cdiv(a.L, rows), B);
  gn_apply_kernel<<<grid, 256, 2 * C * sizeof(float), st>>>(a, rows);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm_cl_launch(ClView x, const float* g, ClView res, ClView out, int B, int L, cudaStream_t st) {
  dim3 grid(cdiv(L, 8), B);
  switch (x.C) {
    case 256: layernorm_cl_kernel<1><<<grid, 256, 0, st>>>(x, g, res, out, L); break;
    case 512: layernorm_cl_kernel<2><<<grid, 256, 0, st>>>(x, g, res, out, L); break;
    case 768: layernorm_cl_kernel<3><<<grid, 256, 0, st>>>(x, g, res, out, L); break;
    case 1024: layernorm_cl_kernel<4><<<grid, 256, 0, st>>>(x, g, res, out, L); break;
    default: LADIFF_REQUIRE(false, LADIFF_ERR_ARG, "layernorm_cl: unsupported C=%d", x.C);
  }
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int linattn_launch(ClView qkv, float* ctx, ClView out, int B, int L, cudaStream_t st) {
  linattn_ctx_kernel<<<dim3(4, B), 256, 0, st>>>(qkv, ctx, L);
  LADIFF_CUDA_OK(cudaGetLastError());
  linattn_out_kernel<<<dim3(cdiv(L, 64), B), 256, 0, st>>>(qkv, ctx, out, L);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int fullattn_launch(ClView qkv, ClView out, int B, int L, cudaStream_t st) {
  fullattn_kernel<<<dim3(cdiv(L, 32), 4, B), 256, 0, st>>>(qkv, out, L);
  LADIFF_CUDA_OK(cudaGetLastError());
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

int ddpm_step_launch(const float* eps, float* x, const float* noise, unsigned long long seed, int step_index, const int* t_dev,
                     DdpmTables tb, ClView xin, int B, int C, int L, cudaStream_t st) {
  LADIFF_REQUIRE(C % 32 == 0, LADIFF_ERR_ARG, "ddpm_step: C=%d", C);
  ddpm_step_kernel<<<dim3(cdiv(L, 32), C / 32, B), dim3(32, 8), 0, st>>>(eps, x, noise, seed, step_index, t_dev, tb, xin, C, L);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int fill_t_launch(int* t_dev, int t, int B, cudaStream_t st) {
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
