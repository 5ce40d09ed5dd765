// fp32 NCL kernels of the SEANet codec path (HBM/latency-bound side of the pipeline; plain SIMT fp32 so that the
// encoder output feeding the RVQ argmax keeps fp32 fidelity).
#include <stdlib.h>
#include <string.h>

#include "codec_ops.cuh"

namespace {

// ------------------------------------------------------------------ Conv1d (implicit GEMM in registers)
// block 256 = 16 (t) x 16 (co) threads, thread tile RC x RT, CTA tile (16 RC) x (16 RT)
template <int RC, int RT>
__global__ void __launch_bounds__(256) conv1d_f32_kernel(ConvF32Args a, int CI_T, int XW) {
  extern __shared__ __align__(16) float sm[];
  constexpr int CO_T = 16 * RC, T_T = 16 * RT, CO_TP = CO_T + 4;
  const int K = a.K;
  float* xs = sm;                                   // [CI_T][XW]
  float* ws = sm + ((CI_T * XW + 3) & ~3);          // [CI_T*K][CO_TP]
  const int b = blockIdx.z, co0 = blockIdx.y * CO_T, t0 = blockIdx.x * T_T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[RC][RT];
#pragma unroll
  for (int i = 0; i < RC; ++i)
#pragma unroll
    for (int j = 0; j < RT; ++j) acc[i][j] = 0.f;
  const float* xb = a.x + (long long)b * a.Cin * a.Lin;
  const int g0 = t0 * a.stride - a.padL;
  for (int ci0 = 0; ci0 < a.Cin; ci0 += CI_T) {
    const int nci = min(CI_T, a.Cin - ci0);
    __syncthreads();
    for (int i = threadIdx.x; i < nci * XW; i += 256) {
      const int ci = i / XW, p = i - ci * XW;
      int gi = g0 + p;
      if (gi < 0) gi = a.pad_reflect ? -gi : -1;
      else if (gi >= a.Lin) gi = a.pad_reflect ? 2 * (a.Lin - 1) - gi : -1;
      float v = 0.f;
      if (gi >= 0 && gi < a.Lin) {
        v = xb[(long long)(ci0 + ci) * a.Lin + gi];
        if (a.act_in == 1) v = v > 0.f ? v : expm1f(v);
      }
      xs[i] = v;
    }
    const int nr = nci * K;
    for (int i = threadIdx.x; i < nr * CO_T; i += 256) {
      const int co = i / nr, r = i - co * nr;
      ws[r * CO_TP + co] = (co0 + co < a.CoutV) ? a.w[((long long)(co0 + co) * a.Cin + ci0) * K + r] : 0.f;
    }
    __syncthreads();
    for (int ci = 0; ci < nci; ++ci) {
      const float* xrow = xs + ci * XW + tx * a.stride;
      const float* wrow = ws + ci * K * CO_TP + ty * RC;
      for (int k = 0; k < K; ++k) {
        float av[RC], bv[RT];
#pragma unroll
        for (int i = 0; i < RC; ++i) av[i] = wrow[k * CO_TP + i];
#pragma unroll
        for (int j = 0; j < RT; ++j) bv[j] = xrow[16 * j * a.stride + k];
#pragma unroll
        for (int i = 0; i < RC; ++i)
#pragma unroll
          for (int j = 0; j < RT; ++j) acc[i][j] += av[i] * bv[j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RC; ++i) {
    const int co = co0 + ty * RC + i;
    if (co >= a.CoutV) continue;
    if (a.il_s == 0) {
      const float bias = a.bias ? a.bias[co] : 0.f;
      const long long base = ((long long)b * a.CoutV + co) * a.LoutV;
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        const int t = t0 + tx + 16 * j;
        if (t < a.LoutV) {
          float v = acc[i][j] + bias;
          if (a.res) v += a.res[base + t];
          a.y[base + t] = v;
        }
      }
    } else {
      const int ph = co / a.il_cout, cr = co - ph * a.il_cout;
      const float bias = a.bias ? a.bias[cr] : 0.f;
      const long long base = ((long long)b * a.il_cout + cr) * a.il_lout;
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        const int t = t0 + tx + 16 * j;
        const int pos = t * a.il_s + ph - a.il_trim;
        if (t < a.LoutV && pos >= 0 && pos < a.il_lout) a.y[base + pos] = acc[i][j] + bias;
      }
    }
  }
}

// ------------------------------------------------------------------ Conv1d, register-tiled (the fast path)
// A strided conv (K = 2*stride in SEANet) is a stride-1 conv over S = stride "phase channels" with KT = K/S taps:
//   y[co][t] = sum_{cv = ci*S + p} sum_{kt < KT} wt[cv*KT + kt][co] * xv[cv][t + kt],   xv[ci*S + p][u] = xpad[ci][u*S + p - padL]
// so one kernel serves plain, strided and (phase-interleaved) transposed convs.  Block 256 = 16 (t) x 16 (co) threads; thread tile
// RC channels x 8 consecutive positions: per phase channel the thread reads its 8+KT-1 inputs once (LDS.128) and reuses them for
// every tap, weights come as broadcast LDS.128 -> ~0.05 shared-memory instructions per FMA.  The output tile goes through
// shared memory so that global stores (and residual loads) are coalesced along time also for the interleaved layout.
template <int RC, int KT>
__global__ void __launch_bounds__(256, RC >= 8 ? 2 : 3) conv1d_f32_v2_kernel(ConvF32Args a, int CI_T, int S, unsigned magic_span, unsigned magic_s) {
  extern __shared__ __align__(16) float sm2[];
  constexpr int CO_T = 16 * RC, T_T = 128, XW = T_T + KT - 1, XWP = (XW + 3) & ~3, NX = 8 + KT - 1;
  float* xs = sm2;                       // [CI_T][XWP]
  float* ws = sm2 + CI_T * XWP;          // [CI_T*KT][CO_T]
  const int b = blockIdx.z, co0 = blockIdx.y * CO_T, t0 = blockIdx.x * T_T;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[RC][8];
#pragma unroll
  for (int i = 0; i < RC; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int CinV = a.Cin * S;
  const float* xb = a.x + (long long)b * a.Cin * a.Lin;
  const long long g_base = (long long)t0 * S - a.padL;
  const int span = XW * S;
  for (int cv0 = 0; cv0 < CinV; cv0 += CI_T) {
    const int nci = min(CI_T, CinV - cv0), nreal = nci / S, cvr0 = cv0 / S;
    __syncthreads();
    // fills: batches of independent loads (all in flight before the first shared-memory store): the layers with long time
    // axes are latency-bound on exactly these loads
    constexpr int FB = 8;
    const int xtotal = nreal * span;
    for (int i0 = tid; i0 < xtotal; i0 += 256 * FB) {
      float v[FB];
      int dst[FB];
#pragma unroll
      for (int qq = 0; qq < FB; ++qq) {
        const int i = i0 + qq * 256;
        v[qq] = 0.f; dst[qq] = -1;
        if (i < xtotal) {
          // (divisions by multiply-high with host-made reciprocals: the generic integer division was 40 % of the kernel's instructions)
          const int ci = (int)__umulhi((unsigned)i, magic_span), g = i - ci * span;
          long long gi = g_base + g;
          if (gi < 0) gi = a.pad_reflect ? -gi : -1;
          else if (gi >= a.Lin) gi = a.pad_reflect ? 2LL * (a.Lin - 1) - gi : -1;
          if (gi >= 0 && gi < a.Lin) v[qq] = __ldg(xb + (long long)(cvr0 + ci) * a.Lin + gi);
          const int u = S == 1 ? g : (int)__umulhi((unsigned)g, magic_s), ph = g - u * S;
          dst[qq] = (ci * S + ph) * XWP + u;
        }
      }
#pragma unroll
      for (int qq = 0; qq < FB; ++qq) {
        if (dst[qq] >= 0) {
          float vv = v[qq];
          if (a.act_in == 1) vv = vv > 0.f ? vv : expm1f(vv);
          xs[dst[qq]] = vv;
        }
      }
    }
    if ((a.CoutV & 3) == 0) {
      constexpr int WB = 4;
      const int wtotal = nci * KT * (CO_T / 4);
      for (int i0 = tid; i0 < wtotal; i0 += 256 * WB) {
        float4 v[WB];
#pragma unroll
        for (int qq = 0; qq < WB; ++qq) {
          const int i = i0 + qq * 256;
          v[qq] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < wtotal) {
            const int r = i / (CO_T / 4), c4 = i - r * (CO_T / 4), co = co0 + 4 * c4;
            if (co < a.CoutV) v[qq] = __ldg(reinterpret_cast<const float4*>(a.wt + (long long)(cv0 * KT + r) * a.CoutV + co));
          }
        }
#pragma unroll
        for (int qq = 0; qq < WB; ++qq) {
          const int i = i0 + qq * 256;
          if (i < wtotal) {
            const int r = i / (CO_T / 4), c4 = i - r * (CO_T / 4);
            *reinterpret_cast<float4*>(ws + r * CO_T + 4 * c4) = v[qq];
          }
        }
      }
    } else {
      for (int i = tid; i < nci * KT * CO_T; i += 256) {
        const int r = i / CO_T, c = i - r * CO_T;
        ws[i] = (co0 + c < a.CoutV) ? __ldg(a.wt + (long long)(cv0 * KT + r) * a.CoutV + co0 + c) : 0.f;
      }
    }
    __syncthreads();
    for (int cv = 0; cv < nci; ++cv) {
      const float* xr = xs + cv * XWP + tx * 8;
      float xv[NX];
      {
        const float4 x0 = *reinterpret_cast<const float4*>(xr), x1 = *reinterpret_cast<const float4*>(xr + 4);
        xv[0] = x0.x; xv[1] = x0.y; xv[2] = x0.z; xv[3] = x0.w; xv[4] = x1.x; xv[5] = x1.y; xv[6] = x1.z; xv[7] = x1.w;
#pragma unroll
        for (int j = 8; j < NX; ++j) xv[j] = xr[j];
      }
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const float* wr = ws + (cv * KT + k) * CO_T + ty * RC;
        float wv[RC];
        if (RC >= 4) {
#pragma unroll
          for (int i4 = 0; i4 < RC / 4; ++i4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + 4 * i4);
            wv[4 * i4] = w4.x; wv[4 * i4 + 1] = w4.y; wv[4 * i4 + 2] = w4.z; wv[4 * i4 + 3] = w4.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < RC; ++i) wv[i] = wr[i];
        }
#pragma unroll
        for (int i = 0; i < RC; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += wv[i] * xv[j + k];
      }
    }
  }
  // ---- epilogue: two passes of 8*RC channels through shared memory ([channel][EPW]), coalesced copy-out
  constexpr int EPW = T_T + 4, ECH = 8 * RC;
  float* es = sm2;
  const int s_il = a.il_s ? a.il_s : 1;
  const int nreal_ch = ECH / s_il;               // real channels per pass (interleave: s_il virtual channels each)
  const int ow = T_T * s_il;                     // output positions per real channel per tile
  for (int h = 0; h < 2; ++h) {
    __syncthreads();
    if ((ty >> 3) == h) {
#pragma unroll
      for (int i = 0; i < RC; ++i) {
        float* er = es + ((ty & 7) * RC + i) * EPW + tx * 8;
        *reinterpret_cast<float4*>(er) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(er + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
    __syncthreads();
    const int v0 = co0 + h * ECH;                // first virtual channel of this pass
    const int etotal = nreal_ch * ow;
    if (a.il_s == 0) {
      constexpr int EB = 4;
      for (int e0 = tid; e0 < etotal; e0 += 256 * EB) {
        float rv[EB], bv[EB];
        long long oo[EB];
#pragma unroll
        for (int qq = 0; qq < EB; ++qq) {
          const int e = e0 + qq * 256;
          oo[qq] = -1; rv[qq] = 0.f; bv[qq] = 0.f;
          if (e < etotal) {
            const int crl = e >> 7, pos = e & (T_T - 1), vch = v0 + crl;       // ow == T_T == 128 here
            if (vch < a.CoutV && t0 + pos < a.LoutV) {
              oo[qq] = ((long long)b * a.CoutV + vch) * a.LoutV + t0 + pos;
              if (a.res) rv[qq] = __ldg(a.res + oo[qq]);
              if (a.bias) bv[qq] = __ldg(a.bias + vch);
            }
          }
        }
#pragma unroll
        for (int qq = 0; qq < EB; ++qq) {
          const int e = e0 + qq * 256;
          if (oo[qq] >= 0) { const int crl = e >> 7, pos = e & (T_T - 1); a.y[oo[qq]] = es[crl * EPW + pos] + bv[qq] + rv[qq]; }
        }
      }
    } else {
      const int sh = 31 - __clz(s_il);             // il_s is 2, 4 or 8 on this path (launcher)
      for (int e = tid; e < etotal; e += 256) {
        const int crl = e >> (7 + sh), op = e & (ow - 1);
        const int pos = op >> sh, ph = op & (s_il - 1);
        const int vch = v0 + crl * s_il + ph;
        if (vch >= a.CoutV || t0 + pos >= a.LoutV) continue;
        const int cr = vch >> sh;
        const long long opos = (long long)(t0 + pos) * s_il + ph - a.il_trim;
        if (opos >= 0 && opos < a.il_lout)
          a.y[((long long)b * a.il_cout + cr) * a.il_lout + opos] = es[(crl * s_il + ph) * EPW + pos] + (a.bias ? __ldg(a.bias + cr) : 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------ fused SEANet residual block (seanet.py:45-63), C = 32 / 64
//   y = conv1x1_shortcut(x) + conv_k1(ELU(conv_k3(ELU(x))))        hidden = C/2, causal reflect padding 2 for the k3 (conv.py:217-232)
// One read of x and one write of y per element (the three-launch form moves x three times, the hidden tensor twice and the shortcut
// twice).  CTA = NTH threads (8 warps for C = 32, 16 for C = 64), TT positions of one clip; shared memory holds raw x, ELU(x) with its 2-sample halo, ELU(hidden) and
// all three weight matrices (K-major).  Thread tile: (channels of its warp) x (groups of 4 consecutive positions, 128 apart):
// activations as LDS.128/.64 (lanes 16 B apart: conflict-free), weights as warp-broadcast vector loads.
template <int C, int TT, int NTH>
__global__ void __launch_bounds__(NTH) seanet_resblock_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w1t,
                                                               const float* __restrict__ b1, const float* __restrict__ wsct,
                                                               const float* __restrict__ bsc, const float* __restrict__ w2t,
                                                               const float* __restrict__ b2, int L) {
  constexpr int HID = C / 2, XP = TT + 8, NG = TT / 128, NW = NTH / 32, R1 = HID / NW, R2 = C / NW;
  extern __shared__ __align__(16) float rsm[];
  float* xe = rsm;                          // [C][XP]   ELU(x), element t0 + i at index i + 2
  float* xraw = xe + C * XP;                // [C][TT]
  float* he = xraw + C * TT;                // [HID][TT] ELU(hidden)
  float* w1s = he + HID * TT;               // [C*3][HID]
  float* wss = w1s + C * 3 * HID;           // [C][C]
  float* w2s = wss + C * C;                 // [HID][C]
  const int b = blockIdx.y, t0 = blockIdx.x * TT, tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const float* xb = x + (long long)b * C * L;
  // ---- stage 0: weights and the x tile
  for (int i = tid; i < C * 3 * HID / 4; i += NTH) reinterpret_cast<float4*>(w1s)[i] = __ldg(reinterpret_cast<const float4*>(w1t) + i);
  for (int i = tid; i < C * C / 4; i += NTH) reinterpret_cast<float4*>(wss)[i] = __ldg(reinterpret_cast<const float4*>(wsct) + i);
  for (int i = tid; i < HID * C / 4; i += NTH) reinterpret_cast<float4*>(w2s)[i] = __ldg(reinterpret_cast<const float4*>(w2t) + i);
  for (int i = tid; i < C * (TT / 4); i += NTH) {
    const int c = i / (TT / 4), p4 = (i - c * (TT / 4)) * 4, t = t0 + p4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t + 3 < L) v = __ldg(reinterpret_cast<const float4*>(xb + (long long)c * L + t));
    else {
      if (t < L) v.x = __ldg(xb + (long long)c * L + t);
      if (t + 1 < L) v.y = __ldg(xb + (long long)c * L + t + 1);
      if (t + 2 < L) v.z = __ldg(xb + (long long)c * L + t + 2);
    }
    *reinterpret_cast<float4*>(xraw + c * TT + p4) = v;
    float* e = xe + c * XP + p4 + 2;
    e[0] = v.x > 0.f ? v.x : expm1f(v.x); e[1] = v.y > 0.f ? v.y : expm1f(v.y);
    e[2] = v.z > 0.f ? v.z : expm1f(v.z); e[3] = v.w > 0.f ? v.w : expm1f(v.w);
  }
  if (tid < 2 * C) {                        // halo: x[t0-2], x[t0-1]; reflect at the clip start (x[-1] = x[1], x[-2] = x[2])
    const int c = tid >> 1, hh = tid & 1;   // hh = 0 -> offset -2, 1 -> offset -1
    int t = t0 - 2 + hh;
    if (t < 0) t = -t;
    float v = t < L ? __ldg(xb + (long long)c * L + t) : 0.f;
    xe[c * XP + hh] = v > 0.f ? v : expm1f(v);
  }
  __syncthreads();
  // ---- stage 1: hidden = conv_k3(ELU(x)) + b1 -> ELU -> he
  {
    float acc[R1][NG * 4];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      const float bb = __ldg(b1 + ty * R1 + r);
#pragma unroll
      for (int j = 0; j < NG * 4; ++j) acc[r][j] = bb;
    }
    for (int ci = 0; ci < C; ++ci) {
      float xv[NG][6];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float* xr = xe + ci * XP + 4 * tx + 128 * g;
        const float4 a4 = *reinterpret_cast<const float4*>(xr);
        const float2 a2 = *reinterpret_cast<const float2*>(xr + 4);
        xv[g][0] = a4.x; xv[g][1] = a4.y; xv[g][2] = a4.z; xv[g][3] = a4.w; xv[g][4] = a2.x; xv[g][5] = a2.y;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float wv[R1];
        const float* wr = w1s + (ci * 3 + k) * HID + ty * R1;
        if (R1 == 2) { const float2 t2 = *reinterpret_cast<const float2*>(wr); wv[0] = t2.x; wv[1] = t2.y; }
        else {
#pragma unroll
          for (int r4 = 0; r4 < R1 / 4; ++r4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wr + 4 * r4);
            wv[4 * r4] = t4.x; wv[4 * r4 + 1] = t4.y; wv[4 * r4 + 2] = t4.z; wv[4 * r4 + 3] = t4.w;
          }
        }
#pragma unroll
        for (int r = 0; r < R1; ++r)
#pragma unroll
          for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][g * 4 + j] = fmaf(wv[r], xv[g][j + k], acc[r][g * 4 + j]);
      }
    }
#pragma unroll
    for (int r = 0; r < R1; ++r)
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        float4 o;
        o.x = acc[r][g * 4 + 0]; o.y = acc[r][g * 4 + 1]; o.z = acc[r][g * 4 + 2]; o.w = acc[r][g * 4 + 3];
        o.x = o.x > 0.f ? o.x : expm1f(o.x); o.y = o.y > 0.f ? o.y : expm1f(o.y);
        o.z = o.z > 0.f ? o.z : expm1f(o.z); o.w = o.w > 0.f ? o.w : expm1f(o.w);
        *reinterpret_cast<float4*>(he + (ty * R1 + r) * TT + 4 * tx + 128 * g) = o;
      }
  }
  __syncthreads();
  // ---- stage 2: y = W_sc x + b_sc + W_2 ELU(hidden) + b_2
  float acc[R2][NG * 4];
#pragma unroll
  for (int r = 0; r < R2; ++r) {
    const float bb = __ldg(bsc + ty * R2 + r) + __ldg(b2 + ty * R2 + r);
#pragma unroll
    for (int j = 0; j < NG * 4; ++j) acc[r][j] = bb;
  }
  auto accumulate = [&](const float* act, const float* wrow) {      // one input channel: act [TT], wrow [C]
    float wv[R2];
#pragma unroll
    for (int r4 = 0; r4 < R2 / 4; ++r4) {
      const float4 t4 = *reinterpret_cast<const float4*>(wrow + ty * R2 + 4 * r4);
      wv[4 * r4] = t4.x; wv[4 * r4 + 1] = t4.y; wv[4 * r4 + 2] = t4.z; wv[4 * r4 + 3] = t4.w;
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const float4 a4 = *reinterpret_cast<const float4*>(act + 4 * tx + 128 * g);
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        acc[r][g * 4 + 0] = fmaf(wv[r], a4.x, acc[r][g * 4 + 0]); acc[r][g * 4 + 1] = fmaf(wv[r], a4.y, acc[r][g * 4 + 1]);
        acc[r][g * 4 + 2] = fmaf(wv[r], a4.z, acc[r][g * 4 + 2]); acc[r][g * 4 + 3] = fmaf(wv[r], a4.w, acc[r][g * 4 + 3]);
      }
    }
  };
  for (int ci = 0; ci < C; ++ci) accumulate(xraw + ci * TT, wss + ci * C);
  for (int hc = 0; hc < HID; ++hc) accumulate(he + hc * TT, w2s + hc * C);
  float* yb = y + (long long)b * C * L;
#pragma unroll
  for (int r = 0; r < R2; ++r)
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int t = t0 + 4 * tx + 128 * g;
      float* yr = yb + (long long)(ty * R2 + r) * L + t;
      if (t + 3 < L) *reinterpret_cast<float4*>(yr) = make_float4(acc[r][g * 4 + 0], acc[r][g * 4 + 1], acc[r][g * 4 + 2], acc[r][g * 4 + 3]);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (t + j < L) yr[j] = acc[r][g * 4 + j];
      }
    }
}

// ------------------------------------------------------------------ SEANet output conv: C -> 1 channel, k = 7, causal reflect pad
// (seanet.py:236; conv.py:217-232).  HBM-bound: every input element is read once; thread = 4 consecutive output samples, the
// ELU'd input tile of the CTA lives in shared memory, weights in registers via broadcast loads.
template <int C>
__global__ void __launch_bounds__(128) conv_out1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                        float* __restrict__ y, int L, int act_in) {
  constexpr int TT = 512, XP = TT + 8;         // 6 halo samples on the left, stored at index i + 6 - t0... (row pitch keeps 16-byte alignment)
  extern __shared__ __align__(16) float osm[];
  float* xs = osm;                               // [C][XP]: sample t0 - 8 + i at index i  (two unused slots keep float4 alignment)
  float* ws = osm + C * XP;                      // [C][8]
  const int b = blockIdx.y, t0 = blockIdx.x * TT, tid = threadIdx.x;
  const float* xb = x + (long long)b * C * L;
  for (int i = tid; i < C * 7; i += 128) ws[(i / 7) * 8 + i % 7] = __ldg(w + i);
  for (int i = tid; i < C * (XP / 4); i += 128) {
    const int c = i / (XP / 4), p4 = (i - c * (XP / 4)) * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int t = t0 - 8 + p4 + j;
      if (t < 0) t = -t;                          // reflect (only the first tile; |t| <= 8 < L)
      float u = t < L ? __ldg(xb + (long long)c * L + t) : 0.f;
      if (act_in == 1) u = u > 0.f ? u : expm1f(u);
      v[j] = u;
    }
    *reinterpret_cast<float4*>(xs + c * XP + p4) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) {
    // outputs t0 + 4 tid + j (j < 4) read samples t - 6 .. t  ->  indices 4 tid + 2 + j .. 4 tid + 8 + j: 12 floats from 4 tid
    const float* xr = xs + c * XP + 4 * tid;
    const float4 a0 = *reinterpret_cast<const float4*>(xr), a1 = *reinterpret_cast<const float4*>(xr + 4), a2 = *reinterpret_cast<const float4*>(xr + 8);
    const float xv[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
    const float4 w0 = *reinterpret_cast<const float4*>(ws + c * 8), w1 = *reinterpret_cast<const float4*>(ws + c * 8 + 4);
    const float wv[7] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z};
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(wv[k], xv[2 + j + k], acc[j]);
  }
  const float bb = bias ? __ldg(bias) : 0.f;
  const int t = t0 + 4 * tid;
  float* yr = y + (long long)b * L + t;
  if (t + 3 < L && (((uintptr_t)yr) & 15) == 0) *reinterpret_cast<float4*>(yr) = make_float4(acc[0] + bb, acc[1] + bb, acc[2] + bb, acc[3] + bb);
  else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (t + j < L) yr[j] = acc[j] + bb;
  }
}

__global__ void conv_w_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int CoutV, int Cin, int K, int S, int perm_s,
                                        int perm_cout) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)CoutV * Cin * K;
  if (i >= total) return;
  const int KT = K / S;
  const int col = (int)(i % CoutV);
  const long long r = i / CoutV;                 // (ci*S + p)*KT + kt
  const int kt = (int)(r % KT);
  const int cv = (int)(r / KT), ci = cv / S, p = cv - ci * S;
  int co = col;
  if (perm_s > 0) { const int cr = col / perm_s, ph = col - cr * perm_s; co = ph * perm_cout + cr; }
  wt[i] = w[((long long)co * Cin + ci) * K + kt * S + p];
}

// ------------------------------------------------------------------ LSTM, small H: one persistent CTA per clip
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

template <int H, int KREG>
__global__ void __launch_bounds__(4 * H) lstm_seq_kernel(const float* __restrict__ pre, const float* __restrict__ whh,
                                                         const float* __restrict__ skip, float* __restrict__ y, int T) {
  constexpr int KS = H - KREG;
  extern __shared__ __align__(16) float wsm[];      // [KS][4H] tail of the recurrent weights
  __shared__ __align__(16) float hs[2][H];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int q = tid & 3, j = tid >> 2, row = q * H + j;   // lanes 4j..4j+3 = gates i,f,g,o of unit j
  float w[KREG];
#pragma unroll
  for (int k = 0; k < KREG; ++k) w[k] = whh[(long long)row * H + k];
  for (int k = 0; k < KS; ++k) wsm[k * 4 * H + tid] = whh[(long long)row * H + KREG + k];
  if (tid < H) hs[0][tid] = 0.f;
  float c = 0.f;
  const float* prow = pre + ((long long)b * 4 * H + row) * T;
  const long long ybase = ((long long)b * H + j) * T;
  __syncthreads();
  int cur = 0;
  for (int t0 = 0; t0 < T; t0 += 8) {
    float pbuf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pbuf[i] = (t0 + i < T) ? prow[t0 + i] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (t0 + i < T) {   // uniform across the block
        float acc0 = pbuf[i], acc1 = 0.f;
        const float4* h4 = reinterpret_cast<const float4*>(hs[cur]);
#pragma unroll
        for (int k4 = 0; k4 < KREG / 4; ++k4) {
          const float4 hv = h4[k4];
          acc0 += w[4 * k4 + 0] * hv.x; acc1 += w[4 * k4 + 1] * hv.y;
          acc0 += w[4 * k4 + 2] * hv.z; acc1 += w[4 * k4 + 3] * hv.w;
        }
        if (KS > 0) {
#pragma unroll 8
          for (int k = 0; k < KS; ++k) acc0 += wsm[k * 4 * H + tid] * hs[cur][KREG + k];
        }
        const float acc = acc0 + acc1;
        const int base = (tid & 31) & ~3;
        const float gi = __shfl_sync(0xffffffffu, acc, base + 0);
        const float gf = __shfl_sync(0xffffffffu, acc, base + 1);
        const float gg = __shfl_sync(0xffffffffu, acc, base + 2);
        const float go = __shfl_sync(0xffffffffu, acc, base + 3);
        c = sigmoid_f(gf) * c + sigmoid_f(gi) * tanhf(gg);
        const float h = sigmoid_f(go) * tanhf(c);
        if (q == 0) {
          hs[cur ^ 1][j] = h;
          y[ybase + t0 + i] = skip ? h + skip[ybase + t0 + i] : h;
        }
        __syncthreads();
        cur ^= 1;
      }
    }
  }
}

// ------------------------------------------------------------------ LSTM, small H: a CTA CLUSTER per clip (gate rows split over SMs)
// The recurrence is serial in time, so its speed is the latency of ONE step.  A cluster of CL = H/32 CTAs (128 threads each) owns a
// clip; CTA r owns hidden units [32r, 32r+32) = 128 gate rows, one row per thread (thread = 4*unit + gate), the row's H recurrent
// weights live in REGISTERS, h_{t-1} in the CTA's shared memory (double-buffered).  Per step: 4-chain dot product against
// broadcast LDS.128 reads of h, quad shuffle -> the four gates of a unit, cell update (every lane of the quad keeps c), then lane q of
// the quad sends h_t[unit] to CTA q with ONE one-way message: st.async into the peer's shared memory that also completes 4 bytes
// on the peer's mbarrier (tx-count), the same mechanism TMA uses.  Every thread then waits for the step's H*4 bytes.  Two
// barriers alternate by step parity, so a fast peer's message for step t+1 can never be counted in step t; a peer only gets to
// step t+1 after it has received THIS CTA's step-t values, i.e. after all of this CTA's warps finished reading the h buffer that
// step t+1 overwrites.  No __syncthreads, no cluster barrier and no release fence inside the loop.
__device__ __forceinline__ uint32_t lc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lc_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
// gates with MUFU-based exp (rel. error ~2^-21): far below the decoder's 5e-5 tolerance; the ENCODER's LSTM (whose output feeds
// the RVQ argmax) runs lstm_persist_kernel with expf / tanhf
__device__ __forceinline__ float lc_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float lc_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

template <int H>
__global__ void __launch_bounds__(128, 1) lstm_cluster_kernel(const float* __restrict__ pre, const float* __restrict__ whh,
                                                              const float* __restrict__ skip, float* __restrict__ y, int T) {
  constexpr int CL = H / 32;
  __shared__ __align__(16) float hs[2][H];
  __shared__ __align__(8) uint64_t bar[2];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int b = blockIdx.x / CL, tid = threadIdx.x;
  const int q = tid & 3, u = tid >> 2, j = 32 * (int)rank + u, row = q * H + j;
  float w[H];
#pragma unroll
  for (int k = 0; k < H; k += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(whh + (long long)row * H + k));
    w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
  }
  for (int i = tid; i < 2 * H; i += 128) (&hs[0][0])[i] = 0.f;
  const uint32_t bar0 = lc_smem_u32(&bar[0]), bar1 = lc_smem_u32(&bar[1]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // steps 0 and 1 expect H*4 bytes each (posted before any peer can send)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(H * 4) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar1), "r"(H * 4) : "memory");
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers' barriers and hs are ready
  const bool sender = q < CL;          // lane q of a unit's quad feeds CTA q of the cluster
  const uint32_t r_bar0 = sender ? lc_mapa(bar0, (uint32_t)q) : 0u, r_bar1 = sender ? lc_mapa(bar1, (uint32_t)q) : 0u;
  const uint32_t r_hs0 = sender ? lc_mapa(lc_smem_u32(&hs[0][j]), (uint32_t)q) : 0u;
  const uint32_t r_hs1 = sender ? lc_mapa(lc_smem_u32(&hs[1][j]), (uint32_t)q) : 0u;
  float c = 0.f;
  const float* prow = pre + ((long long)b * 4 * H + row) * T;
  const long long ybase = ((long long)b * H + j) * T;
  for (int t0 = 0; t0 < T; t0 += 8) {
    float pbuf[8], sbuf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pbuf[i] = (t0 + i < T) ? __ldg(prow + t0 + i) : 0.f;
      sbuf[i] = (skip && q == 3 && t0 + i < T) ? __ldg(skip + ybase + t0 + i) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = t0 + i;
      if (t >= T) break;            // uniform
      const int cur = t & 1;        // step t reads hs[cur], fills hs[cur ^ 1] and completes bar[cur]
      float a0 = pbuf[i], a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float4* h4 = reinterpret_cast<const float4*>(hs[cur]);
#pragma unroll
      for (int k4 = 0; k4 < H / 4; ++k4) {
        const float4 hv = h4[k4];
        a0 = fmaf(w[4 * k4 + 0], hv.x, a0); a1 = fmaf(w[4 * k4 + 1], hv.y, a1);
        a2 = fmaf(w[4 * k4 + 2], hv.z, a2); a3 = fmaf(w[4 * k4 + 3], hv.w, a3);
      }
      const float acc = (a0 + a1) + (a2 + a3);
      const int base = (tid & 31) & ~3;
      const float gi = __shfl_sync(0xffffffffu, acc, base + 0);
      const float gf = __shfl_sync(0xffffffffu, acc, base + 1);
      const float gg = __shfl_sync(0xffffffffu, acc, base + 2);
      const float go = __shfl_sync(0xffffffffu, acc, base + 3);
      c = lc_sigmoid(gf) * c + lc_sigmoid(gi) * lc_tanh(gg);
      const float h = lc_sigmoid(go) * lc_tanh(c);
      if (sender)
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                     ::"r"(cur ? r_hs0 : r_hs1), "r"(__float_as_uint(h)), "r"(cur ? r_bar1 : r_bar0) : "memory");
      if (q == 3) y[ybase + t] = skip ? h + sbuf[i] : h;
      // wait for h_t of every unit of the clip: H*4 bytes on bar[cur]; phase parity = (t >> 1) & 1
      const uint32_t bar_c = cur ? bar1 : bar0, parity = (uint32_t)(t >> 1) & 1u;
      uint32_t ok = 0;
      while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(bar_c), "r"(parity) : "memory");
      }
      // re-arm this barrier for step t + 2 (its phase just completed; nobody sends step t + 2 data before receiving OUR step t + 1)
      if (tid == 0 && t + 2 < T) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_c), "r"(H * 4) : "memory");
    }
  }
  // no CTA may exit while a peer can still write into its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ LSTM, any H: one launch per time step
// grid (H/4, ceil(B/32)), 128 threads; CTA = 4 hidden units (16 gate rows) x 32 clips
__global__ void __launch_bounds__(128) lstm_step_kernel(const float* __restrict__ pre, const float* __restrict__ whh,
                                                        const float* __restrict__ skip, float* __restrict__ y,
                                                        const float* __restrict__ h_prev, float* __restrict__ h_next,
                                                        float* __restrict__ cbuf, int B, int H, int T, int t) {
  constexpr int KC = 64;
  __shared__ float Ws[16][KC + 1];
  __shared__ float hs[32][KC + 1];
  __shared__ float gs[16][33];
  const int j0 = blockIdx.x * 4, b0 = blockIdx.y * 32, tid = threadIdx.x;
  const int rp = tid >> 4, bp = tid & 15;
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
  for (int k0 = 0; k0 < H; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < 16 * KC; i += 128) {
      const int lr = i / KC, k = i - lr * KC;           // local row lr = gate*4 + unit
      const int grow = (lr >> 2) * H + j0 + (lr & 3);
      Ws[lr][k] = (k0 + k < H) ? whh[(long long)grow * H + k0 + k] : 0.f;
    }
    for (int i = tid; i < 32 * KC; i += 128) {
      const int bb = i / KC, k = i - bb * KC;
      hs[bb][k] = (b0 + bb < B && k0 + k < H) ? h_prev[(long long)(b0 + bb) * H + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 16
    for (int k = 0; k < KC; ++k) {
      const float w0 = Ws[2 * rp][k], w1 = Ws[2 * rp + 1][k];
      const float h0 = hs[2 * bp][k], h1 = hs[2 * bp + 1][k];
      a00 += w0 * h0; a01 += w0 * h1; a10 += w1 * h0; a11 += w1 * h1;
    }
  }
  gs[2 * rp][2 * bp] = a00; gs[2 * rp][2 * bp + 1] = a01;
  gs[2 * rp + 1][2 * bp] = a10; gs[2 * rp + 1][2 * bp + 1] = a11;
  __syncthreads();
  const int u = tid >> 5, bb = tid & 31, bg = b0 + bb, j = j0 + u;
  if (bg < B) {
    const long long pbase = (long long)bg * 4 * H * T + t;
    const float gi = gs[0 * 4 + u][bb] + pre[pbase + (long long)(0 * H + j) * T];
    const float gf = gs[1 * 4 + u][bb] + pre[pbase + (long long)(1 * H + j) * T];
    const float gg = gs[2 * 4 + u][bb] + pre[pbase + (long long)(2 * H + j) * T];
    const float go = gs[3 * 4 + u][bb] + pre[pbase + (long long)(3 * H + j) * T];
    const float c = sigmoid_f(gf) * cbuf[(long long)bg * H + j] + sigmoid_f(gi) * tanhf(gg);
    const float h = sigmoid_f(go) * tanhf(c);
    cbuf[(long long)bg * H + j] = c;
    h_next[(long long)bg * H + j] = h;
    const long long yi = ((long long)bg * H + j) * T + t;
    y[yi] = skip ? h + skip[yi] : h;
  }
}

// ------------------------------------------------------------------ LSTM, large H: ONE persistent cooperative kernel per layer
// grid H/4 CTAs x 256 threads, all co-resident (cooperative launch).  CTA c owns hidden units 4c..4c+3 = 16 gate rows, whose
// recurrent weights stay in shared memory ([k][16], 32 B per k) for the whole sequence.  Clips are processed in groups of 32.
// Per time step: every CTA pulls h_{t-1} of the group ([H][32] fp32, k-major, written by all CTAs) from L2 into shared memory,
// warp w multiplies its K-slice (thread tile 4 rows x 4 clips, operands as broadcast LDS.128), the 8 partial tiles are summed
// in a fixed order, 128 threads apply the gates (cell state lives in their registers) and publish h_t; a global arrive/spin
// barrier (monotonic counter) separates the steps.  Replaces T launches of lstm_step_kernel.
constexpr int LP_B = 32;      // clips per group
constexpr int LP_WARPS = 8;

__device__ __forceinline__ void lp_grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 1) lstm_persist_kernel(const float* __restrict__ pre, const float* __restrict__ whh,
                                                               const float* __restrict__ skip, float* __restrict__ y,
                                                               float* __restrict__ hglob, unsigned int* counter, int B, int H, int T) {
  extern __shared__ __align__(16) float lp_sm[];
  float* Wt = lp_sm;                              // [H][16]  (k-major; column lr = gate*4 + unit)
  float* hs = lp_sm + (size_t)H * 16;             // [H][32]  h_{t-1} of the clip group, k-major
  float* part = hs + (size_t)H * LP_B;            // [8][16][32] partial tiles
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j0 = blockIdx.x * 4;
  for (int i = tid; i < 16 * H; i += 256) {
    const int lr = i / H, k = i - lr * H;
    Wt[k * 16 + lr] = whh[(long long)((lr >> 2) * H + j0 + (lr & 3)) * H + k];
  }
  const int kslice = H / LP_WARPS, k_lo = warp * kslice;
  const int ri = lane >> 3, ci = lane & 7;        // thread tile: rows 4ri..4ri+3, clips 4ci..4ci+3
  const int gu = tid >> 5, gb = tid & 31;         // gate threads (tid < 128): unit gu, clip gb of the group
  const int ngroups = (B + LP_B - 1) / LP_B;
  unsigned int epoch = 0;
  for (int g = 0; g < ngroups; ++g) {
    const int bg = g * LP_B + gb;
    float c = 0.f;
    float* hbuf0 = hglob + (size_t)g * 2 * H * LP_B;
    for (int t = 0; t < T; ++t) {
      // gate pre-activations of this step: independent of h, requested before the barrier wait
      float p4[4] = {0.f, 0.f, 0.f, 0.f};
      float sk = 0.f;
      if (tid < 128 && bg < B) {
        const long long pbase = (long long)bg * 4 * H * T + t;
#pragma unroll
        for (int q = 0; q < 4; ++q) p4[q] = __ldg(pre + pbase + (long long)(q * H + j0 + gu) * T);
        if (skip) sk = __ldg(skip + ((long long)bg * H + j0 + gu) * T + t);
      }
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      if (t > 0) {
        lp_grid_barrier(counter, (++epoch) * gridDim.x);     // h_{t-1} of every unit is published
        const float4* src = reinterpret_cast<const float4*>(hbuf0 + (size_t)((t - 1) & 1) * H * LP_B);
        float4* dst = reinterpret_cast<float4*>(hs);
#pragma unroll 8
        for (int i = tid; i < H * LP_B / 4; i += 256) dst[i] = __ldcg(src + i);
        __syncthreads();
        const float4* w4 = reinterpret_cast<const float4*>(Wt) + ri;
        const float4* h4 = reinterpret_cast<const float4*>(hs) + ci;
#pragma unroll 4
        for (int k = k_lo; k < k_lo + kslice; ++k) {
          const float4 w = w4[k * 4], hv = h4[k * 8];
          acc[0][0] += w.x * hv.x; acc[0][1] += w.x * hv.y; acc[0][2] += w.x * hv.z; acc[0][3] += w.x * hv.w;
          acc[1][0] += w.y * hv.x; acc[1][1] += w.y * hv.y; acc[1][2] += w.y * hv.z; acc[1][3] += w.y * hv.w;
          acc[2][0] += w.z * hv.x; acc[2][1] += w.z * hv.y; acc[2][2] += w.z * hv.z; acc[2][3] += w.z * hv.w;
          acc[3][0] += w.w * hv.x; acc[3][1] += w.w * hv.y; acc[3][2] += w.w * hv.z; acc[3][3] += w.w * hv.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(part + ((size_t)warp * 16 + 4 * ri + i) * LP_B + 4 * ci) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      __syncthreads();
      if (tid < 128) {
        float gsum[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a = 0.f;
#pragma unroll
          for (int w = 0; w < LP_WARPS; ++w) a += part[((size_t)w * 16 + q * 4 + gu) * LP_B + gb];
          gsum[q] = a + p4[q];
        }
        c = sigmoid_f(gsum[1]) * c + sigmoid_f(gsum[0]) * tanhf(gsum[2]);
        const float hv = sigmoid_f(gsum[3]) * tanhf(c);
        hbuf0[(size_t)(t & 1) * H * LP_B + (size_t)(j0 + gu) * LP_B + gb] = hv;
        if (bg < B) y[((long long)bg * H + j0 + gu) * T + t] = skip ? hv + sk : hv;
      }
      // `part` and `hs` are rewritten only after the next step's barrier (which begins with __syncthreads)
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ RVQ
// 8 frames per CTA, 256 threads; residual kept in smem across the n_q stages
constexpr int RVQ_FR = 8;
// embed_t [n_q][D][bins] is the transposed codebook: lanes (consecutive codewords) read consecutive addresses.  (Reading codeword
// rows from embed [bins][D] made every warp load touch 32 lines: 670 us at config 2; the sums below are the same fma chains in the
// same order, so the codes are bit-identical.)
__global__ void __launch_bounds__(256) rvq_encode_kernel(const float* __restrict__ z, const float* __restrict__ embed,
                                                         const float* __restrict__ embed_t, const float* __restrict__ embed_sq, int n_q,
                                                         int bins, int D, int B, int F, float* __restrict__ quantized,
                                                         long long* __restrict__ codes) {
  extern __shared__ __align__(16) float sm[];
  float* res = sm;                    // [FR][D]
  float* outv = sm + RVQ_FR * D;      // [FR][D]
  __shared__ float xx[RVQ_FR];
  __shared__ float bestv[RVQ_FR][8];
  __shared__ int besti[RVQ_FR][8];
  __shared__ int chosen[RVQ_FR];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long f0 = (long long)blockIdx.x * RVQ_FR, NF = (long long)B * F;
  for (int i = tid; i < RVQ_FR * D; i += 256) {
    const int f = i / D, d = i - f * D;
    const long long gf = f0 + f;
    float v = 0.f;
    if (gf < NF) { const long long bb = gf / F, ff = gf - bb * F; v = z[(bb * D + d) * F + ff]; }
    res[i] = v; outv[i] = 0.f;
  }
  __syncthreads();
  for (int q = 0; q < n_q; ++q) {
    const float* E = embed + (long long)q * bins * D;
    const float* ES = embed_sq + (long long)q * bins;
    if (warp < RVQ_FR) {   // ||x||^2 per frame (x.pow(2).sum(1))
      float s = 0.f;
      for (int d = lane; d < D; d += 32) s += res[warp * D + d] * res[warp * D + d];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) xx[warp] = s;
    }
    __syncthreads();
    float bv[RVQ_FR]; int bi[RVQ_FR];
#pragma unroll
    for (int f = 0; f < RVQ_FR; ++f) { bv[f] = -INFINITY; bi[f] = 0x7fffffff; }
    const float* ET = embed_t + (long long)q * bins * D;
    for (int jc0 = tid; jc0 < bins; jc0 += 1024) {           // four codewords per thread per sweep over d: jc0 + 256 c
      float dot[4][RVQ_FR];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int f = 0; f < RVQ_FR; ++f) dot[c][f] = 0.f;
      for (int d4 = 0; d4 < D / 4; ++d4) {
        float4 r[RVQ_FR];
#pragma unroll
        for (int f = 0; f < RVQ_FR; ++f) r[f] = *reinterpret_cast<const float4*>(res + f * D + 4 * d4);
        float e[4][4];
#pragma unroll
        for (int dd = 0; dd < 4; ++dd)
#pragma unroll
          for (int c = 0; c < 4; ++c) e[dd][c] = (jc0 + 256 * c < bins) ? __ldg(ET + (long long)(4 * d4 + dd) * bins + jc0 + 256 * c) : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int f = 0; f < RVQ_FR; ++f) {
            dot[c][f] = fmaf(r[f].x, e[0][c], dot[c][f]); dot[c][f] = fmaf(r[f].y, e[1][c], dot[c][f]);
            dot[c][f] = fmaf(r[f].z, e[2][c], dot[c][f]); dot[c][f] = fmaf(r[f].w, e[3][c], dot[c][f]);
          }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jc = jc0 + 256 * c;
        if (jc < bins) {
          const float ee = ES[jc];
#pragma unroll
          for (int f = 0; f < RVQ_FR; ++f) {
            const float dist = -((xx[f] - 2.f * dot[c][f]) + ee);      // core_vq.py:176-180
            if (dist > bv[f]) { bv[f] = dist; bi[f] = jc; }            // ascending jc per thread: first max kept
          }
        }
      }
    }
#pragma unroll
    for (int f = 0; f < RVQ_FR; ++f) {
      float v = bv[f]; int ix = bi[f];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
        if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
      }
      if (lane == 0) { bestv[f][warp] = v; besti[f][warp] = ix; }
    }
    __syncthreads();
    if (tid < RVQ_FR) {
      float v = bestv[tid][0]; int ix = besti[tid][0];
      for (int w = 1; w < 8; ++w) {
        const float ov = bestv[tid][w]; const int oi = besti[tid][w];
        if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
      }
      chosen[tid] = ix;
      const long long gf = f0 + tid;
      if (codes && gf < NF) codes[(long long)q * NF + gf] = ix;
    }
    __syncthreads();
    for (int i = tid; i < RVQ_FR * D; i += 256) {
      const int f = i / D, d = i - f * D;
      const float qv = E[(long long)chosen[f] * D + d];
      res[i] = res[i] - qv;          // residual = residual - quantized   (core_vq.py:334)
      outv[i] = outv[i] + qv;        // quantized_out = quantized_out + quantized (:335), starts from 0.0
    }
    __syncthreads();
  }
  if (quantized) {
    for (int i = tid; i < RVQ_FR * D; i += 256) {
      const int d = i / RVQ_FR, f = i - d * RVQ_FR;     // f fastest -> contiguous F in the NCL output
      const long long gf = f0 + f;
      if (gf < NF) { const long long bb = gf / F, ff = gf - bb * F; quantized[(bb * D + d) * F + ff] = outv[f * D + d]; }
    }
  }
}

// quantized[b][d][f] = sum_q embed[q][codes[q][b][f]][d]   (sum order q = 0.. as core_vq.py:356-362)
__global__ void rvq_decode_kernel(const long long* __restrict__ codes, const float* __restrict__ embed, int n_q, int bins, int D, int B,
                                  int F, float* __restrict__ quantized) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * D * F;
  if (i >= total) return;
  const int f = (int)(i % F);
  const int d = (int)((i / F) % D);
  const long long b = i / ((long long)F * D);
  float s = 0.f;
  for (int q = 0; q < n_q; ++q) {
    long long ix = codes[((long long)q * B + b) * F + f];
    ix = ix < 0 ? 0 : (ix >= bins ? bins - 1 : ix);
    s = s + embed[((long long)q * bins + ix) * D + d];
  }
  quantized[i] = s;
}

__global__ void rowsq_kernel(const float* __restrict__ e, float* __restrict__ sq, int rows, int D) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s += e[(long long)r * D + d] * e[(long long)r * D + d];
  sq[r] = s;
}

// ------------------------------------------------------------------ per-clip normalisation (sample.py:129,133-134)
__device__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}
__device__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s = fmaxf(s, red[w]);
  return s;
}
__global__ void __launch_bounds__(1024) normalize_clips_kernel(float* __restrict__ x, long long n, int mode) {
  __shared__ float red[32];
  float* p = x + (long long)blockIdx.x * n;
  float mx = 0.f, s = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) { const float v = p[i]; mx = fmaxf(mx, fabsf(v)); s += v; }
  mx = block_max(mx, red);
  if (mode == 0 || mode == 2) {
    const float d = mx + (mode == 0 ? 1e-8f : 1e-20f);   // sample.py:129 / unet.py:401-403
    for (long long i = threadIdx.x; i < n; i += blockDim.x) p[i] = p[i] / d;
    return;
  }
  const float mean = block_sum(s, red) / (float)n;
  float q = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) { const float dlt = p[i] - mean; q += dlt * dlt; }
  const float var = block_sum(q, red) / (float)(n - 1);      // torch.std: unbiased
  const float d1 = sqrtf(var) + 1e-8f;
  const float d2 = mx / d1 + 1e-8f;                          // max|x / d1| = max|x| / d1
  for (long long i = threadIdx.x; i < n; i += blockDim.x) p[i] = (p[i] / d1) / d2;
}

// ------------------------------------------------------------------ load-time folds
__global__ void weight_norm_fold_kernel(const float* __restrict__ g, const float* __restrict__ v, float* __restrict__ w, int inner) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float* vr = v + (long long)r * inner;
  float s = 0.f;
  for (int i = threadIdx.x; i < inner; i += blockDim.x) s += vr[i] * vr[i];
  const float nrm = sqrtf(block_sum(s, red));
  const float sc = g[r] / nrm;
  for (int i = threadIdx.x; i < inner; i += blockDim.x) w[(long long)r * inner + i] = vr[i] * sc;
}

__global__ void convtr_pack_kernel(const float* __restrict__ w, float* __restrict__ w2, int Cin, int Cout, int s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)s * Cout * Cin * 2;
  if (i >= total) return;
  const int kk = (int)(i & 1);
  const int ci = (int)((i >> 1) % Cin);
  const long long v = (i >> 1) / Cin;           // ph*Cout + co
  const int ph = (int)(v / Cout), co = (int)(v % Cout);
  const int tap = kk == 0 ? ph + s : ph;        // kk=0 multiplies x[i-1], kk=1 multiplies x[i]
  w2[i] = w[((long long)ci * Cout + co) * (2 * s) + tap];
}

__global__ void add_vec_kernel(const float* a, const float* b, float* c, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = a[i] + b[i];
}

}  // namespace

template <int RC>
static int conv1d_v2_dispatch(const ConvF32Args& a, int KT, int S, int B, cudaStream_t st) {
  constexpr int CO_T = 16 * RC;
  const int XWP = (128 + KT - 1 + 3) & ~3;
  int CI_T = KT == 1 ? 32 : (KT == 2 ? 24 : (KT == 3 ? 16 : 8));
  CI_T = (CI_T / S) * S;
  if (CI_T < S) CI_T = S;
  if (CI_T > a.Cin * S) CI_T = a.Cin * S;
  size_t smem = (size_t)(CI_T * XWP + CI_T * KT * CO_T) * sizeof(float);
  const size_t epi = (size_t)8 * RC * 132 * sizeof(float);
  if (epi > smem) smem = epi;
  LADIFF_REQUIRE(smem <= 100 * 1024, LADIFF_ERR_ARG, "conv1d_f32_v2: smem %zu", smem);
  dim3 grid(cdiv(a.LoutV, 128), cdiv(a.CoutV, CO_T), B);
  // exact for every dividend the fill produces (< 2^16): ceil(2^32 / d) * n >> 32 == n / d  while n * d < 2^32
  const unsigned span = (unsigned)(128 + KT - 1) * S;
  const unsigned magic_span = (unsigned)((0x100000000ULL + span - 1) / span), magic_s = (unsigned)((0x100000000ULL + S - 1) / S);
#define LADIFF_V2_CASE(KTV)                                                                                                   \
  case KTV: {                                                                                                                 \
    static unsigned long long attr = 0;                                                                                       \
    if (ladiff_first_on_device(&attr)) {                                                                                      \
      LADIFF_CUDA_OK(cudaFuncSetAttribute(conv1d_f32_v2_kernel<RC, KTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); \
    }                                                                                                                         \
    conv1d_f32_v2_kernel<RC, KTV><<<grid, 256, smem, st>>>(a, CI_T, S, magic_span, magic_s);                                                       \
    break;                                                                                                                    \
  }
  switch (KT) {
    LADIFF_V2_CASE(1)
    LADIFF_V2_CASE(2)
    LADIFF_V2_CASE(3)
    LADIFF_V2_CASE(7)
    default: LADIFF_REQUIRE(false, LADIFF_ERR_ARG, "conv1d_f32_v2: KT=%d", KT);
  }
#undef LADIFF_V2_CASE
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int C, int TT, int NTH>
static int seanet_resblock_launch_t(const float* x, float* y, const float* w1t, const float* b1, const float* wsct, const float* bsc,
                                    const float* w2t, const float* b2, int B, int L, cudaStream_t st) {
  constexpr int HID = C / 2, XP = TT + 8;
  const size_t smem = (size_t)(C * XP + C * TT + HID * TT + C * 3 * HID + C * C + HID * C) * sizeof(float);
  static unsigned long long attr = 0;
  if (ladiff_first_on_device(&attr))
    LADIFF_CUDA_OK(cudaFuncSetAttribute(seanet_resblock_kernel<C, TT, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  seanet_resblock_kernel<C, TT, NTH><<<dim3(cdiv(L, TT), B), NTH, smem, st>>>(x, y, w1t, b1, wsct, bsc, w2t, b2, L);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
// Fused residual block; returns 1 if this shape has no fused kernel (caller falls back to the three-launch form).
int seanet_resblock_launch(const float* x, float* y, const float* w1t, const float* b1, const float* wsct, const float* bsc, const float* w2t,
                           const float* b2, int C, int B, int L, cudaStream_t st) {
  static const bool off = getenv("LADIFF_NO_FUSED_RESBLOCK") != nullptr;
  if (off || L % 4 != 0 || L < 4 || ((uintptr_t)x % 16) != 0 || ((uintptr_t)y % 16) != 0) return 1;
  if (C == 32) return seanet_resblock_launch_t<32, 256, 256>(x, y, w1t, b1, wsct, bsc, w2t, b2, B, L, st);
  if (C == 64) return seanet_resblock_launch_t<64, 256, 512>(x, y, w1t, b1, wsct, bsc, w2t, b2, B, L, st);
  return 1;
}

int conv1d_f32_launch(const ConvF32Args& a, int B, cudaStream_t st) {
  LADIFF_REQUIRE(a.K >= 1 && a.K <= 64 && a.stride >= 1, LADIFF_ERR_ARG, "conv1d_f32: K=%d stride=%d", a.K, a.stride);
  static const bool no_out1 = getenv("LADIFF_CODEC_NO_OUT1") != nullptr;
  if (!no_out1 && a.CoutV == 1 && a.K == 7 && a.stride == 1 && a.pad_reflect && a.padL == 6 && a.il_s == 0 && !a.res && a.Cin == 32 &&
      a.LoutV == a.Lin && a.Lin > 8) {          // the decoder's output conv
    const size_t smem = (size_t)(32 * (512 + 8) + 32 * 8) * sizeof(float);
    static unsigned long long attr = 0;
    if (ladiff_first_on_device(&attr))
      LADIFF_CUDA_OK(cudaFuncSetAttribute(conv_out1_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_out1_kernel<32><<<dim3(cdiv(a.Lin, 512), B), 128, smem, st>>>(a.x, a.w, a.bias, a.y, a.Lin, a.act_in);
    LADIFF_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (a.tcw) {                                 // tcgen05 3xTF32 kernel (codec_tc.cu) where the weights were prepared for it
    const int rc = codec_tc_launch(a, B, st);
    if (rc <= 0) return rc;
  }
  static const bool no_v2 = getenv("LADIFF_CODEC_V1") != nullptr;
  if (a.wt && !no_v2 && a.K % a.stride == 0) {
    const int S = a.stride, KT = a.K / S;
    const bool il_ok = a.il_s == 0 || (a.il_s <= 8 && (8 % a.il_s) == 0);     // phases of one channel live in one thread tile / pass
    if ((KT == 1 || KT == 2 || KT == 3 || KT == 7) && S <= 8 && il_ok) {
      // channels per thread (RC; CTA = 16 RC channels x 128 positions): the widest tile the channel count allows, narrowed while the
      // grid would leave SMs idle (short time axes: the 120-frame layers of the encoder launch 32-128 CTAs with RC = 8)
      int rc = a.CoutV >= 128 ? 8 : (a.CoutV >= 64 ? 4 : (a.CoutV >= 32 && a.il_s <= 4 ? 2 : 1));
      static const bool no_narrow = getenv("LADIFF_CODEC_NO_NARROW") != nullptr;
      const long tiles_t = cdiv(a.LoutV, 128);
      // (stride-1 convs only: a strided conv's phase-channel fill is repeated by every channel tile and dominates when narrowed)
      while (!no_narrow && S == 1 && rc > 2 && tiles_t * cdiv(a.CoutV, 16 * rc) * B < 2L * tc_num_sms()) rc >>= 1;
      if (rc == 2 && !(a.CoutV >= 32 && a.il_s <= 4)) rc = 4;
      if (rc == 8) return conv1d_v2_dispatch<8>(a, KT, S, B, st);
      if (rc == 4) return conv1d_v2_dispatch<4>(a, KT, S, B, st);
      if (rc == 2) return conv1d_v2_dispatch<2>(a, KT, S, B, st);
      if (a.il_s <= 1) return conv1d_v2_dispatch<1>(a, KT, S, B, st);
    }
  }
  int CI_T = 64 / a.K;
  if (CI_T > 32) CI_T = 32;
  if (CI_T < 1) CI_T = 1;
  if (CI_T > a.Cin) CI_T = a.Cin;
  const int variant = a.CoutV <= 16 ? 2 : (a.CoutV <= 32 ? 1 : 0);
  const int RC = variant == 0 ? 4 : (variant == 1 ? 2 : 1);
  const int RT = variant == 0 ? 4 : 8;
  const int CO_T = 16 * RC, T_T = 16 * RT;
  const int XW = (T_T - 1) * a.stride + a.K;
  const size_t smem = (size_t)(((CI_T * XW + 3) & ~3) + CI_T * a.K * (CO_T + 4)) * sizeof(float);
  LADIFF_REQUIRE(smem <= 48 * 1024, LADIFF_ERR_ARG, "conv1d_f32: smem %zu", smem);
  dim3 grid(cdiv(a.LoutV, T_T), cdiv(a.CoutV, CO_T), B);
  if (variant == 0) conv1d_f32_kernel<4, 4><<<grid, 256, smem, st>>>(a, CI_T, XW);
  else if (variant == 1) conv1d_f32_kernel<2, 8><<<grid, 256, smem, st>>>(a, CI_T, XW);
  else conv1d_f32_kernel<1, 8><<<grid, 256, smem, st>>>(a, CI_T, XW);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int H>
static int lstm_cluster_launch(const float* pre, const float* whh, const float* skip, float* y, int B, int T, cudaStream_t st) {
  constexpr int CL = H / 32;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(B * CL); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  LADIFF_CUDA_OK(cudaLaunchKernelEx(&cfg, lstm_cluster_kernel<H>, pre, whh, skip, y, T));
  return 0;
}

int lstm_seq_launch(const float* pre, const float* whh, const float* skip, float* y, int B, int H, int T, cudaStream_t st) {
  static const bool no_cluster = getenv("LADIFF_LSTM_NO_CLUSTER") != nullptr;
  if (!no_cluster && H == 64) return lstm_cluster_launch<64>(pre, whh, skip, y, B, T, st);
  if (!no_cluster && H == 128) return lstm_cluster_launch<128>(pre, whh, skip, y, B, T, st);
  if (H == 64) {
    lstm_seq_kernel<64, 64><<<B, 256, 0, st>>>(pre, whh, skip, y, T);
  } else if (H == 128) {
    static unsigned long long attr = 0;
    const size_t smem = (size_t)64 * 512 * sizeof(float);
    if (ladiff_first_on_device(&attr)) {
      LADIFF_CUDA_OK(cudaFuncSetAttribute(lstm_seq_kernel<128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    lstm_seq_kernel<128, 64><<<B, 512, smem, st>>>(pre, whh, skip, y, T);
  } else {
    LADIFF_REQUIRE(false, LADIFF_ERR_ARG, "lstm_seq: H=%d unsupported", H);
  }
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int lstm_persist_launch(const float* pre, const float* whh, const float* skip, float* y, float* hglob, unsigned int* counter, int B, int H,
                        int T, cudaStream_t st) {
  LADIFF_REQUIRE(H % 32 == 0 && H >= 32, LADIFF_ERR_ARG, "lstm_persist: H=%d", H);
  const size_t smem = ((size_t)H * 16 + (size_t)H * LP_B + (size_t)LP_WARPS * 16 * LP_B) * sizeof(float);
  static unsigned long long attr = 0;
  if (ladiff_first_on_device(&attr)) {
    LADIFF_CUDA_OK(cudaFuncSetAttribute(lstm_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  LADIFF_REQUIRE(smem <= 200 * 1024 && H / 4 <= tc_num_sms(), LADIFF_ERR_ARG, "lstm_persist: H=%d does not fit one co-resident grid", H);
  LADIFF_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
  void* args[] = {(void*)&pre, (void*)&whh, (void*)&skip, (void*)&y, (void*)&hglob, (void*)&counter, (void*)&B, (void*)&H, (void*)&T};
  LADIFF_CUDA_OK(cudaLaunchCooperativeKernel((const void*)lstm_persist_kernel, dim3(H / 4), dim3(256), args, smem, st));
  return 0;
}
size_t lstm_persist_scratch_floats(int B, int H) { return (size_t)((B + LP_B - 1) / LP_B) * 2 * H * LP_B + 64; }

int lstm_steps_launch(const float* pre, const float* whh, const float* skip, float* y, float* hbuf, float* cbuf, int B, int H, int T,
                      cudaStream_t st, long long* launches) {
  LADIFF_REQUIRE(H % 4 == 0, LADIFF_ERR_ARG, "lstm_steps: H=%d", H);
  LADIFF_CUDA_OK(cudaMemsetAsync(hbuf, 0, sizeof(float) * 2 * B * H, st));
  LADIFF_CUDA_OK(cudaMemsetAsync(cbuf, 0, sizeof(float) * B * H, st));
  dim3 grid(H / 4, cdiv(B, 32));
  for (int t = 0; t < T; ++t) {
    const float* hp = hbuf + (size_t)(t & 1) * B * H;
    float* hn = hbuf + (size_t)((t + 1) & 1) * B * H;
    lstm_step_kernel<<<grid, 128, 0, st>>>(pre, whh, skip, y, hp, hn, cbuf, B, H, T, t);
  }
  LADIFF_CUDA_OK(cudaGetLastError());
  if (launches) *launches += T;
  return 0;
}

__global__ void rvq_transpose_kernel(const float* __restrict__ e, float* __restrict__ et, int n_q, int bins, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_q * bins * D) return;
  const int j = (int)(i % bins);
  const long long r = i / bins;
  const int d = (int)(r % D), q = (int)(r / D);
  et[i] = e[((long long)q * bins + j) * D + d];
}
int rvq_transpose_launch(const float* embed, float* embed_t, int n_q, int bins, int D, cudaStream_t st) {
  const long long total = (long long)n_q * bins * D;
  rvq_transpose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(embed, embed_t, n_q, bins, D);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int rvq_encode_launch(const float* z, const float* embed, const float* embed_t, const float* embed_sq, int n_q, int bins, int D, int B, int F,
                      float* quantized, long long* codes, cudaStream_t st) {
  LADIFF_REQUIRE(D % 4 == 0 && D <= 512, LADIFF_ERR_ARG, "rvq: D=%d", D);
  const long long NF = (long long)B * F;
  const size_t smem = (size_t)2 * RVQ_FR * D * sizeof(float);
  rvq_encode_kernel<<<(unsigned)((NF + RVQ_FR - 1) / RVQ_FR), 256, smem, st>>>(z, embed, embed_t, embed_sq, n_q, bins, D, B, F, quantized, codes);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int rvq_decode_launch(const long long* codes, const float* embed, int n_q, int bins, int D, int B, int F, float* quantized,
                      cudaStream_t st) {
  const long long total = (long long)B * D * F;
  rvq_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(codes, embed, n_q, bins, D, B, F, quantized);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int rowsq_launch(const float* e, float* sq, int rows, int D, cudaStream_t st) {
  rowsq_kernel<<<cdiv(rows, 128), 128, 0, st>>>(e, sq, rows, D);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int normalize_clips_launch(float* x, int B, long long n, int mode, cudaStream_t st) {
  normalize_clips_kernel<<<B, 1024, 0, st>>>(x, n, mode);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int weight_norm_fold_launch(const float* g, const float* v, float* w, int rows, int inner, cudaStream_t st) {
  weight_norm_fold_kernel<<<rows, 128, 0, st>>>(g, v, w, inner);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int conv_w_transpose_launch(const float* w, float* wt, int CoutV, int Cin, int K, int S, int perm_s, int perm_cout, cudaStream_t st) {
  LADIFF_REQUIRE(S >= 1 && K % S == 0, LADIFF_ERR_ARG, "conv_w_transpose: K=%d S=%d", K, S);
  const long long total = (long long)CoutV * Cin * K;
  conv_w_transpose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, wt, CoutV, Cin, K, S, perm_s, perm_cout);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int convtr_pack_launch(const float* w, float* w2, int Cin, int Cout, int s, cudaStream_t st) {
  const long long total = (long long)s * Cout * Cin * 2;
  convtr_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, w2, Cin, Cout, s);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int add_vec_launch(const float* a, const float* b, float* c, int n, cudaStream_t st) {
  add_vec_kernel<<<cdiv(n, 256), 256, 0, st>>>(a, b, c, n);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
