// Element-wise / normalisation / attention kernels of the UNet on channels-last h16 activations.
#pragma once
#include "common.cuh"

// channels-last h16 view: element (b, l, c) at p[b*bstride + l*pitch + c]
struct ClView {
  h16* p;
  long long bstride;
  int pitch;
  int C;
};

// GroupNorm(8) + FiLM + SiLU (+ residual) (+ tanh)            unet.py:145-154, 183-192, 466-467
struct GnApplyArgs {
  ClView y;            // conv output (h16)
  const float2* stats; // [B][n_ntiles][C/32] partial (sum, sumsq)
  int n_ntiles;
  const float* gamma;  // [C]
  const float* beta;   // [C]
  const float* film;   // FiLM table base + block offset, row stride film_stride; null = no FiLM (block2)
  long long film_stride;
  const int* t_dev;    // [B] step index per clip
  ClView res;          // residual to add after SiLU (p null = none)
  ClView out;
  int L;
  int do_tanh;
  // optional fused channel LayerNorm of the result (the attention block's pre-norm, unet.py:82-101): ln_out = LN(out) * ln_g
  const float* ln_g;   // [C] or null
  ClView ln_out;
};
int gn_apply_launch(const GnApplyArgs& a, int B, cudaStream_t st);

// channel LayerNorm (gain only) (+ residual)                    unet.py:82-101, 203-206
int layernorm_cl_launch(ClView x, const float* g, ClView res, ClView out, int B, int L, cudaStream_t st);

// LinearAttention core (between to_qkv and to_out)              unet.py:208-221
// qkv [B][L][384] h16 (q | k | v, each 4 heads x 32) -> out [B][L][128] h16; ctx scratch [B][4][32][32] f32;
// part scratch linattn_part_floats(B, L) f32; counters [4B] int32, zero before the first launch (the kernel re-zeroes them)
int linattn_launch(ClView qkv, float* ctx, float* part, int* counters, ClView out, int B, int L, cudaStream_t st);
size_t linattn_part_floats(int B, int L);
// the two halves separately: context only, and (csrc/attn_tc.cu) everything after it in one kernel — ctx^T softmax(q), the to_out 1x1
// conv on tcgen05, channel LayerNorm and the residual:  out = LN(W_out (ctx^T q~) + b) g + xres
int linattn_ctx_launch(ClView qkv, float* ctx, float* part, int* counters, int B, int L, cudaStream_t st);
// second half of the linear attention on tcgen05 (attn_tc.cu): out = softmax_d(q) ctx, split operands, fp32-kernel accuracy
int linattn_out_tc_launch(ClView qkv, const float* ctx, ClView out, int B, int L, cudaStream_t st);
int linattn_tail_launch(ClView qkv, const float* ctx, const h16* wout, const float* bias, const float* gain, ClView xres, ClView out, int B,
                        int L, int C, cudaStream_t st);
// Attention core (mid block)                                    unet.py:234-245
int fullattn_launch(ClView qkv, ClView out, int B, int L, cudaStream_t st);
int fullattn_launch_impl(ClView qkv, ClView out, int B, int L, int impl, cudaStream_t st);
// the same on the tensor cores (csrc/attn_tc.cu: tcgen05 QK^T and PV, flash-attention style); fullattn_launch picks it for long bottlenecks
int fullattn_tc_launch(ClView qkv, ClView out, int B, int L, cudaStream_t st);

// layout conversion / DDPM
// x NCL f32 [B][C][L] * scale[b] -> channels-last h16 view (channel offset via out.p)
int ncl_to_cl_launch(const float* x, const float* inv_scale /*[B] or null*/, ClView out, int B, int C, int L, cudaStream_t st);
// channels-last f32 [B][L][C] -> NCL f32
int cl_to_ncl_f32_launch(const float* x, float* y, int B, int C, int L, cudaStream_t st);
// per-clip reciprocal of (max|x| + eps):  inv[b] = 1 / (max + eps)
int absmax_inv_launch(const float* x, float* inv, int B, long long n, float eps, cudaStream_t st);

struct DdpmTables {   // device pointers to the (1000,) fp32 buffers of GaussianDiffusion1D (ddpm_loss.py:140-164)
  const float* sqrt_recip_ac;
  const float* sqrt_recipm1_ac;
  const float* coef1;
  const float* coef2;
  const float* logvar;
};
// One fused sampler step (see ddpm_step_kernel): x0 = clamp(a x - b eps, -1, 1);
//   mode 0 (DDPM): x <- (k0 x0 + k1 x) + ks z;  mode 1 (DDIM): x <- (k0 x0 + k1 eps) + ks z;  mode 2: x <- x0.
struct StepCoef { float a, b, k0, k1, ks; int mode; };
// eps channels-last f32 [B][L][C]; x NCL f32 in place; noise NCL f32 (step slice) or null -> Philox(seed, t_abs, clip_offset + b);
// also writes the new x as h16 into xin (channels-last view, channel offset applied by caller; p null = skip).
int ddpm_step_launch(const float* eps, float* x, const float* noise, unsigned long long seed, int t_abs, unsigned long long clip_offset,
                     StepCoef cf, ClView xin, int B, int C, int L, cudaStream_t st);
// x[0..n) <- N(0,1) (uniform = 0) or U[0,1) (uniform = 1) from Philox(seed, elem_offset + i)
int randn_fill_launch(float* x, long long n, unsigned long long seed, unsigned long long elem_offset, int uniform, cudaStream_t st);
// q_sample (ddpm_loss.py:387-393), per clip of n elements
int q_sample_launch(const float* x, const float* noise, const int* t_dev, const float* sqrt_ac, const float* sqrt_1m_ac, float* out, int B,
                    long long n, cudaStream_t st);
// x <- a x + b y (y null: x <- a x)
int axpby_launch(float* x, float a, const float* y, float b, long long n, cudaStream_t st);
// p_losses tail (ddpm_loss.py:404-437): pred_x0 / eps_ncl optional NCL outputs, part [B] scratch, loss [1]
int p_losses_launch(const float* eps, const float* noise, const float* xt, const int* t_dev, const float* recip, const float* recipm1,
                    const float* p2w, float* pred_x0, float* eps_ncl, float* part, float* loss, int B, int C, int L, cudaStream_t st);
// clamp(sd-sdr negative, min clip) per clip (losses_fn.py:54-66)
int sdsdr_launch(const float* est, const float* tgt, float* out, int B, long long n, float clip, cudaStream_t st);
int fill_t_launch(int* t_dev, int t, int B, cudaStream_t st);
int time_to_int_launch(const long long* time, int* t_dev, int B, cudaStream_t st);

// y[M][ldy](col0..) = act_out( act_in(x[M][K]) @ W[N][K]^T + b )     (load-time folds: time MLP, FiLM table)
int linear_f32_launch(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int M, int N, int K,
                      int act_in /*0 none, 1 silu*/, int act_out /*0 none, 1 gelu(erf)*/, cudaStream_t st);
int sinusoid_launch(float* emb, int T, int dim, cudaStream_t st);
