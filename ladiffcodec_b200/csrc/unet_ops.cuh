// Element-wise / normalisation / attention kernels of the UNet on channels-last bf16 activations.
#pragma once
#include "common.cuh"

// channels-last bf16 view: element (b, l, c) at p[b*bstride + l*pitch + c]
struct ClView {
  bf16* p;
  long long bstride;
  int pitch;
  int C;
};

// GroupNorm(8) + FiLM + SiLU (+ residual) (+ tanh)            unet.py:145-154, 183-192, 466-467
struct GnApplyArgs {
  ClView y;            // conv output (bf16)
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
// qkv [B][L][384] bf16 (q | k | v, each 4 heads x 32) -> out [B][L][128] bf16; ctx scratch [B][4][32][32] f32;
// part scratch linattn_part_floats(B, L) f32; counters [4B] int32, zero before the first launch (the kernel re-zeroes them)
int linattn_launch(ClView qkv, float* ctx, float* part, int* counters, ClView out, int B, int L, cudaStream_t st);
size_t linattn_part_floats(int B, int L);
// Attention core (mid block)                                    unet.py:234-245
int fullattn_launch(ClView qkv, ClView out, int B, int L, cudaStream_t st);

// layout conversion / DDPM
// x NCL f32 [B][C][L] * scale[b] -> channels-last bf16 view (channel offset via out.p)
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
// One fused posterior step (ddpm_loss.py:175-179,199-206,233-251):
//   x0 = clamp(a_t x - b_t eps, -1, 1); x <- c1_t x0 + c2_t x + exp(0.5 logvar_t) z   (z = 0 at t == 0)
// eps channels-last f32 [B][L][C]; x NCL f32 in place; noise NCL f32 (step slice) or null -> Philox(seed, step);
// also writes the new x as bf16 into xin (channels-last view, channel offset applied by caller).
int ddpm_step_launch(const float* eps, float* x, const float* noise, unsigned long long seed, int step_index,
                     const int* t_dev, DdpmTables tb, ClView xin, int B, int C, int L, cudaStream_t st);
int fill_t_launch(int* t_dev, int t, int B, cudaStream_t st);
int time_to_int_launch(const long long* time, int* t_dev, int B, cudaStream_t st);

// y[M][ldy](col0..) = act_out( act_in(x[M][K]) @ W[N][K]^T + b )     (load-time folds: time MLP, FiLM table)
int linear_f32_launch(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int M, int N, int K,
                      int act_in /*0 none, 1 silu*/, int act_out /*0 none, 1 gelu(erf)*/, cudaStream_t st);
int sinusoid_launch(float* emb, int T, int dim, cudaStream_t st);
