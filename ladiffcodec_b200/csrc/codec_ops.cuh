// fp32 NCL kernels of the SEANet codec path: Conv1d / ConvTranspose1d, LSTM, RVQ, clip normalisation.
#pragma once
#include "common.cuh"

// y[b,co,t] = bias[co] + sum_{ci,k} w[co,ci,k] * act(xpad[b,ci,t*stride + k - padL])  (+ res[b,co,t])
// xpad: reflect (SConv1d, conv.py:217-232) or zero extension of x outside [0,Lin).
// Weights of one conv prepared for the tcgen05 3xTF32 kernel (codec_tc.cu): hi / lo planes [2][KT][CoutV][CinV] and their tensor map.
struct CodecTcWeights {
  CUtensorMap tm;
  int CinV = 0, KT = 0, CoutV = 0, valid = 0;
};
struct ConvF32Args {
  const float* x; int Cin; int Lin;
  const float* w;          // [CoutV][Cin][K]
  const float* wt;         // optional K-major copy for the register-tiled kernel (conv_w_transpose_launch); null: generic kernel
  const float* bias;       // [Cout real] or null
  float* y;
  int CoutV;               // (virtual) output channels computed
  int LoutV;               // (virtual) output positions computed
  int K, stride, padL;
  int pad_reflect;         // 1 reflect, 0 zeros
  int act_in;              // 0 none, 1 ELU(alpha=1)
  const float* res;        // optional residual, same layout as y
  // transposed-conv interleave (SConvTranspose1d, conv.py:252-274): virtual channel v = ph*il_cout + co at
  // virtual position i lands at y[b, co, i*il_s + ph - il_trim] if inside [0, il_lout).  il_s = 0: plain conv.
  int il_s, il_cout, il_trim, il_lout;
  const CodecTcWeights* tcw;   // optional: prepared tensor-core weights (null: FMA kernels only)
  int tc_precise;              // 1: this conv feeds the vector quantiser (cond encoder): more accumulator chains (codec_tc.cu)
};
// planes: device buffer of 2*CinV*KT*CoutV floats (filled here from wt = the K-major copy of conv_w_transpose_launch)
int codec_tc_prepare(const float* wt, int CinV, int KT, int CoutV, float* planes, CodecTcWeights* out, cudaStream_t st);
// 0: launched; 1: this conv has no tensor-core form (caller falls back to the FMA kernel); < 0: error
int codec_tc_launch(const ConvF32Args& a, int B, cudaStream_t st);
int conv1d_f32_launch(const ConvF32Args& a, int B, cudaStream_t st);
// Fused SEANet residual block (seanet.py:45-63) for C = 32 / 64: y = shortcut_1x1(x) + conv_k1(ELU(conv_k3(ELU(x)))), causal reflect
// padding; w1t [C*3][C/2], wsct [C][C], w2t [C/2][C] are the K-major copies (conv_w_transpose_launch).  Returns 0 when launched,
// 1 when the shape has no fused kernel (caller uses three conv launches), < 0 on error.
int seanet_resblock_launch(const float* x, float* y, const float* w1t, const float* b1, const float* wsct, const float* bsc, const float* w2t,
                           const float* b2, int C, int B, int L, cudaStream_t st);
// wt[((ci*S + p)*KT + kt)][col(co)] = w[co][ci][kt*S + p], S = stride phases (K = KT*S), for the register-tiled kernel.
// perm_s > 0 (transposed-conv virtual channels co = ph*perm_cout + cr): col = cr*perm_s + ph (phase fastest), else col = co.
int conv_w_transpose_launch(const float* w, float* wt, int CoutV, int Cin, int K, int S, int perm_s, int perm_cout, cudaStream_t st);

// nn.LSTM layer recurrence given pre[b][4H][T] = W_ih x_t + b_ih + b_hh (PyTorch gate order i,f,g,o; lstm.py:20)
//   y[b][j][t] = h_t[j] (+ skip[b][j][t]).  small H (64/128): one persistent CTA per clip.
int lstm_seq_launch(const float* pre, const float* whh, const float* skip, float* y, int B, int H, int T, cudaStream_t st);
// any H % 4 == 0: one launch per time step; hbuf [2][B][H], cbuf [B][H] scratch (zeroed here)
int lstm_steps_launch(const float* pre, const float* whh, const float* skip, float* y, float* hbuf, float* cbuf, int B, int H, int T,
                      cudaStream_t st, long long* launches);

// large H (encoder, H = 512): one persistent cooperative kernel per layer; hglob: lstm_persist_scratch_floats(B, H) floats
// (h ping-pong per 32-clip group, k-major), counter: one zero-initialised-here unsigned int
int lstm_persist_launch(const float* pre, const float* whh, const float* skip, float* y, float* hglob, unsigned int* counter, int B, int H,
                        int T, cudaStream_t st);
size_t lstm_persist_scratch_floats(int B, int H);

// Residual VQ (core_vq.py:174-189, 324-362).  z [B][D][F] -> quantized [B][D][F] (nullable), codes [n_q][B][F] int64 (nullable)
// embed [n_q][bins][D], embed_sq [n_q][bins] = sum_d e^2
// embed_t [n_q][D][bins]: transposed copy (rvq_transpose_launch) for coalesced codeword sweeps
int rvq_encode_launch(const float* z, const float* embed, const float* embed_t, const float* embed_sq, int n_q, int bins, int D, int B, int F,
                      float* quantized, long long* codes, cudaStream_t st);
int rvq_transpose_launch(const float* embed, float* embed_t, int n_q, int bins, int D, cudaStream_t st);
int rvq_decode_launch(const long long* codes, const float* embed, int n_q, int bins, int D, int B, int F, float* quantized,
                      cudaStream_t st);
int rowsq_launch(const float* e, float* sq, int rows, int D, cudaStream_t st);

// sample.py:129 (mode 0) and :133-134 (mode 1), per clip
int normalize_clips_launch(float* x, int B, long long n, int mode, cudaStream_t st);

// load-time folds
// weight_norm (conv.py:30): w[o,:,:] = g[o] * v[o,:,:] / ||v[o,:,:]||      rows = dim-0 size, inner = product of the rest
int weight_norm_fold_launch(const float* g, const float* v, float* w, int rows, int inner, cudaStream_t st);
// ConvTranspose1d weight [Cin][Cout][2s] -> virtual conv weight [s*Cout][Cin][2] (see ConvF32Args)
int convtr_pack_launch(const float* w, float* w2, int Cin, int Cout, int s, cudaStream_t st);
int add_vec_launch(const float* a, const float* b, float* c, int n, cudaStream_t st);
