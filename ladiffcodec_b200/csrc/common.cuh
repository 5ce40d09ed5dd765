// Shared definitions for the ladiff_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ladiff_b200.h"

typedef __nv_bfloat16 bf16;

#define LADIFF_CUDA_OK(expr)                                                                        \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ladiff_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));        \
      return LADIFF_ERR_CUDA;                                                                       \
    }                                                                                               \
  } while (0)

#define LADIFF_REQUIRE(cond, code, ...)                                                             \
  do {                                                                                              \
    if (!(cond)) {                                                                                  \
      ladiff_set_error(__VA_ARGS__);                                                                \
      return (code);                                                                                \
    }                                                                                               \
  } while (0)

void ladiff_set_error(const char* fmt, ...);

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------ tcgen05 implicit-GEMM conv (tc_conv.cu)
// out[b, l, m] = bias[m] + sum over K-segments s, channels c:  W[m, kofs_s + c] * X[b, l + shift_s, ch0_s + c]
// X is a channels-last bf16 view [B][Lv][Cv] (row pitch / batch stride in elements), zero outside [0,Lv).
#define TC_MAX_SEG 8
#define TC_BM 128
#define TC_BK 64

struct TcSeg {
  int shift;    // position shift of this tap (in view rows)
  int ch0;      // first channel of this segment in the view
  int nchunk;   // number of 64-channel chunks
  int m_lo;     // segment contributes only to output-channel tiles m0 in [m_lo, m_hi)
  int m_hi;
};

struct TcConvParams {
  CUtensorMap tmW;  // weights  [Cout][Ktot] bf16 row-major, box {64, 128}, SWIZZLE_128B
  CUtensorMap tmX;  // activations view [B][Lv][Cv] bf16, box {64, NT, 1}, SWIZZLE_128B, OOB -> 0
  TcSeg seg[TC_MAX_SEG];
  int nseg;
  int NT;            // positions per tile (multiple of 16, 16..256)
  int stages;        // pipeline depth
  int Lout;          // valid output rows per clip
  int Cout;          // total output channels (multiple of 128)
  const float* bias; // [Cout] or null
  void* out;         // bf16 or f32, channels-last
  long long out_bstride;  // elements between clips
  int out_pitch;     // elements between rows
  int out_ch0;       // channel offset of m = 0
  int out_split;     // channels m >= out_split land at m + out_jump (nearest-x2 upsample interleave); 0 = off
  int out_jump;
  int out_f32;       // 1: float output
  float2* stats;     // optional GroupNorm partials [B][n_ntiles][Cout/32] (sum, sumsq) of the fp32 result
  const bf16* res;   // optional residual added in the epilogue (channels-last bf16, same rows/channels as the output)
  long long res_bstride;
  int res_pitch;
  int res_ch0;
};

// X view description used by the SIMT check kernel (same math, no tensor maps)
struct TcRefView {
  const bf16* x; long long bstride; int pitch; int Lv; int Cv;
  const bf16* w; int Ktot;
};

int tc_conv_launch(const TcConvParams& p, int B, cudaStream_t st);
int tc_conv_ref_launch(const TcConvParams& p, const TcRefView& v, int B, cudaStream_t st);
int tc_make_tmap_w(CUtensorMap* tm, const bf16* w, int Cout, int Ktot);
int tc_make_tmap_x(CUtensorMap* tm, const bf16* x, int B, int Lv, int Cv, int pitch, long long bstride, int NT);
int tc_pick_nt(int L, int* n_tiles);
size_t tc_smem_bytes(int NT, int stages);
int tc_pick_stages(int NT);
