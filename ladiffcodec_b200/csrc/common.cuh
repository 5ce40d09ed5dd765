// Shared definitions for the ladiff_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "../../include/ladiff_b200.h"

// ------------------------------------------------------------------ 16-bit operand / activation type of the UNet
// Tensor-core operands and stored activations are IEEE fp16 (11-bit significand): kind::f16 runs fp16 and bf16 at the same
// rate, and fp16 rounds 8x finer than bf16 (2^-11 vs 2^-8 relative) — one UNet evaluation agrees with the fp32 reference to
// ~1.5e-3 rel-L2 instead of 1.2e-2.  Every activation of this network sits behind a GroupNorm / LayerNorm / softmax or is a
// weight-standardised convolution of such a tensor, i.e. O(1)..O(100), far inside fp16's range; the conversions saturate
// (cvt.rn.satfinite) so an outlier can never become inf/NaN.  -DLADIFF_USE_BF16 builds the bf16 variant (A/B measurements).
#ifdef LADIFF_USE_BF16
typedef __nv_bfloat16 h16;
typedef __nv_bfloat162 h162;
#define TC_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define TC_IDESC_AB_FMT 1u          /* InstrDescriptor a_format / b_format: BF16 */
#define LADIFF_DTYPE_NAME "bf16"
#else
typedef __half h16;
typedef __half2 h162;
#define TC_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define TC_IDESC_AB_FMT 0u          /* F16 */
#define LADIFF_DTYPE_NAME "f16"
#endif
#ifdef __CUDACC__
__device__ __forceinline__ h16 f2h(float v) {
#ifdef LADIFF_USE_BF16
  return __float2bfloat16(v);
#else
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
#endif
}
__device__ __forceinline__ float h2f(h16 v) {
#ifdef LADIFF_USE_BF16
  return __bfloat162float(v);
#else
  return __half2float(v);
#endif
}
__device__ __forceinline__ h162 ff2h2(float a, float b) {      // (a, b) -> packed pair, a in the low half
#ifdef LADIFF_USE_BF16
  return __floats2bfloat162_rn(a, b);
#else
  unsigned int r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return *reinterpret_cast<h162*>(&r);
#endif
}
__device__ __forceinline__ float2 h22ff(h162 v) {
#ifdef LADIFF_USE_BF16
  return __bfloat1622float2(v);
#else
  return __half22float2(v);
#endif
}
__device__ __forceinline__ unsigned short h16_bits(h16 v) { return *reinterpret_cast<unsigned short*>(&v); }
__device__ __forceinline__ h16 h16_from_bits(unsigned short b) { return *reinterpret_cast<h16*>(&b); }
#endif

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

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Kernels of the UNet step are launched with programmatic stream serialization: kernel N+1 may start (block scheduling,
// barrier init, TMEM allocation, descriptor prefetch) while kernel N drains.  Every such kernel executes pdl_wait()
// before it touches global memory another kernel may have written or may still read, and pdl_trigger() as soon as its
// dependents may begin launching.  LADIFF_NO_PDL=1 falls back to plain stream order.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool ladiff_pdl_enabled();
// Every kernel of the step keeps the SM's shared-memory carve-out at its maximum, the configuration the conv kernel needs:
// a kernel that prefers a different L1/shared split cannot become resident next to a running conv CTA and forces an SM
// reconfiguration at every kernel boundary of the ~170-kernel step.
template <typename... KArgs>
static inline void prefer_max_smem_carveout(void (*kernel)(KArgs...)) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
// Function attributes are per device: one-time setup is keyed by the current device ordinal (a second GPU in the same
// process gets its own opt-in).
static inline bool ladiff_first_on_device(unsigned long long* mask) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long bit = 1ull << (d & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}
#define LADIFF_CARVEOUT_ONCE(kernel)                                             \
  do {                                                                           \
    static unsigned long long _done = 0;                                         \
    if (ladiff_first_on_device(&_done)) prefer_max_smem_carveout(kernel);        \
  } while (0)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ladiff_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------ tcgen05 implicit-GEMM conv (tc_conv.cu)
// out[b, l, m] = bias[m] + sum over groups g, taps t of g, channels c < 64*nchunk_g:
//                  W[m, kofs_t + c] * X[b, l + shift_g + row_off_t, ch0_g + c]
// X is a channels-last h16 view [B][Lv][Cv] (row pitch / clip stride in elements), zero outside [0,Lv).
// A group is one activation tile in shared memory (one TMA load per 64-channel chunk) that all of its taps read at
// different row offsets — a k-tap conv loads its activations once, not k times.
#define TC_MAX_GRP 7
#define TC_MAX_TAPS 7
#define TC_BM 128
#define TC_BK 64
#define TC_STAT_PARTS 2   // GroupNorm partials per (clip, position tile): one per epilogue warpgroup

struct TcTap {
  int row_off;  // row offset of this tap inside the group's shared-memory tile (0..7)
  int kofs;     // K offset of this tap's weights (chunk c adds 64*c)
  int m_lo;     // the tap contributes only to output-channel tiles m0 in [m_lo, m_hi)
  int m_hi;
};
struct TcGroup {
  int ch0;      // first channel of the group in the view
  int shift;    // view row of tile row 0 relative to the first output position of the tile
  int nchunk;   // 64-channel chunks
  int ntaps;
  TcTap tap[TC_MAX_TAPS];
};

enum { TC_KIND_PLAIN = 0, TC_KIND_DOWN = 1, TC_KIND_UP = 2 };

// What one conv launch computes (host-side description; tc_conv_plan turns it into launch parameters).
struct TcConvDesc {
  int kind;                 // TC_KIND_*: plain (odd k, zero pad (k-1)/2), stride-2 k=4 pad 1, nearest-x2 + k=3 folded
  int Cin, K;               // real input channels, taps
  int CoutV;                // output channels computed (2*Cout for TC_KIND_UP), multiple of 128
  const h16* w;            // packed [CoutV][Ktot]
  const CUtensorMap* tmW;   // box {64, 128}, SWIZZLE_128B
  int Ktot;
  const float* bias;        // [CoutV] or null
  const h16* x; long long x_bstride; int x_pitch; int Lin;   // input view (Lin rows of Cin channels per clip)
  h16* out; long long out_bstride; int out_pitch;            // h16 output view (Lout rows; 2*Lout rows for UP)
  float* out32;             // if set: fp32 output, contiguous [B][Lout][CoutV] (direct epilogue)
  float2* stats;            // optional GroupNorm partials [B][n_ptiles][TC_STAT_PARTS][CoutV/32]
  const h16* res; long long res_bstride; int res_pitch;      // optional residual (direct epilogue)
  // optional second output (ResnetBlock: res_conv fused into block1's conv): output channels m >= split_m are a 1x1 conv of the
  // same input (weights at K offset 0 of rows [split_m, CoutV)) written to out2; GroupNorm partials cover m < split_m only
  int split_m;
  h16* out2; long long out2_bstride; int out2_pitch;
  int B;
  int tap_share;            // 1: all taps of a chunk read one shared activation tile; 0: one tile load per tap
  int want_nt, want_nclip;  // > 0: force the tile shape (rows per clip region / clip regions per tile); 0: cost model
  int want_two_per_sm;      // 1: shape the launch so that two CTAs share an SM (<= 112 KB shared memory, <= 256 TMEM columns)
  int want_pair;            // 1: CTA pairs along N share every weight tile (each CTA loads half of it and multicasts to both)
  int want_transposed;      // 1: positions-on-M kernel (tc_conv_t_kernel): 128-row position tiles x <= 256-channel tiles
};

struct TcConvParams {
  CUtensorMap tmW;   // weights  [Cout][Ktot] h16 row-major, box {64, 128}, SWIZZLE_128B
  CUtensorMap tmX;   // activations view [B][Lv][Cv] h16, box {64, BOXROWS, 1}, SWIZZLE_128B, OOB -> 0
  CUtensorMap tmY;   // output {Cc, phases, rows, B} h16, box {32, 1, CR, 1} (one epilogue warp's channel slice)
  CUtensorMap tmYr;  // same, box rows = NT % CR (last chunk of a single-clip tile)
  CUtensorMap tmY2;  // second output (channels m >= split_m), box {32, 1, CR, 1}
  CUtensorMap tmY2r;
  TcGroup grp[TC_MAX_GRP];
  int ngrp;
  int NT;            // output rows per clip region of a tile
  int NCLIP;         // clip regions per tile (1 = single-clip position tiles)
  int BOXROWS;       // rows per activation TMA box
  int NMMA;          // MMA N = NCLIP*NT (multiple of 16, <= 256)
  int CR;            // rows per store chunk
  int n_ptiles;      // position tiles per clip (1 when NCLIP > 1)
  int n_ntiles;      // N tiles in total
  int MT;            // Cout/128
  int S;             // pipeline stages; a stage = a_cap weight tiles (16 KB each) + one activation tile
  int a_cap;
  int stage_bytes;
  int b_slot_bytes;
  int direct;        // 1: direct register->global epilogue (fp32 output and/or residual); 0: smem-staged TMA store
  int up_cout;       // TC_KIND_UP: real Cout (output phase = m0 / up_cout); else 0
  int split_m;       // second-output split (0 = none)
  int stat_slots;    // GroupNorm slots per tile row block: (split_m ? split_m : Cout) / 32
  int B, Lout, Cout;
  const float* bias; // [Cout] or null
  float2* stats;     // optional GroupNorm partials [B][n_ptiles][TC_STAT_PARTS][Cout/32] (sum, sumsq) of the fp32 result
  // direct epilogue only
  void* out;
  long long out_bstride;
  int out_pitch;
  int out_f32;
  const h16* res;
  long long res_bstride;
  int res_pitch;
  int pair;                   // 1: launched as clusters of 2 CTAs that take neighbouring position tiles of the same output-channel tile;
                              // each CTA TMA-loads half (64 rows) of every weight tile with .multicast::cluster into both CTAs
  CUtensorMap tmWh;           // weights, box {64, 64} (half a weight tile)
  int minb;                   // CTAs per SM the launch is shaped for (1 or 2): selects the kernel instantiation and the grid
  // positions-on-M variant (tc_conv_t_kernel): D[position, channel]; rows = 128 positions of one clip, columns = NCH channels
  int transposed;
  int NCH;                    // channels per tile (multiple of 16, <= 256, divides Cout and every tap's channel range)
  int n_chtiles;              // Cout / NCH
  int stat_parts;             // GroupNorm partials per (clip, position tile): 2 (epilogue warpgroups) or 4 (epilogue warps)
  CUtensorMap tmWt;           // weights, box {64, NCH}
  void* out2v; long long out2_bstride; int out2_pitch;   // second output (split_m) for the direct-store epilogue
  unsigned long long* prof;   // debug (LADIFF_TC_PROF): per-CTA wait-cycle counters [grid][8]
  int dbg;                    // debug (LADIFF_TC_DBG, only with LADIFF_TC_PROF): ablation bits, see tc_conv_launch
};

// X view description used by the SIMT check kernel (same math, no tensor maps, no tensor cores)
struct TcRefView {
  const h16* x; long long bstride; int pitch; int Lv; int Cv;
  const h16* w; int Ktot;
  h16* out; long long out_bstride; int out_pitch;   // h16 output view incl. the UP phase interleave (host fills)
  h16* out2; long long out2_bstride; int out2_pitch;
};

int tc_conv_plan(const TcConvDesc& d, TcConvParams* p, TcRefView* rv);
int tc_conv_launch(const TcConvParams& p, cudaStream_t st);
int tc_conv_ref_launch(const TcConvParams& p, const TcRefView& v, cudaStream_t st);
int tc_make_tmap_w(CUtensorMap* tm, const h16* w, int Cout, int Ktot);
int tc_num_sms();
