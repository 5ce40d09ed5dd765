// Load-time weight folds for the UNet: weight standardisation (unet.py:72-80) + h16 K-major repack.
#include "fold.cuh"

namespace {

__device__ float block_sum128(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// w [Cout][Cin][K] fp32 -> packed [Cout][K*Cin] h16 with k-index = tap*Cin + c.
// standardize: per output channel (w - mean) * rsqrt(var_biased + 1e-5) over (Cin, K).     one block (128 thr) per o
__global__ void __launch_bounds__(128) pack_conv_kernel(const float* __restrict__ w, h16* __restrict__ out, int Cin, int K,
                                                        int standardize, int out_ld) {
  __shared__ float red[4];
  const int o = blockIdx.x, n = Cin * K;
  const float* wr = w + (long long)o * n;
  float mean = 0.f, rstd = 1.f;
  if (standardize) {
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 128) s += wr[i];
    mean = block_sum128(s, red) / (float)n;
    float q = 0.f;
    for (int i = threadIdx.x; i < n; i += 128) { const float d = wr[i] - mean; q += d * d; }
    const float var = block_sum128(q, red) / (float)n;
    rstd = rsqrtf(var + 1e-5f);
  }
  for (int i = threadIdx.x; i < n; i += 128) {
    const int c = i / K, tap = i - c * K;
    out[(long long)o * out_ld + tap * Cin + c] = f2h((wr[i] - mean) * rstd);
  }
}

// nearest-x2 upsample folded into the k=3 conv that follows it (unet.py:58-62):
//   y[2m]   = W0 x[m-1] + (W1+W2) x[m]            rows [0, Cout)
//   y[2m+1] = (W0+W1) x[m] + W2 x[m+1]            rows [Cout, 2Cout)
// w [Cout][Cin][3] -> packed [2Cout][3*Cin] (tap order -1, 0, +1)
__global__ void pack_up_kernel(const float* __restrict__ w, h16* __restrict__ out, int Cout, int Cin) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)2 * Cout * 3 * Cin;
  if (i >= total) return;
  const int c = (int)(i % Cin);
  const int tap = (int)((i / Cin) % 3);
  const int m = (int)(i / (3LL * Cin));
  const int o = m < Cout ? m : m - Cout;
  const float* wr = w + ((long long)o * Cin + c) * 3;
  float v;
  if (m < Cout) v = tap == 0 ? wr[0] : (tap == 1 ? wr[1] + wr[2] : 0.f);
  else v = tap == 0 ? 0.f : (tap == 1 ? wr[0] + wr[1] : wr[2]);
  out[i] = f2h(v);
}

__global__ void dup_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n) out[i] = b[i % n];
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, h16* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = f2h(x[i]);
}

}  // namespace

int pack_conv_launch(const float* w, h16* out, int Cout, int Cin, int K, int standardize, cudaStream_t st, int out_ld) {
  pack_conv_kernel<<<Cout, 128, 0, st>>>(w, out, Cin, K, standardize, out_ld > 0 ? out_ld : Cin * K);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int pack_up_launch(const float* w, h16* out, int Cout, int Cin, cudaStream_t st) {
  const long long total = (long long)2 * Cout * 3 * Cin;
  pack_up_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, out, Cout, Cin);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int dup_bias_launch(const float* b, float* out, int n, cudaStream_t st) {
  dup_bias_kernel<<<cdiv(2 * n, 256), 256, 0, st>>>(b, out, n);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
int f32_to_bf16_launch(const float* x, h16* y, long long n, cudaStream_t st) {
  f32_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, n);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
