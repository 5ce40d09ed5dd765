// Microbenchmark: warp-level mma.sync issue throughput on sm_100a (the legacy tensor path: HMMA), register operands only.
//   TF32 m16n8k8 (HMMA.1688.F32.TF32) and the 16-bit m16n8k16 (HMMA.16816.F32), 8 independent accumulators per warp,
//   4 / 8 / 16 warps per SM on every SM.  Evidence for DESIGN.md §4: a 3xTF32 conv on mma.sync cannot beat the fp32 FMA kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu && ./mma_sync_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int KIND>   // 0: tf32 m16n8k8, 1: f16 m16n8k16
__global__ void rate_kernel(float* out, int iters) {
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  uint32_t a0 = threadIdx.x * 0x01010101u, a1 = a0 ^ 0x3c003c00u, a2 = a0 + 7u, a3 = a1 + 11u, b0 = a0 ^ 0x12345678u, b1 = b0 + 3u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 123.456f) out[0] = s;
}

template <int KIND>
static void run(const char* name, double flop_per_mma, int warps_per_sm, int sms, float* d) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<KIND><<<sms, 32 * warps_per_sm>>>(d, 100);
  cudaEventRecord(e0);
  rate_kernel<KIND><<<sms, 32 * warps_per_sm>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)sms * warps_per_sm * iters * 8;
  printf("%-22s warps/SM %2d: %8.1f TFLOP/s   %.1f clk per MMA per SM sub-partition at 1.9 GHz\n", name, warps_per_sm,
         mmas * flop_per_mma / (ms * 1e-3) / 1e12, (ms * 1e-3) * 1.9e9 / ((double)iters * 8 * warps_per_sm / 4.0));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d;
  cudaMalloc(&d, 4);
  for (int w : {4, 8, 16}) run<0>("tf32 m16n8k8", 2.0 * 16 * 8 * 8, w, sms, d);
  for (int w : {4, 8, 16}) run<1>("f16 m16n8k16", 2.0 * 16 * 8 * 16, w, sms, d);
  return 0;
}
