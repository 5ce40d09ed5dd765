// Microbenchmark: tcgen05.mma (kind::f16, bf16 operands from shared memory, K-major SWIZZLE_128B) issue throughput per SM
// as a function of N, for cta_group::1 (M=128) and cta_group::2 (M=256 over a CTA pair), with the conv kernel's access
// pattern (3 weight tiles + one row-shifted activation tile per stage).  No TMA traffic: operands are whatever is in smem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
template <int CG> __device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG> __device__ __forceinline__ void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// mode bit0: taps read the activation tile at row offsets 0,1,2 (else all at 0); bit1: one weight tile reused by all taps
template <int CG>
__global__ void __launch_bounds__(256, 1) mma_rate_kernel(int N, int stages, int iters, int mode, long long* out, const uint8_t* gsrc) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_stage, bar_done;
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8];
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    mbar_init(&bar_stage, 1); mbar_init(&bar_done, 1);
    for (int i = 0; i < 8; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int nloc = CG == 2 ? N / 2 : N;                 // activation rows held by this CTA
  const uint32_t b_bytes = (uint32_t)((nloc + 8) * 128 + 1023) & ~1023u;
  const uint32_t stage_bytes = 3 * 16384 + b_bytes;
  const int ntap = ((mode >> 4) & 3) ? ((mode >> 4) & 3) : 3;
  if (threadIdx.x == 32 && (mode & 4)) {        // stand-in producer: waits for a free stage, hands it straight back as full
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty_bar[st], ph ^ 1);
      if (mode & 256) {
        const uint32_t nb = stage_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar[st])), "r"(nb) : "memory");
        const uint8_t* src = gsrc + ((size_t)blockIdx.x * 8 + (it & 7)) * 131072;
        for (uint32_t off = 0; off < nb; off += 16384) {
          const uint32_t sz = nb - off < 16384 ? nb - off : 16384;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(base + (uint32_t)st * stage_bytes + off), "l"(src + off), "r"(sz), "r"(smem_u32(&full_bar[st])) : "memory");
        }
      } else
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[st])) : "memory");
      if (++st == stages) { st = 0; ph ^= 1; }
    }
  }
  if ((mode & 512) && threadIdx.x >= 128) {   // epilogue stand-in: read the accumulator the MMAs are not writing
    const int w = (threadIdx.x >> 5) & 3;
    float sink = 0.f;
    for (int rep = 0; rep < iters / 2; ++rep) {
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + 256u * ((rep & 1) ^ 1) + (uint32_t)c0 + ((uint32_t)(w * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) sink += __uint_as_float(r[i]);
      }
    }
    if (sink == 123.456f) out[0] = 1;
  }
  if (threadIdx.x < 32 && rank == 0 && elect_one()) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    const long long t0 = clock64();
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (mode & 4) { mbar_wait(&full_bar[st], ph); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      const uint32_t a_base = base + (uint32_t)st * stage_bytes;
      const uint64_t adesc = umma_desc(a_base), bdesc = umma_desc(a_base + 3 * 16384);
      const uint32_t d = tmem + (uint32_t)((it & 1) * 256);
#pragma unroll
      for (int tap = 0; tap < 3; ++tap) {
        if (tap >= ntap) break;
        const uint64_t a = adesc + ((mode & 2) ? 0 : tap * (16384 >> 4)) + ((mode & 8) ? tap * 8 : 0);   // bit 3: row-shifted A start
        const uint64_t b = bdesc + ((mode & 1) ? tap * 8 : 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<CG>(d, a + 2 * k, b + 2 * k, idesc, (tap | k) ? 1u : 0u);
      }
      if (mode & 4) commit<CG>(&empty_bar[st]); else commit<CG>(&bar_stage);
      if (++st == stages) { st = 0; ph ^= 1; }
    }
    commit<CG>(&bar_done);
    mbar_wait(&bar_done, 0);
    const long long t1 = clock64();
    out[blockIdx.x / CG] = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (threadIdx.x < 32) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

static uint8_t* g_src = nullptr;
template <int CG> static void run(int N, int mode, int grid_ctas, long long* dout) {
  const int iters = 400;
  const int nloc = CG == 2 ? N / 2 : N;
  const int b_bytes = ((nloc + 8) * 128 + 1023) & ~1023;
  int stages = (200 * 1024 - 2048) / (3 * 16384 + b_bytes);
  if (stages > 4) stages = 4;
  cudaFuncSetAttribute(mma_rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_ctas); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG>, N, stages, iters, mode, dout, (const uint8_t*)g_src);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return; }
  }
  long long h[148];
  cudaMemcpy(h, dout, sizeof(long long) * (grid_ctas / CG), cudaMemcpyDeviceToHost);
  double mx = 0, av = 0;
  for (int i = 0; i < grid_ctas / CG; ++i) { av += (double)h[i]; if ((double)h[i] > mx) mx = (double)h[i]; }
  av /= grid_ctas / CG;
  const int ntap = ((mode >> 4) & 3) ? ((mode >> 4) & 3) : 3;
  const double per = av / (iters * 4.0 * ntap);
  const double ideal = (double)N / 2.0;                 // cycles per MMA per SM at 4096 MAC/clk/SM (M=128 rows per SM)
  printf("cta_group::%d N=%3d mode=%d stages=%d grid=%d: %.1f clk/MMA (max-CTA %.1f)  ideal %.0f  -> %.1f%% of tensor peak\n", CG, N, mode,
         stages, grid_ctas, per, mx / (iters * 4.0 * ntap), ideal, 100.0 * ideal / per);
}

int main(int argc, char** argv) {
  long long* dout;
  cudaMalloc(&dout, sizeof(long long) * 148);
  cudaMalloc(&g_src, (size_t)148 * 8 * 131072);
  cudaMemset(g_src, 0x3c, (size_t)148 * 8 * 131072);
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs: %d\n", nsm);
  const int ns[] = {144, 160, 208, 256};
  printf("-- ring handshake, 12 MMAs per stage, stand-in producer (no copies)\n");
  for (int n : ns) run<1>(n, 1 | 4, nsm, dout);
  printf("-- + the producer really copies the stage bytes (cp.async.bulk global->shared, L2-resident source)\n");
  for (int n : ns) run<1>(n, 1 | 4 | 256, nsm, dout);
  printf("-- + four warps read the idle accumulator with tcgen05.ld (no copies)\n");
  for (int n : ns) run<1>(n, 1 | 4 | 512, nsm, dout);
  printf("-- + both\n");
  for (int n : ns) run<1>(n, 1 | 4 | 256 | 512, nsm, dout);
  printf("-- free-running issue: B start shifted by tap rows (bit0) / A start shifted (bit3) / neither\n");
  for (int n : ns) run<1>(n, 0, nsm, dout);
  for (int n : ns) run<1>(n, 1, nsm, dout);
  for (int n : ns) run<1>(n, 8, nsm, dout);
  for (int n : ns) run<2>(n == 144 ? 128 : n, 8, nsm, dout);
  for (int n : ns) run<2>(n == 144 ? 128 : n, 0, nsm, dout);
  printf("-- both, on ONE SM only\n");
  for (int n : ns) run<1>(n, 1 | 4 | 256 | 512, 1, dout);
  return 0;
}
