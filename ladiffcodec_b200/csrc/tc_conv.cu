// tcgen05 implicit-GEMM Conv1d for the UNet (sm_100a).
//
// Replaces every F.conv1d of Unet1D.forward (reference srcs/modules/unet.py:80,61,65,201,204,232,307,369)
// on channels-last bf16 activations with fp32 accumulation in TMEM.
//
//   D[m, n] (TMEM, fp32; lane = output channel m, column = position n)
//     = sum over K-segments (conv taps / concat halves) and 64-channel chunks of
//       A = W[m0:m0+128, k:k+64]      (bf16, K-major, TMA 2D, SWIZZLE_128B)
//       B = X[b, l0+shift : +NT, c:c+64] (bf16, K-major, TMA 3D, SWIZZLE_128B, OOB rows -> 0 = conv zero padding)
//
// One CTA per (position tile, 128-channel tile, clip).  Warp 0 lane 0 = TMA producer, warp 1 lane 0 =
// MMA issuer (tcgen05.mma cta_group::1 kind::f16, M=128, N=NT, K=16), all four warps = epilogue
// (tcgen05.ld 32x32b -> +bias -> GroupNorm partial sums -> channels-last store).
#include <cudaTypedefs.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Parity wait with a watchdog: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins == 64) t0 = clock64();
    if (spins > 64 && (spins & 1023) == 0 && clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 | LBO(=1, ignored for swizzled K-major)<<16 | SBO(=1024 B: 8 rows x 128 B)>>4 <<32 | version 1 <<46 | layout 2 <<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;  // 16 KB

__global__ void __launch_bounds__(128) tc_conv_kernel(const __grid_constant__ TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt_idx = blockIdx.x, m0 = blockIdx.y * TC_BM, b = blockIdx.z;
  const int l0 = nt_idx * p.NT;
  const uint32_t b_bytes = (uint32_t)p.NT * 128u;
  const uint32_t stage_bytes = A_BYTES + b_bytes;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.NT) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmX) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer
    int it = 0, kofs = 0;
    for (int s = 0; s < p.nseg; ++s) {
      const TcSeg sg = p.seg[s];
      if (m0 >= sg.m_lo && m0 < sg.m_hi) {
        for (int c = 0; c < sg.nchunk; ++c, ++it) {
          const int st = it % p.stages;
          mbar_wait(&empty_bar[st], ((it / p.stages) & 1) ^ 1);
          mbar_expect_tx(&full_bar[st], stage_bytes);
          const uint32_t a_dst = smem_base + st * stage_bytes;
          tma_load_2d(a_dst, &p.tmW, &full_bar[st], kofs + c * TC_BK, m0);
          tma_load_3d(a_dst + A_BYTES, &p.tmX, &full_bar[st], sg.ch0 + c * TC_BK, l0 + sg.shift, b);
        }
      }
      kofs += sg.nchunk * TC_BK;
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer.  Instruction descriptor (InstrDescriptor, mma_sm100_desc.hpp):
    // c_format F32 (1<<4) | a_format BF16 (1<<7) | b_format BF16 (1<<10) | K-major A,B | N>>3 <<17 | M>>4 <<24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    int it = 0;
    for (int s = 0; s < p.nseg; ++s) {
      const TcSeg sg = p.seg[s];
      if (m0 >= sg.m_lo && m0 < sg.m_hi) {
        for (int c = 0; c < sg.nchunk; ++c, ++it) {
          const int st = it % p.stages;
          mbar_wait(&full_bar[st], (it / p.stages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_base + st * stage_bytes;
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16(tmem_base, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc, (it > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty_bar[st]);   // frees the smem slot once these MMAs have read it
        }
      }
    }
    tc_commit(&accum_bar);             // accumulator complete
  }
  __syncwarp();

  // ---------------- epilogue (all 4 warps; warp w owns TMEM lanes 32w..32w+31)
  mbar_wait(&accum_bar, 0);
  tc_fence_after();
  {
    const int ch = m0 + warp * 32 + lane;
    const float bias = p.bias ? p.bias[ch] : 0.f;
    const int oc = p.out_ch0 + ch + ((p.out_split && ch >= p.out_split) ? p.out_jump : 0);
    const long long obase = (long long)b * p.out_bstride + oc;
    float s1 = 0.f, s2 = 0.f;
    const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < p.NT; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tlane + (uint32_t)c0, r);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int l = l0 + c0 + j;
        if (l < p.Lout) {
          float v = __uint_as_float(r[j]) + bias;
          s1 += v; s2 += v * v;
          if (p.res) v += __bfloat162float(p.res[(long long)b * p.res_bstride + (long long)l * p.res_pitch + p.res_ch0 + ch]);
          const long long o = obase + (long long)l * p.out_pitch;
          if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = v;
          else reinterpret_cast<bf16*>(p.out)[o] = __float2bfloat16(v);
        }
      }
    }
    if (p.stats) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0)
        p.stats[((long long)b * gridDim.x + nt_idx) * (p.Cout / 32) + (m0 / 32 + warp)] = make_float2(s1, s2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// SIMT check kernel: identical operands, tiling and epilogue semantics, plain FMA loop.
// grid (n_ntiles, Cout/32, B), 128 threads: lane = channel, warp w handles rows w, w+4, ...
__global__ void __launch_bounds__(128) tc_conv_ref_kernel(TcConvParams p, TcRefView v) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt_idx = blockIdx.x, ch = blockIdx.y * 32 + lane, b = blockIdx.z;
  const int m0 = (ch / TC_BM) * TC_BM;
  const int l0 = nt_idx * p.NT;
  const float bias = p.bias ? p.bias[ch] : 0.f;
  const int oc = p.out_ch0 + ch + ((p.out_split && ch >= p.out_split) ? p.out_jump : 0);
  const long long obase = (long long)b * p.out_bstride + oc;
  float s1 = 0.f, s2 = 0.f;
  for (int r = warp; r < p.NT; r += 4) {
    const int l = l0 + r;
    if (l >= p.Lout) break;
    float acc = 0.f;
    int kofs = 0;
    for (int s = 0; s < p.nseg; ++s) {
      const TcSeg sg = p.seg[s];
      const int row = l + sg.shift;
      if (m0 >= sg.m_lo && m0 < sg.m_hi && row >= 0 && row < v.Lv) {
        const bf16* xr = v.x + (long long)b * v.bstride + (long long)row * v.pitch + sg.ch0;
        const bf16* wr = v.w + (long long)ch * v.Ktot + kofs;
        for (int c = 0; c < sg.nchunk * TC_BK; ++c) acc += __bfloat162float(wr[c]) * __bfloat162float(xr[c]);
      }
      kofs += sg.nchunk * TC_BK;
    }
    float val = acc + bias;
    s1 += val; s2 += val * val;
    if (p.res) val += __bfloat162float(p.res[(long long)b * p.res_bstride + (long long)l * p.res_pitch + p.res_ch0 + ch]);
    const long long o = obase + (long long)l * p.out_pitch;
    if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = val;
    else reinterpret_cast<bf16*>(p.out)[o] = __float2bfloat16(val);
  }
  if (p.stats) {
    __shared__ float red[2][4][32];
    red[0][warp][lane] = s1; red[1][warp][lane] = s2;
    __syncthreads();
    if (warp == 0) {
      float a = red[0][0][lane] + red[0][1][lane] + red[0][2][lane] + red[0][3][lane];
      float c = red[1][0][lane] + red[1][1][lane] + red[1][2][lane] + red[1][3][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
      }
      if (lane == 0) p.stats[((long long)b * gridDim.x + nt_idx) * (p.Cout / 32) + blockIdx.y] = make_float2(a, c);
    }
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

}  // namespace

size_t tc_smem_bytes(int NT, int stages) { return (size_t)stages * (A_BYTES + (size_t)NT * 128) + 1024; }

int tc_pick_stages(int NT) {
  // as deep as fits in ~200 KB, at most 6
  int s = (int)((200 * 1024) / (A_BYTES + NT * 128));
  return s > 6 ? 6 : (s < 2 ? 2 : s);
}

// Position-tile size for clips of length L: multiple of 16, <= 256, minimising padded work.
int tc_pick_nt(int L, int* n_tiles) {
  int best_nt = 0, best_n = 0;
  long best_cost = -1;
  const int nmin = cdiv(L, 256);
  for (int n = nmin; n <= nmin + 3; ++n) {
    int nt = cdiv(cdiv(L, n), 16) * 16;
    if (nt > 256) continue;
    if (nt < 16) nt = 16;
    const long cost = (long)n * nt;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_nt = nt; best_n = cdiv(L, nt); }
  }
  if (n_tiles) *n_tiles = best_n;
  return best_nt;
}

int tc_make_tmap_w(CUtensorMap* tm, const bf16* w, int Cout, int Ktot) {
  auto enc = get_encode();
  LADIFF_REQUIRE(enc != nullptr, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {TC_BK, TC_BM};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LADIFF_REQUIRE(r == CUDA_SUCCESS, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled(W %dx%d) failed: %d", Cout, Ktot, (int)r);
  return 0;
}

int tc_make_tmap_x(CUtensorMap* tm, const bf16* x, int B, int Lv, int Cv, int pitch, long long bstride, int NT) {
  auto enc = get_encode();
  LADIFF_REQUIRE(enc != nullptr, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)Cv, (cuuint64_t)Lv, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)bstride * 2};
  cuuint32_t box[3] = {TC_BK, (cuuint32_t)NT, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LADIFF_REQUIRE(r == CUDA_SUCCESS, LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled(X B%d L%d C%d pitch %d NT %d) failed: %d", B, Lv, Cv,
                 pitch, NT, (int)r);
  return 0;
}

int tc_conv_launch(const TcConvParams& p, int B, cudaStream_t st) {
  LADIFF_REQUIRE(p.Cout % TC_BM == 0 && p.NT % 16 == 0 && p.NT >= 16 && p.NT <= 256 && p.nseg >= 1 && p.nseg <= TC_MAX_SEG,
                 LADIFF_ERR_ARG, "tc_conv: bad tile config Cout=%d NT=%d nseg=%d", p.Cout, p.NT, p.nseg);
  LADIFF_REQUIRE(p.stages >= 2 && p.stages <= 8, LADIFF_ERR_ARG, "tc_conv: stages=%d", p.stages);
  const size_t smem = tc_smem_bytes(p.NT, p.stages);
  LADIFF_REQUIRE(smem <= 226 * 1024, LADIFF_ERR_ARG, "tc_conv: smem %zu too large", smem);
  static size_t attr_set = 0;     // opt-in dynamic shared memory (static barriers take a few hundred bytes of the 227 KB)
  if (smem > attr_set) {
    LADIFF_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = smem;
  }
  dim3 grid(cdiv(p.Lout, p.NT), p.Cout / TC_BM, B);
  tc_conv_kernel<<<grid, 128, smem, st>>>(p);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}

int tc_conv_ref_launch(const TcConvParams& p, const TcRefView& v, int B, cudaStream_t st) {
  dim3 grid(cdiv(p.Lout, p.NT), p.Cout / 32, B);
  tc_conv_ref_kernel<<<grid, 128, 0, st>>>(p, v);
  LADIFF_CUDA_OK(cudaGetLastError());
  return 0;
}
