// tcgen05 / TMEM version of the stand-alone fully-fused MLP forward (tiny-cuda-nn `Network(FullyFusedMLP)`,
// models/network_utils.py:181): 5th-generation tensor cores instead of the warp-level mma.sync path of mlp.cu.
//
// One CTA = 128 threads = one 128-row tile (UMMA M = 128, cta_group::1).  Thread r owns row r end to end:
//   it writes its input row into shared memory in the canonical K-major no-swizzle operand layout
//   (8x16-byte core matrices; LBO = 128 B between core matrices along K, SBO = (K/8)*128 B between 8-row groups),
//   ONE elected thread issues tcgen05.mma (A, B from smem descriptors, D in TMEM, fp32 accumulate) and commits to an
//   mbarrier, every thread then pulls ITS row of the accumulator out of TMEM with tcgen05.ld.32x32b (lane = row),
//   applies the activation, re-packs to fp16 and writes the next layer's A operand row -- no cross-thread shuffles at all.
// Weights are staged once per CTA in the same canonical layout; TMEM: 128 columns (64 hidden + 16 output, power of two).
#include "common.cuh"
#include "mlp_warp.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kRows = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major no-swizzle layout: element (row, k) of a [rows][K] fp16 operand
__device__ __forceinline__ int canon_off_bytes(int row, int k, int K) { return ((row >> 3) * (K >> 3) + (k >> 3)) * 128 + (row & 7) * 16 + (k & 7) * 2; }

// 64-bit shared-memory matrix descriptor (SM100 UMMA): start>>4 | LBO>>4 @16 | SBO>>4 @32 | version 1 @46 | layout NONE @61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// 32-bit instruction descriptor, kind::f16: D = f32 (1 @4), A = B = f16 (0 @7, 0 @10), K-major A and B (0 @15, 0 @16), N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t mbar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t mbar_saddr, uint32_t parity) {
  const long long t_start = clock64();
  for (;;) {
    if (clock64() - t_start > 2000000000ll) break;  // ~1 s: something is wrong with the MMA / commit
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(mbar_saddr), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;  // never signalled: the caller traps instead of hanging the GPU
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// stage a row-major [rows][K] fp16 matrix from global memory into the canonical layout
__device__ __forceinline__ void stage_canonical(uint8_t* dst, const __half* __restrict__ src, int rows, int K) {
  const int vec_per_row = K / 8;
  for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
    const int r = i / vec_per_row, kc = i % vec_per_row;
    *reinterpret_cast<uint4*>(dst + canon_off_bytes(r, kc * 8, K)) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * K) + kc);
  }
}

// one layer on the tensor core: D[128 x N] (TMEM column d_col) = A[128 x K] (smem) * W[N x K]^T (smem)
__device__ __forceinline__ void issue_layer(uint32_t tmem_base, int d_col, uint32_t a_saddr, uint32_t w_saddr, int K, int N, bool swap,
                                            uint32_t mbar_saddr) {
  const uint32_t lbo = 128, sbo = (uint32_t)(K / 8) * 128;
  const uint32_t idesc = make_idesc(128, N);
  for (int kk = 0; kk < K / 16; ++kk) {
    const uint32_t adv = (uint32_t)kk * 256;  // two core matrices along K per MMA
    const uint64_t da = swap ? make_desc(a_saddr + adv, sbo, lbo) : make_desc(a_saddr + adv, lbo, sbo);
    const uint64_t db = swap ? make_desc(w_saddr + adv, sbo, lbo) : make_desc(w_saddr + adv, lbo, sbo);
    umma_f16(tmem_base + (uint32_t)d_col, da, db, idesc, kk > 0 ? 1u : 0u);
  }
  umma_commit(mbar_saddr);
}

template <int KT_IN>
__global__ void __launch_bounds__(kThreads) mlp_fwd_tc_kernel(nsr_mlp_t m, const __half* __restrict__ x, const __half* __restrict__ params,
                                                              __half* __restrict__ out, int64_t n, int swap, int* __restrict__ status) {
  constexpr int IN_PAD = KT_IN * 16;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_mbar;
  __shared__ uint32_t s_tmem;
  uint8_t* sX = smem;                                   // [128][IN_PAD]
  uint8_t* sH = sX + kRows * IN_PAD * 2;                // [128][64]
  uint8_t* sW1 = sH + kRows * 64 * 2;                   // [64][IN_PAD]
  uint8_t* sWh = sW1 + 64 * IN_PAD * 2;                 // (n_hidden-1) x [64][64]
  uint8_t* sWl = sWh + (m.n_hidden - 1) * 64 * 64 * 2;  // [16][64]
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  stage_canonical(sW1, params, 64, IN_PAD);
  {
    size_t off = (size_t)64 * IN_PAD;
    for (int h = 0; h < m.n_hidden - 1; ++h) {
      stage_canonical(sWh + h * 64 * 64 * 2, params + off, 64, 64);
      off += 64 * 64;
    }
    stage_canonical(sWl, params + off, 16, 64);
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t mbar = smem_u32(&s_mbar);
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
  uint32_t parity = 0;

  const int64_t n_tiles = (n + kRows - 1) / kRows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * kRows + tid;
    // ---- my input row -> canonical A operand
#pragma unroll
    for (int kc = 0; kc < IN_PAD / 8; ++kc) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row < n) v = __ldg(reinterpret_cast<const uint4*>(x + row * IN_PAD) + kc);
      *reinterpret_cast<uint4*>(sX + canon_off_bytes(tid, kc * 8, IN_PAD)) = v;
    }
    proxy_fence();
    __syncthreads();
    // ---- hidden layers: D (TMEM cols 0..63) = A * W^T, ReLU, re-pack as the next A operand
    for (int layer = 0; layer < m.n_hidden; ++layer) {
      if (tid == 0) {
        tc_fence_after();
        if (layer == 0)
          issue_layer(tmem, 0, smem_u32(sX), smem_u32(sW1), IN_PAD, 64, swap != 0, mbar);
        else
          issue_layer(tmem, 0, smem_u32(sH), smem_u32(sWh + (layer - 1) * 64 * 64 * 2), 64, 64, swap != 0, mbar);
      }
      if (!mbar_wait(mbar, parity)) {
        if (status) atomicExch(status, 1);
        __trap();
      }
      parity ^= 1u;
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pk[j] = nsr_pack_h2(nsr_apply_act(__uint_as_float(v[2 * j]), m.activation), nsr_apply_act(__uint_as_float(v[2 * j + 1]), m.activation));
        *reinterpret_cast<uint4*>(sH + canon_off_bytes(tid, c0, 64)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(sH + canon_off_bytes(tid, c0 + 8, 64)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      tc_fence_before();
      proxy_fence();
      __syncthreads();
    }
    // ---- output layer: D (TMEM cols 64..79) = H * Wl^T
    if (tid == 0) {
      tc_fence_after();
      issue_layer(tmem, 64, smem_u32(sH), smem_u32(sWl), 64, 16, swap != 0, mbar);
    }
    if (!mbar_wait(mbar, parity)) {
      if (status) atomicExch(status, 1);
      __trap();
    }
    parity ^= 1u;
    tc_fence_after();
    {
      uint32_t v[16];
      tmem_ld16(lane_addr + 64u, v);
      tmem_ld_wait();
      if (row < n) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pk[j] = nsr_pack_h2(nsr_apply_act(__uint_as_float(v[2 * j]), m.out_activation), nsr_apply_act(__uint_as_float(v[2 * j + 1]), m.out_activation));
        uint4* o = reinterpret_cast<uint4*>(out + row * 16);
        o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM columns and the smem operands are free for the next tile
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

}  // namespace

// tcgen05 variant of nsr_mlp_fwd (same arguments).  variant bit 0: swap the LBO / SBO fields of the smem descriptors
// (bring-up switch).  status (device int, may be NULL) is set to 1 if an mbarrier wait timed out (the kernel then traps).
extern "C" int nsr_mlp_fwd_tc(const nsr_mlp_t* m, const void* x_h, const void* params_h, void* out_h, int64_t n, int variant, int* status,
                              void* stream) {
  NSR_REQUIRE(m != nullptr && m->n_hidden >= 1 && m->n_hidden <= 3 && m->n_out >= 1 && m->n_out <= 16 && m->n_in >= 1 && m->n_in <= 64,
              "nsr_mlp_fwd_tc: unsupported network shape");
  if (n == 0) return 0;
  const int in_pad = (m->n_in + 15) / 16 * 16, kt = in_pad / 16;
  const size_t smem = (size_t)kRows * in_pad * 2 + kRows * 64 * 2 + 64 * in_pad * 2 + (size_t)(m->n_hidden - 1) * 64 * 64 * 2 + 16 * 64 * 2 + 128;
  const int64_t tiles = (n + kRows - 1) / kRows;
  // ~40 KB smem and 128 TMEM columns per CTA: 4 CTAs per SM overlap each other's load / MMA / epilogue phases
  const int grid = (int)min((int64_t)nsr_sm_count() * 4, tiles);
#define NSR_LAUNCH_TC(KT)                                                                                                  \
  case KT: {                                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_tc_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e != cudaSuccess) {                                                                                                \
      nsr_set_error("nsr_mlp_fwd_tc: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));                \
      return 2;                                                                                                            \
    }                                                                                                                      \
    mlp_fwd_tc_kernel<KT><<<grid, kThreads, smem, (cudaStream_t)stream>>>(*m, (const __half*)x_h, (const __half*)params_h, \
                                                                          (__half*)out_h, n, variant & 1, status);       \
  } break;
  switch (kt) {
    NSR_LAUNCH_TC(1) NSR_LAUNCH_TC(2) NSR_LAUNCH_TC(3) NSR_LAUNCH_TC(4)
    default: NSR_REQUIRE(false, "nsr_mlp_fwd_tc: unsupported padded input width %d", in_pad);
  }
#undef NSR_LAUNCH_TC
  NSR_CHECK_LAUNCH("nsr_mlp_fwd_tc");
  return 0;
}
