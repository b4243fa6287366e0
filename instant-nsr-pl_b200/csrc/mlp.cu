// Stand-alone fully-fused 64-wide MLP: the tiny-cuda-nn `Network(FullyFusedMLP)` surface used by
// models/network_utils.py:181 (texture / geometry networks).  Whole network in one kernel,
// activations never leave the SM; see mlp_warp.cuh for the warp-level pieces.
// BIAS = true: the same kernels as the reference's VanillaMLP with ReLU (models/network_utils.py:95-139: nn.Linear layers WITH
// biases, fp32 in/out under autocast(False)) for the 64-wide shapes of its configs (neus-dtu.yaml:84-91 background density 32->64->8,
// colour networks): fp16 tensor-core operands, fp32 accumulation starting from the fp32 bias, compact fp32 output [n, n_out],
// fp32 dy [n, n_out] in and bias gradients (column sums of the pre-activation gradient tiles) out.
#include "mlp_warp.cuh"

namespace {

struct MlpSmem {
  int ld1;        // in_pad + 8
  int off_w1;     // halves
  int off_wh;     // (n_hidden-1) x [64][72]
  int off_wl;     // [16][72]
  int w_total;    // halves used by the weights
};

__host__ __device__ inline MlpSmem mlp_smem_layout(int in_pad, int n_hidden) {
  MlpSmem s;
  s.ld1 = in_pad + NSR_LDW_PAD;
  s.off_w1 = 0;
  s.off_wh = NSR_W * s.ld1;
  s.off_wl = s.off_wh + (n_hidden - 1) * NSR_W * NSR_LD64;
  s.w_total = s.off_wl + 16 * NSR_LD64;
  return s;
}

__device__ __forceinline__ void stage_weights(__half* smem, const MlpSmem& L, const __half* __restrict__ params, int in_pad, int n_hidden) {
  const int tid = threadIdx.x, nt = blockDim.x;
  nsr_stage_matrix(smem + L.off_w1, params, NSR_W, in_pad, tid, nt);
  size_t off = (size_t)NSR_W * in_pad;
  for (int h = 0; h < n_hidden - 1; ++h) {
    nsr_stage_matrix(smem + L.off_wh + h * NSR_W * NSR_LD64, params + off, NSR_W, NSR_W, tid, nt);
    off += NSR_W * NSR_W;
  }
  nsr_stage_matrix(smem + L.off_wl, params + off, 16, NSR_W, tid, nt);
}

// accumulators start from zero (FullyFused) or from the fp32 bias of their column (BIAS)
template <bool BIAS, int MT, int NT>
__device__ __forceinline__ void init_acc(float (&acc)[MT][NT][4], const float* __restrict__ bias) {
  if (BIAS) {
    const int c2 = (threadIdx.x & 3) * 2;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float b0 = __ldg(bias + n * 8 + c2), b1 = __ldg(bias + n * 8 + c2 + 1);
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[m][n][0] = b0, acc[m][n][1] = b1, acc[m][n][2] = b0, acc[m][n][3] = b1;
    }
  } else {
    nsr_zero_acc(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// forward: 4 warps x 32 rows
// ------------------------------------------------------------------------------------------------
constexpr int kFwdWarps = 4;

template <int KT_IN, bool BIAS>
__global__ void __launch_bounds__(kFwdWarps * 32) mlp_fwd_kernel(nsr_mlp_t m, const __half* __restrict__ x,
                                                                 const __half* __restrict__ params, const float* __restrict__ bias,
                                                                 __half* __restrict__ out, float* __restrict__ out_f32, int64_t n) {
  extern __shared__ __align__(16) __half smem[];
  constexpr int IN_PAD = KT_IN * 16;
  const MlpSmem L = mlp_smem_layout(IN_PAD, m.n_hidden);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* xt = smem + L.w_total + warp * (32 * L.ld1 + 32 * 24);
  __half* ot = xt + 32 * L.ld1;
  stage_weights(smem, L, params, IN_PAD, m.n_hidden);
  __syncthreads();
  const int64_t n_tiles = (n + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * kFwdWarps + warp; tile < n_tiles; tile += (int64_t)gridDim.x * kFwdWarps) {
    const int64_t row0 = tile * 32;
    const int rows = (int)min((int64_t)32, n - row0);
    constexpr int VPR = IN_PAD / 8;
    for (int i = lane; i < 32 * VPR; i += 32) {
      const int r = i / VPR, v = i % VPR;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (r < rows) val = __ldg(reinterpret_cast<const uint4*>(x + (row0 + r) * IN_PAD) + v);
      *reinterpret_cast<uint4*>(xt + r * L.ld1 + v * 8) = val;
    }
    __syncwarp();
    uint32_t a_in[2][KT_IN][4];
    nsr_load_afrag<2, KT_IN>(a_in, xt, L.ld1, 0);
    float acc[2][8][4];
    init_acc<BIAS>(acc, bias);
    nsr_gemm_w<2, KT_IN, 8>(acc, a_in, smem + L.off_w1, L.ld1);
    uint32_t a_h[2][4][4];
    nsr_acc_to_afrag<2, 8>(acc, a_h, m.activation);
    for (int h = 0; h < m.n_hidden - 1; ++h) {
      init_acc<BIAS>(acc, bias + (h + 1) * NSR_W);
      nsr_gemm_w<2, 4, 8>(acc, a_h, smem + L.off_wh + h * NSR_W * NSR_LD64, NSR_LD64);
      nsr_acc_to_afrag<2, 8>(acc, a_h, m.activation);
    }
    float acco[2][2][4];
    init_acc<BIAS>(acco, bias + m.n_hidden * NSR_W);
    nsr_gemm_w<2, 4, 2>(acco, a_h, smem + L.off_wl, NSR_LD64);
    if (BIAS) {  // VanillaMLP: compact fp32 output straight from the accumulators
      const int g = lane >> 2, c = lane & 3;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = mt * 16 + g + ((i >> 1) << 3), col = nt * 8 + c * 2 + (i & 1);
            if (r < rows && col < m.n_out) out_f32[(row0 + r) * m.n_out + col] = nsr_apply_act(acco[mt][nt][i], m.out_activation);
          }
      __syncwarp();
      continue;
    }
    uint32_t a_o[2][1][4];
    nsr_acc_to_afrag<2, 2>(acco, a_o, m.out_activation);
    nsr_store_afrag<2, 1>(a_o, ot, 24, 0);
    __syncwarp();
    for (int i = lane; i < 64; i += 32) {
      const int r = i >> 1, v = i & 1;
      if (r < rows) *reinterpret_cast<uint4*>(out + (row0 + r) * 16 + v * 8) = *reinterpret_cast<const uint4*>(ot + r * 24 + v * 8);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// backward: 8 warps x 16 rows = 128-row CTA tile; dgrad per warp, wgrad split over warps with
// register accumulators that live for the whole (persistent) kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kBwdWarps = 8;
constexpr int kBwdRows = kBwdWarps * 16;
constexpr int kWgradSlots = 7;

template <int KT_IN, bool BIAS>
__global__ void __launch_bounds__(kBwdWarps * 32) mlp_bwd_kernel(nsr_mlp_t m, const __half* __restrict__ x,
                                                                 const __half* __restrict__ params, const float* __restrict__ bias,
                                                                 const __half* __restrict__ dy, const float* __restrict__ dy_f32,
                                                                 float* __restrict__ grad_params, float* __restrict__ grad_bias,
                                                                 __half* __restrict__ dx, float* __restrict__ dx_f32, float loss_scale,
                                                                 const float* __restrict__ amax_ptr, int64_t n) {
  extern __shared__ __align__(16) __half smem[];
  constexpr int IN_PAD = KT_IN * 16;
  const MlpSmem L = mlp_smem_layout(IN_PAD, m.n_hidden);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  __half* Xt = smem + L.w_total;                       // [128][ld1]
  __half* Ht = Xt + kBwdRows * L.ld1;                  // [n_hidden][128][72]
  __half* Pt = Ht + m.n_hidden * kBwdRows * NSR_LD64;  // [n_hidden][128][72]  d/d(pre-activation)
  __half* Ot = Pt + m.n_hidden * kBwdRows * NSR_LD64;  // [128][24]
  stage_weights(smem, L, params, IN_PAD, m.n_hidden);
  if (BIAS && loss_scale <= 0.f) {  // automatic: bring the largest incoming gradient to ~2^8 (amax = max |dy|, device float)
    const float amax = fmaxf(__ldg(amax_ptr), 1e-30f);
    loss_scale = exp2f(fminf(fmaxf(floorf(log2f(256.f / amax)), -24.f), 60.f));
  }

  // wgrad bookkeeping: pair-tiles (16 out x 16 in) enumerated layer by layer
  const int p_first = 4 * KT_IN, p_hidden = 16 * (m.n_hidden - 1), p_total = p_first + p_hidden + 4;
  float wacc[kWgradSlots][2][4];
#pragma unroll
  for (int s = 0; s < kWgradSlots; ++s)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) wacc[s][j][i] = 0.f;
  float bsum = 0.f;  // BIAS: this thread's column of d(pre-activation) (threads < 64 * n_hidden) or of d(output) (the next 16)

  const int64_t n_tiles = (n + kBwdRows - 1) / kBwdRows;
  const int r0 = warp * 16;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kBwdRows;
    const int rows = (int)min((int64_t)kBwdRows, n - row0);
    __syncthreads();  // previous tile's wgrad has finished reading the smem tiles (also covers weight staging)
    {
      constexpr int VPR = IN_PAD / 8;
      for (int i = threadIdx.x; i < kBwdRows * VPR; i += blockDim.x) {
        const int r = i / VPR, v = i % VPR;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (r < rows) val = __ldg(reinterpret_cast<const uint4*>(x + (row0 + r) * IN_PAD) + v);
        *reinterpret_cast<uint4*>(Xt + r * L.ld1 + v * 8) = val;
      }
    }
    __syncthreads();
    // ---- forward recompute (this warp's 16 rows)
    uint32_t a_in[1][KT_IN][4];
    nsr_load_afrag<1, KT_IN>(a_in, Xt, L.ld1, r0);
    float acc[1][8][4];
    init_acc<BIAS>(acc, bias);
    nsr_gemm_w<1, KT_IN, 8>(acc, a_in, smem + L.off_w1, L.ld1);
    uint32_t a_h[1][4][4];
    nsr_acc_to_afrag<1, 8>(acc, a_h, m.activation);
    nsr_store_afrag<1, 4>(a_h, Ht, NSR_LD64, r0);
    for (int h = 1; h < m.n_hidden; ++h) {
      init_acc<BIAS>(acc, bias + h * NSR_W);
      nsr_gemm_w<1, 4, 8>(acc, a_h, smem + L.off_wh + (h - 1) * NSR_W * NSR_LD64, NSR_LD64);
      nsr_acc_to_afrag<1, 8>(acc, a_h, m.activation);
      nsr_store_afrag<1, 4>(a_h, Ht + h * kBwdRows * NSR_LD64, NSR_LD64, r0);
    }
    float acco[1][2][4];
    init_acc<BIAS>(acco, bias + m.n_hidden * NSR_W);
    nsr_gemm_w<1, 4, 2>(acco, a_h, smem + L.off_wl, NSR_LD64);
    // ---- output-layer gradient: d(pre) = dy * act'(y)
    uint32_t a_do[1][1][4];
    {
      float dpre[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r0 + g + ((i >> 1) << 3), col = nt * 8 + c * 2 + (i & 1);
          float d = 0.f;
          if (BIAS) {
            if (r < rows && col < m.n_out) d = dy_f32[(row0 + r) * m.n_out + col] * loss_scale;
          } else if (r < rows) {
            d = __half2float(dy[(row0 + r) * 16 + col]) * loss_scale;
          }
          const float y = nsr_apply_act(acco[0][nt][i], m.out_activation);
          dpre[nt][i] = d * nsr_act_grad_from_out(y, m.out_activation);
        }
      a_do[0][0][0] = nsr_pack_h2(dpre[0][0], dpre[0][1]);
      a_do[0][0][1] = nsr_pack_h2(dpre[0][2], dpre[0][3]);
      a_do[0][0][2] = nsr_pack_h2(dpre[1][0], dpre[1][1]);
      a_do[0][0][3] = nsr_pack_h2(dpre[1][2], dpre[1][3]);
    }
    nsr_store_afrag<1, 1>(a_do, Ot, 24, r0);
    // ---- dgrad through the hidden layers
    uint32_t a_dp[1][4][4];
    nsr_zero_acc(acc);
    nsr_gemm_wt<1, 1, 8>(acc, a_do, smem + L.off_wl, NSR_LD64);
    for (int h = m.n_hidden - 1; h >= 0; --h) {
      // a_h currently holds H[h] (post-activation) in fragment layout
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2 hv = *reinterpret_cast<const __half2*>(&a_h[0][k][j]);
          const int nt = 2 * k + (j >> 1), i0 = (j & 1) * 2;
          const float d0 = acc[0][nt][i0] * nsr_act_grad_from_out(__low2float(hv), m.activation);
          const float d1 = acc[0][nt][i0 + 1] * nsr_act_grad_from_out(__high2float(hv), m.activation);
          a_dp[0][k][j] = nsr_pack_h2(d0, d1);
        }
      nsr_store_afrag<1, 4>(a_dp, Pt + h * kBwdRows * NSR_LD64, NSR_LD64, r0);
      if (h > 0) {
        nsr_zero_acc(acc);
        nsr_gemm_wt<1, 4, 8>(acc, a_dp, smem + L.off_wh + (h - 1) * NSR_W * NSR_LD64, NSR_LD64);
        nsr_load_afrag<1, 4>(a_h, Ht + (h - 1) * kBwdRows * NSR_LD64, NSR_LD64, r0);
      }
    }
    if (BIAS ? dx_f32 != nullptr : dx != nullptr) {
      float accx[1][2 * KT_IN][4];
      nsr_zero_acc(accx);
      nsr_gemm_wt<1, 4, 2 * KT_IN>(accx, a_dp, smem + L.off_w1, L.ld1);
#pragma unroll
      for (int nt = 0; nt < 2 * KT_IN; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int r = r0 + g + hh * 8;
          if (BIAS) {  // compact fp32 [n, n_in], true (unscaled) gradient
            const float inv = 1.f / loss_scale;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = nt * 8 + c * 2 + e;
              if (r < rows && col < m.n_in) dx_f32[(row0 + r) * m.n_in + col] = accx[0][nt][hh * 2 + e] * inv;
            }
          } else if (r < rows)
            *reinterpret_cast<__half2*>(dx + (row0 + r) * IN_PAD + nt * 8 + c * 2) =
                __floats2half2_rn(accx[0][nt][hh * 2], accx[0][nt][hh * 2 + 1]);
        }
    }
    __syncthreads();
    // ---- wgrad: dW[out][in] += dPre^T * ActIn over the 128 rows of the tile
#pragma unroll
    for (int s = 0; s < kWgradSlots; ++s) {
      const int t = warp + s * kBwdWarps;
      if (t < p_total) {
        const __half *dYt, *Xin;
        int ldy, ldx, m0, n0;
        if (t < p_first) {
          dYt = Pt; ldy = NSR_LD64; Xin = Xt; ldx = L.ld1;
          m0 = (t / KT_IN) * 16; n0 = (t % KT_IN) * 16;
        } else if (t < p_first + p_hidden) {
          const int u = t - p_first, h = 1 + u / 16, v = u % 16;
          dYt = Pt + h * kBwdRows * NSR_LD64; ldy = NSR_LD64;
          Xin = Ht + (h - 1) * kBwdRows * NSR_LD64; ldx = NSR_LD64;
          m0 = (v / 4) * 16; n0 = (v % 4) * 16;
        } else {
          const int v = t - p_first - p_hidden;
          dYt = Ot; ldy = 24; Xin = Ht + (m.n_hidden - 1) * kBwdRows * NSR_LD64; ldx = NSR_LD64;
          m0 = 0; n0 = v * 16;
        }
        nsr_wgrad_tile(wacc[s][0], wacc[s][1], dYt, ldy, m0, Xin, ldx, n0, kBwdRows);
      }
    }
    if (BIAS) {  // bias gradients: column sums of the pre-activation gradient tiles (rows beyond `rows` hold zeros)
      const int t = threadIdx.x;
      if (t < NSR_W * m.n_hidden) {
        const __half* col = Pt + (size_t)(t >> 6) * kBwdRows * NSR_LD64 + (t & 63);
        float sum = 0.f;
#pragma unroll 8
        for (int r = 0; r < kBwdRows; ++r) sum += __half2float(col[(size_t)r * NSR_LD64]);
        bsum += sum;
      } else if (t < NSR_W * m.n_hidden + 16) {
        const __half* col = Ot + (t - NSR_W * m.n_hidden);
        float sum = 0.f;
#pragma unroll 8
        for (int r = 0; r < kBwdRows; ++r) sum += __half2float(col[r * 24]);
        bsum += sum;
      }
    }
  }
  // ---- flush the register accumulators (one atomic per element per CTA)
  const float inv_scale = 1.f / loss_scale;
  if (BIAS && threadIdx.x < NSR_W * m.n_hidden + 16 && bsum != 0.f) atomicAdd(grad_bias + threadIdx.x, bsum * inv_scale);
#pragma unroll
  for (int s = 0; s < kWgradSlots; ++s) {
    const int t = warp + s * kBwdWarps;
    if (t < p_total) {
      size_t base; int in_dim, m0, n0;
      if (t < p_first) {
        base = 0; in_dim = IN_PAD; m0 = (t / KT_IN) * 16; n0 = (t % KT_IN) * 16;
      } else if (t < p_first + p_hidden) {
        const int u = t - p_first, h = 1 + u / 16, v = u % 16;
        base = (size_t)NSR_W * IN_PAD + (size_t)(h - 1) * NSR_W * NSR_W; in_dim = NSR_W; m0 = (v / 4) * 16; n0 = (v % 4) * 16;
      } else {
        const int v = t - p_first - p_hidden;
        base = (size_t)NSR_W * IN_PAD + (size_t)(m.n_hidden - 1) * NSR_W * NSR_W; in_dim = NSR_W; m0 = 0; n0 = v * 16;
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int o = m0 + g + ((i >> 1) << 3), ii = n0 + j * 8 + c * 2 + (i & 1);
          atomicAdd(grad_params + base + (size_t)o * in_dim + ii, wacc[s][j][i] * inv_scale);
        }
    }
  }
}

int check_mlp(const nsr_mlp_t* m) {
  NSR_REQUIRE(m != nullptr, "mlp descriptor is NULL");
  NSR_REQUIRE(m->n_in >= 1 && m->n_in <= 64, "FullyFusedMLP: n_in must be in [1,64], got %d", m->n_in);
  NSR_REQUIRE(m->n_out >= 1 && m->n_out <= 16, "FullyFusedMLP: n_out must be in [1,16], got %d", m->n_out);
  NSR_REQUIRE(m->n_hidden >= 1 && m->n_hidden <= 3, "FullyFusedMLP: n_hidden_layers must be in [1,3], got %d", m->n_hidden);
  NSR_REQUIRE(m->activation == NSR_ACT_NONE || m->activation == NSR_ACT_RELU, "hidden activation %d not implemented", m->activation);
  NSR_REQUIRE(m->out_activation >= 0 && m->out_activation <= 3, "output activation %d not implemented", m->out_activation);
  return 0;
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    nsr_set_error("%s: cannot reserve %zu B of shared memory: %s", name, bytes, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

template <bool BIAS>
int launch_mlp_fwd(const char* who, const nsr_mlp_t* m, const void* x_h, const void* params_h, const float* bias, void* out_h,
                   float* out_f32, int64_t n, void* stream) {
  if (int e = check_mlp(m)) return e;
  if (n == 0) return 0;
  NSR_REQUIRE(!BIAS || (bias != nullptr && out_f32 != nullptr), "%s: bias / output is NULL", who);
  const int in_pad = (m->n_in + 15) / 16 * 16, kt = in_pad / 16;
  const MlpSmem L = mlp_smem_layout(in_pad, m->n_hidden);
  const size_t smem = (size_t)(L.w_total + kFwdWarps * (32 * L.ld1 + 32 * 24)) * sizeof(__half);
  const int64_t tiles = (n + 31) / 32;
  int grid = (int)min((int64_t)nsr_sm_count() * 4, (tiles + kFwdWarps - 1) / kFwdWarps);
  if (grid < 1) grid = 1;
#define NSR_LAUNCH_FWD(KT)                                                                                              \
  case KT:                                                                                                              \
    if (int e = set_smem(mlp_fwd_kernel<KT, BIAS>, smem, who)) return e;                                                \
    mlp_fwd_kernel<KT, BIAS><<<grid, kFwdWarps * 32, smem, (cudaStream_t)stream>>>(*m, (const __half*)x_h, (const __half*)params_h, bias, \
                                                                                   (__half*)out_h, out_f32, n);         \
    break;
  switch (kt) {
    NSR_LAUNCH_FWD(1) NSR_LAUNCH_FWD(2) NSR_LAUNCH_FWD(3) NSR_LAUNCH_FWD(4)
    default: NSR_REQUIRE(false, "%s: unsupported padded input width %d", who, in_pad);
  }
#undef NSR_LAUNCH_FWD
  NSR_CHECK_LAUNCH(who);
  return 0;
}

template <bool BIAS>
int launch_mlp_bwd(const char* who, const nsr_mlp_t* m, const void* x_h, const void* params_h, const float* bias, const void* dy_h,
                   const float* dy_f32, float* grad_params, float* grad_bias, void* dx_h, float* dx_f32, float loss_scale,
                   const float* amax, int64_t n, void* stream) {
  if (int e = check_mlp(m)) return e;
  NSR_REQUIRE(loss_scale > 0.f || (BIAS && amax != nullptr), "%s: loss_scale must be > 0%s", who,
              BIAS ? " (or <= 0 with the amax pointer: automatic)" : "");
  if (n == 0) return 0;
  NSR_REQUIRE(!BIAS || (bias != nullptr && dy_f32 != nullptr && grad_bias != nullptr), "%s: bias / dy / grad_bias is NULL", who);
  const int in_pad = (m->n_in + 15) / 16 * 16, kt = in_pad / 16;
  const MlpSmem L = mlp_smem_layout(in_pad, m->n_hidden);
  const size_t smem = (size_t)(L.w_total + kBwdRows * L.ld1 + 2 * m->n_hidden * kBwdRows * NSR_LD64 + kBwdRows * 24) * sizeof(__half);
  const int64_t tiles = (n + kBwdRows - 1) / kBwdRows;
  int grid = (int)min((int64_t)nsr_sm_count() * (smem <= 110 * 1024 ? 2 : 1), tiles);
  if (grid < 1) grid = 1;
#define NSR_LAUNCH_BWD(KT)                                                                                              \
  case KT:                                                                                                              \
    if (int e = set_smem(mlp_bwd_kernel<KT, BIAS>, smem, who)) return e;                                                \
    mlp_bwd_kernel<KT, BIAS><<<grid, kBwdWarps * 32, smem, (cudaStream_t)stream>>>(*m, (const __half*)x_h, (const __half*)params_h, bias, \
                                                                                   (const __half*)dy_h, dy_f32, grad_params, grad_bias, \
                                                                                   (__half*)dx_h, dx_f32, loss_scale, amax, n); \
    break;
  switch (kt) {
    NSR_LAUNCH_BWD(1) NSR_LAUNCH_BWD(2) NSR_LAUNCH_BWD(3) NSR_LAUNCH_BWD(4)
    default: NSR_REQUIRE(false, "%s: unsupported padded input width %d", who, in_pad);
  }
#undef NSR_LAUNCH_BWD
  NSR_CHECK_LAUNCH(who);
  return 0;
}

}  // namespace

extern "C" int nsr_mlp_fwd(const nsr_mlp_t* m, const void* x_h, const void* params_h, void* out_h, int64_t n, void* stream) {
  return launch_mlp_fwd<false>("nsr_mlp_fwd", m, x_h, params_h, nullptr, out_h, nullptr, n, stream);
}

extern "C" int nsr_mlp_bwd(const nsr_mlp_t* m, const void* x_h, const void* params_h, const void* y_h, const void* dy_h,
                           float* grad_params, void* dx_h, float loss_scale, int64_t n, void* stream) {
  (void)y_h;  // the forward is recomputed on tensor cores; y is accepted for interface symmetry
  return launch_mlp_bwd<false>("nsr_mlp_bwd", m, x_h, params_h, nullptr, dy_h, nullptr, grad_params, nullptr, dx_h, nullptr, loss_scale,
                               nullptr, n, stream);
}

extern "C" int nsr_mlp_vanilla_fwd(const nsr_mlp_t* m, const void* x_h, const void* weights_h, const float* bias, float* out, int64_t n,
                                   void* stream) {
  return launch_mlp_fwd<true>("nsr_mlp_vanilla_fwd", m, x_h, weights_h, bias, nullptr, out, n, stream);
}

extern "C" int nsr_mlp_vanilla_bwd(const nsr_mlp_t* m, const void* x_h, const void* weights_h, const float* bias, const float* dy,
                                   float* grad_weights, float* grad_bias, float* dx, float loss_scale, const float* amax, int64_t n,
                                   void* stream) {
  return launch_mlp_bwd<true>("nsr_mlp_vanilla_bwd", m, x_h, weights_h, bias, nullptr, dy, grad_weights, grad_bias, nullptr, dx, loss_scale,
                              amax, n, stream);
}
