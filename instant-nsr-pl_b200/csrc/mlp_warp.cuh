// Warp-level building blocks of the 64-wide fully-fused MLP (tiny-cuda-nn FullyFusedMLP semantics:
// bias-free, fp16 weights/activations; here with fp32 accumulation).  A warp owns MT*16 samples
// (rows); layer outputs stay in registers as mma accumulators and are re-packed as the next
// layer's A fragments without touching memory.  Weights live in shared memory, row-major [out][in]
// (tcnn's flat-parameter layout, models/network_utils.py:142-173) with a +8-half row pad so that
// ldmatrix is bank-conflict free.
#pragma once
#include "common.cuh"

constexpr int NSR_W = 64;        // hidden width (n_neurons) -- every reference config uses 64
constexpr int NSR_LDW_PAD = 8;   // row padding (halves) of every smem matrix / activation tile
constexpr int NSR_LD64 = NSR_W + NSR_LDW_PAD;  // 72

enum { NSR_ACT_NONE = 0, NSR_ACT_RELU = 1, NSR_ACT_SIGMOID = 2, NSR_ACT_EXP = 3 };

__device__ __forceinline__ float nsr_apply_act(float x, int act) {
  switch (act) {
    case NSR_ACT_RELU: return fmaxf(x, 0.f);
    case NSR_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    case NSR_ACT_EXP: return __expf(x);
    default: return x;
  }
}
// derivative expressed through the post-activation value y
__device__ __forceinline__ float nsr_act_grad_from_out(float y, int act) {
  switch (act) {
    case NSR_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case NSR_ACT_SIGMOID: return y * (1.f - y);
    case NSR_ACT_EXP: return y;
    default: return 1.f;
  }
}

template <int MT, int NT>
__device__ __forceinline__ void nsr_zero_acc(float (&acc)[MT][NT][4]) {
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
}

// A fragments (MT x KT tiles of 16x16) from a row-major fp16 smem tile: rows row0.., cols 0..KT*16
template <int MT, int KT>
__device__ __forceinline__ void nsr_load_afrag(uint32_t (&a)[MT][KT][4], const __half* tile, int ld, int row0) {
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int k = 0; k < KT; ++k)
      nsr_ldmatrix_x4(a[m][k], tile + (size_t)(row0 + m * 16 + (mi & 1) * 8 + r) * ld + k * 16 + (mi >> 1) * 8);
}

// acc[MT][NT] += A[MT][KT] * W^T, W smem row-major [NT*8 rows (out)][ldw], k (in) contiguous
template <int MT, int KT, int NT>
__device__ __forceinline__ void nsr_gemm_w(float (&acc)[MT][NT][4], const uint32_t (&a)[MT][KT][4], const __half* Wsm, int ldw) {
  static_assert(NT % 2 == 0, "NT must be even");
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      nsr_ldmatrix_x4(b, Wsm + (size_t)(np * 16 + (mi >> 1) * 8 + r) * ldw + k * 16 + (mi & 1) * 8);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        nsr_mma16816(acc[m][2 * np], a[m][k], b[0], b[1]);
        nsr_mma16816(acc[m][2 * np + 1], a[m][k], b[2], b[3]);
      }
    }
}

// dgrad: acc[MT][NT (in)] += A[MT][KT (out)] * W, W smem row-major [KT*16 rows (out)][ldw]
template <int MT, int KT, int NT>
__device__ __forceinline__ void nsr_gemm_wt(float (&acc)[MT][NT][4], const uint32_t (&a)[MT][KT][4], const __half* Wsm, int ldw) {
  static_assert(NT % 2 == 0, "NT must be even");
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int k = 0; k < KT; ++k)
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      nsr_ldmatrix_x4_trans(b, Wsm + (size_t)(k * 16 + (mi & 1) * 8 + r) * ldw + np * 16 + (mi >> 1) * 8);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        nsr_mma16816(acc[m][2 * np], a[m][k], b[0], b[1]);
        nsr_mma16816(acc[m][2 * np + 1], a[m][k], b[2], b[3]);
      }
    }
}

// accumulators -> next layer's A fragments (fp16), with activation
template <int MT, int NT>
__device__ __forceinline__ void nsr_acc_to_afrag(const float (&acc)[MT][NT][4], uint32_t (&a)[MT][NT / 2][4], int act) {
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int k = 0; k < NT / 2; ++k) {
      a[m][k][0] = nsr_pack_h2(nsr_apply_act(acc[m][2 * k][0], act), nsr_apply_act(acc[m][2 * k][1], act));
      a[m][k][1] = nsr_pack_h2(nsr_apply_act(acc[m][2 * k][2], act), nsr_apply_act(acc[m][2 * k][3], act));
      a[m][k][2] = nsr_pack_h2(nsr_apply_act(acc[m][2 * k + 1][0], act), nsr_apply_act(acc[m][2 * k + 1][1], act));
      a[m][k][3] = nsr_pack_h2(nsr_apply_act(acc[m][2 * k + 1][2], act), nsr_apply_act(acc[m][2 * k + 1][3], act));
    }
}

// A fragments (fp16, fragment layout) -> row-major smem tile (rows row0.., cols col0..)
template <int MT, int KT>
__device__ __forceinline__ void nsr_store_afrag(const uint32_t (&a)[MT][KT][4], __half* tile, int ld, int row0, int col0 = 0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      __half* p = tile + (size_t)(row0 + m * 16 + g) * ld + col0 + k * 16 + c * 2;
      *reinterpret_cast<uint32_t*>(p) = a[m][k][0];
      *reinterpret_cast<uint32_t*>(p + 8 * ld) = a[m][k][1];
      *reinterpret_cast<uint32_t*>(p + 8) = a[m][k][2];
      *reinterpret_cast<uint32_t*>(p + 8 * ld + 8) = a[m][k][3];
    }
}

// element coordinates of accumulator entry i of tile (m, n) for this lane: row, col
__device__ __forceinline__ int nsr_acc_row(int m, int i) { return m * 16 + ((threadIdx.x & 31) >> 2) + ((i >> 1) << 3); }
__device__ __forceinline__ int nsr_acc_col(int n, int i) { return n * 8 + ((threadIdx.x & 3) << 1) + (i & 1); }

// stage the flat fp16 parameter vector (row-major [out][in] matrices) into padded smem matrices
__device__ __forceinline__ void nsr_stage_matrix(__half* dst, const __half* __restrict__ src, int rows, int cols, int tid, int nthreads) {
  const int ld = cols + NSR_LDW_PAD;
  const int vec_per_row = cols / 8;
  for (int i = tid; i < rows * vec_per_row; i += nthreads) {
    const int r = i / vec_per_row, v = i % vec_per_row;
    *reinterpret_cast<uint4*>(dst + (size_t)r * ld + v * 8) = *reinterpret_cast<const uint4*>(src + (size_t)r * cols + v * 8);
  }
}

// wgrad tile: acc(16 out x 8 in... as m16n8) += dY^T[16 out][ks samples] * X[ks samples][8 in] over `rows` samples.
// dY tile smem row-major [samples][ldy] (cols = out), X tile smem row-major [samples][ldx] (cols = in).
// Computes TWO n8 tiles (n0, n0+8) at once.
__device__ __forceinline__ void nsr_wgrad_tile(float (&acc0)[4], float (&acc1)[4], const __half* dY, int ldy, int m0, const __half* X,
                                               int ldx, int n0, int rows) {
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
  for (int s0 = 0; s0 < rows; s0 += 16) {
    uint32_t a[4], b[4];
    nsr_ldmatrix_x4_trans(a, dY + (size_t)(s0 + (mi >> 1) * 8 + r) * ldy + m0 + (mi & 1) * 8);
    nsr_ldmatrix_x4_trans(b, X + (size_t)(s0 + (mi & 1) * 8 + r) * ldx + n0 + (mi >> 1) * 8);
    nsr_mma16816(acc0, a, b[0], b[1]);
    nsr_mma16816(acc1, a, b[2], b[3]);
  }
}
