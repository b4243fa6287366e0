// Fused NeRF backward through both networks and the hash grid (autograd of VolumeRadiance + VolumeDensity +
// HashGrid, models/texture.py:23-30, models/geometry.py:122-130), one launch.
//
// Per 64-sample CTA tile (4 warps x 16 rows, two CTAs per SM): reload the 64 B/sample encoded features saved by the forward,
// recompute every activation on tensor cores (cheaper than storing 416 B/sample of hidden state), run the dgrad
// chain in registers, scatter dL/d(table) straight from the mma accumulator layout (each thread owns 2 samples x 4
// levels) with 8-byte vector REDs into the fp32 gradient table, then split the five weight-gradient GEMMs
// (dW = dPre^T * Act, K = 64 samples) over the 4 warps with register accumulators that persist for the whole
// kernel; one atomicAdd per weight per CTA at the end (tcnn: split-K CUTLASS GEMMs over K = batch + reduction).
#include <stdlib.h>
#include "nerf_fused.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kRows = kWarps * 16;  // 64-row CTA tile: ~96 KB of smem => 2 CTAs / SM whose phases (MMA / scatter / wgrad) overlap
constexpr int kCtasPerSm = 2;

// smem tile offsets (halves) after the weights
constexpr int T_X0 = 0;                         // [64][40]  encoded features
constexpr int T_H1 = T_X0 + kRows * NF_LD32;    // [64][72]  density hidden (post ReLU)
constexpr int T_CI = T_H1 + kRows * NSR_LD64;   // [64][40]  colour input: out16 | SH16
constexpr int T_G1 = T_CI + kRows * NF_LD32;    // [64][72]
constexpr int T_G2 = T_G1 + kRows * NSR_LD64;   // [64][72]
constexpr int T_DC3 = T_G2 + kRows * NSR_LD64;  // [64][24]  d(rgb pre-activation)
constexpr int T_DG2 = T_DC3 + kRows * 24;       // [64][72]
constexpr int T_DG1 = T_DG2 + kRows * NSR_LD64; // [64][72]
constexpr int T_DO = T_DG1 + kRows * NSR_LD64;  // [64][24]  d(out16)
constexpr int T_DH1 = T_DO + kRows * 24;        // [64][72]
constexpr int T_X0B = T_DH1 + kRows * NSR_LD64; // [64][40]  second encoded-feature buffer (packed mode: cp.async double buffering)
constexpr int T_TOTAL = T_X0B + kRows * NF_LD32;
// packed mode: per-row inputs of two tiles in flight (floats): xyz+dir [64][6], d_sraw [64], d_rgb [64][3]
constexpr int S_XYZ = 0, S_DS = S_XYZ + kRows * 6, S_DRGB = S_DS + kRows, S_ROWF = S_DRGB + kRows * 3;  // 640 floats per buffer
constexpr size_t kSmemBytes = (size_t)(NF_W_TOTAL + T_TOTAL) * sizeof(__half) + 2 * S_ROWF * sizeof(float);

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int kSlots = 40 / kWarps;  // 40 wgrad pair-tiles split over the warps

struct WgradTile {
  int dy_off, ldy, x_off, ldx, m0, n0;  // smem tiles
  int net, base, in_dim;                // destination in the flat parameter gradient (net 0 density, 1 colour)
};

__device__ __forceinline__ WgradTile wgrad_tile(int t) {
  WgradTile w;
  if (t < 8) {          // density W1 [64][32]
    w = {T_DH1, NSR_LD64, T_X0, NF_LD32, (t / 2) * 16, (t % 2) * 16, 0, 0, 32};
  } else if (t < 12) {  // density W2 [16][64]
    w = {T_DO, 24, T_H1, NSR_LD64, 0, (t - 8) * 16, 0, 64 * 32, 64};
  } else if (t < 20) {  // colour W1 [64][32]
    const int u = t - 12;
    w = {T_DG1, NSR_LD64, T_CI, NF_LD32, (u / 2) * 16, (u % 2) * 16, 1, 0, 32};
  } else if (t < 36) {  // colour W2 [64][64]
    const int u = t - 20;
    w = {T_DG2, NSR_LD64, T_G1, NSR_LD64, (u / 4) * 16, (u % 4) * 16, 1, 64 * 32, 64};
  } else {              // colour W3 [16][64]
    w = {T_DC3, 24, T_G2, NSR_LD64, 0, (t - 36) * 16, 1, 64 * 32 + 64 * 64, 64};
  }
  return w;
}

// mask a dgrad accumulator with the ReLU of the post-activation fragments and pack to fp16 A fragments
__device__ __forceinline__ void relu_mask_pack(const float (&acc)[1][8][4], const uint32_t (&post)[1][4][4], uint32_t (&out)[1][4][4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 hv = *reinterpret_cast<const __half2*>(&post[0][k][j]);
      const int nt = 2 * k + (j >> 1), i0 = (j & 1) * 2;
      const float d0 = __low2float(hv) > 0.f ? acc[0][nt][i0] : 0.f;
      const float d1 = __high2float(hv) > 0.f ? acc[0][nt][i0 + 1] : 0.f;
      out[0][k][j] = nsr_pack_h2(d0, d1);
    }
}

// PACKED: enc_save, d_sraw, d_rgb are in packed row order and xyzdir [n,6] holds the unit-cube position and the view direction of
// every row (written by nsr_pack_kept; buffers padded by one tile).  All global inputs of tile t+1 are then fetched with cp.async
// into the second smem buffer while tile t computes: no load of the kernel sits in front of the math any more (ncu before:
// 37 % of the stall samples were long-scoreboard waits on the row_pos -> enc / ray_indices -> rays chains and on the tile's loads).
// SCATTER = false (split backward, nsr_nerf_field_bwd_split): d(encoding) leaves the kernel as fp16 pairs [n][16 levels] (still multiplied by
// the loss scale) and a second, high-occupancy kernel (nerf_table_scatter_kernel) turns it into table REDs with warp-wide run merging.
template <bool PACKED, bool SCATTER>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) nerf_bwd_kernel(const __grid_constant__ nsr_nerf_t P, const float* __restrict__ rays,
                                                               const int32_t* __restrict__ ray_indices, const float* __restrict__ t_starts,
                                                               const float* __restrict__ t_ends, const __half* __restrict__ enc_save,
                                                               const __half* __restrict__ dparams, const __half* __restrict__ cparams,
                                                               const float* __restrict__ d_sraw, const float* __restrict__ d_rgb,
                                                               float* __restrict__ grad_dparams, float* __restrict__ grad_cparams,
                                                               float loss_scale, const float* __restrict__ amax_ptr, int64_t n_cap, const int64_t* __restrict__ n_dev,
                                                               const int64_t* __restrict__ row_pos, const float* __restrict__ xyzdir, uint32_t* __restrict__ denc_out) {
  const int64_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
  extern __shared__ __align__(16) __half smem[];
  __half* T = smem + NF_W_TOTAL;
  float* rowf = reinterpret_cast<float*>(smem + NF_W_TOTAL + T_TOTAL);  // [2][S_ROWF]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int r0 = warp * 16;
  if (loss_scale <= 0.f) {  // automatic: bring the largest incoming gradient to ~2^8
    const float amax = fmaxf(__ldg(amax_ptr), 1e-30f);
    loss_scale = exp2f(fminf(fmaxf(floorf(log2f(256.f / amax)), -24.f), 60.f));
  }
  const float inv_scale = 1.f / loss_scale;
  nf_stage_weights(smem, dparams, cparams, true);
  float* grad_table = grad_dparams + NF_DENSITY_PARAMS;

  float wacc[kSlots][2][4];
#pragma unroll
  for (int s = 0; s < kSlots; ++s)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) wacc[s][j][i] = 0.f;

  const int64_t n_tiles = (n + kRows - 1) / kRows;
  // packed mode: asynchronous fetch of one tile's inputs into buffer `buf` (all threads; 16-byte chunks, L1 bypassed)
  auto fetch_tile = [&](int64_t tile, int buf) {
    const int64_t row0 = tile * kRows;
    __half* xb = T + (buf ? T_X0B : T_X0);
    float* rf = rowf + buf * S_ROWF;
    for (int v = threadIdx.x; v < kRows * 4; v += kThreads)  // 64 rows x 4 chunks of the 64-byte encodings
      cp_async16(xb + (v >> 2) * NF_LD32 + (v & 3) * 8, enc_save + (row0 + (v >> 2)) * 32 + (v & 3) * 8);
    for (int v = threadIdx.x; v < kRows * 6 / 4; v += kThreads) cp_async16(rf + S_XYZ + v * 4, xyzdir + row0 * 6 + v * 4);
    if (threadIdx.x < kRows / 4) cp_async16(rf + S_DS + threadIdx.x * 4, d_sraw + row0 + threadIdx.x * 4);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + kRows * 3 / 4)
      cp_async16(rf + S_DRGB + (threadIdx.x - 64) * 4, d_rgb + row0 * 3 + (threadIdx.x - 64) * 4);
    cp_async_commit();
  };
  int buf = 0;
  if (PACKED && (int64_t)blockIdx.x < n_tiles) fetch_tile(blockIdx.x, 0);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const int64_t row0 = tile * kRows;
    if (PACKED) cp_async_wait_all();  // this tile's inputs have landed (issued one tile ago)
    __syncthreads();  // previous tile's wgrad is done with the smem tiles (first iteration: weights are staged)
    const __half* X0 = T + ((PACKED && buf) ? T_X0B : T_X0);
    const int x0_off = (PACKED && buf) ? T_X0B : T_X0;
    const float* rf = rowf + buf * S_ROWF;
    if (PACKED) {
      if (tile + gridDim.x < n_tiles) fetch_tile(tile + gridDim.x, buf ^ 1);
      if (row0 + kRows > n) {  // last, partial tile: rows >= n hold stale data; zero them so that 0 * garbage never reaches the wgrad sums
        for (int v = threadIdx.x; v < kRows * 4; v += kThreads)
          if (row0 + (v >> 2) >= n) *reinterpret_cast<uint4*>(T + x0_off + (v >> 2) * NF_LD32 + (v & 3) * 8) = make_uint4(0, 0, 0, 0);
        float* rw = rowf + buf * S_ROWF;
        for (int v = threadIdx.x; v < S_ROWF; v += kThreads) {
          const int r = v < S_DS ? v / 6 : (v < S_DRGB ? v - S_DS : (v - S_DRGB) / 3);
          if (row0 + r >= n) rw[v] = 0.f;
        }
        __syncthreads();
      }
    }
    // rows this thread owns in the accumulator layout (g, g+8) and where their per-sample gradients live
    const int64_t ia = row0 + r0 + g, ib = ia + 8;
    const int64_t pa = (!PACKED && row_pos && ia < n) ? row_pos[ia] : ia, pb = (!PACKED && row_pos && ib < n) ? row_pos[ib] : ib;
    // ---- stage this warp's 16 rows: encoded features and SH of the view direction
    if (!PACKED) {
      for (int v = lane; v < 64; v += 32) {
        const int r = v >> 2, q = v & 3;
        const int64_t i = row0 + r0 + r;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (i < n) val = __ldg(reinterpret_cast<const uint4*>(enc_save + (row_pos ? row_pos[i] : i) * 32) + q);
        *reinterpret_cast<uint4*>(T + T_X0 + (r0 + r) * NF_LD32 + q * 8) = val;
      }
    }
    if (lane < 16) {
      const int64_t i = row0 + r0 + lane;
      uint4 s0 = make_uint4(0, 0, 0, 0), s1 = s0;
      if (i < n) {
        float s[16];
        if (PACKED) {
          const float* rr = rf + S_XYZ + (r0 + lane) * 6;
          nsr_sh4(rr[3], rr[4], rr[5], s);
        } else {
          const float* rr = rays + (size_t)ray_indices[i] * 6;
          nsr_sh4(__ldg(rr + 3), __ldg(rr + 4), __ldg(rr + 5), s);
        }
        s0 = make_uint4(nsr_pack_h2(s[0], s[1]), nsr_pack_h2(s[2], s[3]), nsr_pack_h2(s[4], s[5]), nsr_pack_h2(s[6], s[7]));
        s1 = make_uint4(nsr_pack_h2(s[8], s[9]), nsr_pack_h2(s[10], s[11]), nsr_pack_h2(s[12], s[13]), nsr_pack_h2(s[14], s[15]));
      }
      uint4* sp = reinterpret_cast<uint4*>(T + T_CI + (r0 + lane) * NF_LD32 + 16);
      sp[0] = s0;
      sp[1] = s1;
    }
    __syncwarp();

    // ---- forward recompute
    uint32_t a_h1[1][4][4], a_o[1][1][4], a_g1[1][4][4], a_g2[1][4][4];
    float acc[1][8][4], acc16[1][2][4];
    {
      uint32_t a_in[1][2][4];
      nsr_load_afrag<1, 2>(a_in, X0, NF_LD32, r0);
      nsr_zero_acc(acc);
      nsr_gemm_w<1, 2, 8>(acc, a_in, smem + NF_OFF_DW1, NF_LD32);
      nsr_acc_to_afrag<1, 8>(acc, a_h1, NSR_ACT_RELU);
      nsr_store_afrag<1, 4>(a_h1, T + T_H1, NSR_LD64, r0);
      nsr_zero_acc(acc16);
      nsr_gemm_w<1, 4, 2>(acc16, a_h1, smem + NF_OFF_DW2, NSR_LD64);
      nsr_acc_to_afrag<1, 2>(acc16, a_o, NSR_ACT_NONE);
      nsr_store_afrag<1, 1>(a_o, T + T_CI, NF_LD32, r0, 0);
    }
    {
      uint32_t a_c[1][2][4], a_sh[1][1][4];
      nsr_load_afrag<1, 1>(a_sh, T + T_CI + 16, NF_LD32, r0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a_c[0][0][j] = a_o[0][0][j];
        a_c[0][1][j] = a_sh[0][0][j];
      }
      nsr_zero_acc(acc);
      nsr_gemm_w<1, 2, 8>(acc, a_c, smem + NF_OFF_CW1, NF_LD32);
      nsr_acc_to_afrag<1, 8>(acc, a_g1, NSR_ACT_RELU);
      nsr_store_afrag<1, 4>(a_g1, T + T_G1, NSR_LD64, r0);
      nsr_zero_acc(acc);
      nsr_gemm_w<1, 4, 8>(acc, a_g1, smem + NF_OFF_CW2, NSR_LD64);
      nsr_acc_to_afrag<1, 8>(acc, a_g2, NSR_ACT_RELU);
      nsr_store_afrag<1, 4>(a_g2, T + T_G2, NSR_LD64, r0);
      nsr_zero_acc(acc16);
      nsr_gemm_w<1, 4, 2>(acc16, a_g2, smem + NF_OFF_CW3, NSR_LD64);
    }
    // ---- d(rgb pre-activation) = d_rgb * s (1 - s), s = sigmoid(fp16(raw)); columns 0..2 only
    uint32_t a_dc3[1][1][4];
    {
      float dp[4] = {0.f, 0.f, 0.f, 0.f};  // (row g: col c*2, c*2+1), (row g+8: col c*2, c*2+1)
      if (c < 2) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int64_t i = hh ? ib : ia, pi = hh ? pb : pa;
          if (i < n) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = c * 2 + e;
              if (col < 3) {
                const float raw = __half2float(__float2half_rn(acc16[0][0][hh * 2 + e]));
                const float s = 1.f / (1.f + expf(-raw));
                const float dr = PACKED ? rf[S_DRGB + (r0 + g + hh * 8) * 3 + col] : d_rgb[pi * 3 + col];
                dp[hh * 2 + e] = dr * s * (1.f - s) * loss_scale;
              }
            }
          }
        }
      }
      a_dc3[0][0][0] = nsr_pack_h2(dp[0], dp[1]);
      a_dc3[0][0][1] = nsr_pack_h2(dp[2], dp[3]);
      a_dc3[0][0][2] = 0u;
      a_dc3[0][0][3] = 0u;
      nsr_store_afrag<1, 1>(a_dc3, T + T_DC3, 24, r0);
    }
    // ---- dgrad chain
    uint32_t a_d[1][4][4];
    nsr_zero_acc(acc);
    nsr_gemm_wt<1, 1, 8>(acc, a_dc3, smem + NF_OFF_CW3, NSR_LD64);
    relu_mask_pack(acc, a_g2, a_d);
    nsr_store_afrag<1, 4>(a_d, T + T_DG2, NSR_LD64, r0);
    nsr_zero_acc(acc);
    nsr_gemm_wt<1, 4, 8>(acc, a_d, smem + NF_OFF_CW2, NSR_LD64);
    relu_mask_pack(acc, a_g1, a_d);
    nsr_store_afrag<1, 4>(a_d, T + T_DG1, NSR_LD64, r0);
    nsr_zero_acc(acc16);
    nsr_gemm_wt<1, 4, 2>(acc16, a_d, smem + NF_OFF_CW1, NF_LD32);  // first 16 input columns = the geometry features
    if (c == 0) {  // density path: d(out0) += d sigma / d raw (trunc_exp backward folded in by nsr_nerf_ray_bwd)
      if (ia < n) acc16[0][0][0] += (PACKED ? rf[S_DS + r0 + g] : d_sraw[pa]) * loss_scale;
      if (ib < n) acc16[0][0][2] += (PACKED ? rf[S_DS + r0 + g + 8] : d_sraw[pb]) * loss_scale;
    }
    uint32_t a_do[1][1][4];
    nsr_acc_to_afrag<1, 2>(acc16, a_do, NSR_ACT_NONE);
    nsr_store_afrag<1, 1>(a_do, T + T_DO, 24, r0);
    nsr_zero_acc(acc);
    nsr_gemm_wt<1, 1, 8>(acc, a_do, smem + NF_OFF_DW2, NSR_LD64);
    relu_mask_pack(acc, a_h1, a_d);
    nsr_store_afrag<1, 4>(a_d, T + T_DH1, NSR_LD64, r0);
    float accE[1][4][4];
    nsr_zero_acc(accE);
    nsr_gemm_wt<1, 4, 4>(accE, a_d, smem + NF_OFF_DW1, NF_LD32);

    // ---- hash-table scatter straight from the accumulator layout: this thread owns samples (g, g+8) x levels (c, 4+c, 8+c, 12+c).
    // Levels 0..7 (nt < 2): consecutive samples of a ray stay in one cell for several steps, so the 8 lanes that hold the
    // same level for 8 consecutive samples first merge runs of equal cells with a segmented shuffle scan and only the
    // last lane of each run issues the 8 REDs (-40 % REDs overall, far less same-address contention in L2).
    if (!SCATTER) {  // split backward: 4 lanes (c = 0..3) x 4-byte stores = 16 contiguous bytes per (row, nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int64_t i = hh ? ib : ia;
        if (i < n) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) denc_out[i * 16 + nt * 4 + c] = nsr_pack_h2(accE[0][nt][hh * 2], accE[0][nt][hh * 2 + 1]);
        }
      }
    }
#pragma unroll
    for (int hh = 0; hh < (SCATTER ? 2 : 0); ++hh) {
      const int64_t i = hh ? ib : ia;
      const bool ok = i < n;
      float x = 0.f, y = 0.f, z = 0.f, dx, dy, dz;
      if (ok) {
        if (PACKED) {
          const float* rr = rf + S_XYZ + (r0 + g + hh * 8) * 6;
          x = rr[0];
          y = rr[1];
          z = rr[2];
        } else {
          nf_sample_position(P, rays, ray_indices[i], t_starts[i], t_ends[i], x, y, z, dx, dy, dz);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float d0 = ok ? accE[0][nt][hh * 2] * inv_scale : 0.f, d1 = ok ? accE[0][nt][hh * 2 + 1] * inv_scale : 0.f;
        const LevelInfo li = nsr_level(P.grid, nt * 4 + c);
        uint32_t cx, cy, cz, idx[8];
        float fx, fy, fz;
        nsr_pos_fract(x, li.scale, cx, fx);
        nsr_pos_fract(y, li.scale, cy, fy);
        nsr_pos_fract(z, li.scale, cz, fz);
        if (nt < 2) {
          // segmented inclusive scan over g (lanes c, c+4, ..., c+28) keyed by the cell
          const uint32_t key = ok ? (cx + li.res * (cy + li.res * cz)) : (0xFFFFFFF0u + g);
          const uint32_t key_prev = __shfl_up_sync(0xffffffffu, key, 4);
          const bool head = (g == 0) || (key_prev != key);
          const int next_head = __shfl_down_sync(0xffffffffu, (int)head, 4);
          const bool tail = (g == 7) || next_head;
          float v[16];
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float w = nsr_corner_weight(cc, fx, fy, fz);
            v[2 * cc] = w * d0;
            v[2 * cc + 1] = w * d1;
          }
          bool flag = head;
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            const int f_up = __shfl_up_sync(0xffffffffu, (int)flag, 4 * o);
            const bool take = (g >= o) && !flag;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float u = __shfl_up_sync(0xffffffffu, v[e], 4 * o);
              if (take) v[e] += u;
            }
            if (take) flag = f_up;
          }
          if (ok && tail) {
            nsr_corner_indices(li, cx, cy, cz, idx);
            // (pairing x-adjacent corners into one 16-byte RED, nsr_red_corner_pair, wins 27 % in the scatter-only micro-benchmark but
            //  not here: 205 -> 209 us -- at 8 warps/SM the cost is per RED instruction, and the divergent pair test adds instructions)
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
              if (v[2 * cc] != 0.f || v[2 * cc + 1] != 0.f) nsr_red_add_f32x2(grad_table + 2 * (size_t)idx[cc], v[2 * cc], v[2 * cc + 1]);
          }
        } else if (ok && (d0 != 0.f || d1 != 0.f)) {
          nsr_corner_indices(li, cx, cy, cz, idx);
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float w = nsr_corner_weight(cc, fx, fy, fz);
            nsr_red_add_f32x2(grad_table + 2 * (size_t)idx[cc], w * d0, w * d1);
          }
        }
      }
    }
    __syncthreads();
    // ---- weight gradients over the 64 rows of the tile
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const WgradTile w = wgrad_tile(warp + s * kWarps);
      const int xo = (w.x_off == T_X0) ? x0_off : w.x_off;  // the encoded features live in the current double buffer
      nsr_wgrad_tile(wacc[s][0], wacc[s][1], T + w.dy_off, w.ldy, w.m0, T + xo, w.ldx, w.n0, kRows);
    }
  }
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const WgradTile w = wgrad_tile(warp + s * kWarps);
    float* dst = (w.net == 0 ? grad_dparams : grad_cparams) + w.base;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int o = w.m0 + g + ((i >> 1) << 3), ii = w.n0 + j * 8 + c * 2 + (i & 1);
        atomicAdd(dst + (size_t)o * w.in_dim + ii, wacc[s][j][i] * inv_scale);
      }
  }
}

}  // namespace

namespace {

// Table half of the split backward: rows = kept samples in packed (ray-major) order, thread per row, a warp = 32 consecutive rows.
// Per level: corner weights x d(feature pair); on levels < kMergeLevels runs of rows that sit in the same cell (consecutive samples of
// a ray on the coarse levels) are summed with a segmented shuffle scan over the whole warp and only the run's last lane issues REDs;
// x-adjacent corners that are neighbours in memory (hashed levels: cell x even; dense levels: entry index even) leave as ONE 16-byte
// red.global.add.v4.f32.  48 registers, no shared memory: 64 warps per SM keep the RED path of the SM full (tools/gather_bench.py,
// profiles/r2_scatter_microbench.md: 92 us for 267 k samples against 264 us for the plain 8-byte form).
constexpr int kMergeLevels = 8;

template <int MINB>   // resident CTAs per SM the register allocation is sized for: 5 (48 registers), 6 (40), 8 (32, 64 B of spills)
__global__ void __launch_bounds__(256, MINB) nerf_table_scatter_kernel(const __grid_constant__ nsr_grid_t g, const float* __restrict__ xyz, int stride,
                                                                 const __half2* __restrict__ denc, float loss_scale,
                                                                 const float* __restrict__ amax_ptr, float* __restrict__ grad_table,
                                                                 int64_t n_cap, const int64_t* __restrict__ n_dev, int l_begin, int l_end) {
  const int64_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
  if (loss_scale <= 0.f) {  // the same automatic scale as nerf_bwd_kernel derives from the same amax
    const float amax = fmaxf(__ldg(amax_ptr), 1e-30f);
    loss_scale = exp2f(fminf(fmaxf(floorf(log2f(256.f / amax)), -24.f), 60.f));
  }
  const float inv_scale = 1.f / loss_scale;
  const int lane = threadIdx.x & 31;
  const int64_t n32 = (n + 31) & ~31ll;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n32; i += (int64_t)gridDim.x * 256) {
    const bool ok = i < n;
    float x = 0.f, y = 0.f, z = 0.f;
    if (ok) {
      x = xyz[i * stride];
      y = xyz[i * stride + 1];
      z = xyz[i * stride + 2];
    }
#pragma unroll 1
    for (int l = l_begin; l < l_end; ++l) {  // the data-parallel step scatters level groups in separate launches (gradient exchange overlap)
      float2 d = make_float2(0.f, 0.f);
      if (ok) {
        d = __half22float2(denc[i * 16 + l]);
        d.x *= inv_scale;
        d.y *= inv_scale;
      }
      const LevelInfo li = nsr_level(g, l);
      uint32_t cx, cy, cz, idx[8];
      float fx, fy, fz;
      nsr_pos_fract(x, li.scale, cx, fx);
      nsr_pos_fract(y, li.scale, cy, fy);
      nsr_pos_fract(z, li.scale, cz, fz);
      float v[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float w = nsr_corner_weight(c, fx, fy, fz);
        v[2 * c] = w * d.x;
        v[2 * c + 1] = w * d.y;
      }
      bool issue = ok && (d.x != 0.f || d.y != 0.f);
      if (l < kMergeLevels) {  // res^3 < 2^32 on these levels (base 16 .. 32, growth <= 1.45): the cell index is a unique key
        const uint32_t key = ok ? (cx + li.res * (cy + li.res * cz)) : (0xFFFFFF00u + lane);
        const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = lane == 0 || prev != key;
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const int my_head = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));  // first lane of my run
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        int maxrun = lane - my_head + 1;  // the longest run of the warp bounds the number of scan steps (warp-uniform)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxrun = max(maxrun, __shfl_xor_sync(0xffffffffu, maxrun, o));
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          if (o < maxrun) {
            const bool take = lane - o >= my_head;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float u = __shfl_up_sync(0xffffffffu, v[e], o);
              if (take) v[e] += u;
            }
          }
        }
        issue = ok && tail;
      }
      if (issue) {
        nsr_corner_indices(li, cx, cy, cz, idx);
#pragma unroll
        for (int c = 0; c < 8; c += 2) nsr_red_corner_pair(grad_table, idx[c], idx[c + 1], v[2 * c], v[2 * c + 1], v[2 * c + 2], v[2 * c + 3]);
      }
    }
  }
}

int field_bwd_launch(const nsr_nerf_t* f, const float* rays, const int32_t* ray_indices, const float* t_starts, const float* t_ends,
                     const void* enc_save_h, const void* dparams_h, const void* cparams_h, const float* d_sraw, const float* d_rgb,
                     float* grad_dparams, float* grad_cparams, float loss_scale, const float* amax, int64_t k, const int64_t* k_dev,
                     const int64_t* row_pos, const float* xyzdir, void* denc_out, void* stream, const char* who) {
  NSR_REQUIRE(f != nullptr, "%s: field descriptor is NULL", who);
  NSR_REQUIRE(f->grid.n_levels == 16 && f->grid.n_features == 2 && f->feature_dim == 16 && f->density_hidden == 1 && f->color_hidden == 2,
              "%s: fused path needs L=16, F=2, feature_dim=16, hidden layers 1/2", who);
  NSR_REQUIRE(loss_scale > 0.f || amax != nullptr, "%s: loss_scale <= 0 (automatic) needs the amax pointer", who);
  if (k == 0) return 0;
  NSR_REQUIRE(xyzdir == nullptr || row_pos == nullptr, "%s: packed inputs (xyzdir) and row_pos are mutually exclusive", who);
  NSR_REQUIRE(denc_out == nullptr || xyzdir != nullptr, "%s: the split form needs the packed inputs (xyzdir)", who);
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(nerf_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nerf_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nerf_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) {
      nsr_set_error("%s: cannot reserve %zu B shared memory: %s", who, kSmemBytes, cudaGetErrorString(e));
      return 2;
    }
    attr_set = true;
  }
  const int64_t tiles = (k + kRows - 1) / kRows;
  int grid = (int)min((int64_t)nsr_sm_count() * kCtasPerSm, tiles);
  if (k_dev != nullptr) grid = nsr_sm_count() * kCtasPerSm;
#define NSR_BWD_ARGS                                                                                                                        \
  *f, rays, ray_indices, t_starts, t_ends, (const __half*)enc_save_h, (const __half*)dparams_h, (const __half*)cparams_h, d_sraw, d_rgb,    \
      grad_dparams, grad_cparams, loss_scale, amax, k, k_dev, row_pos, xyzdir, (uint32_t*)denc_out
  if (denc_out != nullptr)
    nerf_bwd_kernel<true, false><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(NSR_BWD_ARGS);
  else if (xyzdir != nullptr)
    nerf_bwd_kernel<true, true><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(NSR_BWD_ARGS);
  else
    nerf_bwd_kernel<false, true><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(NSR_BWD_ARGS);
#undef NSR_BWD_ARGS
  NSR_CHECK_LAUNCH(who);
  return 0;
}

}  // namespace

extern "C" int nsr_nerf_field_bwd(const nsr_nerf_t* f, const float* rays, const int32_t* ray_indices, const float* t_starts,
                                  const float* t_ends, const void* enc_save_h, const void* dparams_h, const void* cparams_h,
                                  const float* d_sraw, const float* d_rgb, float* grad_dparams, float* grad_cparams, float loss_scale,
                                  const float* amax, int64_t k, const int64_t* k_dev, const int64_t* row_pos, const float* xyzdir,
                                  void* stream) {
  return field_bwd_launch(f, rays, ray_indices, t_starts, t_ends, enc_save_h, dparams_h, cparams_h, d_sraw, d_rgb, grad_dparams, grad_cparams,
                          loss_scale, amax, k, k_dev, row_pos, xyzdir, nullptr, stream, "nsr_nerf_field_bwd");
}

// The two halves of the split backward as separate entry points (what the Python side calls, so that each half shows up with its own
// duration in bench.py's per-kernel table):
//   nsr_nerf_field_bwd_net   : MLP recompute + dgrad + wgrad over the packed rows; d(encoding) -> denc_h (fp16 [k,32], still multiplied by
//                              the loss scale); grad_dparams receives only the density network's weight gradients
//   nsr_nerf_table_scatter   : levels [level_begin, level_end) of denc_h -> fp32 REDs into grad_table (= grad_dparams + the density network's parameter count); xyz = the
//                              packed unit-cube positions with row stride `stride` floats (6 for the xyzdir buffer of nsr_pack_kept)
extern "C" int nsr_nerf_field_bwd_net(const nsr_nerf_t* f, const void* enc_k_h, const void* dparams_h, const void* cparams_h, const float* d_sraw,
                                      const float* d_rgb, float* grad_dparams, float* grad_cparams, float loss_scale, const float* amax, int64_t k,
                                      const int64_t* k_dev, const float* xyzdir, void* denc_h, void* stream) {
  NSR_REQUIRE(denc_h != nullptr && xyzdir != nullptr, "nsr_nerf_field_bwd_net: denc / xyzdir is NULL");
  return field_bwd_launch(f, nullptr, nullptr, nullptr, nullptr, enc_k_h, dparams_h, cparams_h, d_sraw, d_rgb, grad_dparams, grad_cparams, loss_scale,
                          amax, k, k_dev, nullptr, xyzdir, denc_h, stream, "nsr_nerf_field_bwd_net");
}

extern "C" int nsr_nerf_table_scatter(const nsr_grid_t* g, const float* xyz, int32_t stride, const void* denc_h, float loss_scale, const float* amax,
                                      float* grad_table, int64_t k, const int64_t* k_dev, int32_t level_begin, int32_t level_end, int32_t ctas_per_sm,
                                      void* stream) {
  NSR_REQUIRE(g != nullptr && xyz != nullptr && denc_h != nullptr && grad_table != nullptr, "nsr_nerf_table_scatter: NULL argument");
  NSR_REQUIRE(g->n_levels == 16 && g->n_features == 2, "nsr_nerf_table_scatter: needs L=16, F=2");
  NSR_REQUIRE(loss_scale > 0.f || amax != nullptr, "nsr_nerf_table_scatter: loss_scale <= 0 (automatic) needs the amax pointer");
  NSR_REQUIRE(stride >= 3, "nsr_nerf_table_scatter: stride must be >= 3");
  NSR_REQUIRE(level_begin >= 0 && level_begin <= level_end && level_end <= 16, "nsr_nerf_table_scatter: bad level range [%d, %d)", level_begin, level_end);
  if (k == 0 || level_begin == level_end) return 0;
  const int per_sm = ctas_per_sm >= 1 && ctas_per_sm <= 8 ? ctas_per_sm : 8;   // < 8 leaves room for a kernel running beside it (the exchange)
  const int grid = (int)min((int64_t)nsr_sm_count() * per_sm, (k + 255) / 256);
  static const int occ = [] {   // NSR_SCATTER_CTAS = 5 (default) | 6 | 8: register budget of the instantiation
    const char* v = getenv("NSR_SCATTER_CTAS");
    const int n = v ? atoi(v) : 5;
    return n == 6 || n == 8 ? n : 5;
  }();
  const int g_ = k_dev ? nsr_sm_count() * per_sm : grid;
  cudaStream_t st_ = (cudaStream_t)stream;
  const __half2* de_ = (const __half2*)denc_h;
  if (occ == 6)
    nerf_table_scatter_kernel<6><<<g_, 256, 0, st_>>>(*g, xyz, stride, de_, loss_scale, amax, grad_table, k, k_dev, level_begin, level_end);
  else if (occ == 8)
    nerf_table_scatter_kernel<8><<<g_, 256, 0, st_>>>(*g, xyz, stride, de_, loss_scale, amax, grad_table, k, k_dev, level_begin, level_end);
  else
    nerf_table_scatter_kernel<5><<<g_, 256, 0, st_>>>(*g, xyz, stride, de_, loss_scale, amax, grad_table, k, k_dev, level_begin, level_end);
  NSR_CHECK_LAUNCH("nsr_nerf_table_scatter");
  return 0;
}

// Split form of the same backward (packed inputs only): kernel 1 = MLP recompute + dgrad + wgrad, d(encoding) -> denc_h (fp16 [k,32],
// still multiplied by the loss scale); kernel 2 = nerf_table_scatter_kernel over the same rows (grad_dparams + NF_DENSITY_PARAMS).
extern "C" int nsr_nerf_field_bwd_split(const nsr_nerf_t* f, const void* enc_k_h, const void* dparams_h, const void* cparams_h,
                                        const float* d_sraw, const float* d_rgb, float* grad_dparams, float* grad_cparams, float loss_scale,
                                        const float* amax, int64_t k, const int64_t* k_dev, const float* xyzdir, void* denc_h, void* stream) {
  NSR_REQUIRE(denc_h != nullptr && xyzdir != nullptr, "nsr_nerf_field_bwd_split: denc / xyzdir is NULL");
  const int rc = field_bwd_launch(f, nullptr, nullptr, nullptr, nullptr, enc_k_h, dparams_h, cparams_h, d_sraw, d_rgb, grad_dparams, grad_cparams,
                                  loss_scale, amax, k, k_dev, nullptr, xyzdir, denc_h, stream, "nsr_nerf_field_bwd_split");
  if (rc != 0 || k == 0) return rc;
  const int grid = (int)min((int64_t)nsr_sm_count() * 8, (k + 255) / 256);
  nerf_table_scatter_kernel<5><<<k_dev ? nsr_sm_count() * 8 : grid, 256, 0, (cudaStream_t)stream>>>(f->grid, xyzdir, 6, (const __half2*)denc_h, loss_scale, amax,
                                                                                                 grad_dparams + NF_DENSITY_PARAMS, k, k_dev, 0, 16);
  NSR_CHECK_LAUNCH("nsr_nerf_field_bwd_split (scatter)");
  return 0;
}
