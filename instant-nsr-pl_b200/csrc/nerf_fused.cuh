// Shared pieces of the fused NeRF kernels (nerf_fused_fwd.cu / nerf_fused_bwd.cu): field description,
// shared-memory weight layout, sample -> position, the 16-level gather.
#pragma once
#include "mlp_warp.cuh"

// Weights of the two nerf-blender networks in shared memory (halves), padded rows:
//   density  W1 [64][32+8], W2 [16][64+8]            (geometry.mlp_network_config: 1 hidden layer)
//   colour   W1 [64][32+8], W2 [64][64+8], W3 [16][64+8]   (texture.mlp_network_config: 2 hidden layers)
constexpr int NF_LD32 = 32 + NSR_LDW_PAD;  // 40
constexpr int NF_OFF_DW1 = 0;
constexpr int NF_OFF_DW2 = NF_OFF_DW1 + 64 * NF_LD32;
constexpr int NF_OFF_CW1 = NF_OFF_DW2 + 16 * NSR_LD64;
constexpr int NF_OFF_CW2 = NF_OFF_CW1 + 64 * NF_LD32;
constexpr int NF_OFF_CW3 = NF_OFF_CW2 + 64 * NSR_LD64;
constexpr int NF_W_TOTAL = NF_OFF_CW3 + 16 * NSR_LD64;  // 12032 halves = 24064 B
constexpr int NF_DENSITY_PARAMS = 64 * 32 + 16 * 64;     // 3072
constexpr int NF_COLOR_PARAMS = 64 * 32 + 64 * 64 + 16 * 64;  // 7168

__device__ __forceinline__ void nf_stage_weights(__half* smem, const __half* __restrict__ dparams, const __half* __restrict__ cparams,
                                                 bool with_color) {
  const int tid = threadIdx.x, nt = blockDim.x;
  nsr_stage_matrix(smem + NF_OFF_DW1, dparams, 64, 32, tid, nt);
  nsr_stage_matrix(smem + NF_OFF_DW2, dparams + 64 * 32, 16, 64, tid, nt);
  if (with_color) {
    nsr_stage_matrix(smem + NF_OFF_CW1, cparams, 64, 32, tid, nt);
    nsr_stage_matrix(smem + NF_OFF_CW2, cparams + 64 * 32, 64, 64, tid, nt);
    nsr_stage_matrix(smem + NF_OFF_CW3, cparams + 64 * 32 + 64 * 64, 16, 64, tid, nt);
  }
}

// sample (ray, t0, t1) -> unit-cube position (contract_to_unisphere, AABB: models/geometry.py:17-19)
__device__ __forceinline__ void nf_sample_position(const nsr_nerf_t& P, const float* __restrict__ rays, int ray, float t0, float t1,
                                                   float& x, float& y, float& z, float& dx, float& dy, float& dz) {
  const float* r = rays + (size_t)ray * 6;
  const float ox = __ldg(r + 0), oy = __ldg(r + 1), oz = __ldg(r + 2);
  dx = __ldg(r + 3); dy = __ldg(r + 4); dz = __ldg(r + 5);
  const float mid = (t0 + t1) * 0.5f;
  const float inv = 1.f / (2.f * P.radius);
  x = (fmaf(dx, mid, ox) + P.radius) * inv;
  y = (fmaf(dy, mid, oy) + P.radius) * inv;
  z = (fmaf(dz, mid, oz) + P.radius) * inv;
}

// all levels of one sample; features packed as half2 per level (fp16 = what tcnn's encoding emits)
template <int L>
__device__ __forceinline__ void nf_gather(const nsr_grid_t& g, const __half2* __restrict__ table, float x, float y, float z,
                                          uint32_t (&f)[L]) {
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const LevelInfo li = nsr_level(g, l);
    uint32_t cx, cy, cz, idx[8];
    float fx, fy, fz;
    nsr_pos_fract(x, li.scale, cx, fx);
    nsr_pos_fract(y, li.scale, cy, fy);
    nsr_pos_fract(z, li.scale, cz, fz);
    nsr_corner_indices(li, cx, cy, cz, idx);
    float2 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = nsr_ld_table(table, idx[c]);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = nsr_corner_weight(c, fx, fy, fz);
      a0 = fmaf(w, v[c].x, a0);
      a1 = fmaf(w, v[c].y, a1);
    }
    f[l] = nsr_pack_h2(a0, a1);
  }
}

// write one 32-feature row (16 packed half2) of a [rows][NF_LD32] smem tile
__device__ __forceinline__ void nf_store_row32(__half* tile, int row, const uint32_t (&f)[16]) {
  uint4* p = reinterpret_cast<uint4*>(tile + row * NF_LD32);
  p[0] = make_uint4(f[0], f[1], f[2], f[3]);
  p[1] = make_uint4(f[4], f[5], f[6], f[7]);
  p[2] = make_uint4(f[8], f[9], f[10], f[11]);
  p[3] = make_uint4(f[12], f[13], f[14], f[15]);
}

__device__ __forceinline__ float nf_half_lo(uint32_t v) { return __low2float(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ float nf_half_hi(uint32_t v) { return __high2float(*reinterpret_cast<const __half2*>(&v)); }
