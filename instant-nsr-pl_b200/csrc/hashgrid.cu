// Multiresolution hash-grid encoding + SH4: the tiny-cuda-nn `Encoding` surface used by
// models/network_utils.py:47,90 (HashGrid) and models/texture.py:24-25 (SphericalHarmonics).
// Thread-per-sample: one thread walks all levels of its sample so that (a) consecutive samples of a
// ray, which share coarse-level cells, coalesce inside a warp, and (b) the L*F features of a sample
// sit in one thread for the fused kernels (nerf_fused.cu) that reuse these device functions.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// ---- forward: out[n, L*2] fp16 -------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) hashgrid_fwd_kernel(nsr_grid_t g, const float* __restrict__ x,
                                                                const __half2* __restrict__ table,
                                                                __half2* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float px = x[i * 3 + 0], py = x[i * 3 + 1], pz = x[i * 3 + 2];
    __half2* o = out + i * g.n_levels;
#pragma unroll 2
    for (int l = 0; l < g.n_levels; ++l) {
      const LevelInfo li = nsr_level(g, l);
      uint32_t cx, cy, cz, idx[8];
      float fx, fy, fz;
      nsr_pos_fract(px, li.scale, cx, fx);
      nsr_pos_fract(py, li.scale, cy, fy);
      nsr_pos_fract(pz, li.scale, cz, fz);
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
      o[l] = __floats2half2_rn(a0, a1);
    }
  }
}

// ---- backward w.r.t. table: grad_table (fp32) += w_c * dy -----------------------------------------
__global__ void __launch_bounds__(kThreads) hashgrid_bwd_kernel(nsr_grid_t g, const float* __restrict__ x,
                                                                const __half2* __restrict__ dy,
                                                                float* __restrict__ grad_table, float dy_scale, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float px = x[i * 3 + 0], py = x[i * 3 + 1], pz = x[i * 3 + 2];
    const __half2* d = dy + i * g.n_levels;
#pragma unroll 2
    for (int l = 0; l < g.n_levels; ++l) {
      float2 dv = __half22float2(d[l]);
      if (dv.x == 0.f && dv.y == 0.f) continue;
      dv.x *= dy_scale;
      dv.y *= dy_scale;
      const LevelInfo li = nsr_level(g, l);
      uint32_t cx, cy, cz, idx[8];
      float fx, fy, fz;
      nsr_pos_fract(px, li.scale, cx, fx);
      nsr_pos_fract(py, li.scale, cy, fy);
      nsr_pos_fract(pz, li.scale, cz, fz);
      nsr_corner_indices(li, cx, cy, cz, idx);
#pragma unroll
      for (int c = 0; c < 8; c += 2) {  // x-adjacent corners: one 16-byte RED when they are neighbours in memory
        const float w0 = nsr_corner_weight(c, fx, fy, fz), w1 = nsr_corner_weight(c + 1, fx, fy, fz);
        nsr_red_corner_pair(grad_table, idx[c], idx[c + 1], w0 * dv.x, w0 * dv.y, w1 * dv.x, w1 * dv.y);
      }
    }
  }
}

// ---- backward w.r.t. input: dx[n,3] = sum_lf dy_lf * scale_l * sum_c dw_c/dfrac * table[c][f] ------
__global__ void __launch_bounds__(kThreads) hashgrid_bwd_input_kernel(nsr_grid_t g, const float* __restrict__ x,
                                                                      const __half2* __restrict__ table,
                                                                      const float* __restrict__ dy,
                                                                      float* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float px = x[i * 3 + 0], py = x[i * 3 + 1], pz = x[i * 3 + 2];
    const float2* d = reinterpret_cast<const float2*>(dy) + i * g.n_levels;
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll 2
    for (int l = 0; l < g.n_levels; ++l) {
      const LevelInfo li = nsr_level(g, l);
      const float2 dv = d[l];
      uint32_t cx, cy, cz, idx[8];
      float fx, fy, fz;
      nsr_pos_fract(px, li.scale, cx, fx);
      nsr_pos_fract(py, li.scale, cy, fy);
      nsr_pos_fract(pz, li.scale, cz, fz);
      nsr_corner_indices(li, cx, cy, cz, idx);
      float lx = 0.f, ly = 0.f, lz = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 v = nsr_ld_table(table, idx[c]);
        const float s = v.x * dv.x + v.y * dv.y;
        lx = fmaf(nsr_corner_dweight(c, 0, fx, fy, fz), s, lx);
        ly = fmaf(nsr_corner_dweight(c, 1, fx, fy, fz), s, ly);
        lz = fmaf(nsr_corner_dweight(c, 2, fx, fy, fz), s, lz);
      }
      gx = fmaf(li.scale, lx, gx);
      gy = fmaf(li.scale, ly, gy);
      gz = fmaf(li.scale, lz, gz);
    }
    dx[i * 3 + 0] = gx;
    dx[i * 3 + 1] = gy;
    dx[i * 3 + 2] = gz;
  }
}

// ---- double backward of bwd_input: given ddx = dL/d(dx) --------------------------------------------
//   grad_dy_lf    = scale_l * sum_c (ddx . dw_c/dfrac) * table[c][f]
//   grad_table[c][f] += dy_lf * scale_l * (ddx . dw_c/dfrac)
__global__ void __launch_bounds__(kThreads) hashgrid_bwd_bwd_kernel(nsr_grid_t g, const float* __restrict__ x,
                                                                    const __half2* __restrict__ table,
                                                                    const float* __restrict__ dy,
                                                                    const float* __restrict__ ddx,
                                                                    float* __restrict__ grad_table,
                                                                    float* __restrict__ grad_dy, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float px = x[i * 3 + 0], py = x[i * 3 + 1], pz = x[i * 3 + 2];
    const float vx = ddx[i * 3 + 0], vy = ddx[i * 3 + 1], vz = ddx[i * 3 + 2];
    const float2* d = reinterpret_cast<const float2*>(dy) + i * g.n_levels;
    float2* gd = grad_dy ? reinterpret_cast<float2*>(grad_dy) + i * g.n_levels : nullptr;
#pragma unroll 2
    for (int l = 0; l < g.n_levels; ++l) {
      const LevelInfo li = nsr_level(g, l);
      const float2 dv = d[l];
      uint32_t cx, cy, cz, idx[8];
      float fx, fy, fz;
      nsr_pos_fract(px, li.scale, cx, fx);
      nsr_pos_fract(py, li.scale, cy, fy);
      nsr_pos_fract(pz, li.scale, cz, fz);
      nsr_corner_indices(li, cx, cy, cz, idx);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float dw = li.scale * (vx * nsr_corner_dweight(c, 0, fx, fy, fz) + vy * nsr_corner_dweight(c, 1, fx, fy, fz) +
                                     vz * nsr_corner_dweight(c, 2, fx, fy, fz));
        if (gd) {
          const float2 v = nsr_ld_table(table, idx[c]);
          a0 = fmaf(dw, v.x, a0);
          a1 = fmaf(dw, v.y, a1);
        }
        if (grad_table) nsr_red_add_f32x2(grad_table + 2 * (size_t)idx[c], dw * dv.x, dw * dv.y);
      }
      if (gd) gd[l] = make_float2(a0, a1);
    }
  }
}

// ---- SH degree 4 -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) sh4_kernel(const float* __restrict__ v, __half* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    float s[16];
    nsr_sh4(v[i * 3 + 0] * 2.f - 1.f, v[i * 3 + 1] * 2.f - 1.f, v[i * 3 + 2] * 2.f - 1.f, s);
    uint4* o = reinterpret_cast<uint4*>(out + i * 16);
    uint4 a, b;
    a.x = nsr_pack_h2(s[0], s[1]); a.y = nsr_pack_h2(s[2], s[3]); a.z = nsr_pack_h2(s[4], s[5]); a.w = nsr_pack_h2(s[6], s[7]);
    b.x = nsr_pack_h2(s[8], s[9]); b.y = nsr_pack_h2(s[10], s[11]); b.z = nsr_pack_h2(s[12], s[13]); b.w = nsr_pack_h2(s[14], s[15]);
    o[0] = a;
    o[1] = b;
  }
}

int check_grid(const nsr_grid_t* g) {
  NSR_REQUIRE(g != nullptr, "grid descriptor is NULL");
  NSR_REQUIRE(g->n_features == 2, "only n_features_per_level == 2 is implemented (got %d)", g->n_features);
  NSR_REQUIRE(g->n_levels >= 1 && g->n_levels <= NSR_MAX_LEVELS, "n_levels out of range: %d", g->n_levels);
  for (int l = 0; l < g->n_levels; ++l)
    if (!((g->dense_mask >> l) & 1u))
      NSR_REQUIRE((g->size[l] & (g->size[l] - 1)) == 0, "hashed level %d: size %u is not a power of two", l, g->size[l]);
  return 0;
}

int grid_dim(int64_t n) {
  int64_t b = (n + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)nsr_sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" int nsr_hashgrid_fwd(const nsr_grid_t* g, const float* x, const void* table_h, void* out_h, int64_t n, void* stream) {
  if (int e = check_grid(g)) return e;
  if (n == 0) return 0;
  hashgrid_fwd_kernel<<<grid_dim(n), kThreads, 0, (cudaStream_t)stream>>>(*g, x, (const __half2*)table_h, (__half2*)out_h, n);
  NSR_CHECK_LAUNCH("nsr_hashgrid_fwd");
  return 0;
}

extern "C" int nsr_hashgrid_bwd(const nsr_grid_t* g, const float* x, const void* dy_h, float* grad_table, float dy_scale, int64_t n, void* stream) {
  if (int e = check_grid(g)) return e;
  if (n == 0) return 0;
  hashgrid_bwd_kernel<<<grid_dim(n), kThreads, 0, (cudaStream_t)stream>>>(*g, x, (const __half2*)dy_h, grad_table, dy_scale, n);
  NSR_CHECK_LAUNCH("nsr_hashgrid_bwd");
  return 0;
}

extern "C" int nsr_hashgrid_bwd_input(const nsr_grid_t* g, const float* x, const void* table_h, const float* dy, float* dx,
                                      int64_t n, void* stream) {
  if (int e = check_grid(g)) return e;
  if (n == 0) return 0;
  hashgrid_bwd_input_kernel<<<grid_dim(n), kThreads, 0, (cudaStream_t)stream>>>(*g, x, (const __half2*)table_h, dy, dx, n);
  NSR_CHECK_LAUNCH("nsr_hashgrid_bwd_input");
  return 0;
}

extern "C" int nsr_hashgrid_bwd_bwd(const nsr_grid_t* g, const float* x, const void* table_h, const float* dy, const float* ddx,
                                    float* grad_table, float* grad_dy, int64_t n, void* stream) {
  if (int e = check_grid(g)) return e;
  if (n == 0) return 0;
  hashgrid_bwd_bwd_kernel<<<grid_dim(n), kThreads, 0, (cudaStream_t)stream>>>(*g, x, (const __half2*)table_h, dy, ddx, grad_table,
                                                                              grad_dy, n);
  NSR_CHECK_LAUNCH("nsr_hashgrid_bwd_bwd");
  return 0;
}

extern "C" int nsr_sh4_fwd(const float* v, void* out_h, int64_t n, void* stream) {
  if (n == 0) return 0;
  sh4_kernel<<<grid_dim(n), kThreads, 0, (cudaStream_t)stream>>>(v, (__half*)out_h, n);
  NSR_CHECK_LAUNCH("nsr_sh4_fwd");
  return 0;
}
