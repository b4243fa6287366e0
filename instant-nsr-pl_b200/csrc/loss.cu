// Fused per-ray training loss of systems/nerf.py:68-97: background blend + masked smooth-L1 (beta = 1) over the valid rays,
// forward and backward in one small kernel each (replaces ~30 elementwise / reduction torch kernels per step).
//   comp = acc_rgb + bg * (1 - opacity);  valid = opacity > 0
//   loss = sum_{valid rays, 3 channels} smooth_l1(comp - target) / max(3 * n_valid, 1)
#include "common.cuh"

namespace {

__device__ __forceinline__ float sl1(float d) {
  const float a = fabsf(d);
  return a < 1.f ? 0.5f * d * d : a - 0.5f;
}
__device__ __forceinline__ float sl1_grad(float d) { return fminf(fmaxf(d, -1.f), 1.f); }

// accum[0] = sum of losses, accum[1] = number of valid rays (both zeroed by the caller)
__global__ void __launch_bounds__(256) nerf_loss_fwd_kernel(const float* __restrict__ acc_rgb, const float* __restrict__ opacity,
                                                            const float* __restrict__ bg, const float* __restrict__ target,
                                                            float* __restrict__ comp_rgb, float* __restrict__ accum, int64_t n) {
  float s = 0.f, cnt = 0.f;
  const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float op = opacity[i], k = 1.f - op;
    const float c0 = acc_rgb[i * 3] + b0 * k, c1 = acc_rgb[i * 3 + 1] + b1 * k, c2 = acc_rgb[i * 3 + 2] + b2 * k;
    if (comp_rgb) {
      comp_rgb[i * 3] = c0;
      comp_rgb[i * 3 + 1] = c1;
      comp_rgb[i * 3 + 2] = c2;
    }
    if (op > 0.f) {
      s += sl1(c0 - target[i * 3]) + sl1(c1 - target[i * 3 + 1]) + sl1(c2 - target[i * 3 + 2]);
      cnt += 1.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  __shared__ float ws[8], wc[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    ws[warp] = s;
    wc[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tc = 0.f;
    for (int w = 0; w < 8; ++w) {
      ts += ws[w];
      tc += wc[w];
    }
    if (ts != 0.f) atomicAdd(accum, ts);
    if (tc != 0.f) atomicAdd(accum + 1, tc);
  }
}

// g_acc_rgb[i] = valid * sl1'(comp - target) * g / (3 n_valid);  g_opacity[i] = -sum_c g_acc_rgb[i,c] * bg[c]
__global__ void __launch_bounds__(256) nerf_loss_bwd_kernel(const float* __restrict__ acc_rgb, const float* __restrict__ opacity,
                                                            const float* __restrict__ bg, const float* __restrict__ target,
                                                            const float* __restrict__ accum, const float* __restrict__ g_loss,
                                                            float* __restrict__ g_acc_rgb, float* __restrict__ g_opacity, int64_t n) {
  const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
  const float scale = (g_loss ? g_loss[0] : 1.f) / fmaxf(3.f * accum[1], 1.f);
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float op = opacity[i], k = 1.f - op;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (op > 0.f) {
      g0 = sl1_grad(acc_rgb[i * 3] + b0 * k - target[i * 3]) * scale;
      g1 = sl1_grad(acc_rgb[i * 3 + 1] + b1 * k - target[i * 3 + 1]) * scale;
      g2 = sl1_grad(acc_rgb[i * 3 + 2] + b2 * k - target[i * 3 + 2]) * scale;
    }
    g_acc_rgb[i * 3] = g0;
    g_acc_rgb[i * 3 + 1] = g1;
    g_acc_rgb[i * 3 + 2] = g2;
    g_opacity[i] = -(g0 * b0 + g1 * b1 + g2 * b2);
  }
}

}  // namespace

namespace {
// accum[2] = loss = sum / max(3 n_valid, 1)
__global__ void nerf_loss_finalize_kernel(float* __restrict__ accum) {
  if (threadIdx.x == 0 && blockIdx.x == 0) accum[2] = accum[0] / fmaxf(accum[1] * 3.f, 1.f);
}
}  // namespace

extern "C" int nsr_nerf_loss_fwd(const float* acc_rgb, const float* opacity, const float* bg3, const float* target, float* comp_rgb,
                                 float* accum4, int64_t n_rays, void* stream) {
  NSR_REQUIRE(accum4 != nullptr, "nsr_nerf_loss_fwd: accum is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(accum4, 0, 4 * sizeof(float), st);
  if (n_rays > 0) {
    const int grid = (int)min((int64_t)nsr_sm_count(), (n_rays + 255) / 256);
    nerf_loss_fwd_kernel<<<grid, 256, 0, st>>>(acc_rgb, opacity, bg3, target, comp_rgb, accum4, n_rays);
  }
  nerf_loss_finalize_kernel<<<1, 32, 0, st>>>(accum4);
  NSR_CHECK_LAUNCH("nsr_nerf_loss_fwd");
  return 0;
}

extern "C" int nsr_nerf_loss_bwd(const float* acc_rgb, const float* opacity, const float* bg3, const float* target, const float* accum2,
                                 const float* g_loss, float* g_acc_rgb, float* g_opacity, int64_t n_rays, void* stream) {
  if (n_rays == 0) return 0;
  const int grid = (int)min((int64_t)nsr_sm_count(), (n_rays + 255) / 256);
  nerf_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(acc_rgb, opacity, bg3, target, accum2, g_loss, g_acc_rgb, g_opacity, n_rays);
  NSR_CHECK_LAUNCH("nsr_nerf_loss_bwd");
  return 0;
}

// ---- NeuS training losses (systems/neus.py:98-121) as one reduction kernel + one gradient kernel instead of ~65 torch kernels:
//   rgb mse / l1 over the valid rays, eikonal ((|grad sdf| - 1)^2 mean over samples), mask BCE and "opaque" BCE on the clamped opacity
//   (systems/criterions.py:155-159), sparsity exp(-scale |sdf|) mean.  accum (device float[8], zeroed by the entry point):
//   [0] sum (c-t)^2  [1] sum |c-t|  [2] n_valid  [3] mask BCE sum  [4] opaque BCE sum  [5] eikonal sum  [6] sparsity sum
namespace {

__device__ __forceinline__ void block_accumulate(float (&v)[7], float* __restrict__ accum) {
  __shared__ float ws[8][7];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    if (lane == 0) ws[warp][j] = v[j];
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += ws[w][threadIdx.x];
    if (t != 0.f) atomicAdd(accum + threadIdx.x, t);
  }
}

__global__ void __launch_bounds__(256) neus_loss_fwd_kernel(const float* __restrict__ comp_rgb, const uint8_t* __restrict__ valid,
                                                            const float* __restrict__ target, const float* __restrict__ opacity,
                                                            const float* __restrict__ fg_mask, const float* __restrict__ sdf_grad,
                                                            const float* __restrict__ sdf, float sparsity_scale, float* __restrict__ accum,
                                                            int64_t n_rays, int64_t k_cap, const int64_t* __restrict__ k_dev) {
  const int64_t k = k_dev ? min(*k_dev, k_cap) : k_cap;
  float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n_rays; i += stride) {
    if (valid[i]) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = comp_rgb[i * 3 + c] - target[i * 3 + c];
        v[0] += d * d;
        v[1] += fabsf(d);
      }
      v[2] += 1.f;
    }
    const float o = fminf(fmaxf(opacity[i], 1e-3f), 1.f - 1e-3f);
    const float lo = logf(o), l1o = logf(1.f - o);
    if (fg_mask) {
      const float m = fg_mask[i];
      v[3] -= m * lo + (1.f - m) * l1o;
    }
    v[4] -= o * lo + (1.f - o) * l1o;
  }
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < k; i += stride) {
    if (sdf_grad) {
      const float gx = sdf_grad[i * 3], gy = sdf_grad[i * 3 + 1], gz = sdf_grad[i * 3 + 2];
      const float e = sqrtf(gx * gx + gy * gy + gz * gz) - 1.f;
      v[5] += e * e;
    }
    if (sdf) v[6] += expf(-sparsity_scale * fabsf(sdf[i]));
  }
  block_accumulate(v, accum);
}

// losses[0..5] = rgb_mse, rgb_l1, eikonal, mask, opaque, sparsity;  losses[6] = lambda-weighted total
__global__ void neus_loss_finalize_kernel(const nsr_neus_loss_t P, const float* __restrict__ accum, float* __restrict__ losses, int64_t n_rays,
                                          int64_t k_cap, const int64_t* __restrict__ k_dev) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int64_t k = k_dev ? min(*k_dev, k_cap) : k_cap;
  const float nv3 = fmaxf(accum[2] * 3.f, 1.f), nr = fmaxf((float)n_rays, 1.f), nk = fmaxf((float)k, 1.f);
  const float l[6] = {accum[0] / nv3, accum[1] / nv3, accum[5] / nk, accum[3] / nr, accum[4] / nr, accum[6] / nk};
  const float lam[6] = {P.lambda_rgb_mse, P.lambda_rgb_l1, P.lambda_eikonal, P.lambda_mask, P.lambda_opaque, P.lambda_sparsity};
  float tot = 0.f;
  for (int j = 0; j < 6; ++j) {
    losses[j] = l[j];
    if (lam[j] != 0.f) tot += lam[j] * l[j];
  }
  losses[6] = tot;
}

__global__ void __launch_bounds__(256) neus_loss_bwd_kernel(const nsr_neus_loss_t P, const float* __restrict__ comp_rgb,
                                                            const uint8_t* __restrict__ valid, const float* __restrict__ target,
                                                            const float* __restrict__ opacity, const float* __restrict__ fg_mask,
                                                            const float* __restrict__ sdf_grad, const float* __restrict__ sdf,
                                                            const float* __restrict__ accum, const float* __restrict__ g_loss,
                                                            float* __restrict__ g_comp_rgb, float* __restrict__ g_opacity,
                                                            float* __restrict__ g_sdf_grad, float* __restrict__ g_sdf, int64_t n_rays,
                                                            int64_t k_cap, const int64_t* __restrict__ k_dev) {
  const int64_t k = k_dev ? min(*k_dev, k_cap) : k_cap;
  const float g = g_loss ? __ldg(g_loss) : 1.f;
  const float inv_nv3 = g / fmaxf(accum[2] * 3.f, 1.f), inv_nr = g / fmaxf((float)n_rays, 1.f), inv_nk = g / fmaxf((float)k, 1.f);
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n_rays; i += stride) {
    const bool ok = valid[i] != 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float d = 0.f;
      if (ok) {
        const float e = comp_rgb[i * 3 + c] - target[i * 3 + c];
        d = (P.lambda_rgb_mse * 2.f * e + P.lambda_rgb_l1 * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f))) * inv_nv3;
      }
      g_comp_rgb[i * 3 + c] = d;
    }
    const float op = opacity[i];
    float go = 0.f;
    if (op >= 1e-3f && op <= 1.f - 1e-3f) {  // clamp passes the gradient on the closed interval
      if (fg_mask && P.lambda_mask != 0.f) {
        const float m = fg_mask[i];
        go -= P.lambda_mask * (m / op - (1.f - m) / (1.f - op));
      }
      if (P.lambda_opaque != 0.f) go -= P.lambda_opaque * (logf(op) - logf(1.f - op));
    }
    g_opacity[i] = go * inv_nr;
  }
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < k; i += stride) {
    if (g_sdf_grad) {
      const float gx = sdf_grad[i * 3], gy = sdf_grad[i * 3 + 1], gz = sdf_grad[i * 3 + 2];
      const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
      const float s = nrm > 0.f ? P.lambda_eikonal * 2.f * (nrm - 1.f) / nrm * inv_nk : 0.f;
      g_sdf_grad[i * 3] = s * gx;
      g_sdf_grad[i * 3 + 1] = s * gy;
      g_sdf_grad[i * 3 + 2] = s * gz;
    }
    if (g_sdf) {
      const float x = sdf[i];
      const float sg = x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
      g_sdf[i] = P.lambda_sparsity != 0.f ? -P.lambda_sparsity * P.sparsity_scale * sg * expf(-P.sparsity_scale * fabsf(x)) * inv_nk : 0.f;
    }
  }
}

}  // namespace

extern "C" int nsr_neus_loss_fwd(const nsr_neus_loss_t* p, const float* comp_rgb, const uint8_t* valid, const float* target,
                                 const float* opacity, const float* fg_mask, const float* sdf_grad, const float* sdf, float* accum8,
                                 float* losses7, int64_t n_rays, int64_t k, const int64_t* k_dev, void* stream) {
  NSR_REQUIRE(p != nullptr && accum8 != nullptr && losses7 != nullptr, "nsr_neus_loss_fwd: descriptor / accum / losses is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(accum8, 0, 8 * sizeof(float), st);
  const int64_t nmax = max(n_rays, k);
  if (nmax > 0) {
    const int grid = (int)min((int64_t)nsr_sm_count() * 4, (nmax + 255) / 256);
    neus_loss_fwd_kernel<<<grid, 256, 0, st>>>(comp_rgb, valid, target, opacity, fg_mask, sdf_grad, sdf, p->sparsity_scale, accum8, n_rays, k, k_dev);
  }
  neus_loss_finalize_kernel<<<1, 32, 0, st>>>(*p, accum8, losses7, n_rays, k, k_dev);
  NSR_CHECK_LAUNCH("nsr_neus_loss_fwd");
  return 0;
}

extern "C" int nsr_neus_loss_bwd(const nsr_neus_loss_t* p, const float* comp_rgb, const uint8_t* valid, const float* target,
                                 const float* opacity, const float* fg_mask, const float* sdf_grad, const float* sdf, const float* accum8,
                                 const float* g_loss, float* g_comp_rgb, float* g_opacity, float* g_sdf_grad, float* g_sdf, int64_t n_rays,
                                 int64_t k, const int64_t* k_dev, void* stream) {
  NSR_REQUIRE(p != nullptr && accum8 != nullptr, "nsr_neus_loss_bwd: descriptor / accum is NULL");
  NSR_REQUIRE(g_comp_rgb != nullptr && g_opacity != nullptr, "nsr_neus_loss_bwd: per-ray gradient outputs are NULL");
  const int64_t nmax = max(n_rays, k);
  if (nmax == 0) return 0;
  const int grid = (int)min((int64_t)nsr_sm_count() * 4, (nmax + 255) / 256);
  neus_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, comp_rgb, valid, target, opacity, fg_mask, sdf_grad, sdf, accum8, g_loss,
                                                                g_comp_rgb, g_opacity, g_sdf_grad, g_sdf, n_rays, k, k_dev);
  NSR_CHECK_LAUNCH("nsr_neus_loss_bwd");
  return 0;
}
