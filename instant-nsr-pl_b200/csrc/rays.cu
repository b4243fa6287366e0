// Training-batch front end: pixel -> ray in one kernel.  Replaces preprocess_data of the reference's systems
// (systems/nerf.py:33-76, systems/neus.py:34-84): all_c2w[index], directions[y, x] (or [index, y, x]), get_rays
// (models/ray_utils.py:23-43: rays_d = sum_j directions_j * R_ij, rays_o = c2w[:, 3]), F.normalize, torch.cat, the colour / mask
// gathers all_images[index, y, x], all_fg_masks[index, y, x] and the apply_mask blend -- about a dozen torch kernels and their
// intermediates -- with one thread per ray.  The random draws (index, x, y) stay with the caller's torch generator.
#include "common.cuh"

namespace {

struct RayArgs {
  const float* directions;  // [H, W, 3] or [n_images, H, W, 3]
  const float* c2w;         // [n_images, c2w_rows, 4]
  const float* images;      // [n_images, H, W, channels] or NULL
  const float* masks;       // [n_images, H, W] or NULL
  const int64_t* index;     // [n] or NULL: every ray from image fixed_index, pixel r -> (x = r % W, y = r / W)  (eval: whole image)
  const int64_t* x;
  const int64_t* y;
  const float* bg;          // [3] or NULL
  float* rays;              // [n, 6]
  float* rgb;               // [n, 3] or NULL
  float* fg;                // [n] or NULL
  int64_t n;
  int32_t H, W, n_images, channels, c2w_rows, dirs_per_image, apply_mask, fixed_index;
};

__global__ void __launch_bounds__(256) gather_rays_kernel(const RayArgs a) {
  const int64_t r = blockIdx.x * 256ll + threadIdx.x;
  if (r >= a.n) return;
  int64_t img, px, py;
  if (a.index != nullptr) {
    img = a.index[r], px = a.x[r], py = a.y[r];
  } else {
    img = a.fixed_index, px = r % a.W, py = r / a.W;
  }
  // the caller draws in range (torch.randint); clamp anyway so a bad index can never read outside the dataset tensors
  img = min(max(img, (int64_t)0), (int64_t)a.n_images - 1);
  px = min(max(px, (int64_t)0), (int64_t)a.W - 1);
  py = min(max(py, (int64_t)0), (int64_t)a.H - 1);
  const int64_t pix = py * a.W + px, ipix = img * a.H * a.W + pix;
  const float* d = a.directions + (a.dirs_per_image ? ipix : pix) * 3;
  const float dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
  const float* M = a.c2w + img * a.c2w_rows * 4;
  float o[3], w[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float* row = M + i * 4;
    w[i] = __fadd_rn(__fadd_rn(__fmul_rn(dx, __ldg(row)), __fmul_rn(dy, __ldg(row + 1))), __fmul_rn(dz, __ldg(row + 2)));  // torch: (d * R).sum(-1)
    o[i] = __ldg(row + 3);
  }
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(w[0], w[0]), __fmul_rn(w[1], w[1])), __fmul_rn(w[2], w[2])));
  const float den = fmaxf(nrm, 1e-12f);  // F.normalize(p=2, eps=1e-12)
  float* out = a.rays + r * 6;
  out[0] = o[0], out[1] = o[1], out[2] = o[2];
  out[3] = __fdiv_rn(w[0], den), out[4] = __fdiv_rn(w[1], den), out[5] = __fdiv_rn(w[2], den);
  float m = 1.f;
  if (a.masks != nullptr) m = __ldg(a.masks + ipix);
  if (a.fg != nullptr) a.fg[r] = m;
  if (a.rgb != nullptr) {
    const float* c = a.images + ipix * a.channels;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = __ldg(c + k);
      if (a.apply_mask) v = __fadd_rn(__fmul_rn(v, m), __fmul_rn(__ldg(a.bg + k), __fsub_rn(1.f, m)));  // rgb * fg + bg * (1 - fg)
      a.rgb[r * 3 + k] = v;
    }
  }
}

}  // namespace

extern "C" int nsr_gather_rays(const float* directions, int32_t dirs_per_image, const float* c2w, int32_t c2w_rows, const float* images,
                               int32_t channels, const float* masks, const int64_t* index, const int64_t* x, const int64_t* y,
                               int32_t fixed_index, const float* bg, int32_t apply_mask, int32_t H, int32_t W, int32_t n_images, float* rays,
                               float* rgb, float* fg, int64_t n, void* stream) {
  if (n == 0) return 0;
  NSR_REQUIRE(directions != nullptr && c2w != nullptr && rays != nullptr, "nsr_gather_rays: directions / c2w / rays is NULL");
  NSR_REQUIRE(H > 0 && W > 0 && n_images > 0, "nsr_gather_rays: empty dataset (H %d, W %d, images %d)", H, W, n_images);
  NSR_REQUIRE(c2w_rows == 3 || c2w_rows == 4, "nsr_gather_rays: c2w must be [n,3,4] or [n,4,4]");
  NSR_REQUIRE(index == nullptr || (x != nullptr && y != nullptr), "nsr_gather_rays: index without x / y");
  NSR_REQUIRE(index != nullptr || (fixed_index >= 0 && fixed_index < n_images), "nsr_gather_rays: image %d out of range", fixed_index);
  NSR_REQUIRE(rgb == nullptr || (images != nullptr && channels >= 3), "nsr_gather_rays: rgb output needs images with >= 3 channels");
  NSR_REQUIRE(!apply_mask || rgb == nullptr || (masks != nullptr && bg != nullptr), "nsr_gather_rays: apply_mask needs masks and bg");
  RayArgs a;
  a.directions = directions, a.c2w = c2w, a.images = images, a.masks = masks, a.index = index, a.x = x, a.y = y, a.bg = bg;
  a.rays = rays, a.rgb = rgb, a.fg = fg, a.n = n;
  a.H = H, a.W = W, a.n_images = n_images, a.channels = channels, a.c2w_rows = c2w_rows, a.dirs_per_image = dirs_per_image;
  a.apply_mask = apply_mask, a.fixed_index = fixed_index;
  gather_rays_kernel<<<nsr_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(a);
  NSR_CHECK_LAUNCH("nsr_gather_rays");
  return 0;
}
