// Occupancy-grid refresh (nerfacc 0.3.3 OccupancyGrid._update behind every_n_step; models/nerf.py:45-55, models/neus.py:79-111;
// SURVEY 8f-1 / Appendix A.3):
//   x = contract_inv((cell + jitter) / R)            -> nsr_occgrid_points     (cells, jitter drawn by the caller)
//   occs[cell] = max(occs[cell] * decay, occ(x))     -> nsr_occgrid_update     (+ deterministic partial sums for the mean)
//   binary = occs > min(mean(occs), occ_thre)        -> nsr_occgrid_binarize   (bool grid + the packed bitfield + the coarse field
//                                                                               the marcher reads, no separate packing pass)
// Duplicate cells in one update (uniform + occupied samples overlap) resolve to the MAX of their new values; the reference's
// indexed assignment keeps an arbitrary one of them (nondeterministic on CUDA).
#include "common.cuh"

namespace {

constexpr int kPartials = 1024;

__global__ void __launch_bounds__(256) occgrid_points_kernel(const __grid_constant__ nsr_march_t P, const int64_t* __restrict__ cells,
                                                             const float* __restrict__ jitter, float* __restrict__ x_world,
                                                             uint8_t* __restrict__ valid, int64_t n) {
  const int64_t i = blockIdx.x * 256ll + threadIdx.x;
  if (i >= n) return;
  const int64_t cell = cells ? cells[i] : i;
  const int R = P.res;
  const int ix = (int)(cell / ((int64_t)R * R)), iy = (int)((cell / R) % R), iz = (int)(cell % R);
  float u[3] = {((float)ix + jitter[i * 3]) / (float)R, ((float)iy + jitter[i * 3 + 1]) / (float)R, ((float)iz + jitter[i * 3 + 2]) / (float)R};
  bool ok = true;
  if (P.contraction == 2) {  // UN_BOUNDED_SPHERE: only the inscribed ball of the unit cube maps to space
    const float a = u[0] - 0.5f, b = u[1] - 0.5f, c = u[2] - 0.5f;
    ok = sqrtf(a * a + b * b + c * c) < 0.5f;
    float v[3] = {a * 4.f, b * 4.f, c * 4.f};
    const float mag = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const float s = mag > 1.f ? 1.f / (2.f - mag) / mag : 1.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = (v[d] * s + 1.f) / 2.f;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) x_world[i * 3 + d] = u[d] * (P.roi[3 + d] - P.roi[d]) + P.roi[d];
  if (valid) valid[i] = ok ? 1 : 0;
}

// scratch[cell] = max over this update's samples of occ (scratch pre-filled with -1; occ >= 0 so float order == int order)
__global__ void __launch_bounds__(256) occgrid_scatter_max_kernel(const int64_t* __restrict__ cells, const float* __restrict__ occ,
                                                                  float* __restrict__ scratch, int64_t n) {
  const int64_t i = blockIdx.x * 256ll + threadIdx.x;
  if (i >= n) return;
  const float v = fmaxf(occ[i], 0.f);
  atomicMax(reinterpret_cast<int*>(scratch) + cells[i], __float_as_int(v));
}

// occs = max(occs * decay, new) on the touched cells; partial[b] = sum of occs over block b's cells (fixed order => deterministic mean)
__global__ void __launch_bounds__(256) occgrid_ema_kernel(float* __restrict__ occs, const float* __restrict__ fresh, bool fresh_is_scratch,
                                                          float decay, double* __restrict__ partial, int64_t n_cells) {
  double s = 0.0;
  const int64_t per = (n_cells + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * per, hi = min(n_cells, lo + per);
  for (int64_t c = lo + threadIdx.x; c < hi; c += 256) {
    float o = occs[c];
    const float f = fresh[c];
    if (!fresh_is_scratch || f >= 0.f) {
      o = fmaxf(o * decay, f);
      occs[c] = o;
    }
    s += (double)o;
  }
  __shared__ double ws[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += ws[w];
    partial[blockIdx.x] = t;
  }
}

// one thread per cell; a warp's ballot is one word of the bitfield (32 consecutive cells along z)
__global__ void __launch_bounds__(256) occgrid_binarize_kernel(const float* __restrict__ occs, const double* __restrict__ partial, int n_partial,
                                                               float occ_thre, uint8_t* __restrict__ binary, uint32_t* __restrict__ bits,
                                                               uint32_t* __restrict__ coarse, int R, int64_t n_cells) {
  __shared__ float thr_s;
  __shared__ double red[8];
  {  // mean(occs) from the block partials: same fixed-shape tree in every block => every block derives the same threshold
    double t = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 256) t += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += red[w];
      thr_s = fminf((float)(tot / (double)n_cells), occ_thre);
    }
    __syncthreads();
  }
  const float thr = thr_s;
  const int64_t c = blockIdx.x * 256ll + threadIdx.x;
  const bool on = c < n_cells && occs[c] > thr;
  if (c < n_cells && binary) binary[c] = on ? 1 : 0;
  const uint32_t word = __ballot_sync(0xffffffffu, on);
  const int lane = threadIdx.x & 31;
  if (lane == 0 && (c >> 5) < ((n_cells + 31) >> 5)) bits[c >> 5] = word;
  if (coarse != nullptr && on) {
    // "any bit in the 4^3 block": one atomicOr per group of 4 z-neighbours (the lowest set lane of the group issues it)
    const uint32_t grp = (word >> (lane & ~3)) & 0xFu;
    if ((lane & 3) == __ffs(grp) - 1) {
      const int Rc = R >> 2;
      const int ix = (int)(c / ((int64_t)R * R)), iy = (int)((c / R) % R), iz = (int)(c % R);
      const int64_t ci = ((int64_t)(ix >> 2) * Rc + (iy >> 2)) * Rc + (iz >> 2);
      atomicOr(coarse + (ci >> 5), 1u << (ci & 31));
    }
  }
}

}  // namespace

extern "C" int nsr_occgrid_points(const nsr_march_t* p, const int64_t* cells, const float* jitter, float* x_world, uint8_t* valid, int64_t n,
                                  void* stream) {
  NSR_REQUIRE(p != nullptr && p->res >= 1, "nsr_occgrid_points: grid descriptor is NULL / empty");
  NSR_REQUIRE(p->contraction == 0 || p->contraction == 2, "nsr_occgrid_points: contraction must be AABB (0) or UN_BOUNDED_SPHERE (2)");
  if (n == 0) return 0;
  occgrid_points_kernel<<<nsr_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(*p, cells, jitter, x_world, valid, n);
  NSR_CHECK_LAUNCH("nsr_occgrid_points");
  return 0;
}

extern "C" int nsr_occgrid_update(float* occs, const int64_t* cells, const float* occ, float* scratch, float ema_decay, double* partial,
                                  int64_t n, int64_t n_cells, void* stream) {
  NSR_REQUIRE(occs != nullptr && partial != nullptr && n_cells > 0, "nsr_occgrid_update: occs / partial is NULL or the grid is empty");
  NSR_REQUIRE(cells != nullptr || n == n_cells, "nsr_occgrid_update: cells == NULL means one sample per cell (n == n_cells)");
  NSR_REQUIRE(cells == nullptr || scratch != nullptr, "nsr_occgrid_update: sparse updates need the scratch grid");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)min((int64_t)kPartials, (n_cells + 255) / 256);
  if (cells == nullptr) {
    occgrid_ema_kernel<<<grid, 256, 0, st>>>(occs, occ, false, ema_decay, partial, n_cells);
  } else {
    cudaMemsetAsync(scratch, 0xBF, (size_t)n_cells * sizeof(float), st);  // 0xBFBFBFBF = -1.498: "untouched"
    if (n > 0) occgrid_scatter_max_kernel<<<nsr_blocks(n, 256), 256, 0, st>>>(cells, occ, scratch, n);
    occgrid_ema_kernel<<<grid, 256, 0, st>>>(occs, scratch, true, ema_decay, partial, n_cells);
  }
  NSR_CHECK_LAUNCH("nsr_occgrid_update");
  return 0;
}

extern "C" int nsr_occgrid_binarize(const float* occs, const double* partial, float occ_thre, uint8_t* binary, uint32_t* bits,
                                    uint32_t* coarse_bits, int32_t res, int64_t n_cells, void* stream) {
  NSR_REQUIRE(occs != nullptr && partial != nullptr && bits != nullptr, "nsr_occgrid_binarize: occs / partial / bits is NULL");
  NSR_REQUIRE(n_cells == (int64_t)res * res * res, "nsr_occgrid_binarize: n_cells must be res^3");
  NSR_REQUIRE(coarse_bits == nullptr || (res % 4 == 0), "nsr_occgrid_binarize: the coarse field needs res % 4 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (coarse_bits != nullptr) {
    const int64_t nc = (int64_t)(res / 4) * (res / 4) * (res / 4);
    cudaMemsetAsync(coarse_bits, 0, (size_t)((nc + 31) / 32) * sizeof(uint32_t), st);
  }
  const int n_partial = (int)min((int64_t)kPartials, (n_cells + 255) / 256);
  const int64_t padded = (n_cells + 31) / 32 * 32;
  occgrid_binarize_kernel<<<nsr_blocks(padded, 256), 256, 0, st>>>(occs, partial, n_partial, occ_thre, binary, bits, coarse_bits, res, n_cells);
  NSR_CHECK_LAUNCH("nsr_occgrid_binarize");
  return 0;
}
