// a9 -- occupancy grid update / decay (reference: OccupancyGrid.update, src/core.py:134-145).
//
// The reference loops over the 128 depth slices on the host: CPU jitter -> H2D copy -> sigma_fn ->
// where().  Here the update is two elementwise kernels around one batched sigma_fn call over any
// range of cells (the whole 128^3 grid in one go, or a depth-slice shard per rank):
//   coords: cell -> jittered [-1,1] position, including the reference's (x/D, y/H, z/W) quirk
//   apply : alpha = 1 - exp(-sigma*step); grid = alpha > thr ? 1 : decay*grid
// Every arithmetic step is rounded where the reference's un-fused ops round it.
#include "common.cuh"
#include "nerf_math.cuh"

namespace tnf {
namespace {

__global__ void occ_coords_kernel(int gd, int gh, int gw, long long cell0, long long n,
                                  const float* __restrict__ noise, unsigned long long seed,
                                  unsigned long long offset, float* __restrict__ coords) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long cell = cell0 + i;
  const int x = (int)(cell % gw), y = (int)((cell / gw) % gh), z = (int)(cell / ((long long)gw * gh));
  float u[3];
  if (noise) {
    u[0] = __ldg(noise + 3 * i); u[1] = __ldg(noise + 3 * i + 1); u[2] = __ldg(noise + 3 * i + 2);
  } else {
    const Philox ph(seed);
    const uint4 r = ph(offset + (unsigned long long)cell, 1ull);
    u[0] = u01(r.x); u[1] = u01(r.y); u[2] = u01(r.z);
  }
  // self.coords[z,y,x] = (x,y,z) (flip, src/core.py:119); size = (D,H,W) divides (x,y,z) (src/core.py:109,137)
  const float c[3] = {(float)x, (float)y, (float)z};
  const float sz[3] = {(float)gd, (float)gh, (float)gw};
#pragma unroll
  for (int k = 0; k < 3; ++k)
    coords[3 * i + k] = TNF_ADD(-1.f, TNF_DIV(TNF_MUL(2.f, TNF_ADD(c[k], u[k])), sz[k]));
}

__global__ void occ_apply_kernel(float* __restrict__ grid, long long cell0, long long n,
                                 const float* __restrict__ sigma, float step, float thr,
                                 const float* __restrict__ thr_dev, float decay) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (thr_dev) thr = __ldg(thr_dev);
  const float alpha = TNF_SUB(1.f, expf(TNF_MUL(-__ldg(sigma + i), step)));
  const float old = grid[cell0 + i];
  grid[cell0 + i] = (alpha > thr) ? 1.f : TNF_MUL(decay, old);
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_occ_update_coords(int32_t gd, int32_t gh, int32_t gw, int64_t cell0, int64_t n_cells,
                                     const float* noise, uint64_t seed, uint64_t offset, float* coords,
                                     void* stream) {
  using namespace tnf;
  TNF_REQUIRE(gd > 0 && gh > 0 && gw > 0, "bad grid size");
  TNF_REQUIRE(cell0 >= 0 && n_cells >= 0 && cell0 + n_cells <= (int64_t)gd * gh * gw, "cell range out of grid");
  if (n_cells == 0) return TNF_OK;
  TNF_REQUIRE(coords, "null coords");
  occ_coords_kernel<<<(unsigned)ceil_div(n_cells, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gd, gh, gw, cell0, n_cells, noise, seed, offset, coords);
  TNF_LAUNCH_CHECK("occ_coords_kernel");
  return TNF_OK;
}

extern "C" int tnf_occ_update_apply(float* grid, int64_t cell0, int64_t n_cells, const float* sigma,
                                    float step_size, float threshold, float decay, void* stream) {
  return tnf_occ_update_apply_dev(grid, cell0, n_cells, sigma, step_size, threshold, nullptr, decay, stream);
}

extern "C" int tnf_occ_update_apply_dev(float* grid, int64_t cell0, int64_t n_cells, const float* sigma, float step_size,
                                        float threshold, const float* threshold_dev, float decay, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(cell0 >= 0 && n_cells >= 0, "bad cell range");
  if (n_cells == 0) return TNF_OK;
  TNF_REQUIRE(grid && sigma, "null pointer");
  occ_apply_kernel<<<(unsigned)ceil_div(n_cells, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grid, cell0, n_cells, sigma, step_size, threshold, threshold_dev, decay);
  TNF_LAUNCH_CHECK("occ_apply_kernel");
  return TNF_OK;
}
