// Stand-alone forms of the reference's public helper callables, so that NO method of the host-side mirror runs
// PyTorch arithmetic (the eager bodies were transcribed reference code kept as fallbacks):
//   RayMarcherAABB.__call__            src/core.py:73-88    -> tnf_marcher_aabb
//   ContractionAABB / Mip360.__call__  src/core.py:16-31    -> tnf_contract
//   KPlanesFeaturePlane.forward        src/models.py:105-113 (one 2-D grid_sample)  -> tnf_plane_lookup_fwd / _bwd
//   CobafaGrid.forward                 src/models.py:228-238 (one 3-D grid_sample)  -> tnf_grid3_lookup_fwd / _bwd
//   KPlanesFeaturePlane.loss_l1        src/models.py:120-121 (mean |plane|)          -> tnf_abs_mean_fwd / _bwd
// None of them is on the training hot path (RayProvider fuses march + contraction + occupancy + packing, the feature
// fields fuse all their planes / grids); they exist for callers that use the pieces on their own.  The arithmetic is
// the same device code the fused kernels use (nerf_math.cuh), so the results are the fused path's bit for bit.
#include "common.cuh"
#include "nerf_math.cuh"

namespace tnf {
namespace {

__global__ void marcher_aabb_kernel(MarchConst M, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                    long long n_rays, float* __restrict__ t_values, float* __restrict__ step_sizes) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_rays * M.n_steps) return;
  const long long ray = i / M.n_steps;
  const int j = (int)(i - ray * M.n_steps);
  float o[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = __ldg(rays_o + ray * 3 + c);
    d[c] = __ldg(rays_d + ray * 3 + c);
  }
  const float tmin = aabb_t_min(o, d, M);
  t_values[i] = TNF_ADD(tmin, TNF_MUL((float)j, M.step_size));   // t_min[:, None] + arange(S) * step (src/core.py:84-85)
  step_sizes[i] = M.step_size;
}

__global__ void contract_kernel(MarchConst M, const float* __restrict__ coords, long long n, float* __restrict__ out,
                                uint8_t* __restrict__ mask) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[3], q[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = __ldg(coords + i * 3 + c);
  const bool inside = contract_point(M, p, q);
#pragma unroll
  for (int c = 0; c < 3; ++c) out[i * 3 + c] = q[c];
  if (mask) mask[i] = inside;
}

// ---- one 2-D plane, channels-last [H][W][C]; grid_sample(bilinear, zeros, align_corners=True) ------------------
struct Ax { int i0; float w0, w1; bool ok0, ok1; };
__device__ __forceinline__ Ax ax_setup(float c, int res) {
  Ax a;
  const float i = TNF_MUL(TNF_MUL(TNF_ADD(c, 1.f), 0.5f), (float)(res - 1));
  const float f = floorf(i);
  a.i0 = (int)f;
  a.w0 = TNF_SUB(f + 1.f, i);
  a.w1 = TNF_SUB(i, f);
  a.ok0 = (unsigned)a.i0 < (unsigned)res;
  a.ok1 = (unsigned)(a.i0 + 1) < (unsigned)res;
  return a;
}

template <bool BWD>
__global__ void plane_lookup_kernel(const float* __restrict__ plane, float* __restrict__ gplane, int H, int W, int C,
                                    const float* __restrict__ xy, long long xy_stride, long long n, float* __restrict__ out,
                                    const float* __restrict__ gout) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * C) return;
  const long long s = t / C;
  const int c = (int)(t - s * C);
  const Ax ax = ax_setup(__ldg(xy + s * xy_stride), W), ay = ax_setup(__ldg(xy + s * xy_stride + 1), H);
  // nw, ne, sw, se in torch's order
  const float w[4] = {TNF_MUL(ax.w0, ay.w0), TNF_MUL(ax.w1, ay.w0), TNF_MUL(ax.w0, ay.w1), TNF_MUL(ax.w1, ay.w1)};
  const bool ok[4] = {ax.ok0 && ay.ok0, ax.ok1 && ay.ok0, ax.ok0 && ay.ok1, ax.ok1 && ay.ok1};
  const long long off[4] = {((long long)ay.i0 * W + ax.i0) * C + c, ((long long)ay.i0 * W + ax.i0 + 1) * C + c,
                            ((long long)(ay.i0 + 1) * W + ax.i0) * C + c, ((long long)(ay.i0 + 1) * W + ax.i0 + 1) * C + c};
  if (!BWD) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) acc = TNF_FMA(__ldg(plane + off[k]), w[k], acc);
    out[t] = acc;
  } else {
    const float g = __ldg(gout + t);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) atomicAdd(gplane + off[k], TNF_MUL(w[k], g));
  }
}

// ---- one 3-D grid, channels-last [D][H][W][C]; x -> W, y -> H, z -> D ----------------------------------------
template <bool BWD>
__global__ void grid3_lookup_kernel(const float* __restrict__ grid, float* __restrict__ ggrid, int D, int H, int W, int C,
                                    const float* __restrict__ x, long long x_stride, long long n, float* __restrict__ out,
                                    const float* __restrict__ gout) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * C) return;
  const long long s = t / C;
  const int c = (int)(t - s * C);
  const Ax ax = ax_setup(__ldg(x + s * x_stride), W), ay = ax_setup(__ldg(x + s * x_stride + 1), H),
           az = ax_setup(__ldg(x + s * x_stride + 2), D);
  float acc = 0.f;
  const float g = BWD ? __ldg(gout + t) : 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {   // tnw,tne,tsw,tse,bnw,bne,bsw,bse
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const float w = TNF_MUL(TNF_MUL(dx ? ax.w1 : ax.w0, dy ? ay.w1 : ay.w0), dz ? az.w1 : az.w0);
    const bool ok = (dx ? ax.ok1 : ax.ok0) && (dy ? ay.ok1 : ay.ok0) && (dz ? az.ok1 : az.ok0);
    if (!ok) continue;
    const long long off = ((((long long)(az.i0 + dz)) * H + (ay.i0 + dy)) * W + (ax.i0 + dx)) * C + c;
    if (!BWD) acc = TNF_FMA(__ldg(grid + off), w, acc);
    else atomicAdd(ggrid + off, TNF_MUL(w, g));
  }
  if (!BWD) out[t] = acc;
}

// ---- PositionalEncoding.forward (src/models.py:30-39): x [n][d] -> [n][d * 2 * n_freqs]; per coordinate the n_freqs sines
// then the n_freqs cosines; frequencies fl32(pi) * 2^k, products rounded to fp32 before sincosf (as tnf_color_input) ----
__global__ void positional_encoding_kernel(const float* __restrict__ x, long long x_stride, int d, int n_freqs, long long n,
                                           float* __restrict__ out, long long ld_out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int per_row = d * n_freqs;
  if (t >= n * per_row) return;
  const long long row = t / per_row;
  const int j = (int)(t - row * per_row), c = j / n_freqs, k = j - c * n_freqs;
  const float arg = __fmul_rn(__ldg(x + row * x_stride + c), ldexpf(3.14159274101257324219f, k));
  float sv, cv;
  sincosf(arg, &sv, &cv);
  float* o = out + row * ld_out + c * 2 * n_freqs;
  o[k] = sv;
  o[n_freqs + k] = cv;
}

// ---- mean |x| ---------------------------------------------------------------------------------------------------
__global__ void abs_sum_kernel(const float* __restrict__ x, long long n, double* __restrict__ sum) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += (double)fabsf(__ldg(x + i));
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(kFullMask, acc, d);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
    atomicAdd(sum, t);
  }
}
__global__ void abs_mean_bwd_kernel(const float* __restrict__ x, long long n, const float* __restrict__ gscale,
                                    float* __restrict__ gx) {
  const float g = __ldg(gscale) / (float)n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i);
    gx[i] = v > 0.f ? g : (v < 0.f ? -g : 0.f);   // sign(x) * g / n, like torch's abs backward (0 at 0)
  }
}

int aabb_const(const float* aabb6, float near_, float far_, float step, int n_steps, MarchConst* M) {
  TNF_REQUIRE(aabb6, "null aabb");
  *M = MarchConst{};
  M->scene = 0;
  M->n_steps = n_steps;
  for (int c = 0; c < 3; ++c) {
    M->a0[c] = aabb6[c];
    M->a1[c] = aabb6[3 + c];
    volatile float e = aabb6[3 + c] - aabb6[c];
    M->ext[c] = e;
  }
  M->near_ = near_; M->far_ = far_; M->step_size = step;
  return TNF_OK;
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_marcher_aabb(const float* aabb6, float near_, float far_, float step_size, const float* rays_o,
                                const float* rays_d, int64_t n_rays, int32_t n_steps, float* t_values, float* step_sizes,
                                void* stream) {
  using namespace tnf;
  MarchConst M;
  int rc = aabb_const(aabb6, near_, far_, step_size, n_steps, &M);
  if (rc != TNF_OK) return rc;
  TNF_REQUIRE(n_rays >= 0 && n_steps > 0, "bad sizes");
  if (n_rays == 0) return TNF_OK;
  TNF_REQUIRE(rays_o && rays_d && t_values && step_sizes, "null pointer");
  marcher_aabb_kernel<<<(unsigned)ceil_div(n_rays * n_steps, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      M, rays_o, rays_d, n_rays, t_values, step_sizes);
  TNF_LAUNCH_CHECK("marcher_aabb_kernel");
  return TNF_OK;
}

extern "C" int tnf_contract(int32_t scene, const float* aabb6, const float* coords, int64_t n, float* out, uint8_t* mask,
                            void* stream) {
  using namespace tnf;
  TNF_REQUIRE(scene == TNF_SCENE_AABB || scene == TNF_SCENE_UNBOUNDED, "bad scene %d", scene);
  MarchConst M{};
  if (scene == TNF_SCENE_AABB) {
    int rc = aabb_const(aabb6, 0.f, 0.f, 0.f, 1, &M);
    if (rc != TNF_OK) return rc;
  }
  M.scene = scene;
  TNF_REQUIRE(n >= 0, "negative n");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(coords && out, "null pointer");
  contract_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(M, coords, n, out, mask);
  TNF_LAUNCH_CHECK("contract_kernel");
  return TNF_OK;
}

extern "C" int tnf_plane_lookup_fwd(const float* plane, int32_t h, int32_t w, int32_t channels, const float* xy,
                                    int64_t xy_stride, int64_t n, float* out, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0 && h >= 1 && w >= 1 && channels >= 1 && xy_stride >= 2, "bad sizes");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(plane && xy && out, "null pointer");
  plane_lookup_kernel<false><<<(unsigned)ceil_div(n * channels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      plane, nullptr, h, w, channels, xy, xy_stride, n, out, nullptr);
  TNF_LAUNCH_CHECK("plane_lookup_kernel");
  return TNF_OK;
}

extern "C" int tnf_plane_lookup_bwd(float* grad_plane, int32_t h, int32_t w, int32_t channels, const float* xy,
                                    int64_t xy_stride, int64_t n, const float* grad_out, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0 && h >= 1 && w >= 1 && channels >= 1 && xy_stride >= 2, "bad sizes");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(grad_plane && xy && grad_out, "null pointer");
  plane_lookup_kernel<true><<<(unsigned)ceil_div(n * channels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      nullptr, grad_plane, h, w, channels, xy, xy_stride, n, nullptr, grad_out);
  TNF_LAUNCH_CHECK("plane_lookup_bwd_kernel");
  return TNF_OK;
}

extern "C" int tnf_grid3_lookup_fwd(const float* grid, int32_t d, int32_t h, int32_t w, int32_t channels, const float* x,
                                    int64_t x_stride, int64_t n, float* out, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0 && d >= 1 && h >= 1 && w >= 1 && channels >= 1 && x_stride >= 3, "bad sizes");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(grid && x && out, "null pointer");
  grid3_lookup_kernel<false><<<(unsigned)ceil_div(n * channels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grid, nullptr, d, h, w, channels, x, x_stride, n, out, nullptr);
  TNF_LAUNCH_CHECK("grid3_lookup_kernel");
  return TNF_OK;
}

extern "C" int tnf_grid3_lookup_bwd(float* grad_grid, int32_t d, int32_t h, int32_t w, int32_t channels, const float* x,
                                    int64_t x_stride, int64_t n, const float* grad_out, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0 && d >= 1 && h >= 1 && w >= 1 && channels >= 1 && x_stride >= 3, "bad sizes");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(grad_grid && x && grad_out, "null pointer");
  grid3_lookup_kernel<true><<<(unsigned)ceil_div(n * channels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      nullptr, grad_grid, d, h, w, channels, x, x_stride, n, nullptr, grad_out);
  TNF_LAUNCH_CHECK("grid3_lookup_bwd_kernel");
  return TNF_OK;
}

extern "C" int tnf_positional_encoding(const float* x, int64_t x_stride, int32_t d, int32_t n_freqs, int64_t n, float* out,
                                       int64_t ld_out, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0 && d >= 1 && n_freqs >= 1 && n_freqs <= 24 && x_stride >= d && ld_out >= 2 * d * n_freqs, "bad sizes");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(x && out, "null pointer");
  positional_encoding_kernel<<<(unsigned)ceil_div(n * d * n_freqs, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_stride, d, n_freqs, n, out, ld_out);
  TNF_LAUNCH_CHECK("positional_encoding_kernel");
  return TNF_OK;
}

extern "C" int tnf_abs_mean_fwd(const float* x, int64_t n, double* sum, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0 && sum, "bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TNF_CUDA(cudaMemsetAsync(sum, 0, sizeof(double), st));
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(x, "null x");
  const int64_t want = ceil_div(n, 256 * 8), cap = (int64_t)sm_count() * 8;
  abs_sum_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(x, n, sum);
  TNF_LAUNCH_CHECK("abs_sum_kernel");
  return TNF_OK;
}

extern "C" int tnf_abs_mean_bwd(const float* x, int64_t n, const float* grad_scale, float* grad_x, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0, "negative n");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(x && grad_scale && grad_x, "null pointer");
  const int64_t want = ceil_div(n, 256 * 4), cap = (int64_t)sm_count() * 8;
  abs_mean_bwd_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, grad_scale, grad_x);
  TNF_LAUNCH_CHECK("abs_mean_bwd_kernel");
  return TNF_OK;
}

// ---- ray batches: rows of the scene's ray table picked by the shuffled order (src/run.py:116-122,226-228) --------------
// The reference's DataLoader collates `batch_size` random rows of the pinned ray table on the host and copies them to the
// device.  Here the GPU picks the rows itself: `table` may live in HBM or in PINNED HOST memory (unified addressing makes
// cudaHostAlloc'd memory readable from the device at the same address), in which case the 36-byte rows cross the host link as
// zero-copy reads -- the H2D transfer of exactly the rows the batch needs, with no host gather and no staging buffer.
namespace tnf {
namespace {
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, long long row_floats,
                                                          const long long* __restrict__ idx, long long n_rows, float* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = n_rows * row_floats;
  if (t >= total) return;
  const long long r = t / row_floats, c = t - r * row_floats;
  out[t] = ld_stream_f1(table + __ldg(idx + r) * row_floats + c);
}
}  // namespace
}  // namespace tnf

extern "C" int tnf_gather_rows(const float* table, int64_t n_table_rows, int32_t row_floats, const int64_t* idx, int64_t n_rows,
                               float* out, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n_rows >= 0 && n_table_rows >= 0 && row_floats >= 1, "bad sizes");
  if (n_rows == 0) return TNF_OK;
  TNF_REQUIRE(table && idx && out, "null pointer");
  const long long total = n_rows * (long long)row_floats;
  gather_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(table, row_floats,
      reinterpret_cast<const long long*>(idx), n_rows, out);
  TNF_LAUNCH_CHECK("gather_rows_kernel");
  return TNF_OK;
}
