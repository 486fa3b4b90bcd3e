// a13 + SURVEY 8f rank 1/2 -- plane regulariser and optimiser as single streaming passes.
//
// tnf_tv_fwd/bwd: KPlanesFeaturePlane.loss_tv (src/models.py:115-118) = mse(p[1:]-p[:-1]) along H plus
// along W, for a table of channels-last planes [res][res][C].  The reference does this with strided
// slices of NCHW planes through 4 mse_loss calls per plane (fwd+bwd: ~20 elementwise/reduce kernels per
// plane, each a full pass over 132 MB in total); here forward is one read of the planes and backward one
// read + one write (the 5-point stencil's neighbours come from L1/L2).
//
// tnf_adam_step: torch.optim.Adam (src/run.py:186: lr, betas, eps, L2 weight_decay, no amsgrad) for a
// table of tensors in one launch: p, g, m, v read once, p, m, v written once (28 B/param).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace tnf {
namespace {

constexpr int kMaxPlanes = 24;

struct TVArgs {
  const float* planes[kMaxPlanes];
  float* grads[kMaxPlanes];
  int res[kMaxPlanes];
  long long first_block[kMaxPlanes + 1];  // blocks [first_block[i], first_block[i+1]) work on plane i
  float coef_h[kMaxPlanes], coef_w[kMaxPlanes];  // bwd: 2/(n_h*count), 2/(n_w*count)
  int n_planes;
  int channels;
  double* sums;            // fwd: [n_planes][2] sums of squared differences (H, W)
  const float* gscale;     // bwd: device scalar = upstream gradient of the loss
  int accumulate;          // bwd: add into grads instead of overwriting
};

__device__ __forceinline__ int find_plane(const TVArgs& A, long long b) {
  int i = 0;
  while (i + 1 < A.n_planes && b >= A.first_block[i + 1]) ++i;
  return i;
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float sq4(float4 a, float4 b) {
  const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
  return x * x + y * y + z * z + w * w;
}

// one thread = one float4 of channels at one texel; 256 threads per block
__global__ void __launch_bounds__(256) tv_fwd_kernel(const TVArgs A) {
  __shared__ float s_red[2][8];
  const int pi = find_plane(A, blockIdx.x);
  const int res = A.res[pi], C = A.channels, c4 = C >> 2;
  const long long t = (blockIdx.x - A.first_block[pi]) * 256LL + threadIdx.x;
  const long long n_vec = (long long)res * res * c4;
  float sh = 0.f, sw = 0.f;
  if (t < n_vec) {
    const long long texel = t / c4;
    const int h = (int)(texel / res), w = (int)(texel % res);
    const float* p = A.planes[pi] + t * 4;
    const float4 v = ld4(p);
    if (h + 1 < res) sh = sq4(ld4(p + (long long)res * C), v);
    if (w + 1 < res) sw = sq4(ld4(p + C), v);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sh += __shfl_xor_sync(kFullMask, sh, d);
    sw += __shfl_xor_sync(kFullMask, sw, d);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_red[0][wid] = sh; s_red[1][wid] = sw; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += (double)s_red[threadIdx.x][i];
    atomicAdd(A.sums + 2 * pi + threadIdx.x, acc);
  }
}

// SUMS: also emit the forward's sums of squared differences (the "down" and "right" neighbours are loaded for the
// gradient anyway), so one pass over the planes yields loss value and gradient.
template <bool SUMS>
__global__ void __launch_bounds__(256) tv_bwd_kernel(const TVArgs A) {
  __shared__ float s_red[2][8];
  const int pi = find_plane(A, blockIdx.x);
  const int res = A.res[pi], C = A.channels, c4 = C >> 2;
  const long long t = (blockIdx.x - A.first_block[pi]) * 256LL + threadIdx.x;
  const long long n_vec = (long long)res * res * c4;
  float sh = 0.f, sw = 0.f;
  if (t < n_vec) {
    const long long texel = t / c4;
    const int h = (int)(texel / res), w = (int)(texel % res);
    const float gs = __ldg(A.gscale);
    const float ch = A.coef_h[pi] * gs, cw = A.coef_w[pi] * gs;
    const float* p = A.planes[pi] + t * 4;
    const long long sH = (long long)res * C;
    const float4 v = ld4(p);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    // d/dp[h,w] of sum (p[h+1]-p[h])^2 = 2 (p[h]-p[h-1]) [h>0] - 2 (p[h+1]-p[h]) [h<H-1]
    if (h > 0) { const float4 u = ld4(p - sH); g.x += ch * (v.x - u.x); g.y += ch * (v.y - u.y); g.z += ch * (v.z - u.z); g.w += ch * (v.w - u.w); }
    if (h + 1 < res) {
      const float4 u = ld4(p + sH);
      g.x -= ch * (u.x - v.x); g.y -= ch * (u.y - v.y); g.z -= ch * (u.z - v.z); g.w -= ch * (u.w - v.w);
      if (SUMS) sh = sq4(u, v);
    }
    if (w > 0) { const float4 u = ld4(p - C); g.x += cw * (v.x - u.x); g.y += cw * (v.y - u.y); g.z += cw * (v.z - u.z); g.w += cw * (v.w - u.w); }
    if (w + 1 < res) {
      const float4 u = ld4(p + C);
      g.x -= cw * (u.x - v.x); g.y -= cw * (u.y - v.y); g.z -= cw * (u.z - v.z); g.w -= cw * (u.w - v.w);
      if (SUMS) sw = sq4(u, v);
    }
    float4* out = reinterpret_cast<float4*>(A.grads[pi] + t * 4);
    if (A.accumulate) {
      const float4 o = *out;
      g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
    }
    *out = g;
  }
  if (SUMS) {  // same reduction tree as tv_fwd_kernel: identical sums
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sh += __shfl_xor_sync(kFullMask, sh, d);
      sw += __shfl_xor_sync(kFullMask, sw, d);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][wid] = sh; s_red[1][wid] = sw; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += (double)s_red[threadIdx.x][i];
      atomicAdd(A.sums + 2 * pi + threadIdx.x, acc);
    }
  }
}

// The same pass for 32-channel planes whose resolution is a multiple of 32, organised so that every texel is fetched from
// HBM once: a warp owns 4 adjacent columns (lane = texel-in-group * 8 + channel quad: one row of the group is 512
// contiguous bytes) and marches down a segment of kTvRows rows with the rows above / at / below in registers; the left and
// right neighbours come from the adjacent lane groups by shuffle, only the two edge columns of the group are loaded (lines
// the neighbouring warp of the same block is streaming anyway).  tv_bwd_kernel loads five texels per output texel and
// leans on L1/L2 for the four re-reads (3.1 TB/s of algorithmic traffic); this one issues 1 + 2/4 loads per texel.
// Arithmetic per texel is identical to tv_bwd_kernel; the sums are accumulated per thread in fp32 (128 terms) and across threads in double.
constexpr int kTvRows = 32;
struct TVMarchArgs {
  TVArgs tv;
  long long first_warp[kMaxPlanes + 1];   // warps [first_warp[i], first_warp[i+1]) work on plane i
};
template <bool SUMS, bool ACC>
__global__ void __launch_bounds__(256, 3) tv_march_kernel(const TVMarchArgs M) {
  const TVArgs& A = M.tv;
  __shared__ double s_red[2][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * 8 + wid;
  int pi = 0;
  while (pi + 1 < A.n_planes && gw >= M.first_warp[pi + 1]) ++pi;   // all warps of a block are in one plane (res % 32 == 0)
  const int res = A.res[pi];
  const long long lw = gw - M.first_warp[pi];
  const int groups = res >> 2;                        // column groups per row segment
  const int seg = (int)(lw / groups), grp = (int)(lw % groups);
  const int tg = lane >> 3, quad = lane & 7;
  const int w = 4 * grp + tg, h0 = seg * kTvRows;
  const float gs = __ldg(A.gscale);
  const float ch = A.coef_h[pi] * gs, cw = A.coef_w[pi] * gs;
  const long long sH = (long long)res * 32;           // floats per row
  const float* p = A.planes[pi] + ((long long)h0 * res + w) * 32 + quad * 4;
  float* gout = A.grads[pi] + ((long long)h0 * res + w) * 32 + quad * 4;
  const bool has_l = w > 0, has_r = w + 1 < res;
  const bool edge_l = tg == 0 && has_l, edge_r = tg == 3 && has_r;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 up = h0 > 0 ? ld_stream_f4(p - sH) : zero;
  float4 cur = ld_stream_f4(p);
  float shf = 0.f, swf = 0.f;   // 32 rows x 4 channels per thread: fp32 is plenty; widened before the cross-thread sums
#pragma unroll 1
  for (int r0 = 0; r0 < kTvRows; r0 += 4) {
    float4 nxt[4], el[4], er[4], old[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // every load of the four rows is issued before the first use
      const int h = h0 + r0 + i;
      nxt[i] = (h + 1 < res) ? ld_stream_f4(p + (long long)(r0 + i + 1) * sH) : zero;
      el[i] = edge_l ? ld4(p + (long long)(r0 + i) * sH - 32) : zero;
      er[i] = edge_r ? ld4(p + (long long)(r0 + i) * sH + 32) : zero;
      if (ACC) old[i] = *reinterpret_cast<const float4*>(gout + (long long)(r0 + i) * sH);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int h = h0 + r0 + i;
      const float4 v = cur, dn = nxt[i];
      float4 lf, rt;
      lf.x = __shfl_up_sync(kFullMask, v.x, 8); lf.y = __shfl_up_sync(kFullMask, v.y, 8);
      lf.z = __shfl_up_sync(kFullMask, v.z, 8); lf.w = __shfl_up_sync(kFullMask, v.w, 8);
      rt.x = __shfl_down_sync(kFullMask, v.x, 8); rt.y = __shfl_down_sync(kFullMask, v.y, 8);
      rt.z = __shfl_down_sync(kFullMask, v.z, 8); rt.w = __shfl_down_sync(kFullMask, v.w, 8);
      if (tg == 0) lf = el[i];
      if (tg == 3) rt = er[i];
      float4 g = zero;
      if (h > 0) { g.x += ch * (v.x - up.x); g.y += ch * (v.y - up.y); g.z += ch * (v.z - up.z); g.w += ch * (v.w - up.w); }
      if (h + 1 < res) {
        g.x -= ch * (dn.x - v.x); g.y -= ch * (dn.y - v.y); g.z -= ch * (dn.z - v.z); g.w -= ch * (dn.w - v.w);
        if (SUMS) shf += sq4(dn, v);
      }
      if (has_l) { g.x += cw * (v.x - lf.x); g.y += cw * (v.y - lf.y); g.z += cw * (v.z - lf.z); g.w += cw * (v.w - lf.w); }
      if (has_r) {
        g.x -= cw * (rt.x - v.x); g.y -= cw * (rt.y - v.y); g.z -= cw * (rt.z - v.z); g.w -= cw * (rt.w - v.w);
        if (SUMS) swf += sq4(rt, v);
      }
      if (ACC) { g.x += old[i].x; g.y += old[i].y; g.z += old[i].z; g.w += old[i].w; }
      st_stream_f4(gout + (long long)(r0 + i) * sH, g);
      up = cur;
      cur = dn;
    }
  }
  if (SUMS) {
    double sh = (double)shf, sw = (double)swf;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sh += __shfl_xor_sync(kFullMask, sh, d);
      sw += __shfl_xor_sync(kFullMask, sw, d);
    }
    if (lane == 0) { s_red[0][wid] = sh; s_red[1][wid] = sw; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += s_red[threadIdx.x][i];
      atomicAdd(A.sums + 2 * pi + threadIdx.x, acc);
    }
  }
}
// launches tv_march_kernel when the planes allow it; returns false (nothing launched) otherwise
template <bool SUMS>
bool launch_tv_march(const TVArgs& A, cudaStream_t st) {
  if (A.channels != 32 || variant(kVariantTvTexel) == 1) return false;   // diagnostics / tests: the one-thread-per-texel kernel
  TVMarchArgs M{};
  M.tv = A;
  long long warps = 0;
  for (int i = 0; i < A.n_planes; ++i) {
    if (A.res[i] % 32 != 0 || A.res[i] < 32) return false;
    M.first_warp[i] = warps;
    warps += (long long)(A.res[i] / 4) * (A.res[i] / kTvRows);
  }
  M.first_warp[A.n_planes] = warps;   // a multiple of 8: res/4 is
  if (A.accumulate) tv_march_kernel<SUMS, true><<<(unsigned)(warps / 8), 256, 0, st>>>(M);
  else tv_march_kernel<SUMS, false><<<(unsigned)(warps / 8), 256, 0, st>>>(M);
  return true;
}

int fill_tv(TVArgs* A, const float* const* planes, float* const* grads, const int32_t* res, int n_planes,
            int channels) {
  TNF_REQUIRE(n_planes >= 1 && n_planes <= kMaxPlanes, "n_planes must be in [1,%d]", kMaxPlanes);
  TNF_REQUIRE(channels >= 4 && channels % 4 == 0, "channels must be a positive multiple of 4");
  TNF_REQUIRE(planes && res, "null plane table");
  long long blocks = 0;
  for (int i = 0; i < n_planes; ++i) {
    TNF_REQUIRE(res[i] >= 2, "plane resolution must be >= 2");
    TNF_REQUIRE(planes[i] && (reinterpret_cast<uintptr_t>(planes[i]) & 15u) == 0, "plane %d null/misaligned", i);
    A->planes[i] = planes[i];
    if (grads) {
      TNF_REQUIRE(grads[i] && (reinterpret_cast<uintptr_t>(grads[i]) & 15u) == 0, "grad %d null/misaligned", i);
      A->grads[i] = grads[i];
    }
    A->res[i] = res[i];
    A->first_block[i] = blocks;
    blocks += ceil_div((long long)res[i] * res[i] * (channels / 4), 256);
    const double nh = (double)channels * (res[i] - 1) * res[i];
    A->coef_h[i] = (float)(2.0 / nh);
    A->coef_w[i] = (float)(2.0 / nh);
  }
  A->first_block[n_planes] = blocks;
  A->n_planes = n_planes;
  A->channels = channels;
  TNF_REQUIRE(blocks < (1LL << 31), "too many blocks");
  return TNF_OK;
}

// ---- Adam ---------------------------------------------------------------------------------------
constexpr int kMaxTensors = 48;
struct AdamArgs {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  long long n[kMaxTensors];
  long long first_block[kMaxTensors + 1];
  int n_tensors;
  float lr, beta1, beta2, eps, weight_decay;
  float bias1, bias2_sqrt;  // 1 - beta1^t, sqrt(1 - beta2^t)
  long long total_blocks;   // virtual blocks of 4096 elements; the grid strides over them (grid < total_blocks: persistent form)
};
constexpr int kAdamVec = 4;       // floats per thread per iteration
constexpr int kAdamPerBlock = 256 * kAdamVec * 4;  // 4096 elements per block

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& A) {
  // torch/optim/adam.py (_multi_tensor_adam): grad += wd*p; m.lerp_(g, 1-b1); v = v*b2 + (1-b2) g*g;
  // denom = sqrt(v)/sqrt(bias2) + eps; p -= (lr/bias1) * m/denom
  g = __fmaf_rn(p, A.weight_decay, g);
  m = m + (1.f - A.beta1) * (g - m);
  v = __fmaf_rn((1.f - A.beta2) * g, g, v * A.beta2);
  const float denom = sqrtf(v) / A.bias2_sqrt + A.eps;
  p = p - (A.lr / A.bias1) * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs A) {
  int ti = 0;
  for (long long vb = blockIdx.x; vb < A.total_blocks; vb += gridDim.x) {
  while (ti + 1 < A.n_tensors && vb >= A.first_block[ti + 1]) ++ti;
  const long long base = (vb - A.first_block[ti]) * kAdamPerBlock;
  const long long n = A.n[ti];
  float* p = A.p[ti]; const float* g = A.g[ti]; float* m = A.m[ti]; float* v = A.v[ti];
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const long long i = base + (it * 256 + threadIdx.x) * kAdamVec;
    if (vec && i + 3 < n) {
      float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      const float4 gg = ld_stream_f4(g + i);
      adam_one(pp.x, gg.x, mm.x, vv.x, A); adam_one(pp.y, gg.y, mm.y, vv.y, A);
      adam_one(pp.z, gg.z, mm.z, vv.z, A); adam_one(pp.w, gg.w, mm.w, vv.w, A);
      *reinterpret_cast<float4*>(p + i) = pp; *reinterpret_cast<float4*>(m + i) = mm; *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int k = 0; k < kAdamVec; ++k)
        if (i + k < n) adam_one(p[i + k], g[i + k], m[i + k], v[i + k], A);
    }
  }
  }
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_tv_fwd(const float* const* planes, const int32_t* res, int32_t n_planes, int32_t channels,
                          double* sums, void* stream) {
  using namespace tnf;
  TVArgs A{};
  int rc = fill_tv(&A, planes, nullptr, res, n_planes, channels);
  if (rc != TNF_OK) return rc;
  TNF_REQUIRE(sums, "null sums");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TNF_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * n_planes, st));
  A.sums = sums;
  tv_fwd_kernel<<<(unsigned)A.first_block[n_planes], 256, 0, st>>>(A);
  TNF_LAUNCH_CHECK("tv_fwd_kernel");
  return TNF_OK;
}

extern "C" int tnf_tv_bwd(const float* const* planes, float* const* grads, const int32_t* res, int32_t n_planes,
                          int32_t channels, const float* plane_weight, const float* gscale, int32_t accumulate,
                          void* stream) {
  using namespace tnf;
  TVArgs A{};
  TNF_REQUIRE(grads && gscale, "null grads/gscale");
  int rc = fill_tv(&A, planes, grads, res, n_planes, channels);
  if (rc != TNF_OK) return rc;
  if (plane_weight)
    for (int i = 0; i < n_planes; ++i) { A.coef_h[i] *= plane_weight[i]; A.coef_w[i] *= plane_weight[i]; }
  A.gscale = gscale;
  A.accumulate = accumulate;
  if (!launch_tv_march<false>(A, static_cast<cudaStream_t>(stream)))
    tv_bwd_kernel<false><<<(unsigned)A.first_block[n_planes], 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  TNF_LAUNCH_CHECK("tv_bwd_kernel");
  return TNF_OK;
}

extern "C" int tnf_tv_fwd_bwd(const float* const* planes, float* const* grads, const int32_t* res, int32_t n_planes,
                              int32_t channels, const float* plane_weight, const float* gscale, int32_t accumulate,
                              double* sums, void* stream) {
  using namespace tnf;
  TVArgs A{};
  TNF_REQUIRE(grads && gscale && sums, "null grads/gscale/sums");
  int rc = fill_tv(&A, planes, grads, res, n_planes, channels);
  if (rc != TNF_OK) return rc;
  if (plane_weight)
    for (int i = 0; i < n_planes; ++i) { A.coef_h[i] *= plane_weight[i]; A.coef_w[i] *= plane_weight[i]; }
  A.gscale = gscale;
  A.accumulate = accumulate;
  A.sums = sums;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TNF_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * n_planes, st));
  if (!launch_tv_march<true>(A, st)) tv_bwd_kernel<true><<<(unsigned)A.first_block[n_planes], 256, 0, st>>>(A);
  TNF_LAUNCH_CHECK("tv_fwd_bwd_kernel");
  return TNF_OK;
}

extern "C" int tnf_adam_step_grid(float* const* params, const float* const* grads, float* const* exp_avg,
                                  float* const* exp_avg_sq, const int64_t* numel, int32_t n_tensors, float lr,
                                  float beta1, float beta2, float eps, float weight_decay, int64_t step, int32_t max_blocks,
                                  void* stream);
extern "C" int tnf_adam_step(float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int64_t* numel, int32_t n_tensors, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream) {
  return tnf_adam_step_grid(params, grads, exp_avg, exp_avg_sq, numel, n_tensors, lr, beta1, beta2, eps, weight_decay, step, 0, stream);
}
extern "C" int tnf_adam_step_grid(float* const* params, const float* const* grads, float* const* exp_avg,
                                  float* const* exp_avg_sq, const int64_t* numel, int32_t n_tensors, float lr,
                                  float beta1, float beta2, float eps, float weight_decay, int64_t step, int32_t max_blocks,
                                  void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n_tensors >= 0 && step >= 1, "bad n_tensors/step");
  TNF_REQUIRE(n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && numel), "null table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int t0 = 0; t0 < n_tensors; t0 += kMaxTensors) {
    AdamArgs A{};
    const int cnt = (n_tensors - t0 < kMaxTensors) ? (n_tensors - t0) : kMaxTensors;
    long long blocks = 0;
    for (int i = 0; i < cnt; ++i) {
      TNF_REQUIRE(numel[t0 + i] >= 0, "negative numel");
      TNF_REQUIRE(params[t0 + i] && grads[t0 + i] && exp_avg[t0 + i] && exp_avg_sq[t0 + i], "null tensor %d", t0 + i);
      A.p[i] = params[t0 + i]; A.g[i] = grads[t0 + i]; A.m[i] = exp_avg[t0 + i]; A.v[i] = exp_avg_sq[t0 + i];
      A.n[i] = numel[t0 + i];
      A.first_block[i] = blocks;
      blocks += ceil_div(numel[t0 + i], kAdamPerBlock);
    }
    A.first_block[cnt] = blocks;
    A.n_tensors = cnt;
    A.lr = lr; A.beta1 = beta1; A.beta2 = beta2; A.eps = eps; A.weight_decay = weight_decay;
    A.bias1 = (float)(1.0 - pow((double)beta1, (double)step));
    A.bias2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    if (blocks == 0) continue;
    TNF_REQUIRE(blocks < (1LL << 31), "too many blocks");
    A.total_blocks = blocks;
    const long long grid = (max_blocks > 0 && max_blocks < blocks) ? max_blocks : blocks;
    adam_kernel<<<(unsigned)grid, 256, 0, st>>>(A);
    TNF_LAUNCH_CHECK("adam_kernel");
  }
  return TNF_OK;
}
