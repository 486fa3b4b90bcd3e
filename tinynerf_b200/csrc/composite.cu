// a18 -- compositing: per-ray segment sums of w*rgb and w over packed samples, + background
// (reference: the index_add_ block of NerfRenderer.forward, src/core.py:256-265, which the reference
// itself marks "TODO: cuda kernel this").  The reference builds a [N] ray-index tensor with
// repeat_interleave (host sync) and scatters with float atomics; here packing info gives each ray its
// contiguous segment, so one warp reduces one ray with coalesced reads and no atomics (deterministic).
#include "common.cuh"

namespace tnf {
namespace {

constexpr int kWarpsC = 8;

__global__ void __launch_bounds__(kWarpsC * 32)
composite_fwd_kernel(const float* __restrict__ w, const float* __restrict__ rgb, const int2* __restrict__ info,
                     long long n_samples, long long n_rays, bool has_bg, float bg0, float bg1, float bg2,
                     float* __restrict__ out_rgb, float* __restrict__ out_op) {
  const int lane = threadIdx.x & 31;
  const long long ray = blockIdx.x * (long long)kWarpsC + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int2 e = __ldg(&info[ray]);
  long long k0 = e.x, k1 = (long long)e.x + e.y;
  if (k0 < 0) k0 = 0;
  if (k1 > n_samples) k1 = n_samples;
  float r = 0.f, g = 0.f, b = 0.f, op = 0.f;
  for (long long k = k0 + lane; k < k1; k += 32) {
    const float wk = __ldg(w + k);
    if (wk > 0.f) {  // colours exist only where weights > 0 (the reference's mask, src/core.py:243-250)
      r = __fmaf_rn(wk, __ldg(rgb + 3 * k + 0), r);
      g = __fmaf_rn(wk, __ldg(rgb + 3 * k + 1), g);
      b = __fmaf_rn(wk, __ldg(rgb + 3 * k + 2), b);
    }
    op += wk;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    r += __shfl_xor_sync(kFullMask, r, d);
    g += __shfl_xor_sync(kFullMask, g, d);
    b += __shfl_xor_sync(kFullMask, b, d);
    op += __shfl_xor_sync(kFullMask, op, d);
  }
  if (lane == 0) {
    if (has_bg) {  // rendered + bg * (1 - opacity)   (src/core.py:265)
      const float t = __fsub_rn(1.f, op);
      r = __fadd_rn(r, __fmul_rn(bg0, t));
      g = __fadd_rn(g, __fmul_rn(bg1, t));
      b = __fadd_rn(b, __fmul_rn(bg2, t));
    }
    out_rgb[3 * ray + 0] = r;
    out_rgb[3 * ray + 1] = g;
    out_rgb[3 * ray + 2] = b;
    if (out_op) out_op[ray] = op;
  }
}

__global__ void __launch_bounds__(kWarpsC * 32)
composite_bwd_kernel(const float* __restrict__ w, const float* __restrict__ rgb, const int2* __restrict__ info,
                     long long n_samples, long long n_rays, bool has_bg, float bg0, float bg1, float bg2,
                     const float* __restrict__ go, float* __restrict__ gw, float* __restrict__ grgb) {
  const int lane = threadIdx.x & 31;
  const long long ray = blockIdx.x * (long long)kWarpsC + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int2 e = __ldg(&info[ray]);
  long long k0 = e.x, k1 = (long long)e.x + e.y;
  if (k0 < 0) k0 = 0;
  if (k1 > n_samples) k1 = n_samples;
  const float g0 = __ldg(go + 3 * ray), g1 = __ldg(go + 3 * ray + 1), g2 = __ldg(go + 3 * ray + 2);
  const float gbg = has_bg ? (bg0 * g0 + bg1 * g1 + bg2 * g2) : 0.f;
  for (long long k = k0 + lane; k < k1; k += 32) {
    const float wk = __ldg(w + k);
    const bool on = wk > 0.f;  // masked-out samples carry colour 0 in the reference: no colour term in gw, no grgb
    if (gw) gw[k] = (on ? __ldg(rgb + 3 * k) * g0 + __ldg(rgb + 3 * k + 1) * g1 + __ldg(rgb + 3 * k + 2) * g2 : 0.f) - gbg;
    if (grgb) {
      grgb[3 * k + 0] = on ? wk * g0 : 0.f;
      grgb[3 * k + 1] = on ? wk * g1 : 0.f;
      grgb[3 * k + 2] = on ? wk * g2 : 0.f;
    }
  }
}

// MSE over the union batch of all ranks + its gradient (src/run.py:252 MSELoss, :259 scaled backward):
//   loss_out = sum (r-t)^2 / denom ;  grad[i] = grad_scale * 2 (r_i - t_i) / denom ,  denom = n_rays_global * 3
// One CTA, fixed reduction order: deterministic.
__global__ void __launch_bounds__(1024) mse_loss_grad_kernel(const float* __restrict__ r, const float* __restrict__ t, long long n,
                                                             float n_rays_global, const float* __restrict__ n_rays_global_dev,
                                                             float grad_scale, float* __restrict__ grad, float* __restrict__ loss_out) {
  __shared__ double s_red[32];
  const float denom = (n_rays_global_dev ? __ldg(n_rays_global_dev) : n_rays_global) * 3.f;
  const float gs = grad_scale / denom;   // d(loss*grad_scale)/d(sum) as autograd forms it: scale / denom
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const float d = __ldg(r + i) - __ldg(t + i);
    acc += (double)(d * d);
    if (grad) grad[i] = gs * 2.f * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = s_red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
    if (threadIdx.x == 0 && loss_out) *loss_out = (float)(acc / (double)denom);
  }
}

// composite forward + MSE + composite backward of one iteration in a single pass (src/core.py:256-265, src/run.py:252,259):
// the loss gradient of a ray depends on that ray's rendered colour only (the normaliser is a scalar known beforehand), so
// the warp that reduced a ray's segment turns round and writes the segment's grad_weights / grad_rgbs while its samples
// are still in L1.  Same arithmetic per element as the three separate kernels; the loss is summed in double per block and
// across blocks with one atomic each, the last block to finish writes loss_out and re-zeroes the scratch words.
__global__ void __launch_bounds__(kWarpsC * 32)
composite_loss_kernel(const float* __restrict__ w, const float* __restrict__ rgb, const int2* __restrict__ info,
                      long long n_samples, long long n_rays, bool has_bg, float bg0, float bg1, float bg2,
                      const float* __restrict__ target, float n_rays_global, const float* __restrict__ n_rays_global_dev,
                      float grad_scale, float* __restrict__ out_rgb, float* __restrict__ gw, float* __restrict__ grgb,
                      float* __restrict__ loss_out, double* __restrict__ scratch, const double* __restrict__ extra_terms,
                      const double* __restrict__ extra_coef, int n_extra) {
  __shared__ double s_part[kWarpsC];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long ray = blockIdx.x * (long long)kWarpsC + wid;
  const float denom = (n_rays_global_dev ? __ldg(n_rays_global_dev) : n_rays_global) * 3.f;
  const float gs = grad_scale / denom;
  double part = 0.0;
  if (ray < n_rays) {
    const int2 e = __ldg(&info[ray]);
    long long k0 = e.x, k1 = (long long)e.x + e.y;
    if (k0 < 0) k0 = 0;
    if (k1 > n_samples) k1 = n_samples;
    float r = 0.f, g = 0.f, b = 0.f, op = 0.f;
    for (long long k = k0 + lane; k < k1; k += 32) {
      const float wk = __ldg(w + k);
      if (wk > 0.f) {
        r = __fmaf_rn(wk, __ldg(rgb + 3 * k + 0), r);
        g = __fmaf_rn(wk, __ldg(rgb + 3 * k + 1), g);
        b = __fmaf_rn(wk, __ldg(rgb + 3 * k + 2), b);
      }
      op += wk;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      r += __shfl_xor_sync(kFullMask, r, d);
      g += __shfl_xor_sync(kFullMask, g, d);
      b += __shfl_xor_sync(kFullMask, b, d);
      op += __shfl_xor_sync(kFullMask, op, d);
    }
    if (has_bg) {
      const float t = __fsub_rn(1.f, op);
      r = __fadd_rn(r, __fmul_rn(bg0, t));
      g = __fadd_rn(g, __fmul_rn(bg1, t));
      b = __fadd_rn(b, __fmul_rn(bg2, t));
    }
    const float d0 = r - __ldg(target + 3 * ray), d1 = g - __ldg(target + 3 * ray + 1), d2 = b - __ldg(target + 3 * ray + 2);
    if (lane == 0) {
      out_rgb[3 * ray + 0] = r;
      out_rgb[3 * ray + 1] = g;
      out_rgb[3 * ray + 2] = b;
      part = (double)(d0 * d0) + (double)(d1 * d1) + (double)(d2 * d2);
    }
    const float g0 = gs * 2.f * d0, g1 = gs * 2.f * d1, g2 = gs * 2.f * d2;
    const float gbg = has_bg ? (bg0 * g0 + bg1 * g1 + bg2 * g2) : 0.f;
    for (long long k = k0 + lane; k < k1; k += 32) {
      const float wk = __ldg(w + k);
      const bool on = wk > 0.f;
      gw[k] = (on ? __ldg(rgb + 3 * k) * g0 + __ldg(rgb + 3 * k + 1) * g1 + __ldg(rgb + 3 * k + 2) * g2 : 0.f) - gbg;
      grgb[3 * k + 0] = on ? wk * g0 : 0.f;
      grgb[3 * k + 1] = on ? wk * g1 : 0.f;
      grgb[3 * k + 2] = on ? wk * g2 : 0.f;
    }
  }
  if (lane == 0) s_part[wid] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < kWarpsC; ++i) acc += s_part[i];
    atomicAdd(scratch, acc);
    __threadfence();
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(scratch + 1);
    s_last = atomicAdd(counter, 1ull) == (unsigned long long)gridDim.x - 1;
    if (s_last) {
      __threadfence();
      double loss = atomicAdd(scratch, 0.0) / (double)denom;
      for (int i = 0; i < n_extra; ++i) loss += extra_terms[i] * extra_coef[i];   // e.g. the weighted TV sums of this iteration
      *loss_out = (float)loss;
      *scratch = 0.0;
      *counter = 0ull;
    }
  }
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_composite_loss_fwd_bwd(const float* weights, const float* rgbs, const int32_t* info, int64_t n_samples,
                                          int64_t n_rays, const float* bg, const float* target, float n_rays_global,
                                          const float* n_rays_global_dev, float grad_scale, float* out_rgb,
                                          float* grad_weights, float* grad_rgbs, float* loss_out, void* scratch,
                                          const double* extra_terms, const double* extra_coef, int32_t n_extra, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n_samples >= 0 && n_rays >= 1, "bad sizes");
  TNF_REQUIRE(weights && rgbs && info && target && out_rgb && grad_weights && grad_rgbs && loss_out && scratch, "null pointer");
  TNF_REQUIRE((reinterpret_cast<uintptr_t>(info) & 7u) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 7u) == 0,
              "info / scratch must be 8-byte aligned");
  TNF_REQUIRE(n_rays_global_dev || n_rays_global > 0.f, "n_rays_global must be positive");
  TNF_REQUIRE(n_extra >= 0 && (n_extra == 0 || (extra_terms && extra_coef)), "bad extra loss terms");
  composite_loss_kernel<<<(unsigned)ceil_div(n_rays, kWarpsC), kWarpsC * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      weights, rgbs, reinterpret_cast<const int2*>(info), n_samples, n_rays, bg != nullptr, bg ? bg[0] : 0.f, bg ? bg[1] : 0.f,
      bg ? bg[2] : 0.f, target, n_rays_global, n_rays_global_dev, grad_scale, out_rgb, grad_weights, grad_rgbs, loss_out,
      static_cast<double*>(scratch), extra_terms, extra_coef, n_extra);
  TNF_LAUNCH_CHECK("composite_loss_kernel");
  return TNF_OK;
}

extern "C" int tnf_mse_loss_grad(const float* rendered, const float* target, int64_t n_rays, float n_rays_global,
                                 const float* n_rays_global_dev, float grad_scale, float* grad_rendered, float* loss_out,
                                 void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n_rays >= 0, "negative size");
  TNF_REQUIRE(grad_rendered || loss_out, "nothing requested");
  TNF_REQUIRE(n_rays == 0 || (rendered && target), "null pointer");
  TNF_REQUIRE(n_rays_global_dev || n_rays_global > 0.f, "n_rays_global must be positive");
  mse_loss_grad_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(rendered, target, 3 * n_rays, n_rays_global,
                                                                          n_rays_global_dev, grad_scale, grad_rendered, loss_out);
  TNF_LAUNCH_CHECK("mse_loss_grad_kernel");
  return TNF_OK;
}

extern "C" int tnf_composite_fwd(const float* weights, const float* rgbs, const int32_t* info,
                                 int64_t n_samples, int64_t n_rays, const float* bg, float* out_rgb,
                                 float* out_opacity, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n_samples >= 0 && n_rays >= 0, "negative size");
  if (n_rays == 0) return TNF_OK;
  TNF_REQUIRE(info && out_rgb, "null pointer");
  TNF_REQUIRE(n_samples == 0 || (weights && rgbs), "null sample pointer");
  TNF_REQUIRE((reinterpret_cast<uintptr_t>(info) & 7u) == 0, "info must be 8-byte aligned");
  composite_fwd_kernel<<<(unsigned)ceil_div(n_rays, kWarpsC), kWarpsC * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      weights, rgbs, reinterpret_cast<const int2*>(info), n_samples, n_rays, bg != nullptr, bg ? bg[0] : 0.f,
      bg ? bg[1] : 0.f, bg ? bg[2] : 0.f, out_rgb, out_opacity);
  TNF_LAUNCH_CHECK("composite_fwd_kernel");
  return TNF_OK;
}

extern "C" int tnf_composite_bwd(const float* weights, const float* rgbs, const int32_t* info,
                                 int64_t n_samples, int64_t n_rays, const float* bg, const float* grad_out,
                                 float* grad_weights, float* grad_rgbs, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n_samples >= 0 && n_rays >= 0, "negative size");
  if (n_rays == 0 || n_samples == 0) return TNF_OK;
  TNF_REQUIRE(info && grad_out && weights && rgbs, "null pointer");
  TNF_REQUIRE(grad_weights || grad_rgbs, "no gradient requested");
  TNF_REQUIRE((reinterpret_cast<uintptr_t>(info) & 7u) == 0, "info must be 8-byte aligned");
  composite_bwd_kernel<<<(unsigned)ceil_div(n_rays, kWarpsC), kWarpsC * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      weights, rgbs, reinterpret_cast<const int2*>(info), n_samples, n_rays, bg != nullptr, bg ? bg[0] : 0.f,
      bg ? bg[1] : 0.f, bg ? bg[2] : 0.f, grad_out, grad_weights, grad_rgbs);
  TNF_LAUNCH_CHECK("composite_bwd_kernel");
  return TNF_OK;
}
