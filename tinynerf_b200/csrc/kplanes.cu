// a12 -- K-Planes fused feature lookup (reference: KPlanesFeatureField.forward src/models.py:153-163,
// KPlanesFeaturePlane.forward :105-113 = F.grid_sample(bilinear, zeros, align_corners=True)).
//
// The reference launches 9 grid_sample kernels over NCHW planes (per-corner channel stride H*W*4 B
// = 32 separate sectors for 32 channels), 9 transposes, 6 multiplies and a concat.  Here the planes
// are stored channels-last ([res][res][C]), so one bilinear corner is ONE 128-byte line read by
// C/4 adjacent lanes with a single 128-bit load each; all 9 planes, the 3 Hadamard products and the
// concat happen in registers and the [N, 3*C] feature row is written once.
// Backward recomputes the plane features (cheaper than saving 9x[N,C]) and scatters the plane
// gradients with vector reductions (red.global.add.v4.f32: one L2 op per 16 B instead of four).
#include "common.cuh"
#include "nerf_math.cuh"

namespace tnf {
namespace {

constexpr int kMaxScales = 8;

struct KPArgs {
  const float* planes[kMaxScales * 3];
  float* grads[kMaxScales * 3];
  int res[kMaxScales];
  int n_scales;
  int channels;
  const float* x;
  long long x_stride;
  long long n;
  float* out;
  const float* grad_out;
};

struct Bilinear {
  int x0, y0;
  float w[4];    // nw, ne, sw, se  (torch grid_sampler_2d order)
  bool ok[4];
};

// torch grid_sampler_2d (bilinear, zeros, align_corners=True) index/weight arithmetic
__device__ __forceinline__ Bilinear bilinear_setup(float gx, float gy, int res) {
  Bilinear b;
  const float ix = TNF_MUL(TNF_MUL(TNF_ADD(gx, 1.f), 0.5f), (float)(res - 1));
  const float iy = TNF_MUL(TNF_MUL(TNF_ADD(gy, 1.f), 0.5f), (float)(res - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  b.x0 = (int)fx;
  b.y0 = (int)fy;
  const float wx0 = TNF_SUB((float)(b.x0 + 1), ix), wx1 = TNF_SUB(ix, (float)b.x0);
  const float wy0 = TNF_SUB((float)(b.y0 + 1), iy), wy1 = TNF_SUB(iy, (float)b.y0);
  b.w[0] = TNF_MUL(wx0, wy0);
  b.w[1] = TNF_MUL(wx1, wy0);
  b.w[2] = TNF_MUL(wx0, wy1);
  b.w[3] = TNF_MUL(wx1, wy1);
  const bool bx0 = (unsigned)b.x0 < (unsigned)res, bx1 = (unsigned)(b.x0 + 1) < (unsigned)res;
  const bool by0 = (unsigned)b.y0 < (unsigned)res, by1 = (unsigned)(b.y0 + 1) < (unsigned)res;
  b.ok[0] = bx0 && by0;
  b.ok[1] = bx1 && by0;
  b.ok[2] = bx0 && by1;
  b.ok[3] = bx1 && by1;
  return b;
}

__device__ __forceinline__ long long corner_offset(const Bilinear& b, int k, int res, int C) {
  const int xx = b.x0 + (k & 1), yy = b.y0 + (k >> 1);
  return ((long long)yy * res + xx) * C;
}

__device__ __forceinline__ float4 gather_plane(const float* __restrict__ plane, const Bilinear& b, int res,
                                               int C, int ch) {
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    v[k] = b.ok[k] ? __ldg(reinterpret_cast<const float4*>(plane + corner_offset(b, k, res, C) + ch))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (b.ok[k]) {
      acc.x = TNF_FMA(v[k].x, b.w[k], acc.x);
      acc.y = TNF_FMA(v[k].y, b.w[k], acc.y);
      acc.z = TNF_FMA(v[k].z, b.w[k], acc.z);
      acc.w = TNF_FMA(v[k].w, b.w[k], acc.w);
    }
  }
  return acc;
}

__device__ __forceinline__ float4 mul4(float4 a, float4 b) {
  return make_float4(TNF_MUL(a.x, b.x), TNF_MUL(a.y, b.y), TNF_MUL(a.z, b.z), TNF_MUL(a.w, b.w));
}

// dimension pairs of itertools.combinations(range(3), 2) (src/models.py:145): first -> x/W, second -> y/H
__device__ __constant__ int kPairA[3] = {0, 0, 1};
__device__ __constant__ int kPairB[3] = {1, 2, 2};

template <bool BWD>
__global__ void __launch_bounds__(256) kplanes_kernel(const KPArgs A) {
  const int lps = A.channels >> 2;  // lanes per sample (power of two, <= 8)
  const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n = gt / lps;
  const int ch = (int)(gt % lps) * 4;
  if (n >= A.n) return;
  float xyz[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) xyz[c] = __ldg(A.x + n * A.x_stride + c);
  const int F = A.n_scales * A.channels;
  for (int s = 0; s < A.n_scales; ++s) {
    const int res = A.res[s];
    Bilinear b[3];
    float4 f[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      b[p] = bilinear_setup(xyz[kPairA[p]], xyz[kPairB[p]], res);
      f[p] = gather_plane(A.planes[s * 3 + p], b[p], res, A.channels, ch);
    }
    if (!BWD) {
      // current_scale_features = 1.; *= plane0; *= plane1; *= plane2  (src/models.py:158-160)
      const float4 o = mul4(mul4(f[0], f[1]), f[2]);
      *reinterpret_cast<float4*>(A.out + n * F + s * A.channels + ch) = o;
    } else {
      const float4 g = __ldg(reinterpret_cast<const float4*>(A.grad_out + n * F + s * A.channels + ch));
      // autograd of ((f0*f1)*f2): d f2 = g*(f0*f1); d(f0*f1) = g*f2; d f0 = (g*f2)*f1; d f1 = (g*f2)*f0
      const float4 g01 = mul4(g, f[2]);
      float4 gp[3];
      gp[0] = mul4(g01, f[1]);
      gp[1] = mul4(g01, f[0]);
      gp[2] = mul4(g, mul4(f[0], f[1]));
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        float* gpl = A.grads[s * 3 + p];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (b[p].ok[k]) {
            const float wk = b[p].w[k];
            red_add_f4(gpl + corner_offset(b[p], k, res, A.channels) + ch,
                       make_float4(TNF_MUL(wk, gp[p].x), TNF_MUL(wk, gp[p].y), TNF_MUL(wk, gp[p].z),
                                   TNF_MUL(wk, gp[p].w)));
          }
        }
      }
    }
  }
}

int fill_args(KPArgs* A, const float* const* planes, float* const* grads, const int32_t* res, int n_scales,
              int channels, const float* x, int64_t x_stride, int64_t n) {
  TNF_REQUIRE(n >= 0, "negative n");
  TNF_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "n_scales must be in [1,%d]", kMaxScales);
  TNF_REQUIRE(channels == 4 || channels == 8 || channels == 16 || channels == 32,
              "channels must be 4, 8, 16 or 32 (got %d)", channels);
  TNF_REQUIRE(planes && res, "null plane table");
  TNF_REQUIRE(n == 0 || x, "null x");
  TNF_REQUIRE(x_stride >= 3, "x_stride must be >= 3");
  for (int i = 0; i < n_scales * 3; ++i) {
    TNF_REQUIRE(planes[i] && (reinterpret_cast<uintptr_t>(planes[i]) & 15u) == 0, "plane %d null/misaligned", i);
    A->planes[i] = planes[i];
    if (grads) {
      TNF_REQUIRE(grads[i] && (reinterpret_cast<uintptr_t>(grads[i]) & 15u) == 0, "grad plane %d null/misaligned", i);
      A->grads[i] = grads[i];
    }
  }
  for (int s = 0; s < n_scales; ++s) {
    TNF_REQUIRE(res[s] >= 2, "plane resolution must be >= 2");
    A->res[s] = res[s];
  }
  A->n_scales = n_scales;
  A->channels = channels;
  A->x = x;
  A->x_stride = x_stride;
  A->n = n;
  return TNF_OK;
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_kplanes_fwd(const float* const* planes, const int32_t* res, int32_t n_scales,
                               int32_t channels, const float* x, int64_t x_stride, int64_t n, float* out,
                               void* stream) {
  using namespace tnf;
  KPArgs A{};
  int rc = fill_args(&A, planes, nullptr, res, n_scales, channels, x, x_stride, n);
  if (rc != TNF_OK || n == 0) return rc;
  TNF_REQUIRE(out && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "out null/misaligned");
  A.out = out;
  const long long threads = n * (channels / 4);
  kplanes_kernel<false><<<(unsigned)ceil_div(threads, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  TNF_LAUNCH_CHECK("kplanes_fwd_kernel");
  return TNF_OK;
}

extern "C" int tnf_kplanes_bwd(const float* const* planes, float* const* grad_planes, const int32_t* res,
                               int32_t n_scales, int32_t channels, const float* x, int64_t x_stride,
                               int64_t n, const float* grad_out, void* stream) {
  using namespace tnf;
  KPArgs A{};
  TNF_REQUIRE(grad_planes, "null grad plane table");
  int rc = fill_args(&A, planes, grad_planes, res, n_scales, channels, x, x_stride, n);
  if (rc != TNF_OK || n == 0) return rc;
  TNF_REQUIRE(grad_out && (reinterpret_cast<uintptr_t>(grad_out) & 15u) == 0, "grad_out null/misaligned");
  A.grad_out = grad_out;
  const long long threads = n * (channels / 4);
  kplanes_kernel<true><<<(unsigned)ceil_div(threads, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  TNF_LAUNCH_CHECK("kplanes_bwd_kernel");
  return TNF_OK;
}
