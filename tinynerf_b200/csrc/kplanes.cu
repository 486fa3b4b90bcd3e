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
  int scale0;   // first scale handled by this launch (blockIdx.y counts from it)
};

// One axis of torch's grid_sampler (bilinear, zeros padding, align_corners=True):
//   i = ((c+1)/2)*(res-1); i0 = floor(i); w0 = (i0+1) - i; w1 = i - i0
// Out-of-range corners are skipped by torch; here their axis weight is zeroed and the index clamped,
// which adds an exact +-0 instead (identical values for finite planes) and keeps every load unpredicated.
struct Axis {
  int i0, i1;     // clamped texel indices
  float w0, w1;   // weights of i0 / i1 (0 when that texel is outside the plane)
};
__device__ __forceinline__ Axis axis_setup(float c, int res) {
  Axis a;
  const float i = TNF_MUL(TNF_MUL(TNF_ADD(c, 1.f), 0.5f), (float)(res - 1));
  const float f = floorf(i);
  const int i0 = (int)f;
  a.w0 = ((unsigned)i0 < (unsigned)res) ? TNF_SUB(f + 1.f, i) : 0.f;
  a.w1 = ((unsigned)(i0 + 1) < (unsigned)res) ? TNF_SUB(i, f) : 0.f;
  a.i0 = min(max(i0, 0), res - 1);
  a.i1 = min(max(i0 + 1, 0), res - 1);
  return a;
}

struct Corners {
  int off[4];   // element offsets of nw, ne, sw, se (torch grid_sampler_2d order)
  float w[4];
};
// plane indexed [y][x][C]; ax -> x/W axis, ay -> y/H axis
__device__ __forceinline__ Corners corners(const Axis& ax, const Axis& ay, int res, int C, int ch) {
  Corners c;
  const int r0 = ay.i0 * res, r1 = ay.i1 * res;
  c.off[0] = (r0 + ax.i0) * C + ch;
  c.off[1] = (r0 + ax.i1) * C + ch;
  c.off[2] = (r1 + ax.i0) * C + ch;
  c.off[3] = (r1 + ax.i1) * C + ch;
  c.w[0] = TNF_MUL(ax.w0, ay.w0);
  c.w[1] = TNF_MUL(ax.w1, ay.w0);
  c.w[2] = TNF_MUL(ax.w0, ay.w1);
  c.w[3] = TNF_MUL(ax.w1, ay.w1);
  return c;
}

__device__ __forceinline__ float4 blend(const float4 v[4], const Corners& c) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc.x = TNF_FMA(v[k].x, c.w[k], acc.x);
    acc.y = TNF_FMA(v[k].y, c.w[k], acc.y);
    acc.z = TNF_FMA(v[k].z, c.w[k], acc.z);
    acc.w = TNF_FMA(v[k].w, c.w[k], acc.w);
  }
  return acc;
}

__device__ __forceinline__ float4 mul4(float4 a, float4 b) {
  return make_float4(TNF_MUL(a.x, b.x), TNF_MUL(a.y, b.y), TNF_MUL(a.z, b.z), TNF_MUL(a.w, b.w));
}

// Planes of a scale use the coordinate pairs of itertools.combinations(range(3), 2) (src/models.py:145):
// plane 0 = (x,y), plane 1 = (x,z), plane 2 = (y,z); first coordinate -> W axis, second -> H axis.
// MINB: minimum resident blocks per SM asked of the compiler (register cap 65536 / (256 MINB)); 0 = no cap (44 registers
// forward, 66 backward = 5 / 3 blocks per SM).  Measured at the bench shape (scripts/time_kplanes.py, round 2): forward
// 110.6 us uncapped / 99.0 us at 6 blocks (40 registers) / 100.4 us at 8; backward 276.5 us uncapped / 260.1 us at 4 blocks
// (64 registers) / 279 at 5 / 292 at 6 / 381 at 8 (spills).  Defaults: forward 6, backward 4; tnf_set_variant(2, k) selects
// another build for diagnostics (k = blocks per SM, 1 = uncapped).
template <bool BWD, int MINB>
__global__ void __launch_bounds__(256, MINB) kplanes_kernel(const KPArgs A) {
  const int C = A.channels;
  const int lps = C >> 2;  // lanes per sample (power of two, <= 8)
  const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n = gt / lps;
  const int ch = (int)(gt % lps) * 4;
  if (n >= A.n) return;
  const float* xp = A.x + n * A.x_stride;
  const float cx = __ldg(xp), cy = __ldg(xp + 1), cz = __ldg(xp + 2);
  const int F = A.n_scales * C;
  // One scale per thread (blockIdx.y): three times as many, three times shorter dependent chains (coordinates ->
  // 12 corner lines -> blend) than a loop over the scales, and since the blocks of one scale are scheduled together the
  // live working set is that scale's three planes (6 / 25 / 100 MB for 128 / 256 / 512: each fits the 126 MB L2).
  {
    const int s = blockIdx.y + A.scale0;
    const int res = A.res[s];
    const Axis ax = axis_setup(cx, res), ay = axis_setup(cy, res), az = axis_setup(cz, res);
    Corners c[3];
    c[0] = corners(ax, ay, res, C, ch);
    c[1] = corners(ax, az, res, C, ch);
    c[2] = corners(ay, az, res, C, ch);
    float4 v[3][4];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const float* pl = A.planes[s * 3 + p];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[p][k] = __ldg(reinterpret_cast<const float4*>(pl + c[p].off[k]));
    }
    float4 f[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) f[p] = blend(v[p], c[p]);
    if (!BWD) {
      // current_scale_features = 1.; *= plane0; *= plane1; *= plane2  (src/models.py:158-160)
      const float4 o = mul4(mul4(f[0], f[1]), f[2]);
      *reinterpret_cast<float4*>(A.out + n * F + s * C + ch) = o;
    } else {
      const float4 g = __ldg(reinterpret_cast<const float4*>(A.grad_out + n * F + s * C + ch));
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
          const float wk = c[p].w[k];
          if (wk != 0.f)  // also skips the clamped out-of-range corners, like torch's safe_add_2d
            red_add_f4(gpl + c[p].off[k], make_float4(TNF_MUL(wk, gp[p].x), TNF_MUL(wk, gp[p].y),
                                                      TNF_MUL(wk, gp[p].z), TNF_MUL(wk, gp[p].w)));
        }
      }
    }
  }
}

// ---- sorted scatter for the coarse scales -------------------------------------------------------------------------------
// The scatter sits on the L2's reduction throughput (6.7 TB/s of red.v4 payload, profiles/r02_atomics_ubench.txt): 36 corner
// lines per sample whatever their order.  What CAN be removed is the multiplicity: a batch of 2^18 samples touches every
// texel of a 128^2 plane 65 times, of a 256^2 plane 16 times, of a 512^2 plane 4.6 times -- but only in each plane's OWN 2-D
// projection (3-D neighbours share almost nothing), and the three planes of a scale are coupled by the Hadamard product.
// So, for the scales below the finest:
//   sort    (tnf_kplanes_sort, on the prefetch stream when the batch is packed): for each of the three orientations a counting
//           sort of the samples by their 2x2-blocked cell index at the sorting resolution -> pos[o][n] (slot of sample n in
//           orientation o's order) and the orientation's two coordinates in slot order;
//   phase 1 (kplanes_bwd_p1_kernel): the gather + product of the plain kernel; the finest scale scatters directly as before,
//           the coarser scales WRITE their three per-plane gradient rows (128 B each) to slot pos[o][n] of a workspace;
//   phase 2 (kplanes_bwd_p2_kernel): per (scale, orientation) the rows are streamed in slot order; an 8-lane group walks 32
//           consecutive slots, accumulates the four corner sums in registers while the cell stays the same and issues one
//           red.v4 per corner per RUN instead of per sample (8-17x fewer at 128^2, ~4x at 256^2).
// MEASURED (scripts/time_kplanes_sorted.py, profiles/r02_kplanes_sorted.txt, 271,824 samples): direct scatter 270 us; sorted
// coarse scales: phase 1 219 us + phase 2 107 us = 320 us, plus a 75 us sort; all three scales sorted: 176 + 172 us, sort 216 us.
// The reductions do disappear, but the per-plane rows have to cross HBM twice (random 128-byte row stores cost phase 1
// ~80 us over the 98 us gather, and phase 2 would need >= 32-48 us even at the full HBM rate), which is what the merged
// reductions save.  Correct (tests/test_gpu_fused.py::test_sorted_plane_scatter_equals_direct_scatter) and kept opt-in
// (fused.py: TNF_KPLANES_SORTED=1); the direct scatter stays the default.
constexpr int kSortChunk = 32;   // slots walked by one lane group in phase 2

__device__ __forceinline__ int sort_key(float u, float v, int res) {
  // 2x2-blocked cell index of (u -> x/W axis, v -> y/H axis) at resolution `res` (axis_setup's floor, clamped)
  const float iu = TNF_MUL(TNF_MUL(TNF_ADD(u, 1.f), 0.5f), (float)(res - 1));
  const float iv = TNF_MUL(TNF_MUL(TNF_ADD(v, 1.f), 0.5f), (float)(res - 1));
  const int x0 = min(max((int)floorf(iu), 0), res - 1), y0 = min(max((int)floorf(iv), 0), res - 1);
  return (((y0 >> 1) * ((res + 1) >> 1) + (x0 >> 1)) << 2) | ((y0 & 1) << 1) | (x0 & 1);
}
__host__ __device__ __forceinline__ long long sort_keys(int res) { return 4LL * ((res + 1) >> 1) * ((res + 1) >> 1); }

__global__ void __launch_bounds__(256) ksort_hist_kernel(const float* __restrict__ x, long long xs, long long n, int res,
                                                         int* __restrict__ hist, int* __restrict__ keys) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c[3] = {__ldg(x + i * xs), __ldg(x + i * xs + 1), __ldg(x + i * xs + 2)};
  const long long K = sort_keys(res);
  const int k0 = sort_key(c[0], c[1], res), k1 = sort_key(c[0], c[2], res), k2 = sort_key(c[1], c[2], res);
  keys[i] = k0; keys[n + i] = k1; keys[2 * n + i] = k2;
  atomicAdd(hist + k0, 1); atomicAdd(hist + K + k1, 1); atomicAdd(hist + 2 * K + k2, 1);
}
// exclusive scan of each orientation's histogram, one CTA per orientation
__global__ void __launch_bounds__(1024) ksort_scan_kernel(int* __restrict__ hist, long long K) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  int* h = hist + blockIdx.x * K;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (long long chunk = 0; chunk < K; chunk += 1024 * 8) {
    const long long r0 = chunk + (long long)tid * 8;
    int c[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] = (r0 + i < K) ? h[r0 + i] : 0; sum += c[i]; }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kFullMask, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int v = s_warp[lane];
      int wi = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(kFullMask, wi, d); if (lane >= d) wi += u; }
      s_warp[lane] = wi - v;
    }
    __syncthreads();
    int run = s_base + s_warp[wid] + (incl - sum);
#pragma unroll
    for (int i = 0; i < 8; ++i) { if (r0 + i < K) h[r0 + i] = run; run += c[i]; }
    __syncthreads();
    if (tid == 1023) s_base = run;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) ksort_scatter_kernel(const float* __restrict__ x, long long xs, long long n, int res,
                                                            int* __restrict__ hist, const int* __restrict__ keys,
                                                            int* __restrict__ pos, float2* __restrict__ uv) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c[3] = {__ldg(x + i * xs), __ldg(x + i * xs + 1), __ldg(x + i * xs + 2)};
  const long long K = sort_keys(res);
  const int p0 = atomicAdd(hist + keys[i], 1), p1 = atomicAdd(hist + K + keys[n + i], 1), p2 = atomicAdd(hist + 2 * K + keys[2 * n + i], 1);
  pos[i] = p0; pos[n + i] = p1; pos[2 * n + i] = p2;
  uv[p0] = make_float2(c[0], c[1]); uv[n + p1] = make_float2(c[0], c[2]); uv[2 * n + p2] = make_float2(c[1], c[2]);
}

// phase 1: kplanes_kernel<true> with the scatter of the scales below `sorted_scales` replaced by a row store
__global__ void __launch_bounds__(256, 4) kplanes_bwd_p1_kernel(const KPArgs A, int sorted_scales, const int* __restrict__ pos,
                                                                float* __restrict__ rows) {
  const int C = A.channels;
  const int lps = C >> 2;
  const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n = gt / lps;
  const int ch = (int)(gt % lps) * 4;
  if (n >= A.n) return;
  const float* xp = A.x + n * A.x_stride;
  const float cx = __ldg(xp), cy = __ldg(xp + 1), cz = __ldg(xp + 2);
  const int F = A.n_scales * C;
  const int s = blockIdx.y + A.scale0;
  const int res = A.res[s];
  const Axis ax = axis_setup(cx, res), ay = axis_setup(cy, res), az = axis_setup(cz, res);
  Corners c[3];
  c[0] = corners(ax, ay, res, C, ch);
  c[1] = corners(ax, az, res, C, ch);
  c[2] = corners(ay, az, res, C, ch);
  float4 v[3][4];
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const float* pl = A.planes[s * 3 + p];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[p][k] = __ldg(reinterpret_cast<const float4*>(pl + c[p].off[k]));
  }
  float4 f[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) f[p] = blend(v[p], c[p]);
  const float4 g = __ldg(reinterpret_cast<const float4*>(A.grad_out + n * F + s * C + ch));
  const float4 g01 = mul4(g, f[2]);
  float4 gp[3];
  gp[0] = mul4(g01, f[1]);
  gp[1] = mul4(g01, f[0]);
  gp[2] = mul4(g, mul4(f[0], f[1]));
  if (s < sorted_scales) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const long long slot = __ldg(pos + p * A.n + n);
      st_stream_f4(rows + ((long long)(s * 3 + p) * A.n + slot) * C + ch, gp[p]);
    }
  } else {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float* gpl = A.grads[s * 3 + p];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float wk = c[p].w[k];
        if (wk != 0.f)
          red_add_f4(gpl + c[p].off[k], make_float4(TNF_MUL(wk, gp[p].x), TNF_MUL(wk, gp[p].y), TNF_MUL(wk, gp[p].z), TNF_MUL(wk, gp[p].w)));
      }
    }
  }
}

// phase 2: blockIdx.y = (scale - scale0) * 3 + orientation; an 8-lane group (C = 32) walks kSortChunk consecutive slots
__global__ void __launch_bounds__(256) kplanes_bwd_p2_kernel(const KPArgs A, const float2* __restrict__ uv,
                                                             const float* __restrict__ rows) {
  const int C = A.channels;
  const int lps = C >> 2;
  const int s = A.scale0 + blockIdx.y / 3, p = blockIdx.y % 3;
  const int res = A.res[s];
  const long long grp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / lps;
  const int ch = (int)(threadIdx.x % lps) * 4;
  const long long i0 = grp * kSortChunk;
  if (i0 >= A.n) return;
  const long long i1 = (i0 + kSortChunk < A.n) ? i0 + kSortChunk : A.n;
  const float2* q = uv + (long long)p * A.n;
  const float* src = rows + ((long long)(s * 3 + p) * A.n) * C + ch;
  float* gpl = A.grads[s * 3 + p];
  float4 acc[4];
  int cur_off[4] = {-1, -1, -1, -1};
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto flush = [&]() {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (acc[k].x != 0.f || acc[k].y != 0.f || acc[k].z != 0.f || acc[k].w != 0.f) red_add_f4(gpl + cur_off[k], acc[k]);
      acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float2 t = __ldg(q + i0);
  float4 r = ld_stream_f4(src + i0 * C);
  for (long long i = i0; i < i1; ++i) {
    const float2 tc = t;
    const float4 rc = r;
    if (i + 1 < i1) { t = __ldg(q + i + 1); r = ld_stream_f4(src + (i + 1) * C); }   // next slot in flight
    const Axis ax = axis_setup(tc.x, res), ay = axis_setup(tc.y, res);
    const Corners c = corners(ax, ay, res, C, ch);
    if (c.off[0] != cur_off[0] || c.off[3] != cur_off[3]) {
      if (cur_off[0] >= 0) flush();
#pragma unroll
      for (int k = 0; k < 4; ++k) cur_off[k] = c.off[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float wk = c.w[k];
      acc[k].x = TNF_ADD(acc[k].x, TNF_MUL(wk, rc.x)); acc[k].y = TNF_ADD(acc[k].y, TNF_MUL(wk, rc.y));
      acc[k].z = TNF_ADD(acc[k].z, TNF_MUL(wk, rc.z)); acc[k].w = TNF_ADD(acc[k].w, TNF_MUL(wk, rc.w));
    }
  }
  if (cur_off[0] >= 0) flush();
}

// ---- experimental variants of the scatter (scripts/time_kplanes.py; round-2 exploration) ----
// MODE 1: gather + blend only (no reductions)   MODE 2: reductions only (no plane reads)
// MODE 3: reductions issued by the TMA engine: the weighted rows are staged in shared memory and one
//         cp.reduce.async.bulk (add.f32) per pair of x-adjacent corners (256 B) replaces 16 red.v4 lane operations.
__device__ __forceinline__ void bulk_red_add(float* gdst, const float* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes)
               : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(256) kplanes_bwd_ex_kernel(const KPArgs A) {
  extern __shared__ __align__(128) float stage[];
  const int C = A.channels;
  const int lps = C >> 2;
  const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long n = gt / lps;
  const int ch = (int)(gt % lps) * 4;
  const bool live = n < A.n;
  if (!live) n = A.n - 1;
  const float* xp = A.x + n * A.x_stride;
  const float cx = __ldg(xp), cy = __ldg(xp + 1), cz = __ldg(xp + 2);
  const int F = A.n_scales * C;
  const int s = blockIdx.y + A.scale0;
  const int res = A.res[s];
  const Axis ax = axis_setup(cx, res), ay = axis_setup(cy, res), az = axis_setup(cz, res);
  Corners c[3];
  c[0] = corners(ax, ay, res, C, ch);
  c[1] = corners(ax, az, res, C, ch);
  c[2] = corners(ay, az, res, C, ch);
  float4 f[3];
  if (MODE != 2) {
    float4 v[3][4];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const float* pl = A.planes[s * 3 + p];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[p][k] = __ldg(reinterpret_cast<const float4*>(pl + c[p].off[k]));
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) f[p] = blend(v[p], c[p]);
  } else {
    f[0] = make_float4(cx, cy, cz, cx); f[1] = make_float4(cy, cz, cx, cy); f[2] = make_float4(cz, cx, cy, cz);
  }
  const float4 g = __ldg(reinterpret_cast<const float4*>(A.grad_out + n * F + s * C + ch));
  const float4 g01 = mul4(g, f[2]);
  float4 gp[3];
  gp[0] = mul4(g01, f[1]);
  gp[1] = mul4(g01, f[0]);
  gp[2] = mul4(g, mul4(f[0], f[1]));
  if (MODE == 1) {
    const float t = gp[0].x + gp[1].y + gp[2].z + gp[0].w;
    if (t == 1.2345678e-30f && live) A.grads[0][0] = t;
    return;
  }
  if (MODE == 2) {
    if (!live) return;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float* gpl = A.grads[s * 3 + p];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float wk = c[p].w[k];
        if (wk != 0.f)
          red_add_f4(gpl + c[p].off[k], make_float4(TNF_MUL(wk, gp[p].x), TNF_MUL(wk, gp[p].y), TNF_MUL(wk, gp[p].z), TNF_MUL(wk, gp[p].w)));
      }
    }
    return;
  }
  if (MODE == 3) {
    // group of lps lanes = one sample; row (p, j) of the group's staging area = [x0 texel C floats | x1 texel C floats]
    const int grp = threadIdx.x / lps, l = threadIdx.x % lps;
    float* base = stage + (size_t)grp * 6 * 2 * C;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float wa = c[p].w[2 * j], wb = c[p].w[2 * j + 1];
        float* row = base + (p * 2 + j) * 2 * C;
        *reinterpret_cast<float4*>(row + l * 4) = make_float4(TNF_MUL(wa, gp[p].x), TNF_MUL(wa, gp[p].y), TNF_MUL(wa, gp[p].z), TNF_MUL(wa, gp[p].w));
        *reinterpret_cast<float4*>(row + C + l * 4) = make_float4(TNF_MUL(wb, gp[p].x), TNF_MUL(wb, gp[p].y), TNF_MUL(wb, gp[p].z), TNF_MUL(wb, gp[p].w));
      }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (live) {
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (l == p * 2 + j) {
            float* gpl = A.grads[s * 3 + p];
            const float* row = base + (p * 2 + j) * 2 * C;
            const int o0 = c[p].off[2 * j] - ch, o1 = c[p].off[2 * j + 1] - ch;
            const float wa = c[p].w[2 * j], wb = c[p].w[2 * j + 1];
            if (wa != 0.f && wb != 0.f && o1 == o0 + C) {
              bulk_red_add(gpl + o0, row, 8u * C);
            } else {
              if (wa != 0.f) bulk_red_add(gpl + o0, row, 4u * C);
              if (wb != 0.f) bulk_red_add(gpl + o1, row + C, 4u * C);
            }
          }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
  const dim3 grid((unsigned)ceil_div(threads, 256), (unsigned)n_scales);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (variant(kVariantKplanesOcc)) {
    case 1: kplanes_kernel<false, 0><<<grid, 256, 0, st>>>(A); break;
    case 8: kplanes_kernel<false, 8><<<grid, 256, 0, st>>>(A); break;
    default: kplanes_kernel<false, 6><<<grid, 256, 0, st>>>(A);
  }
  TNF_LAUNCH_CHECK("kplanes_fwd_kernel");
  return TNF_OK;
}

extern "C" int tnf_kplanes_bwd(const float* const* planes, float* const* grad_planes, const int32_t* res,
                               int32_t n_scales, int32_t channels, const float* x, int64_t x_stride,
                               int64_t n, const float* grad_out, void* stream) {
  return tnf_kplanes_bwd_scales(planes, grad_planes, res, n_scales, channels, x, x_stride, n, grad_out, 0, n_scales, stream);
}

extern "C" int tnf_kplanes_bwd_scales(const float* const* planes, float* const* grad_planes, const int32_t* res,
                                      int32_t n_scales, int32_t channels, const float* x, int64_t x_stride, int64_t n,
                                      const float* grad_out, int32_t scale_begin, int32_t scale_end, void* stream) {
  using namespace tnf;
  KPArgs A{};
  TNF_REQUIRE(grad_planes, "null grad plane table");
  TNF_REQUIRE(scale_begin >= 0 && scale_begin <= scale_end && scale_end <= n_scales, "bad scale range [%d,%d)", scale_begin, scale_end);
  if (scale_begin == scale_end) return TNF_OK;
  int rc = fill_args(&A, planes, grad_planes, res, n_scales, channels, x, x_stride, n);
  if (rc != TNF_OK || n == 0) return rc;
  TNF_REQUIRE(grad_out && (reinterpret_cast<uintptr_t>(grad_out) & 15u) == 0, "grad_out null/misaligned");
  A.grad_out = grad_out;
  A.scale0 = scale_begin;
  const long long threads = n * (channels / 4);
  const dim3 grid((unsigned)ceil_div(threads, 256), (unsigned)(scale_end - scale_begin));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (variant(kVariantKplanesOcc)) {
    case 1: kplanes_kernel<true, 0><<<grid, 256, 0, st>>>(A); break;
    case 5: kplanes_kernel<true, 5><<<grid, 256, 0, st>>>(A); break;
    case 6: kplanes_kernel<true, 6><<<grid, 256, 0, st>>>(A); break;
    default: kplanes_kernel<true, 4><<<grid, 256, 0, st>>>(A);
  }
  TNF_LAUNCH_CHECK("kplanes_bwd_kernel");
  return TNF_OK;
}

// Experimental scatter variants (see kplanes_bwd_ex_kernel); not part of the reference-facing surface.
extern "C" int tnf_kplanes_bwd_ex(const float* const* planes, float* const* grad_planes, const int32_t* res,
                                  int32_t n_scales, int32_t channels, const float* x, int64_t x_stride, int64_t n,
                                  const float* grad_out, int32_t mode, void* stream) {
  using namespace tnf;
  KPArgs A{};
  int rc = fill_args(&A, planes, grad_planes, res, n_scales, channels, x, x_stride, n);
  if (rc != TNF_OK || n == 0) return rc;
  A.grad_out = grad_out;
  const long long threads = n * (channels / 4);
  const dim3 grid((unsigned)ceil_div(threads, 256), (unsigned)n_scales);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)(256 / (channels / 4)) * 6 * 2 * channels * sizeof(float);
  if (mode == 1) kplanes_bwd_ex_kernel<1><<<grid, 256, 0, st>>>(A);
  else if (mode == 2) kplanes_bwd_ex_kernel<2><<<grid, 256, 0, st>>>(A);
  else if (mode == 3) {
    cudaFuncSetAttribute(kplanes_bwd_ex_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kplanes_bwd_ex_kernel<3><<<grid, 256, smem, st>>>(A);
  } else TNF_REQUIRE(false, "mode must be 1, 2 or 3");
  TNF_LAUNCH_CHECK("kplanes_bwd_ex_kernel");
  return TNF_OK;
}

// ---- sorted scatter (see kplanes_bwd_p1_kernel / _p2_kernel) ----
extern "C" int64_t tnf_kplanes_sort_scratch_ints(int32_t sort_res, int64_t n) {
  return (sort_res >= 2 && n >= 0) ? 3 * tnf::sort_keys(sort_res) + 3 * n : 0;
}

extern "C" int tnf_kplanes_sort(const float* x, int64_t x_stride, int64_t n, int32_t sort_res, int32_t* scratch, int32_t* pos,
                                float* uv, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0 && n < (1LL << 31) && x_stride >= 3 && sort_res >= 2 && sort_res <= 2048, "bad sizes");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(x && scratch && pos && uv, "null pointer");
  TNF_REQUIRE((reinterpret_cast<uintptr_t>(uv) & 7u) == 0, "uv must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long K = sort_keys(sort_res);
  int* hist = scratch;
  int* keys = scratch + 3 * K;
  TNF_CUDA(cudaMemsetAsync(hist, 0, sizeof(int) * 3 * K, st));
  const unsigned blocks = (unsigned)ceil_div(n, 256);
  ksort_hist_kernel<<<blocks, 256, 0, st>>>(x, x_stride, n, sort_res, hist, keys);
  ksort_scan_kernel<<<3, 1024, 0, st>>>(hist, K);
  ksort_scatter_kernel<<<blocks, 256, 0, st>>>(x, x_stride, n, sort_res, hist, keys, pos, reinterpret_cast<float2*>(uv));
  TNF_LAUNCH_CHECK("kplanes_sort kernels");
  return TNF_OK;
}

extern "C" int tnf_kplanes_bwd_sorted(const float* const* planes, float* const* grad_planes, const int32_t* res, int32_t n_scales,
                                      int32_t channels, const float* x, int64_t x_stride, int64_t n, const float* grad_out,
                                      int32_t sorted_scales, const int32_t* pos, const float* uv, float* rows, int32_t phase,
                                      void* stream) {
  using namespace tnf;
  KPArgs A{};
  TNF_REQUIRE(grad_planes, "null grad plane table");
  TNF_REQUIRE(sorted_scales >= 0 && sorted_scales <= n_scales, "bad sorted_scales");
  TNF_REQUIRE(phase >= 0 && phase <= 2, "phase must be 0 (both), 1 or 2");
  int rc = fill_args(&A, planes, grad_planes, res, n_scales, channels, x, x_stride, n);
  if (rc != TNF_OK || n == 0) return rc;
  TNF_REQUIRE(channels == 32, "the sorted scatter is implemented for 32-channel planes");
  TNF_REQUIRE(sorted_scales == 0 || (pos && uv && rows), "null sort tables / row workspace");
  TNF_REQUIRE((reinterpret_cast<uintptr_t>(rows) & 15u) == 0 && (reinterpret_cast<uintptr_t>(uv) & 7u) == 0, "rows / uv misaligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long threads = n * (channels / 4);
  if (phase != 2) {
    TNF_REQUIRE(grad_out && (reinterpret_cast<uintptr_t>(grad_out) & 15u) == 0, "grad_out null/misaligned");
    A.grad_out = grad_out;
    A.scale0 = 0;
    kplanes_bwd_p1_kernel<<<dim3((unsigned)ceil_div(threads, 256), (unsigned)n_scales), 256, 0, st>>>(A, sorted_scales, pos, rows);
    TNF_LAUNCH_CHECK("kplanes_bwd_p1_kernel");
  }
  if (phase != 1 && sorted_scales > 0) {
    A.scale0 = 0;
    const long long groups = ceil_div(n, kSortChunk);
    kplanes_bwd_p2_kernel<<<dim3((unsigned)ceil_div(groups * (channels / 4), 256), (unsigned)(3 * sorted_scales)), 256, 0, st>>>(
        A, reinterpret_cast<const float2*>(uv), rows);
    TNF_LAUNCH_CHECK("kplanes_bwd_p2_kernel");
  }
  return TNF_OK;
}
