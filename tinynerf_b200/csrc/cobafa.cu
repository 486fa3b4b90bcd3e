// a14 -- Cobafa fused basis/coefficient lookup (reference: CobafaFeatureField.forward
// src/models.py:258-264, CobafaGrid.forward :228-238, SawtoothEncoding :209-214).
//
// Reference: 7 five-dimensional grid_sample launches over NCDHW grids + 6 broadcasts-multiplies +
// concat.  Here grids are channels-last ([r][r][r][c]); 8 lanes cooperate on one sample, lane l
// owning level l: it interpolates coefficient channel l from the coef grid and the c_l basis
// channels of level l at the sawtooth-encoded position, multiplies, and writes its slice of the
// concatenated [N, sum(c_l)] row.  Backward recomputes the forward values and scatters with
// reductions into the (channels-last) gradient grids.
#include "common.cuh"
#include "nerf_math.cuh"

namespace tnf {
namespace {

constexpr int kMaxLevels = 8;
constexpr int kMaxCh = 8;

struct CBArgs {
  const float* basis[kMaxLevels];
  float* gbasis[kMaxLevels];
  int res[kMaxLevels];
  int ch[kMaxLevels];
  int off[kMaxLevels];
  float freq[kMaxLevels];
  int n_levels;
  int feat;  // sum of channels
  const float* coef;
  float* gcoef;
  int coef_res;
  const float* x;
  long long x_stride;
  long long n;
  float* out;
  const float* grad_out;
};

struct Trilinear {
  int x0, y0, z0;
  float w[8];  // tnw,tne,tsw,tse,bnw,bne,bsw,bse
  bool ok[8];
};

__device__ __forceinline__ Trilinear trilinear_setup(float gx, float gy, float gz, int D, int H, int W) {
  Trilinear t;
  const float ix = TNF_MUL(TNF_MUL(TNF_ADD(gx, 1.f), 0.5f), (float)(W - 1));
  const float iy = TNF_MUL(TNF_MUL(TNF_ADD(gy, 1.f), 0.5f), (float)(H - 1));
  const float iz = TNF_MUL(TNF_MUL(TNF_ADD(gz, 1.f), 0.5f), (float)(D - 1));
  t.x0 = (int)floorf(ix);
  t.y0 = (int)floorf(iy);
  t.z0 = (int)floorf(iz);
  const float wx[2] = {TNF_SUB((float)(t.x0 + 1), ix), TNF_SUB(ix, (float)t.x0)};
  const float wy[2] = {TNF_SUB((float)(t.y0 + 1), iy), TNF_SUB(iy, (float)t.y0)};
  const float wz[2] = {TNF_SUB((float)(t.z0 + 1), iz), TNF_SUB(iz, (float)t.z0)};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    t.w[k] = TNF_MUL(TNF_MUL(wx[dx], wy[dy]), wz[dz]);
    t.ok[k] = ((unsigned)(t.x0 + dx) < (unsigned)W) && ((unsigned)(t.y0 + dy) < (unsigned)H) &&
              ((unsigned)(t.z0 + dz) < (unsigned)D);
  }
  return t;
}
__device__ __forceinline__ long long corner3(const Trilinear& t, int k, int H, int W, int C) {
  const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
  return (((long long)(t.z0 + dz) * H + (t.y0 + dy)) * W + (t.x0 + dx)) * C;
}

// torch.remainder(v, 1.) for floats: fmod then sign fix-up (floor-mod)
__device__ __forceinline__ float remainder1(float v) {
  float m = fmodf(v, 1.f);
  if (m != 0.f && m < 0.f) m = TNF_ADD(m, 1.f);
  return m;
}

template <bool BWD>
__global__ void __launch_bounds__(256) cobafa_kernel(const CBArgs A) {
  const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n = gt >> 3;
  const int l = (int)(gt & 7);
  if (n >= A.n || l >= A.n_levels) return;
  float xyz[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) xyz[c] = __ldg(A.x + n * A.x_stride + c);

  // coefficient channel l (coef grid is [r][r][r][L])
  const int cr = A.coef_res, L = A.n_levels;
  const Trilinear tc = trilinear_setup(xyz[0], xyz[1], xyz[2], cr, cr, cr);
  float coef = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (tc.ok[k]) coef = TNF_FMA(__ldg(A.coef + corner3(tc, k, cr, cr, L) + l), tc.w[k], coef);

  // basis level l at the sawtooth-encoded position: 2*((f*x) % 1) - 1
  float e[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) e[c] = TNF_SUB(TNF_MUL(2.f, remainder1(TNF_MUL(A.freq[l], xyz[c]))), 1.f);
  const int r = A.res[l], C = A.ch[l];
  const Trilinear tb = trilinear_setup(e[0], e[1], e[2], r, r, r);
  const float* bg = A.basis[l];
  float val[kMaxCh];
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) val[c] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (!tb.ok[k]) continue;
    const float* p = bg + corner3(tb, k, r, r, C);
    if ((C & 3) == 0) {
#pragma unroll
      for (int c4 = 0; c4 < kMaxCh / 4; ++c4) {
        if (c4 * 4 < C) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p) + c4);
          val[c4 * 4 + 0] = TNF_FMA(v.x, tb.w[k], val[c4 * 4 + 0]);
          val[c4 * 4 + 1] = TNF_FMA(v.y, tb.w[k], val[c4 * 4 + 1]);
          val[c4 * 4 + 2] = TNF_FMA(v.z, tb.w[k], val[c4 * 4 + 2]);
          val[c4 * 4 + 3] = TNF_FMA(v.w, tb.w[k], val[c4 * 4 + 3]);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < kMaxCh; ++c)
        if (c < C) val[c] = TNF_FMA(__ldg(p + c), tb.w[k], val[c]);
    }
  }

  if (!BWD) {
    float* o = A.out + n * A.feat + A.off[l];
#pragma unroll
    for (int c = 0; c < kMaxCh; ++c)
      if (c < C) o[c] = TNF_MUL(val[c], coef);
  } else {
    const float* g = A.grad_out + n * A.feat + A.off[l];
    float gv[kMaxCh];
    float gc = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxCh; ++c) {
      gv[c] = (c < C) ? __ldg(g + c) : 0.f;
      if (c < C) gc = TNF_ADD(gc, TNF_MUL(gv[c], val[c]));  // d coef_l = sum_c g_c * basis_c
    }
    float* gb = A.gbasis[l];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (!tb.ok[k]) continue;
      float* p = gb + corner3(tb, k, r, r, C);
      if ((C & 3) == 0) {
#pragma unroll
        for (int c4 = 0; c4 < kMaxCh / 4; ++c4) {
          if (c4 * 4 < C)
            red_add_f4(p + c4 * 4, make_float4(TNF_MUL(tb.w[k], TNF_MUL(gv[c4 * 4 + 0], coef)),
                                               TNF_MUL(tb.w[k], TNF_MUL(gv[c4 * 4 + 1], coef)),
                                               TNF_MUL(tb.w[k], TNF_MUL(gv[c4 * 4 + 2], coef)),
                                               TNF_MUL(tb.w[k], TNF_MUL(gv[c4 * 4 + 3], coef))));
        }
      } else {
#pragma unroll
        for (int c = 0; c < kMaxCh; ++c)
          if (c < C) atomicAdd(p + c, TNF_MUL(tb.w[k], TNF_MUL(gv[c], coef)));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (tc.ok[k]) atomicAdd(A.gcoef + corner3(tc, k, cr, cr, L) + l, TNF_MUL(tc.w[k], gc));
  }
}

int fill(CBArgs* A, const float* const* basis, float* const* gbasis, const int32_t* res, const int32_t* ch,
         const float* freqs, int n_levels, const float* coef, float* gcoef, int coef_res, const float* x,
         int64_t x_stride, int64_t n, bool bwd) {
  TNF_REQUIRE(n >= 0, "negative n");
  TNF_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels, "n_levels must be in [1,%d]", kMaxLevels);
  TNF_REQUIRE(basis && res && ch && freqs && coef && coef_res >= 2, "null/invalid grid description");
  TNF_REQUIRE(!bwd || (gbasis && gcoef), "null gradient grids");
  TNF_REQUIRE(n == 0 || x, "null x");
  TNF_REQUIRE(x_stride >= 3, "x_stride must be >= 3");
  int off = 0;
  for (int l = 0; l < n_levels; ++l) {
    TNF_REQUIRE(ch[l] >= 1 && ch[l] <= kMaxCh, "level %d: channels must be in [1,%d]", l, kMaxCh);
    TNF_REQUIRE(res[l] >= 2, "level %d: resolution must be >= 2", l);
    TNF_REQUIRE(basis[l] && (reinterpret_cast<uintptr_t>(basis[l]) & 15u) == 0, "basis %d null/misaligned", l);
    A->basis[l] = basis[l];
    if (bwd) {
      TNF_REQUIRE(gbasis[l] && (reinterpret_cast<uintptr_t>(gbasis[l]) & 15u) == 0, "grad basis %d null/misaligned", l);
      A->gbasis[l] = gbasis[l];
    }
    A->res[l] = res[l];
    A->ch[l] = ch[l];
    A->off[l] = off;
    A->freq[l] = freqs[l];
    off += ch[l];
  }
  A->n_levels = n_levels;
  A->feat = off;
  A->coef = coef;
  A->gcoef = gcoef;
  A->coef_res = coef_res;
  A->x = x;
  A->x_stride = x_stride;
  A->n = n;
  return TNF_OK;
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_cobafa_fwd(const float* const* basis, const int32_t* basis_res, const int32_t* basis_ch,
                              const float* freqs, int32_t n_levels, const float* coef, int32_t coef_res,
                              const float* x, int64_t x_stride, int64_t n, float* out, void* stream) {
  using namespace tnf;
  CBArgs A{};
  int rc = fill(&A, basis, nullptr, basis_res, basis_ch, freqs, n_levels, coef, nullptr, coef_res, x, x_stride,
                n, false);
  if (rc != TNF_OK || n == 0) return rc;
  TNF_REQUIRE(out, "null out");
  A.out = out;
  cobafa_kernel<false><<<(unsigned)ceil_div(n * 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  TNF_LAUNCH_CHECK("cobafa_fwd_kernel");
  return TNF_OK;
}

extern "C" int tnf_cobafa_bwd(const float* const* basis, float* const* grad_basis, const int32_t* basis_res,
                              const int32_t* basis_ch, const float* freqs, int32_t n_levels, const float* coef,
                              float* grad_coef, int32_t coef_res, const float* x, int64_t x_stride, int64_t n,
                              const float* grad_out, void* stream) {
  using namespace tnf;
  CBArgs A{};
  int rc = fill(&A, basis, grad_basis, basis_res, basis_ch, freqs, n_levels, coef, grad_coef, coef_res, x,
                x_stride, n, true);
  if (rc != TNF_OK || n == 0) return rc;
  TNF_REQUIRE(grad_out, "null grad_out");
  A.grad_out = grad_out;
  cobafa_kernel<true><<<(unsigned)ceil_div(n * 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  TNF_LAUNCH_CHECK("cobafa_bwd_kernel");
  return TNF_OK;
}
