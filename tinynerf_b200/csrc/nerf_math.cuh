// Per-sample arithmetic of the march/contract/occupancy path, written so that every operation is
// rounded exactly where the reference's un-fused PyTorch elementwise kernels round it
// (src/core.py:16-31, 73-88, 165-176) -- no implicit FMA contraction -- and so that the trilinear
// lookup reproduces torch's CUDA grid_sampler_3d arithmetic (ATen/native/cuda/GridSampler.cuh:21-31
// unnormalize; corner weights as products of differences; 8 fused multiply-adds in
// tnw,tne,tsw,tse,bnw,bne,bsw,bse order; out-of-bounds corners skipped).
//
// The functions are __host__ __device__ so tests/host_check.cu can run them on the CPU of the build
// container (compiled with -ffp-contract=off) before any GPU time is spent.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TNF_HD __host__ __device__ __forceinline__
#else
#define TNF_HD inline
#endif

#ifdef __CUDA_ARCH__
#define TNF_ADD(a, b) __fadd_rn((a), (b))
#define TNF_SUB(a, b) __fsub_rn((a), (b))
#define TNF_MUL(a, b) __fmul_rn((a), (b))
#define TNF_DIV(a, b) __fdiv_rn((a), (b))
#define TNF_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define TNF_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define TNF_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define TNF_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define TNF_DIV(a, b) ((float)((float)(a) / (float)(b)))
#define TNF_FMA(a, b, c) fmaf((a), (b), (c))
#endif

namespace tnf {

struct MarchConst {
  int scene, n_steps;
  float a0[3], a1[3], ext[3];  // aabb min, max, (max-min)
  float near_, far_, step_size;
  const float* t_table;
  const float* step_table;
  const float* grid;
  int gd, gh, gw;
  float thr;
  const float* thr_dev;   // optional device scalar overriding thr (keeps the occupancy update free of host syncs)
  const float* noise;
  int jitter;
  unsigned long long seed, offset;
  const uint32_t* empty_bits;   // optional classifier bitfields (tnf_occ_build_empty_bits), see occ_certainly_empty
};

// NaN-propagating min/max/clamp like torch.amin/amax/clamp
TNF_HD float nmin(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
TNF_HD float nmax(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

// RayMarcherAABB.__call__ (src/core.py:73-81): slab entry distance.
TNF_HD float aabb_t_min(const float o[3], const float d[3], const MarchConst& M) {
  float tmin = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float den = (d[c] == 0.f) ? TNF_ADD(d[c], 1e-9f) : d[c];
    const float i0 = TNF_DIV(TNF_SUB(M.a0[c], o[c]), den);
    const float i1 = TNF_DIV(TNF_SUB(M.a1[c], o[c]), den);
    const float m = nmin(i0, i1);
    tmin = (c == 0) ? m : nmax(tmin, m);
  }
  // torch.clamp(min=near, max=far): min(max(x, near), far), NaN stays NaN
  if (tmin == tmin) {
    tmin = tmin < M.near_ ? M.near_ : tmin;
    tmin = tmin > M.far_ ? M.far_ : tmin;
  }
  return tmin;
}

// grid_sample 5-D, bilinear, zeros padding, align_corners=True on a [D][H][W] fp32 grid.
// c[0] -> W, c[1] -> H, c[2] -> D.
TNF_HD float trilinear_zeros(const float* __restrict__ g, int D, int H, int W, float x, float y, float z) {
  const float ix = TNF_MUL(TNF_MUL(TNF_ADD(x, 1.f), 0.5f), (float)(W - 1));
  const float iy = TNF_MUL(TNF_MUL(TNF_ADD(y, 1.f), 0.5f), (float)(H - 1));
  const float iz = TNF_MUL(TNF_MUL(TNF_ADD(z, 1.f), 0.5f), (float)(D - 1));
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const int x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
  // "surfaces to each neighbour": (far corner - i) / (i - near corner), ints converted to float
  const float wx0 = TNF_SUB((float)x1, ix), wx1 = TNF_SUB(ix, (float)x0);
  const float wy0 = TNF_SUB((float)y1, iy), wy1 = TNF_SUB(iy, (float)y0);
  const float wz0 = TNF_SUB((float)z1, iz), wz1 = TNF_SUB(iz, (float)z0);
  const float tnw = TNF_MUL(TNF_MUL(wx0, wy0), wz0);
  const float tne = TNF_MUL(TNF_MUL(wx1, wy0), wz0);
  const float tsw = TNF_MUL(TNF_MUL(wx0, wy1), wz0);
  const float tse = TNF_MUL(TNF_MUL(wx1, wy1), wz0);
  const float bnw = TNF_MUL(TNF_MUL(wx0, wy0), wz1);
  const float bne = TNF_MUL(TNF_MUL(wx1, wy0), wz1);
  const float bsw = TNF_MUL(TNF_MUL(wx0, wy1), wz1);
  const float bse = TNF_MUL(TNF_MUL(wx1, wy1), wz1);
  const bool bx0 = (unsigned)x0 < (unsigned)W, bx1 = (unsigned)x1 < (unsigned)W;
  const bool by0 = (unsigned)y0 < (unsigned)H, by1 = (unsigned)y1 < (unsigned)H;
  const bool bz0 = (unsigned)z0 < (unsigned)D, bz1 = (unsigned)z1 < (unsigned)D;
  const long long sH = W, sD = (long long)W * H;
  float acc = 0.f;
  if (bz0 && by0 && bx0) acc = TNF_FMA(g[z0 * sD + y0 * sH + x0], tnw, acc);
  if (bz0 && by0 && bx1) acc = TNF_FMA(g[z0 * sD + y0 * sH + x1], tne, acc);
  if (bz0 && by1 && bx0) acc = TNF_FMA(g[z0 * sD + y1 * sH + x0], tsw, acc);
  if (bz0 && by1 && bx1) acc = TNF_FMA(g[z0 * sD + y1 * sH + x1], tse, acc);
  if (bz1 && by0 && bx0) acc = TNF_FMA(g[z1 * sD + y0 * sH + x0], bnw, acc);
  if (bz1 && by0 && bx1) acc = TNF_FMA(g[z1 * sD + y0 * sH + x1], bne, acc);
  if (bz1 && by1 && bx0) acc = TNF_FMA(g[z1 * sD + y1 * sH + x0], bsw, acc);
  if (bz1 && by1 && bx1) acc = TNF_FMA(g[z1 * sD + y1 * sH + x1], bse, acc);
  return acc;
}

// Contraction of one world-space point (src/core.py:16-31): AABB -> affine map to [-1,1] + inside-box flag,
// unbounded -> Mip-NeRF 360 contraction with the infinity norm (no mask: returns true).
TNF_HD bool contract_point(const MarchConst& M, const float p[3], float q[3]) {
  bool inside = true;
  if (M.scene == 0) {  // ContractionAABB (src/core.py:29-30)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      inside = inside && (p[c] >= M.a0[c]) && (p[c] <= M.a1[c]);
      q[c] = TNF_SUB(TNF_MUL(TNF_DIV(TNF_SUB(p[c], M.a0[c]), M.ext[c]), 2.f), 1.f);
    }
  } else {             // ContractionMip360, order=inf (src/core.py:18-19)
    const float n = fmaxf(fmaxf(fabsf(p[0]), fabsf(p[1])), fabsf(p[2]));
    if (n <= 1.f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) q[c] = TNF_MUL(p[c], 0.5f);
    } else {
      const float k = TNF_SUB(2.f, TNF_MUL(TNF_DIV(1.f, n), 1.f));  // 2. - norm.reciprocal()*1.
#pragma unroll
      for (int c = 0; c < 3; ++c) q[c] = TNF_MUL(TNF_DIV(TNF_MUL(k, p[c]), n), 0.5f);
    }
  }
  return inside;
}

// Occupancy classifier, two levels, one buffer (tnf_occ_build_empty_bits):
//   coarse  bit (bz, by, bx): EVERY lattice value a lookup in the 2x2x2-cell block can touch (corners [2b, 2b+2] per axis)
//           is <= threshold * (1 - 1e-5).  [ceil(D/2)][ceil(H/2)][ceil(ceil(W/2)/32)] words, 32 KB for 128^3: the march
//           kernel stages this level in SHARED memory.
//   fine    bit (z, y, x): the 8 corners of cell (x, y, z) are all <= threshold * (1 - 1e-5).  [D][H][ceil(W/32)] words
//           after the coarse level (256 KB for 128^3: L1/L2-resident, one 4-byte load).
// A lookup's result is a convex combination of its 8 corners (weights >= 0, their fp32 sum <= 1 + 5e-7, 8 fused
// multiply-adds: <= 1e-6 relative excess in all), so where a bit is set `trilinear(grid) > threshold` is false whatever
// the weights -- exactly what the float path would return.  Isolated occupied cells ("floaters") defeat the coarse level
// (27 corners) more often than the fine one (8 corners).
TNF_HD int occ_bits_words_per_row(int W) { return (((W + 1) >> 1) + 31) >> 5; }
TNF_HD long long occ_bits_coarse_words(int D, int H, int W) {
  return (long long)((D + 1) >> 1) * ((H + 1) >> 1) * occ_bits_words_per_row(W);
}
TNF_HD long long occ_bits_words(int D, int H, int W) {
  return occ_bits_coarse_words(D, H, W) + (long long)D * H * ((W + 31) >> 5);
}
// coarse: level-1 words (shared or global memory); fine: level-2 words (global) or nullptr
TNF_HD bool occ_certainly_empty(const uint32_t* coarse, const uint32_t* fine, int D, int H, int W, float x, float y, float z) {
  const float ix = TNF_MUL(TNF_MUL(TNF_ADD(x, 1.f), 0.5f), (float)(W - 1));
  const float iy = TNF_MUL(TNF_MUL(TNF_ADD(y, 1.f), 0.5f), (float)(H - 1));
  const float iz = TNF_MUL(TNF_MUL(TNF_ADD(z, 1.f), 0.5f), (float)(D - 1));
  const int x0 = (int)floorf(ix), y0 = (int)floorf(iy), z0 = (int)floorf(iz);
  // only lookups whose near corner is a lattice point (far corners beyond the grid are skipped by the lookup: they only
  // lower the sum); anything else -- outside coordinates, NaN -- takes the float path
  if (!((unsigned)x0 < (unsigned)W && (unsigned)y0 < (unsigned)H && (unsigned)z0 < (unsigned)D)) return false;
  const int bx = x0 >> 1;
  const uint32_t w = coarse[((long long)(z0 >> 1) * ((H + 1) >> 1) + (y0 >> 1)) * occ_bits_words_per_row(W) + (bx >> 5)];
  if ((w >> (bx & 31)) & 1u) return true;
  if (!fine) return false;
#ifdef __CUDA_ARCH__
  const uint32_t f = __ldg(fine + ((long long)z0 * H + y0) * ((W + 31) >> 5) + (x0 >> 5));
#else
  const uint32_t f = fine[((long long)z0 * H + y0) * ((W + 31) >> 5) + (x0 >> 5)];
#endif
  return (f >> (x0 & 31)) & 1u;
}

struct SampleOut {
  float p[3];   // contracted coordinates in [-1,1]
  float step;   // step size of this sample
  bool keep;    // marcher mask & occupancy mask
};

// One lattice point (ray, step j) of RayProvider.__call__ (src/core.py:171-176).
// u = jitter in [0,1) (ignored when !jitter); tmin only used for AABB scenes.
TNF_HD SampleOut march_sample(const MarchConst& M, const float o[3], const float d[3], float tmin, int j,
                              float u, const uint32_t* coarse_bits = nullptr, const uint32_t* fine_bits = nullptr) {
  SampleOut s;
  float t, step;
  if (M.scene == 0) {  // AABB: t = t_min + j*step ; step constant (src/core.py:84-86)
    step = M.step_size;
    t = TNF_ADD(tmin, TNF_MUL((float)j, step));
  } else {             // unbounded: ray-independent tables (src/core.py:52-58)
    step = M.step_table[j];
    t = M.t_table[j];
  }
  if (M.jitter) t = TNF_ADD(t, TNF_MUL(u, step));  // src/core.py:173
  float p[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = TNF_ADD(o[c], TNF_MUL(d[c], t));  // src/core.py:174
  const bool inside = contract_point(M, p, s.p);
  s.step = step;
  // src/core.py:151-156,176: mask = marcher_mask & (trilinear(grid) > thr); the lookup has no side effect,
  // so it is skipped for samples the marcher mask already rejects (about half of an AABB lattice)
  // ... and for samples the classifier bitfields know to be empty (most of the rest in a trained scene)
  s.keep = inside && !(coarse_bits && occ_certainly_empty(coarse_bits, fine_bits, M.gd, M.gh, M.gw, s.p[0], s.p[1], s.p[2])) &&
           (trilinear_zeros(M.grid, M.gd, M.gh, M.gw, s.p[0], s.p[1], s.p[2]) > (M.thr_dev ? *M.thr_dev : M.thr));
  return s;
}

}  // namespace tnf
