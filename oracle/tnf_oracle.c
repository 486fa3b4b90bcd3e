/*
 * oracle/tnf_oracle.c -- CPU restatement of the reference algorithm for the packed-ray hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in tinynerf_b200/ may import, link or call this file; it is the
 * checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Each function follows the cited lines of loicmagne/tinynerf (paths relative to the
 * reference root) literally: serial per-ray order, one rounding per un-fused PyTorch op.
 *
 * Compile: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/build.py).  -ffp-contract=off matters:
 * every fused multiply-add below is written out with fmaf() on purpose.
 *
 * Pinning (see DESIGN.md "oracle"): the reference has no CPU implementation of the weights op
 * (src/cuda.cu:62 rejects CPU tensors), so orc_weights_* is checked (a) here, against an independent
 * float64 autograd cumprod formulation, (b) on the GPU box against the UNMODIFIED reference kernel
 * built into oracle/_ref/_cuda.so.  The march/occupancy functions are checked against golden vectors
 * produced by importing the reference's src/core.py (tests/golden/make_golden.py) and against the
 * reference's only known-answer test (tests/test_core.py:5-38).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ---- a1: src/cuda.cu:3-30 (kernel_compute_weights_fwd), one "thread" per ray ------------------ */
void orc_weights_fwd(const float* sigmas, const float* steps, int64_t steps_stride, const int32_t* info,
                     float threshold, float* weights, int64_t n_samples, int64_t n_rays) {
  memset(weights, 0, sizeof(float) * (size_t)n_samples); /* torch::zeros_like, src/cuda.cu:84 */
  for (int64_t idx = 0; idx < n_rays; ++idx) {
    const int n = info[2 * idx + 1];
    const int ray_start = info[2 * idx];
    const int ray_end = ray_start + n;
    if (n == 0) continue;
    float transmittance = 1.f;
    int k = ray_start;
    while (transmittance > threshold && k < ray_end) { /* early termination, src/cuda.cu:23 */
      const float alpha = expf(-sigmas[k] * steps[(int64_t)k * steps_stride]); /* __expf on the GPU */
      weights[k] = (float)((double)transmittance * (1. - (double)alpha));      /* `1.` is a double */
      transmittance *= alpha;
      k++;
    }
  }
}

/* ---- a2: src/cuda.cu:32-58 (kernel_compute_weights_bwd) -------------------------------------- */
void orc_weights_bwd(const float* sigmas, const float* steps, int64_t steps_stride, const int32_t* info,
                     const float* weights, const float* grad_weights, float* grad_sigmas,
                     int64_t n_samples, int64_t n_rays) {
  memset(grad_sigmas, 0, sizeof(float) * (size_t)n_samples);
  for (int64_t idx = 0; idx < n_rays; ++idx) {
    const int n = info[2 * idx + 1];
    const int ray_start = info[2 * idx];
    const int ray_end = ray_start + n;
    if (n == 0) continue;
    float acc = 0.f, transmittance = 1.f;
    /* nvcc contracts `acc -= w*g` / `acc += w*g` / `acc + T*g` into FFMA (default -fmad=true) */
    for (int k = ray_start; k < ray_end; k++) acc = fmaf(-weights[k], grad_weights[k], acc);
    for (int k = ray_start; k < ray_end; k++) {
      const float st = steps[(int64_t)k * steps_stride];
      acc = fmaf(weights[k], grad_weights[k], acc);
      transmittance *= expf(-sigmas[k] * st);
      grad_sigmas[k] = st * fmaf(transmittance, grad_weights[k], acc);
    }
  }
}

/* ---- a8: OccupancyGrid.forward value, src/core.py:151-155 = F.grid_sample(5-D, bilinear, zeros,
 * align_corners=True); arithmetic of torch's CUDA kernel (ATen/native/cuda/GridSampler.cuh:21-31
 * for the unnormalisation; corner weights are products of differences; the 8 accumulations are
 * `out_acc += v * w`, which nvcc contracts to FFMA -- use_fma=0 gives the uncontracted variant so a
 * test can tell which one a given torch build uses). */
float orc_trilinear(const float* g, int D, int H, int W, float x, float y, float z, int use_fma) {
  const float ix = ((x + 1.f) / 2.f) * (float)(W - 1);
  const float iy = ((y + 1.f) / 2.f) * (float)(H - 1);
  const float iz = ((z + 1.f) / 2.f) * (float)(D - 1);
  const int x0 = (int)floorf(ix), y0 = (int)floorf(iy), z0 = (int)floorf(iz);
  const int x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
  const float ex = (float)x1 - ix, wx = ix - (float)x0;
  const float ey = (float)y1 - iy, wy = iy - (float)y0;
  const float ez = (float)z1 - iz, wz = iz - (float)z0;
  const float w[8] = {ex * ey * ez, wx * ey * ez, ex * wy * ez, wx * wy * ez,
                      ex * ey * wz, wx * ey * wz, ex * wy * wz, wx * wy * wz};
  const int xs[8] = {x0, x1, x0, x1, x0, x1, x0, x1};
  const int ys[8] = {y0, y0, y1, y1, y0, y0, y1, y1};
  const int zs[8] = {z0, z0, z0, z0, z1, z1, z1, z1};
  float acc = 0.f;
  for (int k = 0; k < 8; ++k) {
    if (xs[k] < 0 || xs[k] >= W || ys[k] < 0 || ys[k] >= H || zs[k] < 0 || zs[k] >= D) continue;
    const float v = g[((int64_t)zs[k] * H + ys[k]) * W + xs[k]];
    acc = use_fma ? fmaf(v, w[k], acc) : (acc + v * w[k]);
  }
  return acc;
}

void orc_occ_query(const float* grid, int D, int H, int W, const float* coords, int64_t n, float thr,
                   int use_fma, uint8_t* mask, float* values) {
  for (int64_t i = 0; i < n; ++i) {
    const float v = orc_trilinear(grid, D, H, W, coords[3 * i], coords[3 * i + 1], coords[3 * i + 2], use_fma);
    if (mask) mask[i] = v > thr; /* src/core.py:156, compared in fp32 */
    if (values) values[i] = v;
  }
}

/* ---- a4-a7, a10: RayProvider.__call__, src/core.py:165-188 ------------------------------------
 * scene 0 = RayMarcherAABB (:73-88) + ContractionAABB (:27-31); scene 1 = RayMarcherUnbounded
 * (:48-59; the ray-independent t/step tables are inputs) + ContractionMip360(order=inf) (:16-20).
 * Outputs: mask[R*S] (0/1), info[R][2], packed[n][7] (if packed != NULL, capacity cap rows).
 * Returns the number of packed samples. */
typedef struct {
  int scene, n_steps;
  float aabb[6];
  float near_, far_, step_size;
  const float* t_table;
  const float* step_table;
  const float* grid;
  int gd, gh, gw;
  float thr;
  const float* noise; /* [R][S] or NULL (training=False) */
  int use_fma;
} orc_march_params;

static float nanmin(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
static float nanmax(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

int64_t orc_march(const orc_march_params* P, const float* rays_o, const float* rays_d, int64_t n_rays,
                  uint8_t* mask, int32_t* info, float* packed, int64_t cap) {
  const int S = P->n_steps;
  int64_t total = 0;
  for (int64_t r = 0; r < n_rays; ++r) {
    const float* o = rays_o + 3 * r;
    const float* d = rays_d + 3 * r;
    float t_min = 0.f;
    if (P->scene == 0) {
      /* src/core.py:77-81: (aabb - o) / where(d == 0, d + eps, d); amax over axes of amin over planes */
      for (int c = 0; c < 3; ++c) {
        const float den = (d[c] == 0.f) ? (d[c] + 1e-9f) : d[c];
        const float m = nanmin((P->aabb[c] - o[c]) / den, (P->aabb[3 + c] - o[c]) / den);
        t_min = (c == 0) ? m : nanmax(t_min, m);
      }
      if (t_min == t_min) {
        if (t_min < P->near_) t_min = P->near_;
        if (t_min > P->far_) t_min = P->far_;
      }
    }
    int count = 0;
    info[2 * r] = (int32_t)total;
    for (int j = 0; j < S; ++j) {
      float t, step;
      if (P->scene == 0) { /* src/core.py:84-86 */
        step = P->step_size;
        t = t_min + (float)j * step;
      } else {
        step = P->step_table[j];
        t = P->t_table[j];
      }
      if (P->noise) t = t + P->noise[r * S + j] * step; /* src/core.py:173 */
      float p[3], q[3];
      for (int c = 0; c < 3; ++c) p[c] = o[c] + d[c] * t; /* src/core.py:174, mul then add */
      int keep = 1;
      if (P->scene == 0) { /* src/core.py:29-30 */
        for (int c = 0; c < 3; ++c) {
          keep = keep && (p[c] >= P->aabb[c]) && (p[c] <= P->aabb[3 + c]);
          q[c] = (p[c] - P->aabb[c]) / (P->aabb[3 + c] - P->aabb[c]) * 2.f - 1.f;
        }
      } else { /* src/core.py:18-19 with order = inf */
        const float n = fmaxf(fmaxf(fabsf(p[0]), fabsf(p[1])), fabsf(p[2]));
        for (int c = 0; c < 3; ++c) {
          if (n <= 1.f) {
            q[c] = p[c] / 2.f;
          } else {
            const float rcp = (1.f / n) * 1.f; /* `1./norm` is norm.reciprocal() * 1. */
            q[c] = ((2.f - rcp) * p[c] / n) / 2.f;
          }
        }
      }
      const float v = orc_trilinear(P->grid, P->gd, P->gh, P->gw, q[0], q[1], q[2], P->use_fma);
      keep = keep && (v > P->thr);
      if (mask) mask[r * S + j] = (uint8_t)keep;
      if (keep) {
        if (packed && total < cap) { /* src/core.py:183-186: (coords, dir, step) */
          float* row = packed + 7 * total;
          row[0] = q[0]; row[1] = q[1]; row[2] = q[2];
          row[3] = d[0]; row[4] = d[1]; row[5] = d[2];
          row[6] = step;
        }
        ++total;
        ++count;
      }
    }
    info[2 * r + 1] = count;
  }
  return total;
}

/* ---- a9: OccupancyGrid.update, src/core.py:134-145 (around the sigma_fn call) ---------------- */
void orc_occ_update_coords(int gd, int gh, int gw, int64_t cell0, int64_t n, const float* noise, float* coords) {
  const float size[3] = {(float)gd, (float)gh, (float)gw}; /* self.size = (D,H,W) divides (x,y,z) */
  for (int64_t i = 0; i < n; ++i) {
    const int64_t cell = cell0 + i;
    const float c[3] = {(float)(cell % gw), (float)((cell / gw) % gh), (float)(cell / ((int64_t)gw * gh))};
    for (int k = 0; k < 3; ++k) coords[3 * i + k] = -1.f + 2.f * (c[k] + noise[3 * i + k]) / size[k];
  }
}
void orc_occ_update_apply(float* grid, int64_t cell0, int64_t n, const float* sigma, float step, float thr,
                          float decay) {
  for (int64_t i = 0; i < n; ++i) {
    const float alpha = 1.f - expf(-sigma[i] * step);
    grid[cell0 + i] = (alpha > thr) ? 1.f : decay * grid[cell0 + i];
  }
}

/* ---- a18: compositing, src/core.py:256-265 (index_add_ over repeat_interleave'd ray ids) ------ */
void orc_composite_fwd(const float* w, const float* rgb, const int32_t* info, int64_t n_rays, const float* bg,
                       float* out) {
  for (int64_t r = 0; r < n_rays; ++r) {
    double acc[3] = {0, 0, 0}, op = 0; /* order-free reference value: accumulate in double */
    for (int k = info[2 * r]; k < info[2 * r] + info[2 * r + 1]; ++k) {
      for (int c = 0; c < 3; ++c) acc[c] += (double)(rgb[3 * k + c] * w[k]);
      op += w[k];
    }
    for (int c = 0; c < 3; ++c) out[3 * r + c] = (float)(bg ? acc[c] + (double)bg[c] * (1.0 - op) : acc[c]);
  }
}
void orc_composite_bwd(const float* w, const float* rgb, const int32_t* info, int64_t n_rays, const float* bg,
                       const float* go, float* gw, float* grgb) {
  for (int64_t r = 0; r < n_rays; ++r) {
    const float* g = go + 3 * r;
    const double gbg = bg ? (double)bg[0] * g[0] + (double)bg[1] * g[1] + (double)bg[2] * g[2] : 0.0;
    for (int k = info[2 * r]; k < info[2 * r] + info[2 * r + 1]; ++k) {
      gw[k] = (float)((double)rgb[3 * k] * g[0] + (double)rgb[3 * k + 1] * g[1] + (double)rgb[3 * k + 2] * g[2] - gbg);
      for (int c = 0; c < 3; ++c) grgb[3 * k + c] = w[k] * g[c];
    }
  }
}
