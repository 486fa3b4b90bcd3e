// a4-a8, a10 -- fused ray marching + contraction + occupancy lookup + sample packing
// (reference: RayProvider.__call__ src/core.py:165-188 and the helpers it calls).
//
// The reference materialises the [R,S,3] sample lattice three times and compacts it with
// boolean-mask indexing (two host syncs).  Here one warp owns one ray:
//   count kernel : lanes take steps j = 32*i+lane, evaluate the sample on the fly (nothing but the
//                  24 B ray is read from HBM; the 8 MB occupancy grid is L2-resident), ballot the
//                  keep-mask into a bitfield word and popcount it.
//   scan kernel  : exclusive int32 scan of the per-ray counts -> packing info (+ batch offset).
//   pack kernel  : lanes re-evaluate only the kept samples, stage the [<=32][7] rows of one mask
//                  word in shared memory and stream them out with fully coalesced 128 B stores.
// Sample order is (ray, step) row-major, exactly the order of samples[mask] in the reference.
#include "common.cuh"
#include "nerf_math.cuh"

namespace tnf {
namespace {

constexpr int kWarpsM = 8;

__device__ __forceinline__ float jitter_u(const MarchConst& M, long long ray, int j) {
  if (!M.jitter) return 0.f;
  const long long flat = ray * M.n_steps + j;
  if (M.noise) return __ldg(M.noise + flat);
  const Philox ph(M.seed);
  const uint4 r = ph(M.offset + (unsigned long long)flat, 0ull);
  return u01(r.x);
}

__global__ void __launch_bounds__(kWarpsM * 32)
march_count_kernel(const MarchConst M, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                   long long n_rays, unsigned* __restrict__ mask_bits, int2* __restrict__ info) {
  const int lane = threadIdx.x & 31;
  const long long ray = blockIdx.x * (long long)kWarpsM + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  float o[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = __ldg(rays_o + ray * 3 + c);
    d[c] = __ldg(rays_d + ray * 3 + c);
  }
  const float tmin = (M.scene == 0) ? aabb_t_min(o, d, M) : 0.f;
  const int words = (M.n_steps + 31) >> 5;
  int count = 0;
  for (int i = 0; i < words; ++i) {
    const int j = i * 32 + lane;
    bool keep = false;
    if (j < M.n_steps) keep = march_sample(M, o, d, tmin, j, jitter_u(M, ray, j)).keep;
    const unsigned bits = __ballot_sync(kFullMask, keep);
    if (lane == 0) mask_bits[ray * words + i] = bits;
    count += __popc(bits);
  }
  if (lane == 0) info[ray].y = count;
}

// Variant with the "certainly empty" classifier of the occupancy grid (nerf_math.cuh; tnf_march_params.empty_bits): the
// coarse level (32 KB for 128^3) is staged in shared memory once per CTA -- the CTAs are persistent over groups of kWarpsM
// rays so the staging is paid ~600 times per launch -- and a lattice point in an empty block costs one shared-memory bit
// test, in an empty cell one more 4-byte load of the fine level, instead of 8 grid loads + 11 multiplies + 8 fused
// multiply-adds.  The keep-mask is bit-identical (tests/test_gpu_march.py).  MEASURED SLOWER than the kernel above on the
// bench scene (round 2: 78 -> 83 us in the training step, 1.54 -> 1.95 ms for the 164 M lattice points of an 800x800 pose):
// the 32 lanes of a warp are 32 consecutive steps of one ray, 72 % of the points are certainly empty but almost every
// warp holds at least one that is not, so the float path still runs for the warp and the bit tests come on top; the 8 MB
// grid is L2/L1-resident, the kernel is bound by the position / contraction arithmetic, not by the lookups.  Off by default
// (OccupancyGrid.use_empty_bits); a version that pays would compact the uncertain points of a ray before the float path.
__global__ void __launch_bounds__(kWarpsM * 32)
march_count_staged_kernel(const MarchConst M, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                          long long n_rays, unsigned* __restrict__ mask_bits, int2* __restrict__ info, int bits_words) {
  extern __shared__ uint32_t s_bits[];
  for (int i = threadIdx.x; i < bits_words; i += kWarpsM * 32) s_bits[i] = __ldg(M.empty_bits + i);
  __syncthreads();
  const uint32_t* fbits = M.empty_bits + bits_words;
  const int lane = threadIdx.x & 31;
  for (long long ray = blockIdx.x * (long long)kWarpsM + (threadIdx.x >> 5); ray < n_rays; ray += (long long)gridDim.x * kWarpsM) {
    float o[3], d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = __ldg(rays_o + ray * 3 + c);
      d[c] = __ldg(rays_d + ray * 3 + c);
    }
    const float tmin = (M.scene == 0) ? aabb_t_min(o, d, M) : 0.f;
    const int words = (M.n_steps + 31) >> 5;
    int count = 0;
    for (int i = 0; i < words; ++i) {
      const int j = i * 32 + lane;
      bool keep = false;
      if (j < M.n_steps) keep = march_sample(M, o, d, tmin, j, jitter_u(M, ray, j), s_bits, fbits).keep;
      const unsigned bits = __ballot_sync(kFullMask, keep);
      if (lane == 0) mask_bits[ray * words + i] = bits;
      count += __popc(bits);
    }
    if (lane == 0) info[ray].y = count;
  }
}

// One warp per output word of either level.  coarse bit (bz, by, bx): corners [2b, 2b+2]^3 (clipped to the grid) all
// <= thr * (1 - 1e-5); fine bit (z, y, x): corners [x, x+1] x [y, y+1] x [z, z+1] (clipped) likewise.  thr <= 0: no bits.
__global__ void __launch_bounds__(256)
occ_empty_bits_kernel(const float* __restrict__ grid, int D, int H, int W, float thr_host, const float* __restrict__ thr_dev,
                      uint32_t* __restrict__ bits) {
  const long long n_coarse = occ_bits_coarse_words(D, H, W), n_words = occ_bits_words(D, H, W);
  const long long wi = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wi >= n_words) return;
  const int lane = threadIdx.x & 31;
  const bool coarse = wi < n_coarse;
  const int span = coarse ? 2 : 1;                    // corners [span*b, span*b + span] per axis
  const long long wj = coarse ? wi : wi - n_coarse;
  const int WPR = coarse ? occ_bits_words_per_row(W) : ((W + 31) >> 5);
  const int NH = coarse ? ((H + 1) >> 1) : H, NW = coarse ? ((W + 1) >> 1) : W;
  const int wx = (int)(wj % WPR), by = (int)((wj / WPR) % NH), bz = (int)(wj / ((long long)WPR * NH));
  const int bx = wx * 32 + lane;
  const float thr = thr_dev ? __ldg(thr_dev) : thr_host;
  const float lim = TNF_SUB(thr, TNF_MUL(thr, 1e-5f));
  bool empty = false;
  if (bx < NW && thr > 0.f) {
    empty = true;
    for (int dz = 0; dz <= span && empty; ++dz) {
      const int z = span * bz + dz;
      if (z >= D) break;
      for (int dy = 0; dy <= span && empty; ++dy) {
        const int y = span * by + dy;
        if (y >= H) break;
        for (int dx = 0; dx <= span; ++dx) {
          const int x = span * bx + dx;
          if (x >= W) break;
          const float g = __ldg(grid + ((long long)z * H + y) * W + x);
          if (!(g <= lim)) { empty = false; break; }   // NaN counts as occupied: the float path decides
        }
      }
    }
  }
  const unsigned word = __ballot_sync(kFullMask, empty);
  if (lane == 0) bits[wi] = word;
}

// Single-CTA exclusive scan of info[:,1] into info[:,0] (+offset); total -> n_packed.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;
__global__ void __launch_bounds__(kScanThreads)
scan_counts_kernel(int2* __restrict__ info, long long n_rays, int offset, long long* __restrict__ n_packed) {
  __shared__ int s_warp[32];
  __shared__ long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (long long chunk = 0; chunk < n_rays; chunk += kScanThreads * kScanItems) {
    const long long r0 = chunk + (long long)tid * kScanItems;
    int c[kScanItems];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      c[i] = (r0 + i < n_rays) ? info[r0 + i].y : 0;
      sum += c[i];
    }
    int incl = sum;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const int v = __shfl_up_sync(kFullMask, incl, dlt);
      if (lane >= dlt) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int v = s_warp[lane];
      int wi = v;
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const int u = __shfl_up_sync(kFullMask, wi, dlt);
        if (lane >= dlt) wi += u;
      }
      s_warp[lane] = wi - v;  // exclusive warp offsets
    }
    __syncthreads();
    const long long base = s_base;
    long long run = base + s_warp[wid] + (incl - sum);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      if (r0 + i < n_rays) info[r0 + i].x = (int)(run + offset);
      run += c[i];
    }
    __syncthreads();
    if (tid == kScanThreads - 1) s_base = run;
    __syncthreads();
  }
  if (tid == 0) *n_packed = s_base;
}

__global__ void __launch_bounds__(kWarpsM * 32)
march_pack_kernel(const MarchConst M, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                  long long n_rays, int info_offset, const unsigned* __restrict__ mask_bits,
                  const int2* __restrict__ info, float* __restrict__ packed, float* __restrict__ steps_out,
                  int* __restrict__ ray_idx_out, long long n_packed) {
  __shared__ float s_rows[kWarpsM][32 * 7];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long ray = blockIdx.x * (long long)kWarpsM + wib;
  if (ray >= n_rays) return;
  const int2 e = __ldg(&info[ray]);
  if (e.y == 0) return;
  float o[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = __ldg(rays_o + ray * 3 + c);
    d[c] = __ldg(rays_d + ray * 3 + c);
  }
  const float tmin = (M.scene == 0) ? aabb_t_min(o, d, M) : 0.f;
  const int words = (M.n_steps + 31) >> 5;
  float* rows = s_rows[wib];
  long long dst = (long long)e.x - info_offset;  // first packed row of this ray
  for (int i = 0; i < words; ++i) {
    const unsigned bits = __ldg(mask_bits + ray * words + i);
    if (bits == 0u) continue;
    const int cnt = __popc(bits);
    if ((bits >> lane) & 1u) {
      const int j = i * 32 + lane;
      const SampleOut s = march_sample(M, o, d, tmin, j, jitter_u(M, ray, j));
      const int rank = __popc(bits & ((1u << lane) - 1u));
      float* r = rows + rank * 7;
      r[0] = s.p[0]; r[1] = s.p[1]; r[2] = s.p[2];
      r[3] = d[0];   r[4] = d[1];   r[5] = d[2];
      r[6] = s.step;
      if (steps_out && dst + rank < n_packed) steps_out[dst + rank] = s.step;
      if (ray_idx_out && dst + rank < n_packed) ray_idx_out[dst + rank] = (int)ray;
    }
    __syncwarp();
    if (dst + cnt <= n_packed) {
      float* out = packed + dst * 7;
      for (int k = lane; k < cnt * 7; k += 32) out[k] = rows[k];
    }
    __syncwarp();
    dst += cnt;
  }
}

__global__ void occ_query_kernel(const float* __restrict__ grid, int gd, int gh, int gw,
                                 const float* __restrict__ coords, long long n, float thr,
                                 uint8_t* __restrict__ out_mask, float* __restrict__ values) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = trilinear_zeros(grid, gd, gh, gw, __ldg(coords + i * 3), __ldg(coords + i * 3 + 1),
                                  __ldg(coords + i * 3 + 2));
  if (out_mask) out_mask[i] = v > thr;
  if (values) values[i] = v;
}

int make_const(const tnf_march_params* p, MarchConst* M) {
  TNF_REQUIRE(p, "null params");
  TNF_REQUIRE(p->scene == TNF_SCENE_AABB || p->scene == TNF_SCENE_UNBOUNDED, "bad scene %d", p->scene);
  TNF_REQUIRE(p->n_steps > 0, "n_steps must be positive");
  TNF_REQUIRE(p->grid && p->gd > 0 && p->gh > 0 && p->gw > 0, "bad occupancy grid");
  TNF_REQUIRE(p->scene == TNF_SCENE_AABB || (p->t_table && p->step_table), "unbounded scene needs t/step tables");
  M->scene = p->scene;
  M->n_steps = p->n_steps;
  for (int c = 0; c < 3; ++c) {
    M->a0[c] = p->aabb[c];
    M->a1[c] = p->aabb[3 + c];
    volatile float e = p->aabb[3 + c] - p->aabb[c];  // fp32 subtraction, as torch does on the tensor
    M->ext[c] = e;
  }
  M->near_ = p->near; M->far_ = p->far; M->step_size = p->step_size;
  M->t_table = p->t_table; M->step_table = p->step_table;
  M->grid = p->grid; M->gd = p->gd; M->gh = p->gh; M->gw = p->gw;
  M->thr = p->threshold; M->thr_dev = p->threshold_dev; M->noise = p->noise; M->jitter = p->jitter;
  M->seed = p->seed; M->offset = p->offset;
  M->empty_bits = p->empty_bits;
  return TNF_OK;
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_march_count(const tnf_march_params* p, const float* rays_o, const float* rays_d,
                               int64_t n_rays, int32_t info_offset, uint32_t* mask_bits, int32_t* info,
                               int64_t* n_packed, void* stream) {
  using namespace tnf;
  MarchConst M;
  int rc = make_const(p, &M);
  if (rc != TNF_OK) return rc;
  TNF_REQUIRE(n_rays >= 0 && n_rays < (1LL << 31), "bad n_rays %lld", (long long)n_rays);
  TNF_REQUIRE(n_packed, "null n_packed");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_rays == 0) {
    TNF_CUDA(cudaMemsetAsync(n_packed, 0, sizeof(int64_t), st));
    return TNF_OK;
  }
  TNF_REQUIRE(rays_o && rays_d && mask_bits && info, "null pointer");
  TNF_REQUIRE((reinterpret_cast<uintptr_t>(info) & 7u) == 0, "info must be 8-byte aligned");
  TNF_REQUIRE(n_rays * (int64_t)p->n_steps < (1LL << 31), "n_rays*n_steps must fit int32 packing info");
  const long long groups = ceil_div(n_rays, kWarpsM);
  const long long bits_words = occ_bits_coarse_words(M.gd, M.gh, M.gw);   // the level staged in shared memory
  if (M.empty_bits && bits_words * 4 <= 64 * 1024) {
    // persistent CTAs (4 per SM: 59 registers x 256 threads), each stages the bitfield once
    const size_t smem = (size_t)bits_words * 4;
    static PerDeviceOnce configured{};
    if (configured.pending()) {
      TNF_CUDA(cudaFuncSetAttribute(march_count_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      configured.mark();
    }
    const long long cap = (long long)sm_count() * 4;
    march_count_staged_kernel<<<(unsigned)(groups < cap ? groups : cap), kWarpsM * 32, smem, st>>>(
        M, rays_o, rays_d, n_rays, mask_bits, reinterpret_cast<int2*>(info), (int)bits_words);
  } else {
    march_count_kernel<<<(unsigned)groups, kWarpsM * 32, 0, st>>>(M, rays_o, rays_d, n_rays, mask_bits,
                                                                 reinterpret_cast<int2*>(info));
  }
  TNF_LAUNCH_CHECK("march_count_kernel");
  scan_counts_kernel<<<1, kScanThreads, 0, st>>>(reinterpret_cast<int2*>(info), n_rays, info_offset,
                                                 reinterpret_cast<long long*>(n_packed));
  TNF_LAUNCH_CHECK("scan_counts_kernel");
  return TNF_OK;
}

extern "C" int tnf_march_pack(const tnf_march_params* p, const float* rays_o, const float* rays_d,
                              int64_t n_rays, int32_t info_offset, const uint32_t* mask_bits,
                              const int32_t* info, float* packed, float* steps_out, int32_t* ray_idx_out,
                              int64_t n_packed, void* stream) {
  using namespace tnf;
  MarchConst M;
  int rc = make_const(p, &M);
  if (rc != TNF_OK) return rc;
  TNF_REQUIRE(n_rays >= 0 && n_packed >= 0, "negative size");
  if (n_rays == 0 || n_packed == 0) return TNF_OK;
  TNF_REQUIRE(rays_o && rays_d && mask_bits && info && packed, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  march_pack_kernel<<<(unsigned)ceil_div(n_rays, kWarpsM), kWarpsM * 32, 0, st>>>(
      M, rays_o, rays_d, n_rays, info_offset, mask_bits, reinterpret_cast<const int2*>(info), packed,
      steps_out, ray_idx_out, n_packed);
  TNF_LAUNCH_CHECK("march_pack_kernel");
  return TNF_OK;
}

extern "C" int64_t tnf_occ_empty_bits_words(int32_t gd, int32_t gh, int32_t gw) {
  return (gd > 0 && gh > 0 && gw > 0) ? tnf::occ_bits_words(gd, gh, gw) : 0;
}

extern "C" int tnf_occ_build_empty_bits(const float* grid, int32_t gd, int32_t gh, int32_t gw, float threshold,
                                        const float* threshold_dev, uint32_t* bits, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(grid && bits && gd > 0 && gh > 0 && gw > 0, "bad grid/bits");
  const long long n_words = occ_bits_words(gd, gh, gw);
  occ_empty_bits_kernel<<<(unsigned)ceil_div(n_words, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(grid, gd, gh, gw, threshold,
                                                                                                  threshold_dev, bits);
  TNF_LAUNCH_CHECK("occ_empty_bits_kernel");
  return TNF_OK;
}

extern "C" int tnf_occ_query(const float* grid, int32_t gd, int32_t gh, int32_t gw, const float* coords,
                             int64_t n, float threshold, uint8_t* out_mask, float* values, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0, "negative n");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(grid && coords && gd > 0 && gh > 0 && gw > 0, "bad grid/coords");
  TNF_REQUIRE(out_mask || values, "no output requested");
  occ_query_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grid, gd, gh, gw, coords, n, threshold, out_mask, values);
  TNF_LAUNCH_CHECK("occ_query_kernel");
  return TNF_OK;
}
