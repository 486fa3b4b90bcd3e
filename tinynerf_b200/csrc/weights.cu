// a1/a2 -- NeRF-equation weights over packed variable-length rays (reference: src/cuda.cu:3-58).
//
// Design (see DESIGN.md "weights"): the reference walks each ray with one thread.  Here the sample
// axis is streamed: every warp owns a "task" = all rays whose first sample lies in one tile of the
// sample axis, reads that contiguous sample window with perfectly coalesced 128-bit loads (each
// lane 4 consecutive samples, 512 B per warp instruction) and evaluates the segmented exclusive
// cumprod of a_k = exp(-sigma_k*delta_k) with a warp-segmented scan (shuffle Hillis-Steele over
// per-lane aggregates + a register carry from round to round).  Ray boundaries come from a per-warp
// shared-memory bitfield of ray-head positions built from packing info.  Tasks are aligned to rays,
// so warps never wait on each other (no look-back chain, no deadlock potential).
//
// Exact termination: the reference stops a ray at the first k with T_k <= thr where T_k is the
// *serial* fp32 product.  A scan re-associates the product, so T can differ by <= (k-1)*2^-23
// relative.  Any ray that has a sample whose scanned T lies inside that rigorous band around thr is
// re-evaluated by one lane in the reference's serial order (rare: ~1e-3 of terminating rays), which
// makes the (weights > 0) mask bit-identical to the reference.
#include <stdlib.h>
#include "common.cuh"

namespace tnf {
namespace {

constexpr int kWarps = 8;                  // warps per CTA, each runs independent tasks
constexpr int kMaxTile = 2048;             // max samples per task tile
constexpr int kMaskWords = kMaxTile / 32;  // head bitfield words per warp
constexpr int kQueue = 32;                 // per-task queue of samples needing serial re-evaluation

struct WArgs {
  const float* sigmas;
  const float* steps;
  long long sstride;
  const int2* info;
  float thr;
  float* out;      // weights (fwd) / grad_sigmas (bwd)
  const float* w;  // bwd: saved weights
  const float* g;  // bwd: grad wrt weights
  long long n;
  int n_rays;
  int tile;
  int n_tasks;
  unsigned* status;
  int flags;
};

// ---- per-warp cp.async ring -------------------------------------------------------------------------
// Every lane copies the 16 bytes (4 samples) it will consume itself straight from global to its private slot of a
// shared-memory ring (LDGSTS, no registers held, L1 bypassed) kDepth-1 rounds ahead, and later reads the slot back
// with one LDS.128.  No cross-lane traffic, so cp.async.wait_group is the only synchronisation.  This keeps
// (kDepth-1) x 1-1.5 KB of loads in flight per warp instead of the one round a register prefetch can afford.
constexpr int kDepth = 4;
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool ok) {  // !ok: zero-fill, nothing is read
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(ok ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_ring() { asm volatile("cp.async.wait_group %0;" ::"n"(kDepth - 1) : "memory"); }

// MODE 0: all arrays contiguous + 16B aligned (float4 path); 1: steps strided, rest aligned; 2: scalar.
// `p` points at the task's tile origin, `rel` is the tile-relative index of the lane's first sample and
// `lim` = n_samples - tile0 bounds the reads.
template <int MODE>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, int rel, int lim, float v[4]) {
  if (MODE != 2 && rel + 3 < lim) {
    const float4 t = ld_stream_f4(p + rel);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (rel + i < lim) ? __ldg(p + rel + i) : 0.f;
  }
}
template <int MODE>
__device__ __forceinline__ void load_steps(const float* __restrict__ p, long long stride, int rel, int lim,
                                           float v[4]) {
  if (MODE == 0) {
    load_vec<0>(p, rel, lim, v);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (rel + i < lim) ? __ldg(p + (long long)(rel + i) * stride) : 0.f;
  }
}
// asynchronous variants of load_vec / load_steps into a ring slot (zeros beyond the end of the arrays)
template <int MODE>
__device__ __forceinline__ void ring_vec(float4* slot, const float* __restrict__ p, int rel, int lim) {
  if (MODE != 2 && rel + 3 < lim) {
    cp_async16(slot, p + rel);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) cp_async4(reinterpret_cast<float*>(slot) + i, (rel + i < lim) ? p + rel + i : p, rel + i < lim);
  }
}
template <int MODE>
__device__ __forceinline__ void ring_steps(float4* slot, const float* __restrict__ p, long long stride, int rel, int lim) {
  if (MODE == 0) {
    ring_vec<0>(slot, p, rel, lim);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      cp_async4(reinterpret_cast<float*>(slot) + i, (rel + i < lim) ? p + (long long)(rel + i) * stride : p, rel + i < lim);
  }
}
__device__ __forceinline__ void unpack4(const float4 t, float v[4]) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }

// store the lane's 4 values, only samples in [lo, hi) (tile-relative) belong to this task
template <int MODE>
__device__ __forceinline__ void store_vec(float* p, int rel, int lo, int hi, const float v[4]) {
  if (MODE != 2 && rel >= lo && rel + 3 < hi) {
    st_stream_f4(p + rel, make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (rel + i >= lo && rel + i < hi) p[rel + i] = v[i];
  }
}

// First r in [0,n) with info[r].start >= key (n if none); warp-cooperative 32-ary search, 4 rounds
// for 2^20 rays.  Terminates for arbitrary (unsorted) data.
__device__ __forceinline__ int warp_lower_bound(const int2* __restrict__ info, int n, long long key,
                                                int lane) {
  int lo = 0, hi = n;
  while (hi > lo) {
    const int len = hi - lo;
    const int step = (len + 31) >> 5;
    const long long pl = (long long)lo + (long long)lane * step;
    bool pred = false;
    if (pl < hi) pred = (long long)__ldg(&info[pl].x) < key;
    const unsigned b = __ballot_sync(kFullMask, pred);
    // predicates are a prefix of ones when sorted; for garbage use the first zero as the split
    const int cnt = __ffs(~b) - 1;  // number of leading true lanes (32 if all true -> ffs(0)=0 -> -1)
    const int c = (b == kFullMask) ? 32 : cnt;
    if (c == 0) {
      hi = lo;
    } else {
      const long long q = (long long)lo + (long long)(c - 1) * step;
      const long long nh = q + step;
      lo = (int)(q + 1);
      hi = (int)(nh < hi ? nh : hi);
    }
  }
  return lo;
}

__device__ __forceinline__ int find_head_le(const unsigned* mask, int b, int tile) {
  if (b >= tile) b = tile - 1;
  int w = b >> 5;
  unsigned m = mask[w] & (0xFFFFFFFFu >> (31 - (b & 31)));
  while (m == 0u && w > 0) m = mask[--w];
  return m ? (w << 5) + 31 - __clz(m) : -1;
}
__device__ __forceinline__ int find_head_gt(const unsigned* mask, int b, int tile) {
  ++b;
  if (b >= tile) return -1;
  int w = b >> 5;
  const int nw = tile >> 5;
  unsigned m = mask[w] & (0xFFFFFFFFu << (b & 31));
  while (m == 0u && w + 1 < nw) m = mask[++w];
  return m ? (w << 5) + __ffs(m) - 1 : -1;
}
// bits [rel, rel+5) of the head bitfield (zero beyond the tile)
__device__ __forceinline__ unsigned head_bits5(const unsigned* mask, int rel, int tile) {
  if (rel >= tile) return 0u;
  const unsigned lo = mask[rel >> 5];
  const unsigned hi = (rel + 32 < tile) ? mask[(rel >> 5) + 1] : 0u;
  const unsigned long long both = ((unsigned long long)hi << 32) | lo;
  return (unsigned)(both >> (rel & 31)) & 0x1Fu;
}

// Warp-segmented scans over per-lane aggregates without shuffling flags: the lanes that contain a ray
// head (resp. a ray tail) are known to the whole warp from one ballot, so every lane derives how many
// lanes below (resp. above) it may absorb: `lim`.  Step d of the Hillis-Steele scan applies iff lim >= d.
__device__ __forceinline__ int absorb_limit_up(unsigned heads, int lane) {
  const unsigned below = heads & (0xFFFFFFFFu >> (31 - lane));  // lanes 0..lane
  return below ? lane - (31 - __clz(below)) : lane;
}
__device__ __forceinline__ int absorb_limit_down(unsigned tails, int lane) {
  const unsigned above = tails >> lane;  // bit 0 = this lane
  return above ? __ffs(above) - 1 : 31 - lane;
}
__device__ __forceinline__ float seg_scan_mul_up(float P, int lim) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float Pu = __shfl_up_sync(kFullMask, P, d);
    if (lim >= d) P *= Pu;
  }
  return P;
}
__device__ __forceinline__ float seg_scan_add_down(float P, int lim) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float Pu = __shfl_down_sync(kFullMask, P, d);
    if (lim >= d) P += Pu;
  }
  return P;
}

// a = __expf(x), x = -sigma*delta <= 0 (src/cuda.cu:24).  __expf is ex2.approx(x * log2e) wrapped in a rescaling that
// only acts when the result is below 2^-126 (6 instructions per sample).  FTZ == true drops the wrapper: bit-identical
// for every a >= 2^-126 and 0 instead of a denormal below -- indistinguishable downstream unless the threshold itself
// is below ~2^-99 (the caller keeps the exact form then): T*a is <= thr either way, and 1.-a is 1 in both.
template <bool FTZ>
__device__ __forceinline__ float exp_fast(float x) {
  if (!FTZ) return __expf(x);
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950216293334961f));
  return y;
}

// w = T*(1.-a) exactly as the reference's fp64 expression (src/cuda.cu:25).  For a in [0.5, 2] the fp32
// evaluation is bit-identical: 1-a is exact (Sterbenz) and the 48-bit product is exact in fp64, so both
// round the same real number to fp32 once.  Only a < 0.5 (sigma*delta > 0.69) takes the fp64 path.
__device__ __forceinline__ void weights_from_T(const float T[4], const float a[4], float thr, float w[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = (T[i] > thr) ? T[i] * (1.f - a[i]) : 0.f;
  if (!(fminf(fminf(a[0], a[1]), fminf(a[2], a[3])) >= 0.5f)) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (!(a[i] >= 0.5f) && T[i] > thr) w[i] = (float)((double)T[i] * (1. - (double)a[i]));
  }
}

// The reference's loop, verbatim semantics (src/cuda.cu:19-28), plus the zeros it gets from
// zeros_like (src/cuda.cu:84) for samples after termination.
__device__ __noinline__ void serial_ray_fwd(const float* __restrict__ sig, const float* __restrict__ stp,
                                            long long ss, long long start, long long end, float thr,
                                            float* out) {
  float transmittance = 1.f;
  long long k = start;
  // Eight samples at a time while none of them terminates: the loads and exponentials of a batch are independent, only
  // the transmittance products form a chain, in the reference's order -- same values, ~5x fewer stall cycles for the
  // one lane that walks.  The scalar loop below finishes the ray (termination inside a batch, tail).
  while (k + 8 <= end) {
    float al[8], tr[9];
#pragma unroll
    for (int i = 0; i < 8; ++i) al[i] = __expf(-sig[k + i] * stp[(k + i) * ss]);
    tr[0] = transmittance;
    bool alive = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      alive &= tr[i] > thr;
      tr[i + 1] = tr[i] * al[i];
    }
    if (!alive) break;
#pragma unroll
    for (int i = 0; i < 8; ++i) out[k + i] = tr[i] * (1. - al[i]);  // double multiply, as in the reference
    transmittance = tr[8];
    k += 8;
  }
  while (transmittance > thr && k < end) {
    const float alpha = __expf(-sig[k] * stp[k * ss]);
    out[k] = transmittance * (1. - alpha);  // double multiply, as in the reference
    transmittance *= alpha;
    ++k;
  }
  for (; k < end; ++k) out[k] = 0.f;
}
// src/cuda.cu:49-56 (two serial passes; fused-multiply-adds written out explicitly).
__device__ __noinline__ void serial_ray_bwd(const float* __restrict__ sig, const float* __restrict__ stp,
                                            long long ss, const float* __restrict__ w,
                                            const float* __restrict__ g, long long start, long long end,
                                            float* out) {
  float acc = 0.f, transmittance = 1.f;
  for (long long k = start; k < end; ++k) acc = __fmaf_rn(-w[k], g[k], acc);
  for (long long k = start; k < end; ++k) {
    acc = __fmaf_rn(w[k], g[k], acc);
    transmittance *= __expf(-sig[k] * stp[k * ss]);
    out[k] = stp[k * ss] * __fmaf_rn(transmittance, g[k], acc);
  }
}

struct Task {
  long long tile0, s0, s1;
  int r_lo, r_hi;
  int lo, hi;        // s0 - tile0, s1 - tile0 (tile-relative sample window owned by the task)
  int lim;           // n_samples - tile0 (< 2^31: packing info is int32)
  bool valid;
};

// Common task prologue: optional partition validation, ray range, sample window, head bitfield.
__device__ __forceinline__ Task task_setup(const WArgs& A, int task, unsigned* mask, int lane) {
  Task t;
  const long long N = A.n;
  const int R = A.n_rays;
  if (!(A.flags & TNF_W_TRUSTED_PARTITION)) {
    // every task validates a static slice of the ray list (coalesced 8 B/ray, +~1% traffic)
    const long long rb = (long long)task * R / A.n_tasks, re = (long long)(task + 1) * R / A.n_tasks;
    bool bad = false;
    for (long long r = rb + lane; r < re; r += 32) {
      const int2 e = __ldg(&A.info[r]);
      const long long nxt = (r + 1 < R) ? (long long)__ldg(&A.info[r + 1].x) : N;
      bad |= (e.y < 0) | (e.x < 0) | ((long long)e.x + e.y != nxt) | (r == 0 && e.x != 0);
    }
    if (__any_sync(kFullMask, bad) && lane == 0) atomicOr(A.status, 1u);
  }
  t.tile0 = (long long)task * A.tile;
  const long long tile_end = t.tile0 + A.tile;
  t.r_lo = warp_lower_bound(A.info, R, t.tile0, lane);
  for (int i = lane; i < (A.tile >> 5); i += 32) mask[i] = 0u;
  __syncwarp();
  // The rays of the tile follow r_lo contiguously: one forward sweep over packing info both scatters the ray heads
  // into the bitfield and finds r_hi (first ray starting at or after the tile's end) -- no second binary search.
  // Bounded (64 x 32 rays) so that arbitrary, unsorted info cannot make it run long; beyond that, search.
  t.r_hi = -1;
  int base = t.r_lo;
  for (int it = 0; it < 64; ++it, base += 32) {
    const int r = base + lane;
    int2 e = make_int2(0, 0);
    if (r < R) e = __ldg(&A.info[r]);
    const bool beyond = (r >= R) || ((long long)e.x >= tile_end);
    const unsigned bm = __ballot_sync(kFullMask, beyond);
    const bool mine = bm ? (lane < __ffs(bm) - 1) : true;  // rays before the first "beyond" ray
    const long long b = (long long)e.x - t.tile0;
    if (mine && e.y > 0 && b >= 0 && b < A.tile) atomicOr(&mask[b >> 5], 1u << (b & 31));
    if (bm) { t.r_hi = base + __ffs(bm) - 1; break; }
  }
  if (t.r_hi < 0) {
    t.r_hi = warp_lower_bound(A.info, R, tile_end, lane);
    for (int r = base + lane; r < t.r_hi; r += 32) {
      const int2 e = __ldg(&A.info[r]);
      const long long b = (long long)e.x - t.tile0;
      if (e.y > 0 && b >= 0 && b < A.tile) atomicOr(&mask[b >> 5], 1u << (b & 31));
    }
  }
  long long s0 = (t.r_lo < R) ? (long long)__ldg(&A.info[t.r_lo].x) : N;
  long long s1 = (t.r_hi < R) ? (long long)__ldg(&A.info[t.r_hi].x) : N;
  s0 = s0 < 0 ? 0 : (s0 > N ? N : s0);
  s1 = s1 < 0 ? 0 : (s1 > N ? N : s1);
  t.s0 = s0;
  t.s1 = s1;
  t.lo = (int)(s0 - t.tile0);
  t.hi = (int)(s1 - t.tile0);
  t.lim = (int)(N - t.tile0);
  t.valid = (s0 >= t.tile0) && (s1 > s0) && (t.r_hi > t.r_lo);
  __syncwarp();
  return t;
}

template <int MODE, bool FTZ>
__global__ void __launch_bounds__(kWarps * 32, 4) weights_fwd_kernel(const WArgs A) {
  extern __shared__ float4 s_ring_dyn[];  // [kWarps][kDepth][2 arrays][32 lanes]
  __shared__ unsigned s_mask[kWarps][kMaskWords];
  __shared__ int s_q[kWarps][kQueue];
  __shared__ int s_qn[kWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned* mask = s_mask[wib];
  int* queue = s_q[wib];
  const bool exact = !(A.flags & TNF_W_NO_EXACT_TERMINATION);
  const float thr = A.thr;
  const bool tiny_thr = thr < 1.5777218e-30f;  // 2^-99: transmittance may underflow before stopping
  const int n_warps = gridDim.x * kWarps;

  for (int task = blockIdx.x * kWarps + wib; task < A.n_tasks; task += n_warps) {
    if (!(A.flags & TNF_W_TRUSTED_PARTITION) && *(volatile unsigned*)A.status) return;
    const Task t = task_setup(A, task, mask, lane);
    if (!t.valid) continue;
    if (lane == 0) s_qn[wib] = 0;
    __syncwarp();

    const float* sig = A.sigmas + t.tile0;
    const float* stp = A.steps + t.tile0 * A.sstride;
    float* out = A.out + t.tile0;
    float carry = 1.f;  // product since the last ray head, through the end of the previous round
    const int r_first = t.lo & ~127;
    const int n_rounds = (t.hi - r_first + 127) >> 7;
    float4* ring = s_ring_dyn + wib * (kDepth * 2 * 32) + lane;  // slot (stage, array) at ring[(stage*2+array)*32]
    auto issue = [&](int j) {  // start the copies of round j (an empty group past the end keeps the count uniform)
      if (j < n_rounds) {
        const int rl = r_first + j * 128 + lane * 4;
        float4* slot = ring + ((j & (kDepth - 1)) * 2) * 32;
        ring_vec<MODE>(slot, sig, rl, t.lim);
        ring_steps<MODE>(slot + 32, stp, A.sstride, rl, t.lim);
      }
      cp_async_commit();
    };
#pragma unroll
    for (int j = 0; j < kDepth - 1; ++j) issue(j);
    for (int j = 0; j < n_rounds; ++j) {
      const int rpos = r_first + j * 128;
      const int rel = rpos + lane * 4;
      issue(j + kDepth - 1);  // refills the stage consumed in the previous iteration
      cp_async_wait_ring();   // round j has landed
      float s[4], d[4];
      unpack4(ring[((j & (kDepth - 1)) * 2) * 32], s);
      unpack4(ring[((j & (kDepth - 1)) * 2 + 1) * 32], d);
      float a[4], T[4], w[4];
      const unsigned hb = (rel < A.tile) ? ((mask[rel >> 5] >> (rel & 31)) & 0xFu) : 0u;
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = exp_fast<FTZ>(-s[i] * d[i]);

      // lane aggregate: product of this lane's samples after its last head
      float P = a[0];
#pragma unroll
      for (int i = 1; i < 4; ++i) P = ((hb >> i) & 1u) ? a[i] : P * a[i];
      const unsigned heads = __ballot_sync(kFullMask, hb != 0u);
      if (lane == 0 && !(heads & 1u)) P *= carry;
      P = seg_scan_mul_up(P, absorb_limit_up(heads, lane));
      float E = __shfl_up_sync(kFullMask, P, 1);
      if (lane == 0) E = carry;
      carry = __shfl_sync(kFullMask, P, 31);

      // exclusive transmittance per sample
      T[0] = (hb & 1u) ? 1.f : E;
#pragma unroll
      for (int i = 1; i < 4; ++i) T[i] = ((hb >> i) & 1u) ? 1.f : T[i - 1] * a[i - 1];
      weights_from_T(T, a, thr, w);

      const bool edge = (rpos < t.lo) | (rpos + 128 > t.hi);  // warp-uniform: round not fully owned
      if (exact) {
        // Cheap prefilter, warp-uniform band: |T - thr| <= thr * 1.5*2^-23 * (upper bound on the sample's
        // position in its ray); plus non-monotone (a > 1 / NaN) and underflow cases.
        const float bub = (float)(rpos + 130 - t.lo) * 1.7881393e-07f;
        const float blo = thr - thr * bub, bhi = thr + thr * bub;
        // one compare on the closest approach of the four T to thr (|T - thr| <= thr*bub, the same band up to an ulp
        // of thr; the per-sample test below decides with the exact per-ray bound)
        const float near = fminf(fminf(fabsf(T[0] - thr), fabsf(T[1] - thr)), fminf(fabsf(T[2] - thr), fabsf(T[3] - thr)));
        bool sus = !(fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3])) <= 1.f) | !(near > thr * (bub + 2.4e-7f));
        if (tiny_thr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sus |= T[i] < 7.8886091e-31f;
        }
        if (sus) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = rel + i;
            if (k < t.lo || k >= t.hi) continue;
            bool push = !(a[i] <= 1.f) || (tiny_thr && T[i] < 7.8886091e-31f);
            if (!push && T[i] >= blo && T[i] <= bhi) {
              const int h = find_head_le(mask, k, A.tile);
              const float krel = (float)(k - h + 1);
              push = (h < 0) || (fabsf(T[i] - thr) <= thr * krel * 1.7881393e-07f);  // 1.5 * 2^-23 per factor
            }
            if (push) {
              const int slot = atomicAdd(&s_qn[wib], 1);
              if (slot < kQueue) queue[slot] = k;
            }
          }
        }
      }
      if (MODE != 2 && !edge) st_stream_f4(out + rel, make_float4(w[0], w[1], w[2], w[3]));
      else store_vec<MODE>(out, rel, t.lo, t.hi, w);
    }

    if (exact) {
      __syncwarp();
      const int nq = s_qn[wib];
      if (nq > kQueue) {  // pathological data: redo every ray of the task in reference order
        for (int r = t.r_lo + lane; r < t.r_hi; r += 32) {
          const int2 e = __ldg(&A.info[r]);
          long long st = e.x, en = (long long)e.x + e.y;
          if (e.y > 0 && st >= t.s0 && en <= t.s1) serial_ray_fwd(A.sigmas, A.steps, A.sstride, st, en, thr, A.out);
        }
      } else if (lane < nq) {
        const int b = queue[lane];
        const int hs = find_head_le(mask, b, A.tile);
        const int he = find_head_gt(mask, b, A.tile);
        if (hs >= 0) {
          const long long st = t.tile0 + hs;
          const long long en = (he < 0) ? t.s1 : t.tile0 + he;
          serial_ray_fwd(A.sigmas, A.steps, A.sstride, st, en, thr, A.out);
        }
      }
    }
    __syncwarp();
  }
}

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32, 3) weights_bwd_kernel(const WArgs A) {
  extern __shared__ float4 s_ring_dyn[];  // [kWarps][kDepth][3 arrays][32 lanes]
  __shared__ unsigned s_mask[kWarps][kMaskWords];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned* mask = s_mask[wib];
  const int n_warps = gridDim.x * kWarps;

  for (int task = blockIdx.x * kWarps + wib; task < A.n_tasks; task += n_warps) {
    if (!(A.flags & TNF_W_TRUSTED_PARTITION) && *(volatile unsigned*)A.status) return;
    const Task t = task_setup(A, task, mask, lane);
    if (!t.valid) continue;
    const float* sig = A.sigmas + t.tile0;
    const float* stp = A.steps + t.tile0 * A.sstride;
    const float* wp = A.w + t.tile0;
    const float* gp = A.g + t.tile0;
    float* out = A.out + t.tile0;
    const int first = t.lo & ~127;
    const int last = (t.hi - 1) & ~127;
    const int n_rounds = ((last - first) >> 7) + 1;
    float4* ring = s_ring_dyn + wib * (kDepth * 3 * 32) + lane;  // slot (stage, array) at ring[(stage*3+array)*32]

    // Pass A (descending): S_k = sum_{j>k in ray} w_j*g_j, parked in grad_sigmas.
    {
      float carry = 0.f;
      auto issue = [&](int j) {
        if (j < n_rounds) {
          const int rl = last - j * 128 + lane * 4;
          float4* slot = ring + ((j & (kDepth - 1)) * 3) * 32;
          ring_vec<MODE>(slot, wp, rl, t.lim);
          ring_vec<MODE>(slot + 32, gp, rl, t.lim);
        }
        cp_async_commit();
      };
#pragma unroll
      for (int j = 0; j < kDepth - 1; ++j) issue(j);
      for (int j = 0; j < n_rounds; ++j) {
        const int rpos = last - j * 128;
        const int rel = rpos + lane * 4;
        issue(j + kDepth - 1);
        cp_async_wait_ring();
        float w[4], g[4];
        unpack4(ring[((j & (kDepth - 1)) * 3) * 32], w);
        unpack4(ring[((j & (kDepth - 1)) * 3 + 1) * 32], g);
        const bool edge = (rpos < t.lo) | (rpos + 128 > t.hi);
        float c[4], S[4];
        const unsigned hb5 = head_bits5(mask, rel, A.tile);
        unsigned tail = (hb5 >> 1) & 0xFu;  // bit i: sample rel+i is the last sample of its ray
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = w[i] * g[i];
        if (rpos + 128 >= t.hi) {  // last owned sample closes its ray; samples beyond it contribute nothing
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (rel + i >= t.hi) c[i] = 0.f;
            tail |= (unsigned)(rel + i + 1 >= t.hi) << i;
          }
        }
        // lane aggregate in descending order: inclusive suffix sum of the lane's lowest sample
        float P = c[3];
#pragma unroll
        for (int i = 2; i >= 0; --i) P = ((tail >> i) & 1u) ? c[i] : P + c[i];
        const unsigned tails = __ballot_sync(kFullMask, tail != 0u);
        if (lane == 31 && !(tails >> 31)) P += carry;
        P = seg_scan_add_down(P, absorb_limit_down(tails, lane));
        float E = __shfl_down_sync(kFullMask, P, 1);  // inclusive suffix sum of sample rel+4
        if (lane == 31) E = carry;
        carry = __shfl_sync(kFullMask, P, 0);
        S[3] = ((tail >> 3) & 1u) ? 0.f : E;
#pragma unroll
        for (int i = 2; i >= 0; --i) S[i] = ((tail >> i) & 1u) ? 0.f : S[i + 1] + c[i + 1];
        if (MODE != 2 && !edge) *reinterpret_cast<float4*>(out + rel) = make_float4(S[0], S[1], S[2], S[3]);
        else store_vec<MODE>(out, rel, t.lo, t.hi, S);
      }
    }
    __syncwarp();

    // Pass B (ascending): T_{k+1} = prod_{j<=k in ray} a_j (no termination, src/cuda.cu:52-56);
    // grad_sigma_k = delta_k * (T_{k+1}*g_k - S_k).  Each lane re-reads the S it stored itself.
    {
      float carryT = 1.f;
      auto issue = [&](int j) {
        if (j < n_rounds) {
          const int rl = first + j * 128 + lane * 4;
          float4* slot = ring + ((j & (kDepth - 1)) * 3) * 32;
          ring_vec<MODE>(slot, sig, rl, t.lim);
          ring_steps<MODE>(slot + 32, stp, A.sstride, rl, t.lim);
          ring_vec<MODE>(slot + 64, gp, rl, t.lim);
        }
        cp_async_commit();
      };
      // S is re-read with plain loads (program order after this lane's own stores of pass A), one round ahead
      auto load_S = [&](int rpos, float S[4]) {
        const int rel = rpos + lane * 4;
        const bool edge = (rpos < t.lo) | (rpos + 128 > t.hi);
        if (MODE != 2 && !edge) {
          unpack4(*reinterpret_cast<const float4*>(out + rel), S);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) S[i] = (rel + i >= t.lo && rel + i < t.hi) ? out[rel + i] : 0.f;
        }
      };
#pragma unroll
      for (int j = 0; j < kDepth - 1; ++j) issue(j);
      float S[4], Sn[4];
      load_S(first, S);
      for (int j = 0; j < n_rounds; ++j) {
        const int rpos = first + j * 128;
        const int rel = rpos + lane * 4;
        issue(j + kDepth - 1);
        if (j + 1 < n_rounds) load_S(rpos + 128, Sn);
        cp_async_wait_ring();
        float s[4], d[4], g[4];
        unpack4(ring[((j & (kDepth - 1)) * 3) * 32], s);
        unpack4(ring[((j & (kDepth - 1)) * 3 + 1) * 32], d);
        unpack4(ring[((j & (kDepth - 1)) * 3 + 2) * 32], g);
        const bool edge = (rpos < t.lo) | (rpos + 128 > t.hi);
        float a[4], o[4];
        const unsigned hb = (rel < A.tile) ? ((mask[rel >> 5] >> (rel & 31)) & 0xFu) : 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = exp_fast<true>(-s[i] * d[i]);  // gradients carry a tolerance, not bit parity
        float P = a[0];
#pragma unroll
        for (int i = 1; i < 4; ++i) P = ((hb >> i) & 1u) ? a[i] : P * a[i];
        const unsigned heads = __ballot_sync(kFullMask, hb != 0u);
        if (lane == 0 && !(heads & 1u)) P *= carryT;
        P = seg_scan_mul_up(P, absorb_limit_up(heads, lane));
        float E = __shfl_up_sync(kFullMask, P, 1);
        if (lane == 0) E = carryT;
        carryT = __shfl_sync(kFullMask, P, 31);
        float Tn = ((hb & 1u) ? 1.f : E) * a[0];  // inclusive product T_{k+1}
        o[0] = d[0] * __fmaf_rn(Tn, g[0], -S[0]);
#pragma unroll
        for (int i = 1; i < 4; ++i) {
          Tn = (((hb >> i) & 1u) ? 1.f : Tn) * a[i];
          o[i] = d[i] * __fmaf_rn(Tn, g[i], -S[i]);
        }
        if (MODE != 2 && !edge) st_stream_f4(out + rel, make_float4(o[0], o[1], o[2], o[3]));
        else store_vec<MODE>(out, rel, t.lo, t.hi, o);
#pragma unroll
        for (int i = 0; i < 4; ++i) S[i] = Sn[i];
      }
    }
    __syncwarp();
  }
}

// ---- generic path for packing info that is not a sorted partition (matches the reference for any
// non-overlapping info: thread per ray in reference order, untouched samples are zero) ----------
__global__ void fallback_zero_kernel(const unsigned* status, float* out, long long n) {
  if (!(*status & 1u)) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = 0.f;
}
__global__ void fallback_fwd_kernel(const WArgs A) {
  if (!(*A.status & 1u)) return;
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= A.n_rays) return;
  const int2 e = A.info[r];
  const long long st = e.x, en = (long long)e.x + e.y;
  if (e.y <= 0 || st < 0 || en > A.n) return;
  // reference semantics: nothing written after termination (stays zero from the zero pass)
  float transmittance = 1.f;
  long long k = st;
  while (transmittance > A.thr && k < en) {
    const float alpha = __expf(-A.sigmas[k] * A.steps[k * A.sstride]);
    A.out[k] = transmittance * (1. - alpha);
    transmittance *= alpha;
    ++k;
  }
}
__global__ void fallback_bwd_kernel(const WArgs A) {
  if (!(*A.status & 1u)) return;
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= A.n_rays) return;
  const int2 e = A.info[r];
  const long long st = e.x, en = (long long)e.x + e.y;
  if (e.y <= 0 || st < 0 || en > A.n) return;
  serial_ray_bwd(A.sigmas, A.steps, A.sstride, A.w, A.g, st, en, A.out);
}

int pick_tile(long long n, int warps_per_sm) {
  // Tiles between 256 and kMaxTile samples.  Small inputs get ONE wave: the tile is rounded UP so that the task count
  // does not exceed the resident warps (a second, mostly empty wave would double the latency-bound run time).
  // One task per resident warp and k full waves: the smallest k whose tile fits kMaxTile, then the tile rounded UP
  // so the task count does not spill into a mostly empty extra wave (worth 13% at 2^24 samples).
  const long long slots = (long long)sm_count() * warps_per_sm;
  const long long k = ceil_div(n, slots * kMaxTile);
  long long tile = (ceil_div(n, k * slots) + 127) & ~127LL;
  if (tile < 256) tile = 256;
  if (tile > kMaxTile) tile = kMaxTile;
  return (int)tile;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch(bool bwd, WArgs A, cudaStream_t st) {
  const bool trusted = A.flags & TNF_W_TRUSTED_PARTITION;
  if (!trusted) TNF_CUDA(cudaMemsetAsync(A.status, 0, sizeof(unsigned), st));
  A.tile = pick_tile(A.n, bwd ? 3 * kWarps : 4 * kWarps);
  const long long n_tasks = ceil_div(A.n, A.tile);
  TNF_REQUIRE(n_tasks < (1LL << 30), "too many samples (%lld)", A.n);
  A.n_tasks = (int)n_tasks;
  bool al = aligned16(A.sigmas) && aligned16(A.out);
  if (bwd) al = al && aligned16(A.w) && aligned16(A.g);
  const int mode = !al ? 2 : ((A.sstride == 1 && aligned16(A.steps)) ? 0 : 1);
  const long long ctas = ceil_div(n_tasks, kWarps);  // one task per warp; the block scheduler balances
  const dim3 grid((unsigned)ctas), block(kWarps * 32);
  const size_t ring_fwd = (size_t)kWarps * kDepth * 2 * 32 * sizeof(float4);  // 32 KB
  const size_t ring_bwd = (size_t)kWarps * kDepth * 3 * 32 * sizeof(float4);  // 48 KB
  static PerDeviceOnce configured{};
  if (configured.pending()) {  // the backward ring plus the static arrays exceeds the 48 KB default
    TNF_CUDA(cudaFuncSetAttribute(weights_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bwd));
    TNF_CUDA(cudaFuncSetAttribute(weights_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bwd));
    TNF_CUDA(cudaFuncSetAttribute(weights_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bwd));
    configured.mark();
  }
  if (!bwd) {
    const bool ftz = A.thr >= 1.5777218e-30f;  // 2^-99, the kernel's tiny_thr boundary (see exp_fast)
    if (mode == 0) { if (ftz) weights_fwd_kernel<0, true><<<grid, block, ring_fwd, st>>>(A); else weights_fwd_kernel<0, false><<<grid, block, ring_fwd, st>>>(A); }
    else if (mode == 1) { if (ftz) weights_fwd_kernel<1, true><<<grid, block, ring_fwd, st>>>(A); else weights_fwd_kernel<1, false><<<grid, block, ring_fwd, st>>>(A); }
    else { if (ftz) weights_fwd_kernel<2, true><<<grid, block, ring_fwd, st>>>(A); else weights_fwd_kernel<2, false><<<grid, block, ring_fwd, st>>>(A); }
    TNF_LAUNCH_CHECK("weights_fwd_kernel");
  } else {
    if (mode == 0) weights_bwd_kernel<0><<<grid, block, ring_bwd, st>>>(A);
    else if (mode == 1) weights_bwd_kernel<1><<<grid, block, ring_bwd, st>>>(A);
    else weights_bwd_kernel<2><<<grid, block, ring_bwd, st>>>(A);
    TNF_LAUNCH_CHECK("weights_bwd_kernel");
  }
  if (!trusted) {
    const int zb = (int)(ceil_div(A.n, 256 * 8) < sm_count() * 8 ? ceil_div(A.n, 256 * 8) : sm_count() * 8);
    fallback_zero_kernel<<<zb, 256, 0, st>>>(A.status, A.out, A.n);
    TNF_LAUNCH_CHECK("fallback_zero_kernel");
    const unsigned rb = (unsigned)ceil_div(A.n_rays, 128);
    if (!bwd) fallback_fwd_kernel<<<rb, 128, 0, st>>>(A);
    else fallback_bwd_kernel<<<rb, 128, 0, st>>>(A);
    TNF_LAUNCH_CHECK("fallback_ray_kernel");
  }
  return TNF_OK;
}

int check_common(const float* sigmas, const float* steps, int64_t stride, const int32_t* info,
                 const float* out, int64_t n, int64_t r, int flags, const uint32_t* status) {
  TNF_REQUIRE(n >= 0 && r >= 0, "negative size (n_samples=%lld, n_rays=%lld)", (long long)n, (long long)r);
  TNF_REQUIRE(n < (1LL << 31) && r < (1LL << 31), "sizes must fit int32 (packing info is int32)");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(sigmas && steps && out, "null sample pointer");
  TNF_REQUIRE(stride >= 1, "steps_stride must be >= 1");
  TNF_REQUIRE(r == 0 || info, "null info pointer");
  TNF_REQUIRE((reinterpret_cast<uintptr_t>(info) & 7u) == 0, "info must be 8-byte aligned");
  TNF_REQUIRE((flags & TNF_W_TRUSTED_PARTITION) || status, "status word required unless TNF_W_TRUSTED_PARTITION");
  return TNF_OK;
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_weights_fwd(const float* sigmas, const float* steps, int64_t steps_stride,
                               const int32_t* info, float threshold, float* weights, int64_t n_samples,
                               int64_t n_rays, int flags, uint32_t* status, void* stream) {
  using namespace tnf;
  int rc = check_common(sigmas, steps, steps_stride, info, weights, n_samples, n_rays, flags, status);
  if (rc != TNF_OK || n_samples == 0) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_rays == 0) {  // nothing covers the samples: all zero (src/cuda.cu:84)
    TNF_CUDA(cudaMemsetAsync(weights, 0, sizeof(float) * n_samples, st));
    return TNF_OK;
  }
  WArgs A{};
  A.sigmas = sigmas; A.steps = steps; A.sstride = steps_stride;
  A.info = reinterpret_cast<const int2*>(info);
  A.thr = threshold; A.out = weights; A.n = n_samples; A.n_rays = (int)n_rays;
  A.status = status; A.flags = flags;
  return launch(false, A, st);
}

extern "C" int tnf_weights_bwd(const float* sigmas, const float* steps, int64_t steps_stride,
                               const int32_t* info, const float* weights, const float* grad_weights,
                               float* grad_sigmas, int64_t n_samples, int64_t n_rays, int flags,
                               uint32_t* status, void* stream) {
  using namespace tnf;
  int rc = check_common(sigmas, steps, steps_stride, info, grad_sigmas, n_samples, n_rays, flags, status);
  if (rc != TNF_OK || n_samples == 0) return rc;
  TNF_REQUIRE(weights && grad_weights, "null weights/grad_weights pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_rays == 0) {
    TNF_CUDA(cudaMemsetAsync(grad_sigmas, 0, sizeof(float) * n_samples, st));
    return TNF_OK;
  }
  WArgs A{};
  A.sigmas = sigmas; A.steps = steps; A.sstride = steps_stride;
  A.info = reinterpret_cast<const int2*>(info);
  A.out = grad_sigmas; A.w = weights; A.g = grad_weights; A.n = n_samples; A.n_rays = (int)n_rays;
  A.status = status; A.flags = flags;
  return launch(true, A, st);
}
