// (e) + f1 -- the data-parallel parameter path over NVLink peer memory, one kernel per step instead of
// all-reduce + replicated optimiser (the reference has a single device: src/run.py:98; its update is
// optimizer.step() of torch.optim.Adam at src/run.py:186,258-261).
//
// Every rank keeps its flat gradient buffer and its flat parameter buffer in symmetric memory (same
// offsets on every GPU, mapped into every peer; where the fabric supports it also bound to an NVSwitch
// multicast address).  tnf_dp_reduce_adam_bcast, launched by every rank on its own slice [lo, hi) of the
// flat parameter space, does in ONE pass:
//     start barrier      every rank has finished writing its gradients (and reading its parameters)
//     g   = sum over ranks of grad[i]          multimem.ld_reduce (reduced inside the switch) or P2P loads
//     p,m,v <- Adam(p, g, m, v)                in registers; m, v exist only on the owning rank
//     p  -> every rank                         multimem.st (replicated by the switch) or P2P stores
//     end barrier        every rank's slice has landed everywhere; gradients may be overwritten
// i.e. reduce-scatter + sharded Adam + all-gather without NCCL's channel CTAs, without a second pass over
// the parameters, and with 1/world of the optimiser's HBM traffic per rank.  It runs on a side stream
// beside the heads' weight-gradient kernels (its CTAs need no shared memory and few registers, so they
// co-reside with the tensor-core kernels' one CTA per SM).
//
// Barriers: CTA b of every rank pairs with CTA b of every peer through a flag pad in symmetric memory
// ([cta][source rank] words holding an epoch counter; st.release.sys / ld.acquire.sys).  All CTAs of a
// launch must be co-resident (grid <= SM count, enforced), so no CTA waits on one that cannot start.
// A wait that exceeds ~4 s sets *error and gives up instead of hanging the device.
//
// The per-step ray count of the union batch (the MSE normaliser when ranks hold different ray counts,
// SURVEY 8e) travels the same way: tnf_dp_publish_count stores {step, count} as one 64-bit word into
// every peer's slot table, tnf_dp_sum_counts waits for the world's words of that step and writes the
// float total the loss kernel reads -- two single-warp launches instead of a NCCL all-reduce whose
// spinning CTAs displace the forward's one-CTA-per-SM kernels.
#include "common.cuh"

namespace tnf {
namespace {

constexpr int kMaxRanks = TNF_DP_MAX_RANKS;
constexpr long long kSpinLimit = 8000000000LL;   // clock64 ticks (~4 s at 1.9 GHz)

struct DpArgs {
  const float* mc_grad;              // multicast address of the flat gradients (nullptr: P2P loads)
  float* mc_param;                   // multicast address of the flat parameters (nullptr: P2P stores)
  const float* peer_grad[kMaxRanks];
  float* peer_param[kMaxRanks];
  uint32_t* peer_flags[kMaxRanks];   // [n_ctas][kMaxRanks] per rank, at flag_offset
  float* m;                          // local, indexed like the flat parameter space
  float* v;
  long long lo, hi;                  // slice of this rank (multiples of 4)
  int rank, world;
  uint32_t epoch;                    // start barrier waits for `epoch`, end barrier for `epoch + 1`
  int* error;                        // local device word, set non-zero when a barrier timed out
  float lr, beta1, beta2, eps, weight_decay, bias1, bias2_sqrt;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// CTA-pair barrier across the ranks.  Every thread's earlier writes are ordered before the flag stores by the
// system-scope fence + bar.sync; the acquiring loads order the peers' writes before everything after the barrier.
__device__ __forceinline__ void rank_barrier(const DpArgs& A, uint32_t epoch) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < A.world) {
    const int peer = threadIdx.x;
    st_release_sys(A.peer_flags[peer] + (size_t)blockIdx.x * kMaxRanks + A.rank, epoch);
    const uint32_t* mine = A.peer_flags[A.rank] + (size_t)blockIdx.x * kMaxRanks + peer;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > kSpinLimit) { atomicExch(A.error, 1); break; }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const DpArgs& A) {
  // identical arithmetic to adam_one (tv_adam.cu): torch/optim/adam.py with L2 weight decay, no amsgrad
  g = __fmaf_rn(p, A.weight_decay, g);
  m = m + (1.f - A.beta1) * (g - m);
  v = __fmaf_rn((1.f - A.beta2) * g, g, v * A.beta2);
  const float denom = sqrtf(v) / A.bias2_sqrt + A.eps;
  p = p - (A.lr / A.bias1) * (m / denom);
}

constexpr int kDpThreads = 256;
constexpr int kDpUnroll = 4;     // independent float4 per thread in flight

template <bool MULTIMEM>
__global__ void __launch_bounds__(kDpThreads) dp_reduce_adam_bcast_kernel(const DpArgs A) {
  rank_barrier(A, A.epoch);
  const float* param = A.peer_param[A.rank];
  const long long stride = (long long)gridDim.x * kDpThreads * 4;
  for (long long base = A.lo + ((long long)blockIdx.x * kDpThreads + threadIdx.x) * 4; base < A.hi; base += stride * kDpUnroll) {
    float4 g[kDpUnroll], p[kDpUnroll], m[kDpUnroll], v[kDpUnroll];
#pragma unroll
    for (int u = 0; u < kDpUnroll; ++u) {
      const long long i = base + u * stride;
      if (i < A.hi) {
        if (MULTIMEM) {
          g[u] = multimem_ld_reduce_f4(A.mc_grad + i);
        } else {
          g[u] = ld_relaxed_sys_f4(A.peer_grad[0] + i);
          for (int r = 1; r < A.world; ++r) {
            const float4 t = ld_relaxed_sys_f4(A.peer_grad[r] + i);
            g[u].x += t.x; g[u].y += t.y; g[u].z += t.z; g[u].w += t.w;
          }
        }
        p[u] = *reinterpret_cast<const float4*>(param + i);
        m[u] = *reinterpret_cast<const float4*>(A.m + i);
        v[u] = *reinterpret_cast<const float4*>(A.v + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kDpUnroll; ++u) {
      const long long i = base + u * stride;
      if (i < A.hi) {
        adam_update(p[u].x, g[u].x, m[u].x, v[u].x, A); adam_update(p[u].y, g[u].y, m[u].y, v[u].y, A);
        adam_update(p[u].z, g[u].z, m[u].z, v[u].z, A); adam_update(p[u].w, g[u].w, m[u].w, v[u].w, A);
        *reinterpret_cast<float4*>(A.m + i) = m[u];
        *reinterpret_cast<float4*>(A.v + i) = v[u];
        if (MULTIMEM) {
          multimem_st_f4(A.mc_param + i, p[u]);
        } else {
          for (int r = 0; r < A.world; ++r) st_relaxed_sys_f4(A.peer_param[r] + i, p[u]);
        }
      }
    }
  }
  rank_barrier(A, A.epoch + 1);
}

// ---- ray count of the union batch --------------------------------------------------------------
struct SlotTable { unsigned long long* p[kMaxRanks]; };
__global__ void dp_publish_count_kernel(const SlotTable T, int world, int rank, int slot, uint32_t step, float count) {
  const int peer = threadIdx.x;
  if (peer >= world) return;
  const unsigned long long word = ((unsigned long long)step << 32) | (unsigned long long)__float_as_uint(count);
  unsigned long long* dst = T.p[peer] + (size_t)slot * kMaxRanks + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
}
__global__ void dp_sum_counts_kernel(const unsigned long long* slots, int world, int slot, uint32_t step, float* out, int* error) {
  __shared__ float vals[kMaxRanks];
  const int peer = threadIdx.x;
  if (peer < world) {
    const unsigned long long* src = slots + (size_t)slot * kMaxRanks + peer;
    unsigned long long w;
    const long long t0 = clock64();
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
      if ((uint32_t)(w >> 32) == step) break;
      if (clock64() - t0 > kSpinLimit) { atomicExch(error, 2); w = 0; break; }
    }
    vals[peer] = __uint_as_float((uint32_t)w);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += vals[r];   // ray counts are integers < 2^24: the float sum is exact
    *out = s;
  }
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_dp_reduce_adam_bcast(const float* const* peer_grads, float* const* peer_params, const float* mc_grad,
                                        float* mc_param, float* exp_avg, float* exp_avg_sq, int64_t lo, int64_t hi,
                                        uint32_t* const* peer_flags, int32_t n_ctas, int32_t rank, int32_t world,
                                        uint32_t epoch, int32_t* error, float lr, float beta1, float beta2, float eps,
                                        float weight_decay, int64_t step, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad rank/world (%d/%d, at most %d ranks)", rank, world, kMaxRanks);
  TNF_REQUIRE(peer_grads && peer_params && peer_flags && exp_avg && exp_avg_sq && error, "null pointer");
  TNF_REQUIRE(lo >= 0 && hi >= lo && (lo & 3) == 0 && (hi & 3) == 0, "slice [%lld,%lld) must be 4-element aligned", (long long)lo, (long long)hi);
  TNF_REQUIRE(step >= 1, "step must be >= 1");
  TNF_REQUIRE((mc_grad == nullptr) == (mc_param == nullptr), "multicast addresses must be given for both buffers or neither");
  // CTA b of a rank waits only for CTA b of its peers, and the hardware dispatches CTAs in index order: a CTA that is not
  // resident yet is preceded, on every rank, by CTAs whose partners are resident or will be -- no cycle.  The cap keeps the
  // flag pad ([4 * SM count][TNF_DP_MAX_RANKS]) in bounds.
  TNF_REQUIRE(n_ctas >= 1 && n_ctas <= 4 * sm_count(), "n_ctas must be in [1, 4 x SM count]");
  DpArgs A{};
  for (int r = 0; r < world; ++r) {
    TNF_REQUIRE(peer_grads[r] && peer_params[r] && peer_flags[r], "null peer pointer (rank %d)", r);
    TNF_REQUIRE(((reinterpret_cast<uintptr_t>(peer_grads[r]) | reinterpret_cast<uintptr_t>(peer_params[r])) & 15u) == 0, "peer buffers must be 16-byte aligned");
    A.peer_grad[r] = peer_grads[r]; A.peer_param[r] = peer_params[r]; A.peer_flags[r] = peer_flags[r];
  }
  TNF_REQUIRE(((reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq) | reinterpret_cast<uintptr_t>(mc_grad) |
                reinterpret_cast<uintptr_t>(mc_param)) & 15u) == 0, "buffers must be 16-byte aligned");
  A.mc_grad = mc_grad; A.mc_param = mc_param; A.m = exp_avg; A.v = exp_avg_sq;
  A.lo = lo; A.hi = hi; A.rank = rank; A.world = world; A.epoch = epoch; A.error = error;
  A.lr = lr; A.beta1 = beta1; A.beta2 = beta2; A.eps = eps; A.weight_decay = weight_decay;
  A.bias1 = (float)(1.0 - pow((double)beta1, (double)step));
  A.bias2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mc_grad) dp_reduce_adam_bcast_kernel<true><<<(unsigned)n_ctas, kDpThreads, 0, st>>>(A);
  else dp_reduce_adam_bcast_kernel<false><<<(unsigned)n_ctas, kDpThreads, 0, st>>>(A);
  TNF_LAUNCH_CHECK("dp_reduce_adam_bcast_kernel");
  return TNF_OK;
}

extern "C" int tnf_dp_publish_count(uint64_t* const* peer_slots, int32_t world, int32_t rank, int32_t slot, uint32_t step,
                                    float count, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world && slot >= 0 && slot < TNF_DP_COUNT_SLOTS, "bad rank/world/slot");
  TNF_REQUIRE(peer_slots, "null slot table");
  SlotTable T{};
  for (int r = 0; r < world; ++r) {
    TNF_REQUIRE(peer_slots[r], "null peer slot pointer (rank %d)", r);
    T.p[r] = reinterpret_cast<unsigned long long*>(peer_slots[r]);
  }
  dp_publish_count_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(T, world, rank, slot, step, count);
  TNF_LAUNCH_CHECK("dp_publish_count_kernel");
  return TNF_OK;
}

extern "C" int tnf_dp_sum_counts(const uint64_t* slots, int32_t world, int32_t slot, uint32_t step, float* out, int32_t* error,
                                 void* stream) {
  using namespace tnf;
  TNF_REQUIRE(world >= 1 && world <= kMaxRanks && slot >= 0 && slot < TNF_DP_COUNT_SLOTS && slots && out && error, "bad arguments");
  dp_sum_counts_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const unsigned long long*>(slots), world, slot, step, out, error);
  TNF_LAUNCH_CHECK("dp_sum_counts_kernel");
  return TNF_OK;
}
