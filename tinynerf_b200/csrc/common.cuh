// Shared helpers for the tinynerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tinynerf_b200.h"

namespace tnf {

constexpr unsigned kFullMask = 0xffffffffu;

// Thread-local error message, set by every failing entry point.
void set_error(const char* fmt, ...);

#define TNF_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      ::tnf::set_error(__VA_ARGS__);            \
      return TNF_E_INVALID;                     \
    }                                           \
  } while (0)

#define TNF_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::tnf::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),      \
                       __FILE__, __LINE__);                                         \
      return TNF_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

// After a kernel launch: surface launch-configuration errors (the reference never checks, cuda.cu:86).
#define TNF_LAUNCH_CHECK(name)                                                      \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      ::tnf::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));    \
      return TNF_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

int sm_count();  // cached per device
// Diagnostic kernel variants, selected per calling thread with tnf_set_variant (never from the environment).
constexpr int kVariantWgradSS = 0;   // 1: both-operands-in-shared-memory weight-gradient kernel
constexpr int kVariantTvTexel = 1;   // 1: one-thread-per-texel TV kernel
constexpr int kVariantKplanesOcc = 2; // K-Planes lookup kernels built for another occupancy (value = blocks per SM, 1 = uncapped; 0 = default)
constexpr int kVariantNoWstat = 3;    // 1: 128x128 layers through linear_kernel (weights in shared memory) instead of the weight-stationary kernel
constexpr int kVariantCount = 4;
int variant(int which);

// Layers with 128 outputs and <= 128 inputs with the weights stationary in tensor memory (wstat.cu); mode 0 = forward
// (x [m,k] -> y [m,128], bias, optional ReLU), 1 = data gradient (x = dy [m,128] -> y = dx [m,k], optional ReLU mask [m,k])
bool wstat_linear_supported(int mode, int64_t m, int n, int k, const void* x, int64_t ldx, const void* w, const void* y, int64_t ldy,
                            const void* mask, int64_t ldmask);
int launch_wstat_linear(int mode, const float* x, int64_t ldx, const float* w, int k, const float* bias, int relu, const float* mask,
                        int64_t ldmask, float* y, int64_t ldy, int64_t m, cudaStream_t st);

// Function attributes (cudaFuncSetAttribute) belong to the device/context, not to the calling thread: a call site keeps one
// flag per device and opts in again the first time it runs on another GPU of the process.  The flag is set after the
// attribute call, so a racing thread at worst repeats an idempotent call.
struct PerDeviceOnce {
  unsigned char done[64];
  int dev() const { int d = 0; return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) ? d : -1; }
  bool pending() const { const int d = dev(); return d < 0 || !done[d]; }
  void mark() { const int d = dev(); if (d >= 0) done[d] = 1; }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------------------------

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Streaming (read-once) 128-bit load: keep it out of L1 so gather data can live there.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// Streaming store (no L1 allocation).
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Vector float atomic add without return (sm_90+): one L2 reduction op for 16 bytes.
__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Philox4x32-10 (same generator family as torch's CUDA RNG; used only in perf mode where the
// caller does not inject noise).
struct Philox {
  uint32_t key0, key1;
  __device__ __forceinline__ Philox(uint64_t seed) : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi,
             c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ k0;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ k1;
      c3 = lo0;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// 24-bit mantissa uniform in [0,1) (torch's float path: (x & ((1<<24)-1)) * 2^-24).
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x & 0xFFFFFFu) * 5.9604644775390625e-08f; }

}  // namespace tnf
