// Microbenchmarks behind the K-Planes scatter design: what do 128-byte-line reductions cost
// at L2 (red.global.add.v4.f32 by 8 lanes, scalar red by 32 lanes) against a cache-resident and a DRAM-sized buffer,
// what does the same accumulation cost in shared memory (fp32 atomicAdd = ATOMS.CAST.SPIN loop on sm_100a), and what
// does the gather side cost.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics_bench atomics_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// mode 0: 8 lanes x red.v4 per line; mode 1: 32 lanes x scalar red per line; mode 2: 8 lanes x LDG.128 per line (gather)
// each "item" touches `per` random lines of a buffer of n_lines lines; hot: fraction of touches that go to a 1/64 subset
template <int MODE>
__global__ void l2_kernel(float* buf, unsigned n_lines, unsigned n_items, int per, float* sink) {
  const unsigned gt = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (MODE == 1) {
    const unsigned item = gt >> 5, lane = gt & 31;
    if (item >= n_items) return;
    for (int k = 0; k < per; ++k) {
      const unsigned line = hash32(item * 64 + k) % n_lines;
      atomicAdd(buf + (size_t)line * 32 + lane, 1.f);
    }
  } else {
    const unsigned item = gt >> 3, l = gt & 7;
    if (item >= n_items) return;
    for (int k = 0; k < per; ++k) {
      const unsigned line = hash32(item * 64 + k) % n_lines;
      float* p = buf + (size_t)line * 32 + l * 4;
      if (MODE == 0) red4(p, make_float4(1.f, 1.f, 1.f, 1.f));
      else { float4 v = __ldg(reinterpret_cast<const float4*>(p)); acc += v.x + v.y + v.z + v.w; }
    }
  }
  if (acc == 1.2345e-30f) *sink = acc;
}

// shared-memory accumulation: a CTA owns a tile of `rows` lines (128 B each); every warp adds `per` random rows per item
// with lane = channel (conflict-free banks).  MODE 0: atomicAdd (CAS loop), MODE 1: plain load/add/store (races ignored:
// upper bound of a race-free ownership scheme), MODE 2: red.shared via 8 lanes x 4 scalar atomics (v4-like layout)
template <int MODE>
__global__ void smem_kernel(float* out, int rows, int items_per_warp, int per) {
  extern __shared__ float tile[];
  for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int it = 0; it < items_per_warp; ++it) {
    const unsigned base = (blockIdx.x * 64 + warp) * 4096 + it;
#pragma unroll 4
    for (int k = 0; k < per; ++k) {
      const unsigned row = hash32(base * 16 + k) % rows;
      float* p = tile + row * 32 + lane;
      if (MODE == 0) atomicAdd(p, 1.f);
      else if (MODE == 1) *p += 1.f;
    }
  }
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) s += tile[i];
  if (s == 1.2345e-30f) out[0] = s;
}

template <typename F>
float time_us(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best * 1e3f;
}

int main() {
  float* sink; cudaMalloc(&sink, 4);
  const unsigned n_items = 1u << 18; const int per = 36;
  for (size_t mb : {6, 24, 96, 400}) {
    const unsigned n_lines = (unsigned)(mb * 1000000 / 128);
    float* buf; cudaMalloc(&buf, (size_t)n_lines * 128); cudaMemset(buf, 0, (size_t)n_lines * 128);
    const double bytes = (double)n_items * per * 128;
    float t0 = time_us([&] { l2_kernel<0><<<(n_items * 8 + 255) / 256, 256>>>(buf, n_lines, n_items, per, sink); });
    float t1 = time_us([&] { l2_kernel<1><<<(n_items * 32 + 255) / 256, 256>>>(buf, n_lines, n_items, per, sink); });
    float t2 = time_us([&] { l2_kernel<2><<<(n_items * 8 + 255) / 256, 256>>>(buf, n_lines, n_items, per, sink); });
    printf("buffer %4zu MB: red.v4 x8 lanes %7.1f us (%.2f TB/s) | scalar red x32 lanes %7.1f us (%.2f TB/s) | LDG.128 gather %7.1f us (%.2f TB/s)\n",
           mb, t0, bytes / t0 / 1e6, t1, bytes / t1 / 1e6, t2, bytes / t2 / 1e6);
    cudaFree(buf);
  }
  // shared memory: 148 CTAs x 16 warps, tile of 867 rows (3 planes x 17^2 texels), 9.4 M row updates in total
  const int rows = 867, warps = 16, ctas = 148;
  const int items = 9437184 / 12 / (warps * ctas);   // warp-items of 12 rows each
  cudaFuncSetAttribute(smem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 128);
  cudaFuncSetAttribute(smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 128);
  float s0 = time_us([&] { smem_kernel<0><<<ctas, warps * 32, rows * 128>>>(sink, rows, items, 12); });
  float s1 = time_us([&] { smem_kernel<1><<<ctas, warps * 32, rows * 128>>>(sink, rows, items, 12); });
  const double upd = (double)items * 12 * warps * ctas;
  printf("shared-memory tile (%d rows, %d CTAs x %d warps): %.1f M row updates: atomicAdd %7.1f us (%.2f clk/row/SM @1.9GHz) | plain RMW %7.1f us\n",
         rows, ctas, warps, upd / 1e6, s0, s0 * 1.9e3 / (upd / ctas), s1);
  return 0;
}
