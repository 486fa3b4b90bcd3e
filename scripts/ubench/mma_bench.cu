// Micro-benchmark: cycles per tcgen05.mma.kind::tf32 for the operand layouts the MLP kernels use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu && ./mma_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, int ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, int ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"((uint32_t)acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem),
               "l"(b), "r"(idesc), "r"((uint32_t)acc)
               : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint64_t desc_k(uint32_t atom, int kk) { return smem_desc(atom + kk * 32, 16, 1024, 2); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t atom, int kk, uint32_t lbo) { return smem_desc(atom + kk * 1024, lbo, 512, 1); }
__device__ __forceinline__ uint32_t idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

struct Cfg { int M, N, a_mode /*0 K-major smem, 1 MN-major smem, 2 TMEM*/, b_mn, reps, indep, style /*0 thread0, 1 warp0+elect*/, fixed_k, f16; };
__device__ __forceinline__ bool elect_one() { uint32_t p; asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p)); return p != 0; }
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}


__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((float*)smem)[i] = 1.0f + (i % 7) * 0.125f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) tmem_alloc(&s_tmem, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tmem;
  const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 80 * 1024);
  // kind::f16: bf16 A/B (format 1), fp32 D: idesc bits: c_format=1<<4, a_format=1<<7, b_format=1<<10
  const uint32_t id = c.f16 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24))
                            : idesc(c.M, c.N, c.a_mode == 1, c.b_mn);
  if (c.style == 0) {
    if (threadIdx.x == 0) {
      for (int round = 0; round < 3; ++round) {
        const long long t0 = clock64();
        for (int r = 0; r < c.reps; ++r) {
          const int kk = c.fixed_k ? 0 : (r & 3);
          const uint32_t d = tm + (c.indep ? (r & 1) * 256 : 0);
          uint64_t bd = c.b_mn ? desc_mn(b0, kk, 16384) : desc_k(b0, kk);
          if (c.a_mode == 2) mma_ts(d, tm + 448 + kk * 8, bd, id, r > 1);
          else mma_ss(d, c.a_mode == 1 ? desc_mn(a0, kk, 16384) : desc_k(a0, kk), bd, id, r > 1);
        }
        const long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, round & 1);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  } else if (threadIdx.x < 32) {
    for (int round = 0; round < 3; ++round) {
      const long long t0 = clock64();
      if (elect_one()) {
        if (c.f16) {
#pragma unroll 4
          for (int r = 0; r < c.reps; ++r) mma_f16(tm, desc_k(a0, r & 3), desc_k(b0, r & 3), id, r > 1);
        } else if (c.a_mode == 0 && !c.b_mn) {
#pragma unroll 4
          for (int r = 0; r < c.reps; ++r) mma_ss(tm, desc_k(a0, c.fixed_k ? 0 : (r & 3)), desc_k(b0, c.fixed_k ? 0 : (r & 3)), id, r > 1);
        } else if (c.a_mode == 1) {
#pragma unroll 4
          for (int r = 0; r < c.reps; ++r) mma_ss(tm, desc_mn(a0, r & 3, 16384), desc_mn(b0, r & 3, 16384), id, r > 1);
        } else {
#pragma unroll 4
          for (int r = 0; r < c.reps; ++r) mma_ts(tm, tm + 448 + (r & 3) * 8, desc_k(b0, r & 3), id, r > 1);
        }
      }
      __syncwarp();
      const long long t1 = clock64();
      if (elect_one()) mma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, round & 1);
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const Cfg cfgs[] = {
      {128, 64, 0, 0, 96, 0, 0, 0, 0}, {128, 64, 0, 0, 96, 0, 0, 1, 0}, {128, 64, 0, 0, 96, 0, 1, 0, 0}, {128, 64, 0, 0, 96, 0, 1, 1, 0},
      {128, 64, 1, 1, 96, 0, 1, 0, 0}, {64, 32, 1, 1, 96, 0, 1, 0, 0},  {128, 64, 2, 0, 96, 0, 1, 0, 0}, {128, 256, 0, 0, 96, 0, 1, 0, 0},
      {128, 64, 0, 0, 96, 0, 1, 0, 1}, {128, 256, 0, 0, 96, 0, 1, 0, 1}, {128, 128, 0, 0, 96, 0, 1, 0, 0}, {128, 160, 0, 0, 96, 0, 1, 0, 0},
      {64, 64, 0, 0, 96, 0, 1, 0, 0}, {128, 64, 0, 0, 384, 0, 1, 0, 0}, {128, 32, 0, 0, 96, 0, 1, 0, 0}, {128, 16, 0, 0, 96, 0, 1, 0, 0},
  };
  for (const Cfg& c : cfgs) {
    bench<<<1, 128, 170 * 1024>>>(c, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d A=%s B=%s indep=%d style=%d fixk=%d f16=%d reps=%d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (floor %5.1f)  %s\n", c.M, c.N,
           c.a_mode == 0 ? "K " : (c.a_mode == 1 ? "MN" : "TM"), c.b_mn ? "MN" : "K ", c.indep, c.style, c.fixed_k, c.f16, c.reps, (double)h[0] / c.reps, (double)h[1] / c.reps,
           (c.M > 128 ? c.M : 128) * c.N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
