// tcgen05 / TMEM / mbarrier / cp.async building blocks shared by the tensor-core kernels (mlp.cu, heads.cu).
#pragma once
#include "common.cuh"

namespace tnf {
namespace {


constexpr int kThreads = 128;
constexpr int kAtomBytes = 128 * 128;        // 128 rows x 32 fp32
constexpr int kMaxKAtoms = 5;                // K <= 160

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, int ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, int ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"((uint32_t)accumulate)
      : "memory");
}
// One lane of a converged warp, chosen by the hardware.  Unlike `lane == 0` the compiler keeps the code under it on the
// uniform datapath (descriptors in uniform registers, no R2UR + waterfall loop per tcgen05.mma): measured 48 instead of
// 147 cycles per 128x64x8 tf32 MMA (scripts/ubench/mma_bench.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32-bit, 32 consecutive columns per thread (thread i of warp w reads TMEM lane 32*(w%4)+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// two 4-column reads (thread i of warp w: TMEM lane 32*(w%4)+i, columns [c, c+4) of each address), one wait
__device__ __forceinline__ void tmem_ld4x2(uint32_t taddr_a, uint32_t taddr_b, float a[4], float b[4]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr_a) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr_b) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[4 + i]); }
}

// ---- descriptors (cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor) -------------------
// shared-memory matrix descriptor, version 1 (Blackwell); layout_type 2 = SWIZZLE_128B (16-byte chunks XOR row%8),
// 1 = SWIZZLE_128B_BASE32B (32-byte chunks XOR row%4) -- the only layout available to MN-major tf32 operands
// (cutlass/gemm/collective/builders/sm100_common.inl:92).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | ((uint64_t)layout_type << 61);
}
// K-major operand: 8-row groups are 1024 B apart (SBO); LBO is ignored for swizzled K-major layouts.
// k-step kk (8 tf32 = 32 B) advances the start address inside the 128-byte row.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t atom_saddr, int kk) { return smem_desc(atom_saddr + kk * 32, 16, 1024); }
// MN-major tf32 operand (SWIZZLE_128B_BASE32B): MN blocks of 32 elements are `lbo` bytes apart, 4-row K groups
// 512 B apart (SBO); k-step kk (8 rows) advances the start address by 1024 bytes.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t atom_saddr, int kk, uint32_t lbo = kAtomBytes) {
  return smem_desc(atom_saddr + kk * 1024, lbo, 512, 1);
}
// instruction descriptor: D fp32, A/B tf32, dense
__device__ __forceinline__ uint32_t instr_desc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- operand staging --------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
#ifdef TNF_MASK_TF32
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = __uint_as_float(__float_as_uint(x - hi) & 0xFFFFE000u);
#else
  // kind::tf32 reads the top 19 bits of each 32-bit container (truncation): the raw fp32 value IS the hi operand and
  // the residual x - trunc(x) (exact in fp32) the lo operand, no masking needed (checked against the masked form by
  // tests/test_gpu_mlp.py: identical results)
  hi = x;
  lo = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
#endif
}
// Operand staging is split in two so the global loads of the NEXT atom can be in flight (in registers) while the
// tensor core works on the current one: load_atom_regs issues the 8 coalesced 128-bit loads of a thread,
// store_atom_regs splits hi/lo and writes the swizzled shared-memory images.
// Atom = rows [row0, row0+128) x cols [col0, col0+32) of a row-major matrix (leading dimension ld, `rows` x `cols`
// valid, zero elsewhere).  Thread t covers 16-byte chunk t%8 of rows t/8 + 16 i.
__device__ __forceinline__ void load_atom_regs(const float* __restrict__ g, long long ld, long long row0, long long rows,
                                               int col0, int cols, int tid, float4 v[8], int atom_rows = 128) {
  const int c = tid & 7;
  const int r0 = tid >> 3;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(g) & 15u) == 0);
  const int col = col0 + 4 * c;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + 16 * i;
    const long long row = row0 + r;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < atom_rows && row < rows && col < cols) {
      const float* p = g + row * ld + col;
      if (col + 3 < cols && vec) {
        v[i] = __ldg(reinterpret_cast<const float4*>(p));
      } else if (col + 3 < cols) {
        v[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
      } else {
        v[i].x = __ldg(p);
        if (col + 1 < cols) v[i].y = __ldg(p + 1);
        if (col + 2 < cols) v[i].z = __ldg(p + 2);
      }
    }
  }
}
// Asynchronous variant: the same per-thread 16-byte chunks are copied global -> shared with cp.async into a raw
// ring slot (zero-filled outside the matrix), so several atoms per CTA are in flight without holding registers;
// each thread later reads back exactly the chunks it copied (no cross-thread hazard), splits and writes the operand
// images.  Requires 16-byte aligned rows (ld % 4 == 0, aligned base); otherwise callers use load_atom_regs.
__device__ __forceinline__ void cp_async_atom(const float* __restrict__ g, long long ld, long long row0, long long rows,
                                              int col0, int cols, int tid, uint8_t* raw) {
  const int c = tid & 7;
  const int r0 = tid >> 3;
  const int col = col0 + 4 * c;
  int nbytes = (cols - col) * 4;
  nbytes = nbytes < 0 ? 0 : (nbytes > 16 ? 16 : nbytes);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + 16 * i;
    const long long row = row0 + r;
    const bool ok = row < rows && nbytes > 0;
    const float* src = ok ? (g + row * ld + col) : g;
    const uint32_t dst = smem_u32(raw + r * 128 + c * 16);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? nbytes : 0) : "memory");
  }
}
// Same copy, but straight into an operand image: chunk (r, c) lands at its swizzled position.  The raw fp32 tile IS the
// "hi" operand (kind::tf32 ignores the low 13 mantissa bits), so the global->shared copy needs no register pass at all;
// only the "lo" image (x - trunc(x)) is computed by the loader threads, each from the chunks it copied itself.
template <int NT = 128>  // NT loader threads share the atom: thread t covers chunk t%8 of rows t/8 + (NT/8) i
__device__ __forceinline__ void cp_async_atom_swz(const float* __restrict__ g, long long ld, long long row0, long long rows,
                                                  int col0, int cols, int tid, uint8_t* img, bool mn32) {
  const int c = tid & 7;
  const int r0 = tid >> 3;
  const int col = col0 + 4 * c;
  int nbytes = (cols - col) * 4;
  nbytes = nbytes < 0 ? 0 : (nbytes > 16 ? 16 : nbytes);
#pragma unroll
  for (int i = 0; i < 1024 / NT; ++i) {
    const int r = r0 + (NT / 8) * i;
    const long long row = row0 + r;
    const bool ok = row < rows && nbytes > 0;
    const float* src = ok ? (g + row * ld + col) : g;
    const uint32_t dst = smem_u32(img + (mn32 ? (r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4))) : (r * 128 + ((c ^ (r & 7)) << 4))));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? nbytes : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int pending) {  // pending in 0..6
  if (pending <= 0) cp_async_wait<0>(); else if (pending == 1) cp_async_wait<1>(); else if (pending == 2) cp_async_wait<2>();
  else if (pending == 3) cp_async_wait<3>(); else if (pending == 4) cp_async_wait<4>(); else if (pending == 5) cp_async_wait<5>();
  else cp_async_wait<6>();
}
// lo image of an atom from its hi image: every thread handles the 8 chunks it copied itself (no cross-thread hazard)
template <int NT = 128>
__device__ __forceinline__ void make_lo_atom(const uint8_t* hi_img, uint8_t* lo_img, int tid, bool mn32, float colsum[4]) {
  const int c = tid & 7;
  const int r0 = tid >> 3;
#pragma unroll
  for (int i = 0; i < 1024 / NT; ++i) {
    const int r = r0 + (NT / 8) * i;
    const int off = mn32 ? (r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4))) : (r * 128 + ((c ^ (r & 7)) << 4));
    const float4 v = *reinterpret_cast<const float4*>(hi_img + off);
    if (colsum) { colsum[0] += v.x; colsum[1] += v.y; colsum[2] += v.z; colsum[3] += v.w; }
    float4 l;
    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    *reinterpret_cast<float4*>(lo_img + off) = l;
  }
}
__device__ __forceinline__ void read_raw_atom(const uint8_t* raw, int tid, float4 v[8]) {
  const int c = tid & 7;
  const int r0 = tid >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(raw + (r0 + 16 * i) * 128 + c * 16);
}

// mn32: SWIZZLE_128B_BASE32B image (32-byte chunks XOR row%4) for MN-major reads, else SWIZZLE_128B (16-byte
// chunks XOR row%8) for K-major reads.
__device__ __forceinline__ int swz_off(int r, int c, bool mn32) {
  return mn32 ? (r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4))) : (r * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void store_atom_regs(const float4 v[8], uint8_t* hi_atom, uint8_t* lo_atom, int tid, bool mn32,
                                                float colsum[4] /*optional column sums*/, int atom_rows = 128) {
  const int c = tid & 7;
  const int r0 = tid >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + 16 * i;
    if (r >= atom_rows) break;
    if (colsum) { colsum[0] += v[i].x; colsum[1] += v[i].y; colsum[2] += v[i].z; colsum[3] += v[i].w; }
    float4 h, l;
    split_tf32(v[i].x, h.x, l.x); split_tf32(v[i].y, h.y, l.y); split_tf32(v[i].z, h.z, l.z); split_tf32(v[i].w, h.w, l.w);
    const int off = swz_off(r, c, mn32);
    *reinterpret_cast<float4*>(hi_atom + off) = h;
    *reinterpret_cast<float4*>(lo_atom + off) = l;
  }
}
__device__ __forceinline__ void stage_atom(const float* __restrict__ g, long long ld, long long row0, long long rows,
                                           int col0, int cols, uint8_t* hi_atom, uint8_t* lo_atom, int tid,
                                           float colsum[4], int atom_rows = 128, bool mn32 = false) {
  float4 v[8];
  load_atom_regs(g, ld, row0, rows, col0, cols, tid, v, atom_rows);
  store_atom_regs(v, hi_atom, lo_atom, tid, mn32, colsum, atom_rows);
}


__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA bulk copy global -> shared (no LSU work: one instruction per block), completion counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 2-D TMA tile load (tensor map in kernel parameter space): box lands in shared memory in the map's swizzle, rows/cols
// outside the tensor are zero-filled; completion counted in bytes on the mbarrier.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// A operand from TMEM (lane == row, 32-bit column == k): D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc),
      "r"((uint32_t)accumulate)
      : "memory");
}
// registers -> TMEM: thread i of warp w writes 32 consecutive columns of lane 32*(w%4)+i
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float v[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace
}  // namespace tnf
