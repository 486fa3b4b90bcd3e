// a14 (Cobafa trunk, src/models.py:7-28 MLP(36,128,5)) -- dense layers with 128 outputs and up to 128 inputs (36 -> 128 and
// 128 -> 128 in the trunk), forward and data gradient, with the WEIGHTS STATIONARY IN TENSOR MEMORY.
//
// linear_kernel (mlp.cu) keeps the hi/lo tf32 images of W resident in shared memory.  For a 128 x 128 layer that is 128 KB of
// the SM's 227 KB: three 16 KB slots are left for the operand rings, i.e. ONE atom of global loads in flight per SM, and the
// layer runs at 0.36-0.49 of its HBM bound (107-112 us for 268 MB at 2^18 rows).  Here the GEMM is transposed,
//     fwd  : Y^T [f, s] = sum_k W[f, k]   X [s, k]        dgrad: dX^T[k, s] = sum_f W[f, k] dY[s, f]
// so that the 128 x 128 weight matrix is the A operand -- which tcgen05.mma can take from TENSOR MEMORY (lane = output row,
// column = contraction index): hi image in columns 0..127, lo image in 128..255, written once per CTA.  Shared memory then
// holds nothing but the sample tiles: eight 16 KB TMA landing slots (128 samples x 32 contraction indices, 128-byte swizzle =
// the K-major B operand image, the raw fp32 tile being the tf32 hi operand) plus three lo images, i.e. up to 128 KB of loads in
// flight per SM, issued by one thread.  Accumulators: two 128-column buffers in the remaining tensor memory (512 columns in
// all), so the epilogue of tile t overlaps the MMAs of tile t+1.
// The accumulator comes out transposed (lane = feature, column = sample): a warp-wide store of one register writes 32
// consecutive features of ONE sample row = one full 128-byte line, so the epilogue needs no shared-memory transpose; bias and
// ReLU (fwd) are applied in registers.  The ReLU mask of the data gradient (the saved activation of the producing layer, a
// [M,128] fp32 matrix like the operand) is fetched by the same TMA thread, one 32-feature atom per TMEM lane quarter, and read
// from shared memory by the epilogue (64 per-thread global loads per tile measured 2.6x slower than the whole rest of the
// kernel: 158 us against 60 us for the forward).
// 3xTF32 (W_hi X_hi + W_hi X_lo + W_lo X_hi, fp32 accumulation) as everywhere in mlp.cu: fp32-grade results.
#include <cuda.h>
#include "common.cuh"
#include "tc.cuh"

namespace tnf {
namespace {

constexpr int kSLoWarps = 8;                        // lo-pass warps
constexpr int kSLoThreads = kSLoWarps * 32;
constexpr int kSEpiWarp0 = kSLoWarps;               // 8 epilogue warps (warp % 4 == TMEM lane quarter)
constexpr int kSMmaWarp = kSLoWarps + 8;
constexpr int kSTmaWarp = kSLoWarps + 9;
constexpr int kSThreads = (kSLoWarps + 10) * 32;
constexpr int kSH = 8, kSL = 3;                     // hi (TMA) slots, lo slots (forward)
constexpr int kSHm = 6;                             // hi slots of the data gradient: four more 16 KB slots hold the tile's ReLU-mask atoms
constexpr int kSDim = 128;                          // out_features (= TMEM lanes of the forward's A operand); in_features <= 128

struct WStatArgs {
  const float* W; int K;         // [128,K] row-major (nn.Linear.weight: [out, in]), K = in_features <= 128
  const float* bias;             // fwd: [128] or null
  const float* mask; long long ldmask;   // dgrad: activation whose > 0 gates dX (or null)
  float* Y; long long ldy;       // fwd: Y [M,128]; dgrad: dX [M,K]
  long long M; int n_tiles;
  int relu;
};

template <int MODE>  // 0 fwd, 1 dgrad
__global__ void __launch_bounds__(kSThreads, 1) wstat_linear_kernel(const WStatArgs A, const __grid_constant__ CUtensorMap tm_x,
                                                                     const __grid_constant__ CUtensorMap tm_m) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t s_hfull[kSH], s_hempty[kSH], s_afull[kSL], s_lempty[kSL], s_tfull[2], s_tempty[2], s_mfull[4], s_mempty[4];
  __shared__ uint32_t s_tmem;
  constexpr int H = (MODE == 0) ? kSH : kSHm;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* hi_ring = smem;
  uint8_t* lo_ring = smem + H * kAtomBytes;
  uint8_t* m_ring = lo_ring + kSL * kAtomBytes;       // dgrad: mask atom q (features 32 q .. 32 q + 31) of the tile in flight
  const bool masked = (MODE == 1) && A.mask != nullptr;
  const int T = (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
  const int K = A.K;
  const int atoms = (MODE == 0) ? (K + 31) >> 5 : kSDim / 32;   // contraction atoms per tile: in_features (fwd) / the 128 outputs (dgrad)
  const int lanes = (MODE == 0) ? kSDim : K;                    // valid rows of the A operand = features of the result
  const int n_items = T * atoms;

  if (tid == 0) {
    for (int i = 0; i < H; ++i) { mbar_init(&s_hfull[i], 1); mbar_init(&s_hempty[i], 1); }
    for (int q = 0; q < 4; ++q) { mbar_init(&s_mfull[q], 1); mbar_init(&s_mempty[q], 64); }
    for (int i = 0; i < kSL; ++i) { mbar_init(&s_afull[i], kSLoThreads); mbar_init(&s_lempty[i], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_tfull[b], 1); mbar_init(&s_tempty[b], 256); }
    fence_mbar_init();
  }
  if (warp == kSMmaWarp) tmem_alloc(&s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s_tmem;
  const uint32_t tm_whi = tm, tm_wlo = tm + 128, tm_d = tm + 256;   // W hi / lo images, accumulators at 256 / 384

  // ---- the stationary operand: epilogue thread (quarter q4, lane) owns TMEM lane a = 32 q4 + lane = row a of the A operand;
  // the two warps of a quarter write two of the four 32-column chunks each ----
  if (warp >= kSEpiWarp0 && warp < kSMmaWarp) {
    const int ew = warp - kSEpiWarp0, q4 = ew & 3, g = ew >> 2;
    const int a = q4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
#pragma unroll 1
    for (int ch = 2 * g; ch < 2 * g + 2; ++ch) {
      if (ch >= atoms) break;                 // columns the MMAs never read
      const int c0 = 32 * ch;
      float v[32];
      if (MODE == 0) {   // A[a][k] = W[a][k] (zero for k >= K): the thread's own row
        if ((K & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + i < K) t = __ldg(reinterpret_cast<const float4*>(A.W + (long long)a * K + c0 + i));
            v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (c0 + i < K) ? __ldg(A.W + (long long)a * K + c0 + i) : 0.f;
        }
      } else {           // A[a][f] = W[f][a] (zero rows for a >= K): column a, coalesced across the warp
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (a < K) ? __ldg(A.W + (long long)(c0 + i) * K + a) : 0.f;
      }
      tmem_st32(tm_whi + lane_off + c0, v);   // the tensor core truncates: the fp32 value is the hi operand
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = v[i] - __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
      tmem_st32(tm_wlo + lane_off + c0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == kSTmaWarp) {
    // ===== TMA producer: the sample tiles, `atoms` atoms (32 contraction indices each; columns beyond the matrix arrive as zeros) =====
    if (lane == 0) {
      tma_prefetch_desc(&tm_x);
      if (masked) tma_prefetch_desc(&tm_m);
      // mask atoms of tile tt: slot q is free once the 64 epilogue threads of lane quarter q have read tile tt-1's
      auto issue_mask = [&](int tt) {
        const int row0 = (blockIdx.x + tt * gridDim.x) * 128;
        for (int q = 0; 32 * q < K; ++q) {   // quarters without a valid feature neither wait for nor release a mask atom
          mbar_wait(&s_mempty[q], (tt & 1) ^ 1);
          mbar_expect_tx(&s_mfull[q], kAtomBytes);
          tma_load_2d(m_ring + q * kAtomBytes, &tm_m, 32 * q, row0, &s_mfull[q]);
        }
      };
      for (int tl = 0; tl < T; ++tl) {
        const int row0 = (blockIdx.x + tl * gridDim.x) * 128;
        for (int j = 0; j < atoms; ++j) {
          const int it = tl * atoms + j, h = it % H;
          mbar_wait(&s_hempty[h], ((it / H) & 1) ^ 1);
          mbar_expect_tx(&s_hfull[h], kAtomBytes);
          tma_load_2d(hi_ring + h * kAtomBytes, &tm_x, 32 * j, row0, &s_hfull[h]);
        }
        // the mask follows one tile behind the operand: its slots are released by the epilogue of the tile before it,
        // and the operand ring must not wait for that
        if (masked && tl >= 1) issue_mask(tl - 1);
      }
      if (masked) issue_mask(T - 1);
    }
  } else if (warp < kSLoWarps) {
    // ===== lo-pass warps: lo = x - trunc(x) of every landed atom =====
    for (int it = 0; it < n_items; ++it) {
      const int h = it % H, l = it % kSL;
      mbar_wait(&s_hfull[h], (it / H) & 1);
      mbar_wait(&s_lempty[l], ((it / kSL) & 1) ^ 1);
      make_lo_atom<kSLoThreads>(hi_ring + h * kAtomBytes, lo_ring + l * kAtomBytes, tid, false, nullptr);
      fence_async_smem();
      mbar_arrive(&s_afull[l]);
    }
  } else if (warp == kSMmaWarp) {
    // ===== MMA issuer: per atom 3 x 4 MMAs (M = 128 features, N = 128 samples, K = 8), A from tensor memory =====
    const uint32_t idesc = instr_desc(128, 128, false, false);
    for (int it = 0; it < n_items; ++it) {
      const int h = it % H, l = it % kSL, j = it % atoms, tl = it / atoms, b = tl & 1;
      if (j == 0) mbar_wait(&s_tempty[b], ((tl >> 1) & 1) ^ 1);   // the epilogue has drained accumulator b
      mbar_wait(&s_afull[l], (it / kSL) & 1);                     // hi landed (the lo pass saw it) and lo written
      tc_fence_after();
      if (elect_one()) {
        const uint32_t x_hi = smem_u32(hi_ring + h * kAtomBytes), x_lo = smem_u32(lo_ring + l * kAtomBytes);
        const uint32_t d = tm_d + b * 128;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {        // W_hi X_hi, W_hi X_lo, W_lo X_hi
          const uint32_t wa = (pass == 2 ? tm_wlo : tm_whi) + j * 32;
          const uint32_t xb = (pass == 1) ? x_lo : x_hi;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) mma_tf32_ts(d, wa + kk * 8, desc_kmajor(xb, kk), idesc, (j | pass | kk) != 0);
          if (pass == 1) mma_commit(&s_lempty[l]);    // the lo slot is free once its only pass has read it
        }
        mma_commit(&s_hempty[h]);
        if (j == atoms - 1) mma_commit(&s_tfull[b]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue warps: quarter q4 = features 32 q4 .. +31 (TMEM lanes), group g = samples 64 g .. +63 of the tile =====
    const int ew = warp - kSEpiWarp0, q4 = ew & 3, g = ew >> 2;
    const int a = q4 * 32 + lane;                       // output feature of this thread
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const float bias = (MODE == 0 && A.bias) ? __ldg(A.bias + a) : 0.f;
    const bool quarter_live = 32 * q4 < lanes;          // warp-uniform: this lane quarter holds valid features
    const bool feat_ok = a < lanes;
    for (int tl = 0; tl < T; ++tl) {
      const int b = tl & 1;
      const long long row0 = (long long)(blockIdx.x + tl * gridDim.x) * 128 + 64 * g;   // first sample row of this thread's columns
      // ReLU mask of this thread's 64 outputs as two bit words, from the quarter's mask atom in shared memory (row = sample,
      // 128-byte swizzle: the 32 lanes of a warp read the 32 features of one row, conflict-free)
      unsigned mbits[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
      if (masked && quarter_live) {
        mbar_wait(&s_mfull[q4], tl & 1);
        // 32-bit shared addresses: thread constant (atom, sample half, lane's 16-byte chunk and word) XOR the row's swizzle term
        const uint32_t mbase = smem_u32(m_ring + q4 * kAtomBytes) + (uint32_t)(64 * g) * 128u + ((lane >> 2) << 4) + ((lane & 3) << 2);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          unsigned bits = 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float mv;   // row r = 64 g + 32 ch + i: 16-byte chunks XOR r % 8 (= i % 8)
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mv) : "r"((mbase ^ (uint32_t)((i & 7) << 4)) + (uint32_t)((32 * ch + i) * 128)));
            bits |= (mv > 0.f ? 1u : 0u) << i;
          }
          mbits[ch] = bits;
        }
        mbar_arrive(&s_mempty[q4]);
      }
      mbar_wait(&s_tfull[b], (tl >> 1) & 1);
      tc_fence_after();
      if (!quarter_live) {   // nothing to read or store (dgrad of a narrow input): just hand the accumulator back
        tc_fence_before();
        mbar_arrive(&s_tempty[b]);
        continue;
      }
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float v[32];
        tmem_ld32(tm_d + b * 128 + lane_off + 64 * g + 32 * ch, v);
        if (ch == 1) { tc_fence_before(); mbar_arrive(&s_tempty[b]); }   // last read of accumulator b
        float* yp = A.Y + (row0 + 32 * ch) * A.ldy + a;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float y = v[i];
          if (MODE == 0) {
            y += bias;
            if (A.relu) y = fmaxf(y, 0.f);
          } else {
            y = ((mbits[ch] >> i) & 1u) ? y : 0.f;
          }
          if (feat_ok && row0 + 32 * ch + i < A.M) yp[(long long)i * A.ldy] = y;   // lanes = consecutive features of one row
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSMmaWarp) tmem_dealloc(tm, 512);
}

typedef CUresult (*EncodeTiledFnS)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// row-major fp32 [rows, cols] with leading dimension ld -> 32 x 128 boxes in the 128-byte swizzle (outside the matrix reads 0)
int make_atom_map_s(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  static EncodeTiledFnS encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return reinterpret_cast<EncodeTiledFnS>(fn);
  }();
  TNF_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TNF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TNF_OK;
}

}  // namespace

// fwd: x [m,k], y [m,128]; dgrad: x = dy [m,128], y = dx [m,k], mask [m,k]
bool wstat_linear_supported(int mode, int64_t m, int n, int k, const void* x, int64_t ldx, const void* w, const void* y, int64_t ldy,
                            const void* mask, int64_t ldmask) {
  const int x_cols = mode == 0 ? k : kSDim, y_cols = mode == 0 ? kSDim : k;
  const bool mask_ok = !mask || (ldmask % 4 == 0 && ldmask >= k && (reinterpret_cast<uintptr_t>(mask) & 15u) == 0);
  return n == kSDim && k >= 1 && k <= kSDim && m > 0 && m < (1LL << 31) - 256 && x && w && y && ldx % 4 == 0 && ldx >= x_cols &&
         ldy >= y_cols && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15u) == 0 && mask_ok;
}

int launch_wstat_linear(int mode, const float* x, int64_t ldx, const float* w, int k, const float* bias, int relu, const float* mask,
                        int64_t ldmask, float* y, int64_t ldy, int64_t m, cudaStream_t st) {
  WStatArgs A{};
  A.W = w; A.K = k; A.bias = bias; A.mask = mask; A.ldmask = ldmask; A.Y = y; A.ldy = ldy; A.M = m; A.relu = relu;
  A.n_tiles = (int)ceil_div(m, 128);
  CUtensorMap tm_x, tm_m;
  int rc = make_atom_map_s(&tm_x, x, m, mode == 0 ? k : kSDim, ldx);
  if (rc != TNF_OK) return rc;
  if (mode == 1 && mask) {
    rc = make_atom_map_s(&tm_m, mask, m, k, ldmask);
    if (rc != TNF_OK) return rc;
  } else {
    tm_m = tm_x;
  }
  const size_t smem_fwd = (size_t)(kSH + kSL) * kAtomBytes + 1024;
  const size_t smem_bwd = (size_t)(kSHm + kSL + 4) * kAtomBytes + 1024;
  static PerDeviceOnce configured{};
  if (configured.pending()) {
    TNF_CUDA(cudaFuncSetAttribute(wstat_linear_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fwd));
    TNF_CUDA(cudaFuncSetAttribute(wstat_linear_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bwd));
    configured.mark();
  }
  const int grid = A.n_tiles < sm_count() ? A.n_tiles : sm_count();
  if (mode == 0) wstat_linear_kernel<0><<<grid, kSThreads, smem_fwd, st>>>(A, tm_x, tm_m);
  else wstat_linear_kernel<1><<<grid, kSThreads, smem_bwd, st>>>(A, tm_x, tm_m);
  TNF_LAUNCH_CHECK("wstat_linear_kernel");
  return TNF_OK;
}

}  // namespace tnf
