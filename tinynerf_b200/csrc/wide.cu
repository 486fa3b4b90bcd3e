// a15-a17 -- dense layers that do not fit the resident-weight kernels of mlp.cu (in_features > 160 or out_features > 128):
// the reference's VanillaFeatureMLP(10, 256, 8) (src/models.py:59-68, run.py:131), the decoders' first layers behind it
// (256 -> 64, 307 -> 64; src/models.py:70-89) and the Cobafa colour head's first layer (179 -> 64, run.py:141-147).
// Round 1 sent these shapes to cuBLAS silently; here they run on tcgen05 as well, with BOTH operands streamed.
//
// One kernel, three roles, all 3xTF32 (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM, see mlp.cu):
//   fwd    Y[m,n]  = act(sum_k X[m,k] W[n,k] + b[n])   A = X  (K-major)            B = W  (K-major)      D: 128 rows x n
//   dgrad  dX[m,k] = sum_n dY[m,n] W[n,k] (* mask)     A = dY (K-major)            B = W  (MN-major)     D: 128 rows x k
//   wgrad  dW[n,k] = sum_m dY[m,n] X[m,k]              A = dY (MN-major, 128 n)    B = X  (MN-major)     D: 128 n    x k
// The contraction is walked in steps of 32: per step a stage holds the A block(s) and the B blocks (hi + lo images, 96 KB
// at 256 output columns), two stages alternate so the loads of step i+1 overlap the MMAs of step i.  128 threads: all stage
// (LDG -> split -> STS, swizzled), one elected lane issues the MMAs, thread == TMEM lane in the epilogue.
// Deliberately simple (no warp specialisation, no TMA): these layers belong to the reference's smallest configuration.
#include "common.cuh"
#include "tc.cuh"

namespace tnf {
namespace {

struct WideArgs {
  const float* A; long long lda;     // fwd: X [M,K]   dgrad: dY [M,N]   wgrad: dY [M,N] (+ column offset applied by the host)
  const float* B; long long ldb;     // fwd: W [N,K]   dgrad: W [N,K] (+ column offset)   wgrad: X [M,K] (+ column offset)
  const float* bias;                 // fwd
  float* D; long long ldd;           // fwd: Y   dgrad: dX   wgrad: dW (atomic accumulation)
  const float* mask; long long ldm;  // dgrad: activation whose > 0 gates dX (or null)
  float* dbias;                      // wgrad (or null)
  long long M;                       // rows (samples)
  int nd;                            // output columns of this launch (<= 256)
  int kd;                            // fwd: K   dgrad: N (contraction)   wgrad: unused (contraction is M)
  int ma;                            // wgrad: valid n (rows of D) in this launch (<= 128)
  int act;                           // fwd: 0 none, 1 relu, 2 exp(v-1), 3 sigmoid
  int n_tiles;                       // fwd/dgrad: 128-row tiles; wgrad: 32-row steps
};

constexpr int kBlk = 32 * 128;                 // one [32 rows][32 cols] fp32 block, bytes
constexpr int kStageA = 4 * kBlk;              // 16 KB: fwd/dgrad a 128x32 K-major atom, wgrad 4 MN-major blocks
constexpr int kStageB = 8 * kBlk;              // 32 KB: up to 256 output columns
constexpr int kStage = 2 * (kStageA + kStageB);  // hi + lo

// stage rows [r0, r0+32) x cols [c0, c0+32) of a row-major matrix as one 32x32 block (K-major 16-byte swizzle over 8-row
// groups, or the MN-major 32-byte swizzle); 128 threads, 2 chunks each
__device__ __forceinline__ void stage_block(const float* __restrict__ g, long long ld, long long r0, long long rows, int c0,
                                            int cols, uint8_t* hi, uint8_t* lo, int tid, bool mn32, float* colsum) {
  stage_atom(g, ld, r0, rows, c0, cols, hi, lo, tid, colsum, 32, mn32);
}

template <int MODE>  // 0 fwd, 1 dgrad, 2 wgrad
__global__ void __launch_bounds__(128, 1) wide_kernel(const WideArgs P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t s_free[2], s_done;
  __shared__ uint32_t s_tmem;
  __shared__ float s_db[128];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nd_pad = (P.nd + 15) & ~15;
  const int tcols = nd_pad <= 32 ? 32 : (nd_pad <= 64 ? 64 : (nd_pad <= 128 ? 128 : 256));
  if (tid == 0) {
    mbar_init(&s_free[0], 1); mbar_init(&s_free[1], 1); mbar_init(&s_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, tcols);
  if (MODE == 2) s_db[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = instr_desc(128, nd_pad, MODE == 2, MODE != 0);
  const int nblk_b = (nd_pad + 31) >> 5;
  uint32_t free_phase[2] = {0, 0}, done_phase = 0;
  int issued[2] = {0, 0};   // MMAs committed on each stage so far (a stage is reusable once its last commit has landed)
  float colsum[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) colsum[j][0] = colsum[j][1] = colsum[j][2] = colsum[j][3] = 0.f;

  auto stage_ptr = [&](int s, uint8_t*& a_hi, uint8_t*& a_lo, uint8_t*& b_hi, uint8_t*& b_lo) {
    uint8_t* base = smem + (size_t)s * kStage;
    a_hi = base; a_lo = base + kStageA; b_hi = base + 2 * kStageA; b_lo = base + 2 * kStageA + kStageB;
  };
  // one contraction step: stage, then 4 k-steps x 3 MMAs
  auto mma_step = [&](int s, bool first) {
    uint8_t *a_hi, *a_lo, *b_hi, *b_lo;
    stage_ptr(s, a_hi, a_lo, b_hi, b_lo);
    fence_async_smem();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t dah, dal, dbh, dbl;
        if (MODE == 2) { dah = desc_mnmajor(ah, kk, kBlk); dal = desc_mnmajor(al, kk, kBlk); }
        else           { dah = desc_kmajor(ah, kk);        dal = desc_kmajor(al, kk); }
        if (MODE == 0) { dbh = desc_kmajor(bh, kk);        dbl = desc_kmajor(bl, kk); }
        else           { dbh = desc_mnmajor(bh, kk, kBlk); dbl = desc_mnmajor(bl, kk, kBlk); }
        mma_tf32(tmem, dah, dbh, idesc, !(first && kk == 0));
        mma_tf32(tmem, dal, dbh, idesc, true);
        mma_tf32(tmem, dah, dbl, idesc, true);
      }
      mma_commit(&s_free[s]);
    }
    issued[s]++;
  };
  auto wait_stage = [&](int s) {   // before overwriting stage s: its previous MMAs must have read it
    if (issued[s] > 0) { mbar_wait(&s_free[s], free_phase[s]); free_phase[s] ^= 1; }
  };

  if (MODE != 2) {
    const int ksteps = (P.kd + 31) >> 5;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      const long long row0 = (long long)tile * 128;
      for (int ks = 0; ks < ksteps; ++ks, ++it) {
        const int s = it & 1;
        wait_stage(s);
        uint8_t *a_hi, *a_lo, *b_hi, *b_lo;
        stage_ptr(s, a_hi, a_lo, b_hi, b_lo);
        // A: 128 rows x 32 contraction columns, K-major atom
        stage_atom(P.A, P.lda, row0, P.M, 32 * ks, P.kd, a_hi, a_lo, tid, nullptr, 128, false);
        if (MODE == 0) {
          // B = W rows (output n) x 32 contraction columns, K-major, up to 256 rows = two 128-row atoms
          stage_atom(P.B, P.ldb, 0, P.nd, 32 * ks, P.kd, b_hi, b_lo, tid, nullptr, 128, false);
          if (nd_pad > 128) stage_atom(P.B, P.ldb, 128, P.nd, 32 * ks, P.kd, b_hi + kAtomBytes, b_lo + kAtomBytes, tid, nullptr, 128, false);
        } else {
          // B = W rows (contraction n, 32 of them) x output columns k, MN-major blocks of 32 columns
          for (int j = 0; j < nblk_b; ++j)
            stage_block(P.B, P.ldb, 32 * ks, P.kd, 32 * j, P.nd, b_hi + j * kBlk, b_lo + j * kBlk, tid, true, nullptr);
        }
        mma_step(s, ks == 0);
      }
      // the tile's accumulator is complete when everything issued so far has retired
      if (warp == 0 && elect_one()) mma_commit(&s_done);
      mbar_wait(&s_done, done_phase); done_phase ^= 1;
      tc_fence_after();
      const long long row = row0 + tid;
      for (int c0 = 0; c0 < nd_pad; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        if (row < P.M) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = c0 + i;
            if (c >= P.nd) break;
            float x = v[i];
            if (MODE == 0) {
              if (P.bias) x += __ldg(P.bias + c);
              if (P.act == 1) x = fmaxf(x, 0.f);
              else if (P.act == 2) x = expf(x - 1.f);
              else if (P.act == 3) x = 1.f / (1.f + expf(-x));
            } else if (P.mask) {
              x = __ldg(P.mask + row * P.ldm + c) > 0.f ? x : 0.f;
            }
            P.D[row * P.ldd + c] = x;
          }
        }
      }
      tc_fence_before();
      __syncthreads();   // every lane has read the accumulator before the next tile's first MMA overwrites it
    }
  } else {
    // wgrad: contraction over the rows m in steps of 32; this CTA accumulates its share of the steps in TMEM
    int it = 0;
    for (int step = blockIdx.x; step < P.n_tiles; step += gridDim.x, ++it) {
      const int s = it & 1;
      wait_stage(s);
      uint8_t *a_hi, *a_lo, *b_hi, *b_lo;
      stage_ptr(s, a_hi, a_lo, b_hi, b_lo);
      const long long r0 = (long long)step * 32;
      for (int j = 0; j < 4; ++j)   // A = dY rows (contraction m) x 128 columns n, four MN-major blocks
        stage_block(P.A, P.lda, r0, P.M, 32 * j, P.ma, a_hi + j * kBlk, a_lo + j * kBlk, tid, true, P.dbias ? colsum[j] : nullptr);
      for (int j = 0; j < nblk_b; ++j)   // B = X rows (contraction m) x columns k
        stage_block(P.B, P.ldb, r0, P.M, 32 * j, P.nd, b_hi + j * kBlk, b_lo + j * kBlk, tid, true, nullptr);
      mma_step(s, it == 0);
    }
    if (it > 0) {
      if (warp == 0 && elect_one()) mma_commit(&s_done);
      mbar_wait(&s_done, done_phase); done_phase ^= 1;
      tc_fence_after();
      const int n_out = tid;   // TMEM lane = row of dW
      for (int c0 = 0; c0 < nd_pad; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        if (n_out < P.ma) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < P.nd) atomicAdd(P.D + (long long)n_out * P.ldd + c0 + i, v[i]);
        }
      }
      if (P.dbias) {
        const int c = tid & 7;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) atomicAdd(&s_db[32 * j + 4 * c + e], colsum[j][e]);
        __syncthreads();
        if (tid < P.ma) atomicAdd(P.dbias + tid, s_db[tid]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tcols);
}

template <int MODE>
int launch_wide(const WideArgs& P, cudaStream_t st) {
  static PerDeviceOnce configured{};
  const size_t smem = 2 * (size_t)kStage + 1024;
  if (configured.pending()) {
    TNF_CUDA(cudaFuncSetAttribute(wide_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.mark();
  }
  const int grid = P.n_tiles < sm_count() ? P.n_tiles : sm_count();
  wide_kernel<MODE><<<grid, 128, smem, st>>>(P);
  TNF_LAUNCH_CHECK("wide_kernel");
  return TNF_OK;
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_wide_linear_fwd(const float* x, int64_t ldx, const float* weight, int64_t ldw, const float* bias, float* y,
                                   int64_t ldy, int64_t m, int32_t n, int32_t k, int32_t act, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(m >= 0 && n >= 1 && k >= 1 && act >= 0 && act <= 3, "bad sizes");
  if (m == 0) return TNF_OK;
  TNF_REQUIRE(x && weight && y, "null pointer");
  TNF_REQUIRE(ldx >= k && ldw >= k && ldy >= n, "leading dimension too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int n0 = 0; n0 < n; n0 += 256) {   // output columns in chunks of 256 (one accumulator)
    WideArgs P{};
    P.A = x; P.lda = ldx; P.B = weight + (int64_t)n0 * ldw; P.ldb = ldw; P.bias = bias ? bias + n0 : nullptr;
    P.D = y + n0; P.ldd = ldy; P.M = m; P.nd = n - n0 < 256 ? n - n0 : 256; P.kd = k; P.act = act;
    P.n_tiles = (int)ceil_div(m, 128);
    int rc = launch_wide<0>(P, st);
    if (rc != TNF_OK) return rc;
  }
  return TNF_OK;
}

extern "C" int tnf_wide_linear_bwd_data(const float* dy, int64_t lddy, const float* weight, int64_t ldw, float* dx, int64_t lddx,
                                        const float* relu_src, int64_t ldrs, int64_t m, int32_t n, int32_t k, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(m >= 0 && n >= 1 && k >= 1, "bad sizes");
  if (m == 0) return TNF_OK;
  TNF_REQUIRE(dy && weight && dx, "null pointer");
  TNF_REQUIRE(lddy >= n && ldw >= k && lddx >= k && (!relu_src || ldrs >= k), "leading dimension too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int k0 = 0; k0 < k; k0 += 256) {
    WideArgs P{};
    P.A = dy; P.lda = lddy; P.B = weight + k0; P.ldb = ldw; P.D = dx + k0; P.ldd = lddx;
    P.mask = relu_src ? relu_src + k0 : nullptr; P.ldm = ldrs;
    P.M = m; P.nd = k - k0 < 256 ? k - k0 : 256; P.kd = n; P.n_tiles = (int)ceil_div(m, 128);
    int rc = launch_wide<1>(P, st);
    if (rc != TNF_OK) return rc;
  }
  return TNF_OK;
}

extern "C" int tnf_wide_linear_bwd_weight(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dweight,
                                          int64_t lddw, float* dbias, int64_t m, int32_t n, int32_t k, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(m >= 0 && n >= 1 && k >= 1, "bad sizes");
  if (m == 0) return TNF_OK;
  TNF_REQUIRE(dy && x && dweight, "null pointer");
  TNF_REQUIRE(lddy >= n && ldx >= k && lddw >= k, "leading dimension too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int n0 = 0; n0 < n; n0 += 128)
    for (int k0 = 0; k0 < k; k0 += 256) {
      WideArgs P{};
      P.A = dy + n0; P.lda = lddy; P.B = x + k0; P.ldb = ldx; P.D = dweight + (int64_t)n0 * lddw + k0; P.ldd = lddw;
      P.dbias = (dbias && k0 == 0) ? dbias + n0 : nullptr;
      P.M = m; P.ma = n - n0 < 128 ? n - n0 : 128; P.nd = k - k0 < 256 ? k - k0 : 256; P.n_tiles = (int)ceil_div(m, 32);
      int rc = launch_wide<2>(P, st);
      if (rc != TNF_OK) return rc;
    }
  return TNF_OK;
}
