// a15-a17 -- the MLP heads' dense layers on the 5th-generation tensor cores (tcgen05 + TMEM).
// Reference: MLP / VanillaOpacityDecoder / VanillaColorDecoder / Cobafa trunk, src/models.py:7-28,70-89,
// executed there as fp32 cuBLAS SGEMMs (allow_tf32 is off) plus separate bias/ReLU kernels.
//
// fp32-grade accuracy on TF32 tensor cores: every operand x is split into hi = tf32(x) and
// lo = tf32(x - hi) while it is staged in shared memory and each product is issued as three
// tcgen05.mma.kind::tf32 (hi*hi + lo*hi + hi*lo) accumulating in fp32 in TMEM ("3xTF32", ~2^-21 relative
// per product; the dropped lo*lo term is below fp32 rounding of the sum).
//
// One smem image serves every GEMM of fwd and bwd: an "atom" = 128 rows x 32 fp32 (128 B per row) in the
// canonical 128-byte-swizzle layout (16-byte chunk index XOR row%8).  Read as a K-major operand its rows are
// the M/N index and its 32 columns are K; read as an MN-major operand its rows are K and its columns are M/N:
//   forward   Y[m,n]  = act(sum_k X[m,k] W[n,k] + b[n])  A = X atoms (K-major)      B = W atoms (K-major)
//   dgrad     dX[m,k] = sum_n dY[m,n] W[n,k]             A = dY atoms (K-major)     B = W atoms (MN-major)
//   wgrad     dW[n,k] = sum_m dY[m,n] X[m,k]             A = dY atoms (MN-major)    B = X atoms (MN-major)
// so activations/gradients [rows, features] and nn.Linear weights [out, in] are staged exactly as they lie
// in HBM, with coalesced 128-bit loads, no transposes.
//
// Structure (v1, deliberately simple): 128 threads per CTA, persistent over 128-row tiles, two CTAs per SM so
// one CTA's loads overlap the other's MMAs/epilogue; operands staged by all threads (LDG -> split -> STS),
// one elected thread issues the MMAs, completion through tcgen05.commit -> mbarrier, accumulator read back
// with tcgen05.ld (thread == TMEM lane == output row) for the fused bias / ReLU / head epilogue.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc.cuh"

namespace tnf {
namespace {
// role timing (diagnostics): cycles spent waiting, accumulated per CTA into dbg[role*8 + slot] when set
__device__ long long* g_dbg = nullptr;
#ifdef TNF_ROLE_TIMING
#define TNF_CLK() clock64()
#else
#define TNF_CLK() 0LL
#endif
struct Tm {
  long long t0;
  __device__ __forceinline__ void start() { t0 = TNF_CLK(); }
  __device__ __forceinline__ void stop(long long& acc) { acc += TNF_CLK() - t0; }
};

struct LinArgs {
  const float* X; long long ldx;   // fwd: input [M,K]        dgrad: dY [M,N]          wgrad: dY [M,N]
  const float* W;                  // [N,K] row-major (nn.Linear.weight)
  const float* bias;               // fwd: [N] or null
  float* Y; long long ldy;         // fwd: output [M,N]       dgrad: dX [M,K] (ldy)    wgrad: unused
  const float* X2; long long ldx2; // dgrad: activation whose >0 mask gates dX (or null)   wgrad: X [M,K]
  float* dW; float* db;            // wgrad outputs (accumulated with atomics; caller zeroes)
  const float* head_w; const float* head_b; float* head_out; int n_head; int head_act;  // fwd: fused small head
  long long M; int N, K;
  int relu;
  int n_tiles;
  float* scratch;  // wgrad_tma_kernel, optional: [64][32 kx] partial sums + one counter word, zero on entry and on exit
  int kxa, col_b;  // wgrad_tma_kernel: X atoms [0, kxa) come from the first tensor map and land at dW columns 32 j, atoms
                   // [kxa, kx) from the second map at dW columns col_b + 32 (j - kxa)  (X = [X_a | X_b], col_b = width of X_a)
  int ring;        // hi-image slots of the operand ring (16 KB each), chosen by the host from the smem budget
  int raw;         // lo-image slots (16 KB each); ring - raw atoms of global loads are kept in flight
};

__device__ __forceinline__ float head_activation(float x, int act) {
  if (act == 1) return expf(x - 1.f);                 // truncated_exp(x - 1.)  (src/models.py:74)
  if (act == 2) return 1.f / (1.f + expf(-x));        // sigmoid               (src/models.py:85)
  return x;
}

// ---- forward / dgrad: warp-specialised, persistent, one CTA per SM --------------------------------------------
//   warps 0-3  loaders : cp.async global -> swizzled hi slot (the raw fp32 tile is the hi operand), several atoms in
//                        flight; then lo = x - trunc(x) from the landed chunks -> lo slot -> full[l]
//   warps 4-11 epilogue: tmem_full[b] -> tcgen05.ld (warp w owns TMEM lanes 32*(w%4)..; two warps per quarter split
//                        the column chunks) -> bias/ReLU/head -> warp-private transpose -> coalesced stores -> tmem_empty[b]
//   warp  12   MMA     : full[s] -> 12 x tcgen05.mma.kind::tf32 (3xTF32) -> tcgen05.commit -> empty[s]
//                        (+ tmem_full[b] after the tile's last atom); accumulators double-buffered in TMEM
// so HBM loads, tensor-core work and the epilogue of consecutive tiles overlap inside one CTA.
constexpr int kWsThreads = 416;   // 4 loader + 8 epilogue + 1 MMA warps
constexpr int kMaxStages = 4;
constexpr int kMaxHi = 7;     // fwd/dgrad: hi slots (cp.async targets + MMA operands)
constexpr int kMaxLo = 3;     // fwd/dgrad: lo slots


// dynamic smem: [W hi atoms][W lo atoms][hi ring: H x 16 KB][lo ring: L x 16 KB][epilogue staging 16 KB], 1024-aligned
template <int MODE>  // 0 fwd, 1 dgrad
__global__ void __launch_bounds__(kWsThreads, 1) linear_kernel(const LinArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  long long k_entry = 0;
#ifdef TNF_ROLE_TIMING
  if (g_dbg && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_entry)); if (blockIdx.x == 0) g_dbg[24] = k_entry; }
#endif
  const long long c_entry = TNF_CLK();
  __shared__ uint64_t s_full[kMaxLo], s_lempty[kMaxLo], s_hempty[kMaxHi], s_tfull[2], s_tempty[2];
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // GEMM view: D[128, ND] += sum over KD of A[128, KD] * B[ND, KD]
  const int KD = (MODE == 0) ? A.K : A.N;                 // contraction length
  const int ND = (MODE == 0) ? A.N : ((A.K + 15) & ~15);  // output columns (dgrad: K_in padded to 16)
  const int ka = (KD + 31) >> 5;                          // contraction atoms per tile
  const int w_atoms = (A.K + 31) >> 5;                    // weight image: N rows x K cols -> atoms along K
  const int w_rows = (A.N + 15) & ~15;
  const int w_stride = ((w_rows * 128) + 1023) & ~1023;
  const int H = A.ring, L = A.raw;                        // hi / lo slots
  uint8_t* w_hi = smem;
  uint8_t* w_lo = smem + w_atoms * w_stride;
  uint8_t* hi_ring = smem + 2 * w_atoms * w_stride;
  uint8_t* lo_ring = hi_ring + H * kAtomBytes;
  uint8_t* epi = lo_ring + L * kAtomBytes;
  const int nd_cols = ND <= 32 ? 32 : (ND <= 64 ? 64 : (ND <= 128 ? 128 : 256));  // TMEM columns per accumulator
  const int tmem_cols = 2 * nd_cols;

  if (tid == 0) {
    for (int i = 0; i < L; ++i) { mbar_init(&s_full[i], 128); mbar_init(&s_lempty[i], 1); }
    for (int i = 0; i < H; ++i) mbar_init(&s_hempty[i], 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&s_tfull[b], 1); mbar_init(&s_tempty[b], 256); }
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc(&s_tmem, tmem_cols);
  if (tid < 128)
    for (int j = 0; j < w_atoms; ++j)
      stage_atom(A.W, A.K, 0, A.N, 32 * j, A.K, w_hi + j * w_stride, w_lo + j * w_stride, tid, nullptr, w_rows, MODE == 1);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  const int my_tiles = (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_items = my_tiles * ka;

  if (warp < 4) {
    // ===== loaders =====
    const bool direct = ((A.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(A.X) & 15u) == 0);
    const int D = H - L;  // atoms of global loads in flight
    auto src_row0 = [&](int it) { return (long long)(blockIdx.x + (it / ka) * gridDim.x) * 128; };
    if (direct) {
      long long w_h = 0, w_cp = 0, w_l = 0, w_lo = 0, w_tot = 0;
      Tm tm, tt;
      tt.start();
      if (g_dbg && tid == 0 && blockIdx.x == 0) g_dbg[6] = tt.t0 - c_entry;  // cycles from kernel entry to role start
      auto issue = [&](int it) {
        if (it < n_items) {
          const int h = it % H;
          tm.start();
          mbar_wait(&s_hempty[h], ((it / H) & 1) ^ 1);   // the MMAs that read the slot's previous atom have completed
          tm.stop(w_h);
          cp_async_atom_swz(A.X, A.ldx, src_row0(it), A.M, 32 * (it % ka), KD, tid, hi_ring + h * kAtomBytes, false);
        }
        cp_async_commit();  // (possibly empty) group: keeps the group count uniform
      };
      for (int it = 0; it < D; ++it) issue(it);
      for (int it = 0; it < n_items; ++it) {
        issue(it + D);
        tm.start();
        cp_async_wait_dyn(D);                            // this thread's chunks of item `it` have landed
        tm.stop(w_cp);
        const int l = it % L;
        tm.start();
        mbar_wait(&s_lempty[l], ((it / L) & 1) ^ 1);
        tm.stop(w_l);
        tm.start();
        make_lo_atom(hi_ring + (it % H) * kAtomBytes, lo_ring + l * kAtomBytes, tid, false, nullptr);
        fence_async_smem();
        mbar_arrive(&s_full[l]);
        tm.stop(w_lo);
      }
      cp_async_wait<0>();
      tt.stop(w_tot);
      if (g_dbg && tid == 0 && blockIdx.x == 0) { g_dbg[0] = w_h; g_dbg[1] = w_cp; g_dbg[2] = w_l; g_dbg[3] = w_lo; g_dbg[4] = w_tot; g_dbg[5] = n_items; }
    } else {
      // unaligned rows: registers, one atom ahead
      float4 pre[8];
      if (n_items > 0) load_atom_regs(A.X, A.ldx, src_row0(0), A.M, 0, KD, tid, pre);
      for (int it = 0; it < n_items; ++it) {
        const int h = it % H, l = it % L;
        mbar_wait(&s_hempty[h], ((it / H) & 1) ^ 1);
        mbar_wait(&s_lempty[l], ((it / L) & 1) ^ 1);
        store_atom_regs(pre, hi_ring + h * kAtomBytes, lo_ring + l * kAtomBytes, tid, false, nullptr);
        if (it + 1 < n_items) load_atom_regs(A.X, A.ldx, src_row0(it + 1), A.M, 32 * ((it + 1) % ka), KD, tid, pre);
        fence_async_smem();
        mbar_arrive(&s_full[l]);
      }
    }
  } else if (warp == 12) {
    // ===== MMA issuer =====
    const uint32_t idesc = instr_desc(128, ND, false, MODE == 1);
    long long w_te = 0, w_f = 0, w_tot = 0;
    Tm tm, tt;
    tt.start();
    for (int it = 0; it < n_items; ++it) {
      const int h = it % H, l = it % L, j = it % ka, tl = it / ka, b = tl & 1;
      tm.start();
      if (j == 0) mbar_wait(&s_tempty[b], ((tl >> 1) & 1) ^ 1);   // epilogue has drained accumulator b
      tm.stop(w_te);
      tm.start();
      mbar_wait(&s_full[l], (it / L) & 1);
      tm.stop(w_f);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ah = smem_u32(hi_ring + h * kAtomBytes), al = smem_u32(lo_ring + l * kAtomBytes);
        const uint32_t tmem_d = tmem_base + b * nd_cols;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t aa = (pass == 1) ? al : ah;
          const uint8_t* wb = (pass == 2) ? w_lo : w_hi;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint64_t bd;
            if (MODE == 0) bd = desc_kmajor(smem_u32(wb + j * w_stride), kk);
            else           bd = desc_mnmajor(smem_u32(wb) + j * 4096 /*32 rows of n_out*/, kk, w_stride);
            mma_tf32(tmem_d, desc_kmajor(aa, kk), bd, idesc, (j | pass | kk) != 0);
          }
          if (pass == 1) mma_commit(&s_lempty[l]);   // the lo slot is free once the first two passes have read it
        }
        mma_commit(&s_hempty[h]);                  // hi slot reusable once all three passes have read it
        if (j == ka - 1) mma_commit(&s_tfull[b]);  // accumulator b complete
      }
      __syncwarp();
    }
    tt.stop(w_tot);
    if (g_dbg && lane == 0 && blockIdx.x == 0) { g_dbg[8] = w_te; g_dbg[9] = w_f; g_dbg[10] = w_tot; }
  } else {
    // ===== epilogue (warps 4..11).  Warp w may touch TMEM lanes 32*(w%4) .. +31: two warps share each lane quarter
    // and split the 32-column chunks between them (group g takes chunks g, g+2, ...).  A chunk goes TMEM -> registers
    // (thread == row) -> bias/ReLU/head -> warp-private 4 KB transpose buffer -> coalesced 128-byte row stores; only
    // __syncwarp inside, one named barrier per tile when the fused head needs the two groups' partial sums.
    const int ew = warp - 4, q4 = ew & 3, g = ew >> 2;
    const int r = q4 * 32 + lane;                 // tile row owned by this thread in TMEM
    uint8_t* wbuf = epi + ew * 4096;              // 32 rows x 128 B, 16-byte chunks XOR row%8
    float* s_head = reinterpret_cast<float*>(epi + 8 * 4096);  // [2][128][4] partial head sums of group 1
    const int lr = lane >> 3, c = lane & 7;       // coalesced mapping: rows lr + 4 i of the warp's 32, column chunk c
    const int ncols = (MODE == 0) ? A.N : A.K;
    const bool write_y = (MODE == 1) || (A.Y != nullptr);
    const int n_chunks = (ND + 31) >> 5;
    const bool split_head = (MODE == 0) && A.n_head > 0 && n_chunks > 1;
    long long w_tf = 0, w_tot = 0;
    Tm tm, tt;
    tt.start();
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int b = tl & 1;
      const long long row0 = (long long)(blockIdx.x + tl * gridDim.x) * 128;
      tm.start();
      mbar_wait(&s_tfull[b], (tl >> 1) & 1);
      tm.stop(w_tf);
      tc_fence_after();
      const uint32_t taddr = tmem_base + b * nd_cols + ((uint32_t)(q4 * 32) << 16);
      float head_acc[4] = {0.f, 0.f, 0.f, 0.f};
      bool released = false;
      for (int ch = g; ch < n_chunks; ch += 2) {
        const int c0 = ch * 32;
        const int col = c0 + 4 * c;
        float4 xin[8];
        if (MODE == 1 && A.X2) {  // ReLU-backward mask source: issue the loads now, they land during the TMEM read
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long grow = row0 + q4 * 32 + lr + 4 * i;
            xin[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (grow < A.M && col < ncols) {
              const float* xp = A.X2 + grow * A.ldx2 + col;
              if (col + 3 < ncols) {
                xin[i] = __ldg(reinterpret_cast<const float4*>(xp));
              } else {
                xin[i].x = __ldg(xp);
                if (col + 1 < ncols) xin[i].y = __ldg(xp + 1);
                if (col + 2 < ncols) xin[i].z = __ldg(xp + 2);
              }
            }
          }
        }
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (ch + 2 >= n_chunks) { tc_fence_before(); mbar_arrive(&s_tempty[b]); released = true; }  // last read of accumulator b
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float y = v[i] + (A.bias ? __ldg(A.bias + c0 + i) : 0.f);
            if (A.relu) y = fmaxf(y, 0.f);
            v[i] = y;
          }
          for (int o = 0; o < A.n_head; ++o) {
            float acc = head_acc[o];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc = __fmaf_rn(v[i], __ldg(A.head_w + o * A.N + c0 + i), acc);
            head_acc[o] = acc;
          }
        }
        if (write_y) {
#pragma unroll
          for (int qq = 0; qq < 8; ++qq)
            *reinterpret_cast<float4*>(wbuf + lane * 128 + ((qq ^ (lane & 7)) << 4)) =
                make_float4(v[4 * qq], v[4 * qq + 1], v[4 * qq + 2], v[4 * qq + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = lr + 4 * i;
            const long long grow = row0 + q4 * 32 + rr;
            if (grow < A.M && col < ncols) {
              float4 o = *reinterpret_cast<const float4*>(wbuf + rr * 128 + ((c ^ (rr & 7)) << 4));
              if (MODE == 1 && A.X2) {  // ReLU backward of the layer that produced this layer's input
                o.x = xin[i].x > 0.f ? o.x : 0.f; o.y = xin[i].y > 0.f ? o.y : 0.f;
                o.z = xin[i].z > 0.f ? o.z : 0.f; o.w = xin[i].w > 0.f ? o.w : 0.f;
              }
              float* yp = A.Y + grow * A.ldy + col;
              if (col + 3 < ncols) {
                *reinterpret_cast<float4*>(yp) = o;
              } else {
                yp[0] = o.x;
                if (col + 1 < ncols) yp[1] = o.y;
                if (col + 2 < ncols) yp[2] = o.z;
              }
            }
          }
          __syncwarp();
        }
      }
      if (!released) { tc_fence_before(); mbar_arrive(&s_tempty[b]); }  // a group without chunks still releases
      if (MODE == 0 && A.n_head > 0) {
        const long long row = row0 + r;
        if (split_head) {  // group 1 hands its partial sums to group 0 (double-buffered by tile parity)
          float* hp = s_head + (b * 128 + r) * 4;
          if (g == 1) {
#pragma unroll
            for (int o = 0; o < 4; ++o) hp[o] = head_acc[o];
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (g == 0) {
#pragma unroll
            for (int o = 0; o < 4; ++o) head_acc[o] += hp[o];
          }
        }
        if (g == 0 && row < A.M)
          for (int o = 0; o < A.n_head; ++o)
            A.head_out[row * A.n_head + o] = head_activation(head_acc[o] + __ldg(A.head_b + o), A.head_act);
      }
    }
    tt.stop(w_tot);
    if (g_dbg && tid == 128 && blockIdx.x == 0) { g_dbg[16] = w_tf; g_dbg[17] = w_tot; g_dbg[18] = my_tiles; }
  }
  tc_fence_before();
  __syncthreads();
#ifdef TNF_ROLE_TIMING
  if (g_dbg && threadIdx.x == 0) {
    long long k_exit;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_exit));
    atomicMax((unsigned long long*)&g_dbg[25], (unsigned long long)k_exit);
    atomicMin((unsigned long long*)&g_dbg[26], (unsigned long long)k_entry);
  }
#endif
  if (warp == 12) tmem_dealloc(tmem_base, tmem_cols);
}


// wgrad: dW[n,k] += sum_m dY[m,n] X[m,k] ; db[n] += sum_m dY[m,n].  Warp-specialised like linear_kernel:
//   warps 0-7 loaders: per 128-sample tile the dY atoms (into one of TWO resident sets, so the next tile's dY is staged
//                      while the tensor core still reads this tile's) and then the X atoms (ring), all copied by
//                      cp.async straight into the MN-major operand images (SWIZZLE_128B_BASE32B), a few atoms in
//                      flight; the loaders add the lo images and the column sums of dY (bias gradient)
//   warp  8   MMA    : per X atom j: D[N_out, 32j..32j+32) += dY^T X_j, 16 k-steps x 3 (3xTF32), both operands
//                      MN-major; D lives in TMEM for the whole kernel and is flushed with atomics once per CTA.
constexpr int kWgLoadWarps = 8;
constexpr int kWgLoadThreads = kWgLoadWarps * 32;
constexpr int kWgThreads = kWgLoadThreads + 32;
constexpr int kWgDist = 2;   // atoms of global loads in flight

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const LinArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t s_xfull[kMaxStages], s_xempty[kMaxStages], s_yfull[2], s_yempty[2], s_done;
  __shared__ uint32_t s_tmem;
  __shared__ float s_db[128];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ny = (A.N + 31) >> 5;   // dY atoms (N_out / 32)
  const int kx = (A.K + 31) >> 5;   // X atoms
  const int S = A.ring;             // X stages (hi + lo image each)
  const int n_sets = A.raw;         // resident dY sets (2 when they fit, else 1)
  const int set_bytes = 2 * ny * kAtomBytes;      // [hi atoms][lo atoms]
  uint8_t* ysets = smem;
  uint8_t* xst = smem + n_sets * set_bytes;       // X stage s: hi at +s*2*kAtomBytes, lo right after
  // N_out == 64: ONE MMA per k-step computes all four hi/lo products: A = [dY_hi ; dY_lo] (M = 128: the set's four
  // atoms are consecutive MN blocks), B = [X_hi | X_lo] (N = 64: the stage's two images are consecutive MN blocks).
  // D quadrants (lanes n / n+64, columns c / c+32 of the atom's 64) are summed when the accumulator is flushed.
  const bool stacked = (A.N == 64);
  const int MM = stacked ? 128 : ((A.N <= 64) ? 64 : 128);   // UMMA M
  const int ncol = kx * (stacked ? 64 : 32);
  const int tmem_cols = ncol <= 32 ? 32 : (ncol <= 64 ? 64 : (ncol <= 128 ? 128 : (ncol <= 256 ? 256 : 512)));
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&s_xfull[s], kWgLoadThreads); mbar_init(&s_xempty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_yfull[i], kWgLoadThreads * ny); mbar_init(&s_yempty[i], 1); }
    mbar_init(&s_done, 1);
    fence_mbar_init();
  }
  if (warp == kWgLoadWarps) tmem_alloc(&s_tmem, tmem_cols);
  if (tid < 128) s_db[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = s_tmem;
  const int per_tile = ny + kx;
  const int my_tiles = (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_items = my_tiles * per_tile;

  if (warp < kWgLoadWarps) {
    // ===== loaders =====
    float colsum[kMaxKAtoms - 1][4];
#pragma unroll
    for (int a = 0; a < kMaxKAtoms - 1; ++a) colsum[a][0] = colsum[a][1] = colsum[a][2] = colsum[a][3] = 0.f;
    const bool direct = ((A.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(A.X) & 15u) == 0) && ((A.ldx2 & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(A.X2) & 15u) == 0);
    // item `it` of this CTA: tile it / per_tile, atom it % per_tile (dY atoms first).  X atoms use ring stage xi % S where
    // xi counts X atoms; dY atoms go to set tile % n_sets.
    auto hi_image = [&](int it) -> uint8_t* {
      const int tl = it / per_tile, item = it % per_tile;
      if (item < ny) return ysets + (tl % n_sets) * set_bytes + item * kAtomBytes;
      return xst + ((tl * kx + item - ny) % S) * 2 * kAtomBytes;
    };
    auto wait_free = [&](int it) {   // the MMAs that read the slot's previous content have completed
      const int tl = it / per_tile, item = it % per_tile;
      if (item == 0) mbar_wait(&s_yempty[tl % n_sets], ((tl / n_sets) & 1) ^ 1);
      if (item >= ny) { const int xi = tl * kx + item - ny; mbar_wait(&s_xempty[xi % S], ((xi / S) & 1) ^ 1); }
    };
    auto src_of = [&](int it, const float*& g, long long& ld, int& col0, int& cols, long long& row0) {
      const int t = blockIdx.x + (it / per_tile) * gridDim.x, item = it % per_tile;
      row0 = (long long)t * 128;
      if (item < ny) { g = A.X; ld = A.ldx; col0 = 32 * item; cols = A.N; }
      else { g = A.X2; ld = A.ldx2; col0 = 32 * (item - ny); cols = A.K; }
    };
    auto issue = [&](int it) {
      if (it < n_items) {
        wait_free(it);
        const float* g; long long ld, row0; int col0, cols;
        src_of(it, g, ld, col0, cols, row0);
        cp_async_atom_swz<kWgLoadThreads>(g, ld, row0, A.M, col0, cols, tid, hi_image(it), true);
      }
      cp_async_commit();
    };
    // atoms in flight: bounded by the X ring (the slot of an atom being issued must not belong to an atom this loop has
    // not completed yet); with a single dY set the next tile's dY can only be issued once this tile is consumed -> 0
    const int dist = (n_sets == 2) ? (kWgDist < S - 1 ? kWgDist : S - 1) : 0;
    if (direct) {
      for (int it = 0; it < dist; ++it) issue(it);
    }
    for (int it = 0; it < n_items; ++it) {
      const int tl = it / per_tile, item = it % per_tile;
      uint8_t* hi = hi_image(it);
      uint8_t* lo = (item < ny) ? hi + ny * kAtomBytes : hi + kAtomBytes;
      if (direct) {
        issue(it + dist);
        cp_async_wait_dyn(dist);
        if (item < ny) {
#pragma unroll
          for (int a = 0; a < kMaxKAtoms - 1; ++a)
            if (a == item) make_lo_atom<kWgLoadThreads>(hi, lo, tid, true, colsum[a]);
        } else {
          make_lo_atom<kWgLoadThreads>(hi, lo, tid, true, nullptr);
        }
      } else if (tid < 128) {
        // unaligned rows: 128 threads stage through registers (hi and lo images written together)
        wait_free(it);
        const float* g; long long ld, row0; int col0, cols;
        src_of(it, g, ld, col0, cols, row0);
        float4 v[8];
        load_atom_regs(g, ld, row0, A.M, col0, cols, tid, v);
        if (item < ny) {
#pragma unroll
          for (int a = 0; a < kMaxKAtoms - 1; ++a)
            if (a == item) store_atom_regs(v, hi, lo, tid, true, colsum[a]);
        } else {
          store_atom_regs(v, hi, lo, tid, true, nullptr);
        }
      }
      fence_async_smem();
      if (item < ny) mbar_arrive(&s_yfull[tl % n_sets]);
      else mbar_arrive(&s_xfull[(tl * kx + item - ny) % S]);
    }
    cp_async_wait<0>();
    // bias gradient: column sums of dY gathered while staging
    if (A.db) {
#pragma unroll
      for (int a = 0; a < kMaxKAtoms - 1; ++a)
        if (a < ny)
#pragma unroll
          for (int q = 0; q < 4; ++q) atomicAdd(&s_db[32 * a + 4 * (tid & 7) + q], colsum[a][q]);
    }
  } else {
    // ===== MMA issuer =====
    const uint32_t idesc = instr_desc(MM, stacked ? 64 : 32, true, true);
    int xi = 0;
    long long w_y = 0, w_x = 0, w_iss = 0;
    const long long t_start = TNF_CLK();
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int set = tl % n_sets;
      { const long long t0 = TNF_CLK(); mbar_wait(&s_yfull[set], (tl / n_sets) & 1); w_y += TNF_CLK() - t0; }
      for (int j = 0; j < kx; ++j, ++xi) {
        const int s = xi % S;
        { const long long t0 = TNF_CLK(); mbar_wait(&s_xfull[s], (xi / S) & 1); w_x += TNF_CLK() - t0; }
        const long long ti = TNF_CLK();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t yh = smem_u32(ysets + set * set_bytes), yl = yh + ny * kAtomBytes;
          const uint32_t xh = smem_u32(xst + s * 2 * kAtomBytes), xl = xh + kAtomBytes;
          if (stacked) {
#pragma unroll 4
            for (int kk = 0; kk < 16; ++kk)
              mma_tf32(tmem_d + 64 * j, desc_mnmajor(yh, kk), desc_mnmajor(xh, kk), idesc, !(tl == 0 && kk == 0));
          } else
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t ya = (pass == 1) ? yl : yh;
            const uint32_t xa = (pass == 2) ? xl : xh;
#pragma unroll 1
            for (int kk = 0; kk < 16; ++kk)  // 128 samples = 16 k-steps of 8 rows
              mma_tf32(tmem_d + 32 * j, desc_mnmajor(ya, kk), desc_mnmajor(xa, kk), idesc, !(tl == 0 && pass == 0 && kk == 0));
          }
          mma_commit(&s_xempty[s]);
          if (j == kx - 1) mma_commit(&s_yempty[set]);
        }
        __syncwarp();
        w_iss += TNF_CLK() - ti;
      }
    }
    if (elect_one()) mma_commit(&s_done);
    __syncwarp();
    if (g_dbg && lane == 0 && blockIdx.x == 0) { g_dbg[8] = w_y; g_dbg[9] = w_x; g_dbg[10] = TNF_CLK() - t_start; g_dbg[11] = w_iss; }
  }
  { const long long t0 = TNF_CLK();
  mbar_wait(&s_done, 0);   // every MMA of this CTA has completed
  if (g_dbg && tid == 0 && blockIdx.x == 0) g_dbg[12] = TNF_CLK() - t0; }
  const long long t_flush = TNF_CLK();
  tc_fence_after();
  __syncthreads();
  if (A.db && tid < A.N) atomicAdd(A.db + tid, s_db[tid]);
  if (my_tiles > 0 && warp < 4 && stacked) {
    const int n_out = (warp * 32 + lane) & 63;      // lanes n and n + 64 hold the dY_hi / dY_lo rows of feature n
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int j = 0; j < kx; ++j) {
      float v[32], u[32];
      tmem_ld32(taddr + 64 * j, v);        // x X_hi
      tmem_ld32(taddr + 64 * j + 32, u);   // x X_lo
      float* dst = A.dW + (long long)n_out * A.K + 32 * j;
      if ((A.K & 3) == 0 && (reinterpret_cast<uintptr_t>(A.dW) & 15u) == 0 && 32 * j + 31 < A.K) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)   // one L2 reduction per 16 bytes
          red_add_f4(dst + i, make_float4(v[i] + u[i], v[i + 1] + u[i + 1], v[i + 2] + u[i + 2], v[i + 3] + u[i + 3]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (32 * j + i < A.K) atomicAdd(dst + i, v[i] + u[i]);
      }
    }
  } else if (my_tiles > 0 && warp < 4) {
    // D rows = output features: M=128 -> lane == row; M=64 -> row m sits in lane 32*(m/16) + m%16
    int n_out;
    bool active;
    if (MM == 128) { n_out = warp * 32 + lane; active = true; }
    else { n_out = warp * 16 + lane; active = lane < 16; }
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < ncol; c0 += 32) {
      float v[32];
      tmem_ld32(taddr + c0, v);
      if (active && n_out < A.N) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < A.K) atomicAdd(A.dW + (long long)n_out * A.K + c0 + i, v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (g_dbg && tid == 0 && blockIdx.x == 0) g_dbg[13] = TNF_CLK() - t_flush;
  if (warp == kWgLoadWarps) tmem_dealloc(tmem_d, tmem_cols);
}

// wgrad for 64-wide layers: the A operand [dY_hi ; dY_lo]^T (128 x 128 samples) lives in TENSOR MEMORY.
// The stacked kernel above keeps two dY sets (hi + lo images, 128 KB) in shared memory, which leaves room for two atoms of
// global loads in flight: it is bound by load latency, not by HBM or the tensor pipe.  Here warps transpose the raw dY tile
// shared memory -> registers -> tcgen05.st (thread == TMEM lane == one output feature, hi or lo half; the column sums for
// the bias gradient fall out of the same registers), so shared memory holds only landing zones, and the MMA reads one
// operand from shared memory instead of two (33 instead of 48 cycles per 128x64x8 MMA, scripts/ubench/mma_bench.cu).
// A first version issued the loads with per-thread cp.async from the same eight warps that transposed dY and made the lo
// images (82 -> 55 us for a 64x64 layer at M = 2^18); its role timing showed that serial loader chain -- not HBM latency,
// not the tensor pipe -- setting the pace (per tile: 540 cycles issuing copies per item, 1330 transposing the dY pair, 540
// per X lo image; the MMA warp waited 60 % of the time).  This kernel (55 -> 41 us) splits the stages:
//   warp  9   producer : one thread issues 2-D tensor-map loads, the dY tile as one raw 128 x 64 box, every X atom straight
//                        into its SWIZZLE_128B_ATOM_32B operand image; up to 3 dY tiles + 6 X atoms (192 KB) in flight
//   warps 0-3 dY       : raw tile -> registers -> tcgen05.st (thread == TMEM lane == feature, hi or lo half) + bias sums
//   warps 4-7 X lo     : lo image of each landed X atom into one of two lo slots
//   warp  8   MMA      : per X atom j: D[128, 64j..64j+64) += A_tmem[128, 128 samples] * [X_hi | X_lo] (16 k-steps, B addressed
//                        with a per-atom leading-dimension offset to the lo slot); D stays in TMEM for the whole kernel
// so the three stages run concurrently instead of back to back.
constexpr int kW3XHi = 6, kW3XLo = 2, kW3Y = 3;
constexpr int kW3Threads = 10 * 32;

__global__ void __launch_bounds__(kW3Threads, 1) wgrad_tma_kernel(const LinArgs A, const __grid_constant__ CUtensorMap tm_dy,
                                                                  const __grid_constant__ CUtensorMap tm_x,
                                                                  const __grid_constant__ CUtensorMap tm_xb) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const long long c0 = TNF_CLK();
  __shared__ uint64_t s_pfull[kW3Y], s_pempty[kW3Y], s_xland[kW3XHi], s_xfull[kW3XHi], s_xempty[kW3XHi], s_lempty[kW3XLo],
      s_afull[2], s_aempty[2], s_done;
  __shared__ uint32_t s_tmem;
  __shared__ float s_db[64];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kxa = A.kxa;                                  // X atoms of the first source
  const int kx = kxa + ((A.K - A.col_b + 31) >> 5);       // X atoms per tile (A.col_b == A.K: one source)
  const int n_sets = A.raw;             // A sets in tensor memory (2 when 2*128 + 64*kx <= 512 columns)
  constexpr int S = kW3XHi, L = kW3XLo;
  uint8_t* yraw = smem;                                  // raw dY tile slots: 128 rows x 256 B
  uint8_t* xhi = smem + kW3Y * 2 * kAtomBytes;           // X hi slot s at +s*kAtomBytes
  uint8_t* xlo = xhi + S * kAtomBytes;                   // X lo slot l at +l*kAtomBytes (above every hi slot)
  if (tid == 0) {
    for (int i = 0; i < kW3Y; ++i) { mbar_init(&s_pfull[i], 1); mbar_init(&s_pempty[i], 128); }
    for (int s = 0; s < S; ++s) { mbar_init(&s_xland[s], 1); mbar_init(&s_xfull[s], 128); mbar_init(&s_xempty[s], 1); }
    for (int l = 0; l < L; ++l) mbar_init(&s_lempty[l], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_afull[i], 128); mbar_init(&s_aempty[i], 1); }
    mbar_init(&s_done, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(&s_tmem, 512);
  if (tid < 64) s_db[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_a = s_tmem;                        // A set i: columns [128 i, 128 i + 128)
  const uint32_t tmem_d = s_tmem + 128 * n_sets;         // D: 64 columns per X atom
  const int my_tiles = (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 9) {
    // ===== producer =====
    if (lane == 0) {
      tma_prefetch_desc(&tm_dy);
      tma_prefetch_desc(&tm_x);
      if (kx > kxa) tma_prefetch_desc(&tm_xb);
      int xi = 0;
      long long c_pe = 0, c_xe = 0;
      Tm tm;
      const long long t_start = TNF_CLK();
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int row0 = (int)(((long long)blockIdx.x + (long long)tl * gridDim.x) * 128);
        const int ps = tl % kW3Y;
        tm.start();
        mbar_wait(&s_pempty[ps], ((tl / kW3Y) & 1) ^ 1);            // the dY warps have read the slot's previous tile
        tm.stop(c_pe);
        mbar_expect_tx(&s_pfull[ps], 2 * kAtomBytes);
        tma_load_2d(yraw + ps * 2 * kAtomBytes, &tm_dy, 0, row0, &s_pfull[ps]);
        for (int j = 0; j < kx; ++j, ++xi) {
          const int s = xi % S;
          tm.start();
          mbar_wait(&s_xempty[s], ((xi / S) & 1) ^ 1);              // the MMAs that read the slot's previous atom have completed
          tm.stop(c_xe);
          mbar_expect_tx(&s_xland[s], kAtomBytes);
          if (j < kxa) tma_load_2d(xhi + s * kAtomBytes, &tm_x, 32 * j, row0, &s_xland[s]);
          else tma_load_2d(xhi + s * kAtomBytes, &tm_xb, 32 * (j - kxa), row0, &s_xland[s]);
        }
      }
      if (g_dbg && blockIdx.x == 0) { g_dbg[16] = c_pe; g_dbg[17] = c_xe; g_dbg[18] = TNF_CLK() - t_start; }
    }
  } else if (warp < 4) {
    // ===== dY warps: transpose the raw tile into the A set =====
    const int q = warp;                  // TMEM lane quarter: lanes 32q..32q+31
    const bool is_hi = q < 2;            // quarters 0,1: dY_hi of features 0-31 / 32-63; quarters 2,3: dY_lo of the same
    float bsum = 0.f;
    long long c_pf = 0, c_ae = 0, c_work = 0;
    Tm tm;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int ps = tl % kW3Y, set = tl % n_sets;
      tm.start();
      mbar_wait(&s_pfull[ps], (tl / kW3Y) & 1);
      tm.stop(c_pf); tm.start();
      mbar_wait(&s_aempty[set], ((tl / n_sets) & 1) ^ 1);           // the MMAs that read this A set have completed
      tm.stop(c_ae); tm.start();
      tc_fence_after();
      const uint8_t* src = yraw + ps * 2 * kAtomBytes + (32 * (q & 1) + lane) * 4;
      const uint32_t taddr = tmem_a + ((uint32_t)(32 * q) << 16) + 128 * set;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = *reinterpret_cast<const float*>(src + (32 * c + i) * 256);
          if (is_hi) { bsum += x; v[i] = x; }
          else v[i] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        }
        tmem_st32(taddr + 32 * c, v);
      }
      mbar_arrive(&s_pempty[ps]);                                   // every value of the slot is in registers
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&s_afull[set]);
      tm.stop(c_work);
    }
    if (g_dbg && tid == 0 && blockIdx.x == 0) { g_dbg[19] = c_pf; g_dbg[20] = c_ae; g_dbg[21] = c_work; }
    if (A.db && is_hi) atomicAdd(&s_db[32 * (q & 1) + lane], bsum);
  } else if (warp < 8) {
    // ===== X lo warps =====
    const int t = tid - 128;
    const int n_x = my_tiles * kx;
    long long c_xl = 0, c_le = 0, c_work = 0;
    Tm tm;
    for (int xi = 0; xi < n_x; ++xi) {
      const int s = xi % S, l = xi % L;
      tm.start();
      mbar_wait(&s_xland[s], (xi / S) & 1);
      tm.stop(c_xl); tm.start();
      mbar_wait(&s_lempty[l], ((xi / L) & 1) ^ 1);                  // the MMAs that read the lo slot's previous image have completed
      tm.stop(c_le); tm.start();
      make_lo_atom<128>(xhi + s * kAtomBytes, xlo + l * kAtomBytes, t, true, nullptr);
      fence_async_smem();
      mbar_arrive(&s_xfull[s]);
      tm.stop(c_work);
    }
    if (g_dbg && t == 0 && blockIdx.x == 0) { g_dbg[22] = c_xl; g_dbg[23] = c_le; g_dbg[25] = c_work; }
  } else {
    // ===== MMA issuer =====
    const uint32_t idesc = instr_desc(128, 64, false, true);
    int xi = 0;
    long long w_y = 0, w_x = 0, w_iss = 0;
    const long long t_start = TNF_CLK();
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int set = tl % n_sets;
      { const long long t0 = TNF_CLK(); mbar_wait(&s_afull[set], (tl / n_sets) & 1); w_y += TNF_CLK() - t0; }
      if (tl == 0 && g_dbg && lane == 0 && blockIdx.x == 0) { g_dbg[26] = t_start - c0; g_dbg[27] = TNF_CLK() - c0; }
      for (int j = 0; j < kx; ++j, ++xi) {
        const int s = xi % S;
        { const long long t0 = TNF_CLK(); mbar_wait(&s_xfull[s], (xi / S) & 1); w_x += TNF_CLK() - t0; }
        const long long ti = TNF_CLK();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t xh = smem_u32(xhi + s * kAtomBytes);
          const uint32_t lbo = smem_u32(xlo + (xi % L) * kAtomBytes) - xh;   // MN block 1 of the B operand = the lo image
          const uint32_t aa = tmem_a + 128 * set;
#pragma unroll 4
          for (int kk = 0; kk < 16; ++kk)
            mma_tf32_ts(tmem_d + 64 * j, aa + 8 * kk, desc_mnmajor(xh, kk, lbo), idesc, !(tl == 0 && kk == 0));
          mma_commit(&s_xempty[s]);
          mma_commit(&s_lempty[xi % L]);
          if (j == kx - 1) mma_commit(&s_aempty[set]);
        }
        __syncwarp();
        w_iss += TNF_CLK() - ti;
      }
    }
    if (elect_one()) mma_commit(&s_done);
    __syncwarp();
    if (g_dbg && lane == 0 && blockIdx.x == 0) { g_dbg[8] = w_y; g_dbg[9] = w_x; g_dbg[10] = TNF_CLK() - t_start; g_dbg[11] = w_iss; g_dbg[28] = TNF_CLK() - c0; }
  }
  mbar_wait(&s_done, 0);   // every MMA of this CTA has completed
  tc_fence_after();
  __syncthreads();
  if (g_dbg && tid == 0 && blockIdx.x == 0) g_dbg[29] = TNF_CLK() - c0;
  if (A.db && tid < 64) atomicAdd(A.db + tid, s_db[tid]);
  if (warp < 4) {
    // Flush: every CTA adds its 64 x K partial into dW.  All CTAs get here at about the same time, and reductions to one
    // address are serialised in L2, so each CTA walks the (atom, 16-byte chunk) positions from its own starting point.
    // (Transposing the tile through shared memory for contiguous reductions was measured and is no faster.)
    const int n_out = (warp * 32 + lane) & 63;      // lanes n and n + 64 hold the dY_hi / dY_lo rows of feature n
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    // With a scratch tile (rows of 32 kx floats, atom j at columns 32 j) every reduction is a 16-byte one whatever the
    // alignment of dW's rows and of the second source's first column; the last CTA to finish adds the tile into dW.
    const bool to_scratch = A.scratch != nullptr;
    const bool vec = to_scratch || ((A.K & 3) == 0 && (A.col_b & 3) == 0 && (reinterpret_cast<uintptr_t>(A.dW) & 15u) == 0);
    const int j0 = blockIdx.x % kx, i0 = (blockIdx.x / kx) & 7;
    for (int jj = 0; jj < kx; ++jj) {
      const int j = (j0 + jj) % kx;
      const int cb = to_scratch ? 32 * j : (j < kxa ? 32 * j : A.col_b + 32 * (j - kxa));   // first column of the atom
      const int lim = to_scratch ? 32 * kx : (j < kxa ? A.col_b : A.K);                    // its source's columns end here
      float* dst = to_scratch ? A.scratch + (long long)n_out * (32 * kx) + cb : A.dW + (long long)n_out * A.K + cb;
#pragma unroll 2
      for (int ii = 0; ii < 8; ++ii) {
        const int c4 = 4 * ((i0 + ii) & 7);
        float v[4], u[4];
        tmem_ld4x2(taddr + 64 * j + c4, taddr + 64 * j + 32 + c4, v, u);   // x X_hi, x X_lo
        const int col = cb + c4;
        float* p = dst + c4;
        if (vec && col + 3 < lim) red_add_f4(p, make_float4(v[0] + u[0], v[1] + u[1], v[2] + u[2], v[3] + u[3]));   // one L2 reduction per 16 bytes
        else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e < lim) atomicAdd(p + e, v[e] + u[e]);
        }
      }
    }
  }
  tc_fence_before();
  if (A.scratch) {
    // last CTA: dW += scratch tile, which it leaves zeroed for the next launch (threadfence-reduction pattern)
    __shared__ int s_last;
    unsigned int* counter = reinterpret_cast<unsigned int*>(A.scratch + 64 * 32 * kx);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
      __threadfence();
      const int W = 32 * kx;
      for (int e = tid; e < 64 * W; e += kW3Threads) {
        const int r = e / W, sc = e - r * W, j = sc >> 5;
        const float v = __ldcg(A.scratch + e);
        A.scratch[e] = 0.f;
        const int col = j < kxa ? sc : A.col_b + (sc - 32 * kxa);
        if (col < (j < kxa ? A.col_b : A.K)) A.dW[(long long)r * A.K + col] += v;
      }
      if (tid == 0) *counter = 0u;
    }
  }
  __syncthreads();
  if (g_dbg && tid == 0 && blockIdx.x == 0) g_dbg[30] = TNF_CLK() - c0;
  if (warp == 8) tmem_dealloc(s_tmem, 512);
}

// Several 64-output weight gradients of ONE batch in one launch ("jobs": the heads' dW_i = dh_i^T h_{i-1} layers and the
// density head's first layer).  Each wgrad_tma_kernel launch costs a cold start (tensor-map fetch, TMEM allocation, the first
// loads' latency: ~6 us before the first MMA) and a tail (flush, drain), ~17 us of the 43 us an in-step launch takes, while
// its loop streams at 5 TB/s.  Here the roles of wgrad_tma_kernel simply walk a job list: the operand rings, the two A sets
// and every barrier keep counting across jobs, so the producer prefetches job j+1's first tiles while job j's last MMAs run;
// at a job boundary the dY warps wait for job j's MMAs, flush D (their own TMEM lanes) into dW_j and carry on with job j+1's
// first tile -- the issuer's next wait (A set full) orders the overwrite of D behind that flush.
constexpr int kWgMaxJobs = 4;
struct WgJobs {
  CUtensorMap dy[kWgMaxJobs];   // [M,64] dY of job j: one raw 128 x 64 box
  CUtensorMap x[kWgMaxJobs];    // [M,K_j] X of job j: 32-column atoms in the SWIZZLE_128B_ATOM_32B operand image
  float* dW[kWgMaxJobs];        // [64, K_j]
  float* db[kWgMaxJobs];        // [64] or null
  int K[kWgMaxJobs];
  int n_jobs;
  int n_tiles;
};

__global__ void __launch_bounds__(kW3Threads, 1) wgrad_tma_multi_kernel(const __grid_constant__ WgJobs J) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t s_pfull[kW3Y], s_pempty[kW3Y], s_xland[kW3XHi], s_xfull[kW3XHi], s_xempty[kW3XHi], s_lempty[kW3XLo],
      s_afull[2], s_aempty[2], s_done;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int S = kW3XHi, L = kW3XLo;
  uint8_t* yraw = smem;                                  // raw dY tile slots: 128 rows x 256 B
  uint8_t* xhi = smem + kW3Y * 2 * kAtomBytes;
  uint8_t* xlo = xhi + S * kAtomBytes;
  if (tid == 0) {
    for (int i = 0; i < kW3Y; ++i) { mbar_init(&s_pfull[i], 1); mbar_init(&s_pempty[i], 128); }
    for (int s = 0; s < S; ++s) { mbar_init(&s_xland[s], 1); mbar_init(&s_xfull[s], 128); mbar_init(&s_xempty[s], 1); }
    for (int l = 0; l < L; ++l) mbar_init(&s_lempty[l], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_afull[i], 128); mbar_init(&s_aempty[i], 1); }
    mbar_init(&s_done, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(&s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_a = s_tmem;                        // A set i: columns [128 i, 128 i + 128)
  const uint32_t tmem_d = s_tmem + 256;                  // D: 64 columns per X atom (<= 4 atoms)
  const int my_tiles = (J.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_jobs = J.n_jobs;

  if (warp == 9) {
    // ===== producer =====
    if (lane == 0) {
      for (int j = 0; j < n_jobs; ++j) { tma_prefetch_desc(&J.dy[j]); tma_prefetch_desc(&J.x[j]); }
      int xi = 0, tg = 0;
      for (int job = 0; job < n_jobs; ++job) {
        const int kx = (J.K[job] + 31) >> 5;
        for (int tl = 0; tl < my_tiles; ++tl, ++tg) {
          const int row0 = (int)(((long long)blockIdx.x + (long long)tl * gridDim.x) * 128);
          const int ps = tg % kW3Y;
          mbar_wait(&s_pempty[ps], ((tg / kW3Y) & 1) ^ 1);
          mbar_expect_tx(&s_pfull[ps], 2 * kAtomBytes);
          tma_load_2d(yraw + ps * 2 * kAtomBytes, &J.dy[job], 0, row0, &s_pfull[ps]);
          for (int a = 0; a < kx; ++a, ++xi) {
            const int s = xi % S;
            mbar_wait(&s_xempty[s], ((xi / S) & 1) ^ 1);
            mbar_expect_tx(&s_xland[s], kAtomBytes);
            tma_load_2d(xhi + s * kAtomBytes, &J.x[job], 32 * a, row0, &s_xland[s]);
          }
        }
      }
    }
  } else if (warp < 4) {
    // ===== dY warps: transpose the raw tile into the A set; at the end of a job, flush D =====
    const int q = warp;                  // TMEM lane quarter: lanes 32q..32q+31
    const bool is_hi = q < 2;            // quarters 0,1: dY_hi of features 0-31 / 32-63; quarters 2,3: dY_lo of the same
    const int n_out = (warp * 32 + lane) & 63;
    int tg = 0;
    for (int job = 0; job < n_jobs; ++job) {
      float bsum = 0.f;
      for (int tl = 0; tl < my_tiles; ++tl, ++tg) {
        const int ps = tg % kW3Y, set = tg & 1;
        mbar_wait(&s_pfull[ps], (tg / kW3Y) & 1);
        mbar_wait(&s_aempty[set], ((tg >> 1) & 1) ^ 1);            // the MMAs that read this A set have completed
        tc_fence_after();
        const uint8_t* src = yraw + ps * 2 * kAtomBytes + (32 * (q & 1) + lane) * 4;
        const uint32_t taddr = tmem_a + ((uint32_t)(32 * q) << 16) + 128 * set;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = *reinterpret_cast<const float*>(src + (32 * c + i) * 256);
            if (is_hi) { bsum += x; v[i] = x; }
            else v[i] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
          }
          tmem_st32(taddr + 32 * c, v);
        }
        mbar_arrive(&s_pempty[ps]);                                   // every value of the slot is in registers
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&s_afull[set]);
      }
      // ---- job boundary: this job's MMAs have completed -> add the CTA's 64 x K partial into dW (as wgrad_tma_kernel) ----
      mbar_wait(&s_done, job & 1);
      tc_fence_after();
      if (my_tiles > 0) {
        const int K = J.K[job], kx = (K + 31) >> 5;
        float* dW = J.dW[job];
        if (J.db[job] && is_hi) atomicAdd(J.db[job] + n_out, bsum);
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(dW) & 15u) == 0;
        const int j0 = blockIdx.x % kx, i0 = (blockIdx.x / kx) & 7;
        for (int jj = 0; jj < kx; ++jj) {
          const int a = (j0 + jj) % kx;
          float* dst = dW + (long long)n_out * K + 32 * a;
#pragma unroll 2
          for (int ii = 0; ii < 8; ++ii) {
            const int c4 = 4 * ((i0 + ii) & 7);
            float v[4], u[4];
            tmem_ld4x2(taddr + 64 * a + c4, taddr + 64 * a + 32 + c4, v, u);   // x X_hi, x X_lo
            const int col = 32 * a + c4;
            if (vec && col + 3 < K) red_add_f4(dst + c4, make_float4(v[0] + u[0], v[1] + u[1], v[2] + u[2], v[3] + u[3]));
            else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (col + e < K) atomicAdd(dst + c4 + e, v[e] + u[e]);
            }
          }
        }
      }
      tc_fence_before();
    }
  } else if (warp < 8) {
    // ===== X lo warps =====
    const int t = tid - 128;
    int n_x = 0;
    for (int job = 0; job < n_jobs; ++job) n_x += my_tiles * ((J.K[job] + 31) >> 5);
    for (int xi = 0; xi < n_x; ++xi) {
      const int s = xi % S, l = xi % L;
      mbar_wait(&s_xland[s], (xi / S) & 1);
      mbar_wait(&s_lempty[l], ((xi / L) & 1) ^ 1);
      make_lo_atom<128>(xhi + s * kAtomBytes, xlo + l * kAtomBytes, t, true, nullptr);
      fence_async_smem();
      mbar_arrive(&s_xfull[s]);
    }
  } else {
    // ===== MMA issuer =====
    const uint32_t idesc = instr_desc(128, 64, false, true);
    int xi = 0, tg = 0;
    for (int job = 0; job < n_jobs; ++job) {
      const int kx = (J.K[job] + 31) >> 5;
      for (int tl = 0; tl < my_tiles; ++tl, ++tg) {
        const int set = tg & 1;
        mbar_wait(&s_afull[set], (tg >> 1) & 1);   // (the first of a job also orders this job's MMAs behind the flush of the previous D)
        for (int a = 0; a < kx; ++a, ++xi) {
          const int s = xi % S;
          mbar_wait(&s_xfull[s], (xi / S) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t xh = smem_u32(xhi + s * kAtomBytes);
            const uint32_t lbo = smem_u32(xlo + (xi % L) * kAtomBytes) - xh;   // MN block 1 of the B operand = the lo image
            const uint32_t aa = tmem_a + 128 * set;
#pragma unroll 4
            for (int kk = 0; kk < 16; ++kk)
              mma_tf32_ts(tmem_d + 64 * a, aa + 8 * kk, desc_mnmajor(xh, kk, lbo), idesc, !(tl == 0 && kk == 0));
            mma_commit(&s_xempty[s]);
            mma_commit(&s_lempty[xi % L]);
            if (a == kx - 1) mma_commit(&s_aempty[set]);
          }
          __syncwarp();
        }
      }
      if (elect_one()) mma_commit(&s_done);   // every MMA of this job has completed when this arrives
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(s_tmem, 512);
}

// wgrad for layers with 128 outputs and <= 128 inputs (the Cobafa trunk) in ONE pass over dY and X.
// Two launches of wgrad_tma_kernel (one per 64-row half of dW) read X twice: 768 B per sample and launch where 1,024 B per
// sample would do for the whole layer.  Here all 128 output features are TMEM lanes: the transposed dY tile is written
// twice, as a hi image (columns 0..127 = the tile's 128 samples) and a lo image (128..255), by four warps (thread == lane
// == feature); per X atom the issuer runs, for each of the 16 k-steps,
//     D[:, 64 j .. 64 j + 64) += dY_hi^T [X_hi | X_lo]      (N = 64, the lo image as the second MN block, as above)
//     D[:, 64 j .. 64 j + 32) += dY_lo^T  X_hi              (N = 32)
// i.e. the three 3xTF32 products with two instructions; D (64 columns per atom, <= 256) fills the rest of tensor memory, so
// there is ONE A set: the transposition of tile t+1 waits for the MMAs of tile t (~1.5 of ~5 k cycles per tile, far below
// the tile's HBM time).  Shared memory: two raw dY tiles (64 KB each, two 64-column TMA boxes), four X atoms, two lo images.
constexpr int kG_Y = 2, kG_XH = 4, kG_XL = 2;
constexpr int kGThreads = 10 * 32;

__global__ void __launch_bounds__(kGThreads, 1) wgrad128_tma_kernel(const LinArgs A, const __grid_constant__ CUtensorMap tm_dy,
                                                                    const __grid_constant__ CUtensorMap tm_x) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t s_pfull[kG_Y], s_pempty[kG_Y], s_xland[kG_XH], s_xfull[kG_XH], s_xempty[kG_XH], s_lempty[kG_XL], s_afull,
      s_aempty, s_done;
  __shared__ uint32_t s_tmem;
  __shared__ float s_db[128];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kx = (A.K + 31) >> 5;                        // X atoms per tile (<= 4)
  constexpr int S = kG_XH, L = kG_XL;
  uint8_t* yraw = smem;                                  // raw dY tile slots: two [128 rows x 256 B] boxes (features 0-63 | 64-127)
  uint8_t* xhi = smem + kG_Y * 4 * kAtomBytes;
  uint8_t* xlo = xhi + S * kAtomBytes;
  if (tid == 0) {
    for (int i = 0; i < kG_Y; ++i) { mbar_init(&s_pfull[i], 1); mbar_init(&s_pempty[i], 128); }
    for (int s = 0; s < S; ++s) { mbar_init(&s_xland[s], 1); mbar_init(&s_xfull[s], 128); mbar_init(&s_xempty[s], 1); }
    for (int l = 0; l < L; ++l) mbar_init(&s_lempty[l], 1);
    mbar_init(&s_afull, 128); mbar_init(&s_aempty, 1); mbar_init(&s_done, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(&s_tmem, 512);
  if (tid < 128) s_db[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_ahi = s_tmem, tmem_alo = s_tmem + 128, tmem_d = s_tmem + 256;
  const int my_tiles = (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 9) {
    // ===== producer =====
    if (lane == 0) {
      tma_prefetch_desc(&tm_dy);
      tma_prefetch_desc(&tm_x);
      int xi = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int row0 = (int)(((long long)blockIdx.x + (long long)tl * gridDim.x) * 128);
        const int ps = tl % kG_Y;
        mbar_wait(&s_pempty[ps], ((tl / kG_Y) & 1) ^ 1);
        mbar_expect_tx(&s_pfull[ps], 4 * kAtomBytes);
        tma_load_2d(yraw + ps * 4 * kAtomBytes, &tm_dy, 0, row0, &s_pfull[ps]);
        tma_load_2d(yraw + ps * 4 * kAtomBytes + 2 * kAtomBytes, &tm_dy, 64, row0, &s_pfull[ps]);
        for (int j = 0; j < kx; ++j, ++xi) {
          const int s = xi % S;
          mbar_wait(&s_xempty[s], ((xi / S) & 1) ^ 1);
          mbar_expect_tx(&s_xland[s], kAtomBytes);
          tma_load_2d(xhi + s * kAtomBytes, &tm_x, 32 * j, row0, &s_xland[s]);
        }
      }
    }
  } else if (warp < 4) {
    // ===== dY warps: thread == TMEM lane == output feature 32 q + lane; hi and lo images of the tile's 128 samples =====
    const int q = warp;
    float bsum = 0.f;
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int ps = tl % kG_Y;
      mbar_wait(&s_pfull[ps], (tl / kG_Y) & 1);
      mbar_wait(&s_aempty, (tl & 1) ^ 1);                           // the MMAs of the previous tile have read the A images
      tc_fence_after();
      const uint8_t* src = yraw + ps * 4 * kAtomBytes + (q >> 1) * 2 * kAtomBytes + (32 * (q & 1) + lane) * 4;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = *reinterpret_cast<const float*>(src + (32 * c + i) * 256);
          bsum += v[i];
        }
        tmem_st32(tmem_ahi + lane_off + 32 * c, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = v[i] - __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
        tmem_st32(tmem_alo + lane_off + 32 * c, v);
      }
      mbar_arrive(&s_pempty[ps]);                                   // every value of the slot is in registers
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&s_afull);
    }
    if (A.db) atomicAdd(&s_db[32 * q + lane], bsum);
  } else if (warp < 8) {
    // ===== X lo warps =====
    const int t = tid - 128;
    const int n_x = my_tiles * kx;
    for (int xi = 0; xi < n_x; ++xi) {
      const int s = xi % S, l = xi % L;
      mbar_wait(&s_xland[s], (xi / S) & 1);
      mbar_wait(&s_lempty[l], ((xi / L) & 1) ^ 1);
      make_lo_atom<128>(xhi + s * kAtomBytes, xlo + l * kAtomBytes, t, true, nullptr);
      fence_async_smem();
      mbar_arrive(&s_xfull[s]);
    }
  } else {
    // ===== MMA issuer =====
    const uint32_t idesc64 = instr_desc(128, 64, false, true), idesc32 = instr_desc(128, 32, false, true);
    int xi = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      mbar_wait(&s_afull, tl & 1);
      for (int j = 0; j < kx; ++j, ++xi) {
        const int s = xi % S;
        mbar_wait(&s_xfull[s], (xi / S) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t xh = smem_u32(xhi + s * kAtomBytes);
          const uint32_t lbo = smem_u32(xlo + (xi % L) * kAtomBytes) - xh;   // MN block 1 of the B operand = the lo image
#pragma unroll 4
          for (int kk = 0; kk < 16; ++kk) {
            mma_tf32_ts(tmem_d + 64 * j, tmem_ahi + 8 * kk, desc_mnmajor(xh, kk, lbo), idesc64, !(tl == 0 && kk == 0));
            mma_tf32_ts(tmem_d + 64 * j, tmem_alo + 8 * kk, desc_mnmajor(xh, kk, lbo), idesc32, true);
          }
          mma_commit(&s_xempty[s]);
          mma_commit(&s_lempty[xi % L]);
          if (j == kx - 1) mma_commit(&s_aempty);
        }
        __syncwarp();
      }
    }
    if (elect_one()) mma_commit(&s_done);
    __syncwarp();
  }
  mbar_wait(&s_done, 0);   // every MMA of this CTA has completed
  tc_fence_after();
  __syncthreads();
  if (A.db && tid < 128) atomicAdd(A.db + tid, s_db[tid]);
  if (warp < 4 && my_tiles > 0) {
    // Flush as in wgrad_tma_kernel: every CTA adds its 128 x K partial into dW from its own starting (atom, chunk) position.
    const int n_out = warp * 32 + lane;
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    const bool vec = (A.K & 3) == 0 && (reinterpret_cast<uintptr_t>(A.dW) & 15u) == 0;
    const int j0 = blockIdx.x % kx, i0 = (blockIdx.x / kx) & 7;
    for (int jj = 0; jj < kx; ++jj) {
      const int j = (j0 + jj) % kx;
      float* dst = A.dW + (long long)n_out * A.K + 32 * j;
#pragma unroll 2
      for (int ii = 0; ii < 8; ++ii) {
        const int c4 = 4 * ((i0 + ii) & 7);
        float v[4], u[4];
        tmem_ld4x2(taddr + 64 * j + c4, taddr + 64 * j + 32 + c4, v, u);   // (dY_hi + dY_lo) x X_hi, dY_hi x X_lo
        const int col = 32 * j + c4;
        if (vec && col + 3 < A.K) red_add_f4(dst + c4, make_float4(v[0] + u[0], v[1] + u[1], v[2] + u[2], v[3] + u[3]));
        else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e < A.K) atomicAdd(dst + c4 + e, v[e] + u[e]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(s_tmem, 512);
}

constexpr int kHbU = 4;
template <int NH>  // n_head
__global__ void __launch_bounds__(256, 3) head_bwd_kernel(const float* __restrict__ H, long long ldh, const float* __restrict__ Wh,
                                                       const float* __restrict__ out, const float* __restrict__ dout,
                                                       float* __restrict__ dH, float* __restrict__ dWh, float* __restrict__ dbh,
                                                       long long M, int N, int act) {
  constexpr int n_head = NH;
  // N/4 lanes per row (one float4 of H each), 32/(N/4) rows per warp iteration; N in {32, 64, 128}
  __shared__ float s_dw[4 * 128];
  __shared__ float s_dbh[4];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 4 * 128; i += 256) s_dw[i] = 0.f;
  if (tid < 4) s_dbh[tid] = 0.f;
  __syncthreads();
  const int lpr = N >> 2;            // lanes per row
  const int rpw = 32 / lpr;          // rows per warp iteration
  const int sub = lane / lpr;        // row slot of this lane
  const int c4 = (lane % lpr) * 4;   // first column of this lane
  const long long warps_total = (long long)gridDim.x * 8;
  const long long wid = blockIdx.x * 8LL + (tid >> 5);
  float4 wreg[NH], dwacc[NH];
#pragma unroll
  for (int o = 0; o < NH; ++o) {
    wreg[o] = __ldg(reinterpret_cast<const float4*>(Wh + o * N + c4));
    dwacc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float dbacc[NH];
#pragma unroll
  for (int o = 0; o < NH; ++o) dbacc[o] = 0.f;
  // kHbU independent rows per lane and iteration: all their loads are issued before the first use, so a warp keeps
  // kHbU x 512 B (+ the out/dout scalars) in flight instead of one row's worth
  for (long long m0 = wid * rpw * kHbU; m0 < M; m0 += warps_total * rpw * kHbU) {
    float4 h[kHbU];
    float yv[kHbU][NH], gv[kHbU][NH];
#pragma unroll
    for (int u = 0; u < kHbU; ++u) {
      const long long m = m0 + u * rpw + sub;
      const bool ok = m < M;
      h[u] = ok ? ld_stream_f4(H + m * ldh + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < NH; ++o) {
        yv[u][o] = ok ? __ldg(out + m * n_head + o) : 0.f;
        gv[u][o] = ok ? __ldg(dout + m * n_head + o) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kHbU; ++u) {
      const long long m = m0 + u * rpw + sub;
      if (m >= M) continue;
      float dpre[NH];
#pragma unroll
      for (int o = 0; o < NH; ++o) {
        const float y = yv[u][o], g = gv[u][o];
        float d;
        if (act == 1) {
          // y = exp(x-1); backward multiplies by exp(clamp(x-1, -15, 15)) (src/models.py:52-53)
          d = g * fminf(fmaxf(y, 3.0590232050182579e-07f), 3269017.372472110639f);
        } else if (act == 2) {
          d = g * y * (1.f - y);
        } else {
          d = g;
        }
        dpre[o] = d;
        if (lane % lpr == 0) dbacc[o] += d;
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < NH; ++o) {
        acc.x = __fmaf_rn(dpre[o], wreg[o].x, acc.x); acc.y = __fmaf_rn(dpre[o], wreg[o].y, acc.y);
        acc.z = __fmaf_rn(dpre[o], wreg[o].z, acc.z); acc.w = __fmaf_rn(dpre[o], wreg[o].w, acc.w);
        dwacc[o].x = __fmaf_rn(dpre[o], h[u].x, dwacc[o].x); dwacc[o].y = __fmaf_rn(dpre[o], h[u].y, dwacc[o].y);
        dwacc[o].z = __fmaf_rn(dpre[o], h[u].z, dwacc[o].z); dwacc[o].w = __fmaf_rn(dpre[o], h[u].w, dwacc[o].w);
      }
      st_stream_f4(dH + m * ldh + c4, make_float4(h[u].x > 0.f ? acc.x : 0.f, h[u].y > 0.f ? acc.y : 0.f,
                                                  h[u].z > 0.f ? acc.z : 0.f, h[u].w > 0.f ? acc.w : 0.f));
    }
  }
#pragma unroll
  for (int o = 0; o < NH; ++o) {
    atomicAdd(&s_dw[o * 128 + c4 + 0], dwacc[o].x); atomicAdd(&s_dw[o * 128 + c4 + 1], dwacc[o].y);
    atomicAdd(&s_dw[o * 128 + c4 + 2], dwacc[o].z); atomicAdd(&s_dw[o * 128 + c4 + 3], dwacc[o].w);
    atomicAdd(&s_dbh[o], dbacc[o]);
  }
  __syncthreads();
  for (int i = tid; i < n_head * N; i += 256) atomicAdd(dWh + i, s_dw[(i / N) * 128 + (i % N)]);
  if (tid < n_head) atomicAdd(dbh + tid, s_dbh[tid]);
}

// Input row of VanillaColorDecoder (src/models.py:87): [PE(d) | d | features] written in one pass, 32 lanes per row.
// PE layout per coordinate c: sin(2^k pi d_c), k < n_freqs, then cos(...) (src/models.py:36-39); the frequencies are
// fl32(pi) * 2^k like the reference's `2**arange(n) * torch.pi` buffer, products rounded to fp32 before sin/cos.
__global__ void __launch_bounds__(256) color_input_kernel(const float* __restrict__ dirs, long long ld_dirs,
                                                          const float* __restrict__ feats, long long ld_feats, int n_freqs,
                                                          int feat_dim, float* __restrict__ out, long long ld_out, long long n) {
  const int pe = 6 * n_freqs;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* d = dirs + row * ld_dirs;
  float* o = out + row * ld_out;
  for (int j = lane; j < 3 * n_freqs; j += 32) {  // one (coordinate, frequency) pair per lane: sin and cos together
    const int c = j / n_freqs, k = j % n_freqs;
    const float arg = __fmul_rn(__ldg(d + c), ldexpf(3.14159274101257324219f, k));
    float sv, cv;
    sincosf(arg, &sv, &cv);
    o[c * 2 * n_freqs + k] = sv;
    o[c * 2 * n_freqs + n_freqs + k] = cv;
  }
  for (int j = pe + lane; j < ld_out; j += 32) {
    float v = 0.f;
    if (j < pe + 3) v = __ldg(d + (j - pe));
    else if (j < pe + 3 + feat_dim) v = __ldg(feats + row * ld_feats + (j - pe - 3));
    o[j] = v;
  }
}

// The same row without feature columns ([PE_8(d) | d | 0], 52 floats: the colour input when the feature part is read from the
// feature rows, tnf_heads_fwd xc_cols < k0).  The generic kernel spends ~300 warp instructions per row on loop and index
// overhead with 24 of 32 lanes in sincosf; here four threads share a row: thread c < 3 evaluates the 8 frequencies of
// coordinate c and writes its 16 outputs as four 16-byte stores, thread 3 writes [d | 0].  Same arguments, same sincosf.
__global__ void __launch_bounds__(256) pe8_dirs_kernel(const float* __restrict__ dirs, long long ld_dirs, float* __restrict__ out,
                                                       long long ld_out, long long n) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long row = t >> 2;
  const int c = (int)(t & 3);
  if (row >= n) return;
  const float* d = dirs + row * ld_dirs;
  float* o = out + row * ld_out;
  if (c == 3) {
    *reinterpret_cast<float4*>(o + 48) = make_float4(__ldg(d), __ldg(d + 1), __ldg(d + 2), 0.f);
    for (int j = 52; j < ld_out; ++j) o[j] = 0.f;
    return;
  }
  const float x = __ldg(d + c);
  float sv[8], cv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) sincosf(__fmul_rn(x, ldexpf(3.14159274101257324219f, k)), &sv[k], &cv[k]);
  float4* o4 = reinterpret_cast<float4*>(o + 16 * c);
  o4[0] = make_float4(sv[0], sv[1], sv[2], sv[3]);
  o4[1] = make_float4(sv[4], sv[5], sv[6], sv[7]);
  o4[2] = make_float4(cv[0], cv[1], cv[2], cv[3]);
  o4[3] = make_float4(cv[4], cv[5], cv[6], cv[7]);
}

int check_lin(long long M, int N, int K) {
  TNF_REQUIRE(M >= 0, "negative M");
  TNF_REQUIRE(N >= 8 && N <= 128 && N % 8 == 0, "out_features must be a multiple of 8 in [8,128] (got %d)", N);
  TNF_REQUIRE(K >= 1 && K <= 32 * kMaxKAtoms, "in_features must be in [1,%d] (got %d)", 32 * kMaxKAtoms, K);
  return TNF_OK;
}
bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// row-major fp32 [rows, cols] with leading dimension ld -> tensor map with a box_cols x 128 box (rows/cols outside the matrix
// read as zero)
int make_box_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, CUtensorMapSwizzle swz) {
  static EncodeTiledFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return reinterpret_cast<EncodeTiledFn>(fn);
  }();
  TNF_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TNF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TNF_OK;
}

// 64-wide layers: A operand in tensor memory, loads by the copy engine (wgrad_tma_kernel).  X = [xa (ka columns) | xb (kb columns)].
int launch_wgrad_tma(const float* dy, int64_t lddy, const float* xa, int64_t ldxa, int ka, const float* xb, int64_t ldxb, int kb,
                     float* dweight, float* dbias, float* scratch, int64_t m, cudaStream_t st) {
  LinArgs A{};
  A.dW = dweight; A.db = dbias; A.M = m; A.N = 64; A.K = ka + kb; A.scratch = scratch;
  A.n_tiles = (int)ceil_div(m, 128);
  A.kxa = (ka + 31) / 32;
  A.col_b = ka;
  const int kx = A.kxa + (kb + 31) / 32;
  TNF_REQUIRE(kx >= 1 && kx <= kMaxKAtoms, "wgrad: at most %d input atoms (got %d)", kMaxKAtoms, kx);
  A.raw = (2 * 128 + 64 * kx <= 512) ? 2 : 1;
  CUtensorMap tm_dy, tm_x, tm_xb;
  int rc = make_box_map(&tm_dy, dy, m, 64, lddy, 64, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != TNF_OK) return rc;
  rc = make_box_map(&tm_x, xa, m, ka, ldxa, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != TNF_OK) return rc;
  rc = kb > 0 ? make_box_map(&tm_xb, xb, m, kb, ldxb, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) : TNF_OK;
  if (rc != TNF_OK) return rc;
  if (kb == 0) tm_xb = tm_x;
  const size_t smem = (size_t)(2 * kW3Y + kW3XHi + kW3XLo) * kAtomBytes + 1024;
  static PerDeviceOnce configured_tma{};
  if (configured_tma.pending()) {
    TNF_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured_tma.mark();
  }
  const int grid = A.n_tiles < sm_count() ? A.n_tiles : sm_count();
  wgrad_tma_kernel<<<grid, kW3Threads, smem, st>>>(A, tm_dy, tm_x, tm_xb);
  TNF_LAUNCH_CHECK("linear_wgrad_tma_kernel");
  return TNF_OK;
}

// several 64-output layers of one batch in one launch (wgrad_tma_multi_kernel)
int launch_wgrad_multi(int n_jobs, const float* const* dy, const int64_t* lddy, const float* const* x, const int64_t* ldx,
                       const int32_t* k, float* const* dweight, float* const* dbias, int64_t m, cudaStream_t st) {
  WgJobs J{};
  J.n_jobs = n_jobs;
  J.n_tiles = (int)ceil_div(m, 128);
  for (int j = 0; j < n_jobs; ++j) {
    int rc = make_box_map(&J.dy[j], dy[j], m, 64, lddy[j], 64, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != TNF_OK) return rc;
    rc = make_box_map(&J.x[j], x[j], m, k[j], ldx[j], 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc != TNF_OK) return rc;
    J.dW[j] = dweight[j]; J.db[j] = dbias ? dbias[j] : nullptr; J.K[j] = k[j];
  }
  const size_t smem = (size_t)(2 * kW3Y + kW3XHi + kW3XLo) * kAtomBytes + 1024;
  static PerDeviceOnce configured_multi{};
  if (configured_multi.pending()) {
    TNF_CUDA(cudaFuncSetAttribute(wgrad_tma_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured_multi.mark();
  }
  const int grid = J.n_tiles < sm_count() ? J.n_tiles : sm_count();
  wgrad_tma_multi_kernel<<<grid, kW3Threads, smem, st>>>(J);
  TNF_LAUNCH_CHECK("linear_wgrad_tma_multi_kernel");
  return TNF_OK;
}

// 128 outputs, <= 128 inputs: one pass over dY and X (wgrad128_tma_kernel)
int launch_wgrad128(const float* dy, int64_t lddy, const float* x, int64_t ldx, int k, float* dweight, float* dbias, int64_t m,
                    cudaStream_t st) {
  LinArgs A{};
  A.dW = dweight; A.db = dbias; A.M = m; A.N = 128; A.K = k;
  A.n_tiles = (int)ceil_div(m, 128);
  CUtensorMap tm_dy, tm_x;
  int rc = make_box_map(&tm_dy, dy, m, 128, lddy, 64, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != TNF_OK) return rc;
  rc = make_box_map(&tm_x, x, m, k, ldx, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != TNF_OK) return rc;
  const size_t smem = (size_t)(4 * kG_Y + kG_XH + kG_XL) * kAtomBytes + 1024;
  static PerDeviceOnce configured128{};
  if (configured128.pending()) {
    TNF_CUDA(cudaFuncSetAttribute(wgrad128_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured128.mark();
  }
  const int grid = A.n_tiles < sm_count() ? A.n_tiles : sm_count();
  wgrad128_tma_kernel<<<grid, kGThreads, smem, st>>>(A, tm_dy, tm_x);
  TNF_LAUNCH_CHECK("linear_wgrad128_tma_kernel");
  return TNF_OK;
}

// shared-memory plan of linear_kernel: weight images + H hi slots + L lo slots (16 KB each) + 16 KB epilogue staging.
// L = 3 lets the lo pass run two atoms ahead of the tensor core; every further hi slot is another atom of global loads
// in flight (H - L of them, up to 4).
size_t plan_smem(int n, int k, int* hi_slots, int* lo_slots) {
  const size_t w_stride = ((((size_t)(n + 15) & ~15) * 128) + 1023) & ~(size_t)1023;
  const size_t fixed = 2 * ((k + 31) / 32) * w_stride + 9 * 4096 + 1024;  // + 8 warp transpose buffers + head partials
  const size_t budget = 226 * 1024;
  *hi_slots = *lo_slots = 0;
  if (fixed + 3 * kAtomBytes > budget) return 0;
  const int slots = (int)((budget - fixed) / kAtomBytes);
  int lo = slots >= 6 ? kMaxLo : (slots >= 4 ? 2 : 1);
  int hi = slots - lo;
  if (hi > lo + 4) hi = lo + 4;
  if (hi > kMaxHi) hi = kMaxHi;
  *hi_slots = hi;
  *lo_slots = lo;
  return fixed + (size_t)(hi + lo) * kAtomBytes;
}

template <typename Kern>
int launch_lin(Kern kern, LinArgs& A, cudaStream_t st, const char* name) {
  const size_t smem = plan_smem(A.N, A.K, &A.ring, &A.raw);
  TNF_REQUIRE(A.ring >= 2 && A.raw >= 1, "layer too large for the shared-memory plan (n=%d, k=%d)", A.N, A.K);
  static const void* configured_all[64][8] = {};   // per device (the attribute is per device/context)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  const void** configured = configured_all[dev];
  bool done = false;
  for (int i = 0; i < 8; ++i) done |= (configured[i] == (const void*)kern);
  if (!done) {
    TNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    for (int i = 0; i < 8; ++i) if (!configured[i]) { configured[i] = (const void*)kern; break; }
  }
  const int grid = A.n_tiles < sm_count() ? A.n_tiles : sm_count();
  kern<<<grid, kWsThreads, smem, st>>>(A);
  TNF_LAUNCH_CHECK(name);
  return TNF_OK;
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_debug_role_timing(long long* buf) {  // diagnostics only (not part of the public header)
  return cudaMemcpyToSymbol(tnf::g_dbg, &buf, sizeof(buf)) == cudaSuccess ? 0 : -2;
}

extern "C" int tnf_linear_fwd(const float* x, int64_t ldx, const float* weight, const float* bias, float* y, int64_t ldy,
                              int64_t m, int32_t n, int32_t k, int32_t relu, const float* head_w, const float* head_b,
                              float* head_out, int32_t n_head, int32_t head_act, void* stream) {
  using namespace tnf;
  int rc = check_lin(m, n, k);
  if (rc != TNF_OK || m == 0) return rc;
  TNF_REQUIRE(x && weight && (y || n_head > 0), "null pointer");
  TNF_REQUIRE(n % 32 == 0, "forward needs out_features to be a multiple of 32");
  TNF_REQUIRE(al16(x) && ldx % 4 == 0 && (!y || (al16(y) && ldy % 4 == 0)), "x/y must be 16-byte aligned with ld %% 4 == 0");
  TNF_REQUIRE(n_head >= 0 && n_head <= 4 && (n_head == 0 || (head_w && head_b && head_out)), "bad head");
  if (n_head == 0 && !variant(kVariantNoWstat) && wstat_linear_supported(0, m, n, k, x, ldx, weight, y, ldy, nullptr, 0))
    return launch_wstat_linear(0, x, ldx, weight, k, bias, relu, nullptr, 0, y, ldy, m, static_cast<cudaStream_t>(stream));
  LinArgs A{};
  A.X = x; A.ldx = ldx; A.W = weight; A.bias = bias; A.Y = y; A.ldy = ldy; A.M = m; A.N = n; A.K = k; A.relu = relu;
  A.head_w = head_w; A.head_b = head_b; A.head_out = head_out; A.n_head = n_head; A.head_act = head_act;
  A.n_tiles = (int)ceil_div(m, 128);
  return launch_lin(linear_kernel<0>, A, static_cast<cudaStream_t>(stream), "linear_fwd_kernel");
}

extern "C" int tnf_linear_bwd_data(const float* dy, int64_t lddy, const float* weight, float* dx, int64_t lddx,
                                   const float* relu_src, int64_t ldrs, int64_t m, int32_t n, int32_t k, void* stream) {
  using namespace tnf;
  int rc = check_lin(m, n, k);
  if (rc != TNF_OK || m == 0) return rc;
  TNF_REQUIRE(dy && weight && dx, "null pointer");
  TNF_REQUIRE(n % 32 == 0, "dgrad needs out_features to be a multiple of 32");
  TNF_REQUIRE(al16(dy) && lddy % 4 == 0 && al16(dx) && lddx % 4 == 0, "dy/dx must be 16-byte aligned with ld %% 4 == 0");
  TNF_REQUIRE(!relu_src || (al16(relu_src) && ldrs % 4 == 0), "relu_src must be 16-byte aligned with ld %% 4 == 0");
  if (!variant(kVariantNoWstat) && wstat_linear_supported(1, m, n, k, dy, lddy, weight, dx, lddx, relu_src, ldrs))
    return launch_wstat_linear(1, dy, lddy, weight, k, nullptr, 0, relu_src, ldrs, dx, lddx, m, static_cast<cudaStream_t>(stream));
  LinArgs A{};
  A.X = dy; A.ldx = lddy; A.W = weight; A.Y = dx; A.ldy = lddx; A.X2 = relu_src; A.ldx2 = ldrs; A.M = m; A.N = n; A.K = k;
  A.n_tiles = (int)ceil_div(m, 128);
  return launch_lin(linear_kernel<1>, A, static_cast<cudaStream_t>(stream), "linear_dgrad_kernel");
}

extern "C" int tnf_linear_bwd_weight(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dweight, float* dbias,
                                     int64_t m, int32_t n, int32_t k, void* stream) {
  using namespace tnf;
  int rc = check_lin(m, n, k);
  if (rc != TNF_OK || m == 0) return rc;
  TNF_REQUIRE(dy && x && dweight, "null pointer");
  TNF_REQUIRE(n % 32 == 0 && n <= 128, "wgrad needs out_features in {32,64,96,128}");
  TNF_REQUIRE(al16(dy) && lddy % 4 == 0 && al16(x) && ldx % 4 == 0, "dy/x must be 16-byte aligned with ld %% 4 == 0");
  LinArgs A{};
  A.X = dy; A.ldx = lddy; A.X2 = x; A.ldx2 = ldx; A.dW = dweight; A.db = dbias; A.M = m; A.N = n; A.K = k;
  A.n_tiles = (int)ceil_div(m, 128);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool want_ss = variant(kVariantWgradSS) == 1;   // diagnostics: the both-operands-in-shared-memory kernel
  if (n == 64 && m < (1LL << 31) - 256 && !want_ss)
    return launch_wgrad_tma(dy, lddy, x, ldx, k, nullptr, 0, 0, dweight, dbias, nullptr, m, st);
  if (n == 128 && k <= 128 && m < (1LL << 31) - 256 && !want_ss && !variant(kVariantNoWstat))
    return launch_wgrad128(dy, lddy, x, ldx, k, dweight, dbias, m, st);   // the whole layer in one pass (Cobafa trunk)
  if (n == 128 && (k + 31) / 32 <= kMaxKAtoms && m < (1LL << 31) - 256 && !want_ss) {
    // 128 output features (the Cobafa trunk, src/models.py:254-266) = two 64-row halves of dW through the tensor-memory /
    // TMA kernel: 2 x ~60 us against ~200 us for the both-operands-in-shared-memory kernel below (M = 2^18, K = 128)
    for (int h = 0; h < 2; ++h) {
      rc = launch_wgrad_tma(dy + 64 * h, lddy, x, ldx, k, nullptr, 0, 0, dweight + (size_t)64 * h * k, dbias ? dbias + 64 * h : nullptr,
                            nullptr, m, st);
      if (rc != TNF_OK) return rc;
    }
    return TNF_OK;
  }
  // shared-memory plan: dY sets (hi + lo images of n/32 atoms each; two sets when they leave room for >= 2 X stages)
  // + S X stages (32 KB each)
  const size_t set_bytes = (size_t)(2 * ((n + 31) / 32)) * kAtomBytes;
  const size_t budget = 226 * 1024 - 1024;
  A.raw = (2 * set_bytes + 2 * 2 * kAtomBytes <= budget) ? 2 : 1;
  TNF_REQUIRE((size_t)A.raw * set_bytes + 2 * kAtomBytes <= budget, "layer too large for the wgrad shared-memory plan (n=%d)", n);
  int st_ = (int)((budget - (size_t)A.raw * set_bytes) / (2 * kAtomBytes));
  A.ring = st_ > kMaxStages ? kMaxStages : st_;
  const size_t smem = (size_t)A.raw * set_bytes + (size_t)A.ring * 2 * kAtomBytes + 1024;
  static PerDeviceOnce configured{};
  if (configured.pending()) {
    TNF_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured.mark();
  }
  const int grid = A.n_tiles < sm_count() ? A.n_tiles : sm_count();
  wgrad_kernel<<<grid, kWgThreads, smem, st>>>(A);
  TNF_LAUNCH_CHECK("linear_wgrad_kernel");
  return TNF_OK;
}

extern "C" int64_t tnf_wgrad_cat_scratch_bytes(int32_t ka, int32_t kb) {
  return (int64_t)(64 * 32 * ((ka + 31) / 32 + (kb + 31) / 32) + 4) * 4;
}

extern "C" int tnf_linear_bwd_weight_cat(const float* dy, int64_t lddy, const float* xa, int64_t ldxa, int32_t ka, const float* xb,
                                         int64_t ldxb, int32_t kb, float* dweight, float* dbias, int64_t m, int32_t n,
                                         float* scratch, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(m >= 0 && ka >= 1 && kb >= 1, "bad sizes");
  TNF_REQUIRE(n == 64, "the two-source weight gradient is implemented for 64 output features");
  TNF_REQUIRE(m < (1LL << 31) - 256, "too many rows for the tensor-map coordinates");
  if (m == 0) return TNF_OK;
  TNF_REQUIRE(dy && xa && xb && dweight, "null pointer");
  TNF_REQUIRE(al16(dy) && lddy % 4 == 0 && al16(xa) && ldxa % 4 == 0 && al16(xb) && ldxb % 4 == 0,
              "dy/xa/xb must be 16-byte aligned with ld %% 4 == 0");
  TNF_REQUIRE(!scratch || al16(scratch), "scratch must be 16-byte aligned");
  return launch_wgrad_tma(dy, lddy, xa, ldxa, ka, xb, ldxb, kb, dweight, dbias, scratch, m, static_cast<cudaStream_t>(stream));
}

extern "C" int tnf_linear_bwd_weight_multi(int32_t n_jobs, const float* const* dy, const int64_t* lddy, const float* const* x,
                                           const int64_t* ldx, const int32_t* k, float* const* dweight, float* const* dbias,
                                           int64_t m, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n_jobs >= 1 && n_jobs <= kWgMaxJobs, "n_jobs must be in [1,%d]", kWgMaxJobs);
  TNF_REQUIRE(m >= 0, "negative m");
  if (m == 0) return TNF_OK;
  TNF_REQUIRE(dy && lddy && x && ldx && k && dweight, "null table");
  TNF_REQUIRE(m < (1LL << 31) - 256, "too many rows for the tensor-map coordinates");
  for (int j = 0; j < n_jobs; ++j) {
    TNF_REQUIRE(dy[j] && x[j] && dweight[j], "null pointer in job %d", j);
    TNF_REQUIRE(k[j] >= 1 && k[j] <= 128, "job %d: in_features must be in [1,128] (got %d)", j, k[j]);
    TNF_REQUIRE(al16(dy[j]) && lddy[j] % 4 == 0 && lddy[j] >= 64 && al16(x[j]) && ldx[j] % 4 == 0 && ldx[j] >= k[j],
                "job %d: dy/x must be 16-byte aligned with leading dimensions multiple of 4", j);
  }
  return launch_wgrad_multi(n_jobs, dy, lddy, x, ldx, k, dweight, dbias, m, static_cast<cudaStream_t>(stream));
}

extern "C" int tnf_color_input(const float* dirs, int64_t ld_dirs, const float* feats, int64_t ld_feats, int32_t n_freqs,
                               int32_t feat_dim, float* out, int64_t ld_out, int64_t n, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(n >= 0 && n_freqs >= 0 && n_freqs <= 24 && feat_dim >= 0, "bad sizes");
  if (n == 0) return TNF_OK;
  TNF_REQUIRE(dirs && (feats || feat_dim == 0) && out, "null pointer");
  TNF_REQUIRE(ld_out >= 6 * n_freqs + 3 + feat_dim, "ld_out too small");
  if (n_freqs == 8 && feat_dim == 0 && ld_out >= 52 && ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    pe8_dirs_kernel<<<(unsigned)ceil_div(n * 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dirs, ld_dirs, out, ld_out, n);
    TNF_LAUNCH_CHECK("pe8_dirs_kernel");
    return TNF_OK;
  }
  color_input_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dirs, ld_dirs, feats, ld_feats, n_freqs, feat_dim, out, ld_out, n);
  TNF_LAUNCH_CHECK("color_input_kernel");
  return TNF_OK;
}

extern "C" int tnf_head_bwd(const float* h, int64_t ldh, const float* head_w, const float* out, const float* dout, float* dh,
                            float* dhead_w, float* dhead_b, int64_t m, int32_t n, int32_t n_head, int32_t head_act,
                            void* stream) {
  using namespace tnf;
  TNF_REQUIRE(m >= 0 && (n == 32 || n == 64 || n == 128) && n_head >= 1 && n_head <= 4, "bad head shape");
  TNF_REQUIRE((ldh & 3) == 0 && ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(head_w)) & 15u) == 0,
              "h/dh/head_w must be 16-byte aligned with ldh %% 4 == 0");
  if (m == 0) return TNF_OK;
  TNF_REQUIRE(h && head_w && out && dout && dh && dhead_w && dhead_b, "null pointer");
  // one resident wave: every block strides over the rows, so the per-block flush of the head-weight partials is paid once
  using HeadBwd = void (*)(const float*, long long, const float*, const float*, const float*, float*, float*, float*, long long, int, int);
  static const HeadBwd kerns[4] = {head_bwd_kernel<1>, head_bwd_kernel<2>, head_bwd_kernel<3>, head_bwd_kernel<4>};
  static thread_local int occ[4] = {0, 0, 0, 0};
  const HeadBwd kern = kerns[n_head - 1];
  if (!occ[n_head - 1]) TNF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[n_head - 1], kern, 256, 0));
  const int64_t rows_per_iter = 8 * (32 / (n / 4)) * kHbU;
  const int64_t want = ceil_div(m, rows_per_iter), cap = (int64_t)(occ[n_head - 1] > 0 ? occ[n_head - 1] : 1) * sm_count();
  const int grid = (int)(want < cap ? want : cap);
  kern<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(h, ldh, head_w, out, dout, dh, dhead_w, dhead_b, m, n, head_act);
  TNF_LAUNCH_CHECK("head_bwd_kernel");
  return TNF_OK;
}
