// a15/a16 -- both decoder heads of a sample tile in ONE persistent tcgen05 kernel (forward):
//   sigma  = truncated_exp(W_s1 relu(W_s0 f + b_s0) + b_s1 - 1)                     VanillaOpacityDecoder, src/models.py:70-77
//   rgb    = sigmoid(W_c4 relu(W_c3 relu(W_c2 relu(W_c1 relu(W_c0 x + b)...))))     VanillaColorDecoder,   src/models.py:79-89
// with x = [PE(d) | d | f] (the colour-input row, tnf_color_input) and f the feature row.  The row need not be materialised:
// with xc_cols < k0 the buffer `xc` holds only its first xc_cols columns ([PE(d) | d]) and the rest IS the feature row, whose
// atoms are then multiplied into both heads' layer-0 accumulators (the concatenation of src/models.py:87 becomes a split
// of W_c0's columns).
//
// The per-layer kernels (mlp.cu) stream every hidden activation through HBM and pay a launch per layer.  Here a
// 128-sample tile stays on chip from the inputs to sigma/rgb:
//   * layer 0 of both heads: the input atoms (128 rows x 32 fp32) are copied global -> shared by cp.async straight into
//     the swizzled K-major operand image (the raw fp32 tile IS the tf32 "hi" operand; the loader warps only add the
//     "lo" = x - trunc(x) image) and multiplied SS-mode into two TMEM accumulators (colour, sigma);
//   * hidden layers 1..3: the epilogue warps read the accumulator (tcgen05.ld), add bias, ReLU, split hi/lo and write
//     the result back to TENSOR MEMORY (tcgen05.st) where the next layer's MMAs take it as their A operand (TS-mode):
//     no shared-memory round trip for activations, and the tensor core no longer competes for shared-memory bandwidth;
//   * the weights do not fit next to the rings (224 KB as hi/lo tf32 images), so a tiny pack kernel writes them once per
//     call as 16 KB chunks (64 x 32 hi + lo, swizzled) in the order the tile loop consumes them and the loader warps
//     stream the chunks through a shared-memory ring with cp.async (L2-resident, 224 KB per tile);
//   * the MMA warp follows a static software pipeline: the three hidden layers of tile t are interleaved with layer 0 of
//     tile t+1 (double-buffered accumulators), so the tensor core works while the epilogue warps turn a layer around;
//   * fp32-grade accuracy by 3xTF32 (hi*hi + lo*hi + hi*lo), as in mlp.cu.
// The hidden activations are also written to HBM ([M,64] each) because the backward kernels read them.
#include <cuda.h>
#include "common.cuh"
#include "tc.cuh"

namespace tnf {
namespace {

constexpr int kLoadWarps = 8;        // loader warps (input atoms: cp.async + lo image)
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kMmaWarp = kLoadWarps + 8;
constexpr int kTmaWarp = kLoadWarps + 9;
constexpr int kHThreads = (kLoadWarps + 10) * 32;  // lo-pass warps | 4 colour epilogue | 4 sigma epilogue | MMA issuer | TMA producer
constexpr int kChunkBytes = 16384;  // weight chunk: [hi 64 rows x 128 B][lo 64 rows x 128 B]
constexpr int kAH = 6, kAL = 2, kWN = 4;  // ring slots: input atoms (hi image, TMA target), lo images, weight chunks
constexpr int kHid = 64;

__device__ long long* g_hdbg = nullptr;   // diagnostics: wait cycles per role (CTA 0)
#ifdef TNF_HEADS_TIMING
#define HT0() long long _t0 = clock64()
#define HT1(acc) acc += clock64() - _t0
#define HCLK() clock64()
#else
#define HT0() do {} while (0)
#define HT1(acc) do {} while (0)
#define HCLK() 0LL
#endif

struct HeadsArgs {
  const float* feats; long long ld_feats; int F;
  const float* xc; long long ld_xc; int K0;
  int kxp, kfc;   // layer-0 items: kxp atoms of xc (colour), then kfc feature atoms (colour; 0 when xc holds the whole row), then kf feature atoms (sigma)
  const uint8_t* wimg;
  const float* bias_c[4]; const float* bias_s;
  const float* head_c_w; const float* head_c_b; const float* head_s_w; const float* head_s_b;
  float* h[4]; float* hs; float* rgb; float* sigma;
  long long M; int n_tiles;
};

// ---- the static schedule ---------------------------------------------------------------------------------------------
// Layer-0 work of a tile is n0 = nc + kf "items" (nc colour-head atoms, then kf feature atoms for sigma); the hidden layers are
// three more units.  Units in issue order for a CTA with T tiles:
//   prologue: L0(tile 0) items 0..n0-1
//   tile t  : H1(t) | L0(t+1) items 0,1 | H2(t) | L0(t+1) items 2,3 | H3(t) | L0(t+1) items 4..n0-1     (no L0 after the last tile)
struct Unit { int hidden; int tile; int q; };  // hidden: 0 = layer-0 item q of `tile`, else hidden layer index 1..3 of `tile`
__device__ __forceinline__ int n_units(int T, int n0) { return T <= 0 ? 0 : n0 + (T - 1) * (3 + n0) + 3; }
__device__ __forceinline__ Unit decode_unit(int u, int T, int n0) {
  Unit x;
  if (u < n0) { x.hidden = 0; x.tile = 0; x.q = u; return x; }
  const int U = 3 + n0;
  int v = u - n0;
  int t = v / U, r = v - t * U;
  if (t >= T - 1) { t = T - 1; r = v - t * U; x.hidden = r + 1; x.tile = t; x.q = 0; return x; }
  x.tile = t;
  x.q = 0;
  if (r == 0) x.hidden = 1;
  else if (r <= 2) { x.hidden = 0; x.tile = t + 1; x.q = r - 1; }
  else if (r == 3) x.hidden = 2;
  else if (r <= 5) { x.hidden = 0; x.tile = t + 1; x.q = r - 2; }
  else if (r == 6) x.hidden = 3;
  else { x.hidden = 0; x.tile = t + 1; x.q = r - 3; }
  return x;
}
// chunk index inside the packed image (steady-state order: [W1a W1b][items 0,1][W2a W2b][items 2,3][W3a W3b][items 4..])
__host__ __device__ __forceinline__ int chunk_of_item(int q) { return q < 2 ? 2 + q : (q < 4 ? 4 + q : 6 + q); }
__host__ __device__ __forceinline__ int chunk_of_hidden(int i, int half) { return (i - 1) * 4 + half; }

__global__ void __launch_bounds__(kHThreads, 1) heads_fwd_kernel(const HeadsArgs A, const __grid_constant__ CUtensorMap tm_xc,
                                                                 const __grid_constant__ CUtensorMap tm_feats) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t s_afull[kAL], s_alempty[kAL], s_ahfull[kAH], s_ahempty[kAH], s_wfull[kWN], s_wempty[kWN];
  __shared__ uint64_t s_tfull_c[2], s_tfull_s[2], s_dsempty[2], s_actfull;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[5][kHid];    // colour layers 0..3, sigma layer 0
  __shared__ __align__(16) float s_headw[4][kHid];   // colour head rows 0..2, sigma head
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 5 * kHid) s_bias[tid / kHid][tid % kHid] = __ldg((tid < 4 * kHid ? A.bias_c[tid / kHid] : A.bias_s) + tid % kHid);
  if (tid < 4 * kHid) s_headw[tid / kHid][tid % kHid] = __ldg((tid < 3 * kHid ? A.head_c_w : A.head_s_w - 3 * kHid) + tid);
  uint8_t* ahi = smem;                               // kAH x 16 KB
  uint8_t* alo = ahi + kAH * kAtomBytes;             // kAL x 16 KB
  uint8_t* wring = alo + kAL * kAtomBytes;           // kWN x 16 KB
  uint8_t* epi = wring + kWN * kChunkBytes;          // 8 x 2 KB warp transpose buffers + 4 KB head partial sums
  const int kxp = A.kxp, nc = A.kxp + A.kfc, kf = (A.F + 31) >> 5, n0 = nc + kf;
  const int T = (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int NU = n_units(T, n0);

  if (tid == 0) {
    for (int i = 0; i < kAL; ++i) { mbar_init(&s_afull[i], kLoadThreads); mbar_init(&s_alempty[i], 1); }
    for (int i = 0; i < kAH; ++i) { mbar_init(&s_ahfull[i], 1); mbar_init(&s_ahempty[i], 1); }
    for (int i = 0; i < kWN; ++i) { mbar_init(&s_wfull[i], 1); mbar_init(&s_wempty[i], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_tfull_c[b], 1); mbar_init(&s_tfull_s[b], 1); mbar_init(&s_dsempty[b], 256); }
    mbar_init(&s_actfull, 256);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s_tmem;
  // TMEM columns: colour accumulators 0/64, sigma accumulators 128/192, hidden activation hi 256, lo 320
  const uint32_t tm_dc = tm, tm_ds = tm + 128, tm_ahi = tm + 256, tm_alo = tm + 320;

  if (warp == kTmaWarp) {
    // ===== TMA producer: one thread walks the unit sequence and keeps the rings full.  Input atoms are 2-D tensor-map
    // loads (box 32 x 128 fp32, 128-byte swizzle = the K-major operand image; rows/columns outside the tensor arrive as
    // zeros), weight chunks are 16 KB bulk copies.  No LSU work, and as many loads in flight as there are free slots.
    if (lane == 0) {
      tma_prefetch_desc(&tm_xc);
      tma_prefetch_desc(&tm_feats);
      int ia = 0, iw = 0;
      for (int u = 0; u < NU; ++u) {
        const Unit x = decode_unit(u, T, n0);
        if (x.hidden == 0) {
          const int h = ia % kAH;
          mbar_wait(&s_ahempty[h], ((ia / kAH) & 1) ^ 1);
          const int row0 = (blockIdx.x + x.tile * gridDim.x) * 128;
          mbar_expect_tx(&s_ahfull[h], kAtomBytes);
          if (x.q < kxp) tma_load_2d(ahi + h * kAtomBytes, &tm_xc, 32 * x.q, row0, &s_ahfull[h]);
          else tma_load_2d(ahi + h * kAtomBytes, &tm_feats, 32 * (x.q - (x.q < nc ? kxp : nc)), row0, &s_ahfull[h]);
          ++ia;
        }
        const int n_chunks = x.hidden ? 2 : 1;
        for (int half = 0; half < n_chunks; ++half) {
          const int w = iw % kWN;
          mbar_wait(&s_wempty[w], ((iw / kWN) & 1) ^ 1);
          mbar_expect_tx(&s_wfull[w], kChunkBytes);
          bulk_copy_g2s(wring + w * kChunkBytes, A.wimg + (size_t)(x.hidden ? chunk_of_hidden(x.hidden, half) : chunk_of_item(x.q)) * kChunkBytes,
                        kChunkBytes, &s_wfull[w]);
          ++iw;
        }
      }
    }
  } else if (warp < kLoadWarps) {
    // ===== lo-pass warps: for every input atom, lo = x - trunc(x) from the landed hi image into a lo slot =====
    const int n_atoms = T * n0;
    for (int ca = 0; ca < n_atoms; ++ca) {
      const int h = ca % kAH, l = ca % kAL;
      mbar_wait(&s_ahfull[h], (ca / kAH) & 1);
      mbar_wait(&s_alempty[l], ((ca / kAL) & 1) ^ 1);
      make_lo_atom<kLoadThreads>(ahi + h * kAtomBytes, alo + l * kAtomBytes, tid, false, nullptr);
      fence_async_smem();
      mbar_arrive(&s_afull[l]);
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =====
    const uint32_t idesc = instr_desc(128, kHid, false, false);
    int ca = 0, cw = 0, cact = 0;
    long long d_ds = 0, d_af = 0, d_wf = 0, d_act = 0, d_iss0 = 0, d_issh = 0;
    const long long d_start = HCLK();
    for (int u = 0; u < NU; ++u) {
      const Unit x = decode_unit(u, T, n0);
      const int b = x.tile & 1;
      if (x.hidden == 0) {
        const int h = ca % kAH, l = ca % kAL, w = cw % kWN;
        const bool colour = x.q < nc;
        { HT0(); if (x.q == nc) mbar_wait(&s_dsempty[b], ((x.tile >> 1) & 1) ^ 1); HT1(d_ds); }  // sigma epilogue of tile-2 has drained Ds[b]
        { HT0(); mbar_wait(&s_afull[l], (ca / kAL) & 1); HT1(d_af); }
        { HT0(); mbar_wait(&s_wfull[w], (cw / kWN) & 1); HT1(d_wf); }
        tc_fence_after();
        const long long _ti = HCLK();
        if (elect_one()) {
          const uint32_t a_h = smem_u32(ahi + h * kAtomBytes), a_l = smem_u32(alo + l * kAtomBytes);
          const uint32_t w_h = smem_u32(wring + w * kChunkBytes), w_l = w_h + kChunkBytes / 2;
          const uint32_t d = (colour ? tm_dc : tm_ds) + b * kHid;
          const bool first = (x.q == 0) || (x.q == nc);
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t aa = (pass == 1) ? a_l : a_h;
            const uint32_t ww = (pass == 2) ? w_l : w_h;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_tf32(d, desc_kmajor(aa, kk), desc_kmajor(ww, kk), idesc, !(first && pass == 0 && kk == 0));
            if (pass == 1) mma_commit(&s_alempty[l]);
          }
          mma_commit(&s_ahempty[h]);
          mma_commit(&s_wempty[w]);
          if (x.q == nc - 1) mma_commit(&s_tfull_c[b]);
          if (x.q == n0 - 1) mma_commit(&s_tfull_s[b]);
        }
        __syncwarp();
        d_iss0 += HCLK() - _ti;
        ++ca;
        ++cw;
      } else {
        const int w0 = cw % kWN, w1 = (cw + 1) % kWN;
        { HT0(); mbar_wait(&s_actfull, cact & 1); HT1(d_act); }
        { HT0(); mbar_wait(&s_wfull[w0], (cw / kWN) & 1);
        mbar_wait(&s_wfull[w1], ((cw + 1) / kWN) & 1); HT1(d_wf); }
        tc_fence_after();
        const long long _ti = HCLK();
        if (elect_one()) {
          const uint32_t d = tm_dc + b * kHid;
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t aa = (pass == 1) ? tm_alo : tm_ahi;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t ww = smem_u32(wring + (half ? w1 : w0) * kChunkBytes) + (pass == 2 ? kChunkBytes / 2 : 0);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma_tf32_ts(d, aa + half * 32 + kk * 8, desc_kmajor(ww, kk), idesc, (pass | half | kk) != 0);
            }
          }
          mma_commit(&s_wempty[w0]);
          mma_commit(&s_wempty[w1]);
          mma_commit(&s_tfull_c[b]);
        }
        __syncwarp();
        d_issh += HCLK() - _ti;
        cw += 2;
        ++cact;
      }
    }
    if (g_hdbg && lane == 0 && blockIdx.x == 0) { g_hdbg[8] = d_ds; g_hdbg[9] = d_af; g_hdbg[10] = d_wf; g_hdbg[11] = d_act; g_hdbg[12] = d_iss0; g_hdbg[13] = d_issh; g_hdbg[14] = HCLK() - d_start; }
  } else {
    // ===== epilogue warps (8): warp w owns TMEM lanes 32*(w%4).. (tile rows likewise); the two warps of a lane quarter
    // split the 64 output columns (group g = columns 32g..32g+31).  Every warp serves both chains: the four colour
    // stages of a tile (the critical path: each stage turns one layer's accumulator into the next layer's TMEM operand)
    // and the tile's sigma stage, slotted in after colour layer 0.
    const int ew = warp - kLoadWarps, q4 = ew & 3, g = ew >> 2;
    const int r = q4 * 32 + lane;
    const int c0 = 32 * g;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    uint8_t* wbuf = epi + ew * 2048;              // 32 rows x 64 B (16 columns), 16-byte chunks XOR (row>>1)&3
    float* s_part = reinterpret_cast<float*>(epi + 8 * 2048);   // [2 parity][128 rows][4]: head partial sums of group 1
    // store the warp's 32 rows x 32 columns (v = this thread's row) to dst[M, 64] columns c0..c0+31: two passes of 16
    // columns through the warp-private buffer, written back as 64-byte row segments (8 rows per instruction)
    auto store_rows = [&](float* dst, long long row0, const float v[32]) {
      const int rr0 = lane >> 2, cc = lane & 3;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          *reinterpret_cast<float4*>(wbuf + lane * 64 + ((qq ^ ((lane >> 1) & 3)) << 4)) =
              make_float4(v[16 * p + 4 * qq], v[16 * p + 4 * qq + 1], v[16 * p + 4 * qq + 2], v[16 * p + 4 * qq + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = rr0 + 8 * i;
          const long long grow = row0 + q4 * 32 + rr;
          if (grow < A.M)
            *reinterpret_cast<float4*>(dst + grow * kHid + c0 + 16 * p + 4 * cc) =
                *reinterpret_cast<const float4*>(wbuf + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4));
        }
        __syncwarp();
      }
    };
    // accumulator chunk -> relu(acc + bias) in registers
    auto load_act = [&](uint32_t taddr, const float* bias, float v[32]) {
      tmem_ld32(taddr + lane_off + c0, v);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + i);
        v[i] = fmaxf(v[i] + bb.x, 0.f); v[i + 1] = fmaxf(v[i + 1] + bb.y, 0.f);
        v[i + 2] = fmaxf(v[i + 2] + bb.z, 0.f); v[i + 3] = fmaxf(v[i + 3] + bb.w, 0.f);
      }
    };
    int phc[2] = {0, 0}, phs[2] = {0, 0};
    int n_sync = 0;   // named-barrier uses so far (parity of the partial-sum buffer)
    for (int tl = 0; tl < T; ++tl) {
      const int b = tl & 1;
      const long long row0 = (long long)(blockIdx.x + tl * gridDim.x) * 128;
      const long long row = row0 + r;
      for (int layer = 0; layer < 4; ++layer) {
        mbar_wait(&s_tfull_c[b], phc[b]);
        phc[b] ^= 1;
        tc_fence_after();
        float v[32];
        load_act(tm_dc + b * kHid, s_bias[layer], v);
        if (layer < 3) {
          tmem_st32(tm_ahi + lane_off + c0, v);   // hi operand = the fp32 value itself (the tensor core truncates)
          float lo[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) lo[i] = v[i] - __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
          tmem_st32(tm_alo + lane_off + c0, lo);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&s_actfull);                // the next layer's MMAs may start; the HBM copy follows off the chain
          if (A.h[layer]) store_rows(A.h[layer], row0, v);
        } else {
          if (A.h[layer]) store_rows(A.h[layer], row0, v);
          float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int o = 0; o < 3; ++o)
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[o] = __fmaf_rn(v[i], s_headw[o][c0 + i], acc[o]);
          float* part = s_part + ((n_sync & 1) * 128 + r) * 4;
          if (g == 1) { part[0] = acc[0]; part[1] = acc[1]; part[2] = acc[2]; }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          ++n_sync;
          if (g == 0 && row < A.M) {
#pragma unroll
            for (int o = 0; o < 3; ++o) A.rgb[row * 3 + o] = 1.f / (1.f + expf(-(acc[o] + part[o] + __ldg(A.head_c_b + o))));  // sigmoid
          }
        }
        if (layer == 0) {
          // the tile's density head (its accumulator completes right after colour layer 0's)
          mbar_wait(&s_tfull_s[b], phs[b]);
          phs[b] ^= 1;
          tc_fence_after();
          float u[32];
          load_act(tm_ds + b * kHid, s_bias[4], u);
          tc_fence_before();
          mbar_arrive(&s_dsempty[b]);
          if (A.hs) store_rows(A.hs, row0, u);
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) acc = __fmaf_rn(u[i], s_headw[3][c0 + i], acc);
          float* part = s_part + ((n_sync & 1) * 128 + r) * 4;
          if (g == 1) part[0] = acc;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          ++n_sync;
          if (g == 0 && row < A.M) A.sigma[row] = expf(acc + part[0] + __ldg(A.head_s_b) - 1.f);  // truncated_exp(x - 1.)
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tm, 512);
}

// ---- weight packing: nn.Linear weights -> 16 KB chunks (hi image, lo image; 64 rows x 128 B, 16-byte chunks XOR row%8) ----
struct PackArgs {
  const float* wc0; int K0;   // [64, K0]
  const float* ws0; int F;    // [64, F]
  const float* w123[3];       // [64, 64]
  uint8_t* img;
  int kxp, kfc, kf, xc_cols;  // layer-0 items as in HeadsArgs; xc_cols = columns of W_c0 covered by the xc atoms
};
// layer-0 item q -> the weight matrix, its leading dimension and the column range [k0, k1) the item's 32-wide atom covers
__device__ __forceinline__ void item_weights(const PackArgs& P, int q, const float*& W, int& ld, int& k0, int& k1) {
  if (q < P.kxp) { W = P.wc0; ld = P.K0; k0 = 32 * q; k1 = min(k0 + 32, P.xc_cols); }
  else if (q < P.kxp + P.kfc) { W = P.wc0; ld = P.K0; k0 = P.xc_cols + 32 * (q - P.kxp); k1 = min(k0 + 32, P.K0); }
  else { W = P.ws0; ld = P.F; k0 = 32 * (q - P.kxp - P.kfc); k1 = min(k0 + 32, P.F); }
}
__global__ void __launch_bounds__(256) pack_heads_kernel(const PackArgs P) {
  const int n_chunks = 6 + P.kxp + P.kfc + P.kf;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (chunk, row, 16-byte column chunk)
  if (t >= n_chunks * 64 * 8) return;
  const int chunk = t / 512, r = (t % 512) / 8, c = t % 8;
  // which matrix / k-atom does this chunk hold?
  const float* W = nullptr;
  int ld = 0, k0 = 0, k1 = 0;                       // the chunk holds W[:, k0:k1) (zero-padded to 32 columns)
  const int blk = chunk / 4, pos = chunk % 4;       // blocks of [Wi a, Wi b, item, item] for the first two groups
  if (chunk < 8) {
    if (pos < 2) { W = P.w123[blk]; ld = kHid; k0 = 32 * pos; k1 = k0 + 32; }
    else item_weights(P, blk * 2 + (pos - 2), W, ld, k0, k1);
  } else if (chunk < 10) {
    W = P.w123[2]; ld = kHid; k0 = 32 * (chunk - 8); k1 = k0 + 32;
  } else {
    item_weights(P, chunk - 6, W, ld, k0, k1);
  }
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + c * 4 + i;
    v[i] = (k < k1) ? __ldg(W + (long long)r * ld + k) : 0.f;
  }
  uint8_t* base = P.img + (size_t)chunk * kChunkBytes;
  const int off = r * 128 + ((c ^ (r & 7)) << 4);
  *reinterpret_cast<float4*>(base + off) = make_float4(v[0], v[1], v[2], v[3]);
  float4 lo;
  lo.x = v[0] - __uint_as_float(__float_as_uint(v[0]) & 0xFFFFE000u);
  lo.y = v[1] - __uint_as_float(__float_as_uint(v[1]) & 0xFFFFE000u);
  lo.z = v[2] - __uint_as_float(__float_as_uint(v[2]) & 0xFFFFE000u);
  lo.w = v[3] - __uint_as_float(__float_as_uint(v[3]) & 0xFFFFE000u);
  *reinterpret_cast<float4*>(base + kChunkBytes / 2 + off) = lo;
}

}  // namespace
}  // namespace tnf

extern "C" int tnf_debug_heads_timing(long long* buf) {  // diagnostics only
  return cudaMemcpyToSymbol(tnf::g_hdbg, &buf, sizeof(buf)) == cudaSuccess ? 0 : -2;
}

namespace tnf {
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// row-major fp32 [rows, cols] with leading dimension ld -> tensor map with a 32 x 128 box in the 128-byte swizzle
int make_atom_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  static EncodeTiledFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return reinterpret_cast<EncodeTiledFn>(fn);
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
}  // namespace tnf

extern "C" int64_t tnf_heads_workspace_bytes(int32_t feat_dim, int32_t k0) {
  const int kx = (k0 + 31) / 32, kf = (feat_dim + 31) / 32;
  return (int64_t)(6 + kx + 1 + 2 * kf) * tnf::kChunkBytes;   // enough for either layout of the colour input
}

extern "C" int tnf_heads_fwd(const float* feats, int64_t ld_feats, int32_t feat_dim, const float* xc, int64_t ld_xc, int32_t k0,
                             int32_t xc_cols, const float* const* color_w, const float* const* color_b, const float* const* sigma_w,
                             const float* const* sigma_b, float* const* h_out, float* hs_out, float* rgb, float* sigma,
                             int64_t m, void* workspace, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(m >= 0, "negative m");
  if (m == 0) return TNF_OK;
  TNF_REQUIRE(feats && xc && color_w && color_b && sigma_w && sigma_b && rgb && sigma && workspace, "null pointer");
  TNF_REQUIRE(feat_dim >= 1 && k0 >= 1, "bad widths");
  TNF_REQUIRE(xc_cols >= 1 && (xc_cols == k0 || xc_cols + feat_dim == k0),
              "xc_cols must be k0 (xc holds the whole colour-input row) or k0 - feat_dim (its trailing columns are the feature row)");
  const int kxp = (xc_cols + 31) / 32, kf = (feat_dim + 31) / 32, kfc = (xc_cols == k0) ? 0 : kf;
  TNF_REQUIRE(kxp + kfc + kf >= 4 && kxp + kfc + kf <= 12, "unsupported input widths (k0=%d, feat_dim=%d)", k0, feat_dim);
  TNF_REQUIRE(ld_feats % 4 == 0 && ld_xc % 4 == 0 && (reinterpret_cast<uintptr_t>(feats) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(xc) & 15u) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0,
              "feats/xc/workspace must be 16-byte aligned with leading dimensions multiple of 4");
  for (int i = 0; i < 5; ++i) TNF_REQUIRE(color_w[i] && color_b[i], "null colour layer %d", i);
  for (int i = 0; i < 2; ++i) TNF_REQUIRE(sigma_w[i] && sigma_b[i], "null sigma layer %d", i);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PackArgs P{};
  P.wc0 = color_w[0]; P.K0 = k0; P.ws0 = sigma_w[0]; P.F = feat_dim;
  for (int i = 0; i < 3; ++i) P.w123[i] = color_w[1 + i];
  P.img = static_cast<uint8_t*>(workspace); P.kxp = kxp; P.kfc = kfc; P.kf = kf; P.xc_cols = xc_cols;
  const int pack_threads = (6 + kxp + kfc + kf) * 512;
  pack_heads_kernel<<<(pack_threads + 255) / 256, 256, 0, st>>>(P);
  TNF_LAUNCH_CHECK("pack_heads_kernel");
  HeadsArgs A{};
  A.feats = feats; A.ld_feats = ld_feats; A.F = feat_dim; A.xc = xc; A.ld_xc = ld_xc; A.K0 = k0; A.kxp = kxp; A.kfc = kfc;
  A.wimg = static_cast<const uint8_t*>(workspace);
  for (int i = 0; i < 4; ++i) { A.bias_c[i] = color_b[i]; A.h[i] = h_out ? h_out[i] : nullptr; }
  A.bias_s = sigma_b[0];
  A.head_c_w = color_w[4]; A.head_c_b = color_b[4]; A.head_s_w = sigma_w[1]; A.head_s_b = sigma_b[1];
  A.hs = hs_out; A.rgb = rgb; A.sigma = sigma; A.M = m;
  A.n_tiles = (int)ceil_div(m, 128);
  const size_t smem = (size_t)(kAH + kAL) * kAtomBytes + (size_t)kWN * kChunkBytes + 8 * 2048 + 4096 + 1024;
  static PerDeviceOnce configured{};
  if (configured.pending()) {
    TNF_CUDA(cudaFuncSetAttribute(heads_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.mark();
  }
  TNF_REQUIRE(m < (1LL << 31) - 256, "too many rows for the tensor-map coordinates");
  CUtensorMap tm_xc, tm_feats;
  int rc = make_atom_map(&tm_xc, xc, m, xc_cols, ld_xc);
  if (rc != TNF_OK) return rc;
  rc = make_atom_map(&tm_feats, feats, m, feat_dim, ld_feats);
  if (rc != TNF_OK) return rc;
  const int grid = A.n_tiles < sm_count() ? A.n_tiles : sm_count();
  heads_fwd_kernel<<<grid, kHThreads, smem, st>>>(A, tm_xc, tm_feats);
  TNF_LAUNCH_CHECK("heads_fwd_kernel");
  return TNF_OK;
}
