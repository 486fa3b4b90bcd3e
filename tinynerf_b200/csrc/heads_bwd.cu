// Backward of both decoder heads with respect to their INPUTS (the "dgrad" chain) in one persistent tcgen05 kernel:
//   dh2 = (h2 > 0) * (dh3 W_c3)      dh1 = (h1 > 0) * (dh2 W_c2)      dh0 = (h0 > 0) * (dh1 W_c1)
//   dfeat = dh0 W_c0[:, feature columns] + dhs W_s0
// i.e. the autograd backward of the ReLU/Linear stacks of VanillaColorDecoder / VanillaOpacityDecoder (src/models.py:7-28,
// 70-89) from the gradients at the last hidden layers (dh3, dhs: produced by tnf_head_bwd) down to the feature rows that
// feed both heads.  The per-layer path needs five tnf_linear_bwd_data launches plus an add and streams every dh through
// HBM twice; here a 128-sample tile walks the whole chain on chip, exactly like tnf_heads_fwd walks the forward:
//   * dh3 / dhs tiles arrive by 2-D TMA in the swizzled K-major operand image (raw fp32 = tf32 hi operand), loader warps add lo;
//   * the transposed weights come as pre-packed chunks (hi + lo, B operand K-major with rows = input features) by bulk copy;
//   * every intermediate dh_i is masked by the epilogue warps (mask rows read straight from the saved activations), stored
//     for the weight-gradient kernels and written back to TENSOR MEMORY as the next layer's A operand (TS-mode MMAs);
//   * the last layer of the colour chain (N = feature width) and the density branch accumulate into the same accumulator,
//     so the sum over the two consumers of the features costs nothing;
//   * static software pipeline: layer 0 of tile t+1 is interleaved with the hidden layers of tile t (two accumulators).
// 3xTF32 (hi*hi + lo*hi + hi*lo) as everywhere in mlp.cu.
#include <cuda.h>
#include "common.cuh"
#include "tc.cuh"

namespace tnf {
namespace {

constexpr int kBLoadWarps = 8;
constexpr int kBLoadThreads = kBLoadWarps * 32;
constexpr int kBMmaWarp = kBLoadWarps + 8;
constexpr int kBTmaWarp = kBLoadWarps + 9;
constexpr int kBThreads = (kBLoadWarps + 10) * 32;   // lo-pass warps | 8 epilogue | MMA issuer | TMA producer
constexpr int kBAH = 4, kBAL = 2, kBWN = 4;
constexpr int kBHid = 64;
constexpr int kMaxFeat = 128;
constexpr int kNChunks = 10;                         // packed image: 10 chunks at `slot` stride
// weight ring slot = [hi rows x 128 B][lo rows x 128 B] with rows = max(64, F): 24 KB for the K-Planes feature width

struct HBArgs {
  const uint8_t* wimg;
  const float* mask[3];      // h2, h1, h0  [M,64]
  float* dh_out[3];          // dh2, dh1, dh0 [M,64] (optional)
  float* dfeat; long long ld_dfeat; int F;
  long long M; int n_tiles;
  int slot;                  // bytes per weight ring slot / image chunk
  int wn;                    // weight ring slots (<= kBWN)
};

// Units of a CTA with T tiles (kind: 0 = layer-0 item q (dh3 atom) of `tile`, 1/2 = hidden layer, 3 = final layer (N = F),
// 4 = density item q (dhs atom)):
//   prologue: L0(0) q=0,1 ; tile t < T-1: H1 | L0(t+1) q=0 | H2 | L0(t+1) q=1 | F | S q=0 | S q=1 ; last tile: H1 H2 F S0 S1
struct BUnit { int kind, tile, q; };
__device__ __forceinline__ int b_units(int T) { return T <= 0 ? 0 : 2 + (T - 1) * 7 + 5; }
__device__ __forceinline__ BUnit b_decode(int u, int T) {
  BUnit x;
  if (u < 2) { x.kind = 0; x.tile = 0; x.q = u; return x; }
  int v = u - 2;
  int t = v / 7, r = v - t * 7;
  if (t >= T - 1) {
    t = T - 1; r = v - t * 7;                    // H1 H2 F S0 S1
    x.tile = t;
    x.kind = r < 2 ? r + 1 : (r == 2 ? 3 : 4);
    x.q = r < 3 ? 0 : r - 3;
    return x;
  }
  x.tile = t; x.q = 0;
  switch (r) {
    case 0: x.kind = 1; break;
    case 1: x.kind = 0; x.tile = t + 1; x.q = 0; break;
    case 2: x.kind = 2; break;
    case 3: x.kind = 0; x.tile = t + 1; x.q = 1; break;
    case 4: x.kind = 3; break;
    default: x.kind = 4; x.q = r - 5; break;
  }
  return x;
}
// chunk indices in the packed image: W3^T atoms 0/1 (layer 0), W2^T a/b (hidden 1), W1^T a/b (hidden 2), W0f^T a/b (final),
// Ws0^T a/b (density)
__host__ __device__ __forceinline__ int b_chunk(int kind, int q_or_half) {
  return kind == 0 ? q_or_half : (kind == 1 ? 2 + q_or_half : (kind == 2 ? 4 + q_or_half : (kind == 3 ? 6 + q_or_half : 8 + q_or_half)));
}

__global__ void __launch_bounds__(kBThreads, 1) heads_bwd_data_kernel(const HBArgs A, const __grid_constant__ CUtensorMap tm_dh3,
                                                                      const __grid_constant__ CUtensorMap tm_dhs) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t s_afull[kBAL], s_alempty[kBAL], s_ahfull[kBAH], s_ahempty[kBAH], s_wfull[kBWN], s_wempty[kBWN];
  __shared__ uint64_t s_tfull[2], s_actfull;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* ahi = smem;
  uint8_t* alo = ahi + kBAH * kAtomBytes;
  uint8_t* wring = alo + kBAL * kAtomBytes;
  const int kSlotBytes = A.slot, WN = A.wn;
  uint8_t* epi = wring + WN * kSlotBytes;            // 8 x 2 KB warp transpose buffers
  const int F = A.F;
  const int T = (A.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int NU = b_units(T);

  if (tid == 0) {
    for (int i = 0; i < kBAL; ++i) { mbar_init(&s_afull[i], kBLoadThreads); mbar_init(&s_alempty[i], 1); }
    for (int i = 0; i < kBAH; ++i) { mbar_init(&s_ahfull[i], 1); mbar_init(&s_ahempty[i], 1); }
    for (int i = 0; i < kBWN; ++i) { mbar_init(&s_wfull[i], 1); mbar_init(&s_wempty[i], 1); }
    mbar_init(&s_tfull[0], 1); mbar_init(&s_tfull[1], 1);
    mbar_init(&s_actfull, 256);
    fence_mbar_init();
  }
  if (warp == kBMmaWarp) tmem_alloc(&s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s_tmem;
  const uint32_t tm_d = tm, tm_ahi = tm + 256, tm_alo = tm + 320;   // accumulators at 0 / 128, activation hi / lo

  if (warp == kBTmaWarp) {
    // ===== TMA producer =====
    if (lane == 0) {
      tma_prefetch_desc(&tm_dh3);
      tma_prefetch_desc(&tm_dhs);
      int ia = 0, iw = 0;
      for (int u = 0; u < NU; ++u) {
        const BUnit x = b_decode(u, T);
        const bool item = (x.kind == 0 || x.kind == 4);
        if (item) {
          const int h = ia % kBAH;
          mbar_wait(&s_ahempty[h], ((ia / kBAH) & 1) ^ 1);
          const int row0 = (blockIdx.x + x.tile * gridDim.x) * 128;
          mbar_expect_tx(&s_ahfull[h], kAtomBytes);
          tma_load_2d(ahi + h * kAtomBytes, x.kind == 0 ? &tm_dh3 : &tm_dhs, 32 * x.q, row0, &s_ahfull[h]);
          ++ia;
        }
        const int n_chunks = item ? 1 : 2;
        for (int half = 0; half < n_chunks; ++half) {
          const int w = iw % WN;
          const int chunk = b_chunk(x.kind, item ? x.q : half);
          const uint32_t bytes = (uint32_t)((x.kind >= 3 ? F : kBHid) * 256);
          mbar_wait(&s_wempty[w], ((iw / WN) & 1) ^ 1);
          mbar_expect_tx(&s_wfull[w], bytes);
          bulk_copy_g2s(wring + w * kSlotBytes, A.wimg + (size_t)chunk * kSlotBytes, bytes, &s_wfull[w]);
          ++iw;
        }
      }
    }
  } else if (warp < kBLoadWarps) {
    // ===== lo-pass warps =====
    const int n_atoms = T * 4;
    for (int ca = 0; ca < n_atoms; ++ca) {
      const int h = ca % kBAH, l = ca % kBAL;
      mbar_wait(&s_ahfull[h], (ca / kBAH) & 1);
      mbar_wait(&s_alempty[l], ((ca / kBAL) & 1) ^ 1);
      make_lo_atom<kBLoadThreads>(ahi + h * kAtomBytes, alo + l * kAtomBytes, tid, false, nullptr);
      fence_async_smem();
      mbar_arrive(&s_afull[l]);
    }
  } else if (warp == kBMmaWarp) {
    // ===== MMA issuer =====
    const uint32_t idesc64 = instr_desc(128, kBHid, false, false), idescF = instr_desc(128, F, false, false);
    int ca = 0, cw = 0, cact = 0;
    for (int u = 0; u < NU; ++u) {
      const BUnit x = b_decode(u, T);
      const int b = x.tile & 1;
      const uint32_t d = tm_d + b * 128;
      if (x.kind == 0 || x.kind == 4) {
        const int h = ca % kBAH, l = ca % kBAL, w = cw % WN;
        mbar_wait(&s_afull[l], (ca / kBAL) & 1);
        mbar_wait(&s_wfull[w], (cw / WN) & 1);
        tc_fence_after();
        if (elect_one()) {
          const int rows = x.kind == 0 ? kBHid : F;
          const uint32_t a_h = smem_u32(ahi + h * kAtomBytes), a_l = smem_u32(alo + l * kAtomBytes);
          const uint32_t w_h = smem_u32(wring + w * kSlotBytes), w_l = w_h + rows * 128;
          const uint32_t idesc = x.kind == 0 ? idesc64 : idescF;
          const bool first = (x.kind == 0 && x.q == 0);   // density items accumulate onto the final colour layer
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
          if (x.q == 1) mma_commit(&s_tfull[b]);   // layer 0 complete (kind 0) / whole tile complete (kind 4)
        }
        __syncwarp();
        ++ca;
        ++cw;
      } else {
        const int w0 = cw % WN, w1 = (cw + 1) % WN;
        mbar_wait(&s_actfull, cact & 1);
        mbar_wait(&s_wfull[w0], (cw / WN) & 1);
        mbar_wait(&s_wfull[w1], ((cw + 1) / WN) & 1);
        tc_fence_after();
        if (elect_one()) {
          const int rows = x.kind == 3 ? F : kBHid;
          const uint32_t idesc = x.kind == 3 ? idescF : idesc64;
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t aa = (pass == 1) ? tm_alo : tm_ahi;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t ww = smem_u32(wring + (half ? w1 : w0) * kSlotBytes) + (pass == 2 ? rows * 128 : 0);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma_tf32_ts(d, aa + half * 32 + kk * 8, desc_kmajor(ww, kk), idesc, (pass | half | kk) != 0);
            }
          }
          mma_commit(&s_wempty[w0]);
          mma_commit(&s_wempty[w1]);
          if (x.kind != 3) mma_commit(&s_tfull[b]);   // the final layer's accumulator completes with the density items
        }
        __syncwarp();
        cw += 2;
        ++cact;
      }
    }
  } else {
    // ===== epilogue warps (8): quarter q4 = TMEM lanes / tile rows 32 q4.., group g = 32-column chunks g, g+2, ... =====
    const int ew = warp - kBLoadWarps, q4 = ew & 3, g = ew >> 2;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    uint8_t* wbuf = epi + ew * 2048;
    auto store_rows = [&](float* dst, long long ld, long long row0, int c0, const float v[32]) {
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
            *reinterpret_cast<float4*>(dst + grow * ld + c0 + 16 * p + 4 * cc) =
                *reinterpret_cast<const float4*>(wbuf + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4));
        }
        __syncwarp();
      }
    };
    int ph[2] = {0, 0};
    for (int tl = 0; tl < T; ++tl) {
      const int b = tl & 1;
      const long long row0 = (long long)(blockIdx.x + tl * gridDim.x) * 128;
      const long long row = row0 + q4 * 32 + lane;
      const int c0 = 32 * g;
      for (int stage = 0; stage < 3; ++stage) {
        // mask row of this thread (the saved activation of the layer whose input gradient this is), issued before the wait
        float4 mk[8];
        const float* mrow = A.mask[stage] + (row < A.M ? row : 0) * kBHid + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i) mk[i] = __ldg(reinterpret_cast<const float4*>(mrow) + i);
        mbar_wait(&s_tfull[b], ph[b]);
        ph[b] ^= 1;
        tc_fence_after();
        float v[32];
        tmem_ld32(tm_d + b * 128 + lane_off + c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[4 * i] = mk[i].x > 0.f ? v[4 * i] : 0.f;         v[4 * i + 1] = mk[i].y > 0.f ? v[4 * i + 1] : 0.f;
          v[4 * i + 2] = mk[i].z > 0.f ? v[4 * i + 2] : 0.f; v[4 * i + 3] = mk[i].w > 0.f ? v[4 * i + 3] : 0.f;
        }
        tmem_st32(tm_ahi + lane_off + c0, v);
        float lo[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) lo[i] = v[i] - __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
        tmem_st32(tm_alo + lane_off + c0, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&s_actfull);
        if (A.dh_out[stage]) store_rows(A.dh_out[stage], kBHid, row0, c0, v);
      }
      // final stage: dfeat (F columns) = colour chain + density branch, already summed in the accumulator
      mbar_wait(&s_tfull[b], ph[b]);
      ph[b] ^= 1;
      tc_fence_after();
      for (int cf = c0; cf < F; cf += 64) {
        float v[32];
        tmem_ld32(tm_d + b * 128 + lane_off + cf, v);
        store_rows(A.dfeat, A.ld_dfeat, row0, cf, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBMmaWarp) tmem_dealloc(tm, 512);
}

// ---- weight packing: chunk = B operand of one k-atom: rows = input features (GEMM N), 32 columns = output features (GEMM K),
// value W[nout][kin]; hi image (rows x 128 B, 16-byte chunks XOR row%8) followed by the lo image ----
struct BPackArgs {
  const float* w3; const float* w2; const float* w1;   // [64,64]
  const float* w0; int K0; int feat_col0;               // [64,K0]; the features are columns feat_col0 .. feat_col0+F-1
  const float* ws0;                                     // [64,F]
  int F;
  uint8_t* img;
  int slot;
};
__global__ void __launch_bounds__(256) pack_heads_bwd_kernel(const BPackArgs P) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (chunk, row, 16-byte column chunk)
  if (t >= kNChunks * kMaxFeat * 8) return;
  const int chunk = t / (kMaxFeat * 8), r = (t % (kMaxFeat * 8)) / 8, c = t % 8;
  const int layer = chunk >> 1, atom = chunk & 1;        // 0: W3, 1: W2, 2: W1, 3: W0 feature part, 4: Ws0
  const int rows = layer >= 3 ? P.F : kBHid;
  if (r >= rows) return;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nout = atom * 32 + c * 4 + i;              // GEMM K index
    float w;
    if (layer == 0) w = __ldg(P.w3 + nout * kBHid + r);
    else if (layer == 1) w = __ldg(P.w2 + nout * kBHid + r);
    else if (layer == 2) w = __ldg(P.w1 + nout * kBHid + r);
    else if (layer == 3) w = __ldg(P.w0 + (long long)nout * P.K0 + P.feat_col0 + r);
    else w = __ldg(P.ws0 + (long long)nout * P.F + r);
    v[i] = w;
  }
  uint8_t* base = P.img + (size_t)chunk * P.slot;
  const int off = r * 128 + ((c ^ (r & 7)) << 4);
  *reinterpret_cast<float4*>(base + off) = make_float4(v[0], v[1], v[2], v[3]);
  float4 lo;
  lo.x = v[0] - __uint_as_float(__float_as_uint(v[0]) & 0xFFFFE000u);
  lo.y = v[1] - __uint_as_float(__float_as_uint(v[1]) & 0xFFFFE000u);
  lo.z = v[2] - __uint_as_float(__float_as_uint(v[2]) & 0xFFFFE000u);
  lo.w = v[3] - __uint_as_float(__float_as_uint(v[3]) & 0xFFFFE000u);
  *reinterpret_cast<float4*>(base + rows * 128 + off) = lo;
}

typedef CUresult (*EncodeTiledFnB)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_atom_map_b(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  static EncodeTiledFnB encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return reinterpret_cast<EncodeTiledFnB>(fn);
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

extern "C" int64_t tnf_heads_bwd_workspace_bytes(int32_t feat_dim) {
  return (int64_t)tnf::kNChunks * (feat_dim > tnf::kBHid ? feat_dim : tnf::kBHid) * 256;
}

extern "C" int tnf_heads_bwd_data(const float* dh3, const float* dhs, const float* const* masks, const float* const* color_w,
                                  int32_t k0, int32_t feat_col0, const float* sigma_w0, int32_t feat_dim, float* const* dh_out,
                                  float* dfeat, int64_t ld_dfeat, int64_t m, void* workspace, void* stream) {
  using namespace tnf;
  TNF_REQUIRE(m >= 0, "negative m");
  if (m == 0) return TNF_OK;
  TNF_REQUIRE(dh3 && dhs && masks && color_w && sigma_w0 && dfeat && workspace, "null pointer");
  TNF_REQUIRE(feat_dim >= 32 && feat_dim <= kMaxFeat && feat_dim % 32 == 0, "feat_dim must be a multiple of 32 in [32,128]");
  TNF_REQUIRE(feat_col0 >= 0 && feat_col0 + feat_dim <= k0, "feature columns outside the colour input");
  TNF_REQUIRE(ld_dfeat % 4 == 0 && ld_dfeat >= feat_dim, "ld_dfeat must be a multiple of 4 >= feat_dim");
  for (int i = 0; i < 3; ++i) TNF_REQUIRE(masks[i] && (reinterpret_cast<uintptr_t>(masks[i]) & 15u) == 0, "mask %d null/misaligned", i);
  for (int i = 0; i < 4; ++i) TNF_REQUIRE(color_w[i], "null colour weight %d", i);
  TNF_REQUIRE(((reinterpret_cast<uintptr_t>(dh3) | reinterpret_cast<uintptr_t>(dhs) | reinterpret_cast<uintptr_t>(dfeat) |
                reinterpret_cast<uintptr_t>(workspace)) & 15u) == 0, "dh3/dhs/dfeat/workspace must be 16-byte aligned");
  TNF_REQUIRE(m < (1LL << 31) - 256, "too many rows for the tensor-map coordinates");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BPackArgs P{};
  P.w3 = color_w[3]; P.w2 = color_w[2]; P.w1 = color_w[1]; P.w0 = color_w[0]; P.K0 = k0; P.feat_col0 = feat_col0;
  P.ws0 = sigma_w0; P.F = feat_dim; P.img = static_cast<uint8_t*>(workspace);
  const int slot = (feat_dim > kBHid ? feat_dim : kBHid) * 256;
  P.slot = slot;
  pack_heads_bwd_kernel<<<(kNChunks * kMaxFeat * 8 + 255) / 256, 256, 0, st>>>(P);
  TNF_LAUNCH_CHECK("pack_heads_bwd_kernel");
  HBArgs A{};
  A.wimg = static_cast<const uint8_t*>(workspace);
  for (int i = 0; i < 3; ++i) {
    A.mask[i] = masks[i];
    A.dh_out[i] = dh_out ? dh_out[i] : nullptr;
    TNF_REQUIRE(!A.dh_out[i] || (reinterpret_cast<uintptr_t>(A.dh_out[i]) & 15u) == 0, "dh_out %d misaligned", i);
  }
  A.dfeat = dfeat; A.ld_dfeat = ld_dfeat; A.F = feat_dim; A.M = m;
  A.n_tiles = (int)ceil_div(m, 128);
  A.slot = slot;
  const size_t fixed = (size_t)(kBAH + kBAL) * kAtomBytes + 8 * 2048 + 1024 + 2048;
  A.wn = (int)((226 * 1024 - fixed) / slot);
  if (A.wn > kBWN) A.wn = kBWN;
  TNF_REQUIRE(A.wn >= 3, "feature width too large for the weight ring");
  CUtensorMap tm_dh3, tm_dhs;
  int rc = make_atom_map_b(&tm_dh3, dh3, m, kBHid, kBHid);
  if (rc != TNF_OK) return rc;
  rc = make_atom_map_b(&tm_dhs, dhs, m, kBHid, kBHid);
  if (rc != TNF_OK) return rc;
  const size_t smem = (size_t)(kBAH + kBAL) * kAtomBytes + (size_t)A.wn * slot + 8 * 2048 + 1024;
  static PerDeviceOnce configured{};
  if (configured.pending()) {
    TNF_CUDA(cudaFuncSetAttribute(heads_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured.mark();
  }
  const int grid = A.n_tiles < sm_count() ? A.n_tiles : sm_count();
  heads_bwd_data_kernel<<<grid, kBThreads, smem, st>>>(A, tm_dh3, tm_dhs);
  TNF_LAUNCH_CHECK("heads_bwd_data_kernel");
  return TNF_OK;
}
