// Library-level entry points of the C ABI (include/tinynerf_b200.h): version, errors, device info.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace tnf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local int g_sm_budget = 0;   // tnf_set_sm_budget: persistent grids of this thread use at most this many SMs

static thread_local int g_variant[kVariantCount] = {0};   // tnf_set_variant: diagnostic kernel selection of this thread
int variant(int which) { return (which >= 0 && which < kVariantCount) ? g_variant[which] : 0; }

int sm_count() {
  // per-device cache; benign race (same value written)
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return (g_sm_budget > 0 && g_sm_budget < cache[dev]) ? g_sm_budget : cache[dev];
}

}  // namespace tnf

extern "C" int tnf_version(void) { return 1001; }   // 1001: tnf_march_params.empty_bits

extern "C" int tnf_set_sm_budget(int n_sms) {
  const int prev = tnf::g_sm_budget;
  tnf::g_sm_budget = n_sms > 0 ? n_sms : 0;
  return prev;
}

extern "C" int tnf_set_variant(int which, int value) {
  if (which < 0 || which >= tnf::kVariantCount) return -1;
  const int prev = tnf::g_variant[which];
  tnf::g_variant[which] = value;
  return prev;
}

extern "C" const char* tnf_last_error(void) { return tnf::g_err; }

extern "C" int tnf_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  using namespace tnf;
  TNF_REQUIRE(sm_count && cc_major && cc_minor, "null output pointer");
  int dev = 0;
  TNF_CUDA(cudaGetDevice(&dev));
  TNF_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  TNF_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  TNF_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return TNF_OK;
}

// ---- host-side ray shuffling (the reference shuffles with DataLoader(shuffle=True), src/run.py:116-122) --------------------
// Incremental Fisher-Yates: `perm` (host, n entries, initially 0..n-1) is shuffled lazily, `count` positions per call, so an
// epoch never starts with an O(n) randperm (tens of milliseconds for millions of rays, seconds for a full-resolution scene)
// landing inside one training step.  Positions are a global counter g (slot g % n); g < *fresh_from have been drawn before
// (rays handed back by the dynamic-batch accumulator) and are re-read without consuming random numbers, so a replay returns
// the same rays.  With `world` ranks every rank walks the same sequence and keeps the positions g % world == rank.
static inline uint64_t tnf_splitmix64(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t tnf_bounded(uint64_t* s, uint64_t range) {   // uniform in [0, range), Lemire's method (unbiased)
  uint64_t x = tnf_splitmix64(s);
  __uint128_t m = (__uint128_t)x * range;
  uint64_t lo = (uint64_t)m;
  if (lo < range) {
    const uint64_t t = (0 - range) % range;
    while (lo < t) {
      x = tnf_splitmix64(s);
      m = (__uint128_t)x * range;
      lo = (uint64_t)m;
    }
  }
  return (uint64_t)(m >> 64);
}
extern "C" int tnf_shuffle_next(int64_t* perm, int64_t n, int64_t pos, int64_t count, int32_t rank, int32_t world,
                                int64_t* fresh_from, uint64_t* rng_state, int64_t* out) {
  using namespace tnf;
  TNF_REQUIRE(perm && fresh_from && rng_state && out, "null pointer");
  TNF_REQUIRE(n >= 1 && pos >= 0 && count >= 0 && world >= 1 && rank >= 0 && rank < world, "bad sizes");
  // Blocks of 256 positions: the swap partners of a block are drawn first and their cache lines prefetched, then the swaps
  // run -- early in an epoch every partner is a random line of an array far larger than the caches, and the misses of a
  // block overlap instead of queueing one behind the other (the draws depend on the generator only, never on the data).
  int64_t k = 0;
  const int64_t end = pos + count * world;
  int64_t js[256];
  for (int64_t g0 = pos; g0 < end; g0 += 256) {
    const int64_t g1 = g0 + 256 < end ? g0 + 256 : end;
    for (int64_t g = g0; g < g1; ++g) {
      const int64_t i = g % n;
      int64_t j = -1;
      if (g >= *fresh_from) {
        j = i + (int64_t)tnf_bounded(rng_state, (uint64_t)(n - i));
        *fresh_from = g + 1;
        __builtin_prefetch(perm + j, 1, 0);
      }
      js[g - g0] = j;
    }
    for (int64_t g = g0; g < g1; ++g) {
      const int64_t i = g % n, j = js[g - g0];
      if (j >= 0) {
        const int64_t t = perm[i];
        perm[i] = perm[j];
        perm[j] = t;
      }
      if (g % world == rank) out[k++] = perm[i];
    }
  }
  return TNF_OK;
}
