#include <stdio.h>
#include <stdint.h>
#include "common.cuh"
int main() {
  unsigned long long seeds[] = {0ull, 1ull, 12345ull, 0xDEADBEEFCAFEull, 0xFFFFFFFFFFFFFFFFull, (7ull << 32) | 5ull};
  long long bad = 0, n = 0;
  double kept = 0;
  for (unsigned long long seed : seeds) {
    const uint32_t s32 = tasr_seed_mix(seed);
    unsigned long long starts[] = {0ull, 1000ull, 0xFFFFFFF0ull, (1ull << 33) + 77ull, 0x123456789ull};
    for (unsigned long long p0 : starts) {
      const uint32_t base = tasr_hash_pair_base_s32(s32, p0);
      if (base != tasr_hash_pair_base(seed, p0)) ++bad;
      for (uint32_t j = 0; j < 4096; ++j) {
        // fast path of a run starting at p0 vs the generic hash of pair p0 + j (valid while lo32 does not wrap)
        if ((uint32_t)p0 + j < (uint32_t)p0) break;
        const uint32_t t = tasr_hash_finish(base + j * TASR_HASH_C1);
        const uint32_t packed = (t & 0xFFFF0000u) | ((t * TASR_HASH_GOLD) >> 16);
        if (packed != tasr_hash_pair(seed, p0 + j)) ++bad;
        const uint32_t th = tasr_drop_thresh16(0.1f);
        kept += ((packed & 0xFFFFu) >= th) + ((packed >> 16) >= th);
        n += 2;
      }
    }
  }
  printf("%lld %lld %.6f\n", bad, n, kept / n);
  return 0;
}
