// sb_host.h -- small internal helpers shared by the host-side translation units.
#ifndef SB_HOST_H
#define SB_HOST_H
#include <cstdint>

namespace sb {
// A=0 C=1 G=2 T=3 (either case), anything else -1 (utils.h:73-86 maps those to 4)
int base_code(char c);

// splitmix64 / xoshiro256** : small, fast, reproducible across platforms
struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t &x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed, uint64_t stream = 0) {
    uint64_t x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
    for (int i = 0; i < 4; ++i) s[i] = splitmix(x);
  }
  static uint64_t rotl(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double uniform_pos() { return ((double)(next() >> 11) + 1.0) * (1.0 / 9007199254740992.0); }  // (0,1]
  uint64_t below(uint64_t n) { return (uint64_t)(uniform() * (double)n); }
};
}  // namespace sb
#endif
