// mt19937.hpp -- host-side MT19937 with the two seeding rules and the four draw rules the reference's
// host RNGs use on this path:
//   CPython `random`      : seed(int) = init_by_array(32-bit limbs of |seed|); getrandbits(k<=32) = top k bits;
//                           _randbelow(n) = getrandbits(n.bit_length()) with rejection   (Lib/random.py)
//   legacy numpy RandomState: seed(u32) = init_genrand; randint = masked rejection on 32-bit draws;
//                           uniform = 53-bit double from two draws                      (randomkit.c)
// Reference call sites: random.sample(...) earl_benchmark/envs/tabletop_manipulation.py:66;
// np.random.uniform :115-117; np.random.randint envs/sawyer_peg.py:147,151, envs/kitchen.py:123.
#pragma once
#include <cstdint>

namespace earl {

class MT19937 {
 public:
  static constexpr int N = 624, M = 397;

  void init_genrand(uint32_t s) {
    mt_[0] = s;
    for (int i = 1; i < N; ++i) mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + (uint32_t)i;
    idx_ = N;
  }

  void init_by_array(const uint32_t* key, int len) {
    init_genrand(19650218u);
    int i = 1, j = 0;
    int k = N > len ? N : len;
    for (; k; --k) {
      mt_[i] = (mt_[i] ^ ((mt_[i - 1] ^ (mt_[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
      if (++i >= N) { mt_[0] = mt_[N - 1]; i = 1; }
      if (++j >= len) j = 0;
    }
    for (k = N - 1; k; --k) {
      mt_[i] = (mt_[i] ^ ((mt_[i - 1] ^ (mt_[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
      if (++i >= N) { mt_[0] = mt_[N - 1]; i = 1; }
    }
    mt_[0] = 0x80000000u;
    idx_ = N;
  }

  uint32_t next() {
    if (idx_ >= N) refill();
    uint32_t y = mt_[idx_++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }

  // CPython: getrandbits(k), 0 < k <= 32
  uint32_t getrandbits(int k) { return next() >> (32 - k); }

  // CPython random._randbelow_with_getrandbits(n), n >= 1
  uint32_t py_randbelow(uint32_t n) {
    int k = 32 - __builtin_clz(n);  // n.bit_length()
    uint32_t r = getrandbits(k);
    while (r >= n) r = getrandbits(k);
    return r;
  }

  // legacy numpy rk_interval(max = n-1): masked rejection; max == 0 consumes nothing
  uint32_t np_randint(uint32_t n) {
    uint32_t mx = n - 1;
    if (mx == 0) return 0;
    uint32_t mask = mx;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = next() & mask; } while (v > mx);
    return v;
  }

  // legacy numpy rk_double
  double np_double() {
    uint32_t a = next() >> 5, b = next() >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }

 private:
  void refill() {
    static const uint32_t mag[2] = {0u, 0x9908b0dfu};
    int kk = 0;
    for (; kk < N - M; ++kk) {
      uint32_t y = (mt_[kk] & 0x80000000u) | (mt_[kk + 1] & 0x7fffffffu);
      mt_[kk] = mt_[kk + M] ^ (y >> 1) ^ mag[y & 1u];
    }
    for (; kk < N - 1; ++kk) {
      uint32_t y = (mt_[kk] & 0x80000000u) | (mt_[kk + 1] & 0x7fffffffu);
      mt_[kk] = mt_[kk + (M - N)] ^ (y >> 1) ^ mag[y & 1u];
    }
    uint32_t y = (mt_[N - 1] & 0x80000000u) | (mt_[0] & 0x7fffffffu);
    mt_[N - 1] = mt_[M - 1] ^ (y >> 1) ^ mag[y & 1u];
    idx_ = 0;
  }

  uint32_t mt_[N];
  int idx_ = N + 1;
};

}  // namespace earl
