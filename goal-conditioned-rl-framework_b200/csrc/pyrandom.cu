// Host-side mirror of CPython's `random` draws on the hot path (no CUDA in this file).
//
// The reference draws every future index with random.randint(t+1, T-1) (src/buffer.py:153) and every
// minibatch with random.sample(deque, B) (src/buffer.py:124) from the interpreter's global
// Mersenne-Twister.  Bit-identical batches need that exact stream, and the Python-level loops cost
// ~0.15 ms per committed episode and ~0.1 ms per update.  These functions advance a copy of the
// interpreter's MT19937 state (random.getstate()) exactly as CPython 3.12 does:
//   getrandbits(k <= 32) = genrand_uint32() >> (32 - k)
//   _randbelow(n)        = k = n.bit_length(); r = getrandbits(k); while r >= n: r = getrandbits(k)
//   randint(a, b)        = a + _randbelow(b - a + 1)
//   sample(range(n), k)  = the pool algorithm for n <= setsize, else the selection-set algorithm
// and the caller writes the state back with random.setstate().
#include <cmath>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kN = 624, kM = 397;

struct MT {
  uint32_t *mt;
  int pos;
  uint32_t next() {
    if (pos >= kN) {
      int kk;
      uint32_t y;
      for (kk = 0; kk < kN - kM; ++kk) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + kM] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      for (; kk < kN - 1; ++kk) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (kM - kN)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      y = (mt[kN - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
      mt[kN - 1] = mt[kM - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      pos = 0;
    }
    uint32_t y = mt[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  // n in [1, 2^32)
  uint64_t randbelow(uint64_t n) {
    int k = 0;
    for (uint64_t t = n; t; t >>= 1) ++k;          // n.bit_length()
    uint64_t r;
    do {
      r = uint64_t(next()) >> (32 - k);
    } while (r >= n);
    return r;
  }
};

}  // namespace

extern "C" {

int gcrl_pyrandom_randint(uint32_t *mt624, int *pos, int64_t count, const int32_t *lo, const int32_t *hi,
                          int32_t *out) {
  GCRL_API_BEGIN
  GCRL_REQUIRE(mt624 && pos && (count == 0 || (lo && hi && out)), "NULL argument");
  GCRL_REQUIRE(*pos >= 0 && *pos <= kN, "bad Mersenne-Twister position");
  MT g{mt624, *pos};
  for (int64_t i = 0; i < count; ++i) {
    GCRL_REQUIRE(hi[i] >= lo[i], "empty range for randint");
    out[i] = lo[i] + int32_t(g.randbelow(uint64_t(int64_t(hi[i]) - lo[i] + 1)));
  }
  *pos = g.pos;
  GCRL_API_END
}

int gcrl_pyrandom_sample_range(uint32_t *mt624, int *pos, int64_t n, int64_t k, int64_t *out) {
  GCRL_API_BEGIN
  GCRL_REQUIRE(mt624 && pos && (k == 0 || out), "NULL argument");
  GCRL_REQUIRE(*pos >= 0 && *pos <= kN, "bad Mersenne-Twister position");
  GCRL_REQUIRE(k >= 0 && k <= n, "Sample larger than population or is negative");
  GCRL_REQUIRE(n < (int64_t(1) << 32), "population too large for the 32-bit getrandbits path");
  MT g{mt624, *pos};
  int64_t setsize = 21;
  if (k > 5) setsize += int64_t(std::llround(std::pow(4.0, std::ceil(std::log(double(k) * 3.0) / std::log(4.0)))));
  if (n <= setsize) {
    std::vector<int64_t> pool(size_t(n), 0);
    for (int64_t i = 0; i < n; ++i) pool[size_t(i)] = i;
    for (int64_t i = 0; i < k; ++i) {
      const uint64_t j = g.randbelow(uint64_t(n - i));
      out[i] = pool[j];
      pool[j] = pool[size_t(n - i - 1)];
    }
  } else {
    // membership set of the positions drawn so far (CPython uses a set; only membership matters): open
    // addressing over a power-of-two table, no allocation per element
    size_t cap = 16;
    while (cap < size_t(k) * 4) cap <<= 1;
    std::vector<int64_t> table(cap, -1);
    const size_t mask = cap - 1;
    auto slot_of = [&](int64_t j) {
      size_t h = size_t(uint64_t(j) * 0x9E3779B97F4A7C15ull >> 32) & mask;
      while (table[h] != -1 && table[h] != j) h = (h + 1) & mask;
      return h;
    };
    for (int64_t i = 0; i < k; ++i) {
      int64_t j = int64_t(g.randbelow(uint64_t(n)));
      size_t h = slot_of(j);
      while (table[h] == j) {
        j = int64_t(g.randbelow(uint64_t(n)));
        h = slot_of(j);
      }
      table[h] = j;
      out[i] = j;
    }
  }
  *pos = g.pos;
  GCRL_API_END
}

}  // extern "C"
