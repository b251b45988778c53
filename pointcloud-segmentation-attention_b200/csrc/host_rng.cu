// Host half of the whole-scene chunker: numpy's LEGACY shuffle stream, restated in C.
//
// The reference shuffles the point indices of every 1.5 m cell with np.random.shuffle on numpy's global RandomState
// (data/complete_scene_loader.py:17-18) -- MT19937 plus the legacy bounded-integer draw, a stream NumPy freezes for
// reproducibility (NEP 19).  The chunks, and with them every downstream prediction, depend on it bit for bit, so the
// stream cannot be replaced; but numpy's generic shuffle costs ~20 ns per element (memcpy swaps through a buffer), 3-6 ms
// per ~150 k-point scan and the largest share of the chunker's host time.  pc_host_legacy_shuffle continues the stream
// from a copied state (np.random.get_state()), writes the permutation of arange(n) numpy would have produced and
// returns the advanced state (np.random.set_state()) -- same draws, ~4 ns per element.
//
// Algorithm (numpy/random/mtrand.pyx RandomState.shuffle, 1-d ndarray path; numpy/random/src/distributions/
// distributions.c random_interval; numpy/random/src/mt19937/mt19937.c):
//   for i = n-1 .. 1:  j = random_interval(i);  swap(x[i], x[j])
//   random_interval(max): mask = smallest 2^k - 1 >= max;  draw next_uint32() & mask until <= max   (max < 2^32)
//   next_uint32(): MT19937 with the standard tempering; the 624-word state is regenerated when pos reaches 624.
// No device code in this file.
#include <stdint.h>
#include "common.cuh"

namespace {

constexpr int kMtN = 624, kMtM = 397;

inline void mt_regenerate(uint32_t *key) {
  constexpr uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrix = 0x9908b0dfu;
  int i = 0;
  for (; i < kMtN - kMtM; ++i) {
    const uint32_t y = (key[i] & kUpper) | (key[i + 1] & kLower);
    key[i] = key[i + kMtM] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
  }
  for (; i < kMtN - 1; ++i) {
    const uint32_t y = (key[i] & kUpper) | (key[i + 1] & kLower);
    key[i] = key[i + (kMtM - kMtN)] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
  }
  const uint32_t y = (key[kMtN - 1] & kUpper) | (key[0] & kLower);
  key[kMtN - 1] = key[kMtM - 1] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
}

inline uint32_t mt_next(uint32_t *key, int &pos) {
  if (pos == kMtN) {
    mt_regenerate(key);
    pos = 0;
  }
  uint32_t y = key[pos++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

}  // namespace

extern "C" int pc_host_legacy_shuffle(uint32_t *mt_key, int *mt_pos, int n, int *perm) {
  if (n < 0 || !mt_key || !mt_pos || (n > 0 && !perm) || *mt_pos < 0 || *mt_pos > kMtN) return PC_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < n; ++i) perm[i] = i;
  int pos = *mt_pos;
  for (int i = n - 1; i >= 1; --i) {
    uint32_t mask = (uint32_t)i;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t j;
    do {
      j = mt_next(mt_key, pos) & mask;
    } while (j > (uint32_t)i);
    const int t = perm[i];
    perm[i] = perm[j];
    perm[j] = t;
  }
  *mt_pos = pos;
  return PC_OK;
}

// np.random.choice(high, count, replace=True) = RandomState.randint(0, high, size=count) on the legacy stream
// (complete_scene_loader.py:87: the fill-up indices of a cell's last chunk): numpy/random/_bounded_integers.pyx
// _rand_int64 -> random_bounded_uint64_fill(off = 0, rng = high - 1, use_masked = 1), which for rng < 2^32 draws
// next_uint32() & mask until <= rng -- one masked-rejection draw per element, no buffering for 32-bit values.
extern "C" int pc_host_legacy_randint(uint32_t *mt_key, int *mt_pos, int high, int count, int *out) {
  if (high < 1 || count < 0 || !mt_key || !mt_pos || (count > 0 && !out) || *mt_pos < 0 || *mt_pos > kMtN)
    return PC_ERR_INVALID_ARGUMENT;
  const uint32_t rng = (uint32_t)(high - 1);
  if (rng == 0) {                      // numpy: no draw at all when the range is a single value
    for (int i = 0; i < count; ++i) out[i] = 0;
    return PC_OK;
  }
  uint32_t mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  int pos = *mt_pos;
  for (int i = 0; i < count; ++i) {
    uint32_t v;
    do {
      v = mt_next(mt_key, pos) & mask;
    } while (v > rng);
    out[i] = (int)v;
  }
  *mt_pos = pos;
  return PC_OK;
}
