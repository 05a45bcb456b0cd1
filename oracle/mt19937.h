/* TEST INFRASTRUCTURE (CPU oracle) -- not part of the product path.
 *
 * numpy.random.RandomState draw recipes restated in plain C.  The reference draws everything
 * from the legacy generator (mitty/simulation/readgenerate.py:147-156, illumina.py:56-58,70-73,
 * 93,151-153, readcorrupt.py:31,36,84).  numpy itself is a third-party dependency not vendored
 * under /root/reference (setup.py:16 pins numpy>=1.9.0; installed here: 2.3.5).  RandomState's
 * stream is frozen by numpy policy (NEP 19), so the published algorithm is restated:
 *   - MT19937 (Matsumoto & Nishimura 1998), init_genrand seeding with the Knuth multiplier,
 *   - random_sample  = ((u32>>5)*2^26 + (u32>>6)) / 2^53,
 *   - bounded ints   = masked rejection on one 32-bit word per attempt,
 *   - int8 bounded   = 4 values per 32-bit word, low byte first,
 *   - geometric(p<1/3) = ceil(log(1-U)/log(1-p)),
 *   - shuffle        = Fisher-Yates from the top with random_interval(i).
 * tests/test_oracle_rng.py pins every recipe against numpy on this machine.
 */
#ifndef ORACLE_MT19937_H
#define ORACLE_MT19937_H
#include <stdint.h>
#include <math.h>

typedef struct { uint32_t mt[624]; int mti; } mt_state;

static inline void mt_seed(mt_state *s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; i++)
    s->mt[i] = 1812433253U * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->mti = 624;
}

static inline void mt_refill(mt_state *s) {
  uint32_t *mt = s->mt, y;
  int k;
  for (k = 0; k < 624 - 397; k++) {
    y = (mt[k] & 0x80000000U) | (mt[k + 1] & 0x7fffffffU);
    mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
  }
  for (; k < 623; k++) {
    y = (mt[k] & 0x80000000U) | (mt[k + 1] & 0x7fffffffU);
    mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
  }
  y = (mt[623] & 0x80000000U) | (mt[0] & 0x7fffffffU);
  mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
  s->mti = 0;
}

static inline uint32_t mt_u32(mt_state *s) {
  if (s->mti >= 624) mt_refill(s);
  uint32_t y = s->mt[s->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680U;
  y ^= (y << 15) & 0xefc60000U;
  y ^= (y >> 18);
  return y;
}

/* RandomState.random_sample / rand */
static inline double mt_double(mt_state *s) {
  uint32_t a = mt_u32(s) >> 5, b = mt_u32(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* RandomState.randint(low=0, high=max+1) for max < 2^32-1 : masked rejection, one word per try */
static inline uint32_t mt_bounded(mt_state *s, uint32_t max) {
  if (max == 0) return 0;
  uint32_t mask = max, v;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  do { v = mt_u32(s) & mask; } while (v > max);
  return v;
}

/* RandomState.geometric(p), p < 1/3 (every p read_model_params can produce is <= 0.1) */
static inline int64_t mt_geometric(mt_state *s, double p) {
  return (int64_t)ceil(log(1.0 - mt_double(s)) / log(1.0 - p));
}

/* RandomState.randint(2, size=n, dtype='i1') */
static inline void mt_bits_i8(mt_state *s, int8_t *out, int64_t n) {
  uint32_t buf = 0; int left = 0;
  for (int64_t i = 0; i < n; i++) {
    if (!left) { buf = mt_u32(s); left = 4; } else { buf >>= 8; }
    left--;
    out[i] = (int8_t)(buf & 1U);
  }
}

/* RandomState.shuffle on a 1-D int64 array */
static inline void mt_shuffle_i64(mt_state *s, int64_t *x, int64_t n) {
  for (int64_t i = n - 1; i >= 1; i--) {
    int64_t j = (int64_t)mt_bounded(s, (uint32_t)i);
    int64_t t = x[i]; x[i] = x[j]; x[j] = t;
  }
}
#endif
