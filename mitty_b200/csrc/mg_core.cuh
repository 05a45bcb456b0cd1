// Per-thread logic of the read-generation engine, shared by every kernel.
//
// Everything here is __host__ __device__ and free of CUDA-only constructs so that
// tests/emul/ can compile the very same code with g++ and check it against the oracle on
// the build box (which has no GPU).  The product never runs these on the CPU: the only
// product callers are the kernels in mg_kernels.cu.
//
// Reference semantics restated (paths under /root/reference):
//   node lookup      rpc.get_begin_end_nodes   mitty/simulation/rpc.py:119-130
//   pos/cigar/v_list rpc.generate_read         mitty/simulation/rpc.py:133-160
//   qname / record   readgenerate.fastq_lines  mitty/simulation/readgenerate.py:222-230
//   N filter/revcomp read_generating_worker    mitty/simulation/readgenerate.py:198-210
//   corruption       illumina.corrupt_single_read  mitty/simulation/illumina.py:131-162
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD inline
#endif
// MG_NI: deliberately NOT inlined.  The emit kernel is instruction-cache bound when every helper
// is inlined at every call site (13 k SASS instructions); the formatting helpers are shared.
#if defined(__CUDACC__)
#define MG_NI static __host__ __device__ __noinline__ __attribute__((unused))
#else
#define MG_NI static __attribute__((noinline, unused))
#endif
#if defined(__CUDA_ARCH__)
#define MG_UNROLL _Pragma("unroll")
#define MG_NOUNROLL _Pragma("unroll 1")
#else
#define MG_UNROLL
#define MG_NOUNROLL
#endif

// ------------------------------------------------------------------------------------------
// Data layout in HBM

// One node of a chromosome copy (rpc.Node, rpc.py:5-35) in 16 bytes -> one 128-bit load.
//   key   = ps - p_min (+1 for 'D' nodes: the searchsorted key of rpc.py:127)
//   pr    = 1-based reference position
//   op    = '=', 'X', 'I' or 'D' (ASCII)
struct alignas(16) MgNode {
  uint32_t key;
  int32_t pr;
  int32_t oplen;
  uint32_t op;
};

// Maximal run of one non-ACGT byte on the haplotype, in sample-relative coordinates -- or, with
// byte == MG_EXC_CASE, a maximal run of lower-case a/c/g/t (a soft-masked stretch): those bases keep
// their 2-bit codes in the packed sequence and the run only records the case.
#define MG_EXC_CASE 1u
struct alignas(16) MgExc {
  uint32_t start;
  uint32_t len;
  uint32_t byte;
  uint32_t pad;
};

// ------------------------------------------------------------------------------------------
// Bit tricks with portable fall-backs (the device versions map to single instructions)

MG_HD uint32_t mg_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // ((hi:lo) >> sh) low word, sh in [0,31]
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}

MG_HD uint32_t mg_funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) {  // ((hi:lo) << sh) high word, sh in [0,31]
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, sh);
#else
  return sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
#endif
}

MG_HD uint32_t mg_brev(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
  return (x >> 16) | (x << 16);
#endif
}

MG_HD uint32_t mg_prmt(uint32_t a, uint32_t sel) {  // byte i of result = byte (nibble i of sel & 3) of a
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, 0u, sel);
#else
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= ((a >> (8 * ((sel >> (4 * i)) & 3))) & 0xFFu) << (8 * i);
  return r;
#endif
}

// 16 consecutive 2-bit codes starting at base index s (s >= 0) of a packed sequence
// (16 bases per 32-bit word, base i in bits [2i, 2i+1] of word i/16).
template <class P>
MG_HD uint32_t mg_codes16(P seq, int64_t s) {
  int64_t w = s >> 4;
  uint32_t sh = (uint32_t)(s & 15) * 2u;
  uint32_t lo = seq[w];
  uint32_t hi = sh ? seq[w + 1] : 0u;
  return mg_funnel_r(lo, hi, sh);
}

// reverse the order of the 16 codes in a word and complement them (A<->T, C<->G is code ^ 3)
MG_HD uint32_t mg_revcomp16(uint32_t x) {
  uint32_t y = mg_brev(x);
  y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
  return ~y;
}

// four 2-bit codes (low byte of b) -> four ASCII bases
MG_HD uint32_t mg_chars4(uint32_t b) {
  uint32_t y = (b | (b << 4)) & 0x0F0Fu;
  uint32_t z = (y | (y << 2)) & 0x3333u;
  return mg_prmt(0x54474341u /* 'A','C','G','T' little-endian */, z);
}

MG_HD uint32_t mg_base_code(uint8_t c) {  // ACGT -> 0..3, anything else -> 4
  return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), the counter-based generator of the production mode

struct MgPhilox { uint32_t v[4]; };

MG_HD void mg_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
}

template <int ROUNDS>
MG_HD MgPhilox mg_philox_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  MG_UNROLL
  for (int r = 0; r < ROUNDS; r++) {
    uint32_t h0, l0, h1, l1;
    mg_mulhilo(0xD2511F53u, c0, h0, l0);
    mg_mulhilo(0xCD9E8D57u, c2, h1, l1);
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  MgPhilox o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// Philox4x32-10: template sampling (gaps, template length, file order)
MG_HD MgPhilox mg_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  return mg_philox_r<10>(c0, c1, c2, c3, k0, k1);
}

// Philox4x32-7 (the fewest rounds Salmon et al. report as passing BigCrush): the per-base
// corruption stream, where the generator is ~60 % of the instruction count
#define MG_CORRUPT_ROUNDS 7
MG_HD MgPhilox mg_philox_corrupt(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t k0, uint32_t k1) {
  return mg_philox_r<MG_CORRUPT_ROUNDS>(c0, c1, c2, 0x636f7272u /* MG_STREAM_CORRUPT */, k0, k1);
}

// 53-bit uniform in [0,1) from two words, the same construction numpy's random_sample uses
MG_HD double mg_u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

// Philox stream tags (counter word 3)
#define MG_STREAM_GAP 0x67617073u
#define MG_STREAM_TLEN 0x746c656eu
#define MG_STREAM_CORRUPT 0x636f7272u

// Keyed pseudo-random permutation of [0, n): balanced Feistel network on the next even power of
// two with cycle walking.  Replaces RandomState.shuffle (illumina.py:71) in production mode.
MG_HD uint32_t mg_feistel_round(uint32_t r, uint32_t k) {
  uint32_t x = r * 0x9E3779B1u + k;
  x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; x *= 0xC2B2AE3Du; x ^= x >> 16;
  return x;
}

MG_HD uint32_t mg_permute(uint32_t i, uint32_t n, uint32_t half_bits, uint32_t k0, uint32_t k1) {
  const uint32_t hm = (1u << half_bits) - 1u;
  do {
    uint32_t l = i >> half_bits, r = i & hm;
    for (int t = 0; t < 6; t++) {
      uint32_t f = mg_feistel_round(r, (t & 1) ? k1 + (uint32_t)t : k0 + (uint32_t)t) & hm;
      uint32_t nl = r; r = l ^ f; l = nl;
    }
    i = (l << half_bits) | r;
  } while (i >= n);
  return i;
}

// ------------------------------------------------------------------------------------------
// Decimal helpers

MG_HD uint32_t mg_pow10(int d) {   // 10^d for d in 0..9
  switch (d) {
    case 0: return 1u; case 1: return 10u; case 2: return 100u; case 3: return 1000u; case 4: return 10000u;
    case 5: return 100000u; case 6: return 1000000u; case 7: return 10000000u; case 8: return 100000000u;
    default: return 1000000000u;
  }
}

MG_NI int mg_ndigits32(uint32_t v) {   // no division: bit length -> digit estimate -> one correction
#if defined(__CUDA_ARCH__)
  const int bits = 32 - __clz((int)(v | 1u));
#else
  const int bits = 32 - __builtin_clz(v | 1u);
#endif
  const int g = (bits * 1233) >> 12;          // floor(bits * log10(2)), g in 0..9
  return g + ((v >= mg_pow10(g)) || v == 0u ? 1 : 0);
}

MG_HD int mg_ndigits(uint64_t v) {
  if (v <= 0xFFFFFFFFull) return mg_ndigits32((uint32_t)v);
  int d = 1;
  while (v >= 10) { v /= 10; d++; }
  return d;
}


// sum of the decimal lengths of 1..m  (closed form; used to place records whose qname carries a
// serial number that is only known after the block/grid scan):  d*(m+1) - 11..1 (d ones)
MG_NI uint64_t mg_digit_sum(uint64_t m) {
  if (m == 0) return 0;
  if (m <= 0xFFFFFFFFull) {
    const int d = mg_ndigits32((uint32_t)m);
    const uint64_t ones = d < 10 ? (uint64_t)((mg_pow10(d) - 1u) / 9u) : 1111111111ull;
    return (uint64_t)d * (m + 1) - ones;
  }
  const int d = mg_ndigits(m);
  uint64_t ones = 0, p = 1;
  for (int k = 0; k < d; k++) { ones += p; p *= 10; }
  return (uint64_t)d * (m + 1) - ones;
}

// ------------------------------------------------------------------------------------------
// Writers.  CountWriter sizes a record, WordStream writes it (any byte alignment) with aligned
// 32-bit stores in the interior and byte stores only for the first / last partial word.

struct MgCountWriter {
  static constexpr bool is_bytes = false;
  uint32_t n;
  MG_HD void put(uint8_t) { n++; }
  MG_HD void put_word(uint32_t) { n += 4; }
};

// Address-space policies.  The emit kernel formats records into its warp's shared-memory stage:
// MgSharedSpace addresses it with 32-bit offsets into the kernel's dynamic shared memory, so every
// store is a plain STS (generic 64-bit pointers cost 3-4 instructions per store).  MgGenericSpace
// is any memory: host emulation, and the rare record that is larger than the stage.
struct MgGenericSpace {
  static constexpr bool is_generic = true;
  typedef uint8_t *ptr;
  static MG_HD void st8(ptr p, uint8_t v) { *p = v; }
  static MG_HD void st32(ptr p, uint32_t v) { *reinterpret_cast<uint32_t *>(p) = v; }
  static MG_HD uint32_t ld32(ptr p) { return *reinterpret_cast<const uint32_t *>(p); }
  static MG_HD uint32_t low2(ptr p) { return (uint32_t)((uintptr_t)p & 3); }
};
#if defined(__CUDACC__)
struct MgSharedSpace {
  static constexpr bool is_generic = false;
  // a byte address in the shared window (cvta.to.shared of the kernel's stage): plain 32-bit
  // st.shared / ld.shared with no generic-address arithmetic at the use sites
  typedef uint32_t ptr;
  static __device__ __forceinline__ void st8(ptr p, uint8_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(p), "r"((uint32_t)v)); }
  static __device__ __forceinline__ void st32(ptr p, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(p), "r"(v)); }
  static __device__ __forceinline__ uint32_t ld32(ptr p) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(p));
    return v;
  }
  static __device__ __forceinline__ uint32_t low2(ptr p) { return p & 3u; }
};
#endif

// plain byte stores: used for the qname, whose per-lane byte counts differ (a shared word
// stream would flush on a different iteration in every lane and serialise the warp)
template <class SP>
struct MgByteWriter {
  static constexpr bool is_bytes = true;
  typedef SP space;
  typename SP::ptr p;
  MG_HD void put(uint8_t c) { SP::st8(p, c); p += 1; }
};

template <class SP> MG_NI void mg_store_tail(typename SP::ptr b, uint32_t carry, uint32_t nb) {
  for (uint32_t i = 0; i < nb; i++) SP::st8(b + i, (uint8_t)(carry >> (8 * i)));
}

template <class SP>
struct MgWordStream {
  static constexpr bool is_bytes = false;
  typename SP::ptr wp;   // next aligned word
  uint32_t prev;         // the pending bytes are the TOP sh / 8 bytes of prev
  uint32_t sh;           // 8 * pending byte count: 0, 8, 16 or 24

  // The bytes before dst inside its word were written earlier BY THIS THREAD (the tail of its own
  // qname / separator): they are read back and carried, so every flush is a plain word store.
  MG_HD void begin_rmw(typename SP::ptr dst) {
    const uint32_t a = SP::low2(dst);
    wp = dst - a;
    sh = 8 * a;
    prev = a ? (SP::ld32(wp) << (32 - sh)) : 0u;
  }
  MG_HD void flush_word(uint32_t w) { SP::st32(wp, w); wp += 4; }
  MG_HD void put(uint8_t c) {
    prev = (prev >> 8) | ((uint32_t)c << 24);
    sh += 8;
    if (sh == 32) { flush_word(prev); sh = 0; }
  }
  MG_HD void put_word(uint32_t w) {  // four bytes, little-endian order: one funnel shift, one store
    flush_word(mg_funnel_l(prev, w, sh));    // (w << sh) | (prev >> (32 - sh)); sh == 0 gives w
    prev = w;
  }
  // the last partial word is shared with the NEXT record (another thread): byte stores
  MG_HD void end() { if (sh) mg_store_tail<SP>(wp, prev >> (32 - sh), sh >> 3); sh = 0; }
};

// decimal digits of v at p (byte stores), returns the advanced pointer
template <class SP> MG_NI typename SP::ptr mg_put_u32_p(typename SP::ptr p, uint32_t v) {
  uint64_t acc = 0;
  int n = 0;
  do { uint32_t q = v / 10u; acc = (acc << 4) | (v - q * 10u); v = q; n++; } while (v);
  for (; n; n--) { SP::st8(p, (uint8_t)('0' + ((uint32_t)acc & 15u))); p += 1; acc >>= 4; }
  return p;
}

template <class W>
MG_HD void mg_put_u32(W &w, uint32_t v) {
  if constexpr (W::is_bytes) { w.p = mg_put_u32_p<typename W::space>(w.p, v); return; }
  else {
    // digits are stacked as nibbles in a register pair (no local-memory array); /10 is a multiply
    uint64_t acc = 0;
    int n = 0;
    do { uint32_t q = v / 10u; acc = (acc << 4) | (v - q * 10u); v = q; n++; } while (v);
    for (; n; n--) { w.put((uint8_t)('0' + ((uint32_t)acc & 15u))); acc >>= 4; }
  }
}

template <class W>
MG_HD void mg_put_uint(W &w, uint64_t v) {   // v < 10^16
  if (v <= 0xFFFFFFFFull) { mg_put_u32(w, (uint32_t)v); return; }
  uint64_t acc = 0;
  int n = 0;
  do { acc = (acc << 4) | (v % 10); v /= 10; n++; } while (v);
  for (; n; n--) { w.put((uint8_t)('0' + (acc & 15))); acc >>= 4; }
}

template <class W>
MG_HD void mg_put_bytes(W &w, const uint8_t *s, int n) {
  for (int i = 0; i < n; i++) w.put(s[i]);
}

// ------------------------------------------------------------------------------------------
// Node lookup: searchsorted(keys, x, 'right') - 1 (rpc.py:127-130), accelerated by a block table
// blk[b] = last node whose key <= (b << blk_shift).

template <class NP, class BP>
MG_NI int mg_find_node(NP nodes, BP blk, int blk_shift, int n_blk, int n_nodes, uint32_t x) {
  uint32_t b = x >> blk_shift;
  if ((int)b >= n_blk) b = (uint32_t)(n_blk - 1);
  int lo = (int)blk[b];
  int hi = ((int)b + 1 < n_blk) ? (int)blk[b + 1] : n_nodes - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (nodes[mid].key <= x) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Per-read geometry shared by the sizing pass and the formatting pass.
//   pos (rpc.py:148-158), whether the read lies inside one insertion (rpc.py:149-154) and the
//   last node n1 = searchsorted(keys, x+L-1, 'right') - 1, found by walking forward from n0.
template <class NP>
MG_HD int mg_last_node(NP nodes, int n0, int n_nodes, uint32_t x, int L) {
  const uint32_t last = x + (uint32_t)L - 1u;
  int n1 = n0;
  while (n1 + 1 < n_nodes && nodes[n1 + 1].key <= last) n1++;
  return n1;
}

MG_HD int32_t mg_read_pos(const MgNode &f, bool single, uint32_t x) {
  if (f.op == 'I') return single ? f.pr - 1 : f.pr;                         // rpc.py:148-156
  const int64_t ps0 = (int64_t)f.key - (f.op == 'D' ? 1 : 0);
  return (int32_t)((int64_t)x - ps0 + (int64_t)f.pr);                        // rpc.py:158
}

MG_HD int32_t mg_cigar_len(const MgNode &n, uint32_t x, int L) {             // rpc.py:145
  if (n.op == 'D') return n.oplen;
  int64_t a = (int64_t)x - (int64_t)n.key; if (a < 0) a = 0;
  int64_t b = (int64_t)x + L - (int64_t)n.key; if ((int64_t)n.oplen < b) b = n.oplen;
  return (int32_t)(b - a);
}

template <class W>
MG_HD void mg_put_i32(W &w, int32_t v) {
  if (v < 0) { w.put('-'); mg_put_u32(w, (uint32_t)(-(int64_t)v)); } else mg_put_u32(w, (uint32_t)v);
}

MG_HD int mg_nchars_i32(int32_t v) { return v < 0 ? 1 + mg_ndigits32((uint32_t)(-(int64_t)v)) : mg_ndigits32((uint32_t)v); }

// length of '|strand|pos|rlen|cigar|vlist' for one read; L_nd = decimal digits of L
template <class NP>
MG_NI uint32_t mg_read_fields_len(NP nodes, int n0, int n1, uint32_t x, int L, int L_nd) {
  const MgNode f = nodes[n0];
  const bool single = (n0 == n1);
  uint32_t n = 2u + 1u + (uint32_t)mg_nchars_i32(mg_read_pos(f, single, x)) + 1u + (uint32_t)L_nd + 1u + 1u;
  if (single && f.op == '=') return n + (uint32_t)L_nd + 1u;                 // "<L>=" and an empty v_list
  if (single && f.op == 'I')                                                 // ">p:<L>I" and "<oplen>"
    return n + 1u + (uint32_t)mg_nchars_i32((int32_t)((int64_t)x - (int64_t)f.key)) + 1u + (uint32_t)L_nd + 1u + (uint32_t)mg_nchars_i32(f.oplen);
  uint32_t nv = 0;
  for (int k = n0; k <= n1; k++) {
    const MgNode nd = nodes[k];
    n += (uint32_t)mg_nchars_i32(mg_cigar_len(nd, x, L)) + 1u;
    if (nd.op != '=') {
      n += (nv ? 1u : 0u) + (nd.op == 'X' ? 1u : nd.op == 'I' ? (uint32_t)mg_nchars_i32(nd.oplen) : (uint32_t)mg_nchars_i32(-nd.oplen));
      nv++;
    }
  }
  return n;
}

// '|strand|pos|rlen|cigar|vlist' for one read (fastq_lines, readgenerate.py:224-225, over
// generate_read, rpc.py:144-158).  x = read start relative to p_min, L = read length.
template <class W, class NP>
MG_HD void mg_fmt_read(W &w, NP nodes, int n0, int n1, uint32_t x, int L, int strand) {
  const MgNode f = nodes[n0];
  const bool single = (n0 == n1);
  w.put('|'); w.put((uint8_t)('0' + strand));
  w.put('|'); mg_put_i32(w, mg_read_pos(f, single, x));
  w.put('|'); mg_put_u32(w, (uint32_t)L);
  w.put('|');
  if (single && f.op == '=') {                                               // the common case: "<L>=" + empty v_list
    mg_put_u32(w, (uint32_t)L); w.put('='); w.put('|');
    return;
  }
  if (single && f.op == 'I') {                                               // rpc.py:154
    w.put('>'); mg_put_i32(w, (int32_t)((int64_t)x - (int64_t)f.key)); w.put(':'); mg_put_u32(w, (uint32_t)L); w.put('I');
  } else {
    for (int k = n0; k <= n1; k++) {                                         // rpc.py:145
      const MgNode n = nodes[k];
      mg_put_i32(w, mg_cigar_len(n, x, L)); w.put((uint8_t)n.op);
    }
  }
  w.put('|');
  bool first = true;
  for (int k = n0; k <= n1; k++) {                                           // rpc.py:144
    const MgNode n = nodes[k];
    if (n.op == '=') continue;
    if (!first) w.put(',');
    first = false;
    if (n.op == 'X') w.put('0');
    else mg_put_i32(w, n.op == 'I' ? n.oplen : -n.oplen);
  }
}

// Whole qname line without the trailing newline:
//   '@' stub ':' cnt '|' chrom '|' cpy  + per read in FILE order '|strand|pos|rlen|cigar|vlist'
// prefix = "@<sample>:<worker>:<ps>:"   mid = "|<chrom>|<cpy>"
struct MgReadRef { uint32_t x; int n0, n1, strand; };

template <class W, class NP>
MG_HD void mg_fmt_qname(W &w, const uint8_t *prefix, int prefix_len, uint64_t cnt, bool with_cnt,
                        const uint8_t *mid, int mid_len, NP nodes, MgReadRef first, MgReadRef second, int L) {
  mg_put_bytes(w, prefix, prefix_len);
  if (with_cnt) { if (cnt <= 0xFFFFFFFFull) mg_put_u32(w, (uint32_t)cnt); else mg_put_uint(w, cnt); }
  mg_put_bytes(w, mid, mid_len);
  MG_NOUNROLL
  for (int r = 0; r < 2; r++) {
    const MgReadRef R = r ? second : first;
    mg_fmt_read(w, nodes, R.n0, R.n1, R.x, L, R.strand);
  }
}

// qname + newline as bytes at dst (one shared copy per kernel), returns the advanced pointer
template <class SP, class NP>
MG_NI typename SP::ptr mg_qname_bytes(typename SP::ptr dst, const uint8_t *prefix, int prefix_len, uint64_t cnt,
                                      const uint8_t *mid, int mid_len, NP nodes, MgReadRef first, MgReadRef second, int L) {
  MgByteWriter<SP> bw; bw.p = dst;
  mg_fmt_qname(bw, prefix, prefix_len, cnt, true, mid, mid_len, nodes, first, second, L);
  bw.put('\n');
  return bw.p;
}

// Register window over a read: the 2-bit words covering it are fetched up front (independent
// loads, all in flight at once) and, for the reverse strand, reversed and complemented word by
// word in descending order, so that BOTH strands become a forward extraction
//   codes(c) = funnel(w[c], w[c+1], 2 * off)
// with compile-time word indices.  MAXW - 1 >= ceil((15 + L) / 16).
template <int MAXW>
struct MgWin { uint32_t w[MAXW]; uint32_t off; };

template <int MAXW, class HP>
MG_HD void mg_win_load(MgWin<MAXW> &W, HP hap, uint32_t x, int L, int strand) {
  const uint32_t last = x + (uint32_t)L - 1u;
  const int64_t w_first = (int64_t)(x >> 4), w_last = (int64_t)(last >> 4);
  W.off = strand ? 15u - (last & 15u) : (x & 15u);
  const int nw = (int)((W.off + (uint32_t)L + 15u) >> 4);
  MG_UNROLL
  for (int i = 0; i < MAXW; i++) {
    uint32_t v = 0;
    if (i < nw) v = hap[strand ? w_last - i : w_first + i];
    W.w[i] = strand ? mg_revcomp16(v) : v;
  }
}

template <int MAXW>
MG_HD uint32_t mg_win_codes(const MgWin<MAXW> &W, int c) {  // c must be a compile-time constant after unrolling
  return mg_funnel_r(W.w[c], W.w[c + 1], 2u * W.off);
}

// The emission loops run over chunk PAIRS with a rolled loop (a fully unrolled read is ~1000
// instructions per site and thrashes the instruction cache): chunks (w[0],w[1]) and (w[1],w[2]) are
// consumed, then the window slides down by two registers.
template <int MAXW>
MG_HD void mg_win_slide2(MgWin<MAXW> &W) {
  MG_UNROLL
  for (int i = 0; i < MAXW; i++) W.w[i] = (i + 2 < MAXW) ? W.w[i + 2] : 0u;
}

template <class W>
MG_HD void mg_emit_fill(W &w, uint8_t c, int n) {
  uint32_t cw = 0x01010101u * c;
  int k = 0;
  for (; k + 4 <= n; k += 4) w.put_word(cw);
  for (; k < n; k++) w.put(c);
}

// Exception runs (non-ACGT bytes) overlapping [x, x+L): first run with start+len > x.
template <class EP>
MG_HD int mg_exc_first(EP exc, int n_exc, uint32_t x) {
  int lo = 0, hi = n_exc;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((uint64_t)exc[mid].start + exc[mid].len <= (uint64_t)x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// number of 'N' in the read (seq.count('N'), readgenerate.py:204)
template <class EP>
MG_NI int mg_count_N(EP exc, int n_exc, uint32_t x, int L) {
  int cnt = 0;
  for (int k = mg_exc_first(exc, n_exc, x); k < n_exc; k++) {
    MgExc e = exc[k];
    if ((uint64_t)e.start >= (uint64_t)x + L) break;
    if (e.byte != 'N') continue;
    uint64_t a = e.start > x ? e.start : x;
    uint64_t b = (uint64_t)e.start + e.len < (uint64_t)x + L ? (uint64_t)e.start + e.len : (uint64_t)x + L;
    cnt += (int)(b - a);
  }
  return cnt;
}

// overwrite the bases of a written read (seq points at its first byte) that fall in exception
// runs.  The reference's translate table only maps ATCGN (readgenerate.py:56), so an exception
// byte -- lower-case bases included -- is copied unchanged on either strand; only its position is
// mirrored on strand 1.
// the forward-strand byte of haplotype position i under exception run e: the run's byte, or the
// lower-case letter of the packed code for a case run
template <class HP>
MG_HD uint8_t mg_exc_byte(const MgExc &e, HP hap, uint64_t i) {
  if (e.byte != MG_EXC_CASE) return (uint8_t)e.byte;
  return (uint8_t)("acgt"[(hap[i >> 4] >> (2 * (i & 15))) & 3u]);
}

template <class SP, class EP, class HP>
MG_NI void mg_patch_exc(typename SP::ptr seq, EP exc, int n_exc, HP hap, uint32_t x, int L, int strand) {
  for (int k = mg_exc_first(exc, n_exc, x); k < n_exc; k++) {
    MgExc e = exc[k];
    if ((uint64_t)e.start >= (uint64_t)x + L) break;
    uint64_t a = e.start > x ? e.start : x;
    uint64_t b = (uint64_t)e.start + e.len < (uint64_t)x + L ? (uint64_t)e.start + e.len : (uint64_t)x + L;
    for (uint64_t i = a; i < b; i++) {
      int idx = (int)(i - x);
      SP::st8(seq + (uint32_t)(strand ? (L - 1 - idx) : idx), mg_exc_byte(e, hap, i));
    }
  }
}

// ------------------------------------------------------------------------------------------
// Corruption of one base call (illumina.py:155-160)

MG_HD uint8_t mg_base_rot(uint8_t c, int r) {  // illumina.py:131-136: A->CTG C->ATG T->ACG G->ACT else N
  uint32_t t;
  switch (c) {
    case 'A': t = 'C' | ('T' << 8) | ('G' << 16); break;
    case 'C': t = 'A' | ('T' << 8) | ('G' << 16); break;
    case 'T': t = 'A' | ('C' << 8) | ('G' << 16); break;
    case 'G': t = 'A' | ('C' << 8) | ('T' << 16); break;
    default: t = 'N' | ('N' << 8) | ('N' << 16); break;
  }
  return (uint8_t)(t >> (8 * r));
}

// searchsorted(row, u) side='left' over n doubles
template <class DP>
MG_HD int mg_lower_bound_f64(DP row, int n, double u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (row[mid] < u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------
// Source of a read's 2-bit codes, 16 bases per chunk.  MAXW > 0: register window (MgWin, loads
// issued up front, both strands a forward extraction); MAXW == 0: streamed from memory (any L).
template <int MAXW, class HP>
struct MgSeqSrc {
  MgWin<(MAXW > 0 ? MAXW : 2)> W;
  HP hap; uint32_t x; int L, strand;
  MG_HD void load(HP hap_, uint32_t x_, int L_, int strand_) {
    hap = hap_; x = x_; L = L_; strand = strand_;
    if constexpr (MAXW > 0) mg_win_load(W, hap, x, L, strand);
  }
  // streaming source only (MAXW == 0)
  MG_HD uint32_t codes(int c) const {
    if (strand == 0) return mg_codes16(hap, (int64_t)x + 16 * c);
    // output bases [16c, 16c+16) are the complement of forward bases [L-16c-16, L-16c) reversed
    return mg_revcomp16(mg_codes16(hap, (int64_t)x + L - 16 * (int64_t)c - 16));   // may dip 15 below x: front pad
  }
};

// for_each_chunk(S, fn): fn(codes, c) for every 16-base chunk of the read, in order.  Consumes the
// register window of S.
template <int MAXW, class HP, class FN>
MG_HD void mg_for_each_chunk(MgSeqSrc<MAXW, HP> &S, FN fn) {
  const int L = S.L;
  if constexpr (MAXW > 0) {
    MG_NOUNROLL
    for (int c = 0; 16 * c < L; c += 2) {
      fn(mg_win_codes(S.W, 0), c);
      if (16 * (c + 1) < L) fn(mg_win_codes(S.W, 1), c + 1);
      mg_win_slide2(S.W);
    }
  } else {
    for (int c = 0; 16 * c < L; c++) fn(S.codes(c), c);
  }
}

// four output bases of chunk word q (n0 = first base index) through a writer
template <class WR>
MG_HD void mg_put_bases4(WR &w, uint32_t b4, int n0, int L) {
  const uint32_t ch = mg_chars4(b4);
  if (n0 + 4 <= L) w.put_word(ch);
  else for (int j = 0; n0 + j < L; j++) w.put((uint8_t)(ch >> (8 * j)));
}

template <class WR, int MAXW, class HP>
MG_HD void mg_emit_seq_src(WR &w, MgSeqSrc<MAXW, HP> &S) {
  const int L = S.L;
  mg_for_each_chunk(S, [&](uint32_t codes, int c) {
    if (16 * c + 16 <= L) {              // a whole chunk: four words, no bounds logic
      MG_NOUNROLL
      for (int q = 0; q < 4; q++, codes >>= 8) w.put_word(mg_chars4(codes & 0xFFu));
      return;
    }
    MG_NOUNROLL
    for (int q = 0; q < 4; q++) if (16 * c + 4 * q < L) mg_put_bases4(w, (codes >> (8 * q)) & 0xFFu, 16 * c + 4 * q, L);
  });
}

// One FASTQ record (fastq_lines, readgenerate.py:227-230):  qname \n SEQ \n+\n ~~~~ \n
// S = the read that goes into this file (already loaded); first/second give the qname's file order.
// qlen = length of the qname line without its newline (known from the sizing pass).
template <class SP, int MAXW, class NP, class HP, class EP>
MG_HD void mg_emit_record(typename SP::ptr dst, uint32_t qlen, const uint8_t *prefix, int prefix_len, uint64_t cnt,
                          const uint8_t *mid, int mid_len, NP nodes, MgReadRef first, MgReadRef second,
                          MgSeqSrc<MAXW, HP> &S, EP exc, int n_exc) {
  const int L = S.L;
  MgWordStream<SP> ws;
  ws.begin_rmw(mg_qname_bytes<SP>(dst, prefix, prefix_len, cnt, mid, mid_len, nodes, first, second, L));
  mg_emit_seq_src(ws, S);
  ws.put('\n'); ws.put('+'); ws.put('\n');
  mg_emit_fill(ws, '~', L);
  ws.put('\n');
  ws.end();
  if (n_exc) mg_patch_exc<SP>(dst + (qlen + 1), exc, n_exc, S.hap, S.x, L, S.strand);
}

// The other file's record has the same qname, the same offsets and (for perfect reads) the same
// quality line: only the L sequence bytes are rewritten in place.
template <class SP, int MAXW, class HP, class EP>
MG_HD void mg_rewrite_seq(typename SP::ptr seq_dst, MgSeqSrc<MAXW, HP> &S, EP exc, int n_exc) {
  MgWordStream<SP> ws;
  ws.begin_rmw(seq_dst);
  mg_emit_seq_src(ws, S);
  ws.end();
  if (n_exc) mg_patch_exc<SP>(seq_dst, exc, n_exc, S.hap, S.x, S.L, S.strand);
}

// qname line + the three separator newlines of a record whose SEQ / QUAL lines are written by
// mg_emit_seq_corrupt (fused corruption).
template <class SP, class NP>
MG_HD void mg_emit_frame(typename SP::ptr dst, uint32_t qlen, const uint8_t *prefix, int prefix_len, uint64_t cnt,
                         const uint8_t *mid, int mid_len, NP nodes, MgReadRef first, MgReadRef second, int L) {
  mg_qname_bytes<SP>(dst, prefix, prefix_len, cnt, mid, mid_len, nodes, first, second, L);
  const typename SP::ptr p = dst + (qlen + 1 + (uint32_t)L);
  SP::st8(p, '\n'); SP::st8(p + 1, '+'); SP::st8(p + 2, '\n'); SP::st8(p + (3 + (uint32_t)L), '\n');
}

// ------------------------------------------------------------------------------------------
// Production-mode corruption (Philox draws, alias-method quality sampling).
//
// The reference draws a quality per cycle and then a miscall with probability phred_p[quality]
// (illumina.py:151-160).  The same joint distribution is sampled in the other order, which takes
// the table lookup out of the miscall decision: per (file, cycle) the model gives
//     perr = sum_q P(q) phred_p[q],      P(q | miscall) = P(q) phred_p[q] / perr,
//                                        P(q | correct) = P(q) (1 - phred_p[q]) / (1 - perr)
// so the miscall is decided from a per-cycle threshold the whole warp shares, and only the
// quality comes from an alias row -- the row of (file, cycle, miscall).
//
// Draw layout (the specification tests/philox_ref.py restates in numpy): for template serial s
// (0-based count within the unit), file f and cycle pair q = n / 2,
//     r = Philox4x32-7(counter = (s, f, q, MG_STREAM_CORRUPT), key = (k0, k1))
// cycle 2q uses (r[0], r[1]), cycle 2q+1 uses (r[2], r[3]) as (w_bq, w_call):
//     T = thr[f][n] = min(2^32-1, floor(perr * 2^32));   miscall iff w_call < T
//     substituted base = base_rot[base][(w_call >= T/3) + (w_call >= floor(2T/3))]   (illumina.py:131-136,160;
//     given a miscall, w_call is uniform on [0, T), so no third draw is needed)
//     idx = w_bq >> (32 - kshift); frac = (w_bq << kshift) >> 8        (24 bits)
//     e = alias[((f * n_cycles + n) * 2 + miscall) << kshift | idx]  (entry = prob24 << 8 | alias);
//     bq = frac < (e >> 8) ? idx : (e & 255), evaluated as (w_bq << kshift) < (e & ~255)
struct MgCorruptCtx {
  const uint32_t *alias;   // [n_mates][n_cycles][2][1 << kshift]
  const uint32_t *thr;     // [n_mates][n_cycles] miscall thresholds
  int kshift, n_cycles, n_mates;
  uint32_t k0, k1;
  uint32_t thr_s, lp;      // device hot path: shared-window address of the staged thresholds, planes [T][T/3][2T/3], each [file][lp]
};

MG_HD uint32_t mg_ctz4(uint32_t m) {   // index of the lowest set bit of a non-zero 4-bit mask
#if defined(__CUDA_ARCH__)
  return (uint32_t)__ffs((int)m) - 1u;
#else
  return (m & 1u) ? 0u : (m & 2u) ? 1u : (m & 4u) ? 2u : 3u;
#endif
}

// which of the three alternatives: w_call is uniform on [0, thr) given a miscall
MG_HD uint32_t mg_sub_index(uint32_t w_call, uint32_t thr) {
  const uint32_t q = thr / 3u, r = thr - 3u * q;          // floor(2 thr / 3) = 2 q + (r == 2), without 64-bit arithmetic
  return (uint32_t)(w_call >= q) + (uint32_t)(w_call >= 2u * q + (r >> 1));
}

// one base -> bit 2 = substitution happened, bits 0-1 = which of the three alternatives; qual = ASCII quality
MG_HD uint32_t mg_corrupt_draw(const MgCorruptCtx &C, uint32_t f, int n, uint32_t w_bq, uint32_t w_call, uint32_t &qual) {
  const uint32_t cyc = f * (uint32_t)C.n_cycles + (uint32_t)n;
  const uint32_t T = C.thr[cyc];
  const uint32_t miss = (uint32_t)(w_call < T);
  const uint32_t idx = w_bq >> (32 - C.kshift);
  const uint32_t e = C.alias[((2u * cyc + miss) << C.kshift) | idx];
  qual = ((w_bq << C.kshift) < (e & 0xFFFFFF00u) ? idx : (e & 255u)) + 33u;
  return miss ? (4u | mg_sub_index(w_call, T)) : 0u;
}

MG_HD void mg_corrupt_one(const MgCorruptCtx &C, uint32_t f, int n, uint32_t w_bq, uint32_t w_call, uint32_t &base, uint32_t &qual) {
  const uint32_t d = mg_corrupt_draw(C, f, n, w_bq, w_call, qual);
  if (d) base = mg_base_rot((uint8_t)base, (int)(d & 3u));
}

// base_rot on 2-bit codes (A=0 C=1 G=2 T=3): A->CTG, C->ATG, G->ACT, T->ACG as 2-bit triples
#define MG_ROT_TBL (45u | (44u << 6) | (52u << 12) | (36u << 18))

// four bases (chunk word q, first base index n0) corrupted on the 2-bit codes -> new codes + ASCII
// qualities, in three steps so that everything that does not need the alias rows runs while their
// loads are in flight: draw (Philox, miscall bits, loads issued), subst (the miscalled codes are
// replaced; the caller also writes the previous group's output here), qual (the loaded entries
// become quality bytes).  Cycles >= L read a valid row and are masked out.
struct MgDraw4 { uint32_t wb[4], wc[4], e[4], T[4], miss; };

template <bool FULL, bool ES, int KS>   // FULL: all four cycles are inside the read (n0 + 4 <= L); ES: thresholds staged in
                                        // shared memory; KS: kshift known at compile time (0 = read it from the context)
MG_HD void mg_corrupt4_draw(const MgCorruptCtx &C, uint32_t serial, uint32_t f, int n0, int L, MgDraw4 &D) {
  const uint32_t cyc = f * (uint32_t)C.n_cycles + (uint32_t)n0;
  bool staged = false;
#if defined(__CUDA_ARCH__)
  if constexpr (ES) {   // one 16-byte load: the four cycles' thresholds (n0 is a multiple of 4, lp too)
    asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(D.T[0]), "=r"(D.T[1]), "=r"(D.T[2]), "=r"(D.T[3]) : "r"(C.thr_s + 4u * (f * C.lp + (uint32_t)n0)));
    staged = true;
  }
#endif
  if (!staged) {
    MG_UNROLL
    for (int j = 0; j < 4; j++) D.T[j] = C.thr[cyc + ((FULL || n0 + j < L) ? (uint32_t)j : 0u)];
  }
  const MgPhilox r0 = mg_philox_corrupt(serial, f, (uint32_t)(n0 >> 1), C.k0, C.k1);
  const MgPhilox r1 = mg_philox_corrupt(serial, f, (uint32_t)(n0 >> 1) + 1u, C.k0, C.k1);
  D.wb[0] = r0.v[0]; D.wb[1] = r0.v[2]; D.wb[2] = r1.v[0]; D.wb[3] = r1.v[2];
  D.wc[0] = r0.v[1]; D.wc[1] = r0.v[3]; D.wc[2] = r1.v[1]; D.wc[3] = r1.v[3];
  const uint32_t ks = KS ? (uint32_t)KS : (uint32_t)C.kshift, K = 1u << ks;
  D.miss = 0;
  MG_UNROLL
  for (int j = 0; j < 4; j++) {
    const bool in = FULL || n0 + j < L;
    const uint32_t nj = in ? (uint32_t)j : 0u;            // stay inside the table at the read's end
    const bool m = D.wc[j] < D.T[j];
    // (row of (cycle, correct) << ks) | idx in one funnel shift, + K for the miscall row; 32-bit index
    // arithmetic, so the address is one IMAD.WIDE
    uint32_t ix = mg_funnel_l(D.wb[j], 2u * (cyc + nj), ks);
    if (m) ix += K;
    D.e[j] = C.alias[ix];
    D.miss |= (in && m ? 1u : 0u) << j;
  }
}

// substitutions (2-7 % of the bases with the shipped models): each lane walks ITS OWN miscall bits, so
// the warp runs this loop as often as its worst lane has miscalls among the four bases.  In the
// emit kernel the thirds of the thresholds are staged next to them ([T][T/3][2T/3] planes), so the
// loop only selects its w_call.
template <bool ES>
MG_HD void mg_corrupt4_subst(const MgCorruptCtx &C, const MgDraw4 &D, uint32_t f, int n0, uint32_t &b4) {
  uint32_t any = D.miss;
  while (any) {
    const uint32_t j = mg_ctz4(any);
    any &= any - 1u;
    const uint32_t w = j == 0 ? D.wc[0] : j == 1 ? D.wc[1] : j == 2 ? D.wc[2] : D.wc[3];
    uint32_t sub;
    bool staged = false;
#if defined(__CUDA_ARCH__)
    if constexpr (ES) {
      uint32_t t1, t2;
      const uint32_t a = C.thr_s + 4u * ((2u + f) * C.lp + (uint32_t)n0 + j);
      asm("ld.shared.b32 %0, [%1];" : "=r"(t1) : "r"(a));
      asm("ld.shared.b32 %0, [%1];" : "=r"(t2) : "r"(a + 8u * C.lp));
      sub = (uint32_t)(w >= t1) + (uint32_t)(w >= t2);
      staged = true;
    }
#endif
    if (!staged) sub = mg_sub_index(w, j == 0 ? D.T[0] : j == 1 ? D.T[1] : j == 2 ? D.T[2] : D.T[3]);
    const uint32_t code = (b4 >> (2u * j)) & 3u;
    const uint32_t nc = (MG_ROT_TBL >> (6u * code + 2u * sub)) & 3u;
    b4 ^= (code ^ nc) << (2u * j);
  }
}

template <bool FULL, int KS>
MG_HD uint32_t mg_corrupt4_qual(const MgCorruptCtx &C, const MgDraw4 &D, int n0, int L) {
  const uint32_t ks = KS ? (uint32_t)KS : (uint32_t)C.kshift;
  uint32_t bq[4];
  MG_UNROLL
  for (int j = 0; j < 4; j++) {
    bq[j] = (D.wb[j] << ks) < (D.e[j] & 0xFFFFFF00u) ? (D.wb[j] >> (32u - ks)) : (D.e[j] & 255u);   // frac24 < prob24
  }
  if (FULL) return (bq[0] + (bq[1] << 8)) + ((bq[2] + (bq[3] << 8)) << 16) + 0x21212121u;
  uint32_t qw = 0;
  MG_UNROLL
  for (int j = 0; j < 4; j++) if (n0 + j < L) qw |= (bq[j] + 33u) << (8 * j);
  return qw;
}

// SEQ and QUAL lines of one read, corrupted on the fly: every thread of a warp is at the same
// cycle of its own record, so the alias row (one 256-byte line pair) is shared by the warp.
template <class SP, int KS = 0, int MAXW, class HP, class EP>
MG_HD void mg_emit_seq_corrupt(typename SP::ptr seq_dst, typename SP::ptr qual_dst, MgSeqSrc<MAXW, HP> &S, EP exc, int n_exc,
                               const MgCorruptCtx &C, uint32_t serial, uint32_t f) {
  const int L = S.L;
  const struct { uint32_t x; int strand; HP hap; } mine = {S.x, S.strand, S.hap};
  constexpr bool ES = !SP::is_generic;   // staged in shared memory <=> running in k_unit_emit with the threshold table staged too
  MgWordStream<SP> ws, wq;
  ws.begin_rmw(seq_dst); wq.begin_rmw(qual_dst);
  // the output of a group is written one group late, between the next group's table loads and
  // their first use, together with this group's substitutions: independent work under the L2 latency
  uint32_t pb4 = 0, pqw = 0;
  bool pend = false;
  mg_for_each_chunk(S, [&](uint32_t codes, int c) {
    if (16 * c + 16 <= L) {              // a whole chunk: four full groups, no per-group bounds logic
      int n0 = 16 * c;
      MG_NOUNROLL
      for (int q = 0; q < 4; q++, n0 += 4, codes >>= 8) {
        uint32_t b4 = codes & 0xFFu;
        MgDraw4 D;
        mg_corrupt4_draw<true, ES, KS>(C, serial, f, n0, L, D);
        if (pend) { ws.put_word(mg_chars4(pb4)); wq.put_word(pqw); }
        mg_corrupt4_subst<ES>(C, D, f, n0, b4);
        pb4 = b4; pqw = mg_corrupt4_qual<true, KS>(C, D, n0, L); pend = true;
      }
      return;
    }
    MG_NOUNROLL
    for (int q = 0; q < 4; q++) {        // the last, partial chunk
      const int n0 = 16 * c + 4 * q;
      if (n0 < L) {
        uint32_t b4 = (codes >> (8 * q)) & 0xFFu, qw;
        MgDraw4 D;
        if (n0 + 4 <= L) {
          mg_corrupt4_draw<true, ES, KS>(C, serial, f, n0, L, D);
          if (pend) { ws.put_word(mg_chars4(pb4)); wq.put_word(pqw); }
          mg_corrupt4_subst<ES>(C, D, f, n0, b4);
          pb4 = b4; pqw = mg_corrupt4_qual<true, KS>(C, D, n0, L); pend = true;
        } else {
          if (pend) { ws.put_word(mg_chars4(pb4)); wq.put_word(pqw); pend = false; }
          mg_corrupt4_draw<false, ES, KS>(C, serial, f, n0, L, D);
          mg_corrupt4_subst<ES>(C, D, f, n0, b4);
          qw = mg_corrupt4_qual<false, KS>(C, D, n0, L);
          const uint32_t ch = mg_chars4(b4);
          for (int j = 0; n0 + j < L; j++) { ws.put((uint8_t)(ch >> (8 * j))); wq.put((uint8_t)(qw >> (8 * j))); }
        }
      }
    }
  });
  if (pend) { ws.put_word(mg_chars4(pb4)); wq.put_word(pqw); }
  ws.end(); wq.end();
  if (n_exc) {
    // bases in exception runs: the reference substitutes 'N' for any non-ACGT base on an error
    // (base_rot.get(base, 'NNN'), illumina.py:160) and leaves it alone otherwise
    for (int k = mg_exc_first(exc, n_exc, mine.x); k < n_exc; k++) {
      const MgExc e = exc[k];
      if ((uint64_t)e.start >= (uint64_t)mine.x + L) break;
      uint64_t a = e.start > mine.x ? e.start : mine.x;
      uint64_t b = (uint64_t)e.start + e.len < (uint64_t)mine.x + L ? (uint64_t)e.start + e.len : (uint64_t)mine.x + L;
      for (uint64_t i = a; i < b; i++) {
        const int idx = (int)(i - mine.x), n = mine.strand ? (L - 1 - idx) : idx;
        const MgPhilox r = mg_philox_corrupt(serial, f, (uint32_t)(n >> 1), C.k0, C.k1);
        uint32_t base = mg_exc_byte(e, mine.hap, i), qual;
        mg_corrupt_one(C, f, n, (n & 1) ? r.v[2] : r.v[0], (n & 1) ? r.v[3] : r.v[1], base, qual);
        SP::st8(seq_dst + (uint32_t)n, (uint8_t)base);
      }
    }
  }
}

// One base call with explicit draws (deterministic mode: the reference's own numpy draws)
template <class DP>
MG_HD void mg_corrupt_call(uint8_t *seq, uint8_t *qual, int n, DP cum_row, int n_bq, DP phred,
                           double u_bq, double u_call, int rot) {
  int bq = mg_lower_bound_f64(cum_row, n_bq, u_bq);                // illumina.py:156
  if (bq > 93) bq = 93;
  if (u_call < phred[bq]) seq[n] = mg_base_rot(seq[n], rot);       // illumina.py:159-160
  qual[n] = (uint8_t)(bq + 33);
}

