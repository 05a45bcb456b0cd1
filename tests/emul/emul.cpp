// TEST HARNESS: runs the engine's per-thread device logic (mitty_b200/csrc/mg_core.cuh, the very
// header the kernels include) on the CPU, sequentially, so the build box (no GPU) can check
// formatting / lookup / revcomp / placement arithmetic against the oracle.  Not part of the product
// library and never shipped: tests/test_emul_core.py compiles it with g++ on the fly.
#include <cstring>
#include "../../mitty_b200/csrc/mg_core.cuh"

extern "C" {

// mirrors k_unit_emit phase 1 + phase 2 for explicit templates (ts_rel, tl), one template at a time
int64_t emul_unit(const uint32_t *hap, uint32_t hap_len, const MgNode *nodes, int n_nodes, const uint32_t *blk,
                  int blk_shift, int n_blk, const MgExc *exc, int n_exc, int L, int64_t n, const int64_t *ts_rel,
                  const int64_t *tl_in, const int8_t *fo_in, const char *prefix, const char *mid, uint8_t *out1,
                  uint8_t *out2, int64_t cap, int64_t *n_bytes) {
  const int pl = (int)strlen(prefix), ml = (int)strlen(mid);
  uint64_t sz_sum = 0, cnt1 = 0, cnt2 = 0;
  for (int64_t j = 0; j < n; j++) {
    int64_t tl = tl_in[j] < L ? L : tl_in[j];
    int64_t te = ts_rel[j] + tl;
    if (!(te < (int64_t)hap_len && ts_rel[j] >= 0)) continue;
    uint32_t fo = (uint32_t)(fo_in[cnt1] & 1);
    cnt1++;
    uint32_t xa = (uint32_t)ts_rel[j], xb = (uint32_t)(te - L);
    if (n_exc && !(mg_count_N(exc, n_exc, xa, L) <= 2 && mg_count_N(exc, n_exc, xb, L) <= 2)) continue;
    MgReadRef ra = {xa, mg_find_node(nodes, blk, blk_shift, n_blk, n_nodes, xa), mg_find_node(nodes, blk, blk_shift, n_blk, n_nodes, xa + L - 1), 0};
    MgReadRef rb = {xb, mg_find_node(nodes, blk, blk_shift, n_blk, n_nodes, xb), mg_find_node(nodes, blk, blk_shift, n_blk, n_nodes, xb + L - 1), 1};
    MgCountWriter cw; cw.n = 0;
    mg_fmt_read(cw, nodes, ra.n0, ra.n1, ra.x, L, 0);
    mg_fmt_read(cw, nodes, rb.n0, rb.n1, rb.x, L, 1);
    uint32_t sz = (uint32_t)(pl + ml) + cw.n + 2u * (uint32_t)L + 5u;
    uint64_t cnt = cnt2 + 1;
    uint64_t off = sz_sum + mg_digit_sum(cnt2);         // the kernel's placement formula
    uint32_t qlen = sz + (uint32_t)mg_ndigits(cnt) - (2u * (uint32_t)L + 5u);
    uint64_t rec = sz + (uint64_t)mg_ndigits(cnt);
    if ((int64_t)(off + rec) > cap) return -1;
    MgReadRef first = fo ? rb : ra, second = fo ? ra : rb;
    mg_emit_record(out1 + off, qlen, (const uint8_t *)prefix, pl, cnt, (const uint8_t *)mid, ml, nodes, first, second, first, L, hap, exc, n_exc);
    mg_emit_record(out2 + off, qlen, (const uint8_t *)prefix, pl, cnt, (const uint8_t *)mid, ml, nodes, first, second, second, L, hap, exc, n_exc);
    sz_sum += sz; cnt2++;
  }
  *n_bytes = (int64_t)(sz_sum + mg_digit_sum(cnt2));
  return (int64_t)cnt2;
}

void emul_permute(uint32_t n, uint32_t half_bits, uint32_t k0, uint32_t k1, uint32_t *out) {
  for (uint32_t i = 0; i < n; i++) out[i] = mg_permute(i, n, half_bits, k0, k1);
}

void emul_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
  MgPhilox r = mg_philox(c0, c1, c2, c3, k0, k1);
  for (int i = 0; i < 4; i++) out[i] = r.v[i];
}

uint64_t emul_digit_sum(uint64_t m) { return mg_digit_sum(m); }

// deterministic-mode corruption of one read in place (seq, qual of length L)
void emul_corrupt_det(uint8_t *seq, uint8_t *qual, int L, const double *cum_rows, int n_bq, const double *phred,
                      const double *bq_rnd, const double *call_rnd, const uint8_t *base_rnd) {
  for (int n = 0; n < L; n++)
    mg_corrupt_call(seq, qual, n, cum_rows + (size_t)n * n_bq, n_bq, phred, bq_rnd[n], call_rnd[n], (int)base_rnd[n]);
}
}
