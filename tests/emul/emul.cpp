// TEST HARNESS: runs the engine's per-thread device logic (mitty_b200/csrc/mg_core.cuh, the very
// header the kernels include) on the CPU, sequentially, so the build box (no GPU) can check
// formatting / lookup / revcomp / placement arithmetic against the oracle.  Not part of the product
// library and never shipped: tests/test_emul_core.py compiles it with g++ on the fly.
#include <cstring>
#include "../../mitty_b200/csrc/mg_core.cuh"

// mirrors k_unit_emit phase 1 + phase 2 for explicit templates (ts_rel, tl), one template at a time.
// Like a staged tile of the kernel: the file-0 record is written into a stage buffer and copied
// out, then only the sequence (and, when corrupting, the quality) bytes are rewritten for file 1.
template <int MAXW>
static int64_t emul_unit_t(const uint32_t *hap, uint32_t hap_len, const MgNode *nodes, int n_nodes, const uint32_t *blk,
                  int blk_shift, int n_blk, const MgExc *exc, int n_exc, int L, int64_t n, const int64_t *ts_rel,
                  const int64_t *tl_in, const int8_t *fo_in, const char *prefix, const char *mid, uint8_t *out1,
                  uint8_t *out2, int64_t cap, int64_t *n_bytes,
                  int corrupt, const uint32_t *alias, int kshift, int code9, int n_cycles, uint32_t k0, uint32_t k1) {
  const int pl = (int)strlen(prefix), ml = (int)strlen(mid);
  const int L_nd = mg_ndigits32((uint32_t)L);
  MgCorruptCtx cor; cor.alias = alias; cor.kshift = kshift; cor.code9 = code9; cor.n_cycles = n_cycles; cor.n_mates = 2; cor.k0 = k0; cor.k1 = k1;
  MgQnConst Q;
  mg_qn_const(Q, (const uint8_t *)prefix, pl, (const uint8_t *)mid, ml, L);
  uint64_t sz_sum = 0, cnt1 = 0, cnt2 = 0;
  static uint8_t stage_raw[1 << 16];
  for (int64_t j = 0; j < n; j++) {
    int64_t tl = tl_in[j] < L ? L : tl_in[j];
    int64_t te = ts_rel[j] + tl;
    if (!(te < (int64_t)hap_len && ts_rel[j] >= 0)) continue;
    uint32_t fo = (uint32_t)(fo_in[cnt1] & 1);
    cnt1++;
    uint32_t xa = (uint32_t)ts_rel[j], xb = (uint32_t)(te - L);
    bool ta = false, tb = false;
    if (n_exc && !(mg_count_N(exc, n_exc, xa, L, ta) <= 2 && mg_count_N(exc, n_exc, xb, L, tb) <= 2)) continue;
    MgReadRef ra = {xa, mg_find_node(nodes, blk, blk_shift, n_blk, n_nodes, xa), 0, 0};
    MgReadRef rb = {xb, mg_find_node(nodes, blk, blk_shift, n_blk, n_nodes, xb), 0, 1};
    ra.n1 = mg_last_node(nodes, ra.n0, n_nodes, xa, L);
    rb.n1 = mg_last_node(nodes, rb.n0, n_nodes, xb, L);
    if (ra.n1 != mg_find_node(nodes, blk, blk_shift, n_blk, n_nodes, xa + L - 1)) return -2;
    if (rb.n1 != mg_find_node(nodes, blk, blk_shift, n_blk, n_nodes, xb + L - 1)) return -2;
    MgCountWriter cw; cw.n = 0;
    mg_fmt_read(cw, nodes, ra.n0, ra.n1, ra.x, L, 0);
    mg_fmt_read(cw, nodes, rb.n0, rb.n1, rb.x, L, 1);
    if (cw.n != mg_read_fields_len(nodes, ra.n0, ra.n1, xa, L, L_nd) + mg_read_fields_len(nodes, rb.n0, rb.n1, xb, L, L_nd)) return -3;
    uint32_t sz = (uint32_t)(pl + ml) + cw.n + 2u * (uint32_t)L + 5u;
    uint64_t cnt = cnt2 + 1;
    uint64_t off = sz_sum + mg_digit_sum(cnt2);         // the kernel's placement formula
    uint32_t qlen = sz + (uint32_t)mg_ndigits(cnt) - (2u * (uint32_t)L + 5u);
    uint64_t rec = sz + (uint64_t)mg_ndigits(cnt);
    if ((int64_t)(off + rec) > cap || rec + 8 > sizeof stage_raw) return -1;
    MgReadRef first = fo ? rb : ra, second = fo ? ra : rb;
    const int ne_first = (fo ? tb : ta) ? n_exc : 0, ne_second = (fo ? ta : tb) ? n_exc : 0;   // the plan's flags: patch only reads that touch a run
    uint8_t *dst = stage_raw + (off & 3);                // same word phase as the final destination
    MgSeqSrc<MAXW, const uint32_t *> S;
    S.load(hap, first.x, L, first.strand);
    if (corrupt) {
      mg_emit_frame_qname<MgGenericSpace>(dst, Q, Q, (uint32_t)cnt, nodes, first, second, L);
      mg_emit_frame_seps<MgGenericSpace>(dst, qlen, L);
      if (code9) mg_emit_seq_corrupt<MgGenericSpace, true>(dst + qlen + 1, dst + qlen + 1 + L + 3, S, exc, ne_first, cor, (uint32_t)(cnt - 1), 0u, 0u);
      else mg_emit_seq_corrupt<MgGenericSpace, false>(dst + qlen + 1, dst + qlen + 1 + L + 3, S, exc, ne_first, cor, (uint32_t)(cnt - 1), 0u, 0u);
      memcpy(out1 + off, dst, rec);
      S.load(hap, second.x, L, second.strand);
      if (code9) mg_emit_seq_corrupt<MgGenericSpace, true>(dst + qlen + 1, dst + qlen + 1 + L + 3, S, exc, ne_second, cor, (uint32_t)(cnt - 1), 1u, 1u);
      else mg_emit_seq_corrupt<MgGenericSpace, false>(dst + qlen + 1, dst + qlen + 1 + L + 3, S, exc, ne_second, cor, (uint32_t)(cnt - 1), 1u, 1u);
      memcpy(out2 + off, dst, rec);
    } else {
      MgStream<MgGenericSpace> ws;
      mg_emit_record<MgGenericSpace>(ws, dst, Q, Q, (uint32_t)cnt, nodes, first, second, S);
      ws.end();
      if (ne_first) mg_patch_exc<MgGenericSpace>(dst + (qlen + 1), exc, ne_first, S.hap, S.x, L, S.strand);
      memcpy(out1 + off, dst, rec);
      S.load(hap, second.x, L, second.strand);
      mg_rewrite_seq<MgGenericSpace>(dst + qlen + 1, S, exc, ne_second);
      memcpy(out2 + off, dst, rec);
    }
    sz_sum += sz; cnt2++;
  }
  *n_bytes = (int64_t)(sz_sum + mg_digit_sum(cnt2));
  return (int64_t)cnt2;
}

extern "C" {

// maxw selects the code-source variant the kernel would use: 12 / 21 register windows, 0 streaming
int64_t emul_unit(const uint32_t *hap, uint32_t hap_len, const MgNode *nodes, int n_nodes, const uint32_t *blk,
                  int blk_shift, int n_blk, const MgExc *exc, int n_exc, int L, int64_t n, const int64_t *ts_rel,
                  const int64_t *tl_in, const int8_t *fo_in, const char *prefix, const char *mid, uint8_t *out1,
                  uint8_t *out2, int64_t cap, int64_t *n_bytes,
                  int corrupt, const uint32_t *alias, int kshift, int code9, int n_cycles, uint32_t k0, uint32_t k1, int maxw) {
#define EMUL_ARGS hap, hap_len, nodes, n_nodes, blk, blk_shift, n_blk, exc, n_exc, L, n, ts_rel, tl_in, fo_in, prefix, mid, out1, out2, cap, n_bytes, corrupt, alias, kshift, code9, n_cycles, k0, k1
  if (maxw == 12 && L <= 161) return emul_unit_t<12>(EMUL_ARGS);
  if (maxw == 21 && L <= 305) return emul_unit_t<21>(EMUL_ARGS);
  if (maxw == 0) return emul_unit_t<0>(EMUL_ARGS);
  return -9;
}

void emul_permute(uint32_t n, uint32_t k0, uint32_t k1, uint32_t *out) {
  for (uint32_t i = 0; i < n; i++) out[i] = mg_permute(i, n, mg_perm_bits(n), k0, k1);
}

void emul_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
  MgPhilox r = mg_philox(c0, c1, c2, c3, k0, k1);
  for (int i = 0; i < 4; i++) out[i] = r.v[i];
}

uint64_t emul_digit_sum(uint64_t m) { return mg_digit_sum(m); }

// pre (npre bytes) + decimal digits of v through the token path -> bytes written
int emul_put_num(uint32_t v, uint32_t pre, uint32_t npre, int phase, uint8_t *out, int small) {
  uint8_t buf[64];
  memset(buf, 0xEE, sizeof buf);
  MgStream<MgGenericSpace> w;
  w.begin(buf + 8 + phase);
  if (small) mg_put_small(w, v, pre, npre); else mg_put_num(w, v, pre, npre);
  uint8_t *wp = w.wp; uint32_t sh = w.sh;
  w.end();
  const int n = (int)(wp - (buf + 8 + phase)) + (int)(sh >> 3);   // wp is word aligned: negative part = the start's own phase
  for (int i = 0; i < 8 + phase; i++) if (buf[i] != 0xEE) return -1;        // nothing before the start may be touched
  for (int i = 8 + phase + n; i < 64; i++) if (buf[i] != 0xEE) return -2;   // nor after the end
  memcpy(out, buf + 8 + phase, n);
  return n;
}

// deterministic-mode corruption of one read in place (seq, qual of length L)
void emul_corrupt_det(uint8_t *seq, uint8_t *qual, int L, const double *cum_rows, int n_bq, const double *phred,
                      const double *bq_rnd, const double *call_rnd, const uint8_t *base_rnd) {
  for (int n = 0; n < L; n++)
    mg_corrupt_call(seq, qual, n, cum_rows + (size_t)n * n_bq, n_bq, phred, bq_rnd[n], call_rnd[n], (int)base_rnd[n]);
}
}
