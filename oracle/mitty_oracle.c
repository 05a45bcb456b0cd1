/* TEST INFRASTRUCTURE (CPU oracle) -- not part of the product path.
 *
 * Plain-C restatement of Mitty's read-generation / read-corruption hot path, following the
 * reference function by function (citations are /root/reference paths).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Pinned by: tests/test_oracle_golden.py (every KAT of the reference's own test_rpc.py /
 * test_vcfio.py, plus golden FASTQ produced by running the UNMODIFIED reference in the build
 * container, tests/golden/make_golden.py) and tests/test_oracle_rng.py (numpy draw recipes).
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "mt19937.h"

#define SEED_MAX 0xFFFFFFFFu  /* (1<<32)-1, readgenerate.py:54, illumina.py:9 */

/* ------------------------------------------------------------------------------------------ */
/* RNG recipes exported for the pinning tests                                                  */

void orc_rand(uint32_t seed, double *out, int64_t n) {
  mt_state s; mt_seed(&s, seed);
  for (int64_t i = 0; i < n; i++) out[i] = mt_double(&s);
}
void orc_randint(uint32_t seed, uint32_t high_exclusive, int64_t *out, int64_t n) {
  mt_state s; mt_seed(&s, seed);
  for (int64_t i = 0; i < n; i++) out[i] = mt_bounded(&s, high_exclusive - 1);
}
void orc_geometric(uint32_t seed, double p, int64_t *out, int64_t n) {
  mt_state s; mt_seed(&s, seed);
  for (int64_t i = 0; i < n; i++) out[i] = mt_geometric(&s, p);
}
void orc_bits_i8(uint32_t seed, int8_t *out, int64_t n) {
  mt_state s; mt_seed(&s, seed); mt_bits_i8(&s, out, n);
}
void orc_shuffle_i64(uint32_t seed, int64_t *x, int64_t n) {
  mt_state s; mt_seed(&s, seed); mt_shuffle_i64(&s, x, n);
}

/* ------------------------------------------------------------------------------------------ */
/* a6: get_data_for_workers, readgenerate.py:129-159                                           */
/* out_seed[k], k in nested (region, copy, pass) order; out_order = shuffled index list        */
void orc_unit_schedule(uint32_t seed, int64_t n_units, uint32_t *out_seed, int64_t *out_order) {
  mt_state s; mt_seed(&s, seed);
  uint32_t shuffle_seed = mt_bounded(&s, SEED_MAX - 1);
  for (int64_t k = 0; k < n_units; k++) { out_seed[k] = mt_bounded(&s, SEED_MAX - 1); out_order[k] = k; }
  mt_state sh; mt_seed(&sh, shuffle_seed);
  mt_shuffle_i64(&sh, out_order, n_units);
}

/* a8: generate_reads seed split, illumina.py:56-58 */
void orc_unit_seeds(uint32_t seed, uint32_t *out4) {
  mt_state s; mt_seed(&s, seed);
  for (int k = 0; k < 4; k++) out4[k] = mt_bounded(&s, SEED_MAX - 1);
}

/* a16: per-worker corruption seeds, readcorrupt.py:31,36 */
void orc_corrupt_worker_seeds(uint32_t seed, int n_workers, uint32_t *out) {
  mt_state s; mt_seed(&s, seed);
  for (int k = 0; k < n_workers; k++) out[k] = mt_bounded(&s, SEED_MAX - 1);
}

/* ------------------------------------------------------------------------------------------ */
/* a7: create_node_list, rpc.py:38-116                                                         */

typedef struct {
  int64_t ps, pr;       /* 1-based sample / reference position, rpc.py:5-12 */
  char op;              /* '=', 'X', 'I', 'D' */
  int64_t oplen;
  const char *seq;      /* points into ref_seq or the alt pool */
  int64_t seqlen;
  int has_v; int64_t v; /* rpc.py:15-20 */
} onode;

typedef struct { onode *n; int64_t cnt, cap; } onode_list;

static void nl_push(onode_list *l, int64_t ps, int64_t pr, char op, int64_t oplen,
                    const char *seq, int64_t seqlen) {
  if (l->cnt == l->cap) { l->cap = l->cap ? l->cap * 2 : 1024; l->n = (onode *)realloc(l->n, l->cap * sizeof(onode)); }
  onode *x = &l->n[l->cnt++];
  x->ps = ps; x->pr = pr; x->op = op; x->oplen = oplen; x->seq = seq; x->seqlen = seqlen;
  x->has_v = (op != '=');
  x->v = (op == 'X') ? 0 : (op == 'I') ? oplen : (op == 'D') ? -oplen : 0;
}

/* Python slice ref_seq[a:b] with a,b >= 0 clamps to the string */
static void py_slice(const char *s, int64_t len, int64_t a, int64_t b, const char **o, int64_t *ol) {
  if (a > len) a = len;
  if (b > len) b = len;
  if (b < a) b = a;
  *o = s + a; *ol = b - a;
}

/* variants: pos (1-based), op ('X','I','D'), oplen, alt string = alt_pool[alt_off[i]:alt_off[i+1]] */
static void build_nodes(onode_list *l, const char *ref_seq, int64_t ref_len, int64_t ref_start_pos,
                        int64_t n_var, const int64_t *vpos, const char *vop, const int64_t *voplen,
                        const char *alt_pool, const int64_t *alt_off) {
  int64_t samp_pos = ref_start_pos, ref_pos = ref_start_pos;   /* rpc.py:48 */
  const char *sq; int64_t sl;
  for (int64_t i = 0; i < n_var; i++) {
    if (vpos[i] < ref_pos) continue;                            /* rpc.py:55 */
    const char *alt = alt_pool + alt_off[i];
    int64_t altlen = alt_off[i + 1] - alt_off[i];
    if (vop[i] == 'X') {                                        /* snp, rpc.py:75-87 */
      int64_t delta = vpos[i] - ref_pos;
      if (delta > 0) {
        py_slice(ref_seq, ref_len, ref_pos - ref_start_pos, vpos[i] - ref_start_pos, &sq, &sl);
        nl_push(l, samp_pos, ref_pos, '=', delta, sq, sl);
        ref_pos = vpos[i]; samp_pos += delta;
      }
      nl_push(l, samp_pos, ref_pos, 'X', 1, alt, altlen);
      ref_pos += 1; samp_pos += 1;
    } else if (vop[i] == 'I') {                                 /* insertion, rpc.py:90-102 */
      int64_t delta = vpos[i] + 1 - ref_pos;
      if (delta > 0) {
        py_slice(ref_seq, ref_len, ref_pos - ref_start_pos, vpos[i] + 1 - ref_start_pos, &sq, &sl);
        nl_push(l, samp_pos, ref_pos, '=', delta, sq, sl);
        samp_pos += delta;
      }
      ref_pos = vpos[i] + 1;
      nl_push(l, samp_pos, ref_pos, 'I', voplen[i], alt + (altlen > 0 ? 1 : 0), altlen > 0 ? altlen - 1 : 0);
      samp_pos += voplen[i];
    } else {                                                    /* deletion, rpc.py:105-116 */
      int64_t delta = vpos[i] + 1 - ref_pos;
      if (delta > 0) {
        py_slice(ref_seq, ref_len, ref_pos - ref_start_pos, vpos[i] + 1 - ref_start_pos, &sq, &sl);
        nl_push(l, samp_pos, ref_pos, '=', delta, sq, sl);
        samp_pos += delta;
      }
      ref_pos = vpos[i] + 1 + voplen[i];
      nl_push(l, samp_pos - 1, ref_pos, 'D', voplen[i], "", 0);
    }
  }
  int64_t offset = ref_pos - ref_start_pos;                     /* rpc.py:58-61 */
  if (offset <= ref_len)
    nl_push(l, samp_pos, ref_pos, '=', ref_len - offset, ref_seq + offset, ref_len - offset);
}

/* exported: node table as flat arrays (for the test_rpc KATs). Returns node count, or -1 if cap
 * is too small.  seq_off indexes seq_pool. */
int64_t orc_create_node_list(const char *ref_seq, int64_t ref_len, int64_t ref_start_pos,
                             int64_t n_var, const int64_t *vpos, const char *vop, const int64_t *voplen,
                             const char *alt_pool, const int64_t *alt_off,
                             int64_t cap, int64_t *ps, int64_t *pr, char *op, int64_t *oplen,
                             int64_t *v, int8_t *has_v, int64_t *seq_off, char *seq_pool, int64_t pool_cap) {
  onode_list l = {0, 0, 0};
  build_nodes(&l, ref_seq, ref_len, ref_start_pos, n_var, vpos, vop, voplen, alt_pool, alt_off);
  int64_t ret = l.cnt, o = 0;
  if (l.cnt > cap) ret = -1;
  else {
    for (int64_t i = 0; i < l.cnt; i++) {
      ps[i] = l.n[i].ps; pr[i] = l.n[i].pr; op[i] = l.n[i].op; oplen[i] = l.n[i].oplen;
      v[i] = l.n[i].v; has_v[i] = (int8_t)l.n[i].has_v; seq_off[i] = o;
      if (o + l.n[i].seqlen > pool_cap) { ret = -1; break; }
      memcpy(seq_pool + o, l.n[i].seq, l.n[i].seqlen); o += l.n[i].seqlen;
    }
    if (ret >= 0) seq_off[l.cnt] = o;
  }
  free(l.n);
  return ret;
}

/* ------------------------------------------------------------------------------------------ */
/* a11: get_begin_end_nodes, rpc.py:119-130  (searchsorted side='right' minus 1)               */
static int64_t ss_right_keys(const onode *n, int64_t cnt, int64_t x) {
  int64_t lo = 0, hi = cnt;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    int64_t key = n[mid].ps + (n[mid].op == 'D' ? 1 : 0);      /* rpc.py:127 */
    if (key <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* growable char buffer */
typedef struct { char *p; int64_t len, cap; } sbuf;
static void sb_need(sbuf *b, int64_t extra) {
  if (b->len + extra > b->cap) { while (b->len + extra > b->cap) b->cap = b->cap ? b->cap * 2 : 256; b->p = (char *)realloc(b->p, b->cap); }
}
static void sb_put(sbuf *b, const char *s, int64_t n) { sb_need(b, n); memcpy(b->p + b->len, s, n); b->len += n; }
static void sb_putc(sbuf *b, char c) { sb_need(b, 1); b->p[b->len++] = c; }
static void sb_int(sbuf *b, int64_t v) { char t[32]; int n = snprintf(t, sizeof t, "%lld", (long long)v); sb_put(b, t, n); }

/* a12: generate_read, rpc.py:133-160.  Appends to cigar / vlist / seq buffers; returns pos. */
static int64_t gen_read(int64_t p, int64_t l, int64_t n0, int64_t n1, const onode *nodes,
                        sbuf *cigar, sbuf *vlist, sbuf *seq) {
  cigar->len = vlist->len = seq->len = 0;
  int first_v = 1;
  for (int64_t k = n0; k <= n1; k++) {
    const onode *n = &nodes[k];
    if (n->has_v) {                                             /* rpc.py:144 */
      if (!first_v) sb_putc(vlist, ',');
      sb_int(vlist, n->v); first_v = 0;
    }
    int64_t a = p - n->ps; if (a < 0) a = 0;                    /* max(0, p - n.ps) */
    int64_t b = p + l - n->ps; if (n->oplen < b) b = n->oplen;  /* min(p + l - n.ps, n.oplen) */
    sb_int(cigar, n->op != 'D' ? (b - a) : n->oplen);           /* rpc.py:145 */
    sb_putc(cigar, n->op);
    /* rpc.py:146: n.seq[a:b] with Python slice clamping (b may be <= 0 only in degenerate calls) */
    int64_t sa = a, sb_ = b;
    if (sb_ < 0) { sb_ += n->seqlen; if (sb_ < 0) sb_ = 0; }
    if (sa > n->seqlen) sa = n->seqlen;
    if (sb_ > n->seqlen) sb_ = n->seqlen;
    if (sb_ > sa) sb_put(seq, n->seq + sa, sb_ - sa);
  }
  int64_t pos;
  if (nodes[n0].op == 'I') {
    if (n0 == n1) {                                             /* rpc.py:149-154 */
      pos = nodes[n0].pr - 1;
      cigar->len = 0;
      sb_putc(cigar, '>'); sb_int(cigar, p - nodes[n0].ps); sb_putc(cigar, ':'); sb_int(cigar, l); sb_putc(cigar, 'I');
    } else pos = nodes[n0].pr;                                  /* rpc.py:156 */
  } else pos = p - nodes[n0].ps + nodes[n0].pr;                 /* rpc.py:158 */
  return pos;
}

/* exported single-read form for the KATs: builds nodes, finds n0/n1 (or takes them), returns
 * pos; cigar / vlist / seq are written NUL-terminated into caller buffers of size cap. */
int64_t orc_generate_read(const char *ref_seq, int64_t ref_len, int64_t ref_start_pos,
                          int64_t n_var, const int64_t *vpos, const char *vop, const int64_t *voplen,
                          const char *alt_pool, const int64_t *alt_off,
                          int64_t p, int64_t l, int64_t *n0_io, int64_t *n1_io,
                          char *cigar_out, char *vlist_out, char *seq_out, int64_t cap) {
  onode_list nl = {0, 0, 0};
  build_nodes(&nl, ref_seq, ref_len, ref_start_pos, n_var, vpos, vop, voplen, alt_pool, alt_off);
  int64_t n0 = *n0_io, n1 = *n1_io;
  if (n0 < 0) { n0 = ss_right_keys(nl.n, nl.cnt, p) - 1; n1 = ss_right_keys(nl.n, nl.cnt, p + l - 1) - 1; }
  *n0_io = n0; *n1_io = n1;
  sbuf c = {0, 0, 0}, v = {0, 0, 0}, s = {0, 0, 0};
  int64_t pos = gen_read(p, l, n0, n1, nl.n, &c, &v, &s);
  if (c.len + 1 > cap || v.len + 1 > cap || s.len + 1 > cap) pos = -1;
  else {
    memcpy(cigar_out, c.p, c.len); cigar_out[c.len] = 0;
    memcpy(vlist_out, v.p, v.len); vlist_out[v.len] = 0;
    memcpy(seq_out, s.p, s.len); seq_out[s.len] = 0;
  }
  free(c.p); free(v.p); free(s.p); free(nl.n);
  return pos;
}

/* ------------------------------------------------------------------------------------------ */
/* a9 + a10: _templates_for_region / _reads_for_template_in_region, illumina.py:66-110          */

static int64_t ss_left_f64(const double *a, int64_t n, double x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
  return lo;
}

/* Returns kept count; ts/te/fo sized for est_block_size (returned through *n_est). */
int64_t orc_templates(double p, int64_t rlen, const double *cum_tlen, int64_t n_tlen,
                      int64_t p_min, int64_t p_max, uint32_t unit_seed,
                      int64_t cap, int64_t *ts_out, int64_t *te_out, int8_t *fo_out, int64_t *n_est) {
  uint32_t sd[4]; orc_unit_seeds(unit_seed, sd);                /* tloc, tlen, shuffle, file_order */
  mt_state tloc, tlen, shuf, ford;
  mt_seed(&tloc, sd[0]); mt_seed(&tlen, sd[1]); mt_seed(&shuf, sd[2]); mt_seed(&ford, sd[3]);
  int64_t N = (int64_t)((double)(p_max - p_min) * p * 1.2);     /* illumina.py:69 */
  if (N < 0) N = 0;
  *n_est = N;
  if (N > cap) return -1;
  int64_t *ts = (int64_t *)malloc((N ? N : 1) * sizeof(int64_t));
  int64_t acc = 0;
  for (int64_t i = 0; i < N; i++) { acc += mt_geometric(&tloc, p); ts[i] = acc + p_min + 1; }  /* :70 */
  mt_shuffle_i64(&shuf, ts, N);                                 /* :71 */
  int64_t k = 0;
  for (int64_t i = 0; i < N; i++) {
    int64_t tl = ss_left_f64(cum_tlen, n_tlen, mt_double(&tlen)); /* :72 */
    if (tl < rlen) tl = rlen;                                   /* :73 clip(min) */
    int64_t te = ts[i] + tl;
    if (te < p_max) { ts_out[k] = ts[i]; te_out[k] = te; k++; } /* :75-76 */
  }
  mt_bits_i8(&ford, fo_out, k);                                 /* :93 */
  free(ts);
  return k;
}

/* ------------------------------------------------------------------------------------------ */
/* a13 + a14: worker loop + fastq_lines, readgenerate.py:183-230                                */

static char complement(char c) {                                /* readgenerate.py:56 */
  switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return c; }
}

typedef struct { int strand; int64_t pos, len; sbuf cigar, vlist, seq; } oread;

/* Generates one work unit.  Templates are given explicitly (ts, te, fo of kept templates, i.e.
 * the output of orc_templates or of the reference's generate_reads), so the same routine serves
 * the deterministic and the explicit-template parity tests.
 * qname stub = "<sample>:<worker_id>:<ps>" (readgenerate.py:195).
 * Returns number of templates written; bytes written per file in *len1 / *len2; -1 on overflow. */
int64_t orc_generate_unit(const char *ref_seq, int64_t ref_len, int64_t ref_start_pos,
                          int64_t n_var, const int64_t *vpos, const char *vop, const int64_t *voplen,
                          const char *alt_pool, const int64_t *alt_off,
                          int64_t rlen, int64_t n_t, const int64_t *ts, const int64_t *te, const int8_t *fo,
                          const char *stub, const char *chrom, int cpy,
                          char *out1, char *out2, int64_t cap, int64_t *len1, int64_t *len2) {
  onode_list nl = {0, 0, 0};
  build_nodes(&nl, ref_seq, ref_len, ref_start_pos, n_var, vpos, vop, voplen, alt_pool, alt_off);
  oread rd[2]; memset(rd, 0, sizeof rd);
  sbuf q = {0, 0, 0};
  int64_t o1 = 0, o2 = 0, cnt = 0, ret = 0;
  for (int64_t t = 0; t < n_t; t++) {
    oread *slot[2] = {0, 0};
    int ok = 1;
    for (int s = 0; s < 2; s++) {                               /* readgenerate.py:202 */
      int64_t p = (s == 0) ? ts[t] : te[t] - rlen;              /* illumina.py:95-96 */
      int f = (s == 0) ? fo[t] : 1 - fo[t];
      int64_t n0 = ss_right_keys(nl.n, nl.cnt, p) - 1, n1 = ss_right_keys(nl.n, nl.cnt, p + rlen - 1) - 1;
      oread *r = &rd[s];
      r->strand = s; r->len = rlen;
      r->pos = gen_read(p, rlen, n0, n1, nl.n, &r->cigar, &r->vlist, &r->seq);
      int64_t nN = 0;
      for (int64_t i = 0; i < r->seq.len; i++) nN += (r->seq.p[i] == 'N');
      if (nN > 2) { ok = 0; break; }                            /* :204 */
      if (s == 1) {                                             /* :205-206 */
        for (int64_t i = 0, j = r->seq.len - 1; i <= j; i++, j--) {
          char a = complement(r->seq.p[i]), b = complement(r->seq.p[j]);
          r->seq.p[i] = b; r->seq.p[j] = a;
        }
      }
      slot[f] = r;                                              /* :207 */
    }
    if (!ok) continue;
    cnt++;                                                      /* :209 */
    q.len = 0;                                                  /* fastq_lines, :222-230 */
    sb_putc(&q, '@'); sb_put(&q, stub, (int64_t)strlen(stub)); sb_putc(&q, ':'); sb_int(&q, cnt);
    sb_putc(&q, '|'); sb_put(&q, chrom, (int64_t)strlen(chrom)); sb_putc(&q, '|'); sb_int(&q, cpy);
    for (int f = 0; f < 2; f++) {
      oread *r = slot[f];
      sb_putc(&q, '|'); sb_int(&q, r->strand); sb_putc(&q, '|'); sb_int(&q, r->pos);
      sb_putc(&q, '|'); sb_int(&q, r->len); sb_putc(&q, '|'); sb_put(&q, r->cigar.p, r->cigar.len);
      sb_putc(&q, '|'); sb_put(&q, r->vlist.p, r->vlist.len);
    }
    for (int f = 0; f < 2; f++) {
      oread *r = slot[f];
      char *out = f ? out2 : out1; int64_t *o = f ? &o2 : &o1;
      int64_t need = q.len + 1 + r->seq.len + 3 + r->len + 1;
      if (*o + need > cap) { ret = -1; goto done; }
      memcpy(out + *o, q.p, q.len); *o += q.len; out[(*o)++] = '\n';
      memcpy(out + *o, r->seq.p, r->seq.len); *o += r->seq.len;
      memcpy(out + *o, "\n+\n", 3); *o += 3;
      memset(out + *o, '~', r->len); *o += r->len; out[(*o)++] = '\n';
    }
  }
  ret = cnt;
done:
  *len1 = o1; *len2 = o2;
  for (int s = 0; s < 2; s++) { free(rd[s].cigar.p); free(rd[s].vlist.p); free(rd[s].seq.p); }
  free(q.p); free(nl.n);
  return ret;
}

/* ------------------------------------------------------------------------------------------ */
/* a16-a18: corrupt-reads, readcorrupt.py:18-118 + illumina.py:113-162 (single worker)          */

static const char *base_rot(char c) {                           /* illumina.py:131-136 */
  switch (c) { case 'A': return "CTG"; case 'C': return "ATG"; case 'T': return "ACG"; case 'G': return "ACT"; default: return "NNN"; }
}

typedef struct { const char *name; int64_t name_len; const char *seq; int64_t seq_len; } fq_rec;

/* next 4-line record; name = header after '@' up to first whitespace (FastxFile .name) */
static int fq_next(const char *buf, int64_t len, int64_t *off, fq_rec *r) {
  if (*off >= len) return 0;
  int64_t o = *off, e;
  const char *nl = (const char *)memchr(buf + o, '\n', len - o); e = nl ? nl - buf : len;
  int64_t ns = o + 1, ne = ns;
  while (ne < e && buf[ne] != ' ' && buf[ne] != '\t' && buf[ne] != '\r') ne++;
  r->name = buf + ns; r->name_len = ne - ns;
  o = e + 1; if (o > len) o = len;
  nl = (const char *)memchr(buf + o, '\n', len - o); e = nl ? nl - buf : len;
  r->seq = buf + o; r->seq_len = e - o;
  o = e + 1;
  for (int k = 0; k < 2; k++) { if (o > len) o = len; nl = (const char *)memchr(buf + o, '\n', len - o); e = nl ? nl - buf : len; o = e + 1; }
  *off = o > len ? len : o;
  return 1;
}

/* cum_bq_mat: [2][n_cycles][n_bq] doubles; phred_p: 100 doubles = 10**(-arange(100)/10).
 * worker_seed = the seed handed to the (single) worker.  in2/out2 may be NULL (single-end).
 * Returns templates processed, -1 on overflow, -2 if a read is longer than n_cycles. */
int64_t orc_corrupt_fastq(const double *cum_bq_mat, int64_t n_cycles, int64_t n_bq, const double *phred_p,
                          uint32_t worker_seed,
                          const char *in1, int64_t in1_len, const char *in2, int64_t in2_len,
                          char *out1, char *out2, int64_t cap, int64_t *len1, int64_t *len2) {
  mt_state rng; mt_seed(&rng, worker_seed);                     /* readcorrupt.py:84 */
  int64_t off1 = 0, off2 = 0, o[2] = {0, 0}, cnt = 0;
  int64_t bufcap = 1024;
  double *bq_rnd = (double *)malloc(bufcap * sizeof(double)), *call_rnd = (double *)malloc(bufcap * sizeof(double));
  uint8_t *base_rnd = (uint8_t *)malloc(bufcap);
  fq_rec r[2];
  for (;;) {
    if (!fq_next(in1, in1_len, &off1, &r[0])) break;
    int nm = 1;
    if (in2) { if (!fq_next(in2, in2_len, &off2, &r[1])) break; nm = 2; }   /* zip(), readcorrupt.py:53 */
    for (int m = 0; m < nm; m++) {                              /* illumina.py:125-128 */
      int64_t L = r[m].seq_len;
      if (L > n_cycles) { cnt = -2; goto done; }
      if (L > bufcap) { bufcap = L; bq_rnd = (double *)realloc(bq_rnd, L * sizeof(double)); call_rnd = (double *)realloc(call_rnd, L * sizeof(double)); base_rnd = (uint8_t *)realloc(base_rnd, L); }
      for (int64_t n = 0; n < L; n++) bq_rnd[n] = mt_double(&rng);        /* illumina.py:151 */
      for (int64_t n = 0; n < L; n++) call_rnd[n] = mt_double(&rng);      /* :152 */
      for (int64_t n = 0; n < L; n++) base_rnd[n] = (uint8_t)mt_bounded(&rng, 2);  /* :153 */
      char *out = m ? out2 : out1;
      int64_t need = 1 + r[0].name_len + 1 + L + 3 + L + 1;
      if (o[m] + need > cap) { cnt = -1; goto done; }
      char *w = out + o[m];
      *w++ = '@'; memcpy(w, r[0].name, r[0].name_len); w += r[0].name_len; *w++ = '\n';   /* readcorrupt.py:113 */
      char *sq = w, *bq = w + L + 3;
      const double *mat = cum_bq_mat + (int64_t)m * n_cycles * n_bq;
      for (int64_t n = 0; n < L; n++) {
        int64_t b = ss_left_f64(mat + n * n_bq, n_bq, bq_rnd[n]); if (b > 93) b = 93;     /* illumina.py:156 */
        char c = r[m].seq[n];
        if (call_rnd[n] < phred_p[b]) c = base_rot(c)[base_rnd[n]];                       /* :159-160 */
        sq[n] = c; bq[n] = (char)(b + 33);
      }
      memcpy(sq + L, "\n+\n", 3); bq[L] = '\n';
      o[m] += need;
    }
    cnt++;
  }
done:
  *len1 = o[0]; *len2 = o[1];
  free(bq_rnd); free(call_rnd); free(base_rnd);
  return cnt;
}

/* p_min / p_max of a (region, copy): readgenerate.py:192.  Returns node count. */
int64_t orc_node_span(const char *ref_seq, int64_t ref_len, int64_t ref_start_pos,
                      int64_t n_var, const int64_t *vpos, const char *vop, const int64_t *voplen,
                      const char *alt_pool, const int64_t *alt_off, int64_t *p_min, int64_t *p_max, int *last_op) {
  onode_list l = {0, 0, 0};
  build_nodes(&l, ref_seq, ref_len, ref_start_pos, n_var, vpos, vop, voplen, alt_pool, alt_off);
  int64_t n = l.cnt;
  if (n > 0) { *p_min = l.n[0].ps; *p_max = l.n[n - 1].ps + l.n[n - 1].oplen; *last_op = l.n[n - 1].op; }
  free(l.n);
  return n;
}
