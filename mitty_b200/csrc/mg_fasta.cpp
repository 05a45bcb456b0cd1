// FASTA front end of generate-reads: what pysam.FastaFile(...).fetch(reference=, start=, end=) does for the
// reference (mitty/simulation/readgenerate.py:181, 186; htslib faidx underneath) as native code inside the
// library -- no GIL, no Python objects, so the worker threads of several GPUs fetch their regions at once.
//
// The file is mapped, the '>' headers are found with memchr, and every contig whose lines all have the same
// width (every FASTA a genome is distributed as) is addressed arithmetically: base i of a contig sits at
// body + i + (i / width) * eol.  A fetch is then one memcpy per line, split over a few threads for large
// regions.  A contig with ragged lines is stripped of its line ends once and kept.  Bytes come back as they are
// in the file (case and IUPAC codes preserved), as the reference's fetch returns them.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mitty_b200.h"

namespace {

struct Contig {
  std::string name;
  int64_t body = 0, body_end = 0;     // file offsets of the sequence lines (trailing line ends trimmed)
  int64_t width = 0, eol = 1;         // uniform layout: bases per line, bytes of a line end (1 = \n, 2 = \r\n)
  bool uniform = false;
  int64_t length = 0;                 // bases
  std::vector<uint8_t> flat;          // ragged contig: its bases, filled on first use
  bool flat_ready = false;
};

}  // namespace

struct mg_fasta {
  int fd = -1;
  const uint8_t *data = nullptr;
  int64_t size = 0;
  std::vector<Contig> contigs;
  std::mutex mu;
  std::string err;
};

namespace {

// strip the line ends of [a, b) into out -> bases written
int64_t strip_lines(const uint8_t *p, int64_t a, int64_t b, uint8_t *out) {
  int64_t n = 0;
  while (a < b) {
    const uint8_t *nl = static_cast<const uint8_t *>(memchr(p + a, '\n', (size_t)(b - a)));
    int64_t e = nl ? (int64_t)(nl - p) : b;
    int64_t le = e;
    if (le > a && p[le - 1] == '\r') le--;
    if (out) memcpy(out + n, p + a, (size_t)(le - a));
    n += le - a;
    a = e + 1;
  }
  return n;
}

void index_contig(const uint8_t *p, Contig &c) {
  int64_t s = c.body, e = c.body_end;
  while (e > s && (p[e - 1] == '\n' || p[e - 1] == '\r')) e--;
  c.body_end = e;
  if (e == s) { c.uniform = true; c.width = 1; c.eol = 1; c.length = 0; return; }
  const uint8_t *nl = static_cast<const uint8_t *>(memchr(p + s, '\n', (size_t)(e - s)));
  if (!nl) { c.uniform = true; c.width = e - s; c.eol = 1; c.length = e - s; return; }       // one line
  int64_t first = (int64_t)(nl - (p + s));
  c.eol = (first > 0 && p[s + first - 1] == '\r') ? 2 : 1;
  c.width = first - (c.eol - 1);
  if (c.width <= 0) { c.uniform = false; c.length = strip_lines(p, s, e, nullptr); return; }
  const int64_t lb = c.width + c.eol, body = e - s;
  const int64_t n_full = body / lb, rem = body - n_full * lb;
  bool ok = rem <= c.width;
  for (int64_t k = 1; k <= n_full && ok; k++) {                   // every line ends where the first one says
    const int64_t at = s + k * lb - 1;
    ok = p[at] == '\n' && (c.eol == 1 || p[at - 1] == '\r');
  }
  // a line end anywhere else (two short lines adding up to one full line) would pass the strided test: the
  // number of line ends of a uniform body is exactly n_full
  if (ok) {
    int64_t cnt = 0;
    for (int64_t a = s; a < e;) {
      const uint8_t *q = static_cast<const uint8_t *>(memchr(p + a, '\n', (size_t)(e - a)));
      if (!q) break;
      cnt++; a = (int64_t)(q - p) + 1;
      if (cnt > n_full) break;
    }
    ok = cnt == n_full;
  }
  c.uniform = ok;
  c.length = ok ? n_full * c.width + rem : strip_lines(p, s, e, nullptr);
}

void copy_uniform(const uint8_t *p, const Contig &c, int64_t start, int64_t end, uint8_t *out) {
  const int64_t lb = c.width + c.eol;
  int64_t i = start;
  while (i < end) {
    const int64_t line = i / c.width, col = i - line * c.width;
    const int64_t n = std::min(end - i, c.width - col);
    memcpy(out + (i - start), p + c.body + line * lb + col, (size_t)n);
    i += n;
  }
}

}  // namespace

extern "C" {

int mg_fasta_open(const char *path, mg_fasta **out) {
  if (!path || !out) return MG_EINVAL;
  *out = nullptr;
  const int fd = open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) return MG_EVALUE;
  struct stat st;
  if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); return MG_EVALUE; }
  mg_fasta *f = new mg_fasta();
  f->fd = fd; f->size = (int64_t)st.st_size;
  if (f->size > 0) {
    void *m = mmap(nullptr, (size_t)f->size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { close(fd); delete f; return MG_EVALUE; }
    f->data = static_cast<const uint8_t *>(m);
    madvise(m, (size_t)f->size, MADV_WILLNEED);
  }
  const uint8_t *p = f->data;
  if (f->size >= 2 && p[0] == 0x1f && p[1] == 0x8b) { mg_fasta_close(f); return MG_EVALUE; }   // gzip: the caller's own reader
  // headers: '>' at a line start
  std::vector<int64_t> hdr;
  int64_t at = 0;
  while (at < f->size) {
    if (p[at] == '>') hdr.push_back(at);
    const uint8_t *q = static_cast<const uint8_t *>(memchr(p + at, '\n', (size_t)(f->size - at)));
    if (!q) break;
    at = (int64_t)(q - p) + 1;
    // jump from header to header: the next "\n>" (memmem would do; memchr on '>' then a look back is as fast)
    while (at < f->size && p[at] != '>') {
      const uint8_t *g = static_cast<const uint8_t *>(memchr(p + at, '>', (size_t)(f->size - at)));
      if (!g) { at = f->size; break; }
      at = (int64_t)(g - p);
      if (p[at - 1] == '\n') break;
      at++;
    }
  }
  f->contigs.resize(hdr.size());
  for (size_t k = 0; k < hdr.size(); k++) {
    Contig &c = f->contigs[k];
    const int64_t h = hdr[k], lim = k + 1 < hdr.size() ? hdr[k + 1] : f->size;
    const uint8_t *q = static_cast<const uint8_t *>(memchr(p + h, '\n', (size_t)(lim - h)));
    const int64_t eol = q ? (int64_t)(q - p) : lim;
    int64_t ne = h + 1;                                             // the name: up to the first white space
    while (ne < eol && p[ne] != ' ' && p[ne] != '\t' && p[ne] != '\r') ne++;
    c.name.assign(reinterpret_cast<const char *>(p + h + 1), (size_t)(ne - h - 1));
    c.body = std::min(eol + 1, lim); c.body_end = lim;
  }
  // indexing touches every byte once (the line-end count): a few threads share the contigs
  const unsigned nt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([f, p, t, nt]() { for (size_t k = t; k < f->contigs.size(); k += nt) index_contig(p, f->contigs[k]); });
  for (auto &x : th) x.join();
  *out = f;
  return MG_OK;
}

void mg_fasta_close(mg_fasta *f) {
  if (!f) return;
  if (f->data) munmap(const_cast<uint8_t *>(f->data), (size_t)f->size);
  if (f->fd >= 0) close(f->fd);
  delete f;
}

int64_t mg_fasta_n_contigs(mg_fasta *f) { return f ? (int64_t)f->contigs.size() : 0; }

int mg_fasta_contig(mg_fasta *f, int64_t i, const char **name, int64_t *length, int32_t *uniform) {
  if (!f || i < 0 || i >= (int64_t)f->contigs.size()) return MG_EINVAL;
  if (name) *name = f->contigs[(size_t)i].name.c_str();
  if (length) *length = f->contigs[(size_t)i].length;
  if (uniform) *uniform = f->contigs[(size_t)i].uniform ? 1 : 0;
  return MG_OK;
}

int64_t mg_fasta_fetch(mg_fasta *f, const char *name, int64_t start, int64_t end, uint8_t *out, int32_t threads) {
  if (!f || !name) return MG_EINVAL;
  Contig *c = nullptr;
  for (Contig &x : f->contigs) if (x.name == name) { c = &x; break; }        // the first record of that name, as faidx
  if (!c) return MG_EINDEX;
  if (start < 0) start = 0;
  if (end > c->length) end = c->length;
  if (end <= start) return 0;
  if (!out) return MG_EINVAL;
  if (!c->uniform) {
    {
      std::lock_guard<std::mutex> lk(f->mu);
      if (!c->flat_ready) { c->flat.resize((size_t)c->length); strip_lines(f->data, c->body, c->body_end, c->flat.data()); c->flat_ready = true; }
    }
    memcpy(out, c->flat.data() + start, (size_t)(end - start));
    return end - start;
  }
  const int64_t n = end - start;
  int nt = threads < 1 ? 1 : (threads > 16 ? 16 : threads);
  if (n < (8ll << 20)) nt = 1;
  if (nt == 1) { copy_uniform(f->data, *c, start, end, out); return n; }
  std::vector<std::thread> th;
  const int64_t per = (n + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    const int64_t a = start + t * per, b = std::min(end, a + per);
    if (a >= b) break;
    th.emplace_back([f, c, a, b, start, out]() { copy_uniform(f->data, *c, a, b, out + (a - start)); });
  }
  for (auto &x : th) x.join();
  return n;
}

}  // extern "C"
