// farms_textio.cpp -- see include/farms_textio.h.
#include "farms_textio.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

int pick_threads(int n) {
  if (n > 0) return n;
  unsigned h = std::thread::hardware_concurrency();
  return (int)std::max(1u, h);
}

struct Line {  // what one line yields: how many leading fields parsed, and their values
  int nf;
  long long v[4];
};

inline Line parse_line(const char *q, const char *eol) {
  Line L;
  L.nf = 0;
  while (L.nf < 4) {
    while (q < eol && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
    if (q >= eol) break;
    bool neg = false;
    if (*q == '-' || *q == '+') {
      neg = *q == '-';
      q++;
    }
    if (q >= eol || *q < '0' || *q > '9') break;
    unsigned long long v = 0;
    while (q < eol && *q >= '0' && *q <= '9') v = v * 10 + (unsigned long long)(*q++ - '0');
    L.v[L.nf++] = neg ? -(long long)v : (long long)v;
  }
  return L;
}

template <class F>
void parallel_for(int nthreads, size_t nitems, F f) {
  if (nthreads <= 1 || nitems <= 1) {
    for (size_t i = 0; i < nitems; i++) f(i);
    return;
  }
  std::vector<std::thread> th;
  const int nt = (int)std::min<size_t>((size_t)nthreads, nitems);
  for (int k = 0; k < nt; k++)
    th.emplace_back([&, k]() {
      for (size_t i = (size_t)k; i < nitems; i += (size_t)nt) f(i);
    });
  for (auto &t : th) t.join();
}

inline char *put_int(char *p, long long v) {
  char tmp[24];
  int n = 0;
  unsigned long long u = v < 0 ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
  if (v < 0) *p++ = '-';
  do {
    tmp[n++] = (char)('0' + u % 10);
    u /= 10;
  } while (u);
  while (n) *p++ = tmp[--n];
  return p;
}

inline char *put_g(char *p, double v) {
  if (v == 0.0 && !std::signbit(v)) {  // the most common value (rows without flow)
    *p++ = '0';
    return p;
  }
  // "%g" == chars_format::general with precision 6 (C++17 [charconv.to.chars]); exact and ~10x faster than printf
  if (std::isfinite(v)) {
    auto r = std::to_chars(p, p + 32, v, std::chars_format::general, 6);
    return r.ptr;
  }
  return p + std::snprintf(p, 32, "%g", v);
}

}  // namespace

extern "C" int farms_text_read(const char *path, uint64_t max_events, int nthreads, farms_events *out, char *err,
                               size_t errlen) {
  std::memset(out, 0, sizeof *out);
  auto fail = [&](const std::string &m) {
    if (err && errlen) std::snprintf(err, errlen, "%s", m.c_str());
    return -1;
  };
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) return fail(std::string("Unable to open file ") + path);
  struct stat st;
  if (::fstat(fd, &st) != 0) {
    ::close(fd);
    return fail(std::string("Unable to open file ") + path);
  }
  const size_t got = (size_t)st.st_size;
  void *map = got ? ::mmap(nullptr, got, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
  ::close(fd);
  if (got && map == MAP_FAILED) return fail(std::string("Unable to map file ") + path);
  struct Unmap {
    void *p; size_t n;
    ~Unmap() { if (p) ::munmap(p, n); }
  } unmap{map, got};
  const char *base = (const char *)map, *end = base + got;
  if (!got) base = end = "";
  nthreads = pick_threads(nthreads);

  // 1. cut into pieces at line boundaries and count the lines of each piece
  const size_t npieces = std::max<size_t>(1, std::min<size_t>((size_t)nthreads * 4, got / (1 << 16) + 1));
  std::vector<const char *> cut(npieces + 1);
  cut[0] = base;
  cut[npieces] = end;
  for (size_t k = 1; k < npieces; k++) {
    const char *p = base + got / npieces * k;
    const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
    cut[k] = nl ? nl + 1 : end;
  }
  for (size_t k = 1; k <= npieces; k++) cut[k] = std::max(cut[k], cut[k - 1]);
  std::vector<uint64_t> nlines(npieces + 1, 0);
  parallel_for(nthreads, npieces, [&](size_t k) {
    uint64_t c = 0;
    const char *s = cut[k];
    while (s < cut[k + 1]) {
      const char *nl = (const char *)std::memchr(s, '\n', (size_t)(cut[k + 1] - s));
      c++;
      if (!nl) break;
      s = nl + 1;
    }
    nlines[k + 1] = c;
  });
  for (size_t k = 0; k < npieces; k++) nlines[k + 1] += nlines[k];
  const uint64_t n = std::min<uint64_t>(nlines[npieces], max_events);

  out->n = n;
  const size_t na = (size_t)std::max<uint64_t>(n, 1);
  out->x = (uint16_t *)std::malloc(na * 2);
  out->y = (uint16_t *)std::malloc(na * 2);
  out->t = (uint64_t *)std::malloc(na * 8);
  out->xi = (int32_t *)std::malloc(na * 4);
  out->yi = (int32_t *)std::malloc(na * 4);
  out->pol = (int32_t *)std::malloc(na * 4);
  uint8_t *nf = (uint8_t *)std::malloc(na);
  long long *raw_p = (long long *)std::malloc(na * 8);  // polarity before clamping (needed by the stale-value fix-up)
  struct Free2 {
    void *a, *b;
    ~Free2() { std::free(a); std::free(b); }
  } free2{nf, raw_p};
  if (!nf || !raw_p || !out->x || !out->y || !out->t || !out->xi || !out->yi || !out->pol) {
    farms_text_free(out);
    return fail("out of memory");
  }

  // 2. parse every piece into place; remember lines with fewer than 4 fields
  std::vector<uint8_t> piece_short(npieces, 0);
  parallel_for(nthreads, npieces, [&](size_t k) {
    uint64_t i = nlines[k];
    const char *s = cut[k];
    while (s < cut[k + 1] && i < n) {
      const char *nl = (const char *)std::memchr(s, '\n', (size_t)(cut[k + 1] - s));
      const char *eol = nl ? nl : cut[k + 1];
      const Line L = parse_line(s, eol);
      nf[i] = (uint8_t)L.nf;
      if (L.nf < 4) piece_short[k] = 1;
      out->xi[i] = L.nf > 0 ? (int32_t)L.v[0] : 0;
      out->yi[i] = L.nf > 1 ? (int32_t)L.v[1] : 0;
      out->t[i] = L.nf > 2 ? (uint64_t)L.v[2] : 0;
      raw_p[i] = L.nf > 3 ? L.v[3] : 0;
      i++;
      if (!nl) break;
      s = nl + 1;
    }
  });
  // 3. short lines inherit the missing fields from the line before (sequential, rare)
  long long px = 0, py = 0, pp = 0;
  uint64_t pt = 0;
  bool any_short = false;
  for (size_t k = 0; k < npieces; k++) any_short |= piece_short[k] != 0;
  if (any_short)
    for (uint64_t i = 0; i < n; i++) {
      if (nf[i] < 1) out->xi[i] = (int32_t)px;
      if (nf[i] < 2) out->yi[i] = (int32_t)py;
      if (nf[i] < 3) out->t[i] = pt;
      if (nf[i] < 4) raw_p[i] = pp;
      px = out->xi[i]; py = out->yi[i]; pt = out->t[i]; pp = raw_p[i];
    }
  // 4. device-ready columns, range check
  std::vector<int64_t> bad(npieces, -1);
  parallel_for(nthreads, npieces, [&](size_t k) {
    const uint64_t a = n * k / npieces, b = n * (k + 1) / npieces;
    for (uint64_t i = a; i < b; i++) {
      const int32_t x = out->xi[i], y = out->yi[i];
      if ((x < 0 || x > 65535 || y < 0 || y > 65535) && bad[k] < 0) bad[k] = (int64_t)i;
      out->x[i] = (uint16_t)x;
      out->y[i] = (uint16_t)y;
      out->pol[i] = raw_p[i] < 0 ? 0 : (int32_t)raw_p[i];  // src/vFlow.cpp:246-247
    }
  });
  for (size_t k = 0; k < npieces; k++)
    if (bad[k] >= 0) {
      const std::string m = "event " + std::to_string(bad[k]) + " has coordinates outside the sensor";
      farms_text_free(out);
      return fail(m);
    }
  return 0;
}

extern "C" void farms_text_free(farms_events *ev) {
  if (!ev) return;
  std::free(ev->x); std::free(ev->y); std::free(ev->t); std::free(ev->xi); std::free(ev->yi); std::free(ev->pol);
  std::memset(ev, 0, sizeof *ev);
}

extern "C" int farms_text_write(const char *path11, const char *path8, uint64_t n, const int32_t *xi, const int32_t *yi,
                                const uint32_t *t_rel, const int32_t *pol, const double *gr, const double *gth,
                                const double *vx, const double *vy, const double *lr, const double *lth,
                                const uint8_t *scale, int nthreads) {
  FILE *f11 = std::fopen(path11, "wb");
  FILE *f8 = path8 ? std::fopen(path8, "wb") : nullptr;
  if (!f11 || (path8 && !f8)) {
    if (f11) std::fclose(f11);
    if (f8) std::fclose(f8);
    return -1;
  }
  nthreads = pick_threads(nthreads);
  const uint64_t BLOCK = 1 << 16;  // rows per formatting task
  const uint64_t nblocks = (n + BLOCK - 1) / BLOCK;
  std::mutex mu;
  std::condition_variable cv;
  uint64_t next_to_write = 0, next_task = 0;
  bool io_error = false;
  auto worker = [&]() {
    std::vector<char> b11(BLOCK * 200), b8(f8 ? BLOCK * 160 : 1);
    for (;;) {
      uint64_t blk;
      {
        std::lock_guard<std::mutex> g(mu);
        blk = next_task++;
      }
      if (blk >= nblocks) return;
      char *p = b11.data(), *q = b8.data();
      const uint64_t a = blk * BLOCK, e = std::min(n, a + BLOCK);
      for (uint64_t i = a; i < e; i++) {
        char head[64], *h = head;
        h = put_int(h, xi[i]); *h++ = ' ';
        h = put_int(h, yi[i]); *h++ = ' ';
        h = put_int(h, (int32_t)t_rel[i]); *h++ = ' ';  // T_out is a vector<int> (src/vFlow.cpp:136, 373)
        h = put_int(h, pol[i]); *h++ = ' ';
        const size_t hl = (size_t)(h - head);
        std::memcpy(p, head, hl); p += hl;
        char *g0 = p;
        p = put_g(p, gr[i]); *p++ = ' ';
        p = put_g(p, gth[i]); *p++ = ' ';
        const size_t gl = (size_t)(p - g0);
        p = put_g(p, vx[i]); *p++ = ' ';
        p = put_g(p, vy[i]); *p++ = ' ';
        char *l0 = p;
        p = put_g(p, lr[i]); *p++ = ' ';
        p = put_g(p, lth[i]);
        const size_t ll = (size_t)(p - l0);
        *p++ = ' ';
        p = put_int(p, scale[i]);
        *p++ = '\n';
        if (f8) {
          std::memcpy(q, head, hl); q += hl;
          std::memcpy(q, g0, gl); q += gl;
          std::memcpy(q, l0, ll); q += ll;
          *q++ = '\n';
        }
      }
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return next_to_write == blk; });
      if (std::fwrite(b11.data(), 1, (size_t)(p - b11.data()), f11) != (size_t)(p - b11.data())) io_error = true;
      if (f8 && std::fwrite(b8.data(), 1, (size_t)(q - b8.data()), f8) != (size_t)(q - b8.data())) io_error = true;
      next_to_write++;
      lk.unlock();
      cv.notify_all();
    }
  };
  std::vector<std::thread> th;
  const int nt = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)nthreads, nblocks));
  for (int k = 1; k < nt; k++) th.emplace_back(worker);
  worker();
  for (auto &t : th) t.join();
  if (std::fclose(f11) != 0) io_error = true;
  if (f8 && std::fclose(f8) != 0) io_error = true;
  return io_error ? -1 : 0;
}

// ---------------------------------------------------------------------------------------------------
// binary side-format (include/farms_textio.h): little-endian SoA columns behind an 8-byte magic and a count
// ---------------------------------------------------------------------------------------------------
namespace {

template <class T>
bool read_col(FILE *f, T *dst, uint64_t n_file, uint64_t n_take) {
  if (n_take && std::fread(dst, sizeof(T), (size_t)n_take, f) != (size_t)n_take) return false;
  return std::fseek(f, (long)((n_file - n_take) * sizeof(T)), SEEK_CUR) == 0;
}

template <class T>
bool write_col(FILE *f, const T *src, uint64_t n) {
  return n == 0 || std::fwrite(src, sizeof(T), (size_t)n, f) == (size_t)n;
}

}  // namespace

extern "C" int farms_bin_read(const char *path, uint64_t max_events, farms_events *out, char *err, size_t errlen) {
  auto fail = [&](const std::string &m) {
    if (err && errlen) std::snprintf(err, errlen, "%s", m.c_str());
    return -1;
  };
  if (!out) return fail("null output");
  std::memset(out, 0, sizeof *out);
  FILE *f = std::fopen(path, "rb");
  if (!f) return fail(std::string("Unable to open file ") + path);
  char magic[8];
  uint64_t n_file = 0;
  if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "FARMSEV1", 8) != 0 || std::fread(&n_file, 8, 1, f) != 1) {
    std::fclose(f);
    return fail(std::string(path) + ": not a FARMSEV1 event file");
  }
  const uint64_t n = std::min(n_file, max_events);
  const size_t cap = (size_t)std::max<uint64_t>(n, 1);
  out->x = (uint16_t *)std::malloc(cap * 2);
  out->y = (uint16_t *)std::malloc(cap * 2);
  out->t = (uint64_t *)std::malloc(cap * 8);
  out->xi = (int32_t *)std::malloc(cap * 4);
  out->yi = (int32_t *)std::malloc(cap * 4);
  out->pol = (int32_t *)std::malloc(cap * 4);
  std::vector<uint8_t> p((size_t)n);
  bool ok = out->x && out->y && out->t && out->xi && out->yi && out->pol;
  ok = ok && read_col(f, out->x, n_file, n) && read_col(f, out->y, n_file, n) && read_col(f, out->t, n_file, n) &&
       read_col(f, p.data(), n_file, n);
  std::fclose(f);
  if (!ok) {
    farms_text_free(out);
    return fail(std::string(path) + ": truncated event file");
  }
  for (uint64_t i = 0; i < n; i++) {
    out->xi[i] = out->x[i];
    out->yi[i] = out->y[i];
    out->pol[i] = p[i];  // u8: already >= 0 (src/vFlow.cpp:246-247 clamps negatives)
  }
  out->n = n;
  return 0;
}

extern "C" int farms_bin_write_events(const char *path, uint64_t n, const uint16_t *x, const uint16_t *y,
                                      const uint64_t *t, const uint8_t *p) {
  FILE *f = std::fopen(path, "wb");
  if (!f) return -1;
  bool ok = std::fwrite("FARMSEV1", 1, 8, f) == 8 && std::fwrite(&n, 8, 1, f) == 1 && write_col(f, x, n) &&
            write_col(f, y, n) && write_col(f, t, n) && write_col(f, p, n);
  ok = std::fclose(f) == 0 && ok;
  return ok ? 0 : -1;
}

extern "C" int farms_bin_write(const char *path, uint64_t n, const int32_t *xi, const int32_t *yi,
                               const uint32_t *t_rel, const int32_t *pol, const double *gr, const double *gth,
                               const double *vx, const double *vy, const double *lr, const double *lth,
                               const uint8_t *scale) {
  FILE *f = std::fopen(path, "wb");
  if (!f) return -1;
  std::vector<uint16_t> x16((size_t)n), y16((size_t)n);
  std::vector<uint8_t> p8((size_t)n);
  for (uint64_t i = 0; i < n; i++) {
    x16[i] = (uint16_t)xi[i];
    y16[i] = (uint16_t)yi[i];
    p8[i] = (uint8_t)pol[i];
  }
  bool ok = std::fwrite("FARMSOU1", 1, 8, f) == 8 && std::fwrite(&n, 8, 1, f) == 1 && write_col(f, x16.data(), n) &&
            write_col(f, y16.data(), n) && write_col(f, t_rel, n) && write_col(f, p8.data(), n) &&
            write_col(f, scale, n) && write_col(f, gr, n) && write_col(f, gth, n) && write_col(f, vx, n) &&
            write_col(f, vy, n) && write_col(f, lr, n) && write_col(f, lth, n);
  ok = std::fclose(f) == 0 && ok;
  return ok ? 0 : -1;
}
