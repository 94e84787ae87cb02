// tools/farms_synth.cpp -- see farms_synth.h.  Scenes follow SURVEY.md section 8(d).
#include "farms_synth.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr uint64_t BUCKET_US = 1024;
constexpr uint64_t T_OFFSET = 1000;

inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t hash4(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  return mix64(mix64(mix64(mix64(seed) ^ a) ^ b) ^ c);
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
// bounded, near-Gaussian unit-variance jitter: Irwin-Hall(4) from one 64-bit hash, |g| <= 2*sqrt(3)
inline double jitter(uint64_t h) {
  double s = (double)(h & 0xFFFF) + (double)((h >> 16) & 0xFFFF) + (double)((h >> 32) & 0xFFFF) +
             (double)((h >> 48) & 0xFFFF);
  return (s * (1.0 / 65536.0) - 2.0) * 1.7320508075688772;
}
constexpr double JMAX = 3.4641016151377544;  // 2*sqrt(3)

struct Ev {
  uint32_t dt;  // time within bucket
  uint16_t x, y;
  uint8_t p;
};

struct Point {  // a moving edge point
  double x0, y0, vx, vy;  // linear: px/s, torus wrap
  double rho, phi;        // rotation: radius, initial angle
  uint8_t pol;
};

}  // namespace

struct farms_synth {
  int config = 0;
  uint64_t seed = 0;
  int W = 0, H = 0, fs = 5;
  double sigma_us = 0;     // timestamp jitter
  double noise_frac = 0;   // background noise as a fraction of the signal rate
  double rate = 0;         // approx events/s (signal + noise)
  int kind = 0;            // 0 bar, 1 torus points, 2 rotating points
  // bar
  double bar_c = 0, bar_s = 0, bar_speed = 0, bar_width = 0, bar_period_us = 0, bar_margin = 0;
  // rotation
  double cx = 0, cy = 0, omega = 0;
  std::vector<Point> pts;
  int noise_per_bucket = 0;
};

namespace {

void emit(std::vector<Ev> &out, const farms_synth *s, uint64_t b, double t_nom_us, uint64_t h, long px,
          long py, uint8_t pol, bool torus) {
  double tf = std::floor(t_nom_us + s->sigma_us * jitter(h) + 0.5);
  if (tf < 0) tf = 0;
  uint64_t ti = (uint64_t)tf;
  if (ti / BUCKET_US != b) return;
  if (torus) {
    px %= s->W; if (px < 0) px += s->W;
    py %= s->H; if (py < 0) py += s->H;
  } else if (px < 0 || px >= s->W || py < 0 || py >= s->H) {
    return;
  }
  out.push_back(Ev{(uint32_t)(ti - b * BUCKET_US), (uint16_t)px, (uint16_t)py, pol});
}

// generate bucket b (unsorted, deterministic generation order)
void gen_bucket(const farms_synth *s, uint64_t b, std::vector<Ev> &out) {
  out.clear();
  const double pad = s->sigma_us * JMAX + 1.0;
  const double ta = (double)(b * BUCKET_US) - pad, tb = (double)((b + 1) * BUCKET_US) + pad;  // us
  if (s->kind == 0) {
    // translating bar: pixel (x,y) sees the leading edge (p=1) when x*c + y*s + margin = speed*t
    // (t within the sweep period) and the trailing edge (p=0) bar_width later.
    for (int edge = 0; edge < 2; edge++) {
      const double off = s->bar_margin + (edge ? s->bar_width : 0.0);
      // sweeps m whose window intersects [ta, tb)
      long m0 = (long)std::floor(ta / s->bar_period_us) - 1, m1 = (long)std::floor(tb / s->bar_period_us);
      for (long m = std::max(0l, m0); m <= m1; m++) {
        const double base = (double)m * s->bar_period_us;
        // d = x*c + y*s + off ; t_us = base + d / speed * 1e6  in [ta, tb)
        const double dlo = (ta - base) * 1e-6 * s->bar_speed - off, dhi = (tb - base) * 1e-6 * s->bar_speed - off;
        for (int x = 0; x < s->W; x++) {
          double ylo = (dlo - x * s->bar_c) / s->bar_s, yhi = (dhi - x * s->bar_c) / s->bar_s;
          long y0 = std::max(0l, (long)std::floor(ylo)), y1 = std::min((long)s->H - 1, (long)std::ceil(yhi));
          for (long y = y0; y <= y1; y++) {
            double d = x * s->bar_c + y * s->bar_s + off;
            double t = base + d / s->bar_speed * 1e6;
            if (t < ta || t >= tb) continue;
            emit(out, s, b, t, hash4(s->seed, (uint64_t)m * 2 + edge, (uint64_t)x, (uint64_t)y), x, y,
                 edge ? 0 : 1, false);
          }
        }
      }
    }
  } else if (s->kind == 1) {
    const double ta_s = std::max(0.0, ta) * 1e-6, tb_s = tb * 1e-6;
    for (size_t k = 0; k < s->pts.size(); k++) {
      const Point &q = s->pts[k];
      for (int axis = 0; axis < 2; axis++) {
        const double p0 = axis ? q.y0 : q.x0, v = axis ? q.vy : q.vx;
        if (v == 0) continue;
        // gridline crossings g (integers) with p0 + v*t = g, t in [ta_s, tb_s)
        double ga = p0 + v * ta_s, gb = p0 + v * tb_s;
        long g0 = (long)std::floor(std::min(ga, gb)), g1 = (long)std::ceil(std::max(ga, gb));
        for (long g = g0; g <= g1; g++) {
          double t = ((double)g - p0) / v;
          if (t < ta_s || t >= tb_s || t <= 0) continue;
          long px, py;
          if (axis == 0) {
            px = v > 0 ? g : g - 1;
            py = (long)std::floor(q.y0 + q.vy * t);
          } else {
            py = v > 0 ? g : g - 1;
            px = (long)std::floor(q.x0 + q.vx * t);
          }
          emit(out, s, b, t * 1e6, hash4(s->seed, k, (uint64_t)axis, (uint64_t)(g + (1l << 40))), px, py, q.pol, true);
        }
      }
    }
  } else {
    // rotating points: fixed global time grid of STEP_US; event when the pixel changes between steps
    const double STEP_US = 8.0;
    long k0 = std::max(0l, (long)std::floor(ta / STEP_US)), k1 = (long)std::ceil(tb / STEP_US);
    for (size_t k = 0; k < s->pts.size(); k++) {
      const Point &q = s->pts[k];
      double a = q.phi + s->omega * ((double)k0 * STEP_US * 1e-6);
      long lx = (long)std::floor(s->cx + q.rho * std::cos(a)), ly = (long)std::floor(s->cy + q.rho * std::sin(a));
      for (long st = k0 + 1; st <= k1; st++) {
        a = q.phi + s->omega * ((double)st * STEP_US * 1e-6);
        long nx = (long)std::floor(s->cx + q.rho * std::cos(a)), ny = (long)std::floor(s->cy + q.rho * std::sin(a));
        if (nx != lx || ny != ly)
          emit(out, s, b, (double)st * STEP_US, hash4(s->seed, k, 7, (uint64_t)st), nx, ny, q.pol, false);
        lx = nx; ly = ny;
      }
    }
  }
  for (int i = 0; i < s->noise_per_bucket; i++) {
    uint64_t h = hash4(s->seed ^ 0xA5A5A5A5ull, b, (uint64_t)i, 1), h2 = hash4(s->seed ^ 0x5A5A5A5Aull, b, (uint64_t)i, 2);
    out.push_back(Ev{(uint32_t)(h % BUCKET_US), (uint16_t)((h >> 16) % (uint64_t)s->W),
                     (uint16_t)((h2 >> 8) % (uint64_t)s->H), (uint8_t)(h2 & 1)});
  }
}

// stable counting sort by dt
void sort_bucket(std::vector<Ev> &ev, std::vector<Ev> &tmp) {
  uint32_t cnt[BUCKET_US + 1];
  std::memset(cnt, 0, sizeof cnt);
  for (const Ev &e : ev) cnt[e.dt + 1]++;
  for (uint32_t i = 0; i < BUCKET_US; i++) cnt[i + 1] += cnt[i];
  tmp.resize(ev.size());
  for (const Ev &e : ev) tmp[cnt[e.dt]++] = e;
  ev.swap(tmp);
}

// a straight edge = a chain of points ~1 px apart sharing one velocity and polarity
void add_segment(farms_synth *s, double xa, double ya, double xb, double yb, double vx, double vy, uint8_t pol) {
  double len = std::hypot(xb - xa, yb - ya);
  long n = std::max(2l, (long)std::floor(len) + 1);
  for (long i = 0; i < n; i++) {
    double f = (double)i / (double)(n - 1);
    Point q{};
    q.x0 = xa + f * (xb - xa);
    q.y0 = ya + f * (yb - ya);
    q.vx = vx; q.vy = vy;
    q.pol = pol;
    s->pts.push_back(q);
  }
}
// rotating-frame segment: points stored as (rho, phi) about the disk centre
void add_segment_polar(farms_synth *s, double xa, double ya, double xb, double yb, uint8_t pol) {
  double len = std::hypot(xb - xa, yb - ya);
  long n = std::max(2l, (long)std::floor(len) + 1);
  for (long i = 0; i < n; i++) {
    double f = (double)i / (double)(n - 1);
    double px = xa + f * (xb - xa), py = ya + f * (yb - ya);
    Point q{};
    q.rho = std::hypot(px, py);
    q.phi = std::atan2(py, px);
    q.pol = pol;
    s->pts.push_back(q);
  }
}

}  // namespace

extern "C" farms_synth *farms_synth_open(int config, uint64_t seed) {
  farms_synth *s = new farms_synth();
  s->config = config;
  const double PI = 3.14159265358979323846;
  double sig_rate = 0;
  switch (config) {
    case 1: {  // translating bar 320x320
      s->seed = seed ? seed : 0xFA1; s->W = 320; s->H = 320; s->fs = 5; s->kind = 0; s->sigma_us = 20;
      s->bar_c = std::cos(20 * PI / 180); s->bar_s = std::sin(20 * PI / 180);
      s->bar_speed = 2000; s->bar_width = 10; s->bar_margin = 4;
      double extent = 320 * s->bar_c + 320 * s->bar_s + s->bar_width + 2 * s->bar_margin;
      s->bar_period_us = std::ceil(extent / s->bar_speed * 1e6 / 1000.0) * 1000.0;
      sig_rate = 2.0 * 320 * 320 / (s->bar_period_us * 1e-6);
      break;
    }
    case 2: {  // rotating textured disk, ATIS 304x240
      s->seed = seed ? seed : 0xFA2; s->W = 304; s->H = 240; s->fs = 5; s->kind = 2; s->sigma_us = 10;
      s->cx = 152; s->cy = 120; s->omega = 20;
      // texture = 40 random chords of the r=100 disk, rotating rigidly (~2000 edge points)
      for (int i = 0; i < 40; i++) {
        double r1 = 100.0 * std::sqrt(u01(hash4(s->seed, 21, (uint64_t)i, 0))), a1 = 2 * PI * u01(hash4(s->seed, 22, (uint64_t)i, 0));
        double ang = 2 * PI * u01(hash4(s->seed, 24, (uint64_t)i, 0)), L = 30 + 40 * u01(hash4(s->seed, 25, (uint64_t)i, 0));
        double xa = r1 * std::cos(a1), ya = r1 * std::sin(a1), xb = xa + L * std::cos(ang), yb = ya + L * std::sin(ang);
        double rb = std::hypot(xb, yb);
        if (rb > 100.0) { xb *= 100.0 / rb; yb *= 100.0 / rb; }
        add_segment_polar(s, xa, ya, xb, yb, (uint8_t)(hash4(s->seed, 23, (uint64_t)i, 0) & 1));
      }
      for (const Point &q : s->pts) sig_rate += s->omega * q.rho * 1.27;  // mean |cos|+|sin| = 4/pi
      break;
    }
    case 3: {  // multi-object random texture, DAVIS346 346x260, filtersize 7
      s->seed = seed ? seed : 0xFA3; s->W = 346; s->H = 260; s->fs = 7; s->kind = 1; s->sigma_us = 10;
      s->noise_frac = 0.05;
      for (int o = 0; o < 12; o++) {
        double w = 20 + 60 * u01(hash4(s->seed, 31, o, 0)), h = 20 + 60 * u01(hash4(s->seed, 32, o, 0));
        double x0 = 346 * u01(hash4(s->seed, 33, o, 0)), y0 = 260 * u01(hash4(s->seed, 34, o, 0));
        double sp = 100 + 2900 * u01(hash4(s->seed, 35, o, 0)), hd = 2 * PI * u01(hash4(s->seed, 36, o, 0));
        size_t before = s->pts.size();
        double vx = sp * std::cos(hd), vy = sp * std::sin(hd);
        // a patch = its rectangular outline + 3 random internal edges, all moving together
        add_segment(s, x0, y0, x0 + w, y0, vx, vy, 1);
        add_segment(s, x0, y0 + h, x0 + w, y0 + h, vx, vy, 0);
        add_segment(s, x0, y0, x0, y0 + h, vx, vy, 1);
        add_segment(s, x0 + w, y0, x0 + w, y0 + h, vx, vy, 0);
        for (int e = 0; e < 3; e++) {
          double ax = x0 + w * u01(hash4(s->seed, 37, o, e)), ay = y0 + h * u01(hash4(s->seed, 38, o, e));
          double bx = x0 + w * u01(hash4(s->seed, 39, o, e)), by = y0 + h * u01(hash4(s->seed, 40, o, e));
          add_segment(s, ax, ay, bx, by, vx, vy, (uint8_t)(hash4(s->seed, 41, o, e) & 1));
        }
        sig_rate += (double)(s->pts.size() - before) * (std::fabs(sp * std::cos(hd)) + std::fabs(sp * std::sin(hd)));
      }
      break;
    }
    case 4:
    case 5: {  // high-rate pan, Prophesee Gen4 1280x720
      s->seed = seed ? seed : 0xFA4; s->W = 1280; s->H = 720; s->fs = 5; s->kind = 1; s->sigma_us = 5;
      s->noise_frac = 0.02;
      // global texture = 720 random 64-px edges (~46k edge points) panning at (4000, 500) px/s
      for (int i = 0; i < 720; i++) {
        double xa = 1280 * u01(hash4(s->seed, 42, (uint64_t)i, 0)), ya = 720 * u01(hash4(s->seed, 43, (uint64_t)i, 0));
        double ang = 2 * PI * u01(hash4(s->seed, 44, (uint64_t)i, 0));
        add_segment(s, xa, ya, xa + 63 * std::cos(ang), ya + 63 * std::sin(ang), 4000, 500,
                    (uint8_t)(hash4(s->seed, 45, (uint64_t)i, 0) & 1));
      }
      sig_rate = (double)s->pts.size() * 4500.0;
      break;
    }
    default:
      delete s;
      return nullptr;
  }
  s->noise_per_bucket = (int)(sig_rate * s->noise_frac * (double)BUCKET_US * 1e-6);
  s->rate = sig_rate + (double)s->noise_per_bucket / ((double)BUCKET_US * 1e-6);
  return s;
}

extern "C" void farms_synth_close(farms_synth *s) { delete s; }

extern "C" void farms_synth_info(const farms_synth *s, int *w, int *h, int *fs, double *rate) {
  if (w) *w = s->W;
  if (h) *h = s->H;
  if (fs) *fs = s->fs;
  if (rate) *rate = s->rate;
}

extern "C" int64_t farms_synth_range(farms_synth *s, uint64_t t_begin, uint64_t t_end, uint16_t *x, uint16_t *y,
                                     uint64_t *t, uint8_t *p, int64_t cap, int nthreads) {
  if (t_end <= t_begin) return 0;
  const uint64_t b0 = t_begin / BUCKET_US, b1 = (t_end - 1) / BUCKET_US;
  const size_t nb = (size_t)(b1 - b0 + 1);
  if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
  nthreads = (int)std::min<size_t>((size_t)nthreads, nb);
  std::vector<std::vector<Ev>> buckets(nb);
  std::atomic<size_t> next{0};
  auto work = [&]() {
    std::vector<Ev> tmp;
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= nb) break;
      gen_bucket(s, b0 + i, buckets[i]);
      sort_bucket(buckets[i], tmp);
      // trim to [t_begin, t_end)
      uint64_t base = (b0 + i) * BUCKET_US;
      std::vector<Ev> &v = buckets[i];
      if (base < t_begin || base + BUCKET_US > t_end) {
        size_t w = 0;
        for (const Ev &e : v) if (base + e.dt >= t_begin && base + e.dt < t_end) v[w++] = e;
        v.resize(w);
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(work);
  work();
  for (auto &q : th) q.join();
  int64_t total = 0;
  for (auto &v : buckets) total += (int64_t)v.size();
  if (total > cap) return -total;
  int64_t o = 0;
  for (size_t i = 0; i < nb; i++) {
    uint64_t base = (b0 + i) * BUCKET_US + T_OFFSET;
    for (const Ev &e : buckets[i]) {
      x[o] = e.x; y[o] = e.y; t[o] = base + e.dt; p[o] = e.p; o++;
    }
  }
  return total;
}
