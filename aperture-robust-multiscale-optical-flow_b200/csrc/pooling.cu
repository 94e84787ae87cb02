// pooling.cu -- K4: multi-scale pooling of recent local flows (aperture-robust scale selection).
//
// Replaces computeTrueFlow(x, y, time, pol) (src/vFlow.cpp:952-1210): for the 11 nested squares of
// half-width s = 0,5,...,50 around the event, average |flow|, |flow|cos(theta), |flow|sin(theta) over the
// pixels whose latest event has flow (len > 0) and is younger than 500 us, pick the scale with the
// largest mean |flow| (first maximum) and report that scale's mean vector.
//
// The reference scans 39,611 surface cells per event.  Here the flow events are binned by
// (128-us time slab, 16x16-pixel tile) in stream order; an event only inspects the bins that can hold a
// contributor (<= 5 slabs x the tiles under its 101x101 window), keeps the ones that are still the latest
// event of their pixel (next-at-same-pixel index > i), and accumulates them ONCE into the ring between
// consecutive scales; scale sums are prefix sums over rings.  An empty ring adds exactly nothing, so exact
// ties between scales resolve to the smaller scale like the reference's strict '>' (src/vFlow.cpp:1054).
//
// Flat-index rule: the reference bounds the window's second coordinate by width-1
// (src/vFlow.cpp:1000, 1113) and indexes _data[i*H + j] unchecked (include/EventMatrix.h:32-34), so a
// logical cell (i, j >= H) aliases pixel (i + j/H, j mod H), and indices past W*H read as "no flow".
#include "farms_dev.cuh"

namespace {

constexpr int PW = 4;  // warps per CTA

__global__ void k_cell_keys(const uint16_t *__restrict__ ex, const uint16_t *__restrict__ ey,
                            const uint32_t *__restrict__ em, const uint32_t *__restrict__ excl, size_t m,
                            PoolGeom g, uint32_t *__restrict__ keys, uint32_t *__restrict__ idx,
                            uint32_t *__restrict__ slab_ids) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint32_t sid = em[j] >> FARMS_SLAB_SHIFT;
  const bool first = j == 0 || (em[j - 1] >> FARMS_SLAB_SHIFT) != sid;
  const uint32_t dense = excl[j] + ((j > 0 && first) ? 1u : 0u);
  if (first) slab_ids[dense] = sid;
  const uint32_t tile = (uint32_t)(ex[j] >> g.tile_shift) * (uint32_t)g.nty + (uint32_t)(ey[j] >> g.tile_shift);
  keys[j] = dense * (uint32_t)(g.ntx * g.nty) + tile;
  idx[j] = (uint32_t)j;
}

// rec[pos] = {x | y<<16, t, idx | (len>0)<<31, next}; pay = SoA {len, lcx, lcy}; CSR over cell keys.
__global__ void k_build_records(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ sidx, size_t m,
                                const uint16_t *__restrict__ ex, const uint16_t *__restrict__ ey,
                                const uint32_t *__restrict__ et, const int32_t *__restrict__ nextp,
                                const double *__restrict__ len, const double *__restrict__ lcx,
                                const double *__restrict__ lcy, uint4 *__restrict__ rec, double *__restrict__ pay,
                                uint32_t *__restrict__ cell_start, size_t ncells) {
  size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= m) return;
  const uint32_t j = sidx[pos];
  const double l = len[j];
  rec[pos] = make_uint4((uint32_t)ex[j] | ((uint32_t)ey[j] << 16), et[j], j | (l > 0.0 ? 0x80000000u : 0u),
                        (uint32_t)nextp[j]);
  pay[pos] = l;
  pay[m + pos] = lcx[j];
  pay[2 * m + pos] = lcy[j];
  const long long k = skeys[pos];
  const long long kprev = pos > 0 ? (long long)skeys[pos - 1] : -1ll;
  for (long long c = kprev + 1; c <= k; c++) cell_start[c] = (uint32_t)pos;
  if (pos == m - 1)
    for (long long c = k + 1; c <= (long long)ncells; c++) cell_start[c] = (uint32_t)m;
}

struct PoolArgs {
  const uint4 *rec;
  const double *pay;
  const uint32_t *cell_start;
  const uint32_t *skeys;
  const uint32_t *slab_ids;
  size_t m;
  int h;
  PoolGeom g;
  double *global_r, *global_theta;
  uint8_t *scale;
  unsigned int *work_counter;
  unsigned long long *cand_count;
};

__global__ void __launch_bounds__(PW * 32) k_pooling(PoolArgs A) {
  __shared__ double acc[PW][3][FARMS_NSCALES][32];
  __shared__ uint32_t cnt[PW][FARMS_NSCALES][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = A.g.W, H = A.g.H, ts = A.g.tile_shift, nty = A.g.nty, NT = A.g.ntx * A.g.nty;
  const size_t m = A.m;
  const double *pay_len = A.pay, *pay_cx = A.pay + m, *pay_cy = A.pay + 2 * m;
  unsigned long long ncand = 0;

  for (;;) {
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(A.work_counter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= m) break;
    const size_t pos = (size_t)base + lane;
    uint4 r = make_uint4(0, 0, 0, 0);
    uint32_t key = 0;
    if (pos < m) {
      r = A.rec[pos];
      key = A.skeys[pos];
    }
    // targets: events of this batch (not the halo) with valid local flow.  valid => len > 0 unless
    // the squared speed underflows; such an event falls back to its own (zero) flow below.
    const bool tgt = pos < m && (r.z >> 31) && (int)(r.z & 0x7fffffffu) >= A.h;
    unsigned mask = __ballot_sync(0xffffffffu, tgt);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const int xi = (int)(__shfl_sync(0xffffffffu, r.x, src) & 0xffffu);
      const int yi = (int)(__shfl_sync(0xffffffffu, r.x, src) >> 16);
      const uint32_t ti = __shfl_sync(0xffffffffu, r.y, src);
      const int ii = (int)(__shfl_sync(0xffffffffu, r.z, src) & 0x7fffffffu);
      const int dhi = (int)(__shfl_sync(0xffffffffu, key, src) / (uint32_t)NT);

#pragma unroll
      for (int k = 0; k < FARMS_NSCALES; k++) {
        acc[warp][0][k][lane] = 0.0;
        acc[warp][1][k][lane] = 0.0;
        acc[warp][2][k][lane] = 0.0;
        cnt[warp][k][lane] = 0u;
      }
      // time slabs that can hold an event with |ti - tj| < 500 and index <= ii
      const uint32_t lo_id = (ti >= (uint32_t)(FARMS_KILL_OLD_FLOW_TIME - 1) ? ti - (FARMS_KILL_OLD_FLOW_TIME - 1) : 0u) >>
                             FARMS_SLAB_SHIFT;
      int dlo = dhi;
      while (dlo > 0 && A.slab_ids[dlo - 1] >= lo_id) dlo--;

      const int xlo = max(0, xi - FARMS_MAX_WINDOW), xhi = min(xi + FARMS_MAX_WINDOW, W - 1);   // :998
      const int jlo = max(0, yi - FARMS_MAX_WINDOW), jhi = min(yi + FARMS_MAX_WINDOW, W - 1);   // :1000 (sic)
      const int kmax = jhi >= 0 ? jhi / H : -1;
      for (int k = 0; k <= kmax; k++) {
        // logical cells (ix, j) with j in [k*H, (k+1)*H) live at pixel (ix + k, j - k*H)
        const int pxlo = xlo + k, pxhi = min(xhi + k, W - 1);
        const int jl = max(jlo, k * H), jh = min(jhi, (k + 1) * H - 1);
        if (pxlo > pxhi || jl > jh) continue;
        const int pylo = jl - k * H, pyhi = jh - k * H;
        const int tylo = pylo >> ts, tyhi = pyhi >> ts;
        for (int d = dlo; d <= dhi; d++) {
          for (int tx = pxlo >> ts; tx <= (pxhi >> ts); tx++) {
            const size_t cb = (size_t)d * NT + (size_t)tx * nty;
            const uint32_t s = A.cell_start[cb + tylo], e = A.cell_start[cb + tyhi + 1];
            ncand += (lane == 0) ? (e - s) : 0;
            for (uint32_t p = s + lane; p < e; p += 32) {
              const uint4 c = A.rec[p];
              const int cx = (int)(c.x & 0xffffu), cy = (int)(c.x >> 16);
              const int cj = (int)(c.z & 0x7fffffffu);
              const long long dt = (long long)ti - (long long)c.y;
              const bool ok = (c.z >> 31) && cj <= ii && (int)c.w > ii && cx >= pxlo && cx <= pxhi && cy >= pylo &&
                              cy <= pyhi && dt < FARMS_KILL_OLD_FLOW_TIME && dt > -FARMS_KILL_OLD_FLOW_TIME;  // :1002
              if (ok) {
                const int dx = abs(cx - k - xi), dy = abs(cy + k * H - yi);
                const int ring = (max(dx, dy) + FARMS_WINDOW_JUMP - 1) / FARMS_WINDOW_JUMP;
                acc[warp][0][ring][lane] += pay_len[p];
                acc[warp][1][ring][lane] += pay_cx[p];
                acc[warp][2][ring][lane] += pay_cy[p];
                cnt[warp][ring][lane] += 1u;
              }
            }
          }
        }
      }
      __syncwarp();
      // lane k < 11 reduces ring k over the 32 per-lane partials (rotated start: no bank conflicts)
      double rl = 0.0, rx = 0.0, ry = 0.0;
      uint32_t rn = 0;
      if (lane < FARMS_NSCALES) {
        for (int q = 0; q < 32; q++) {
          const int qq = (q + lane) & 31;
          rl += acc[warp][0][lane][qq];
          rx += acc[warp][1][lane][qq];
          ry += acc[warp][2][lane][qq];
          rn += cnt[warp][lane][qq];
        }
      }
      __syncwarp();
      // prefix over rings = nested squares; arg-max of the mean length, first maximum (:1047-1059)
      double Sl = 0.0, Sx = 0.0, Sy = 0.0, best = 0.0, bvx = 0.0, bvy = 0.0;
      uint32_t Sn = 0;
      int bk = -1;
#pragma unroll
      for (int k = 0; k < FARMS_NSCALES; k++) {
        const uint32_t n_k = __shfl_sync(0xffffffffu, rn, k);
        const double l_k = __shfl_sync(0xffffffffu, rl, k);
        const double x_k = __shfl_sync(0xffffffffu, rx, k);
        const double y_k = __shfl_sync(0xffffffffu, ry, k);
        if (n_k) {  // an empty ring leaves the sums bit-identical
          Sl += l_k;
          Sx += x_k;
          Sy += y_k;
          Sn += n_k;
        }
        const double dn = (double)Sn;
        const double mean = Sn ? Sl / dn : 0.0;  // :1023-1036
        if (mean > best) {
          best = mean;
          bk = k;
          bvx = Sx / dn;
          bvy = Sy / dn;
        }
      }
      if (bk < 0) {  // :1085-1094 fallback: the event's own flow
        bvx = __shfl_sync(0xffffffffu, pos < m ? pay_cx[pos] : 0.0, src);
        bvy = __shfl_sync(0xffffffffu, pos < m ? pay_cy[pos] : 0.0, src);
        bk = 0;
      }
      if (lane == 0) {
        const int o = ii - A.h;
        A.global_r[o] = __dsqrt_rn(__dadd_rn(__dmul_rn(bvy, bvy), __dmul_rn(bvx, bvx)));  // src/vFlow.cpp:365
        A.global_theta[o] = atan2(bvy, bvx);                                               // :366
        A.scale[o] = (uint8_t)(bk * FARMS_WINDOW_JUMP);
      }
    }
  }
  if (lane == 0 && ncand) atomicAdd(A.cand_count, ncand);
}

inline unsigned nb(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void launch_cell_keys(const uint16_t *ex, const uint16_t *ey, const uint32_t *em, const uint32_t *excl, size_t m,
                      PoolGeom g, uint32_t *keys, uint32_t *idx, uint32_t *slab_ids, cudaStream_t s) {
  if (m) k_cell_keys<<<nb(m, 256), 256, 0, s>>>(ex, ey, em, excl, m, g, keys, idx, slab_ids);
}

void launch_build_records(const uint32_t *skeys, const uint32_t *sidx, size_t m, const uint16_t *ex,
                          const uint16_t *ey, const uint32_t *et, const int32_t *nextp, const double *len,
                          const double *lcx, const double *lcy, uint4 *rec, double *pay, uint32_t *cell_start,
                          size_t ncells, cudaStream_t s) {
  if (m) k_build_records<<<nb(m, 256), 256, 0, s>>>(skeys, sidx, m, ex, ey, et, nextp, len, lcx, lcy, rec, pay,
                                                   cell_start, ncells);
}

void launch_pooling(const uint4 *rec, const double *pay, const uint32_t *cell_start, const uint32_t *skeys,
                    const uint32_t *slab_ids, size_t m, int h, PoolGeom g, double *global_r, double *global_theta,
                    uint8_t *scale, unsigned int *work_counter, unsigned long long *cand_count, int num_sms,
                    cudaStream_t s) {
  if (!m) return;
  PoolArgs A;
  A.rec = rec; A.pay = pay; A.cell_start = cell_start; A.skeys = skeys; A.slab_ids = slab_ids;
  A.m = m; A.h = h; A.g = g;
  A.global_r = global_r; A.global_theta = global_theta; A.scale = scale;
  A.work_counter = work_counter; A.cand_count = cand_count;
  // persistent warps pulling 32-slot groups from a global counter: grid = SMs x resident CTAs
  unsigned grid = (unsigned)num_sms * 5u;
  unsigned need = nb(m, 32 * PW);
  if (grid > need) grid = need;
  k_pooling<<<grid, PW * 32, 0, s>>>(A);
}
