// pooling.cu -- K4: multi-scale pooling of recent local flows (aperture-robust scale selection).
//
// Replaces computeTrueFlow(x, y, time, pol) (src/vFlow.cpp:952-1210): for the 11 nested squares of
// half-width s = 0,5,...,50 around the event, average |flow|, |flow|cos(theta), |flow|sin(theta) over the
// pixels whose latest event has flow (len > 0) and is younger than 500 us, pick the scale with the
// largest mean |flow| (first maximum) and report that scale's mean vector.
//
// The reference scans 39,611 surface cells per event.  Here the events that HAVE flow are binned by
// (128-us time slab, 16x16-pixel tile) in stream order (the pooling index: rec/pay sorted by cell + CSR).
// A contributor of event i is a flow event j with  j <= i < next_same_pixel(j)  (it is still the latest
// event of its pixel), |t_i - t_j| < 500 and inside the window; each one is added ONCE, into the ring
// between consecutive scales, and the scale sums are prefix sums over rings.  An empty ring adds exactly
// nothing, so exact ties resolve to the smaller scale like the reference's strict '>' (src/vFlow.cpp:1054).
//
// Two kernels share that definition:
//   k_pool_tile  -- the fast path.  A CTA owns a 32x32-pixel tile for a run of time slabs and keeps the flow
//                   events of the (32+100)^2 region of the last <= 5 slabs staged in shared memory; every
//                   warp pools one event of the tile against the staged set.  Needs sorted timestamps
//                   (so that the age test folds into an index bound) and windows that stay inside rows < H.
//   k_pool_any   -- the general path straight from the global index: unsorted timestamps, windows whose
//                   second coordinate runs past H (the reference's width-1 bound, below), overflowed tiles.
//
// Flat-index rule: the reference bounds the window's second coordinate by width-1
// (src/vFlow.cpp:1000, 1113) and indexes _data[i*H + j] unchecked (include/EventMatrix.h:32-34), so a
// logical cell (i, j >= H) aliases pixel (i + j/H, j mod H), and indices past W*H read as "no flow".
#include <algorithm>

#include "farms_dev.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// index construction
// ------------------------------------------------------------------------------------------------

// Only events WITH flow (len > 0) can contribute (src/vFlow.cpp:1002) or be pooled (src/vFlow.cpp:315,
// 362), so the pooling index holds just those; the others get the sentinel key `ncells` and sort last.
__global__ void k_cell_keys(const uint16_t *__restrict__ ex, const uint16_t *__restrict__ ey,
                            const uint32_t *__restrict__ em, const uint32_t *__restrict__ excl,
                            const double *__restrict__ len, size_t m, PoolGeom g, uint32_t ncells,
                            uint32_t *__restrict__ keys, uint32_t *__restrict__ idx,
                            uint32_t *__restrict__ slab_ids) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint32_t sid = em[j] >> FARMS_SLAB_SHIFT;
  const bool first = j == 0 || (em[j - 1] >> FARMS_SLAB_SHIFT) != sid;
  const uint32_t dense = excl[j] + ((j > 0 && first) ? 1u : 0u);
  if (first) slab_ids[dense] = sid;
  const uint32_t tile = (uint32_t)(ex[j] >> g.tile_shift) * (uint32_t)g.nty + (uint32_t)(ey[j] >> g.tile_shift);
  keys[j] = len[j] > 0.0 ? dense * (uint32_t)(g.ntx * g.nty) + tile : ncells;
  idx[j] = (uint32_t)j;
}

// rec[pos] = {x | y<<16, t, idx, end}; pay = SoA {len, lcx, lcy}; CSR cell_start over cell keys.
// end = min(next event at the same pixel, first event that is >= 500 us younger): for sorted timestamps
// "j is a contributor of i" is exactly  idx <= i < end.  (For unsorted input end = next and k_pool_any
// tests the age itself.)
__global__ void k_build_records(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ sidx, size_t m,
                                const uint16_t *__restrict__ ex, const uint16_t *__restrict__ ey,
                                const uint32_t *__restrict__ et, const int32_t *__restrict__ nextp,
                                const double *__restrict__ len, const double *__restrict__ lcx,
                                const double *__restrict__ lcy, int monotone, uint4 *__restrict__ rec,
                                double *__restrict__ pay, uint32_t *__restrict__ cell_start, uint32_t ncells) {
  size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= m) return;
  const uint32_t k = skeys[pos];
  if (k >= ncells) return;  // no flow: not part of the index
  const uint32_t j = sidx[pos];
  uint32_t end = (uint32_t)nextp[j];
  const uint32_t tj = et[j];
  if (monotone) {
    // first index u > j with et[u] >= tj + 500
    uint32_t lo = j + 1, hi = end < (uint32_t)m ? end : (uint32_t)m;
    const uint64_t lim = (uint64_t)tj + FARMS_KILL_OLD_FLOW_TIME;
    while (lo < hi) {
      uint32_t mid = lo + (hi - lo) / 2;
      if ((uint64_t)et[mid] >= lim) hi = mid; else lo = mid + 1;
    }
    if (lo < end) end = lo;
  }
  rec[pos] = make_uint4((uint32_t)ex[j] | ((uint32_t)ey[j] << 16), tj, j, end);
  pay[pos] = len[j];
  pay[m + pos] = lcx[j];
  pay[2 * m + pos] = lcy[j];
  const long long kprev = pos > 0 ? (long long)skeys[pos - 1] : -1ll;
  for (long long c = kprev + 1; c <= (long long)k; c++) cell_start[c] = (uint32_t)pos;
  if (pos == m - 1 || skeys[pos + 1] >= ncells)
    for (long long c = (long long)k + 1; c <= (long long)ncells; c++) cell_start[c] = (uint32_t)(pos + 1);
}

struct PoolArgs {
  const uint4 *rec;
  const double *pay;
  const uint32_t *cell_start;
  const uint32_t *slab_ids;
  uint8_t *done;       // per index position: 1 once the fast path has pooled that event
  size_t m;            // stride of the pay arrays
  uint32_t ncells;     // cell_start[ncells] = entries in the index (events with flow)
  int h;               // halo events (indices < h) are contributors only
  int nslabs;
  PoolGeom g;
  double *global_r, *global_theta;
  uint8_t *scale;
  unsigned int *work_counter;
  unsigned long long *cand_count;
};

// ring sums of one event -> nested-square means -> arg-max scale -> outputs.  rl/rx/ry/rn hold ring k's
// totals in lane k (k < 11).  All lanes must call.  The prefix over rings is sequential (an empty ring adds
// exactly 0.0, so a scale whose outer ring is empty has bit-identical sums and loses the strict '>').
template <int WIDTH>
__device__ __forceinline__ void finish_event(const PoolArgs &A, int lane, double rl, double rx, double ry,
                                             double rn, double own_cx, double own_cy, int out_index, bool write) {
  double Sl = 0.0, Sx = 0.0, Sy = 0.0, Sn = 0.0;
  double myl = 0.0, myx = 0.0, myy = 0.0, myn = 0.0;
#pragma unroll
  for (int k = 0; k < FARMS_NSCALES; k++) {
    Sn += __shfl_sync(0xffffffffu, rn, k, WIDTH);
    Sl += __shfl_sync(0xffffffffu, rl, k, WIDTH);
    Sx += __shfl_sync(0xffffffffu, rx, k, WIDTH);
    Sy += __shfl_sync(0xffffffffu, ry, k, WIDTH);
    if (lane == k) {
      myl = Sl; myx = Sx; myy = Sy; myn = Sn;
    }
  }
  // lane k: mean length of scale k (src/vFlow.cpp:1023-1036)
  const double mean = (lane < FARMS_NSCALES && myn > 0.0) ? myl / myn : 0.0;
  // arg-max with strict '>' from 0, first maximum (:1047-1059)
  double best = 0.0;
  int bk = -1;
#pragma unroll
  for (int k = 0; k < FARMS_NSCALES; k++) {
    const double mk = __shfl_sync(0xffffffffu, mean, k, WIDTH);
    if (mk > best) {
      best = mk;
      bk = k;
    }
  }
  const int srcl = bk < 0 ? 0 : bk;
  const double wx = __shfl_sync(0xffffffffu, myx, srcl, WIDTH), wy = __shfl_sync(0xffffffffu, myy, srcl, WIDTH),
               wn = __shfl_sync(0xffffffffu, myn, srcl, WIDTH);
  if (lane == 0 && write) {
    double bvx, bvy;
    if (bk < 0) {  // :1085-1094 fallback: the event's own flow
      bvx = own_cx;
      bvy = own_cy;
      bk = 0;
    } else {
      bvx = wx / wn;
      bvy = wy / wn;
    }
    A.global_r[out_index] = __dsqrt_rn(__dadd_rn(__dmul_rn(bvy, bvy), __dmul_rn(bvx, bvx)));  // src/vFlow.cpp:365
    A.global_theta[out_index] = atan2(bvy, bvx);                                               // :366
    A.scale[out_index] = (uint8_t)(bk * FARMS_WINDOW_JUMP);
  }
}

// ------------------------------------------------------------------------------------------------
// general path
// ------------------------------------------------------------------------------------------------
constexpr int PW = 4;  // warps per CTA

__global__ void __launch_bounds__(PW * 32) k_pool_any(PoolArgs A) {
  __shared__ double acc[PW][4][FARMS_NSCALES][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = A.g.W, H = A.g.H, ts = A.g.tile_shift, nty = A.g.nty, NT = A.g.ntx * A.g.nty;
  const size_t m = A.m;
  const uint32_t mi = A.cell_start[A.ncells];
  const double *pay_len = A.pay, *pay_cx = A.pay + m, *pay_cy = A.pay + 2 * m;
  unsigned long long ncand = 0;

  for (;;) {
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(A.work_counter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= mi) break;
    const uint32_t pos = base + lane;
    uint4 r = make_uint4(0, 0, 0, 0);
    bool tgt = false;
    if (pos < mi) {
      r = A.rec[pos];
      tgt = (int)r.z >= A.h && !A.done[pos];
    }
    unsigned mask = __ballot_sync(0xffffffffu, tgt);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const uint32_t rx_ = __shfl_sync(0xffffffffu, r.x, src);
      const int xi = (int)(rx_ & 0xffffu), yi = (int)(rx_ >> 16);
      const uint32_t ti = __shfl_sync(0xffffffffu, r.y, src);
      const int ii = (int)__shfl_sync(0xffffffffu, r.z, src);
      const uint32_t tpos = base + src;
      // dense slab of the target: the cell whose CSR range holds tpos -- recover it from the event's time
      // slab id by walking the (short) slab table downwards from the end is not possible here, so the
      // builder stored it implicitly: binary search the slab whose first cell starts at or before tpos.
      int dhi;
      {
        int lo = 0, hi = A.nslabs - 1;
        while (lo < hi) {
          int mid = (lo + hi + 1) >> 1;
          if (A.cell_start[(size_t)mid * NT] <= tpos) lo = mid; else hi = mid - 1;
        }
        dhi = lo;
      }
#pragma unroll
      for (int k = 0; k < FARMS_NSCALES; k++) {
        acc[warp][0][k][lane] = 0.0;
        acc[warp][1][k][lane] = 0.0;
        acc[warp][2][k][lane] = 0.0;
        acc[warp][3][k][lane] = 0.0;
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
              const long long dt = (long long)ti - (long long)c.y;
              const bool ok = (int)c.z <= ii && (int)c.w > ii && cx >= pxlo && cx <= pxhi && cy >= pylo &&
                              cy <= pyhi && dt < FARMS_KILL_OLD_FLOW_TIME && dt > -FARMS_KILL_OLD_FLOW_TIME;  // :1002
              if (ok) {
                const int dx = abs(cx - k - xi), dy = abs(cy + k * H - yi);
                const int ring = (max(dx, dy) + FARMS_WINDOW_JUMP - 1) / FARMS_WINDOW_JUMP;
                acc[warp][0][ring][lane] += pay_len[p];
                acc[warp][1][ring][lane] += pay_cx[p];
                acc[warp][2][ring][lane] += pay_cy[p];
                acc[warp][3][ring][lane] += 1.0;
              }
            }
          }
        }
      }
      __syncwarp();
      // lane k < 11 reduces ring k over the 32 per-lane partials (rotated start: no bank conflicts)
      double rl = 0.0, rx = 0.0, ry = 0.0, rn = 0.0;
      if (lane < FARMS_NSCALES) {
        for (int q = 0; q < 32; q++) {
          const int qq = (q + lane) & 31;
          rl += acc[warp][0][lane][qq];
          rx += acc[warp][1][lane][qq];
          ry += acc[warp][2][lane][qq];
          rn += acc[warp][3][lane][qq];
        }
      }
      __syncwarp();
      finish_event<32>(A, lane, rl, rx, ry, rn, pay_cx[tpos], pay_cy[tpos], ii - A.h, true);
    }
  }
  if (lane == 0 && ncand) atomicAdd(A.cand_count, ncand);
}

// ------------------------------------------------------------------------------------------------
// fast path: owner tiles with shared-memory staging, FP32 ring partials, exact decisions
// ------------------------------------------------------------------------------------------------
// Ring partial sums are kept in FP32 (half the shared memory of FP64 => twice the resident warps) and
// combined in FP64.  That perturbs a scale's mean by < 1e-5 relative, so an event is only finished here when
// its arg-max over scales is decided by a margin > 2e-5 and its mean vector is not a cancellation residue;
// everything else (measured: well under 1 % of events) is left to k_pool_any, which is exact.  Scales whose
// extra rings are empty have bit-identical sums in both arithmetics, so exact ties behave like the reference.
constexpr int OT_SHIFT = 5, OT = 1 << OT_SHIFT;  // owner tile edge (pixels)
constexpr int TK_NSL = 4;      // consecutive slabs pooled per round (more events per round => fewer, fuller waves)
constexpr int TK_RING = 4 + TK_NSL;  // a 500-us window touches <= 5 slabs, so a round's windows span <= 4 + NSL
constexpr int TK_PAD = 64;     // the pooling loop reads 4 x 16 records at a time without bounds checks
#ifndef FARMS_TK_SEG
#define FARMS_TK_SEG 64
#endif
constexpr int TK_SEG = FARMS_TK_SEG;  // slabs per work item
constexpr int TK_MAXT = 128;   // targets handled per round and slab
constexpr int TK_MAXRUN = 24;  // tile-column runs of a region: <= 10 for rows < H plus <= 10 aliased
constexpr float TK_TIE_TOL = 2e-5f;

template <int WARPS, int CAP>
struct TileSmem {
  uint4 ra[TK_RING][CAP + TK_PAD];        // {x | y<<16 (logical window coordinates), idx, end - idx, len as f32}
  float2 rb[TK_RING][CAP];                // lcx, lcy
  float4 acc[WARPS][FARMS_NSCALES][32];   // per-lane ring partials: len, lcx, lcy, count
  uint32_t tlist[TK_NSL][TK_MAXT];
  uint32_t run_s[TK_MAXRUN], run_o[TK_MAXRUN + 1];
  uint32_t wcount[WARPS];
  int tag[TK_RING];
  int count[TK_RING];
  int overflow[TK_RING];
  unsigned int ntg[TK_NSL], tnext, item;
};

struct Region {  // pixels an owner tile can reach, as physical rectangles
  int rx0, rx1, ry0, ry1;  // rows < H (k = 0)
  int ax0, ax1, ay1;       // aliased part (k = 1): physical x in [ax0, ax1], y in [0, ay1]; empty if ay1 < 0
};

// Stage the flow events of dense slab `s` inside region R into ring slot `slot`, preserving index order
// (ordered compaction => deterministic summation order).  Aliased events are stored with their LOGICAL
// window coordinates (x - 1, y + H) so that the pooling loop needs no special case.
template <class SM, int WARPS, int CAP>
__device__ void stage_slab(const PoolArgs &A, SM &S, int s, int slot, const Region &R) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ts = A.g.tile_shift, nty = A.g.nty, NT = A.g.ntx * A.g.nty, H = A.g.H;
  const int tx0 = R.rx0 >> ts, tx1 = R.rx1 >> ts, ty0 = R.ry0 >> ts, ty1 = R.ry1 >> ts;
  const int nrun0 = tx1 - tx0 + 1;
  const int atx0 = R.ax0 >> ts, atx1 = R.ax1 >> ts;
  const int nrun1 = R.ay1 >= 0 ? atx1 - atx0 + 1 : 0;
  const int nrun = nrun0 + nrun1;
  __syncthreads();  // previous users of run_s/run_o/wcount and of this slot are done
  if (tid < nrun) {
    const int c = tid;
    uint32_t a, b;
    if (c < nrun0) {
      const size_t cb = (size_t)s * NT + (size_t)(tx0 + c) * nty;
      a = A.cell_start[cb + ty0];
      b = A.cell_start[cb + ty1 + 1];
    } else {
      const size_t cb = (size_t)s * NT + (size_t)(atx0 + c - nrun0) * nty;
      a = A.cell_start[cb];
      b = A.cell_start[cb + (R.ay1 >> ts) + 1];
    }
    S.run_s[c] = a;
    S.run_o[c + 1] = b - a;  // lengths first, prefix below
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t o = 0;
    S.run_o[0] = 0;
    for (int c = 0; c < nrun; c++) {
      o += S.run_o[c + 1];
      S.run_o[c + 1] = o;
    }
  }
  __syncthreads();
  const uint32_t total = S.run_o[nrun];
  const double *pay_len = A.pay, *pay_cx = A.pay + A.m, *pay_cy = A.pay + 2 * A.m;
  uint32_t out_base = 0;
  for (uint32_t r0 = 0; r0 < total; r0 += WARPS * 32) {
    const uint32_t f = r0 + tid;
    bool pass = false;
    uint32_t pos = 0;
    uint4 rec = make_uint4(0, 0, 0, 0);
    if (f < total) {
      int c = 0;
      while (c + 1 < nrun && S.run_o[c + 1] <= f) c++;
      pos = S.run_s[c] + (f - S.run_o[c]);
      rec = A.rec[pos];
      int x = (int)(rec.x & 0xffffu), y = (int)(rec.x >> 16);
      if (c < nrun0) {
        pass = x >= R.rx0 && x <= R.rx1 && y >= R.ry0 && y <= R.ry1;
      } else {
        pass = x >= R.ax0 && x <= R.ax1 && y <= R.ay1;
        x -= 1;
        y += H;
      }
      rec.x = (uint32_t)x | ((uint32_t)y << 16);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) S.wcount[warp] = __popc(bal);
    __syncthreads();
    uint32_t pre = 0, all = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
      const uint32_t cw = S.wcount[w];
      if (w < warp) pre += cw;
      all += cw;
    }
    const uint32_t o = out_base + pre + __popc(bal & ((1u << lane) - 1u));
    if (pass && o < (uint32_t)CAP) {
      S.ra[slot][o] = make_uint4(rec.x, rec.z, rec.w - rec.z, __float_as_uint(__double2float_rn(pay_len[pos])));
      S.rb[slot][o] = make_float2(__double2float_rn(pay_cx[pos]), __double2float_rn(pay_cy[pos]));
    }
    out_base += all;
    __syncthreads();
  }
  const uint32_t cnt = min(out_base, (uint32_t)CAP);
  // entries the unrolled loop may touch past the end: span 0 never passes
  if (tid < TK_PAD) S.ra[slot][cnt + tid] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    S.tag[slot] = s;
    S.count[slot] = (int)cnt;
    S.overflow[slot] = out_base > (uint32_t)CAP;
  }
}

// Like finish_event, but from FP32 partial sums: writes only when the decision is safe; returns (in every lane
// of the segment) whether the event was finished.
__device__ __forceinline__ bool finish_event_checked(const PoolArgs &A, int sub, double rl, double rx, double ry,
                                                     double rn, int out_index, bool have) {
  double Sl = 0.0, Sx = 0.0, Sy = 0.0, Sn = 0.0;
  double myl = 0.0, myx = 0.0, myy = 0.0, myn = 0.0;
#pragma unroll
  for (int k = 0; k < FARMS_NSCALES; k++) {
    Sn += __shfl_sync(0xffffffffu, rn, k, 16);
    Sl += __shfl_sync(0xffffffffu, rl, k, 16);
    Sx += __shfl_sync(0xffffffffu, rx, k, 16);
    Sy += __shfl_sync(0xffffffffu, ry, k, 16);
    if (sub == k) {
      myl = Sl; myx = Sx; myy = Sy; myn = Sn;
    }
  }
  const double mean = (sub < FARMS_NSCALES && myn > 0.0) ? myl / myn : 0.0;
  double best = 0.0, bn = 0.0;
  int bk = -1;
#pragma unroll
  for (int k = 0; k < FARMS_NSCALES; k++) {
    const double mk = __shfl_sync(0xffffffffu, mean, k, 16);
    const double nk = __shfl_sync(0xffffffffu, myn, k, 16);
    if (mk > best) {
      best = mk;
      bk = k;
      bn = nk;
    }
  }
  // is any other scale (with a different contributor set) within the FP32 noise of the winner?
  const bool rival = sub < FARMS_NSCALES && myn != bn && fabs(mean - best) <= (double)TK_TIE_TOL * best;
  const unsigned seg = 0xffffu << (threadIdx.x & 16);
  const bool any_rival = (__ballot_sync(0xffffffffu, rival) & seg) != 0u;
  const int srcl = bk < 0 ? 0 : bk;
  const double wx = __shfl_sync(0xffffffffu, myx, srcl, 16), wy = __shfl_sync(0xffffffffu, myy, srcl, 16),
               wn = __shfl_sync(0xffffffffu, myn, srcl, 16);
  bool safe = have && bk >= 0 && !any_rival && best > 1e-30 && best < 1e30;
  double bvx = 0.0, bvy = 0.0, r2 = 0.0;
  if (safe) {
    bvx = wx / wn;
    bvy = wy / wn;
    r2 = __dadd_rn(__dmul_rn(bvy, bvy), __dmul_rn(bvx, bvx));
    // mean vector much shorter than the mean length: the FP32 sums cancelled, let the exact path do it
    safe = r2 > 1e-4 * best * best;
  }
  if (sub == 0 && safe) {
    A.global_r[out_index] = __dsqrt_rn(r2);        // src/vFlow.cpp:365
    A.global_theta[out_index] = atan2(bvy, bvx);   // :366
    A.scale[out_index] = (uint8_t)(bk * FARMS_WINDOW_JUMP);
  }
  return safe;
}

template <int WARPS, int CAP>
__global__ void __launch_bounds__(WARPS * 32, 1) k_pool_tile(PoolArgs A, int otx_n, int oty_n, int nseg) {
  using SM = TileSmem<WARPS, CAP>;
  constexpr int THREADS = WARPS * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4, sub = lane & 15;
  const int W = A.g.W, H = A.g.H, nty = A.g.nty, NT = A.g.ntx * A.g.nty;
  const unsigned int nitems = (unsigned int)otx_n * oty_n * nseg;
  unsigned long long ncand = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) S.item = atomicAdd(A.work_counter, 1u);
    if (tid < TK_RING) S.tag[tid] = -1;
    __syncthreads();
    const unsigned int item = S.item;
    if (item >= nitems) break;
    // items are ordered segment-major so that CTAs running together work on the same time span (L2 reuse)
    const int seg = item / (otx_n * oty_n), ot = item % (otx_n * oty_n);
    const int TX = ot / oty_n, TY = ot % oty_n;
    const int X0 = TX << OT_SHIFT, Y0 = TY << OT_SHIFT;
    Region R;
    R.rx0 = max(X0 - FARMS_MAX_WINDOW, 0);
    R.rx1 = min(X0 + OT - 1 + FARMS_MAX_WINDOW, W - 1);                       // src/vFlow.cpp:998
    R.ry0 = max(Y0 - FARMS_MAX_WINDOW, 0);
    const int jmax = min(min(Y0 + OT - 1, H - 1) + FARMS_MAX_WINDOW, W - 1);  // :1000 (sic: width - 1)
    R.ry1 = min(jmax, H - 1);
    // logical rows j in [H, 2H) alias pixel (i + 1, j - H); rows >= 2H are left to k_pool_any
    R.ay1 = min(jmax, 2 * H - 1) - H;
    R.ax0 = R.rx0 + 1;
    R.ax1 = min(R.rx1 + 1, W - 1);
    if (R.ax0 > R.ax1) R.ay1 = -1;
    // index tiles (16x16) of the owner tile: 2 columns x 2 rows, clipped
    const int itx0 = X0 >> 4, itx1 = min((X0 + OT - 1) >> 4, A.g.ntx - 1);
    const int ity0 = Y0 >> 4, ity1 = min((Y0 + OT - 1) >> 4, nty - 1);
    const int d_begin = seg * TK_SEG, d_end = min(d_begin + TK_SEG, A.nslabs);

    // TK_NSL consecutive slabs per round
    for (int d = d_begin; d < d_end; d += TK_NSL) {
      const int nd = min(TK_NSL, d_end - d);
      // ---- targets of this round: flow events of the owner tile in slabs d .. d+nd-1 ----
      uint32_t ta[TK_NSL][2], tb[TK_NSL][2], nraw[TK_NSL];
      uint32_t nraw_all = 0, nmax = 0;
#pragma unroll
      for (int w = 0; w < TK_NSL; w++) {
        nraw[w] = 0;
#pragma unroll
        for (int c = 0; c < 2; c++) {
          ta[w][c] = tb[w][c] = 0;
          if (w < nd && itx0 + c <= itx1) {
            const size_t cb = (size_t)(d + w) * NT + (size_t)(itx0 + c) * nty;
            ta[w][c] = A.cell_start[cb + ity0];
            tb[w][c] = A.cell_start[cb + ity1 + 1];
          }
          nraw[w] += tb[w][c] - ta[w][c];
        }
        nraw_all += nraw[w];
        nmax = max(nmax, nraw[w]);
      }
      if (nraw_all == 0) continue;  // uniform across the CTA

      // ---- make sure the slabs of all windows of the round are staged ----
      int dlo[TK_NSL], dhi[TK_NSL];
      int s_first = 0x7fffffff, s_last = -1;
#pragma unroll
      for (int w = 0; w < TK_NSL; w++) {
        const int dd = min(d + w, d_end - 1);
        const uint32_t t_first = A.slab_ids[dd] << FARMS_SLAB_SHIFT;
        const uint32_t lo_id =
            (t_first >= (uint32_t)(FARMS_KILL_OLD_FLOW_TIME - 1) ? t_first - (FARMS_KILL_OLD_FLOW_TIME - 1) : 0u) >> FARMS_SLAB_SHIFT;
        int l = dd;
        while (l > 0 && dd - l < 4 && A.slab_ids[l - 1] >= lo_id) l--;
        dlo[w] = l;
        dhi[w] = dd;
        if (nraw[w]) {
          s_first = min(s_first, l);
          s_last = max(s_last, dd);
        }
      }
      for (int s = s_first; s <= s_last; s++) {
        const int slot = s % TK_RING;
        if (S.tag[slot] != s) stage_slab<SM, WARPS, CAP>(A, S, s, slot, R);  // uniform: tag is read after a barrier
        __syncthreads();
      }
      bool ovf[TK_NSL];
#pragma unroll
      for (int w = 0; w < TK_NSL; w++) {
        ovf[w] = false;
        for (int s = dlo[w]; s <= dhi[w]; s++) ovf[w] |= S.overflow[s % TK_RING] != 0;
      }

      for (uint32_t t0 = 0; t0 < nmax; t0 += TK_MAXT) {
        __syncthreads();
        if (tid < TK_NSL) S.ntg[tid] = 0;
        if (tid == 0) S.tnext = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < TK_NSL; w++)
          for (uint32_t f = t0 + tid; f < min(nraw[w], t0 + TK_MAXT); f += THREADS) {
            const uint32_t n0 = tb[w][0] - ta[w][0];
            const uint32_t pos = f < n0 ? ta[w][0] + f : ta[w][1] + (f - n0);
            const uint4 r = A.rec[pos];
            const int yi = (int)(r.x >> 16);
            // fast-path conditions: not a halo event, window rows stay below 2H, staging complete
            const bool ok = (int)r.z >= A.h && min(yi + FARMS_MAX_WINDOW, W - 1) <= 2 * H - 1 && !ovf[w];
            if (ok) S.tlist[w][atomicAdd(&S.ntg[w], 1u)] = pos;
          }
        __syncthreads();
        // tasks = pairs of targets of the same slab (both halves of a warp then share their loop bounds)
        uint32_t tstart[TK_NSL + 1];
        tstart[0] = 0;
#pragma unroll
        for (int w = 0; w < TK_NSL; w++) tstart[w + 1] = tstart[w] + ((S.ntg[w] + 1) >> 1);
        const uint32_t tasks = tstart[TK_NSL];

        // ---- two targets per warp: lanes 0-15 pool one event, lanes 16-31 the next ----
        for (;;) {
          uint32_t k = 0;
          if (lane == 0) k = atomicAdd(&S.tnext, 1u);
          k = __shfl_sync(0xffffffffu, k, 0);
          if (k >= tasks) break;
          int w = 0;
#pragma unroll
          for (int q = 1; q < TK_NSL; q++) w += (k >= tstart[q]) ? 1 : 0;
          const uint32_t kk = (k - tstart[w]) * 2 + half, nt = S.ntg[w];
          const bool have = kk < nt;
          const uint32_t tpos = S.tlist[w][have ? kk : nt - 1];
          const uint4 r = A.rec[tpos];
          const int xi = (int)(r.x & 0xffffu), yi = (int)(r.x >> 16);
          const uint32_t ii = r.z;
          const int jlo = max(0, yi - FARMS_MAX_WINDOW), jhi = min(yi + FARMS_MAX_WINDOW, W - 1);  // :1000 (sic)
          const bool rows_ok = have && jhi >= jlo;  // jhi < jlo when W < H: no cell qualifies
          const uint32_t jspan = rows_ok ? (uint32_t)(jhi - jlo) : 0u;
          const int ylo = rows_ok ? jlo : 0x7fff0000;  // makes (cy - ylo) huge => fails
          const int xoff = FARMS_MAX_WINDOW - xi;
#pragma unroll
          for (int q = 0; q < FARMS_NSCALES; q++) S.acc[warp][q][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
          int sl = dlo[0], sh = dhi[0];
#pragma unroll
          for (int q = 1; q < TK_NSL; q++)
            if (w == q) {
              sl = dlo[q];
              sh = dhi[q];
            }
          for (int s = sl; s <= sh; s++) {
            const int slot = s % TK_RING;
            const int n = S.count[slot];
            ncand += (sub == 0 && have) ? n : 0;
            for (int q0 = sub; q0 < n; q0 += 64) {
              uint4 c[4];
#pragma unroll
              for (int u = 0; u < 4; u++) c[u] = S.ra[slot][q0 + 16 * u];  // padded: no bounds check
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const int cx = (int)(c[u].x & 0xffffu), cy = (int)(c[u].x >> 16);
                // still the latest event of its pixel AND younger than 500 us  <=>  idx <= ii < end
                const bool ok = (ii - c[u].y) < c[u].z && (uint32_t)(cx + xoff) <= 2u * FARMS_MAX_WINDOW &&
                                (uint32_t)(cy - ylo) <= jspan;
                if (ok) {
                  const int m = max(abs(cx - xi), abs(cy - yi));
                  const int ring = ((m + FARMS_WINDOW_JUMP - 1) * 205) >> 10;  // /5 for values <= 54
                  const float2 l = S.rb[slot][q0 + 16 * u];
                  float4 v = S.acc[warp][ring][lane];
                  v.x += __uint_as_float(c[u].w);
                  v.y += l.x;
                  v.z += l.y;
                  v.w += 1.f;
                  S.acc[warp][ring][lane] = v;
                }
              }
            }
          }
          __syncwarp();
          // sub-lane k < 11 of each half combines ring k's 16 per-lane partials in FP64
          double rl = 0.0, rx = 0.0, ry = 0.0, rn = 0.0;
          if (sub < FARMS_NSCALES) {
#pragma unroll 8
            for (int q = 0; q < 16; q++) {
              const float4 v = S.acc[warp][sub][(half << 4) | ((q + sub) & 15)];
              rl += (double)v.x;
              rx += (double)v.y;
              ry += (double)v.z;
              rn += (double)v.w;
            }
          }
          __syncwarp();
          const bool fin = finish_event_checked(A, sub, rl, rx, ry, rn, (int)ii - A.h, have);
          if (sub == 0 && fin) A.done[tpos] = 1;
        }
      }
    }
  }
  if ((lane & 15) == 0 && ncand) atomicAdd(A.cand_count, ncand);
}

template <int WARPS, int CAP>
void launch_tile(const PoolArgs &A0, int nslabs, int num_sms, cudaStream_t s) {
  PoolArgs A = A0;
  using SM = TileSmem<WARPS, CAP>;
  auto kern = k_pool_tile<WARPS, CAP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
    attr_set = true;
  }
  const int otx = (A.g.W + OT - 1) >> OT_SHIFT, oty = (A.g.H + OT - 1) >> OT_SHIFT;
  const int nseg = (nslabs + TK_SEG - 1) / TK_SEG;
  const long long items = (long long)otx * oty * nseg;
  unsigned grid = (unsigned)std::min<long long>(items, (long long)num_sms);
  kern<<<grid, WARPS * 32, sizeof(SM), s>>>(A, otx, oty, nseg);
}

inline unsigned nb(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void launch_cell_keys(const uint16_t *ex, const uint16_t *ey, const uint32_t *em, const uint32_t *excl,
                      const double *len, size_t m, PoolGeom g, uint32_t ncells, uint32_t *keys, uint32_t *idx,
                      uint32_t *slab_ids, cudaStream_t s) {
  if (m) k_cell_keys<<<nb(m, 256), 256, 0, s>>>(ex, ey, em, excl, len, m, g, ncells, keys, idx, slab_ids);
}

void launch_build_records(const uint32_t *skeys, const uint32_t *sidx, size_t m, const uint16_t *ex,
                          const uint16_t *ey, const uint32_t *et, const int32_t *nextp, const double *len,
                          const double *lcx, const double *lcy, int monotone, uint4 *rec, double *pay,
                          uint32_t *cell_start, uint32_t ncells, cudaStream_t s) {
  if (m) k_build_records<<<nb(m, 256), 256, 0, s>>>(skeys, sidx, m, ex, ey, et, nextp, len, lcx, lcy, monotone, rec,
                                                   pay, cell_start, ncells);
}

int pool_tile_smem_bytes() { return (int)sizeof(TileSmem<16, 640>); }

// Launches the fast path (when `fast` is set) and then the general, exact path for whatever is left.
// work_counter: two zeroed words.  done: m zeroed bytes.
int launch_pooling(const uint4 *rec, const double *pay, const uint32_t *cell_start, const uint32_t *slab_ids,
                   uint8_t *done, size_t m, uint32_t ncells, int h, int nslabs, PoolGeom g, int fast,
                   double flow_per_slab, double *global_r, double *global_theta, uint8_t *scale,
                   unsigned int *work_counter, unsigned long long *cand_count, int num_sms, cudaStream_t s) {
  if (!m) return 0;
  (void)flow_per_slab;
  int launches = 0;
  PoolArgs A;
  A.rec = rec; A.pay = pay; A.cell_start = cell_start; A.slab_ids = slab_ids; A.done = done;
  A.m = m; A.ncells = ncells; A.h = h; A.nslabs = nslabs; A.g = g;
  A.global_r = global_r; A.global_theta = global_theta; A.scale = scale;
  A.cand_count = cand_count;
  if (fast && g.tile_shift == 4) {
    A.work_counter = work_counter;
    launch_tile<16, 640>(A, nslabs, num_sms, s);  // 16 warps, 1 CTA per SM, ~224 KB of shared memory
    launches++;
  }
  A.work_counter = work_counter + 1;
  // persistent warps pulling 32-slot groups from a global counter: grid = SMs x resident CTAs
  unsigned grid = (unsigned)num_sms * 5u;
  unsigned need = nb(m, 32 * PW);
  if (grid > need) grid = need;
  k_pool_any<<<grid, PW * 32, 0, s>>>(A);
  launches++;
  return launches;
}
