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
// The kernels that share that definition:
//   k_pool_tile  -- the fast path.  A CTA owns a 32x32-pixel tile for a run of time slabs and keeps the flow
//                   events of the (32+100)^2 region of the slabs its current round's windows span staged in
//                   shared memory; every half-warp pools one event of the tile against the staged set with FP32
//                   partial sums, and pools it again exactly (FP64) when the scale decision is not clear-cut.
//                   Needs sorted timestamps (so that the age test folds into an index bound) and windows that
//                   stay inside rows < 2H.  A flagged second pass with larger slots takes the rounds whose
//                   staging overflowed; k_pool_finish turns the sums it leaves into globalR / globalTheta.
//   k_pool_any   -- the general, exact path straight from the global index: unsorted timestamps, windows whose
//                   second coordinate runs past 2H (the reference's width-1 bound, below), whatever is left.
//   k_pool_bits  -- a measured alternative to k_pool_tile built on prefix bit tables (FARMS_POOL_IMPL=bits).
//
// Flat-index rule: the reference bounds the window's second coordinate by width-1
// (src/vFlow.cpp:1000, 1113) and indexes _data[i*H + j] unchecked (include/EventMatrix.h:32-34), so a
// logical cell (i, j >= H) aliases pixel (i + j/H, j mod H), and indices past W*H read as "no flow".
#include <algorithm>
#include <cstdlib>

#include "../../include/farms_b200.h"
#include "farms_dev.cuh"

namespace {

FARMS_CHK_DECL

// ------------------------------------------------------------------------------------------------
// index construction
// ------------------------------------------------------------------------------------------------

// Only events WITH flow (len > 0) can contribute (src/vFlow.cpp:1002) or be pooled (src/vFlow.cpp:315,
// 362), so the pooling index holds just those; the others get the sentinel key `ncells` and sort last.
__global__ void k_cell_keys(const uint16_t *__restrict__ ex, const uint16_t *__restrict__ ey,
                            const uint32_t *__restrict__ em, const uint32_t *__restrict__ excl,
                            const double *__restrict__ len, size_t m, PoolGeom g, uint32_t ncells,
                            uint32_t *__restrict__ keys, uint32_t *__restrict__ idx,
                            uint32_t *__restrict__ slab_ids, uint32_t *__restrict__ slab_first) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint32_t sid = em[j] >> g.slab_shift;
  const bool first = j == 0 || (em[j - 1] >> g.slab_shift) != sid;
  const uint32_t dense = excl[j] + ((j > 0 && first) ? 1u : 0u);
  if (first) {
    slab_ids[dense] = sid;
    slab_first[dense] = (uint32_t)j;
  }
  const uint32_t tile = (uint32_t)(ex[j] >> g.tile_shift) * (uint32_t)g.nty + (uint32_t)(ey[j] >> g.tile_shift);
  keys[j] = len[j] > 0.0 ? dense * (uint32_t)(g.ntx * g.nty) + tile : ncells;
  idx[j] = (uint32_t)j;
}

__global__ void k_fill_u32(uint32_t *__restrict__ p, uint32_t n, uint32_t v) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// one writer per table entry: the first event of every distinct timestamp fills the entries since the previous one
__global__ void k_time_table(const uint32_t *__restrict__ et, size_t m, uint32_t tt_base,
                             uint32_t *__restrict__ table, uint32_t tt_size) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint32_t t = et[j];
  uint32_t from;  // first table offset this event is the answer for
  if (j == 0) from = 0;
  else {
    const uint32_t tp = et[j - 1];
    if (tp == t) return;
    from = tp - tt_base + 1;
  }
  for (uint32_t u = from; u <= t - tt_base && u < tt_size; u++) table[u] = (uint32_t)j;
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
                                double *__restrict__ pay, uint32_t *__restrict__ cell_start, uint32_t ncells,
                                uint32_t h, unsigned int *__restrict__ n_targets,
                                const uint32_t *__restrict__ time_table, uint32_t tt_base, uint32_t tt_size) {
  size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool entry = pos < m && skeys[pos] < ncells;  // no flow: not part of the index
  {  // events the pooling kernels must produce: index entries that are not halo
    const unsigned bal = __ballot_sync(0xffffffffu, entry && sidx[pos] >= h);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_targets, (unsigned int)__popc(bal));
  }
  if (!entry) return;
  const uint32_t k = skeys[pos];
  const uint32_t j = sidx[pos];
  uint32_t end = (uint32_t)nextp[j];
  const uint32_t tj = et[j];
  if (monotone && time_table) {
    // first index u with et[u] >= tj + 500: one look-up in the per-microsecond table (k_time_table)
    const uint64_t off = (uint64_t)tj + FARMS_KILL_OLD_FLOW_TIME - tt_base;
    const uint32_t lo = off < tt_size ? time_table[off] : (uint32_t)m;
    if (lo < end) end = lo;
  } else if (monotone) {
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
  const uint32_t *slab_first;  // index of the first event of every dense slab
  const double *ev_len, *ev_lcx, *ev_lcy;  // FP64 flow values by event index (exact re-pooling of undecided events)
  uint32_t *fin;               // per output event: contributor count | scale index << 16 of a fast-path result
  uint32_t *item_ovf;          // per (owner tile, slab segment) item: bit (slab - first slab) / 2 = that pair of slabs
                               // had targets the first pass could not stage (slot overflow)
  const uint8_t *own_ok;  // serial semantics only (else null): per event, does its own pixel pass the age test
  uint8_t *done;       // per index position: 1 once the fast path has pooled that event
  size_t m;            // stride of the pay arrays
  uint32_t ncells;     // cell_start[ncells] = entries in the index (events with flow)
  int h;               // halo events (indices < h) are contributors only
  int nslabs;
  PoolGeom g;
  double *global_r, *global_theta;
  uint8_t *scale;
  unsigned int *work_counter;
  // per batch, zeroed by the host: [0] some round's staging overflowed (the second pass has work), [1] events of
  // this batch that must be pooled (index entries that are not halo), [2] of those, finished by the fast passes
  unsigned int *batch_words;
  unsigned long long *cand_count;
  unsigned long long *path_count;  // events pooled by: [0] first fast pass, [1] flagged second pass, [2] k_pool_any
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
  if (lane == 0 && write && FARMS_CHK(out_index >= 0 && (size_t)out_index < A.m - (size_t)A.h, 122)) {
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
  unsigned long long ncand = 0, npooled = 0;
  if (A.batch_words[1] == A.batch_words[2]) return;  // the fast passes finished every event of the batch

  constexpr unsigned int CHUNK = 32;  // index positions per grab: few atomics even when almost nothing is left to pool
  for (unsigned int base = 0, chunk_end = 0;; base += 32) {
    if (base >= chunk_end) {
      if (lane == 0) base = atomicAdd(A.work_counter, CHUNK);
      base = __shfl_sync(0xffffffffu, base, 0);
      chunk_end = base + CHUNK;
    }
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
                             A.g.slab_shift;
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
            (void)FARMS_CHK(cb + tyhi + 1 <= A.ncells && e >= s && e <= mi, 131);
            for (uint32_t p = s + lane; p < e; p += 32) {
              const uint4 c = A.rec[p];
              const int cx = (int)(c.x & 0xffffu), cy = (int)(c.x >> 16);
              const long long dt = (long long)ti - (long long)c.y;
              bool ok = (int)c.z <= ii && (int)c.w > ii && cx >= pxlo && cx <= pxhi && cy >= pylo &&
                        cy <= pyhi && dt < FARMS_KILL_OLD_FLOW_TIME && dt > -FARMS_KILL_OLD_FLOW_TIME;  // :1002
              // serial semantics: the event's own cell is tested with its pixel's previous time (src/vFlow.cpp:790)
              if (A.own_ok && (int)c.z == ii) ok = ok && A.own_ok[ii] != 0;
              if (ok) {
                const int dx = abs(cx - k - xi), dy = abs(cy + k * H - yi);
                const int ring = (max(dx, dy) + FARMS_WINDOW_JUMP - 1) / FARMS_WINDOW_JUMP;
                (void)FARMS_CHK(ring >= 0 && ring < FARMS_NSCALES, 132);
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
      npooled++;
    }
  }
  if (lane == 0 && ncand) atomicAdd(A.cand_count, ncand);
  if (lane == 0 && npooled) atomicAdd(A.path_count + 2, npooled);
}

// ------------------------------------------------------------------------------------------------
// fast path: owner tiles with shared-memory staging, FP32 ring partials, exact decisions
// ------------------------------------------------------------------------------------------------
// Ring partial sums are kept in FP32 (half the shared memory of FP64 => twice the resident warps) and
// combined in FP64.  That perturbs a scale's mean by < 1e-5 relative, so an event is only finished from them when
// its arg-max over scales is decided by a margin > 2e-5 and its mean vector is not a cancellation residue;
// everything else (measured: about 2 % of events) is pooled again in FP64 by the same half-warp.  Scales whose
// extra rings are empty hold the same contributors (equal counts) and never compete, so exact ties behave like
// the reference's strict '>'.
constexpr int OT_SHIFT = 5, OT = 1 << OT_SHIFT;  // owner tile edge (pixels)
// NSL (template parameter): consecutive slabs pooled per round (more events per round => fewer, fuller waves).
// A 500-us window reaches back TK_LB slabs at most, so a round's windows span <= TK_LB + NSL staged slabs (the slot
// ring).
constexpr int TK_LB = FARMS_SLAB_LOOKBACK;
constexpr int TK_OVF_SHIFT = 1;  // slabs per bit of the overflow bitmap, as a shift (32 bits per item of 64 slabs)
constexpr int TK_PAD = 64;     // the pooling loop reads 4 x 16 records at a time without bounds checks
#ifndef FARMS_TK_SEG
#define FARMS_TK_SEG 64
#endif
constexpr int TK_SEG = FARMS_TK_SEG;  // slabs per work item
constexpr int TK_MAXT = 128;   // targets handled per round and slab
constexpr int TK_MAXRUN = 20;  // tile-column runs of a region: <= 10 for rows < H plus <= 10 aliased
constexpr float TK_TIE_TOL = 2e-5f;

template <int WARPS, int CAP, int NSL>
struct TileSmem {
  // {x | y<<16 (logical window coordinates), idx, end - idx, |flow|cos as f32} and |flow|sin; |flow| itself is
  // recomputed from the two (4 bytes per staged record buy 23 % more records per slot)
  uint4 ra[(TK_LB + NSL)][CAP + TK_PAD];
  float rb[(TK_LB + NSL)][CAP];
  float4 acc[WARPS][FARMS_NSCALES][32];   // per-lane ring partials: len, lcx, lcy, count
  uint32_t tlist[NSL][TK_MAXT];
  // staging pass: run descriptors of all slabs being staged (start in the index, flat offset, slab | aliased << 7)
  uint32_t run_s[(TK_LB + NSL) * TK_MAXRUN], run_o[(TK_LB + NSL) * TK_MAXRUN + 1];
  uint8_t run_info[(TK_LB + NSL) * TK_MAXRUN];
  uint32_t slab_f[(TK_LB + NSL) + 1], slab_pre[(TK_LB + NSL) + 1];  // per staged slab: first flat position, passes before it
  uint32_t wcount[WARPS];
  int tag[(TK_LB + NSL)];
  int count[(TK_LB + NSL)];
  int overflow[(TK_LB + NSL)];
  int dlo[NSL], dhi[NSL], ovf[NSL];  // per target slab of the round: staged slabs its windows span
  unsigned int ntg[NSL], tnext, item;
};

struct Region;
[[maybe_unused]] __device__ __forceinline__ bool xi_in_region(uint32_t xy, const Region &R);

struct Region {  // pixels an owner tile can reach, as physical rectangles
  int rx0, rx1, ry0, ry1;  // rows < H (k = 0)
  int ax0, ax1, ay1;       // aliased part (k = 1): physical x in [ax0, ax1], y in [0, ay1]; empty if ay1 < 0
};

[[maybe_unused]] __device__ __forceinline__ bool xi_in_region(uint32_t xy, const Region &R) {  // (checked build) a target lies in its tile
  const int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
  return x >= R.rx0 && x <= R.rx1 && y >= R.ry0;
}

// Stage the flow events of the dense slabs s0 .. s1 inside region R into their ring slots (slab s -> slot
// s mod ring) in ONE pass over the concatenation of their index runs, preserving index order inside every slab
// (ordered compaction => deterministic summation order).  One pass for all slabs of a round matters for sparse
// streams, where a slab holds a few dozen records and the fixed cost of a pass (barriers, L2 latency) dominates.
// Aliased events are stored with their LOGICAL window coordinates (x - 1, y + H) so that the pooling loop needs
// no special case.
template <class SM, int WARPS, int CAP, int NSL>
__device__ void stage_slabs(const PoolArgs &A, SM &S, int s0, int s1, const Region &R, uint32_t i_round) {
  constexpr int THREADS = WARPS * 32, RING = TK_LB + NSL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ts = A.g.tile_shift, nty = A.g.nty, NT = A.g.ntx * A.g.nty, H = A.g.H;
  const int tx0 = R.rx0 >> ts, tx1 = R.rx1 >> ts, ty0 = R.ry0 >> ts, ty1 = R.ry1 >> ts;
  // On a sensor taller than wide the reference's row bound (width - 1) can lie above the whole region
  // (ry1 < ry0): no logical row of such an owner tile is reachable and nothing is staged.
  const int nrun0 = R.ry1 >= R.ry0 ? tx1 - tx0 + 1 : 0;
  const int atx0 = R.ax0 >> ts, atx1 = R.ax1 >> ts;
  const int nrun1 = R.ay1 >= 0 ? atx1 - atx0 + 1 : 0;
  const int nrun = nrun0 + nrun1;
  const int nsl = s1 - s0 + 1, nruns = nsl * nrun;
  __syncthreads();  // previous users of the run tables, of wcount and of these slots are done
  for (int q = tid; q < nruns; q += THREADS) {
    const int sl = q / nrun, c = q - sl * nrun;
    uint32_t a, b;
    if (c < nrun0) {
      const size_t cb = (size_t)(s0 + sl) * NT + (size_t)(tx0 + c) * nty;
      a = A.cell_start[cb + ty0];
      b = A.cell_start[cb + ty1 + 1];
    } else {
      const size_t cb = (size_t)(s0 + sl) * NT + (size_t)(atx0 + c - nrun0) * nty;
      a = A.cell_start[cb];
      b = A.cell_start[cb + (R.ay1 >> ts) + 1];
    }
    S.run_s[q] = a;
    S.run_o[q + 1] = b - a;  // lengths first, offsets below
    S.run_info[q] = (uint8_t)(sl | (c >= nrun0 ? 0x80 : 0));
  }
  __syncthreads();
  if (warp == 0) {  // exclusive offsets of the runs in the flat list: consecutive runs per lane + a warp scan
    constexpr int RPL = (RING * TK_MAXRUN + 31) / 32;
    uint32_t loc[RPL], sum = 0;
#pragma unroll
    for (int j = 0; j < RPL; j++) {
      const int q = lane * RPL + j;
      loc[j] = q < nruns ? S.run_o[q + 1] : 0u;
      sum += loc[j];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    uint32_t run = inc - sum;
#pragma unroll
    for (int j = 0; j < RPL; j++) {
      const int q = lane * RPL + j;
      run += loc[j];
      if (q < nruns) S.run_o[q + 1] = run;
    }
    if (lane == 0) S.run_o[0] = 0;
  }
  __syncthreads();
  const uint32_t total = S.run_o[nruns];
  if (tid <= nsl) {
    S.slab_f[tid] = tid < nsl ? S.run_o[tid * nrun] : total;
    S.slab_pre[tid] = 0;
  }
  const double *pay_cx = A.pay + A.m, *pay_cy = A.pay + 2 * A.m;
  uint32_t out_base = 0;
  // two records per thread and trip (flat positions r0 + tid and r0 + THREADS + tid): the index record and both
  // payload values of a record are requested together
  for (uint32_t r0 = 0; r0 < total; r0 += 2 * THREADS) {
    bool pass[2] = {false, false};
    uint4 rec[2];
    double cxv[2] = {0.0, 0.0}, cyv[2] = {0.0, 0.0};
    uint32_t info[2] = {0, 0};
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const uint32_t f = r0 + e * THREADS + tid;
      rec[e] = make_uint4(0, 0, 0, 0);
      if (f < total) {
        int lo = 0, hi = nruns - 1;  // last run whose offset is <= f
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (S.run_o[mid] <= f) lo = mid; else hi = mid - 1;
        }
        const uint32_t pos = S.run_s[lo] + (f - S.run_o[lo]);
        rec[e] = A.rec[pos];
        cxv[e] = pay_cx[pos];
        cyv[e] = pay_cy[pos];
        info[e] = S.run_info[lo];
        pass[e] = true;
      }
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
      int x = (int)(rec[e].x & 0xffffu), y = (int)(rec[e].x >> 16);
      if (!(info[e] & 0x80u)) {
        pass[e] = pass[e] && x >= R.rx0 && x <= R.rx1 && y >= R.ry0 && y <= R.ry1;
      } else {
        pass[e] = pass[e] && x >= R.ax0 && x <= R.ax1 && y <= R.ay1;
        x -= 1;
        y += H;
      }
      // superseded at its pixel (or past 500 us) before the first event of the round: dead for every target
      pass[e] = pass[e] && rec[e].w > i_round;
      rec[e].x = (uint32_t)x | ((uint32_t)y << 16);
    }
    const unsigned bal0 = __ballot_sync(0xffffffffu, pass[0]), bal1 = __ballot_sync(0xffffffffu, pass[1]);
    if (lane == 0) S.wcount[warp] = (uint32_t)__popc(bal0) | ((uint32_t)__popc(bal1) << 16);
    __syncthreads();
    uint32_t pre0 = 0, pre1 = 0, all0 = 0, all1 = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
      const uint32_t cw = S.wcount[w];
      if (w < warp) {
        pre0 += cw & 0xffffu;
        pre1 += cw >> 16;
      }
      all0 += cw & 0xffffu;
      all1 += cw >> 16;
    }
    const uint32_t lt = (1u << lane) - 1u;
    // passes before this record in the flat order
    const uint32_t g0 = out_base + pre0 + __popc(bal0 & lt), g1 = out_base + all0 + pre1 + __popc(bal1 & lt);
    // the thread that holds the first flat position of a slab publishes how many passes precede that slab
    // (several empty slabs can share one first position)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const uint32_t f = r0 + e * THREADS + tid;
      if (f < total)
        for (int sl = 0; sl < nsl; sl++)
          if (S.slab_f[sl] == f) S.slab_pre[sl] = e ? g1 : g0;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int sl = (int)(info[e] & 0x7fu);
      const uint32_t o = (e ? g1 : g0) - S.slab_pre[sl];
      if (pass[e] && o < (uint32_t)CAP) {
        const int slot = (s0 + sl) % RING;
        S.ra[slot][o] = make_uint4(rec[e].x, rec[e].z, rec[e].w - rec[e].z, __float_as_uint(__double2float_rn(cxv[e])));
        S.rb[slot][o] = __double2float_rn(cyv[e]);
      }
    }
    out_base += all0 + all1;
  }
  __syncthreads();
  // slabs without any raw record at or after their first position (trailing empties) start at the end
  if (tid <= nsl && S.slab_f[tid] >= total) S.slab_pre[tid] = out_base;
  __syncthreads();
  if (tid < nsl) {
    const uint32_t raw = S.slab_pre[tid + 1] - S.slab_pre[tid];
    const int slot = (s0 + tid) % RING;
    S.tag[slot] = s0 + tid;
    S.count[slot] = (int)min(raw, (uint32_t)CAP);
    S.overflow[slot] = raw > (uint32_t)CAP;
  }
  // entries the unrolled pooling loop may touch past the end of a slot: span 0 never passes
  for (int q = tid; q < nsl * TK_PAD; q += THREADS) {
    const int sl = q / TK_PAD;
    const uint32_t cnt = min(S.slab_pre[sl + 1] - S.slab_pre[sl], (uint32_t)CAP);
    S.ra[(s0 + sl) % RING][cnt + (q - sl * TK_PAD)] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Like finish_event, but from FP32 partial sums: writes only when the decision is safe; returns (in every lane
// of the segment) whether the event was finished.
// The nested-square sums are an inclusive scan over the ring lanes (4 shuffle steps instead of 11 broadcasts).
// A scan associates differently from the reference's running sum, so "an empty ring adds exactly nothing" is
// enforced through the counts instead: a scale whose count equals the previous scale's holds the same
// contributors, has the same mean in the reference's arithmetic and can never beat it under the strict '>'
// (src/vFlow.cpp:1054); only scales that add contributors compete.
__device__ __forceinline__ bool finish_event_checked(const PoolArgs &A, int sub, double rl, double rx, double ry,
                                                     int rn, int out_index, bool have) {
  double Sl = rl, Sx = rx, Sy = ry;
  int Sn = rn;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) {
    const double pl = __shfl_up_sync(0xffffffffu, Sl, o, 16), px = __shfl_up_sync(0xffffffffu, Sx, o, 16),
                 py = __shfl_up_sync(0xffffffffu, Sy, o, 16);
    const int pn = __shfl_up_sync(0xffffffffu, Sn, o, 16);
    if (sub >= o) {
      Sl += pl;
      Sx += px;
      Sy += py;
      Sn += pn;
    }
  }
  const int nprev = __shfl_up_sync(0xffffffffu, Sn, 1, 16);
  const bool scale_lane = sub < FARMS_NSCALES && Sn > 0;
  const float mean = scale_lane ? __fdiv_rn((float)Sl, (float)Sn) : 0.f;
  const bool competes = scale_lane && (sub == 0 || Sn != nprev);
  // first maximum over the competing lanes (strict '>' from 0, :1047-1059)
  float best = competes ? mean : 0.f;
  int bk = (competes && mean > 0.f) ? sub : 99;
#pragma unroll
  for (int o = 8; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o, 16);
    const int ok = __shfl_xor_sync(0xffffffffu, bk, o, 16);
    if (ob > best || (ob == best && ok < bk)) {
      best = ob;
      bk = ok;
    }
  }
  const bool found = bk < FARMS_NSCALES;
  const int srcl = found ? bk : 0;
  const int bn = __shfl_sync(0xffffffffu, Sn, srcl, 16);
  // is any other scale (with a different contributor set) within the FP32 noise of the winner?
  const bool rival = scale_lane && Sn != bn && fabsf(mean - best) <= TK_TIE_TOL * best;
  const unsigned seg = 0xffffu << (threadIdx.x & 16);
  const bool any_rival = (__ballot_sync(0xffffffffu, rival) & seg) != 0u;
  const double wx = __shfl_sync(0xffffffffu, Sx, srcl, 16), wy = __shfl_sync(0xffffffffu, Sy, srcl, 16);
  bool safe = have && found && !any_rival && best > 1e-30f && best < 1e30f;
  // mean vector much shorter than the mean length: the FP32 sums cancelled, let the exact path do it
  const double bl = (double)best * (double)bn;
  safe = safe && (wx * wx + wy * wy) > 1e-4 * bl * bl;
  if (sub == 0 && safe && FARMS_CHK(out_index >= 0 && (size_t)out_index < A.m - (size_t)A.h, 121)) {
    // k_pool_finish divides by the count and takes sqrt / atan2 (src/vFlow.cpp:365-366)
    A.global_r[out_index] = wx;
    A.global_theta[out_index] = wy;
    A.fin[out_index] = (uint32_t)bn | ((uint32_t)bk << 16);
  }
  return safe;
}

// SECOND: a later pass over the same items with larger slots; only rounds that still hold undone targets (their
// staging overflowed the slots of the first pass: locally dense scenes) do any work.
template <int WARPS, int CAP, int NSL, int CTAS, bool SECOND>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k_pool_tile(PoolArgs A, int otx_n, int oty_n, int nseg) {
  using SM = TileSmem<WARPS, CAP, NSL>;
  constexpr int THREADS = WARPS * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4, sub = lane & 15;
  const int W = A.g.W, H = A.g.H, nty = A.g.nty, NT = A.g.ntx * A.g.nty;
  const unsigned int nitems = (unsigned int)otx_n * oty_n * nseg;
  unsigned long long ncand = 0;
  unsigned int npooled = 0;

  if (SECOND && A.batch_words[0] == 0u) return;  // no round overflowed in the first pass: nothing to do

  for (;;) {
    __syncthreads();
    if (tid == 0) S.item = atomicAdd(A.work_counter, 1u);
    if (tid < (TK_LB + NSL)) S.tag[tid] = -1;
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

    // NSL consecutive slabs per round
    for (int d = d_begin; d < d_end; d += NSL) {
      const int nd = min(NSL, d_end - d);
      // ---- targets of this round: flow events of the owner tile in slabs d .. d+nd-1 ----
      uint32_t ta[NSL][2], tb[NSL][2], nraw[NSL];
      uint32_t nraw_all = 0, nmax = 0;
#pragma unroll
      for (int w = 0; w < NSL; w++) {
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
      if (SECOND) {  // only rounds the first pass flagged do any work here
        const uint32_t bits = A.item_ovf[item];
        const int b0 = (d - d_begin) >> TK_OVF_SHIFT, b1 = (d + nd - 1 - d_begin) >> TK_OVF_SHIFT;
        if (((bits >> b0) & ((2u << (b1 - b0)) - 1u)) == 0u) continue;
      }

      // ---- make sure the slabs of all windows of the round are staged ----
      __syncthreads();  // the previous round is done with S.dlo / S.dhi / S.ovf
      if (tid < NSL) {
        const int dd = min(d + tid, d_end - 1);
        const uint32_t t_first = A.slab_ids[dd] << A.g.slab_shift;
        const uint32_t lo_id =
            (t_first >= (uint32_t)(FARMS_KILL_OLD_FLOW_TIME - 1) ? t_first - (FARMS_KILL_OLD_FLOW_TIME - 1) : 0u) >> A.g.slab_shift;
        int l = dd;
        while (l > 0 && dd - l < TK_LB && A.slab_ids[l - 1] >= lo_id) l--;
        S.dlo[tid] = l;
        S.dhi[tid] = dd;
      }
      const uint32_t i_round = A.slab_first[d];
      __syncthreads();
      int s_first = 0x7fffffff, s_last = -1;
#pragma unroll
      for (int w = 0; w < NSL; w++) {
        if (nraw[w]) {
          s_first = min(s_first, S.dlo[w]);
          s_last = max(s_last, S.dhi[w]);
        }
      }
      // staged slabs are a contiguous range ending at the last staged slab, so what is missing is a suffix
      int s_new = s_first;
      while (s_new <= s_last && S.tag[s_new % (TK_LB + NSL)] == s_new) s_new++;  // uniform: tags are read after a barrier
      if (s_new <= s_last) stage_slabs<SM, WARPS, CAP, NSL>(A, S, s_new, s_last, R, i_round);
      __syncthreads();
      if (tid < NSL) {
        int o = 0;
        for (int s = S.dlo[tid]; s <= S.dhi[tid]; s++) o |= S.overflow[s % (TK_LB + NSL)];
        S.ovf[tid] = o;
        if (!SECOND && o && nraw[tid]) {
          atomicOr(&A.item_ovf[item], 1u << ((d + tid - d_begin) >> TK_OVF_SHIFT));
          A.batch_words[0] = 1u;
        }
      }

      for (uint32_t t0 = 0; t0 < nmax; t0 += TK_MAXT) {
        __syncthreads();
        if (tid < NSL) S.ntg[tid] = 0;
        if (tid == 0) S.tnext = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NSL; w++)
          for (uint32_t f = t0 + tid; f < min(nraw[w], t0 + TK_MAXT); f += THREADS) {
            const uint32_t n0 = tb[w][0] - ta[w][0];
            const uint32_t pos = f < n0 ? ta[w][0] + f : ta[w][1] + (f - n0);
            const uint4 r = A.rec[pos];
            const int yi = (int)(r.x >> 16);
            // fast-path conditions: not a halo event, window rows stay below 2H, staging complete
            const bool ok = (int)r.z >= A.h && min(yi + FARMS_MAX_WINDOW, W - 1) <= 2 * H - 1 && !S.ovf[w] &&
                            (!SECOND || !A.done[pos]);
            if (ok) S.tlist[w][atomicAdd(&S.ntg[w], 1u)] = pos;

          }
        __syncthreads();
        // Tasks: first pairs of targets of the same slab (lanes 0-15 pool one event, lanes 16-31 the next; both
        // halves share their loop bounds), then "solo" targets pooled by the two halves together, each half taking
        // every other 64-record trip (half the time of a pair task).  The odd target of a slab is a solo, and so
        // are the pairs that would not divide evenly among the warps when they are at most half a wave: the
        // round then ends half a task later instead of a whole one.
        // (Rounds of many slabs -- the sparse-stream variant -- keep the plain pairing: there an odd target shares
        // its warp with an idle half.)
        constexpr bool SOLOS = NSL <= 2;
        uint32_t pstart[NSL + 1], sstart[NSL + 1], npair[NSL];
        uint32_t P = 0;
#pragma unroll
        for (int w = 0; w < NSL; w++) {
          npair[w] = SOLOS ? S.ntg[w] >> 1 : (S.ntg[w] + 1) >> 1;
          P += npair[w];
        }
        uint32_t extra = SOLOS ? P % WARPS : 0u;
        if (extra > WARPS / 2) extra = 0;
#pragma unroll
        for (int w = NSL - 1; w >= 0; w--) {
          const uint32_t take = min(extra, npair[w]);
          npair[w] -= take;
          extra -= take;
        }
        pstart[0] = sstart[0] = 0;
#pragma unroll
        for (int w = 0; w < NSL; w++) {
          pstart[w + 1] = pstart[w] + npair[w];
          sstart[w + 1] = sstart[w] + (SOLOS ? S.ntg[w] - 2 * npair[w] : 0u);
        }
        const uint32_t tasks_p = pstart[NSL], tasks = tasks_p + sstart[NSL];

        for (;;) {
          uint32_t k = 0;
          if (lane == 0) k = atomicAdd(&S.tnext, 1u);
          k = __shfl_sync(0xffffffffu, k, 0);
          if (k >= tasks) break;
          const bool solo = k >= tasks_p;
          int w = 0;
          uint32_t kk;
          if (!solo) {
#pragma unroll
            for (int q = 1; q < NSL; q++) w += (k >= pstart[q]) ? 1 : 0;
            kk = (k - pstart[w]) * 2 + half;
          } else {
            const uint32_t j = k - tasks_p;
#pragma unroll
            for (int q = 1; q < NSL; q++) w += (j >= sstart[q]) ? 1 : 0;
            kk = 2 * (pstart[w + 1] - pstart[w]) + (j - sstart[w]);
          }
          const uint32_t nt = S.ntg[w];
          const bool have = solo ? half == 0 : kk < nt;  // this half owns a target's result
          const uint32_t tpos = S.tlist[w][kk < nt ? kk : nt - 1];
          const uint4 r = A.rec[tpos];
          const int xi = (int)(r.x & 0xffffu), yi = (int)(r.x >> 16);
          const uint32_t ii = r.z;
          const int jlo = max(0, yi - FARMS_MAX_WINDOW), jhi = min(yi + FARMS_MAX_WINDOW, W - 1);  // :1000 (sic)
          const bool rows_ok = jhi >= jlo;  // jhi < jlo when W < H: no cell qualifies
          const uint32_t jspan = rows_ok ? (uint32_t)(jhi - jlo) : 0u;
          const int ylo = rows_ok ? jlo : 0x7fff0000;  // makes (cy - ylo) huge => fails
          const int xoff = FARMS_MAX_WINDOW - xi;
#pragma unroll
          for (int q = 0; q < FARMS_NSCALES; q++) S.acc[warp][q][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int sl = S.dlo[w], sh = S.dhi[w];
          for (int s = sl; s <= sh; s++) {
            const int slot = s % (TK_LB + NSL);
            const int n = S.count[slot];
            ncand += (sub == 0 && have) ? n : 0;
            for (int q0 = sub + (solo ? 64 * half : 0); q0 < n; q0 += solo ? 128 : 64) {
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
                  const float fx = __uint_as_float(c[u].w), fy = S.rb[slot][q0 + 16 * u];
                  float fl;
                  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(fl) : "f"(fx * fx + fy * fy));
                  float4 v = S.acc[warp][ring][lane];
                  v.x += fl;
                  v.y += fx;
                  v.z += fy;
                  v.w += 1.f;
                  S.acc[warp][ring][lane] = v;
                }
              }
            }
          }
          __syncwarp();
          // sub-lane k < 11 of each half combines ring k's 16 per-lane partials: four at a time in FP32, the four
          // group sums in FP64 (the counts are exact either way)
          double rl = 0.0, rx = 0.0, ry = 0.0;
          float rnf = 0.f;
          if (sub < FARMS_NSCALES) {
#pragma unroll
            for (int q4 = 0; q4 < 4; q4++) {
              float4 g = S.acc[warp][sub][(half << 4) | ((4 * q4 + sub) & 15)];
#pragma unroll
              for (int q = 1; q < 4; q++) {
                const float4 v = S.acc[warp][sub][(half << 4) | ((4 * q4 + q + sub) & 15)];
                g.x += v.x;
                g.y += v.y;
                g.z += v.z;
                g.w += v.w;
              }
              rl += (double)g.x;
              rx += (double)g.y;
              ry += (double)g.z;
              rnf += g.w;
            }
          }
          __syncwarp();
          if (solo) {  // the two halves pooled disjoint trips of the same target
            rl += __shfl_xor_sync(0xffffffffu, rl, 16);
            rx += __shfl_xor_sync(0xffffffffu, rx, 16);
            ry += __shfl_xor_sync(0xffffffffu, ry, 16);
            rnf += __shfl_xor_sync(0xffffffffu, rnf, 16);
          }
          const bool fin = finish_event_checked(A, sub, rl, rx, ry, (int)rnf, (int)ii - A.h, have);

          if (sub == 0 && fin) {
            A.done[tpos] = 1;
            npooled++;
          }
          // ---- undecided targets (a rival scale within the FP32 noise, cancelling vectors): pool them again
          // exactly, FP64 partials and the FP64 flow values of the contributors, one half-warp at a time
          // (the FP64 partials of one target take the warp's whole accumulator space) ----
          const unsigned need = __ballot_sync(0xffffffffu, have && !fin && sub == 0);
          for (int hsel = 0; hsel < 2; hsel++) {
            if (!((need >> (16 * hsel)) & 1u)) continue;  // uniform across the warp
            __syncwarp();
            double4 *dacc = reinterpret_cast<double4 *>(&S.acc[warp][0][0]);  // [ring][16 lanes] {len, lcx, lcy, n}
            if (half == hsel) {
#pragma unroll
              for (int q = 0; q < FARMS_NSCALES; q++) dacc[q * 16 + sub] = make_double4(0.0, 0.0, 0.0, 0.0);
              for (int s = sl; s <= sh; s++) {
                const int slot = s % (TK_LB + NSL);
                const int n = S.count[slot];
                for (int q0 = sub; q0 < n; q0 += 16) {
                  const uint4 c = S.ra[slot][q0];
                  const int cx = (int)(c.x & 0xffffu), cy = (int)(c.x >> 16);
                  const bool ok = (ii - c.y) < c.z && (uint32_t)(cx + xoff) <= 2u * FARMS_MAX_WINDOW &&
                                  (uint32_t)(cy - ylo) <= jspan;
                  if (ok) {
                    const int mch = max(abs(cx - xi), abs(cy - yi));
                    const int ring = ((mch + FARMS_WINDOW_JUMP - 1) * 205) >> 10;
                    double4 v = dacc[ring * 16 + sub];
                    v.x += A.ev_len[c.y];
                    v.y += A.ev_lcx[c.y];
                    v.z += A.ev_lcy[c.y];
                    v.w += 1.0;
                    dacc[ring * 16 + sub] = v;
                  }
                }
              }
            }
            __syncwarp();
            double el = 0.0, ex2 = 0.0, ey2 = 0.0, en = 0.0;
            if (half == hsel && sub < FARMS_NSCALES) {
#pragma unroll 4
              for (int q = 0; q < 16; q++) {
                const double4 v = dacc[sub * 16 + ((q + sub) & 15)];
                el += v.x;
                ex2 += v.y;
                ey2 += v.z;
                en += v.w;
              }
            }
            __syncwarp();
            finish_event<16>(A, sub, el, ex2, ey2, en, A.ev_lcx[ii], A.ev_lcy[ii], (int)ii - A.h, half == hsel);
            if (half == hsel && sub == 0) {
              A.done[tpos] = 1;
              npooled++;
            }
          }
        }
      }
    }
  }
  if ((lane & 15) == 0 && ncand) atomicAdd(A.cand_count, ncand);
  if ((lane & 15) == 0 && npooled) {
    atomicAdd(A.path_count + (SECOND ? 1 : 0), (unsigned long long)npooled);
    atomicAdd(A.batch_words + 2, npooled);
  }
}

template <int WARPS, int CAP, int NSL, int CTAS, bool SECOND>
void launch_tile(const PoolArgs &A0, int nslabs, int num_sms, cudaStream_t s) {
  PoolArgs A = A0;
  using SM = TileSmem<WARPS, CAP, NSL>;
  static_assert(sizeof(SM) <= (CTAS == 2 ? 115712 : 232448), "shared memory of the tile kernel: 227 KB per CTA, 228 KB per SM");
  auto kern = k_pool_tile<WARPS, CAP, NSL, CTAS, SECOND>;
  // per device/context, so set on every launch (a process-wide flag would leave every device but the first
  // without the opt-in to > 48 KB of dynamic shared memory)
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  const int otx = (A.g.W + OT - 1) >> OT_SHIFT, oty = (A.g.H + OT - 1) >> OT_SHIFT;
  const int nseg = (nslabs + TK_SEG - 1) / TK_SEG;
  const long long items = (long long)otx * oty * nseg;
  unsigned grid = (unsigned)std::min<long long>(items, (long long)num_sms * CTAS);
  kern<<<grid, WARPS * 32, sizeof(SM), s>>>(A, otx, oty, nseg);
}


// ------------------------------------------------------------------------------------------------
// fast path, two-phase: k_pool_warp
// ------------------------------------------------------------------------------------------------
// Same owner-tile / slab-ring structure as k_pool_tile, with a different inner loop.  In k_pool_tile a lane tests
// a staged record and, if it passes, immediately adds it to its ring accumulator in shared memory: 38 % of the
// lanes pass, so the 22-instruction accumulate body runs at 38 % lane efficiency and dominates the kernel.  Here a
// whole warp pools one event in two phases:
//   A  every lane tests records (8-byte packed test record: one LDS.64, one VABSDIFF4 for the window, a 24-bit
//      index/span compare for "still the latest event of its pixel and younger than 500 us") and appends the slot
//      address of each passing record to its own lane-private queue (one predicated STS.U16, no ballots);
//   B  every lane drains its queue: payload (|flow|cos, |flow|sin as float2), ring from the packed coordinates,
//      read-modify-write of its float4 ring accumulator -- all lanes busy (a lane holds 17 +- 4 entries).
// Staged records are 16 bytes instead of 20 (coordinates relative to the tile's region in one byte each, the index
// relative to the slab's first event in 24 bits, the life span in 24 bits; a slab that does not fit 24 bits is
// flagged as overflowed and left to the exact path), so the queues cost no slot capacity.
#ifndef FARMS_SOLOS_MAX_NSL
#define FARMS_SOLOS_MAX_NSL 3
#endif
constexpr int WQ_DEPTH = 32;   // queue entries per lane before a drain is forced
constexpr int WP_PAD = 32;     // zeroed test records behind a slot's last record (span 0 never passes)

template <int WARPS, int CAP, int NSL>
struct WarpSmem {
  static constexpr int RING = TK_LB + NSL, PAD = WP_PAD, STRIDE = CAP + PAD;
  uint2 ta[RING * STRIDE];   // {x_rel | y_rel << 8 | idx_rel[15:0] << 16,  idx_rel[23:16] | span << 8}
  float2 pb[RING * STRIDE];  // |flow|cos(theta), |flow|sin(theta)
  float4 acc[WARPS][FARMS_NSCALES][32];  // per-lane ring partials: len, lcx, lcy, count
  uint16_t queue[WARPS][WQ_DEPTH][32];
  uint32_t tlist[NSL][TK_MAXT];
  uint32_t run_s[RING * TK_MAXRUN], run_o[RING * TK_MAXRUN + 1];
  uint8_t run_info[RING * TK_MAXRUN];
  uint32_t slab_f[RING + 1], slab_pre[RING + 1];
  uint32_t wcount[WARPS];
  uint32_t slot_base[RING];  // index of the first event of the slab a slot holds (idx_rel is relative to it)
  int tag[RING], count[RING], overflow[RING];
  int dlo[NSL], dhi[NSL], ovf[NSL];
  unsigned int ntg[NSL], tnext, item;
};

// stage_slabs for the packed layout (same one-pass ordered compaction over the concatenated index runs)
// XCULL: also leave, per slot, where the staged records of each tile-column run begin (S.col_off: a target then
// visits only the columns its window touches).
template <class SM, int WARPS, int CAP, int NSL, bool XCULL = false>
__device__ void stage_slabs_packed(const PoolArgs &A, SM &S, int s0, int s1, const Region &R, uint32_t i_round) {
  constexpr int THREADS = WARPS * 32, RING = TK_LB + NSL, STRIDE = SM::STRIDE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ts = A.g.tile_shift, nty = A.g.nty, NT = A.g.ntx * A.g.nty, H = A.g.H;
  const int tx0 = R.rx0 >> ts, tx1 = R.rx1 >> ts, ty0 = R.ry0 >> ts, ty1 = R.ry1 >> ts;
  const int nrun0 = R.ry1 >= R.ry0 ? tx1 - tx0 + 1 : 0;  // empty on tall sensors whose row bound is above the region
  const int atx0 = R.ax0 >> ts, atx1 = R.ax1 >> ts;
  const int nrun1 = R.ay1 >= 0 ? atx1 - atx0 + 1 : 0;
  const int nrun = nrun0 + nrun1;
  const int nsl = s1 - s0 + 1, nruns = nsl * nrun;
  __syncthreads();  // previous users of the run tables, of wcount and of these slots are done
  (void)FARMS_CHK(nsl >= 1 && nsl <= RING && nrun <= TK_MAXRUN && s0 >= 0 && s1 < A.nslabs, 103);
  for (int q = tid; q < nruns; q += THREADS) {
    const int sl = q / nrun, c = q - sl * nrun;
    uint32_t a, b;
    if (c < nrun0) {
      const size_t cb = (size_t)(s0 + sl) * NT + (size_t)(tx0 + c) * nty;
      (void)FARMS_CHK(cb + ty1 + 1 <= A.ncells && ty0 <= ty1 + 1, 104);
      a = A.cell_start[cb + ty0];
      b = A.cell_start[cb + ty1 + 1];
    } else {
      const size_t cb = (size_t)(s0 + sl) * NT + (size_t)(atx0 + c - nrun0) * nty;
      (void)FARMS_CHK(cb + (R.ay1 >> ts) + 1 <= A.ncells, 105);
      a = A.cell_start[cb];
      b = A.cell_start[cb + (R.ay1 >> ts) + 1];
    }
    (void)FARMS_CHK(b >= a, 106);
    S.run_s[q] = a;
    S.run_o[q + 1] = b - a;  // lengths first, offsets below
    S.run_info[q] = (uint8_t)(sl | (c >= nrun0 ? 0x80 : 0));
  }
  if (tid < nsl) {
    const int slot = (s0 + tid) % RING;
    S.slot_base[slot] = A.slab_first[s0 + tid];
    S.overflow[slot] = 0;
  }
  __syncthreads();
  if (warp == 0) {  // exclusive offsets of the runs in the flat list: consecutive runs per lane + a warp scan
    constexpr int RPL = (RING * TK_MAXRUN + 31) / 32;
    uint32_t loc[RPL], sum = 0;
#pragma unroll
    for (int j = 0; j < RPL; j++) {
      const int q = lane * RPL + j;
      loc[j] = q < nruns ? S.run_o[q + 1] : 0u;
      sum += loc[j];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    uint32_t run = inc - sum;
#pragma unroll
    for (int j = 0; j < RPL; j++) {
      const int q = lane * RPL + j;
      run += loc[j];
      if (q < nruns) S.run_o[q + 1] = run;
    }
    if (lane == 0) S.run_o[0] = 0;
  }
  __syncthreads();
  const uint32_t total = S.run_o[nruns];
  if (tid <= nsl) {
    S.slab_f[tid] = tid < nsl ? S.run_o[tid * nrun] : total;
    S.slab_pre[tid] = 0;
  }
  const double *pay_cx = A.pay + A.m, *pay_cy = A.pay + 2 * A.m;
  uint32_t out_base = 0;
  int lo_hint = 0;  // this thread's flat positions only grow: the run search resumes where it stopped
  for (uint32_t r0 = 0; r0 < total; r0 += 2 * THREADS) {
    bool pass[2] = {false, false};
    uint4 rec[2];
    double cxv[2] = {0.0, 0.0}, cyv[2] = {0.0, 0.0};
    uint32_t info[2] = {0, 0};
    int first_of[2] = {-1, -1};  // XCULL: the run whose first raw record this thread holds
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const uint32_t f = r0 + e * THREADS + tid;
      rec[e] = make_uint4(0, 0, 0, 0);
      if (f < total) {
        int lo = lo_hint, hi = nruns - 1;  // last run whose offset is <= f
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (S.run_o[mid] <= f) lo = mid; else hi = mid - 1;
        }
        lo_hint = lo;
        if (XCULL && S.run_o[lo] == f) first_of[e] = lo;
        const uint32_t pos = S.run_s[lo] + (f - S.run_o[lo]);
        if (FARMS_CHK(lo >= 0 && lo < nruns && pos < A.cell_start[A.ncells], 101)) {
          rec[e] = A.rec[pos];
          cxv[e] = pay_cx[pos];
          cyv[e] = pay_cy[pos];
          info[e] = S.run_info[lo];
          pass[e] = true;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
      int x = (int)(rec[e].x & 0xffffu), y = (int)(rec[e].x >> 16);
      if (!(info[e] & 0x80u)) {
        pass[e] = pass[e] && x >= R.rx0 && x <= R.rx1 && y >= R.ry0 && y <= R.ry1;
      } else {
        pass[e] = pass[e] && x >= R.ax0 && x <= R.ax1 && y <= R.ay1;
        x -= 1;   // logical window coordinates of the aliased cell
        y += H;
      }
      // superseded at its pixel (or past 500 us) before the first event of the round: dead for every target
      pass[e] = pass[e] && rec[e].w > i_round;
      rec[e].x = (uint32_t)(x - R.rx0) | ((uint32_t)(y - R.ry0) << 8);  // both < 132 for a passing record
    }
    const unsigned bal0 = __ballot_sync(0xffffffffu, pass[0]), bal1 = __ballot_sync(0xffffffffu, pass[1]);
    if (lane == 0) S.wcount[warp] = (uint32_t)__popc(bal0) | ((uint32_t)__popc(bal1) << 16);
    __syncthreads();
    uint32_t pre0 = 0, pre1 = 0, all0 = 0, all1 = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
      const uint32_t cw = S.wcount[w];
      if (w < warp) {
        pre0 += cw & 0xffffu;
        pre1 += cw >> 16;
      }
      all0 += cw & 0xffffu;
      all1 += cw >> 16;
    }
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t g0 = out_base + pre0 + __popc(bal0 & lt), g1 = out_base + all0 + pre1 + __popc(bal1 & lt);
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const uint32_t f = r0 + e * THREADS + tid;
      if (f < total)
        for (int sl = 0; sl < nsl; sl++)
          if (S.slab_f[sl] == f) S.slab_pre[sl] = e ? g1 : g0;
    }
    __syncthreads();
    if constexpr (XCULL) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        // the compacted position of a run's first raw record is where the run's staged records begin; empty runs
        // just before it (possibly the tail of the previous slab) begin -- and end -- at the same place
        for (int q = first_of[e]; q >= 0 && (q == first_of[e] || S.run_o[q] == S.run_o[first_of[e]]); q--) {
          const int sl = q / nrun, c = q - sl * nrun;
          S.col_off[(s0 + sl) % RING][c] = (uint16_t)min((e ? g1 : g0) - S.slab_pre[sl], (uint32_t)CAP);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int sl = (int)(info[e] & 0x7fu);
      const uint32_t o = (e ? g1 : g0) - S.slab_pre[sl];
      if (pass[e] && o < (uint32_t)CAP) {
        const int slot = (s0 + sl) % RING;
        const uint32_t rel = rec[e].z - S.slot_base[slot], span = rec[e].w - rec[e].z;
        if ((rel | span) >> 24) S.overflow[slot] = 1;  // does not fit the packed record: leave the slab to the exact path
        if (FARMS_CHK(slot >= 0 && slot < RING && (rec[e].x & 0xffu) < 132u && ((rec[e].x >> 8) & 0xffu) < 132u &&
                          (rec[e].x >> 16) == 0u && rec[e].z >= S.slot_base[slot], 102)) {
          S.ta[slot * STRIDE + o] = make_uint2(rec[e].x | (rel << 16), ((rel >> 16) & 0xffu) | (span << 8));
          S.pb[slot * STRIDE + o] = make_float2(__double2float_rn(cxv[e]), __double2float_rn(cyv[e]));
        }
      }
    }
    out_base += all0 + all1;
  }
  __syncthreads();
  // slabs without any raw record at or after their first position (trailing empties) start at the end
  if (tid <= nsl && S.slab_f[tid] >= total) S.slab_pre[tid] = out_base;
  __syncthreads();
  if (tid < nsl) {
    const uint32_t raw = S.slab_pre[tid + 1] - S.slab_pre[tid];
    const int slot = (s0 + tid) % RING;
    S.tag[slot] = s0 + tid;
    S.count[slot] = (int)min(raw, (uint32_t)CAP);
    if (raw > (uint32_t)CAP) S.overflow[slot] = 1;
    if constexpr (XCULL) S.col_off[slot][nrun] = (uint16_t)min(raw, (uint32_t)CAP);
  }
  if constexpr (XCULL) {  // runs with no raw record at or after their start: they begin at the end of everything
    for (int q = tid; q < nruns; q += THREADS)
      if (S.run_o[q] >= total) {
        const int sl = q / nrun, c = q - sl * nrun;
        S.col_off[(s0 + sl) % RING][c] = (uint16_t)min(out_base - S.slab_pre[sl], (uint32_t)CAP);
      }
  }
  // test records the unrolled pooling loop may touch past the end of a slot: span 0 never passes
  for (int q = tid; q < nsl * SM::PAD; q += THREADS) {
    const int sl = q / SM::PAD;
    const uint32_t cnt = min(S.slab_pre[sl + 1] - S.slab_pre[sl], (uint32_t)CAP);
    S.ta[((s0 + sl) % RING) * STRIDE + cnt + (q - sl * SM::PAD)] = make_uint2(0u, 0u);
  }
}

// Phase B of k_pool_warp: every lane adds the records it queued to its ring accumulators.
template <class SM>
__device__ __forceinline__ void drain_queue(SM &S, int warp, int lane, int cnt, uint32_t tw) {
  const int mx = __reduce_max_sync(0xffffffffu, cnt);
  for (int e = 0; e < mx; e++) {
    if (e < cnt) {
      const uint32_t a = S.queue[warp][e][lane];
      const uint32_t d = __vabsdiffu4(S.ta[a].x, tw);  // |dx| in byte 0, |dy| in byte 1
      const float2 p = S.pb[a];
      const uint32_t mch = max(d & 0xffu, (d >> 8) & 0xffu);
      const uint32_t ring = ((mch + FARMS_WINDOW_JUMP - 1) * 205u) >> 10;  // /5 for values <= 54
      float fl;
      asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(fl) : "f"(p.x * p.x + p.y * p.y));
      float4 v = S.acc[warp][ring][lane];
      v.x += fl;
      v.y += p.x;
      v.z += p.y;
      v.w += 1.f;
      S.acc[warp][ring][lane] = v;
    }
  }
}

template <int WARPS, int CAP, int NSL, int CTAS, bool SECOND>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k_pool_warp(PoolArgs A, int otx_n, int oty_n, int nseg) {
  using SM = WarpSmem<WARPS, CAP, NSL>;
  constexpr int THREADS = WARPS * 32, RING = SM::RING, STRIDE = SM::STRIDE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 15;
  const int W = A.g.W, H = A.g.H, nty = A.g.nty, NT = A.g.ntx * A.g.nty;
  const unsigned int nitems = (unsigned int)otx_n * oty_n * nseg;
  unsigned long long ncand = 0;
  unsigned int npooled = 0;

  if (SECOND && A.batch_words[0] == 0u) return;  // no round overflowed in the first pass: nothing to do

  for (;;) {
    __syncthreads();
    if (tid == 0) S.item = atomicAdd(A.work_counter, 1u);
    if (tid < RING) S.tag[tid] = -1;
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

    for (int d = d_begin; d < d_end; d += NSL) {
      const int nd = min(NSL, d_end - d);
      // ---- targets of this round: flow events of the owner tile in slabs d .. d+nd-1 ----
      uint32_t ta_[NSL][2], tb_[NSL][2], nraw[NSL];
      uint32_t nraw_all = 0, nmax = 0;
#pragma unroll
      for (int w = 0; w < NSL; w++) {
        nraw[w] = 0;
#pragma unroll
        for (int c = 0; c < 2; c++) {
          ta_[w][c] = tb_[w][c] = 0;
          if (w < nd && itx0 + c <= itx1) {
            const size_t cb = (size_t)(d + w) * NT + (size_t)(itx0 + c) * nty;
            ta_[w][c] = A.cell_start[cb + ity0];
            tb_[w][c] = A.cell_start[cb + ity1 + 1];
          }
          nraw[w] += tb_[w][c] - ta_[w][c];
        }
        nraw_all += nraw[w];
        nmax = max(nmax, nraw[w]);
      }
      if (nraw_all == 0) continue;  // uniform across the CTA
      if (SECOND) {  // only rounds the first pass flagged do any work here
        const uint32_t bits = A.item_ovf[item];
        const int b0 = (d - d_begin) >> TK_OVF_SHIFT, b1 = (d + nd - 1 - d_begin) >> TK_OVF_SHIFT;
        if (((bits >> b0) & ((2u << (b1 - b0)) - 1u)) == 0u) continue;
      }

      // ---- make sure the slabs of all windows of the round are staged ----
      __syncthreads();  // the previous round is done with S.dlo / S.dhi / S.ovf
      if (tid < NSL) {
        const int dd = min(d + tid, d_end - 1);
        const uint32_t t_first = A.slab_ids[dd] << A.g.slab_shift;
        const uint32_t lo_id =
            (t_first >= (uint32_t)(FARMS_KILL_OLD_FLOW_TIME - 1) ? t_first - (FARMS_KILL_OLD_FLOW_TIME - 1) : 0u) >> A.g.slab_shift;
        int l = dd;
        while (l > 0 && dd - l < TK_LB && A.slab_ids[l - 1] >= lo_id) l--;
        S.dlo[tid] = l;
        S.dhi[tid] = dd;
      }
      const uint32_t i_round = A.slab_first[d];
      __syncthreads();
      int s_first = 0x7fffffff, s_last = -1;
#pragma unroll
      for (int w = 0; w < NSL; w++) {
        if (nraw[w]) {
          s_first = min(s_first, S.dlo[w]);
          s_last = max(s_last, S.dhi[w]);
        }
      }
      // staged slabs are a contiguous range ending at the last staged slab, so what is missing is a suffix
      int s_new = s_first;
      while (s_new <= s_last && S.tag[s_new % RING] == s_new) s_new++;  // uniform: tags are read after a barrier
      if (s_new <= s_last) stage_slabs_packed<SM, WARPS, CAP, NSL>(A, S, s_new, s_last, R, i_round);
      __syncthreads();
      if (tid < NSL) {
        int o = 0;
        for (int s = S.dlo[tid]; s <= S.dhi[tid]; s++) o |= S.overflow[s % RING];
        S.ovf[tid] = o;
        if (!SECOND && o && nraw[tid]) {
          atomicOr(&A.item_ovf[item], 1u << ((d + tid - d_begin) >> TK_OVF_SHIFT));
          A.batch_words[0] = 1u;
        }
      }

      for (uint32_t t0 = 0; t0 < nmax; t0 += TK_MAXT) {
        __syncthreads();
        if (tid < NSL) S.ntg[tid] = 0;
        if (tid == 0) S.tnext = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NSL; w++)
          for (uint32_t f = t0 + tid; f < min(nraw[w], t0 + TK_MAXT); f += THREADS) {
            const uint32_t n0 = tb_[w][0] - ta_[w][0];
            const uint32_t pos = f < n0 ? ta_[w][0] + f : ta_[w][1] + (f - n0);
            const uint4 r = A.rec[pos];
            const int yi = (int)(r.x >> 16);
            // fast-path conditions: not a halo event, window rows stay below 2H, staging complete
            const bool ok = (int)r.z >= A.h && min(yi + FARMS_MAX_WINDOW, W - 1) <= 2 * H - 1 && !S.ovf[w] &&
                            (!SECOND || !A.done[pos]);
            if (ok) S.tlist[w][atomicAdd(&S.ntg[w], 1u)] = pos;
          }
        __syncthreads();
        uint32_t tstart[NSL + 1];
        tstart[0] = 0;
#pragma unroll
        for (int w = 0; w < NSL; w++) tstart[w + 1] = tstart[w] + S.ntg[w];
        const uint32_t tasks = tstart[NSL];

        // ---- one warp pools one event ----
        for (;;) {
          uint32_t k = 0;
          if (lane == 0) k = atomicAdd(&S.tnext, 1u);
          k = __shfl_sync(0xffffffffu, k, 0);
          if (k >= tasks) break;
          int w = 0;
#pragma unroll
          for (int q = 1; q < NSL; q++) w += (k >= tstart[q]) ? 1 : 0;
          const uint32_t tpos = S.tlist[w][k - tstart[w]];
          const uint4 r = A.rec[tpos];
          const int xi = (int)(r.x & 0xffffu), yi = (int)(r.x >> 16);
          const uint32_t ii = r.z;
          // the event's window in region coordinates; staged records all lie inside the sensor and inside the
          // reference's row bound (width - 1), so "in the window" is |dx| <= 50 and |dy| <= 50
          const uint32_t tw = (uint32_t)(xi - R.rx0) | ((uint32_t)(yi - R.ry0) << 8);
#pragma unroll
          for (int q = 0; q < FARMS_NSCALES; q++) S.acc[warp][q][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int sl = S.dlo[w], sh = S.dhi[w];
          int cnt = 0;
          for (int s = sl; s <= sh; s++) {
            const int slot = s % RING;
            const int n = S.count[slot];
            const uint32_t iir = ii - S.slot_base[slot];
            ncand += (lane == 0) ? n : 0;
            const uint2 *recs = &S.ta[slot * STRIDE];
            for (int g0 = 0; g0 < n; g0 += 128) {
              if (__any_sync(0xffffffffu, cnt > WQ_DEPTH - 4)) {  // rare: a lane's queue could overflow this trip
                __syncwarp();
                drain_queue(S, warp, lane, cnt, tw);
                cnt = 0;
              }
              uint2 c[4];
#pragma unroll
              for (int u = 0; u < 4; u++)  // the last trip reads into the zeroed padding (span 0)
                c[u] = (g0 + 32 * u < n) ? recs[g0 + 32 * u + lane] : make_uint2(0u, 0u);
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const uint32_t d4 = __vabsdiffu4(c[u].x, tw);
                const uint32_t rel = __funnelshift_r(c[u].x, c[u].y, 16) & 0x00ffffffu;
                // still the latest event of its pixel AND younger than 500 us  <=>  idx <= ii < end
                const bool ok = (iir - rel) < (c[u].y >> 8) && ((d4 + 0x4d4du) & 0x8080u) == 0u;
                if (ok) {
                  S.queue[warp][cnt][lane] = (uint16_t)(slot * STRIDE + g0 + 32 * u + lane);
                  cnt++;
                }
              }
            }
          }
          __syncwarp();
          drain_queue(S, warp, lane, cnt, tw);
          __syncwarp();
          // lanes 0..10 and 11..21 each combine 16 of ring (lane % 11)'s 32 per-lane partials: four at a time in
          // FP32, the group sums in FP64 (the counts are exact either way)
          double rl = 0.0, rx = 0.0, ry = 0.0;
          float rnf = 0.f;
          if (lane < 2 * FARMS_NSCALES) {
            const int rg = lane < FARMS_NSCALES ? lane : lane - FARMS_NSCALES, hb = lane < FARMS_NSCALES ? 0 : 16;
#pragma unroll
            for (int q4 = 0; q4 < 4; q4++) {
              float4 g = S.acc[warp][rg][hb | ((4 * q4 + rg) & 15)];
#pragma unroll
              for (int q = 1; q < 4; q++) {
                const float4 v = S.acc[warp][rg][hb | ((4 * q4 + q + rg) & 15)];
                g.x += v.x;
                g.y += v.y;
                g.z += v.z;
                g.w += v.w;
              }
              rl += (double)g.x;
              rx += (double)g.y;
              ry += (double)g.z;
              rnf += g.w;
            }
          }
          __syncwarp();
          {
            const double ol = __shfl_down_sync(0xffffffffu, rl, FARMS_NSCALES), ox = __shfl_down_sync(0xffffffffu, rx, FARMS_NSCALES),
                         oy = __shfl_down_sync(0xffffffffu, ry, FARMS_NSCALES);
            const float on = __shfl_down_sync(0xffffffffu, rnf, FARMS_NSCALES);
            if (lane < FARMS_NSCALES) {
              rl += ol;
              rx += ox;
              ry += oy;
              rnf += on;
            } else {
              rl = rx = ry = 0.0;
              rnf = 0.f;
            }
          }
          const bool lower = lane < 16;
          const bool fin = finish_event_checked(A, sub, rl, rx, ry, (int)rnf, (int)ii - A.h, lower);
          const bool fin0 = __shfl_sync(0xffffffffu, fin, 0);
          if (lane == 0 && fin0) {
            A.done[tpos] = 1;
            npooled++;
          }
          if (!fin0) {
            // ---- undecided (a rival scale within the FP32 noise, cancelling vectors): pool it again exactly, FP64
            // partials and the FP64 flow values of the contributors (lanes 0..15; the FP64 partials of one target
            // take the warp's whole accumulator space) ----
            __syncwarp();
            double4 *dacc = reinterpret_cast<double4 *>(&S.acc[warp][0][0]);  // [ring][16 lanes] {len, lcx, lcy, n}
            if (lower) {
#pragma unroll
              for (int q = 0; q < FARMS_NSCALES; q++) dacc[q * 16 + sub] = make_double4(0.0, 0.0, 0.0, 0.0);
              for (int s = sl; s <= sh; s++) {
                const int slot = s % RING;
                const int n = S.count[slot];
                const uint32_t base = S.slot_base[slot], iir = ii - base;
                for (int q0 = sub; q0 < n; q0 += 16) {
                  const uint2 c = S.ta[slot * STRIDE + q0];
                  const uint32_t d4 = __vabsdiffu4(c.x, tw);
                  const uint32_t rel = __funnelshift_r(c.x, c.y, 16) & 0x00ffffffu;
                  const bool ok = (iir - rel) < (c.y >> 8) && ((d4 + 0x4d4du) & 0x8080u) == 0u;
                  if (ok) {
                    const uint32_t mch = max(d4 & 0xffu, (d4 >> 8) & 0xffu);
                    const int ring = (int)(((mch + FARMS_WINDOW_JUMP - 1) * 205u) >> 10);
                    const uint32_t j = base + rel;
                    double4 v = dacc[ring * 16 + sub];
                    v.x += A.ev_len[j];
                    v.y += A.ev_lcx[j];
                    v.z += A.ev_lcy[j];
                    v.w += 1.0;
                    dacc[ring * 16 + sub] = v;
                  }
                }
              }
            }
            __syncwarp();
            double el = 0.0, ex2 = 0.0, ey2 = 0.0, en = 0.0;
            if (lower && sub < FARMS_NSCALES) {
#pragma unroll 4
              for (int q = 0; q < 16; q++) {
                const double4 v = dacc[sub * 16 + ((q + sub) & 15)];
                el += v.x;
                ex2 += v.y;
                ey2 += v.z;
                en += v.w;
              }
            }
            __syncwarp();
            finish_event<16>(A, sub, el, ex2, ey2, en, A.ev_lcx[ii], A.ev_lcy[ii], (int)ii - A.h, lower);
            if (lane == 0) {
              A.done[tpos] = 1;
              npooled++;
            }
            __syncwarp();
          }
        }
      }
    }
  }
  if (lane == 0 && ncand) atomicAdd(A.cand_count, ncand);
  if (lane == 0 && npooled) {
    atomicAdd(A.path_count + (SECOND ? 1 : 0), (unsigned long long)npooled);
    atomicAdd(A.batch_words + 2, npooled);
  }
}

template <int WARPS, int CAP, int NSL, int CTAS, bool SECOND>
void launch_warp(const PoolArgs &A0, int nslabs, int num_sms, cudaStream_t s) {
  PoolArgs A = A0;
  using SM = WarpSmem<WARPS, CAP, NSL>;
  static_assert(sizeof(SM) <= (CTAS == 2 ? 115712 : 232448), "shared memory of the warp kernel: 227 KB per CTA, 228 KB per SM");
  static_assert(SM::RING * SM::STRIDE <= 65536, "queue entries are 16-bit slot addresses");
  auto kern = k_pool_warp<WARPS, CAP, NSL, CTAS, SECOND>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  const int otx = (A.g.W + OT - 1) >> OT_SHIFT, oty = (A.g.H + OT - 1) >> OT_SHIFT;
  const int nseg = (nslabs + TK_SEG - 1) / TK_SEG;
  const long long items = (long long)otx * oty * nseg;
  unsigned grid = (unsigned)std::min<long long>(items, (long long)num_sms * CTAS);
  kern<<<grid, WARPS * 32, sizeof(SM), s>>>(A, otx, oty, nseg);
}

template <int WARPS, int CAP, int NSL, bool XCULL = false>
struct PackedSmem {
  static constexpr int RING = TK_LB + NSL, PAD = TK_PAD, STRIDE = CAP + PAD;
  uint2 ta[RING * STRIDE];   // {x_rel | y_rel << 8 | idx_rel[15:0] << 16,  idx_rel[23:16] | span << 8}
  float2 pb[RING * STRIDE];  // |flow|cos(theta), |flow|sin(theta)
  float4 acc[WARPS][FARMS_NSCALES][32];  // per-lane ring partials: len, lcx, lcy, count
  uint32_t tlist[NSL][TK_MAXT];
  uint32_t run_s[RING * TK_MAXRUN], run_o[RING * TK_MAXRUN + 1];
  uint8_t run_info[RING * TK_MAXRUN];
  uint32_t slab_f[RING + 1], slab_pre[RING + 1];
  uint32_t wcount[WARPS];
  uint32_t slot_base[RING];
  int tag[RING], count[RING], overflow[RING];
  int dlo[NSL], dhi[NSL], ovf[NSL];
  unsigned int ntg[NSL], tnext, item;
  // XCULL: per slot, where the staged records of tile-column run c begin ([nrun] = the slot's count)
  uint16_t col_off[XCULL ? TK_LB + NSL : 1][TK_MAXRUN + 2];
};

// k_pool_tile16: k_pool_tile on the 16-byte packed records of stage_slabs_packed (8-byte test record: one LDS.64,
// VABSDIFF4 window test, 24-bit index / span; float2 payload).  A fifth less shared memory per staged record buys
// four slabs per round at dense-stream slot sizes (twice the tasks per round, half the rounds).
// SECOND: a later pass over the same items with larger slots; only rounds that still hold undone targets (their
// staging overflowed the slots of the first pass: locally dense scenes) do any work.
// XCULL: a slot keeps its records in tile-column runs (16 pixels wide, the order of the pooling index); a target
// tests only the runs its 101-pixel window touches -- 7 or 8 of the region's 9 or 10 -- instead of the whole slot.
// The two halves of a warp walk the union of their two ranges (same shared-memory addresses for both: separate
// ranges measured 3 % slower than no culling at all, the record loads then cost two wavefronts), and the targets of
// a round are listed in index order (tile column by tile column) so that the two targets of a pair mostly want the
// same columns.
template <int WARPS, int CAP, int NSL, int CTAS, bool SECOND, bool XCULL>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k_pool_tile16(PoolArgs A, int otx_n, int oty_n, int nseg) {
  using SM = PackedSmem<WARPS, CAP, NSL, XCULL>;
  constexpr int STRIDE = SM::STRIDE;
  constexpr int THREADS = WARPS * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4, sub = lane & 15;
  const int W = A.g.W, H = A.g.H, nty = A.g.nty, NT = A.g.ntx * A.g.nty;
  const unsigned int nitems = (unsigned int)otx_n * oty_n * nseg;
  unsigned long long ncand = 0;
  unsigned int npooled = 0;

  if (SECOND && A.batch_words[0] == 0u) return;  // no round overflowed in the first pass: nothing to do

  for (;;) {
    __syncthreads();
    if (tid == 0) S.item = atomicAdd(A.work_counter, 1u);
    if (tid < (TK_LB + NSL)) S.tag[tid] = -1;
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

    // NSL consecutive slabs per round
    for (int d = d_begin; d < d_end; d += NSL) {
      const int nd = min(NSL, d_end - d);
      // ---- targets of this round: flow events of the owner tile in slabs d .. d+nd-1 ----
      uint32_t ta[NSL][2], tb[NSL][2], nraw[NSL];
      uint32_t nraw_all = 0, nmax = 0;
#pragma unroll
      for (int w = 0; w < NSL; w++) {
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
      if (SECOND) {  // only rounds the first pass flagged do any work here
        const uint32_t bits = A.item_ovf[item];
        const int b0 = (d - d_begin) >> TK_OVF_SHIFT, b1 = (d + nd - 1 - d_begin) >> TK_OVF_SHIFT;
        if (((bits >> b0) & ((2u << (b1 - b0)) - 1u)) == 0u) continue;
      }

      // ---- make sure the slabs of all windows of the round are staged ----
      __syncthreads();  // the previous round is done with S.dlo / S.dhi / S.ovf
      if (tid < NSL) {
        const int dd = min(d + tid, d_end - 1);
        const uint32_t t_first = A.slab_ids[dd] << A.g.slab_shift;
        const uint32_t lo_id =
            (t_first >= (uint32_t)(FARMS_KILL_OLD_FLOW_TIME - 1) ? t_first - (FARMS_KILL_OLD_FLOW_TIME - 1) : 0u) >> A.g.slab_shift;
        int l = dd;
        while (l > 0 && dd - l < TK_LB && A.slab_ids[l - 1] >= lo_id) l--;
        S.dlo[tid] = l;
        S.dhi[tid] = dd;
      }
      const uint32_t i_round = A.slab_first[d];
      __syncthreads();
      int s_first = 0x7fffffff, s_last = -1;
#pragma unroll
      for (int w = 0; w < NSL; w++) {
        if (nraw[w]) {
          s_first = min(s_first, S.dlo[w]);
          s_last = max(s_last, S.dhi[w]);
        }
      }
      // staged slabs are a contiguous range ending at the last staged slab, so what is missing is a suffix
      int s_new = s_first;
      while (s_new <= s_last && S.tag[s_new % (TK_LB + NSL)] == s_new) s_new++;  // uniform: tags are read after a barrier
      if (s_new <= s_last) stage_slabs_packed<SM, WARPS, CAP, NSL, XCULL>(A, S, s_new, s_last, R, i_round);
      __syncthreads();
      if (tid < NSL) {
        int o = 0;
        for (int s = S.dlo[tid]; s <= S.dhi[tid]; s++) o |= S.overflow[s % (TK_LB + NSL)];
        S.ovf[tid] = o;
        if (!SECOND && o && nraw[tid]) {
          atomicOr(&A.item_ovf[item], 1u << ((d + tid - d_begin) >> TK_OVF_SHIFT));
          A.batch_words[0] = 1u;
        }
      }

      for (uint32_t t0 = 0; t0 < nmax; t0 += TK_MAXT) {
        __syncthreads();
        if (tid < NSL) S.ntg[tid] = 0;
        if (tid == 0) S.tnext = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NSL; w++) {
          const uint32_t lim = min(nraw[w], t0 + TK_MAXT), n0 = tb[w][0] - ta[w][0];
          // fast-path conditions: not a halo event, window rows stay below 2H, staging complete
          auto wanted = [&](const uint32_t pos) {
            const uint4 r = A.rec[pos];
            const int yi = (int)(r.x >> 16);
            return (int)r.z >= A.h && min(yi + FARMS_MAX_WINDOW, W - 1) <= 2 * H - 1 && !S.ovf[w] &&
                   (!SECOND || !A.done[pos]);
          };
          if (XCULL) {  // a warp lists its 32 consecutive index positions in order
            for (uint32_t fb = t0 + warp * 32; fb < lim; fb += THREADS) {
              const uint32_t f = fb + lane;
              const uint32_t pos = f < n0 ? ta[w][0] + f : ta[w][1] + (f - n0);
              const bool ok = f < lim && wanted(pos);
              const unsigned bal = __ballot_sync(0xffffffffu, ok);
              unsigned int base = 0;
              if (lane == 0 && bal) base = atomicAdd(&S.ntg[w], (unsigned int)__popc(bal));
              base = __shfl_sync(0xffffffffu, base, 0);
              const unsigned int slot_t = base + __popc(bal & ((1u << lane) - 1u));
              if (ok && FARMS_CHK(slot_t < (unsigned int)TK_MAXT && pos < A.cell_start[A.ncells], 111))
                S.tlist[w][slot_t] = pos;
            }
          } else {
            for (uint32_t f = t0 + tid; f < lim; f += THREADS) {
              const uint32_t pos = f < n0 ? ta[w][0] + f : ta[w][1] + (f - n0);
              if (wanted(pos)) {
                const unsigned int slot_t = atomicAdd(&S.ntg[w], 1u);
                if (FARMS_CHK(slot_t < (unsigned int)TK_MAXT && pos < A.cell_start[A.ncells], 111)) S.tlist[w][slot_t] = pos;
              }
            }
          }
        }
        __syncthreads();
        // Tasks: first pairs of targets of the same slab (lanes 0-15 pool one event, lanes 16-31 the next; both
        // halves share their loop bounds), then "solo" targets pooled by the two halves together, each half taking
        // every other 64-record trip (half the time of a pair task).  The odd target of a slab is a solo, and so
        // are the pairs that would not divide evenly among the warps when they are at most half a wave: the
        // round then ends half a task later instead of a whole one.
        // (Rounds of many slabs -- the sparse-stream variant -- keep the plain pairing: there an odd target shares
        // its warp with an idle half.)
        constexpr bool SOLOS = NSL <= FARMS_SOLOS_MAX_NSL;
        uint32_t pstart[NSL + 1], sstart[NSL + 1], npair[NSL];
        uint32_t P = 0;
#pragma unroll
        for (int w = 0; w < NSL; w++) {
          npair[w] = SOLOS ? S.ntg[w] >> 1 : (S.ntg[w] + 1) >> 1;
          P += npair[w];
        }
        uint32_t extra = SOLOS ? P % WARPS : 0u;
        if (extra > WARPS / 2) extra = 0;
#pragma unroll
        for (int w = NSL - 1; w >= 0; w--) {
          const uint32_t take = min(extra, npair[w]);
          npair[w] -= take;
          extra -= take;
        }
        pstart[0] = sstart[0] = 0;
#pragma unroll
        for (int w = 0; w < NSL; w++) {
          pstart[w + 1] = pstart[w] + npair[w];
          sstart[w + 1] = sstart[w] + (SOLOS ? S.ntg[w] - 2 * npair[w] : 0u);
        }
        const uint32_t tasks_p = pstart[NSL], tasks = tasks_p + sstart[NSL];

        for (;;) {
          uint32_t k = 0;
          if (lane == 0) k = atomicAdd(&S.tnext, 1u);
          k = __shfl_sync(0xffffffffu, k, 0);
          if (k >= tasks) break;
          const bool solo = k >= tasks_p;
          int w = 0;
          uint32_t kk;
          if (!solo) {
#pragma unroll
            for (int q = 1; q < NSL; q++) w += (k >= pstart[q]) ? 1 : 0;
            kk = (k - pstart[w]) * 2 + half;
          } else {
            const uint32_t j = k - tasks_p;
#pragma unroll
            for (int q = 1; q < NSL; q++) w += (j >= sstart[q]) ? 1 : 0;
            kk = 2 * (pstart[w + 1] - pstart[w]) + (j - sstart[w]);
          }
          const uint32_t nt = S.ntg[w];
          const bool have = solo ? half == 0 : kk < nt;  // this half owns a target's result
          (void)FARMS_CHK(nt >= 1 && nt <= (uint32_t)TK_MAXT && w >= 0 && w < NSL, 112);
          const uint32_t tpos = S.tlist[w][kk < nt ? kk : nt - 1];
          const uint4 r = A.rec[tpos];
          (void)FARMS_CHK((int)r.z >= A.h && (size_t)r.z < A.m && xi_in_region(r.x, R), 113);
          const int xi = (int)(r.x & 0xffffu), yi = (int)(r.x >> 16);
          const uint32_t ii = r.z;
          // the event's window in region coordinates; staged records all lie inside the sensor and inside the
          // reference's row bound (width - 1, :1000), so "in the window" is |dx| <= 50 and |dy| <= 50
          const uint32_t tw = (uint32_t)(xi - R.rx0) | ((uint32_t)(yi - R.ry0) << 8);
#pragma unroll
          for (int q = 0; q < FARMS_NSCALES; q++) S.acc[warp][q][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int sl = S.dlo[w], sh = S.dhi[w];
          // XCULL: the tile-column runs this target's window touches.  Runs 0 .. xr_n0-1 hold the region's columns for
          // rows < H, the runs after them the columns of the aliased pixels (i + 1, j - H), whose PHYSICAL x is one
          // more than the window coordinate they are tested with.
          int xc_lo = 0, xc_hi = 0, xa_lo = 0, xa_hi = -1, nparts = 1;
          if (XCULL) {
            const int ts = A.g.tile_shift, tx0 = R.rx0 >> ts, tx1 = R.rx1 >> ts;
            const int xr_n0 = R.ry1 >= R.ry0 ? tx1 - tx0 + 1 : 0;
            xc_lo = max((xi - FARMS_MAX_WINDOW) >> ts, tx0) - tx0;
            xc_hi = xr_n0 ? min((xi + FARMS_MAX_WINDOW) >> ts, tx1) - tx0 + 1 : 0;  // (one past the last run)
            if (!xr_n0) xc_lo = 0;
            if (R.ay1 >= 0) {
              const int atx0 = R.ax0 >> ts, atx1 = R.ax1 >> ts;
              xa_lo = xr_n0 + max((xi + 1 - FARMS_MAX_WINDOW) >> ts, atx0) - atx0;
              xa_hi = xr_n0 + min((xi + 1 + FARMS_MAX_WINDOW) >> ts, atx1) - atx0 + 1;
              nparts = 2;
            }
            // both halves walk the union of their two ranges: they then read the same records (one shared-memory
            // wavefront per load instead of two) and the trip loops stay uniform across the warp
            xc_lo = min(xc_lo, __shfl_xor_sync(0xffffffffu, xc_lo, 16));
            xc_hi = max(xc_hi, __shfl_xor_sync(0xffffffffu, xc_hi, 16));
            xa_lo = min(xa_lo, __shfl_xor_sync(0xffffffffu, xa_lo, 16));
            xa_hi = max(xa_hi, __shfl_xor_sync(0xffffffffu, xa_hi, 16));
          }
          for (int s = sl; s <= sh; s++) {
            const int slot = s % (TK_LB + NSL);
            const int n = S.count[slot];
            const uint32_t iir = ii - S.slot_base[slot];
            const uint2 *recs = &S.ta[slot * STRIDE];
            const float2 *pays = &S.pb[slot * STRIDE];
            if (!XCULL) ncand += (sub == 0 && have) ? n : 0;
            // one staged record against this half-warp's event: test, and on a pass add it to the lane's ring partials
            auto pool_one = [&](const uint2 c, const int q) {
              const uint32_t d4 = __vabsdiffu4(c.x, tw);  // |dx| in byte 0, |dy| in byte 1
              const uint32_t rel = __funnelshift_r(c.x, c.y, 16) & 0x00ffffffu;
              // still the latest event of its pixel AND younger than 500 us  <=>  idx <= ii < end
              const bool ok = (iir - rel) < (c.y >> 8) && ((d4 + 0x4d4du) & 0x8080u) == 0u;
              if (ok) {
                const uint32_t m = max(d4 & 0xffu, (d4 >> 8) & 0xffu);
                const uint32_t ring = ((m + FARMS_WINDOW_JUMP - 1) * 205u) >> 10;  // /5 for values <= 54
                (void)FARMS_CHK(ring < (uint32_t)FARMS_NSCALES && q < n, 115);
                const float2 f = pays[q];
                float fl;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(fl) : "f"(__fmaf_rn(f.y, f.y, f.x * f.x)));  // (file builds with -fmad=false)
                float4 v = S.acc[warp][ring][lane];
                v.x += fl;
                v.y += f.x;
                v.z += f.y;
                v.w += 1.f;
                S.acc[warp][ring][lane] = v;
              }
            };
            // groups of 16 records: four per trip while four remain, the last one to three singly (a slot holds
            // ~260 records: a fifth trip of four groups would test 60 padding records)
            for (int part = 0; part < nparts; part++) {  // XCULL: the window's columns, then its aliased columns
              int beg = 0, end = n;
              if (XCULL) {
                beg = S.col_off[slot][part ? xa_lo : xc_lo];
                end = max((int)S.col_off[slot][part ? xa_hi : xc_hi], beg);
                (void)FARMS_CHK(end <= n && (part ? xa_hi : xc_hi) <= TK_MAXRUN, 118);
                ncand += (sub == 0 && have) ? end - beg : 0;
              }
              // (the last trip of four may read up to 15 records past `end`: zeroed padding, or records of the next
              // column, which lies outside the window -- except where aliased runs follow: there only whole groups)
              const int ngroups = (XCULL && nparts == 2 && part == 0) ? (end - beg) >> 4 : (end - beg + 15) >> 4;
              const int nfull = beg + ((ngroups & ~3) << 4);
              for (int q0 = beg + sub + (solo ? 64 * half : 0); q0 < nfull; q0 += solo ? 128 : 64) {
                uint2 c[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                  (void)FARMS_CHK(q0 + 16 * u < STRIDE && slot >= 0 && slot < SM::RING, 114);
                  c[u] = recs[q0 + 16 * u];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) pool_one(c[u], q0 + 16 * u);
              }
              // (a solo target's two halves take alternate tail groups)
              for (int q0 = nfull + sub + (solo ? 16 * half : 0); q0 < end; q0 += solo ? 32 : 16) {
                (void)FARMS_CHK(q0 < STRIDE, 117);
                pool_one(recs[q0], q0);
              }
            }
          }
          __syncwarp();
          // sub-lane k < 11 of each half combines ring k's 16 per-lane partials: four at a time in FP32, the four
          // group sums in FP64 (the counts are exact either way)
          double rl = 0.0, rx = 0.0, ry = 0.0;
          float rnf = 0.f;
          if (sub < FARMS_NSCALES) {
#pragma unroll
            for (int q4 = 0; q4 < 4; q4++) {
              float4 g = S.acc[warp][sub][(half << 4) | ((4 * q4 + sub) & 15)];
#pragma unroll
              for (int q = 1; q < 4; q++) {
                const float4 v = S.acc[warp][sub][(half << 4) | ((4 * q4 + q + sub) & 15)];
                g.x += v.x;
                g.y += v.y;
                g.z += v.z;
                g.w += v.w;
              }
              rl += (double)g.x;
              rx += (double)g.y;
              ry += (double)g.z;
              rnf += g.w;
            }
          }
          __syncwarp();
          if (solo) {  // the two halves pooled disjoint trips of the same target
            rl += __shfl_xor_sync(0xffffffffu, rl, 16);
            rx += __shfl_xor_sync(0xffffffffu, rx, 16);
            ry += __shfl_xor_sync(0xffffffffu, ry, 16);
            rnf += __shfl_xor_sync(0xffffffffu, rnf, 16);
          }
          const bool fin = finish_event_checked(A, sub, rl, rx, ry, (int)rnf, (int)ii - A.h, have);

          if (sub == 0 && fin) {
            A.done[tpos] = 1;
            npooled++;
          }
          // ---- undecided targets (a rival scale within the FP32 noise, cancelling vectors): pool them again
          // exactly, FP64 partials and the FP64 flow values of the contributors, one half-warp at a time
          // (the FP64 partials of one target take the warp's whole accumulator space) ----
          const unsigned need = __ballot_sync(0xffffffffu, have && !fin && sub == 0);
          for (int hsel = 0; hsel < 2; hsel++) {
            if (!((need >> (16 * hsel)) & 1u)) continue;  // uniform across the warp
            __syncwarp();
            double4 *dacc = reinterpret_cast<double4 *>(&S.acc[warp][0][0]);  // [ring][16 lanes] {len, lcx, lcy, n}
            if (half == hsel) {
#pragma unroll
              for (int q = 0; q < FARMS_NSCALES; q++) dacc[q * 16 + sub] = make_double4(0.0, 0.0, 0.0, 0.0);
              for (int s = sl; s <= sh; s++) {
                const int slot = s % (TK_LB + NSL);
                const int n = S.count[slot];
                const uint32_t base = S.slot_base[slot], iir = ii - base;
                for (int q0 = sub; q0 < n; q0 += 16) {
                  const uint2 c = S.ta[slot * STRIDE + q0];
                  const uint32_t d4 = __vabsdiffu4(c.x, tw);
                  const uint32_t rel = __funnelshift_r(c.x, c.y, 16) & 0x00ffffffu;
                  const bool ok = (iir - rel) < (c.y >> 8) && ((d4 + 0x4d4du) & 0x8080u) == 0u;
                  if (ok) {
                    const uint32_t mch = max(d4 & 0xffu, (d4 >> 8) & 0xffu);
                    const int ring = (int)(((mch + FARMS_WINDOW_JUMP - 1) * 205u) >> 10);
                    const uint32_t j = base + rel;
                    (void)FARMS_CHK(ring < FARMS_NSCALES && (size_t)j < A.m && j <= ii, 116);
                    double4 v = dacc[ring * 16 + sub];
                    v.x += A.ev_len[j];
                    v.y += A.ev_lcx[j];
                    v.z += A.ev_lcy[j];
                    v.w += 1.0;
                    dacc[ring * 16 + sub] = v;
                  }
                }
              }
            }
            __syncwarp();
            double el = 0.0, ex2 = 0.0, ey2 = 0.0, en = 0.0;
            if (half == hsel && sub < FARMS_NSCALES) {
#pragma unroll 4
              for (int q = 0; q < 16; q++) {
                const double4 v = dacc[sub * 16 + ((q + sub) & 15)];
                el += v.x;
                ex2 += v.y;
                ey2 += v.z;
                en += v.w;
              }
            }
            __syncwarp();
            finish_event<16>(A, sub, el, ex2, ey2, en, A.ev_lcx[ii], A.ev_lcy[ii], (int)ii - A.h, half == hsel);
            if (half == hsel && sub == 0) {
              A.done[tpos] = 1;
              npooled++;
            }
          }
        }
      }
    }
  }
  if ((lane & 15) == 0 && ncand) atomicAdd(A.cand_count, ncand);
  if ((lane & 15) == 0 && npooled) {
    atomicAdd(A.path_count + (SECOND ? 1 : 0), (unsigned long long)npooled);
    atomicAdd(A.batch_words + 2, npooled);
  }
}

template <int WARPS, int CAP, int NSL, int CTAS, bool SECOND, bool XCULL = false>
void launch_tile16(const PoolArgs &A0, int nslabs, int num_sms, cudaStream_t s) {
  PoolArgs A = A0;
  using SM = PackedSmem<WARPS, CAP, NSL, XCULL>;
  static_assert(sizeof(SM) <= (CTAS == 2 ? 115712 : 232448), "shared memory of the tile kernel: 227 KB per CTA, 228 KB per SM");
  auto kern = k_pool_tile16<WARPS, CAP, NSL, CTAS, SECOND, XCULL>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  const int otx = (A.g.W + OT - 1) >> OT_SHIFT, oty = (A.g.H + OT - 1) >> OT_SHIFT;
  const int nseg = (nslabs + TK_SEG - 1) / TK_SEG;
  const long long items = (long long)otx * oty * nseg;
  unsigned grid = (unsigned)std::min<long long>(items, (long long)num_sms * CTAS);
  kern<<<grid, WARPS * 32, sizeof(SM), s>>>(A, otx, oty, nseg);
}

// ------------------------------------------------------------------------------------------------
// fast path, bit-parallel: prefix bitmask tables over the staged records
// ------------------------------------------------------------------------------------------------
// A CTA owns a 32x32 owner tile and advances through its time slabs in rounds.  Per round it stages the flow
// events of the tile's (32+100)^2 region (<= 2048 records: the look-back slabs plus the round's own slabs) and
// builds three bit tables over the staged POSITIONS:
//   PX[j]  records whose window column is < j          (prefix-OR over 133 rows)
//   PY[j]  records whose (logical) window row is < j
//   AL[q]  records that are contributors of query q in time:  idx <= i_q < end   (prefix-XOR over the
//          queries sorted by index: a record toggles its bit at the first and past-the-last query it serves)
// The contributors of query q inside scale k's square are then  (PX[xb]&~PX[xa]) & (PY[yb]&~PY[ya]) & AL[q]
// -- a few 128-bit loads and LOP3s for 256 records per lane -- and ring k is that mask minus scale k-1's.
// Ring sums are accumulated in registers (the ring is static per pass) by walking the set bits.
// Eight lanes serve one query (lane i owns the records with position = i mod 8), four queries per warp.
// FP32 partial sums, FP64 combination and the same safety margin as k_pool_tile: an event whose scale decision
// is not clear-cut is left to k_pool_any.
constexpr int BP_WARPS = 8, BP_THREADS = BP_WARPS * 32;
constexpr int BP_LW = 8;                    // mask words per lane
constexpr int BP_CAP = BP_LW * 256;         // staged records per round
constexpr int BP_PITCH = BP_LW * 8 + 4;     // words per table row (+4: conflict-free 128-bit column access)
constexpr int BP_ROWQ = BP_PITCH / 4;       // uint4 per row
constexpr int BP_NROW = OT + 2 * FARMS_MAX_WINDOW + 1;
constexpr int BP_MAXQ = 64;                 // queries per round
constexpr int BP_NDMAX = 8;  // query slabs per round at most
constexpr int BP_MAXSLAB = TK_LB + BP_NDMAX;  // + look-back slabs
constexpr int BP_RUNS_PER_SLAB = 20;        // <= 10 tile columns for rows < H plus <= 10 aliased
constexpr int BP_MAXRUNS = 256;
static_assert(BP_MAXSLAB * BP_RUNS_PER_SLAB + 2 * BP_NDMAX <= BP_THREADS, "one thread per run descriptor");

struct BitsSmem {
  uint32_t px[BP_NROW][BP_PITCH];
  uint32_t py[BP_NROW][BP_PITCH];
  uint32_t al[BP_MAXQ + 1][BP_PITCH];
  // payload by staged position: |flow|cos, |flow|sin (|flow| is recomputed); entry BP_CAP stays (0, 0)
  float2 pxy[BP_CAP + 2];
  struct {
    uint32_t s[BP_MAXRUNS], n[BP_MAXRUNS], b[BP_MAXRUNS];  // start in the index, raw length | aliased << 31, count -> base
  } run;
  uint32_t q_pos[BP_MAXQ], q_xy[BP_MAXQ], q_ii[BP_MAXQ];                  // queries as loaded
  uint32_t s_pos[BP_MAXQ], s_xy[BP_MAXQ], s_ii[BP_MAXQ], s_ok[BP_MAXQ];   // sorted by event index
  uint32_t qrun_s[2 * BP_NDMAX], qrun_n[2 * BP_NDMAX];
  uint32_t slab_ub[BP_MAXSLAB], slab_cnt[BP_MAXSLAB];
  int nd, nq, count, skip;
  float fest;
  unsigned int item;
};

// first dense slab a query of dense slab dd can need (|dt| < 500 us, at most 4 slabs back)
__device__ __forceinline__ int slab_lookback(const PoolArgs &A, int dd) {
  const uint32_t t_first = A.slab_ids[dd] << A.g.slab_shift;
  const uint32_t lo_id =
      (t_first >= (uint32_t)(FARMS_KILL_OLD_FLOW_TIME - 1) ? t_first - (FARMS_KILL_OLD_FLOW_TIME - 1) : 0u) >> A.g.slab_shift;
  int l = dd;
  while (l > 0 && dd - l < TK_LB && A.slab_ids[l - 1] >= lo_id) l--;
  return l;
}

__device__ __forceinline__ uint4 and4(uint4 a, uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
__device__ __forceinline__ uint4 andn4(uint4 a, uint4 b) { return make_uint4(a.x & ~b.x, a.y & ~b.y, a.z & ~b.z, a.w & ~b.w); }
__device__ __forceinline__ uint4 or4(uint4 a, uint4 b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
__device__ __forceinline__ uint4 xor4(uint4 a, uint4 b) { return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }
template <bool XOR>
__device__ __forceinline__ uint4 comb4(uint4 a, uint4 b) { return XOR ? xor4(a, b) : or4(a, b); }
__device__ __forceinline__ uint4 shfl_up4(uint4 v, int d) {
  return make_uint4(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d),
                    __shfl_up_sync(0xffffffffu, v.z, d), __shfl_up_sync(0xffffffffu, v.w, d));
}

// Inclusive prefix (OR or XOR) down the rows of one uint4 column of a table; one warp, RPL consecutive rows
// per lane.  Row pitch 17 uint4 => the eight lanes of a quarter-warp hit eight different 16-byte bank groups.
template <bool XOR, int RPL>
__device__ __forceinline__ void prefix_column(uint32_t *table, int rows, int col4, int lane) {
  uint4 v[RPL];
  uint4 *base = reinterpret_cast<uint4 *>(table) + col4;
#pragma unroll
  for (int j = 0; j < RPL; j++) {
    const int r = lane * RPL + j;
    v[j] = r < rows ? base[r * BP_ROWQ] : make_uint4(0u, 0u, 0u, 0u);
    if (j) v[j] = comb4<XOR>(v[j], v[j - 1]);
  }
  uint4 tot = v[RPL - 1];
#pragma unroll
  for (int dlt = 1; dlt < 32; dlt <<= 1) {
    const uint4 o = shfl_up4(tot, dlt);
    if (lane >= dlt) tot = comb4<XOR>(tot, o);
  }
  uint4 carry = shfl_up4(tot, 1);
  if (lane == 0) carry = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int j = 0; j < RPL; j++) {
    const int r = lane * RPL + j;
    if (r < rows) base[r * BP_ROWQ] = comb4<XOR>(v[j], carry);
  }
}

// word offset inside a table row of staged position r: lane i = r & 7 owns it, as bit (r >> 3) of its LW words
__device__ __forceinline__ int bit_word(int r) {
  const int i = r & 7, k = r >> 8;
  return (((k >> 2) << 3) + i) * 4 + (k & 3);
}

__global__ void __launch_bounds__(BP_THREADS, 2) k_pool_bits(PoolArgs A, int otx_n, int oty_n, int nseg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BitsSmem &S = *reinterpret_cast<BitsSmem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 3, li = lane & 7;
  const int W = A.g.W, H = A.g.H, ts = A.g.tile_shift, nty = A.g.nty, NT = A.g.ntx * A.g.nty;
  const unsigned int nitems = (unsigned int)otx_n * oty_n * nseg;
  const double *pay_cx = A.pay + A.m, *pay_cy = A.pay + 2 * A.m;
  unsigned long long ncand = 0;
  unsigned int npooled = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) S.item = atomicAdd(A.work_counter, 1u);
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
    const int tx0 = R.rx0 >> ts, tx1 = R.rx1 >> ts, ty0 = R.ry0 >> ts, ty1 = R.ry1 >> ts;
    const int nrun0 = R.ry1 >= R.ry0 ? tx1 - tx0 + 1 : 0;  // empty on tall sensors whose row bound is above the region
    const int atx0 = R.ax0 >> ts, atx1 = R.ax1 >> ts;
    const int nrun1 = R.ay1 >= 0 ? atx1 - atx0 + 1 : 0;
    const int nrun = nrun0 + nrun1;
    const int aty1 = R.ay1 >= 0 ? (R.ay1 >> ts) : 0;
    // index tiles (16x16) of the owner tile: 2 columns x 2 rows, clipped
    const int itx0 = X0 >> 4, itx1 = min((X0 + OT - 1) >> 4, A.g.ntx - 1);
    const int ity0 = Y0 >> 4, ity1 = min((Y0 + OT - 1) >> 4, nty - 1);
    const int d_begin = seg * TK_SEG, d_end = min(d_begin + TK_SEG, A.nslabs);
    float fest = 0.75f;  // staged records / raw records of the covering index tiles, refined every round

    int d = d_begin;
    while (d < d_end) {
      // ---- A: run descriptors of the candidate slabs, query runs, clear the tables ----
      const int s_lo = slab_lookback(A, d);
      const int lb = d - s_lo;
      const int ncs = min(BP_NDMAX, d_end - d);
      const int nsl = lb + ncs;
      __syncthreads();  // the previous round is done with shared memory
      if (tid < nsl * nrun) {
        const int sl = tid / nrun, c = tid - sl * nrun;
        uint32_t a, b;
        if (c < nrun0) {
          const size_t cb = (size_t)(s_lo + sl) * NT + (size_t)(tx0 + c) * nty;
          a = A.cell_start[cb + ty0];
          b = A.cell_start[cb + ty1 + 1];
        } else {
          const size_t cb = (size_t)(s_lo + sl) * NT + (size_t)(atx0 + c - nrun0) * nty;
          a = A.cell_start[cb];
          b = A.cell_start[cb + aty1 + 1];
        }
        S.run.s[tid] = a;
        S.run.n[tid] = (b - a) | (c >= nrun0 ? 0x80000000u : 0u);
      }
      if (tid >= BP_THREADS - 2 * BP_NDMAX) {
        const int q = tid - (BP_THREADS - 2 * BP_NDMAX), w = q >> 1, c = q & 1;
        uint32_t a = 0, b = 0;
        if (w < ncs && itx0 + c <= itx1) {
          const size_t cb = (size_t)(d + w) * NT + (size_t)(itx0 + c) * nty;
          a = A.cell_start[cb + ity0];
          b = A.cell_start[cb + ity1 + 1];
        }
        S.qrun_s[q] = a;
        S.qrun_n[q] = b - a;
      }
      {
        uint4 *z = reinterpret_cast<uint4 *>(&S.px[0][0]);
        constexpr int NZ = (2 * BP_NROW + BP_MAXQ + 1) * BP_ROWQ;
        for (int q = tid; q < NZ; q += BP_THREADS) z[q] = make_uint4(0u, 0u, 0u, 0u);
        if (tid < 2) S.pxy[BP_CAP + tid] = make_float2(0.f, 0.f);
      }
      __syncthreads();
      // ---- A2 (warp 0): skip slabs without queries; candidate slab count from the raw-size estimate ----
      if (warp == 0) {
        if (lane < nsl) {
          uint32_t ub = 0;
          for (int c = 0; c < nrun; c++) ub += S.run.n[lane * nrun + c] & 0x7fffffffu;
          S.slab_ub[lane] = ub;
        }
        __syncwarp();
        if (lane == 0) {
          int w0 = 0;
          while (w0 < ncs && S.qrun_n[2 * w0] + S.qrun_n[2 * w0 + 1] == 0) w0++;
          if (w0 > 0) {  // nothing to pool in the first w0 slabs
            S.nd = w0;
            S.skip = 1;
          } else {
            uint32_t ubsum = 0, qsum = 0;
            for (int sl = 0; sl < lb; sl++) ubsum += S.slab_ub[sl];
            int ndc = 0;
            for (int w = 0; w < ncs; w++) {
              ubsum += S.slab_ub[lb + w];
              qsum += S.qrun_n[2 * w] + S.qrun_n[2 * w + 1];
              if (w > 0 && (qsum > (uint32_t)BP_MAXQ || fest * (float)ubsum > 1.04f * (float)BP_CAP)) break;
              ndc = w + 1;
            }
            S.nd = ndc;
            S.skip = 0;
          }
        }
      }
      __syncthreads();
      if (S.skip) {  // uniform
        d += S.nd;
        continue;
      }
      const int ndc = S.nd;
      const int nruns_c = (lb + ndc) * nrun;
      // ---- B: exact number of region records per run ----
      // a record whose successor at its pixel (or whose 500-us expiry) precedes the round is dead for every query
      const uint32_t i_round = A.slab_first[d];
      for (int run = warp; run < nruns_c; run += BP_WARPS) {
        const uint32_t s0 = S.run.s[run], nn = S.run.n[run];
        const uint32_t n = nn & 0x7fffffffu;
        const bool alias = (nn >> 31) != 0u;
        uint32_t cnt = 0;
        for (uint32_t o = 0; o < n; o += 32) {
          bool pass = false;
          if (o + lane < n) {
            const uint4 rec = A.rec[s0 + o + lane];
            const int x = (int)(rec.x & 0xffffu), y = (int)(rec.x >> 16);
            pass = !alias ? (x >= R.rx0 && x <= R.rx1 && y >= R.ry0 && y <= R.ry1)
                          : (x >= R.ax0 && x <= R.ax1 && y <= R.ay1);
            pass = pass && rec.w > i_round;
          }
          cnt += __popc(__ballot_sync(0xffffffffu, pass));
        }
        if (lane == 0) S.run.b[run] = cnt;
      }
      __syncthreads();
      // ---- B2 (warp 0): slabs that fit, positions of the runs ----
      if (warp == 0) {
        if (lane < lb + ndc) {
          uint32_t c = 0;
          for (int q = 0; q < nrun; q++) c += S.run.b[lane * nrun + q];
          S.slab_cnt[lane] = c;
        }
        __syncwarp();
        if (lane == 0) {
          uint32_t tot = 0, ub = 0, q = 0;
          for (int sl = 0; sl < lb; sl++) {
            tot += S.slab_cnt[sl];
            ub += S.slab_ub[sl];
          }
          int nd = 0;
          for (int w = 0; w < ndc; w++) {
            const uint32_t t2 = tot + S.slab_cnt[lb + w], q2 = q + S.qrun_n[2 * w] + S.qrun_n[2 * w + 1];
            if (w > 0 && (t2 > (uint32_t)BP_CAP || q2 > (uint32_t)BP_MAXQ)) break;
            tot = t2;
            q = q2;
            ub += S.slab_ub[lb + w];
            nd = w + 1;
          }
          S.nd = nd;
          S.nq = (int)min(q, (uint32_t)BP_MAXQ);  // a slab with more queries than that leaves the rest to k_pool_any
          S.count = (int)tot;
          S.skip = tot > (uint32_t)BP_CAP;        // even one slab does not fit: k_pool_any pools it
          S.fest = ub ? fminf(1.0f, (float)tot / (float)ub) : 0.75f;
        }
        __syncwarp();
        // exclusive prefix of the run counts over the runs in use
        const int nru = (lb + S.nd) * nrun;
        uint32_t loc[BP_MAXRUNS / 32];
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < BP_MAXRUNS / 32; j++) {
          const int run = lane * (BP_MAXRUNS / 32) + j;
          loc[j] = run < nru ? S.run.b[run] : 0u;
          sum += loc[j];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, inc, dlt);
          if (lane >= dlt) inc += o;
        }
        uint32_t pos0 = inc - sum;
#pragma unroll
        for (int j = 0; j < BP_MAXRUNS / 32; j++) {
          const int run = lane * (BP_MAXRUNS / 32) + j;
          if (run < nru) S.run.b[run] = pos0;
          pos0 += loc[j];
        }
      }
      __syncthreads();
      const int nd = S.nd, nq = S.nq, count = S.count;
      fest = S.fest;
      if (S.skip || nq == 0) {  // uniform
        d += nd;
        continue;
      }
      // ---- C: the round's queries = flow events of the owner tile in slabs d .. d+nd-1 ----
      if (tid < nq) {
        uint32_t f = (uint32_t)tid, pos = 0;
        for (int q = 0; q < 2 * nd; q++) {
          const uint32_t n = S.qrun_n[q];
          if (f < n) {
            pos = S.qrun_s[q] + f;
            break;
          }
          f -= n;
        }
        const uint4 r = A.rec[pos];
        S.q_pos[tid] = pos;
        S.q_xy[tid] = r.x;
        S.q_ii[tid] = r.z;
      }
      __syncthreads();
      if (tid < nq) {
        const uint32_t my = S.q_ii[tid];
        int rank = 0;
        for (int g = 0; g < nq; g++) rank += S.q_ii[g] < my ? 1 : 0;
        const uint32_t xy = S.q_xy[tid];
        const int yi = (int)(xy >> 16);
        // fast-path conditions: not a halo event, window rows stay below 2H, the event's own row is inside the
        // reference's width-1 row bound (always when W >= H)
        const bool ok = (int)my >= A.h && min(yi + FARMS_MAX_WINDOW, W - 1) <= 2 * H - 1 && yi <= W - 1;
        S.s_pos[rank] = S.q_pos[tid];
        S.s_xy[rank] = xy;
        S.s_ii[rank] = my;
        S.s_ok[rank] = ok ? 1u : 0u;
      }
      __syncthreads();
      // ---- D: stage the region records in run order (deterministic positions) and set their table bits ----
      {
        const int nru = (lb + nd) * nrun;
        const uint32_t ii_first = S.s_ii[0], ii_last = S.s_ii[nq - 1];
        for (int run = warp; run < nru; run += BP_WARPS) {
          const uint32_t s0 = S.run.s[run], nn = S.run.n[run];
          const uint32_t n = nn & 0x7fffffffu;
          const bool alias = (nn >> 31) != 0u;
          uint32_t base = S.run.b[run];
          for (uint32_t o = 0; o < n; o += 32) {
            bool pass = false;
            uint4 rec = make_uint4(0u, 0u, 0u, 0u);
            const uint32_t pos = s0 + o + lane;
            int x = 0, y = 0;
            if (o + lane < n) {
              rec = A.rec[pos];
              x = (int)(rec.x & 0xffffu);
              y = (int)(rec.x >> 16);
              if (!alias) {
                pass = x >= R.rx0 && x <= R.rx1 && y >= R.ry0 && y <= R.ry1;
              } else {
                pass = x >= R.ax0 && x <= R.ax1 && y <= R.ay1;
                x -= 1;   // logical window coordinates of the aliased cell
                y += H;
              }
              pass = pass && rec.w > i_round;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (pass) {
              const int r = (int)(base + __popc(bal & ((1u << lane) - 1u)));
              S.pxy[r] = make_float2(__double2float_rn(pay_cx[pos]), __double2float_rn(pay_cy[pos]));
              const int wo = bit_word(r);
              const uint32_t bit = 1u << ((r >> 3) & 31);
              atomicOr(&S.px[x - R.rx0 + 1][wo], bit);
              atomicOr(&S.py[y - R.ry0 + 1][wo], bit);
              // contributor of query q  <=>  idx <= i_q < end  <=>  qa <= q < qb  in index order
              const uint32_t idx = rec.z, end = rec.w;
              int qa = 0, qb = nq;
              if (idx > ii_first) {  // first query with i_q >= idx
                int lo = 0, hi = nq;
                while (lo < hi) {
                  const int mid = (lo + hi) >> 1;
                  if (S.s_ii[mid] >= idx) hi = mid; else lo = mid + 1;
                }
                qa = lo;
              }
              if (end <= ii_last) {  // first query with i_q >= end
                int lo = 0, hi = nq;
                while (lo < hi) {
                  const int mid = (lo + hi) >> 1;
                  if (S.s_ii[mid] >= end) hi = mid; else lo = mid + 1;
                }
                qb = lo;
              }
              if (qa < qb) {
                atomicXor(&S.al[qa][wo], bit);
                atomicXor(&S.al[qb][wo], bit);
              }
            }
            base += __popc(bal);
          }
        }
      }
      __syncthreads();
      // ---- E: prefix passes over the uint4 columns in use ----
      {
        const int k4n = (count + 1023) >> 10;  // groups of 4 mask words in use (1024 positions each)
        const int ncol = k4n * 8;
        for (int task = warp; task < 3 * ncol; task += BP_WARPS) {
          const int tab = task / ncol, col = task - tab * ncol;
          if (tab == 0) prefix_column<false, (BP_NROW + 31) / 32>(&S.px[0][0], BP_NROW, col, lane);
          else if (tab == 1) prefix_column<false, (BP_NROW + 31) / 32>(&S.py[0][0], BP_NROW, col, lane);
          else prefix_column<true, (BP_MAXQ + 1 + 31) / 32>(&S.al[0][0], nq + 1, col, lane);
        }
      }
      __syncthreads();
      // ---- F: four queries per warp, eight lanes each ----
      for (int task = warp; task * 4 < nq; task += BP_WARPS) {
        const int q = task * 4 + grp;
        const int qc = min(q, nq - 1);
        const bool act = q < nq && S.s_ok[qc] != 0u;
        const uint32_t xy = S.s_xy[qc];
        // lanes without a query walk the tables from the tile corner with an empty mask
        const int xi = act ? (int)(xy & 0xffffu) : X0, yi = act ? (int)(xy >> 16) : Y0;
        uint4 T4[BP_LW / 4], P4[BP_LW / 4];
#pragma unroll
        for (int k4 = 0; k4 < BP_LW / 4; k4++) {
          T4[k4] = reinterpret_cast<const uint4 *>(&S.al[qc][0])[k4 * 8 + li];
          if (!act) T4[k4] = make_uint4(0u, 0u, 0u, 0u);
          P4[k4] = make_uint4(0u, 0u, 0u, 0u);
        }
        double Sl = 0.0, Sx = 0.0, Sy = 0.0;
        int Sn = 0;
        float best = 0.f, mean[FARMS_NSCALES];
        int ncum[FARMS_NSCALES];
        int bk = -1, bn = 0;
        double bx = 0.0, by = 0.0;
        // byte address of this lane's payload column; the dummy entry BP_CAP holds (0, 0)
        const char *pay0 = reinterpret_cast<const char *>(&S.pxy[li]);
        const char *payz = reinterpret_cast<const char *>(&S.pxy[BP_CAP]);
#pragma unroll
        for (int k = 0; k < FARMS_NSCALES; k++) {
          const int s = k * FARMS_WINDOW_JUMP;
          const int xa = max(xi - s, 0) - R.rx0, xb = min(xi + s, W - 1) - R.rx0 + 1;   // src/vFlow.cpp:998
          const int ya = max(yi - s, 0) - R.ry0, yb = max(min(yi + s, W - 1) - R.ry0 + 1, 0);   // :1000 (sic)
          const uint4 *rxa = reinterpret_cast<const uint4 *>(&S.px[xa][0]) + li;
          const uint4 *rxb = reinterpret_cast<const uint4 *>(&S.px[xb][0]) + li;
          const uint4 *rya = reinterpret_cast<const uint4 *>(&S.py[ya][0]) + li;
          const uint4 *ryb = reinterpret_cast<const uint4 *>(&S.py[yb][0]) + li;
          float sl = 0.f, sx = 0.f, sy = 0.f;
          int cn = 0;
#pragma unroll
          for (int k4 = 0; k4 < BP_LW / 4; k4++) {
            const uint4 sq = and4(andn4(rxb[k4 * 8], rxa[k4 * 8]), andn4(ryb[k4 * 8], rya[k4 * 8]));
            const uint4 m4 = and4(andn4(sq, P4[k4]), T4[k4]);
            P4[k4] = sq;
            const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
              uint32_t mm = mw[j];
              cn += __popc(mm);
              // staged position of bit b of this word: ((k4*4+j)*32 + b)*8 + li; 8 bytes of payload each
              const char *wbase = pay0 + (k4 * 4 + j) * 256 * 8;
              while (mm) {  // two contributors per trip (independent loads); a missing second one reads (0, 0)
                const float2 c0 = *reinterpret_cast<const float2 *>(wbase + ((__ffs(mm) - 1) << 6));
                mm &= mm - 1;
                const float2 c1 = *reinterpret_cast<const float2 *>(mm ? wbase + ((__ffs(mm) - 1) << 6) : payz);
                mm &= mm - 1;
                float l0, l1;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(c0.x * c0.x + c0.y * c0.y));
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(c1.x * c1.x + c1.y * c1.y));
                sl += l0;
                sx += c0.x;
                sy += c0.y;
                sl += l1;
                sx += c1.x;
                sy += c1.y;
              }
            }
          }
          // ring totals over the eight lanes, in FP64
          double dl = (double)sl, dx = (double)sx, dy = (double)sy;
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            dl += __shfl_xor_sync(0xffffffffu, dl, o);
            dx += __shfl_xor_sync(0xffffffffu, dx, o);
            dy += __shfl_xor_sync(0xffffffffu, dy, o);
            cn += __shfl_xor_sync(0xffffffffu, cn, o);
          }
          // nested-square sums are prefix sums over rings (src/vFlow.cpp:1023-1036); an empty ring adds exactly 0
          Sl += dl;
          Sx += dx;
          Sy += dy;
          Sn += cn;
          const float mk = Sn > 0 ? __fdiv_rn((float)Sl, (float)Sn) : 0.f;
          mean[k] = mk;
          ncum[k] = Sn;
          if (mk > best) {  // strict '>' from 0: first maximum (:1047-1059)
            best = mk;
            bk = k;
            bn = Sn;
            bx = Sx;
            by = Sy;
          }
        }
        // is any other scale (with a different contributor set) within the FP32 noise of the winner?
        bool rival = false;
#pragma unroll
        for (int k = 0; k < FARMS_NSCALES; k++) rival |= ncum[k] != bn && fabsf(mean[k] - best) <= TK_TIE_TOL * best;
        bool safe = act && bk >= 0 && !rival && best > 1e-30f && best < 1e30f;
        // mean vector much shorter than the mean length: the FP32 sums cancelled, let the exact path do it
        const double bl = (double)best * (double)bn;
        safe = safe && (bx * bx + by * by) > 1e-4 * bl * bl;
        if (li == 0 && act) ncand += (unsigned long long)Sn;
        if (li == 0 && safe) {
          // k_pool_finish divides by the count and takes sqrt / atan2 (src/vFlow.cpp:365-366)
          const int out_index = (int)S.s_ii[qc] - A.h;
          A.global_r[out_index] = bx;
          A.global_theta[out_index] = by;
          A.fin[out_index] = (uint32_t)bn | ((uint32_t)bk << 16);
          A.done[S.s_pos[qc]] = 1;
          npooled++;
        }
      }
      d += nd;
    }
  }
  if (ncand) atomicAdd(A.cand_count, ncand);
  if (npooled) {
    atomicAdd(A.path_count, (unsigned long long)npooled);
    atomicAdd(A.batch_words + 2, npooled);
  }
}

// Second half of the fast path's output: mean vector = sums / count, then length and angle (src/vFlow.cpp:365-366).
__global__ void k_pool_finish(const uint32_t *__restrict__ fin, size_t n, double *__restrict__ gr,
                              double *__restrict__ gth, uint8_t *__restrict__ scale) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t v = fin[e];
  if (!v) return;
  const double wn = (double)(v & 0xffffu);
  const double bvx = gr[e] / wn, bvy = gth[e] / wn;
  gr[e] = __dsqrt_rn(__dadd_rn(__dmul_rn(bvy, bvy), __dmul_rn(bvx, bvx)));
  gth[e] = atan2(bvy, bvx);
  scale[e] = (uint8_t)((v >> 16) * FARMS_WINDOW_JUMP);
}

void launch_bits(const PoolArgs &A0, int nslabs, int num_sms, cudaStream_t s) {
  PoolArgs A = A0;
  cudaFuncSetAttribute(k_pool_bits, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BitsSmem));
  cudaFuncSetAttribute(k_pool_bits, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  const int otx = (A.g.W + OT - 1) >> OT_SHIFT, oty = (A.g.H + OT - 1) >> OT_SHIFT;
  const int nseg = (nslabs + TK_SEG - 1) / TK_SEG;
  const long long items = (long long)otx * oty * nseg;
  unsigned grid = (unsigned)std::min<long long>(items, 2ll * num_sms);
  k_pool_bits<<<grid, BP_THREADS, sizeof(BitsSmem), s>>>(A, otx, oty, nseg);
}

inline unsigned nb(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void launch_cell_keys(const uint16_t *ex, const uint16_t *ey, const uint32_t *em, const uint32_t *excl,
                      const double *len, size_t m, PoolGeom g, uint32_t ncells, uint32_t *keys, uint32_t *idx,
                      uint32_t *slab_ids, uint32_t *slab_first, cudaStream_t s) {
  if (m) k_cell_keys<<<nb(m, 256), 256, 0, s>>>(ex, ey, em, excl, len, m, g, ncells, keys, idx, slab_ids, slab_first);
}

void launch_build_records(const uint32_t *skeys, const uint32_t *sidx, size_t m, const uint16_t *ex,
                          const uint16_t *ey, const uint32_t *et, const int32_t *nextp, const double *len,
                          const double *lcx, const double *lcy, int monotone, uint4 *rec, double *pay,
                          uint32_t *cell_start, uint32_t ncells, uint32_t h, unsigned int *n_targets,
                          const uint32_t *time_table, uint32_t tt_base, uint32_t tt_size, cudaStream_t s) {
  if (m) k_build_records<<<nb(m, 256), 256, 0, s>>>(skeys, sidx, m, ex, ey, et, nextp, len, lcx, lcy, monotone, rec,
                                                   pay, cell_start, ncells, h, n_targets, time_table, tt_base, tt_size);
}

// time_table[u] = first event index whose timestamp is >= tt_base + u (sorted timestamps, integer microseconds), m
// past the last event: turns "first event at least 500 us younger" (the age bound of src/vFlow.cpp:1002 folded into
// the index interval of a pooling record) into one look-up instead of a 20-step binary search per flow event.
void launch_time_table(const uint32_t *et, size_t m, uint32_t tt_base, uint32_t *time_table, uint32_t tt_size,
                       cudaStream_t s) {
  if (!m || !tt_size) return;
  k_fill_u32<<<nb(tt_size, 256), 256, 0, s>>>(time_table, tt_size, (uint32_t)m);
  k_time_table<<<nb(m, 256), 256, 0, s>>>(et, m, tt_base, time_table, tt_size);
}

int pool_tile_smem_bytes() { return (int)sizeof(BitsSmem); }

// Launches the fast path (when `fast` is set) and then the general, exact path for whatever is left.
// work_counter: two zeroed words.  done: m zeroed bytes.
size_t pool_item_words(int W, int H, int nslabs) {
  const size_t otx = (size_t)(W + OT - 1) >> OT_SHIFT, oty = (size_t)(H + OT - 1) >> OT_SHIFT;
  return otx * oty * (size_t)((nslabs + TK_SEG - 1) / TK_SEG);
}

int launch_pooling(const uint4 *rec, const double *pay, const uint32_t *cell_start, const uint32_t *slab_ids,
                   const uint32_t *slab_first, uint32_t *fin, uint32_t *item_ovf, uint8_t *done, size_t m, uint32_t ncells, int h,
                   const double *ev_len, const double *ev_lcx, const double *ev_lcy, int nslabs, PoolGeom g, int fast,
                   double flow_per_slab, double *global_r, double *global_theta, uint8_t *scale,
                   unsigned int *work_counter, unsigned long long *cand_count, int num_sms, cudaStream_t s,
                   unsigned *kernels_used, const uint8_t *own_ok) {
  if (!m) return 0;
  int launches = 0;
  PoolArgs A;
  A.own_ok = own_ok;
  A.batch_words = work_counter + 3;
  A.path_count = cand_count + 1;
  A.rec = rec; A.pay = pay; A.cell_start = cell_start; A.slab_ids = slab_ids; A.done = done;
  A.slab_first = slab_first; A.fin = fin; A.item_ovf = item_ovf;
  A.ev_len = ev_len; A.ev_lcx = ev_lcx; A.ev_lcy = ev_lcy;
  A.m = m; A.ncells = ncells; A.h = h; A.nslabs = nslabs; A.g = g;
  A.global_r = global_r; A.global_theta = global_theta; A.scale = scale;
  A.cand_count = cand_count;
  if (fast && g.tile_shift == 4) {
    A.work_counter = work_counter;
    if (fast == 2) {
      launch_bits(A, nslabs, num_sms, s);  // bit-table variant: 8 warps, 2 CTAs per SM, ~110 KB each
      if (kernels_used) *kernels_used |= FARMS_POOLK_BITS;
    } else if (fast == 3) {
      launch_tile<16, 768, 4, 1, false>(A, nslabs, num_sms, s);  // 16 warps, 1 CTA per SM, ~222 KB
      if (kernels_used) *kernels_used |= FARMS_POOLK_TILE_ONE_CTA;
    } else if (fast >= 5 && fast <= 7) {
      // k_pool_tile on 16-byte packed records: 2 slabs per round with 640-record slots, 3 with 512, or 4 with 416
      const double per_region = flow_per_slab * 17424.0 / ((double)g.W * (double)g.H);
      if (per_region < 200.0) launch_tile16<8, 416, 4, 2, false>(A, nslabs, num_sms, s);
      else if (fast == 5) launch_tile16<8, 640, 2, 2, false>(A, nslabs, num_sms, s);
      else if (fast == 7) launch_tile16<8, 512, 3, 2, false>(A, nslabs, num_sms, s);
      else launch_tile16<8, 416, 4, 2, false>(A, nslabs, num_sms, s);
      if (kernels_used) *kernels_used |= per_region < 200.0 ? FARMS_POOLK_TILE16_SPARSE : FARMS_POOLK_TILE16_DENSE;
      A.work_counter = work_counter + 2;
      launch_tile16<16, 960, 4, 1, true>(A, nslabs, num_sms, s);
      if (kernels_used) *kernels_used |= FARMS_POOLK_TILE16_SECOND;
      launches++;
    } else if (fast == 8) {
      // variant 7 with column-culled trips (XCULL)
      const double per_region = flow_per_slab * 17424.0 / ((double)g.W * (double)g.H);
      if (per_region < 200.0) launch_tile16<8, 416, 4, 2, false, true>(A, nslabs, num_sms, s);
      else launch_tile16<8, 512, 3, 2, false, true>(A, nslabs, num_sms, s);
      if (kernels_used)
        *kernels_used |= (per_region < 200.0 ? FARMS_POOLK_TILE16_SPARSE : FARMS_POOLK_TILE16_DENSE) | FARMS_POOLK_TILE16_XCULL;
      A.work_counter = work_counter + 2;
      launch_tile16<16, 960, 4, 1, true, true>(A, nslabs, num_sms, s);
      if (kernels_used) *kernels_used |= FARMS_POOLK_TILE16_SECOND;
      launches++;
    } else if (fast == 4) {
      // two-phase kernel (k_pool_warp), same dense / sparse split and flagged second pass
      const double per_region = flow_per_slab * 17424.0 / ((double)g.W * (double)g.H);
      if (per_region < 200.0) {
        launch_warp<8, 352, 4, 2, false>(A, nslabs, num_sms, s);
        if (kernels_used) *kernels_used |= FARMS_POOLK_WARP_SPARSE;
      } else {
        launch_warp<8, 480, 2, 2, false>(A, nslabs, num_sms, s);
        if (kernels_used) *kernels_used |= FARMS_POOLK_WARP_DENSE;
      }
      A.work_counter = work_counter + 2;
      launch_warp<16, 768, 4, 1, true>(A, nslabs, num_sms, s);
      if (kernels_used) *kernels_used |= FARMS_POOLK_WARP_SECOND;
      launches++;
    } else {
      // flow events a slab holds inside one (32+100)^2 region, from the batch average
      const double per_region = flow_per_slab * 17424.0 / ((double)g.W * (double)g.H);
      if (per_region < 200.0) {  // thin slabs: 4 per round keep the round's task list full (320-record slots)
        launch_tile<8, 320, 4, 2, false>(A, nslabs, num_sms, s);
        if (kernels_used) *kernels_used |= FARMS_POOLK_TILE_SPARSE;
      } else {
        launch_tile<8, 512, 2, 2, false>(A, nslabs, num_sms, s);  // 8 warps, 2 CTAs per SM, ~112 KB each
        if (kernels_used) *kernels_used |= FARMS_POOLK_TILE_DENSE;
      }
      // rounds whose staging overflowed those slots (locally dense scenes) get a second chance with 768-record
      // slots before the general kernel takes what is left
      A.work_counter = work_counter + 2;
      launch_tile<16, 768, 4, 1, true>(A, nslabs, num_sms, s);
      if (kernels_used) *kernels_used |= FARMS_POOLK_TILE_SECOND;
      launches++;
    }
    const size_t nout = m - (size_t)h;
    // the fast kernels leave sums and counts; this pass turns them into globalR / globalTheta / scale
    if (nout) k_pool_finish<<<nb(nout, 256), 256, 0, s>>>(fin, nout, global_r, global_theta, scale);
    launches += 2;
  }
  A.work_counter = work_counter + 1;
  // persistent warps pulling 32-slot groups from a global counter: grid = SMs x resident CTAs
  unsigned grid = (unsigned)num_sms * 5u;
  unsigned need = nb(m, 32 * PW);
  if (grid > need) grid = need;
  k_pool_any<<<grid, PW * 32, 0, s>>>(A);
  if (kernels_used) *kernels_used |= FARMS_POOLK_ANY;
  launches++;
  return launches;
}

unsigned int farms_chk_pooling(cudaStream_t s) {
#ifdef FARMS_CHECKED
  unsigned int v[2] = {0, 0}, z[2] = {0, 0};
  cudaMemcpyFromSymbolAsync(v, g_farms_chk, sizeof v, 0, cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  if (v[0]) cudaMemcpyToSymbolAsync(g_farms_chk, z, sizeof z, 0, cudaMemcpyHostToDevice, s);
  return v[0];
#else
  (void)s;
  return 0;
#endif
}
