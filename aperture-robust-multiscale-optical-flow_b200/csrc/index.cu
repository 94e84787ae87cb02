// index.cu -- K1 ingest and the link half of K2 (per-pixel event-history index).
//
// K1 replaces the per-event prologue of the reference loop (src/vFlow.cpp:238-247): rebase the timestamp
// by t0 into an unsigned 32-bit value, check the pixel against the sensor, and derive the flat pixel key
// x*H + y (include/EventMatrix.h:32-34).  After the stable sort by pixel (sort.cu) the link kernel turns
// each pixel's run into prev/next pointers, which is what lets every event find "the latest event at
// pixel q with index <= i" without replaying the stream (planefit.cu: sae_lookup).
#include "farms_dev.cuh"

namespace {

FARMS_CHK_DECL

__global__ void k_ingest(const uint16_t *__restrict__ x, const uint16_t *__restrict__ y,
                         const uint64_t *__restrict__ t, uint64_t t0, size_t n, int W, int H,
                         uint16_t *__restrict__ ex, uint16_t *__restrict__ ey, uint32_t *__restrict__ et,
                         uint32_t *__restrict__ pix, uint32_t *__restrict__ idx, uint32_t idx_base, int ghost,
                         int *__restrict__ err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t xx = x[i], yy = y[i];
  if (xx >= (uint32_t)W || yy >= (uint32_t)H) {
    atomicOr(err, 1);
    xx = 0;
    yy = 0;
  }
  ex[i] = (uint16_t)xx;
  ey[i] = (uint16_t)yy;
  et[i] = (uint32_t)(t[i] - t0);  // unsigned wrap like `time_ = time_ - t0` (src/vFlow.cpp:241)
  // the ghost event keeps its coordinates (for the output row) but hashes to the dummy pixel W*H
  pix[i] = (long long)i == (long long)ghost ? (uint32_t)W * (uint32_t)H : xx * (uint32_t)H + yy;
  idx[i] = idx_base + (uint32_t)i;
}

__global__ void k_halo_keys(const uint16_t *__restrict__ ex, const uint16_t *__restrict__ ey, size_t h, int H,
                            uint32_t *__restrict__ pix, uint32_t *__restrict__ idx) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h) return;
  pix[i] = (uint32_t)ex[i] * (uint32_t)H + ey[i];
  idx[i] = (uint32_t)i;
}

// skeys/svals: events sorted by pixel, stream order inside a pixel.  One thread per sorted slot.
__global__ void k_links(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ svals,
                        const uint32_t *__restrict__ et, const uint2 *__restrict__ sae, size_t m,
                        int2 *__restrict__ prevp, int32_t *__restrict__ nextp) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  const uint32_t key = skeys[s];
  const uint32_t j = svals[s];
  if (!FARMS_CHK((size_t)j < m && (s == 0 || skeys[s - 1] <= key), 301)) return;  // sorted by pixel, indices in range
  int2 pp;
  if (s > 0 && skeys[s - 1] == key) {
    uint32_t pj = svals[s - 1];
    pp.x = (int)pj;
    pp.y = (int)et[pj];
  } else {
    uint2 c = sae[key];  // state left by earlier batches: index code is SAE_OLD or SAE_NEVER
    pp.x = (int)c.y;
    pp.y = (int)c.x;
  }
  prevp[j] = pp;
  nextp[j] = (s + 1 < m && skeys[s + 1] == key) ? (int32_t)svals[s + 1] : NEXT_NONE;
}

// Serial semantics (src/vFlow.cpp:465-826).  lastEventTime[x][y] is written after pooling (:790), so when event i
// is pooled its own pixel still holds the time of the previous event there: own_ok[i] = |t_i - T_prev| < 500 with
// T_prev = 0 for a never-hit pixel, or the RAW first timestamp for the pixel of the stream's first event (:558).
// The first event itself (the ghost) gets an all-zero row.
__global__ void k_serial_fix(const int2 *__restrict__ prevp, const uint32_t *__restrict__ et,
                             const uint32_t *__restrict__ pix, size_t h, size_t m, int ghost_index, uint32_t ghost_pix,
                             uint64_t ghost_raw_t, int ghost_prev_pending, FitOut fo, uint8_t *__restrict__ own_ok,
                             uint32_t *__restrict__ ghost_consumed) {
  const size_t i = h + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int2 pp = prevp[i];
  double tprev = pp.x == SAE_NEVER ? 0.0 : (double)(uint32_t)pp.y;
  if (ghost_prev_pending && pp.x == SAE_NEVER && pix[i] == ghost_pix) {
    tprev = (double)ghost_raw_t;  // lastEventTime[x][y] = time_ before the rebase (:558)
    *ghost_consumed = 1u;         // (a never-hit pixel has one first event: a single writer)
  }
  own_ok[i] = fabs((double)et[i] - tprev) < (double)FARMS_KILL_OLD_FLOW_TIME ? 1 : 0;
  if ((long long)i == (long long)ghost_index) {
    fo.vx[i] = 0.0; fo.vy[i] = 0.0; fo.len[i] = 0.0; fo.theta[i] = 0.0; fo.lcx[i] = 0.0; fo.lcy[i] = 0.0;
    fo.valid[i] = 0; fo.best_window[i] = -1; fo.inliers[i] = 0;
    if (fo.det) fo.det[i] = __longlong_as_double(0x7ff8000000000000ll);
  }
}

// flags[j] = 1 where a new time slab starts; *nonmono != 0 if any timestamp runs backwards; *regress = the largest
// step back (running maximum - own time, us) among the NEW events j >= h of the batch.
__global__ void k_slab_flags(const uint32_t *__restrict__ em, const uint32_t *__restrict__ et, size_t m, size_t h,
                             int slab_shift, uint32_t *__restrict__ flags, uint32_t *__restrict__ nonmono,
                             uint32_t *__restrict__ regress) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t back = 0, back_new = 0;
  if (j < m) {
    flags[j] = (j > 0 && (em[j] >> slab_shift) != (em[j - 1] >> slab_shift)) ? 1u : 0u;
    back = em[j] - et[j];
    back_new = j >= h ? back : 0u;
  }
  if (__any_sync(0xffffffffu, back != 0)) {  // (never taken on a sorted stream)
    back_new = __reduce_max_sync(0xffffffffu, back_new);
    if ((threadIdx.x & 31) == 0) {
      atomicOr(nonmono, 1u);
      if (back_new) atomicMax(regress, back_new);
    }
  }
}

__global__ void k_slice_surface(const uint16_t *__restrict__ x, const uint16_t *__restrict__ y,
                                const uint64_t *__restrict__ t, size_t n, uint32_t index_base, uint64_t t0, int W, int H,
                                unsigned long long *__restrict__ packed, int *__restrict__ err) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  // Walk the slice backwards: blocks are scheduled roughly in order, so the latest events arrive first and
  // almost every earlier event of the same pixel sees a larger value already and skips its atomic.
  const size_t i = n - 1 - k;
  const uint32_t tr = (uint32_t)(t[i] - t0);
  if (x[i] >= (uint32_t)W || y[i] >= (uint32_t)H) {  // same rule as k_ingest: FARMS_ERR_RANGE, never an OOB write
    atomicOr(err, 1);
    return;
  }
  unsigned long long *cell = &packed[(size_t)x[i] * H + y[i]];
  // later index wins: index in the high word
  const unsigned long long mine = ((unsigned long long)(index_base + (uint32_t)i + 1u) << 32) | tr;
  if (*(volatile unsigned long long *)cell < mine) atomicMax(cell, mine);
}

__global__ void k_unpack_surface(const unsigned long long *__restrict__ packed, size_t npx,
                                 uint32_t *__restrict__ last_t, uint8_t *__restrict__ hit) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npx) return;
  unsigned long long v = packed[q];
  last_t[q] = (uint32_t)v;
  hit[q] = v != 0ull;
}

__global__ void k_pack4(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c,
                        const double *__restrict__ d, size_t n, float4 *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4((float)a[i], (float)b[i], (float)c[i], (float)d[i]);
}

inline unsigned nb(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void launch_pack4(const double *a, const double *b, const double *c, const double *d, size_t n, float4 *out,
                  cudaStream_t s) {
  if (n) k_pack4<<<nb(n, 256), 256, 0, s>>>(a, b, c, d, n, out);
}

void launch_ingest(const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t t0, size_t n, int W, int H,
                   uint16_t *ex, uint16_t *ey, uint32_t *et, uint32_t *pix, uint32_t *idx, uint32_t idx_base,
                   int ghost, int *err_flag, cudaStream_t s) {
  if (n) k_ingest<<<nb(n, 256), 256, 0, s>>>(x, y, t, t0, n, W, H, ex, ey, et, pix, idx, idx_base, ghost, err_flag);
}

void launch_serial_fix(const int2 *prevp, const uint32_t *et, const uint32_t *pix, size_t h, size_t m, int ghost_index,
                       uint32_t ghost_pix, uint64_t ghost_raw_t, int ghost_prev_pending, FitOut fo, uint8_t *own_ok,
                       uint32_t *ghost_consumed, cudaStream_t s) {
  if (m > h)
    k_serial_fix<<<nb(m - h, 256), 256, 0, s>>>(prevp, et, pix, h, m, ghost_index, ghost_pix, ghost_raw_t,
                                               ghost_prev_pending, fo, own_ok, ghost_consumed);
}
void launch_halo_keys(const uint16_t *ex, const uint16_t *ey, size_t h, int H, uint32_t *pix, uint32_t *idx,
                      cudaStream_t s) {
  if (h) k_halo_keys<<<nb(h, 256), 256, 0, s>>>(ex, ey, h, H, pix, idx);
}
void launch_links(const uint32_t *skeys, const uint32_t *svals, const uint32_t *et, const uint2 *sae, size_t m,
                  int2 *prevp, int32_t *nextp, cudaStream_t s) {
  if (m) k_links<<<nb(m, 256), 256, 0, s>>>(skeys, svals, et, sae, m, prevp, nextp);
}
void launch_slab_flags(const uint32_t *em, const uint32_t *et, size_t m, size_t h, int slab_shift, uint32_t *flags,
                       uint32_t *nonmono, uint32_t *regress, cudaStream_t s) {
  if (m) k_slab_flags<<<nb(m, 256), 256, 0, s>>>(em, et, m, h, slab_shift, flags, nonmono, regress);
}
void launch_slice_surface(const uint16_t *x, const uint16_t *y, const uint64_t *t, size_t n, uint32_t index_base,
                          uint64_t t0, int W, int H, unsigned long long *packed, int *err_flag, cudaStream_t s) {
  if (n) k_slice_surface<<<nb(n, 256), 256, 0, s>>>(x, y, t, n, index_base, t0, W, H, packed, err_flag);
}
void launch_unpack_surface(const unsigned long long *packed, size_t npx, uint32_t *last_t, uint8_t *hit,
                           cudaStream_t s) {
  k_unpack_surface<<<nb(npx, 256), 256, 0, s>>>(packed, npx, last_t, hit);
}

unsigned int farms_chk_index(cudaStream_t s) {
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
