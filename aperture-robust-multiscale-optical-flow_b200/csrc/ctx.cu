// ctx.cu -- host side of the C ABI (include/farms_b200.h): context, device memory, batch pipeline.
//
// Replaces vFlowManager's constructor and the runFileCopy event loop (src/vFlow.cpp:22-108, 223-414).
// A submit is cut into internal batches; each batch is [halo | new] where the halo is the tail of the
// stream already processed (the events young enough to still matter for pooling, src/vFlow.cpp:1002).
// Per batch:  K1 ingest -> K2 sort by pixel + links -> K3 (SAE advance, plane fit) per chunk ->
//             K4a bin by (time slab, tile) -> K4b pooling -> copy results out -> keep the new tail.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "farms_ctx.cuh"

static int vfail(farms_ctx *c, int code, const char *fmt, va_list ap) {
  char buf[512];
  vsnprintf(buf, sizeof buf, fmt, ap);
  if (c) c->err = buf;
  return code;
}

// records the message for farms_last_error and returns the status code (shared with comm.cu)
int farms_fail(farms_ctx *c, int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  code = vfail(c, code, fmt, ap);
  va_end(ap);
  return code;
}

namespace {

#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return farms_fail(c, e_ == cudaErrorMemoryAllocation ? FARMS_ERR_NOMEM : FARMS_ERR_CUDA, "%s: %s (%s:%d)", \
                  #call, cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
  } while (0)

template <class T>
int dalloc(farms_ctx *c, WorkSet &w, T **p, size_t count) {
  void *q = nullptr;
  CU(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  *p = (T *)q;
  w.owned.push_back(q);
  return 0;
}

void free_owned(WorkSet &w) {
  for (void *p : w.owned) cudaFree(p);
  w.owned.clear();
  w.cap = 0;
}

int ensure(farms_ctx *c, DevBuf &b, size_t bytes) {
  if (b.bytes >= bytes) return 0;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  size_t want = bytes + bytes / 4 + 4096;
  CU(cudaMalloc(&b.p, want));
  b.bytes = want;
  return 0;
}

// (re)allocate the per-event working arrays of one set for `cap` events
int alloc_working(farms_ctx *c, WorkSet &w, size_t cap) {
  if (cap <= w.cap) return 0;
  CU(cudaDeviceSynchronize());
  free_owned(w);
  int rc = 0;
#define A(ptr, n) if ((rc = dalloc(c, w, &w.ptr, (n)))) return rc;
  A(ex, cap) A(ey, cap) A(et, cap) A(em, cap) A(keyA, cap) A(valA, cap) A(keyB, cap) A(valB, cap)
  A(pixkeep, cap) A(flags, cap) A(slab_ids, cap + 1) A(slab_first, cap + 1) A(fin, cap) A(prevp, cap) A(nextp, cap)
  A(vx, cap) A(vy, cap) A(len, cap) A(theta, cap) A(lcx, cap) A(lcy, cap) A(gr, cap) A(gth, cap)
  A(pay, 3 * cap) A(done, cap) A(valid, cap) A(scale, cap) A(bw, cap) A(inl, cap) A(rec, cap)
  if (c->serial) { A(own, cap) } else w.own = nullptr;
  if (c->cfg.flags & FARMS_FLAG_DEBUG_DET) { A(det, cap) } else w.det = nullptr;
#undef A
  if ((rc = ensure(c, c->sort_temp, radix_sort_temp_bytes(cap)))) return rc;
  if ((rc = ensure(c, c->scan_temp, scan_temp_bytes(cap)))) return rc;
  w.cap = cap;
  return 0;
}

int bits_for(uint64_t n) {  // bits needed to represent values in [0, n)
  int b = 1;
  while (b < 32 && (1ull << b) < n) b++;
  return b;
}

__global__ void k_tail_start(const uint32_t *__restrict__ em, uint32_t m, uint32_t window, uint32_t *out) {
  // first j with em[j] + window > em[m-1]  (em is non-decreasing)
  const uint32_t last = em[m - 1];
  uint32_t lo = 0, hi = m;  // answer in [lo, hi)
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo) / 2;
    if ((uint64_t)em[mid] + window > last) hi = mid; else lo = mid + 1;
  }
  out[0] = lo;
  out[1] = last;
}

__global__ void k_nslabs(const uint32_t *__restrict__ em, const uint32_t *__restrict__ excl, uint32_t m, int slab_shift,
                         uint32_t *out) {
  const bool first = m == 1 || (em[m - 2] >> slab_shift) != (em[m - 1] >> slab_shift);
  out[0] = excl[m - 1] + ((m > 1 && first) ? 1u : 0u) + 1u;
}

__global__ void k_time_span(const uint32_t *__restrict__ em, uint32_t m, uint32_t *out) {
  out[0] = em[0];
  out[1] = em[m - 1];
}

// first event of a serial-semantics stream: flat pixel and raw timestamp, straight into pinned host memory
__global__ void k_ghost_info(const uint16_t *__restrict__ x, const uint16_t *__restrict__ y, const uint64_t *__restrict__ t,
                             int H, uint32_t *host_dst) {
  host_dst[0] = (uint32_t)x[0] * (uint32_t)H + y[0];
  host_dst[1] = 0;
  host_dst[2] = (uint32_t)t[0];
  host_dst[3] = (uint32_t)(t[0] >> 32);
  __threadfence_system();
}

// Small device->host read-backs (error flag, slab count, tail position) are written straight into pinned host
// memory by a kernel: a cudaMemcpy on the compute stream would queue behind the large result copies of the
// previous batch on the device->host copy engine and stall the kernels.
__global__ void k_publish(uint32_t *__restrict__ host_dst, const uint32_t *__restrict__ src, int nwords) {
  if ((int)threadIdx.x < nwords) host_dst[threadIdx.x] = src[threadIdx.x];
  __threadfence_system();
}

template <class T>
int copy_out(farms_ctx *c, T *dst, const T *src, size_t n, bool to_device, cudaStream_t st) {
  if (!dst || !n) return 0;
  CU(cudaMemcpyAsync(dst, src, n * sizeof(T), to_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  return 0;
}

// stage times of the batch whose events were recorded in set c->stage_pending (all completed by now)
int collect_stage_times(farms_ctx *c, float *stage_ms) {
  if (c->stage_pending < 0) return 0;
  static const int order[] = {EV_H2D, EV_INGEST, EV_INDEX, EV_FIT, EV_BIN, EV_POOL, EV_END};
  float ms = 0;
  for (int k = 1; k < 7; k++) {
    CU(cudaEventElapsedTime(&ms, c->evb[c->stage_pending][order[k - 1]], c->evb[c->stage_pending][order[k]]));
    stage_ms[k] += ms;
  }
  c->stage_pending = -1;
  return 0;
}

// the output columns that exist once the plane fit of a batch is done (everything but globalR / globalTheta / scale)
int copy_fit_columns(farms_ctx *c, WorkSet &w, const farms_out *out, size_t hh, size_t n_out, size_t out_off,
                     bool out_device, cudaStream_t st) {
  int rc;
#define COL(field, src) \
  if ((rc = copy_out(c, out->field ? out->field + out_off : nullptr, (src) + hh, n_out, out_device, st))) return rc;
  COL(t_rel, w.et) COL(vx, w.vx) COL(vy, w.vy) COL(local_r, w.len) COL(local_theta, w.theta) COL(valid, w.valid)
  COL(best_window, w.bw) COL(inliers, w.inl)
#undef COL
  if (out->det && w.det)
    if ((rc = copy_out(c, out->det + out_off, w.det + hh, n_out, out_device, st))) return rc;
  return 0;
}

// One internal batch: n new events already on the device at (dx, dy, dt).  The first `skip` of them are history
// only (a time slice's causal halo): they are fitted (their flow is state) but neither pooled nor returned.
// Results go out on `out_stream` (the compute stream itself, or the device->host stream of the host path).
int run_batch(farms_ctx *c, int set, const uint16_t *dx, const uint16_t *dy, const uint64_t *dt, size_t n, size_t skip,
              const farms_out *out, size_t out_off, bool out_device, cudaStream_t out_stream, float *stage_ms,
              const FarmsBatchHook *hook) {
  cudaStream_t s = c->stream;
  WorkSet &w = c->ws[set];
  const size_t h = c->halo, m = h + n;
  const size_t hh = h + skip, n_out = n - skip;  // first batch-local index with outputs, and how many
  int rc;
  if (m > w.cap && (rc = alloc_working(c, w, m + m / 8 + 1024))) return rc;
  // this set's previous results must have left the device before it is overwritten
  if (c->d2h_pending[set]) {
    CU(cudaStreamWaitEvent(s, c->ev_d2h[set], 0));
    c->d2h_pending[set] = false;
  }
  uint64_t *L = &c->tm.kernel_launches;

  // ---- halo to the front of the working arrays ----
  if (h) {
    CU(cudaMemcpyAsync(w.ex, c->hx, h * 2, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(w.ey, c->hy, h * 2, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(w.et, c->ht, h * 4, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(w.em, c->hm, h * 4, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(w.len, c->hlen, h * 8, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(w.lcx, c->hlcx, h * 8, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(w.lcy, c->hlcy, h * 8, cudaMemcpyDeviceToDevice, s));
  }
  CU(cudaEventRecord(c->evb[c->ev_set][EV_H2D], s));

  // ---- K1 ingest ----
  CU(cudaMemsetAsync(c->d_err, 0, sizeof(int), s));
  // serial semantics: the very first event of the stream only sets t0 (src/vFlow.cpp:531-558)
  const int ghost = (c->serial && !c->ghost_seen) ? 0 : -1;
  launch_ingest(dx, dy, dt, c->t0, n, c->W, c->H, w.ex + h, w.ey + h, w.et + h, w.keyA + h, w.valA + h,
                (uint32_t)h, ghost, c->d_err, s);
  if (ghost == 0) {
    // its pixel and raw timestamp, for the first later event at that pixel (pinned words 40..43)
    k_ghost_info<<<1, 1, 0, s>>>(dx, dy, dt, c->H, c->h_small + 40);
    *L += 1;
  }
  launch_halo_keys(w.ex, w.ey, h, c->H, w.keyA, w.valA, s);
  inclusive_max_scan_u32(w.et + h, w.em + h, n, c->last_M, c->scan_temp.p, s, L);
  // where the history kept for the next batch starts (events younger than 500 us + slack at the end of this one):
  // known as soon as the running maximum of the timestamps is, and read back at the mid-batch synchronisation
  const uint32_t window = FARMS_KILL_OLD_FLOW_TIME + (c->cfg.reorder_slack_us ? c->cfg.reorder_slack_us : DEFAULT_SLACK_US);
  k_tail_start<<<1, 1, 0, s>>>(w.em, (uint32_t)m, window, c->d_small + 12);
  k_publish<<<1, 32, 0, s>>>(c->h_small + 48, c->d_small + 12, 2);
  *L += 2;
  CU(cudaMemcpyAsync(w.pixkeep, w.keyA, m * 4, cudaMemcpyDeviceToDevice, s));
  *L += 2;
  k_publish<<<1, 32, 0, s>>>(c->h_small + 8, (const uint32_t *)c->d_err, 1);  // read at the mid-batch sync below
  CU(cudaEventRecord(c->evb[c->ev_set][EV_INGEST], s));
  CU(cudaEventRecord(c->ev_ingest[set], s));  // the input staging buffers of this set are free again

  // ---- K2 history index: stable sort by pixel, prev/next links ----
  int which = radix_sort_pairs(w.keyA, w.valA, w.keyB, w.valB, m, bits_for(c->npx + 1), c->sort_temp.p, s, L);
  const uint32_t *skeys = which ? w.keyB : w.keyA, *svals = which ? w.valB : w.valA;
  launch_links(skeys, svals, w.et, c->sae, m, w.prevp, w.nextp, s);
  *L += 1;
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->evb[c->ev_set][EV_INDEX], s));


  // ---- K3 plane fit, chunk by chunk against the chunk-end SAE snapshot ----
  // FIT_WAYS surfaces and streams: chunk k works on surface k mod FIT_WAYS in stream k mod FIT_WAYS, so
  // neighbouring chunks overlap (their kernels are short: launch gaps and tails would otherwise idle the GPU).
  // A surface skipped the chunks since its last turn, so its advance step applies them together with chunk k.
  const size_t FIT_CHUNK = (size_t)c->fit_chunk;
  const size_t scratch_bytes = (plane_fit_scratch_bytes(c->r, FIT_CHUNK) + 255) & ~(size_t)255;
  if ((rc = ensure(c, c->fit_scratch, FIT_WAYS * scratch_bytes))) return rc;
  FitParams fp{c->W, c->H, c->r, c->P, c->min_inl};
  FitOut fo{w.vx, w.vy, w.len, w.theta, w.lcx, w.lcy, w.valid, w.bw, w.inl, w.det};
  for (int q = 0; q < FIT_WAYS - 1; q++)
    CU(cudaMemcpyAsync(c->sae_x[q], c->sae, (c->npx + 1) * sizeof(uint2), cudaMemcpyDeviceToDevice, s));
  CU(cudaEventRecord(c->ev_c0, s));
  for (int q = 0; q < FIT_WAYS - 1; q++) CU(cudaStreamWaitEvent(c->fit_streams[q], c->ev_c0, 0));
  size_t nchunks = 0;
  for (size_t c0 = 0; c0 < m; c0 += FIT_CHUNK, nchunks++) {
    const size_t c1 = std::min(m, c0 + FIT_CHUNK);
    const int way = (int)(nchunks % FIT_WAYS);
    cudaStream_t st = way ? c->fit_streams[way - 1] : s;
    uint2 *surf = way ? c->sae_x[way - 1] : c->sae;
    // first event this surface has not seen yet: the chunk after its previous turn
    const size_t from = nchunks >= (size_t)FIT_WAYS ? c0 - (size_t)(FIT_WAYS - 1) * FIT_CHUNK : 0;
    launch_sae_advance(surf, w.pixkeep, w.et, w.nextp, (int)from, (int)c1, st);
    *L += 1;
    if (c1 > h) {
      *L += launch_plane_fit(surf, w.prevp, w.ex, w.ey, w.et, (int)std::max(c0, h), (int)c1, fp, fo, c->d_counters,
                             (char *)c->fit_scratch.p + (size_t)way * scratch_bytes, st);
    }
  }
  for (int q = 0; q < FIT_WAYS - 1; q++) {
    CU(cudaEventRecord(c->ev_fit[q], c->fit_streams[q]));
    CU(cudaStreamWaitEvent(s, c->ev_fit[q], 0));
  }
  if (nchunks) {  // the surface of the last chunk holds the whole batch: it becomes the persistent one
    const int way = (int)((nchunks - 1) % FIT_WAYS);
    if (way) std::swap(c->sae, c->sae_x[way - 1]);
  }
  launch_sae_finalize(c->sae, w.pixkeep, w.nextp, (int)m, s);
  *L += 1;
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->evb[c->ev_set][EV_FIT], s));

  // ---- the columns the plane fit produced can leave now, under the pooling of this batch ----
  const bool early_copy = out && n_out && out_stream != s && !c->serial;  // (serial: the ghost row is fixed later)
  if (early_copy) {
    CU(cudaEventRecord(c->ev_fitdone[set], s));
    CU(cudaStreamWaitEvent(out_stream, c->ev_fitdone[set], 0));
    if ((rc = copy_fit_columns(c, w, out, hh, n_out, out_off, out_device, out_stream))) return rc;
  }

  // ---- K4a pooling index: dense time slabs x tiles ----
  // Slab length from the density of flow events: the shortest of 128 us, 256 us, ... that puts at least 70 flow
  // events of a (32+100)^2 tile region into a slab (shorter slabs cut fewer useless candidates than their fixed
  // cost per slab and round; a sparse stream would otherwise spend its time on near-empty slabs).
  k_time_span<<<1, 1, 0, s>>>(w.em, (uint32_t)m, c->d_small + 2);
  k_publish<<<1, 32, 0, s>>>(c->h_small + 4, c->d_small + 2, 2);
  k_publish<<<1, 32, 0, s>>>(c->h_small + 2, (const uint32_t *)c->d_counters, 2);
  *L += 3;
  CU(cudaStreamSynchronize(s));
  // (k_ingest clamps an out-of-range event to pixel (0,0), so the kernels above were safe to run)
  if (((int *)c->h_small)[8]) return farms_fail(c, FARMS_ERR_RANGE, "event outside the %dx%d sensor", c->W, c->H);
  const size_t ts = c->h_small[48];
  const uint32_t last_M = c->h_small[49];
  // more than HALO_CAP events inside the pooling window (+ slack): truncating the halo would silently lose
  // contributors for the next batch, so this is an error (timestamps in the wrong unit, or a pathological burst) --
  // raised before the pooling stage, whose cost grows with the number of events per pixel inside the window
  if (m - ts > HALO_CAP)
    return farms_fail(c, FARMS_ERR_STATE, "%zu events within the last %u us exceed the %zu-event history kept across batches",
                m - ts, window, HALO_CAP);
  if ((rc = collect_stage_times(c, stage_ms))) return rc;  // of the previous batch: its events have all completed
  if (c->serial) {
    if (ghost == 0) {
      c->ghost_seen = true;
      c->ghost_prev_pending = true;
      c->ghost_pix = c->h_small[40];
      memcpy(&c->ghost_raw_t, c->h_small + 42, sizeof(uint64_t));
    }
    CU(cudaMemsetAsync(c->d_small + 8, 0, 4, s));
    launch_serial_fix(w.prevp, w.et, w.pixkeep, h, m, ghost == 0 ? (int)h : -1, c->ghost_pix, c->ghost_raw_t,
                      c->ghost_prev_pending ? 1 : 0, fo, w.own, c->d_small + 8, s);
    k_publish<<<1, 32, 0, s>>>(c->h_small + 44, c->d_small + 8, 1);  // read at the end-of-batch sync
    *L += 2;
  }
  // flow events of this batch's new part (the fit ran above); the halo is assumed to have the same density
  const unsigned long long valid_total = *(const unsigned long long *)(c->h_small + 2);
  const double flow_frac = n ? (double)(valid_total - c->valid_seen) / (double)n : 0.0;
  c->valid_seen = valid_total;
  PoolGeom g{};
  g.W = c->W;
  g.H = c->H;
  g.slab_shift = FARMS_SLAB_SHIFT_MIN;
  {
    const double span_us = (double)(c->h_small[5] - c->h_small[4]) + 1.0;
    const double per_region_us = flow_frac * (double)m * (132.0 * 132.0 / (double)c->npx) / span_us;
    const double target = c->cfg.slab_target ? (double)c->cfg.slab_target : 70.0;
    while (g.slab_shift < FARMS_SLAB_SHIFT_MAX && per_region_us * (double)(1u << g.slab_shift) < target) g.slab_shift++;
  }
  CU(cudaMemsetAsync(c->d_small, 0, 12, s));
  launch_slab_flags(w.em, w.et, m, hh, g.slab_shift, w.flags, c->d_small + 1, c->d_small + 2, s);
  exclusive_scan_u32(w.flags, w.flags, m, c->scan_temp.p, s, L);
  k_nslabs<<<1, 1, 0, s>>>(w.em, w.flags, (uint32_t)m, g.slab_shift, c->d_small);
  *L += 2;
  k_publish<<<1, 32, 0, s>>>(c->h_small, c->d_small, 2);
  k_publish<<<1, 32, 0, s>>>(c->h_small + 10, c->d_small + 2, 1);
  *L += 1;
  CU(cudaStreamSynchronize(s));
  const size_t nslabs = c->h_small[0];
  const int monotone = c->h_small[1] == 0;
  // History older than 500 us + slack behind the running maximum was dropped at an earlier batch boundary (or was
  // never there: a time slice's halo).  A new event whose timestamp steps back further than the slack could have
  // had contributors in it, so the result would depend on where the batches were cut: an error, not a silent loss.
  if (c->history_cut && c->h_small[10] > window - FARMS_KILL_OLD_FLOW_TIME)
    return farms_fail(c, FARMS_ERR_STATE,
                "a timestamp runs %u us behind the stream's maximum, more than reorder_slack_us = %u allows across "
                "a batch boundary (raise reorder_slack_us or max_batch)",
                c->h_small[10], window - FARMS_KILL_OLD_FLOW_TIME);
  g.tile_shift = 4;
  for (;;) {
    g.ntx = (c->W + (1 << g.tile_shift) - 1) >> g.tile_shift;
    g.nty = (c->H + (1 << g.tile_shift) - 1) >> g.tile_shift;
    if (nslabs * (size_t)g.ntx * g.nty <= CSR_BUDGET || (g.ntx == 1 && g.nty == 1)) break;
    g.tile_shift++;
  }
  const size_t ncells = nslabs * (size_t)g.ntx * g.nty;
  if (ncells >= (1ull << 31)) return farms_fail(c, FARMS_ERR_NOMEM, "pooling index too large (%zu cells)", ncells);
  if ((rc = ensure(c, c->cell_start, (ncells + 1) * sizeof(uint32_t)))) return rc;
  CU(cudaMemsetAsync(c->cell_start.p, 0, (ncells + 1) * sizeof(uint32_t), s));
  launch_cell_keys(w.ex, w.ey, w.em, w.flags, w.len, m, g, (uint32_t)ncells, w.keyA, w.valA, w.slab_ids, w.slab_first, s);
  which = radix_sort_pairs(w.keyA, w.valA, w.keyB, w.valB, m, bits_for(ncells + 1), c->sort_temp.p, s, L);
  skeys = which ? w.keyB : w.keyA;
  svals = which ? w.valB : w.valA;
  CU(cudaMemsetAsync(c->d_work, 0, 8 * sizeof(unsigned int), s));
  // sorted timestamps: a per-microsecond table "first event at or after time u" replaces the binary search for the
  // 500-us age bound of every pooling record (skipped for batches spanning more than 2^26 us: the search is used)
  const uint32_t tt_base = c->h_small[4];
  const uint64_t tt_span = (uint64_t)c->h_small[5] - c->h_small[4] + FARMS_KILL_OLD_FLOW_TIME + 2;
  const uint32_t *time_table = nullptr;
  if (monotone && tt_span <= (1ull << 26)) {
    if ((rc = ensure(c, c->time_table, tt_span * sizeof(uint32_t)))) return rc;
    launch_time_table(w.et, m, tt_base, (uint32_t *)c->time_table.p, (uint32_t)tt_span, s);
    time_table = (const uint32_t *)c->time_table.p;
    *L += 2;
  }
  launch_build_records(skeys, svals, m, w.ex, w.ey, w.et, w.nextp, w.len, w.lcx, w.lcy, monotone, w.rec,
                       w.pay, (uint32_t *)c->cell_start.p, (uint32_t)ncells, (uint32_t)hh, c->d_work + 4, time_table,
                       tt_base, (uint32_t)tt_span, s);
  *L += 2;
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->evb[c->ev_set][EV_BIN], s));

  // ---- K4b pooling ----
  CU(cudaMemsetAsync(w.gr, 0, n_out * 8, s));
  CU(cudaMemsetAsync(w.gth, 0, n_out * 8, s));
  CU(cudaMemsetAsync(w.scale, 0, n_out, s));
  
  CU(cudaMemsetAsync(w.done, 0, m, s));
  CU(cudaMemsetAsync(w.fin, 0, n_out * sizeof(uint32_t), s));
  {
    const size_t iw = pool_item_words(c->W, c->H, (int)nslabs);
    if ((rc = ensure(c, c->item_ovf, iw * sizeof(uint32_t)))) return rc;
    CU(cudaMemsetAsync(c->item_ovf.p, 0, iw * sizeof(uint32_t), s));
  }
  const int fast = (monotone && !c->serial && !(c->cfg.flags & FARMS_FLAG_GENERIC_POOLING)) ? c->pool_impl : 0;
  *L += launch_pooling(w.rec, w.pay, (const uint32_t *)c->cell_start.p, w.slab_ids, w.slab_first, w.fin, (uint32_t *)c->item_ovf.p, w.done, m, (uint32_t)ncells,
                       (int)hh, w.len, w.lcx, w.lcy, (int)nslabs, g, fast, flow_frac * (double)m / (double)nslabs, w.gr, w.gth, w.scale,
                       c->d_work, c->d_counters + 1, c->num_sms, s, &c->pool_kernels, c->serial ? w.own : nullptr);
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->evb[c->ev_set][EV_POOL], s));

  // ---- results of the new events ----
  if (hook && hook->fn && n_out) {
    FarmsBatchView v{w.gr, w.gth, w.len + hh, w.theta + hh, n_out, out_off, s};
    if ((rc = hook->fn(hook->user, c, &v))) return rc;
  }
  if (out && n_out) {
    if (out_stream != s) {
      CU(cudaEventRecord(c->ev_pool[set], s));
      CU(cudaStreamWaitEvent(out_stream, c->ev_pool[set], 0));
    }
    if (!early_copy && (rc = copy_fit_columns(c, w, out, hh, n_out, out_off, out_device, out_stream))) return rc;
    if ((rc = copy_out(c, out->global_r ? out->global_r + out_off : nullptr, w.gr, n_out, out_device, out_stream))) return rc;
    if ((rc = copy_out(c, out->global_theta ? out->global_theta + out_off : nullptr, w.gth, n_out, out_device, out_stream))) return rc;
    if ((rc = copy_out(c, out->scale ? out->scale + out_off : nullptr, w.scale, n_out, out_device, out_stream))) return rc;
  }
  if (out_stream != s) {  // the set may be reused once these copies are done
    CU(cudaEventRecord(c->ev_d2h[set], out_stream));
    c->d2h_pending[set] = true;
  }

  // ---- new tail -> halo store (no host synchronisation: the host goes on to enqueue the next batch) ----
  CU(cudaEventRecord(c->evb[c->ev_set][EV_END], s));
  c->last_M = last_M;
  const size_t nh = m - ts;
  CU(cudaMemcpyAsync(c->hx, w.ex + ts, nh * 2, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(c->hy, w.ey + ts, nh * 2, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(c->ht, w.et + ts, nh * 4, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(c->hm, w.em + ts, nh * 4, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(c->hlen, w.len + ts, nh * 8, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(c->hlcx, w.lcx + ts, nh * 8, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(c->hlcy, w.lcy + ts, nh * 8, cudaMemcpyDeviceToDevice, s));
  c->halo = nh;
  if (ts > 0) c->history_cut = true;
  c->stage_pending = c->ev_set;  // its stage times are read at the next synchronisation
  c->ev_set ^= 1;
  return 0;
}

}  // namespace

int farms_process_impl(farms_ctx *c, const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t n,
                       const farms_out *out, bool in_device, bool out_device, uint64_t n_skip,
                       const FarmsBatchHook *hook) {
  if (!c) return FARMS_ERR_ARG;
  c->err.clear();
  if (n == 0) return FARMS_OK;
  if (!x || !y || !t) return farms_fail(c, FARMS_ERR_ARG, "null event array");
  if (n_skip > n) return farms_fail(c, FARMS_ERR_ARG, "n_skip exceeds n");
  if (n_skip) c->history_cut = true;  // a time slice: what lies before its halo is not here
  CU(cudaSetDevice(c->cfg.device));
  cudaStream_t s = c->stream;
  if (!c->have_t0) {  // src/vFlow.cpp:194
    if (in_device) {
      CU(cudaMemcpyAsync(c->h_small + 12, t, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      memcpy(&c->t0, c->h_small + 12, sizeof(uint64_t));
    } else {
      c->t0 = t[0];
    }
    c->have_t0 = true;
  }
  const uint64_t maxb = c->cfg.max_batch ? c->cfg.max_batch : DEFAULT_MAX_BATCH;
  farms_timings &tm = c->tm;
  tm = farms_timings{};
  CU(cudaMemsetAsync(c->d_counters, 0, 8 * sizeof(unsigned long long), s));
  c->valid_seen = 0;
  c->pool_kernels = 0;
  c->stage_pending = -1;
  float stage[8] = {0};
  CU(cudaEventRecord(c->ev[EV_START], s));
  const uint64_t nbatch = (n + maxb - 1) / maxb;
  if (!in_device && c->cap_in < std::min<uint64_t>(n, maxb)) {
    const size_t want = (size_t)std::min<uint64_t>(n, maxb);
    CU(cudaDeviceSynchronize());
    for (int k = 0; k < 2; k++) {
      if (c->in_x[k]) { cudaFree(c->in_x[k]); cudaFree(c->in_y[k]); cudaFree(c->in_t[k]); }
      c->in_x[k] = nullptr; c->in_y[k] = nullptr; c->in_t[k] = nullptr;
    }
    c->cap_in = 0;
    for (int k = 0; k < 2; k++) {
      CU(cudaMalloc((void **)&c->in_x[k], want * 2));
      CU(cudaMalloc((void **)&c->in_y[k], want * 2));
      CU(cudaMalloc((void **)&c->in_t[k], want * 8));
    }
    c->cap_in = want;
  }
  // host input: batch k+1 is uploaded while batch k computes; host output: the results of batch k go down while
  // batch k+1 computes (three streams, two working sets)
  const bool two_sets = !in_device || !out_device;
  auto upload = [&](uint64_t k) -> int {
    const int set = (int)(k & 1);
    const uint64_t off = k * maxb;
    const size_t nb = (size_t)std::min<uint64_t>(maxb, n - off);
    if (k >= 2) CU(cudaStreamWaitEvent(c->h2d_stream, c->ev_ingest[set], 0));
    CU(cudaMemcpyAsync(c->in_x[set], x + off, nb * 2, cudaMemcpyHostToDevice, c->h2d_stream));
    CU(cudaMemcpyAsync(c->in_y[set], y + off, nb * 2, cudaMemcpyHostToDevice, c->h2d_stream));
    CU(cudaMemcpyAsync(c->in_t[set], t + off, nb * 8, cudaMemcpyHostToDevice, c->h2d_stream));
    CU(cudaEventRecord(c->ev_h2d[set], c->h2d_stream));
    return 0;
  };
  auto drain = [&]() {
    cudaDeviceSynchronize();
    c->d2h_pending[0] = c->d2h_pending[1] = false;
    c->stage_pending = -1;
  };
  if (!in_device) {
    int rc = upload(0);
    if (rc) return rc;
  }
  for (uint64_t k = 0; k < nbatch; k++) {
    const uint64_t off = k * maxb;
    const size_t nb = (size_t)std::min<uint64_t>(maxb, n - off);
    const int set = two_sets ? (int)(k & 1) : 0;
    const uint16_t *dx = x + off, *dy = y + off;
    const uint64_t *dt = t + off;
    if (!in_device) {
      if (k + 1 < nbatch) {
        int rc = upload(k + 1);
        if (rc) { drain(); return rc; }
      }
      CU(cudaStreamWaitEvent(s, c->ev_h2d[set], 0));
      dx = c->in_x[set];
      dy = c->in_y[set];
      dt = c->in_t[set];
    }
    const size_t skip = (size_t)(off >= n_skip ? 0 : std::min<uint64_t>(n_skip - off, nb));
    const size_t out_off = (size_t)(off >= n_skip ? off - n_skip : 0);
    int rc = run_batch(c, set, dx, dy, dt, nb, skip, out, out_off, out_device, out_device ? s : c->d2h_stream, stage, hook);
    if (rc) {
      drain();
      return rc;
    }
    c->total_events += nb;
  }
  if (!out_device) CU(cudaStreamSynchronize(c->d2h_stream));
  if (!in_device) CU(cudaStreamSynchronize(c->h2d_stream));
  c->d2h_pending[0] = c->d2h_pending[1] = false;
  CU(cudaEventRecord(c->ev[EV_END], s));
  CU(cudaStreamSynchronize(s));
  {
    int rc = collect_stage_times(c, stage);
    if (rc) return rc;
  }
  float total = 0;
  CU(cudaEventElapsedTime(&total, c->ev[EV_START], c->ev[EV_END]));
  CU(cudaMemcpyAsync(c->h_small + 16, c->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  unsigned long long counters[8];
  memcpy(counters, c->h_small + 16, sizeof counters);
  tm.total_ms = total;
  tm.ingest_ms = stage[1];
  tm.index_ms = stage[2];
  tm.fit_ms = stage[3];
  tm.bin_ms = stage[4];
  tm.pool_ms = stage[5];
  tm.events = n;
  tm.valid_events = counters[0];
  tm.pool_candidates = counters[1];
  tm.pool_kernels = c->pool_kernels;
  for (int k = 0; k < 3; k++) tm.pool_events[k] = counters[2 + k];
#ifdef FARMS_CHECKED
  {
    const unsigned int a = farms_chk_index(s), b = farms_chk_planefit(s), d = farms_chk_pooling(s);
    if (a | b | d) return farms_fail(c, FARMS_ERR_STATE, "self-check failed: index %u, plane fit %u, pooling %u (csrc/*.cu FARMS_CHK codes)", a, b, d);
  }
#endif
  return FARMS_OK;
}

extern "C" {

int farms_abi_version(void) { return FARMS_B200_ABI_VERSION; }

int farms_build_is_checked(void) {
#ifdef FARMS_CHECKED
  return 1;
#else
  return 0;
#endif
}

int farms_normalize_filtersize(int fs, int32_t *radius, int32_t *plane_size) {
  if (fs < 5) fs = 3;        // src/vFlow.cpp:33
  if (!(fs % 2)) fs--;       // :34
  if (radius) *radius = fs / 2;        // :36
  if (plane_size) *plane_size = fs * fs;  // :38
  return fs;
}

int farms_create(farms_ctx **out, const farms_config *cfg) {
  if (!out || !cfg) return FARMS_ERR_ARG;
  *out = nullptr;
  if (cfg->width <= 0 || cfg->height <= 0 || cfg->width > 65535 || cfg->height > 65535) return FARMS_ERR_ARG;
  if ((uint64_t)cfg->width * (uint64_t)cfg->height >= (1ull << 31)) return FARMS_ERR_ARG;
  farms_ctx *c = new (std::nothrow) farms_ctx();
  if (!c) return FARMS_ERR_NOMEM;
  c->cfg = *cfg;
  c->W = cfg->width;
  c->H = cfg->height;
  int32_t rr = 0, pp = 0;
  c->fs = farms_normalize_filtersize(cfg->filtersize, &rr, &pp);
  c->r = rr;
  c->P = pp;
  c->min_inl = cfg->inlier_check;
  c->npx = (size_t)c->W * c->H;
  c->serial = (cfg->flags & FARMS_FLAG_SERIAL_SEMANTICS) != 0;
  auto bail = [&](int code) {
    farms_destroy(c);
    return code;
  };
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
    cudaGetLastError();
    return bail(FARMS_ERR_CUDA);  // no CPU fallback
  }
  if (cudaSetDevice(cfg->device) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  if (prop.major < 10) return bail(FARMS_ERR_CUDA);  // built for sm_100a only
  c->num_sms = prop.multiProcessorCount;
  c->fit_chunk = FIT_CHUNK_MIN;
  while (c->fit_chunk < FIT_CHUNK_MAX && (size_t)c->fit_chunk * 6 < c->npx) c->fit_chunk *= 2;
  if (cfg->fit_chunk) c->fit_chunk = (int)std::min<uint32_t>(std::max<uint32_t>(cfg->fit_chunk, 1024u), 1u << 20);
  if (cfg->pool_variant > 8) return bail(FARMS_ERR_ARG);
  if (cfg->pool_variant) c->pool_impl = (int)cfg->pool_variant;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  for (int q = 0; q < FIT_WAYS - 1; q++) {
    if (cudaStreamCreateWithFlags(&c->fit_streams[q], cudaStreamNonBlocking) != cudaSuccess) return bail(FARMS_ERR_CUDA);
    if (cudaEventCreateWithFlags(&c->ev_fit[q], cudaEventDisableTiming) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  }
  if (cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  for (int i = 0; i < EV_COUNT; i++)
    if (cudaEventCreate(&c->ev[i]) != cudaSuccess || cudaEventCreate(&c->evb[0][i]) != cudaSuccess ||
        cudaEventCreate(&c->evb[1][i]) != cudaSuccess)
      return bail(FARMS_ERR_CUDA);
  {
    cudaEvent_t *evs[] = {&c->ev_h2d[0], &c->ev_h2d[1], &c->ev_ingest[0], &c->ev_ingest[1], &c->ev_pool[0],
                          &c->ev_pool[1], &c->ev_d2h[0], &c->ev_d2h[1], &c->ev_c0, &c->ev_c1, &c->ev_fitdone[0],
                          &c->ev_fitdone[1]};
    for (cudaEvent_t *e : evs)
      if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  }
  bool ok = true;
  // (+1: the dummy pixel W*H that the ghost event of a serial-semantics stream hashes to)
  ok &= cudaMalloc((void **)&c->sae, (c->npx + 1) * sizeof(uint2)) == cudaSuccess;
  for (int q = 0; q < FIT_WAYS - 1; q++) ok &= cudaMalloc((void **)&c->sae_x[q], (c->npx + 1) * sizeof(uint2)) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->hx, HALO_CAP * 2) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->hy, HALO_CAP * 2) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->ht, HALO_CAP * 4) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->hm, HALO_CAP * 4) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->hlen, HALO_CAP * 8) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->hlcx, HALO_CAP * 8) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->hlcy, HALO_CAP * 8) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->d_err, sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->d_counters, 8 * sizeof(unsigned long long)) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->d_work, 8 * sizeof(unsigned int)) == cudaSuccess;
  ok &= cudaMalloc((void **)&c->d_small, 64) == cudaSuccess;
  ok &= cudaMallocHost((void **)&c->h_small, 256) == cudaSuccess;
  if (!ok) return bail(FARMS_ERR_NOMEM);
  launch_sae_init(c->sae, c->npx + 1, c->stream);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  *out = c;
  return FARMS_OK;
}

void farms_destroy(farms_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaDeviceSynchronize();
  free_owned(c->ws[0]);
  free_owned(c->ws[1]);
  for (int q = 0; q < FIT_WAYS - 1; q++) {
    if (c->sae_x[q]) cudaFree(c->sae_x[q]);
    if (c->fit_streams[q]) cudaStreamDestroy(c->fit_streams[q]);
    if (c->ev_fit[q]) cudaEventDestroy(c->ev_fit[q]);
  }
  void *ps[] = {c->sae, c->hx, c->hy, c->ht, c->hm, c->hlen, c->hlcx, c->hlcy, c->d_err, c->d_counters, c->d_work,
                c->d_small, c->in_x[0], c->in_y[0], c->in_t[0], c->in_x[1], c->in_y[1], c->in_t[1], c->sort_temp.p,
                c->scan_temp.p, c->cell_start.p, c->fit_scratch.p, c->surf_tmp.p, c->item_ovf.p, c->io_x.p, c->io_y.p,
                c->io_t.p, c->io_surf_t.p, c->io_surf_hit.p, c->time_table.p};
  for (void *p : ps)
    if (p) cudaFree(p);
  if (c->h_small) cudaFreeHost(c->h_small);
  for (int i = 0; i < EV_COUNT; i++) {
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->evb[0][i]) cudaEventDestroy(c->evb[0][i]);
    if (c->evb[1][i]) cudaEventDestroy(c->evb[1][i]);
  }
  cudaEvent_t evs[] = {c->ev_h2d[0], c->ev_h2d[1], c->ev_ingest[0], c->ev_ingest[1], c->ev_pool[0], c->ev_pool[1],
                       c->ev_d2h[0], c->ev_d2h[1], c->ev_c0, c->ev_c1, c->ev_fitdone[0], c->ev_fitdone[1]};
  for (cudaEvent_t e : evs)
    if (e) cudaEventDestroy(e);
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *farms_last_error(const farms_ctx *c) { return c ? c->err.c_str() : "null context"; }

int farms_get_params(const farms_ctx *c, int32_t *filtersize, int32_t *radius, int32_t *plane_size) {
  if (!c) return FARMS_ERR_ARG;
  if (filtersize) *filtersize = c->fs;
  if (radius) *radius = c->r;
  if (plane_size) *plane_size = c->P;
  return FARMS_OK;
}

int farms_process_host(farms_ctx *c, const uint16_t *x, const uint16_t *y, const uint64_t *t, const uint8_t *p,
                       uint64_t n, const farms_out *out) {
  (void)p;
  return farms_process_impl(c, x, y, t, n, out, false, false, 0, nullptr);
}

int farms_process_device(farms_ctx *c, const uint16_t *x, const uint16_t *y, const uint64_t *t, const uint8_t *p,
                         uint64_t n, const farms_out *out) {
  (void)p;
  return farms_process_impl(c, x, y, t, n, out, true, true, 0, nullptr);
}

int farms_reserve(farms_ctx *c, uint64_t n, int host_io) {
  if (!c) return FARMS_ERR_ARG;
  if (n == 0) return FARMS_OK;
  CU(cudaSetDevice(c->cfg.device));
  const uint64_t maxb = c->cfg.max_batch ? c->cfg.max_batch : DEFAULT_MAX_BATCH;
  const size_t nb = (size_t)std::min<uint64_t>(n, maxb);
  // a later batch carries the halo of the one before: room for it, as run_batch would make on demand
  const size_t m = nb + (n > maxb ? std::min<size_t>(HALO_CAP, nb) / 4 : 0);
  int rc;
  const int nsets = (host_io && n > 0) ? 2 : 1;
  for (int k = 0; k < nsets; k++)
    if ((rc = alloc_working(c, c->ws[k], m + m / 8 + 1024))) return rc;
  const size_t scratch_bytes = (plane_fit_scratch_bytes(c->r, (size_t)c->fit_chunk) + 255) & ~(size_t)255;
  if ((rc = ensure(c, c->fit_scratch, FIT_WAYS * scratch_bytes))) return rc;
  if (host_io && c->cap_in < nb) {
    CU(cudaDeviceSynchronize());
    for (int k = 0; k < 2; k++) {
      if (c->in_x[k]) { cudaFree(c->in_x[k]); cudaFree(c->in_y[k]); cudaFree(c->in_t[k]); }
      c->in_x[k] = nullptr; c->in_y[k] = nullptr; c->in_t[k] = nullptr;
    }
    c->cap_in = 0;
    for (int k = 0; k < 2; k++) {
      CU(cudaMalloc((void **)&c->in_x[k], nb * 2));
      CU(cudaMalloc((void **)&c->in_y[k], nb * 2));
      CU(cudaMalloc((void **)&c->in_t[k], nb * 8));
    }
    c->cap_in = nb;
  }
  return FARMS_OK;
}

uint64_t farms_num_events(const farms_ctx *c) { return c ? c->total_events : 0; }

int farms_get_timings(const farms_ctx *c, farms_timings *out) {
  if (!c || !out) return FARMS_ERR_ARG;
  *out = c->tm;
  return FARMS_OK;
}

int farms_reset(farms_ctx *c) {
  if (!c) return FARMS_ERR_ARG;
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaDeviceSynchronize());
  c->d2h_pending[0] = c->d2h_pending[1] = false;
  launch_sae_init(c->sae, c->npx + 1, c->stream);
  CU(cudaStreamSynchronize(c->stream));
  c->ghost_seen = c->ghost_prev_pending = false;
  c->have_t0 = false;
  c->t0 = 0;
  c->total_events = 0;
  c->last_M = 0;
  c->halo = 0;
  c->history_cut = false;
  c->err.clear();
  return FARMS_OK;
}

int farms_set_t0(farms_ctx *c, uint64_t t0) {
  if (!c) return FARMS_ERR_ARG;
  if (c->total_events) return farms_fail(c, FARMS_ERR_STATE, "farms_set_t0 after events were processed");
  c->t0 = t0;
  c->have_t0 = true;
  return FARMS_OK;
}

int farms_state_export(farms_ctx *c, uint32_t *d_last_t, uint8_t *d_hit) {
  if (!c || !d_last_t || !d_hit) return FARMS_ERR_ARG;
  CU(cudaSetDevice(c->cfg.device));
  launch_sae_export(c->sae, c->npx, d_last_t, d_hit, c->stream);
  CU(cudaStreamSynchronize(c->stream));
  return FARMS_OK;
}

int farms_state_fold(farms_ctx *c, const uint32_t *d_last_t, const uint8_t *d_hit) {
  if (!c || !d_last_t || !d_hit) return FARMS_ERR_ARG;
  CU(cudaSetDevice(c->cfg.device));
  launch_sae_fold(c->sae, c->npx, d_last_t, d_hit, c->stream);
  CU(cudaStreamSynchronize(c->stream));
  return FARMS_OK;
}

// last event per pixel of a device-resident slice, accumulated into the packed surface (surf_tmp); `first` clears it
static int slice_surface_accumulate(farms_ctx *c, const uint16_t *d_x, const uint16_t *d_y, const uint64_t *d_t,
                                    uint64_t n, uint64_t index_base, uint64_t t0, bool first) {
  int rc;
  if ((rc = ensure(c, c->surf_tmp, c->npx * sizeof(unsigned long long)))) return rc;
  if (first) {
    CU(cudaMemsetAsync(c->surf_tmp.p, 0, c->npx * sizeof(unsigned long long), c->stream));
    CU(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
  }
  launch_slice_surface(d_x, d_y, d_t, (size_t)n, (uint32_t)index_base, t0, c->W, c->H,
                       (unsigned long long *)c->surf_tmp.p, c->d_err, c->stream);
  CU(cudaGetLastError());
  return FARMS_OK;
}

static int slice_surface_finish(farms_ctx *c, uint32_t *d_last_t, uint8_t *d_hit) {
  launch_unpack_surface((const unsigned long long *)c->surf_tmp.p, c->npx, d_last_t, d_hit, c->stream);
  k_publish<<<1, 32, 0, c->stream>>>(c->h_small + 8, (const uint32_t *)c->d_err, 1);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  if (((int *)c->h_small)[8]) return farms_fail(c, FARMS_ERR_RANGE, "event outside the %dx%d sensor", c->W, c->H);
  return FARMS_OK;
}

int farms_slice_surface(farms_ctx *c, const uint16_t *d_x, const uint16_t *d_y, const uint64_t *d_t, uint64_t n,
                        uint64_t t0, uint32_t *d_last_t, uint8_t *d_hit) {
  if (!c || !d_last_t || !d_hit || (n && (!d_x || !d_y || !d_t))) return FARMS_ERR_ARG;
  if (n >= (1ull << 32) - 1) return farms_fail(c, FARMS_ERR_ARG, "slice too long");
  CU(cudaSetDevice(c->cfg.device));
  int rc;
  if ((rc = slice_surface_accumulate(c, d_x, d_y, d_t, n, 0, t0, true))) return rc;
  return slice_surface_finish(c, d_last_t, d_hit);
}

// Host arrays: the slice goes through the context's input staging buffers in pieces (all on the compute stream:
// every copy is ordered before the kernel that reads it), the two surfaces come back the same way.
int farms_slice_surface_host(farms_ctx *c, const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t n,
                             uint64_t t0, uint32_t *last_t, uint8_t *hit) {
  if (!c || !last_t || !hit || (n && (!x || !y || !t))) return FARMS_ERR_ARG;
  if (n >= (1ull << 32) - 1) return farms_fail(c, FARMS_ERR_ARG, "slice too long");
  CU(cudaSetDevice(c->cfg.device));
  int rc;
  const size_t piece = 4u << 20;
  if ((rc = ensure(c, c->io_x, piece * 2)) || (rc = ensure(c, c->io_y, piece * 2)) || (rc = ensure(c, c->io_t, piece * 8)) ||
      (rc = ensure(c, c->io_surf_t, c->npx * 4)) || (rc = ensure(c, c->io_surf_hit, c->npx)))
    return rc;
  if ((rc = slice_surface_accumulate(c, nullptr, nullptr, nullptr, 0, 0, t0, true))) return rc;
  for (uint64_t off = 0; off < n; off += piece) {
    const size_t nb = (size_t)std::min<uint64_t>(piece, n - off);
    CU(cudaMemcpyAsync(c->io_x.p, x + off, nb * 2, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->io_y.p, y + off, nb * 2, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->io_t.p, t + off, nb * 8, cudaMemcpyHostToDevice, c->stream));
    if ((rc = slice_surface_accumulate(c, (const uint16_t *)c->io_x.p, (const uint16_t *)c->io_y.p,
                                       (const uint64_t *)c->io_t.p, nb, off, t0, false)))
      return rc;
  }
  if ((rc = slice_surface_finish(c, (uint32_t *)c->io_surf_t.p, (uint8_t *)c->io_surf_hit.p))) return rc;
  CU(cudaMemcpyAsync(last_t, c->io_surf_t.p, c->npx * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(hit, c->io_surf_hit.p, c->npx, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return FARMS_OK;
}

int farms_state_fold_host(farms_ctx *c, const uint32_t *last_t, const uint8_t *hit) {
  if (!c || !last_t || !hit) return FARMS_ERR_ARG;
  CU(cudaSetDevice(c->cfg.device));
  int rc;
  if ((rc = ensure(c, c->io_surf_t, c->npx * 4)) || (rc = ensure(c, c->io_surf_hit, c->npx))) return rc;
  CU(cudaMemcpyAsync(c->io_surf_t.p, last_t, c->npx * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->io_surf_hit.p, hit, c->npx, cudaMemcpyHostToDevice, c->stream));
  launch_sae_fold(c->sae, c->npx, (const uint32_t *)c->io_surf_t.p, (const uint8_t *)c->io_surf_hit.p, c->stream);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  return FARMS_OK;
}

void *farms_host_alloc(uint64_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, (size_t)std::max<uint64_t>(bytes, 1), cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void farms_host_free(void *p) {
  if (p) cudaFreeHost(p);
}
int farms_host_register(void *p, uint64_t bytes) {
  if (!p || !bytes) return FARMS_ERR_ARG;
  if (cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable) != cudaSuccess) {
    cudaGetLastError();
    return FARMS_ERR_CUDA;
  }
  return FARMS_OK;
}
int farms_host_unregister(void *p) {
  if (!p) return FARMS_ERR_ARG;
  if (cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return FARMS_ERR_CUDA;
  }
  return FARMS_OK;
}

int farms_pack4_f32(farms_ctx *c, const double *d_a, const double *d_b, const double *d_c, const double *d_d,
                    uint64_t n, float *d_out4) {
  if (!c || (n && (!d_a || !d_b || !d_c || !d_d || !d_out4))) return FARMS_ERR_ARG;
  CU(cudaSetDevice(c->cfg.device));
  launch_pack4(d_a, d_b, d_c, d_d, (size_t)n, (float4 *)d_out4, c->stream);
  CU(cudaStreamSynchronize(c->stream));
  return FARMS_OK;
}

}  // extern "C"
