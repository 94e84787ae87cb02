// farms_ctx.cuh -- the context object behind the C ABI, shared by ctx.cu (batch pipeline) and comm.cu (multi-GPU).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/farms_b200.h"
#include "farms_dev.cuh"

// events per internal batch: 32 Mi measured best at 1280x720 (200 M events: 8 Mi 397, 16 Mi 410, 32 Mi 419, 64 Mi 420
// Mevents/s device-resident; end to end 16 Mi 396, 32 Mi 400, 64 Mi 392: longer pipeline fill and drain)
constexpr uint64_t DEFAULT_MAX_BATCH = 32ull << 20;
constexpr uint32_t DEFAULT_SLACK_US = 1000;
constexpr size_t HALO_CAP = 8ull << 20;       // events carried across a batch boundary at most
#ifndef FARMS_FIT_CHUNK_LOG2
#define FARMS_FIT_CHUNK_LOG2 16
#endif
constexpr int FIT_CHUNK_MAX = 1 << FARMS_FIT_CHUNK_LOG2;  // events per SAE snapshot at most
constexpr int FIT_CHUNK_MIN = 1 << 13;
#ifndef FARMS_FIT_WAYS
#define FARMS_FIT_WAYS 4
#endif
constexpr int FIT_WAYS = FARMS_FIT_WAYS;  // plane-fit chunks in flight
constexpr size_t CSR_BUDGET = 96ull << 20;    // max (slab, tile) cells of the pooling index per batch

enum { EV_START, EV_H2D, EV_INGEST, EV_INDEX, EV_FIT, EV_BIN, EV_POOL, EV_END, EV_COUNT };

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

// The per-event working arrays of one internal batch.  The host path keeps two sets so that the result
// copies of batch k (device->host) overlap the kernels of batch k+1.
struct WorkSet {
  size_t cap = 0;  // capacity in events
  uint16_t *ex = nullptr, *ey = nullptr;
  uint32_t *et = nullptr, *em = nullptr, *keyA = nullptr, *valA = nullptr, *keyB = nullptr, *valB = nullptr,
           *pixkeep = nullptr, *flags = nullptr, *slab_ids = nullptr, *slab_first = nullptr, *fin = nullptr;
  int2 *prevp = nullptr;
  int32_t *nextp = nullptr;
  double *vx = nullptr, *vy = nullptr, *len = nullptr, *theta = nullptr, *lcx = nullptr, *lcy = nullptr,
         *det = nullptr, *gr = nullptr, *gth = nullptr, *pay = nullptr;
  uint8_t *valid = nullptr, *scale = nullptr, *done = nullptr, *own = nullptr;
  int8_t *bw = nullptr;
  uint16_t *inl = nullptr;
  uint4 *rec = nullptr;
  std::vector<void *> owned;
};

struct farms_ctx {
  farms_config cfg{};
  int W = 0, H = 0, fs = 0, r = 0, P = 0, min_inl = 0;
  size_t npx = 0;
  int num_sms = 148;
  // Events per SAE snapshot.  A fit thread walks back one history link for every footprint cell that was hit
  // again later in its chunk, so the chunk is kept to about a sixth of an event per pixel, at most 2^16 events:
  // 2^16 at 1280x720 (measured with the two-stream overlap: 2^15 11.3 ms, 2^16 10.1, 2^17 10.7 per 20 M events),
  // 2^14 at 346x260 (2^13 38.9 ms, 2^14 31.7, 2^15 33.4; without the overlap 2^17 took 66 ms).  Small chunks are
  // launch-bound, large ones walk many dependent history links per event.
  int fit_chunk = FIT_CHUNK_MAX;
  // fast pooling kernel (farms_config.pool_variant): 7 = k_pool_tile16, three slabs per round (the default: measured
  // fastest), 1 = k_pool_tile, 2 = k_pool_bits, 3 = k_pool_tile one CTA per SM, 4 = k_pool_warp, 5 / 6 = k_pool_tile16
  // with two / four slabs per round, 8 = 7 with column-culled trips (measured 3 % slower)
  int pool_impl = 7;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[EV_COUNT]{};          // EV_START / EV_END of a whole call
  cudaEvent_t evb[2][EV_COUNT]{};      // stage events of the internal batches, two sets used alternately
  int ev_set = 0, stage_pending = -1;  // set being recorded; set whose stage times have not been read yet
  std::string err;
  bool have_t0 = false;
  uint64_t t0 = 0;
  // FARMS_FLAG_SERIAL_SEMANTICS: the stream's first event ("ghost") is not processed; its pixel keeps its raw time
  bool serial = false, ghost_seen = false, ghost_prev_pending = false;
  uint32_t ghost_pix = 0;
  uint64_t ghost_raw_t = 0;
  uint64_t total_events = 0;
  uint32_t last_M = 0;
  unsigned long long valid_seen = 0;  // flow events counted so far in the current process call
  size_t halo = 0;  // events in the halo store
  bool history_cut = false;  // events older than the halo were dropped (or, for a time slice, never supplied)
  size_t cap_in = 0;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_h2d[2]{}, ev_ingest[2]{}, ev_pool[2]{}, ev_d2h[2]{}, ev_fitdone[2]{}, ev_c0 = nullptr, ev_c1 = nullptr;
  bool d2h_pending[2] = {false, false};
  farms_timings tm{};

  // surface of active events, plus the extra copies and streams of the overlapped plane fit (chunk k works on
  // surface k mod FIT_WAYS in stream k mod FIT_WAYS; way 0 is `sae` on the main stream)
  uint2 *sae = nullptr, *sae_x[FIT_WAYS - 1] = {};
  cudaStream_t fit_streams[FIT_WAYS - 1] = {};
  cudaEvent_t ev_fit[FIT_WAYS - 1] = {};
  WorkSet ws[2];
  // halo store
  uint16_t *hx = nullptr, *hy = nullptr;
  uint32_t *ht = nullptr, *hm = nullptr;
  double *hlen = nullptr, *hlcx = nullptr, *hlcy = nullptr;
  // staging for the host path
  uint16_t *in_x[2] = {nullptr, nullptr}, *in_y[2] = {nullptr, nullptr};
  uint64_t *in_t[2] = {nullptr, nullptr};
  // misc
  DevBuf sort_temp, scan_temp, cell_start, fit_scratch, surf_tmp, item_ovf, time_table;
  DevBuf io_x, io_y, io_t, io_surf_t, io_surf_hit;  // staging of the host-array state helpers
  int *d_err = nullptr;
  unsigned long long *d_counters = nullptr;  // [0] valid events, [1] pool candidates, [2..4] events per pooling path
  unsigned pool_kernels = 0;
  unsigned int *d_work = nullptr;
  uint32_t *d_small = nullptr;  // device scratch words
  uint32_t *h_small = nullptr;  // pinned host scratch words
};


// What a batch hook sees once the pooling of one internal batch has been enqueued on the compute stream: the device
// columns of the events that get outputs (n_out of them, the first is output number out_off of the call).
struct FarmsBatchView {
  const double *gr, *gth, *lr, *lth;
  size_t n_out, out_off;
  cudaStream_t stream;  // the context's compute stream; the arrays are overwritten by later work on it
};
struct FarmsBatchHook {
  int (*fn)(void *user, farms_ctx *c, const FarmsBatchView *v) = nullptr;
  void *user = nullptr;
};

int farms_fail(farms_ctx *c, int code, const char *fmt, ...);
// The event loop behind farms_process_host / _device.  in_device / out_device: where x,y,t and the `out` columns
// live.  The first n_skip events are history only (processed, no outputs): `out` holds n - n_skip entries.
int farms_process_impl(farms_ctx *c, const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t n,
                       const farms_out *out, bool in_device, bool out_device, uint64_t n_skip,
                       const FarmsBatchHook *hook);
