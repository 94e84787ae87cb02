/* farms_b200.h -- C ABI of the B200-native FARMS (Fast Aperture-Robust Multi-Scale) event-flow path.
 *
 * The reference (himstien/aperture-robust-multiscale-optical-flow) has no FFI/plugin interface; its
 * only internal seam is the public surface of `vFlowManager` (reference include/vFlow.h:99-114):
 *     vFlowManager(int height, int width, int filterSize, int minEvtsOnPlane, std::string fileName);
 *     long runFileCopy(unsigned long numEvents);   // batch path, src/vFlow.cpp:111-460
 *     double getNumEvents();
 * Each entry point below cites the piece of that seam it replaces.  Plain pointers and sizes only;
 * no C++ or torch types cross the boundary.  All functions return 0 (FARMS_OK) or a negative
 * farms_status and never throw.  A context is not re-entrant (one per host thread, like the
 * reference object, whose scratch matrices are mutable members: include/vFlow.h:76-79).  The caller
 * owns every buffer it passes in; the context owns all device memory, streams and events.
 *
 * There is no CPU fallback: farms_create fails with FARMS_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef FARMS_B200_H
#define FARMS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FARMS_B200_ABI_VERSION 2

typedef enum {
  FARMS_OK = 0,
  FARMS_ERR_ARG = -1,      /* bad argument (NULL, non-positive size, ...)                          */
  FARMS_ERR_RANGE = -2,    /* an event lies outside the width x height sensor (the reference would
                              index out of bounds: src/vFlow.cpp:264-267)                           */
  FARMS_ERR_CUDA = -3,     /* CUDA runtime failure; farms_last_error() has the text                 */
  FARMS_ERR_NOMEM = -4,    /* host or device allocation failed                                     */
  FARMS_ERR_STATE = -5,    /* call order violated, or the stream cannot be carried across a batch boundary
                              exactly (history overflow, timestamps behind by more than reorder_slack_us) */
  FARMS_ERR_COMM = -6      /* NCCL failure (or libnccl.so.2 cannot be loaded); farms_last_error() has the text */
} farms_status;

typedef struct farms_ctx farms_ctx;

/* Replaces the vFlowManager constructor arguments (src/main.cpp:186-187, src/vFlow.cpp:22-40).
 * filtersize and inlier_check take the raw CLI values; filtersize is normalised exactly like
 * src/vFlow.cpp:32-36 (<5 -> 3, even -> -1). Zero-initialise, then set what you need. */
typedef struct {
  int32_t width;            /* --width   (default 320, src/main.cpp:22)                            */
  int32_t height;           /* --height  (default 320, src/main.cpp:21)                            */
  int32_t filtersize;       /* --filtersize (default 3, src/main.cpp:23)                           */
  int32_t inlier_check;     /* --inlierCheck = minEvtsOnPlane (default 5, src/main.cpp:24)         */
  int32_t device;           /* CUDA device ordinal                                                 */
  uint32_t flags;           /* FARMS_FLAG_*                                                        */
  uint64_t max_batch;       /* events per internal device batch; 0 = default (32 Mi)               */
  uint32_t reorder_slack_us;/* extra history (us) kept across batch boundaries for streams whose
                               timestamps are not perfectly sorted; 0 = default (1000).  A timestamp that runs
                               further behind the stream's maximum than this, after history was dropped at a
                               batch boundary, is FARMS_ERR_STATE (never a silently different result)  */
  /* Tuning / test selectors (0 = the library's own choice; never read from the environment): */
  uint32_t pool_variant;    /* fast pooling kernel: 0 = default (7); 1 k_pool_tile (20-byte staged records, 2 slabs
                               per round), 2 bit-table k_pool_bits, 3 k_pool_tile with one CTA per SM, 4 two-phase
                               k_pool_warp, 5 / 6 / 7 k_pool_tile16 (16-byte packed records) with 2 / 4 / 3 slabs per
                               round, 8 = 7 with column-culled trips.  All are parity-tested against the oracle;
                               7 measured fastest.                                                           */
  uint32_t fit_chunk;       /* events per plane-fit chunk (SAE snapshot interval); 0 = from the sensor size  */
  uint32_t slab_target;     /* flow events per (tile region, time slab) the slab length is chosen for; 0 = 70 */
  uint32_t reserved[4];
} farms_config;

#define FARMS_FLAG_DEBUG_DET 1u       /* also produce the determinant column (farms_out.det)       */
#define FARMS_FLAG_EXACT_POOLING 2u   /* pool every event with the general FP64 kernel: slower, sums
                                         accurate to ~1e-15 instead of ~1e-7..2e-6 relative.  The default fast
                                         path keeps FP32 ring partials: its scale decision is accepted only when
                                         the winning mean beats every rival by a 2e-5 margin, otherwise the event is
                                         pooled again in FP64 (see csrc/pooling.cu)                         */
#define FARMS_FLAG_GENERIC_POOLING FARMS_FLAG_EXACT_POOLING
#define FARMS_FLAG_SERIAL_SEMANTICS 4u /* the semantics of the reference's DEFAULT driver vFlowManager::run
                                         (src/vFlow.cpp:465-826) instead of runFileCopy: the first event of the
                                         stream only sets t0 -- it is not inserted into the surface of active events
                                         and leaves its RAW timestamp as its pixel's lastEventTime (:531-558); and
                                         lastEventTime[x][y] is written only AFTER pooling (:790), so an event's own
                                         pixel is pooled with the time of the PREVIOUS event there (the own-flow
                                         fallback of :1085-1094 becomes reachable).  Implies exact pooling.  The
                                         reference computes these numbers but writes no file in that mode.        */

/* Per-event results, structure of arrays, n entries each, caller-allocated.  Any pointer may be
 * NULL (that column is skipped).  Columns follow the reference's batch output row
 * `x y t p globalR globalTheta Vx Vy localR localTheta scale` (src/vFlow.cpp:438); x, y, p are
 * echoes of the input and stay with the caller.  For events without valid flow every column is 0
 * except vx/vy, which carry the raw local result (src/vFlow.cpp:386-396). */
typedef struct {
  uint32_t *t_rel;       /* column 3: (uint32)(t - t0)                   src/vFlow.cpp:241, 373    */
  double *global_r;      /* column 5                                     src/vFlow.cpp:365         */
  double *global_theta;  /* column 6                                     src/vFlow.cpp:366         */
  double *vx;            /* column 7                                     src/vFlow.cpp:378, 394    */
  double *vy;            /* column 8                                                               */
  double *local_r;       /* column 9                                     src/vFlow.cpp:324         */
  double *local_theta;   /* column 10                                    src/vFlow.cpp:325         */
  uint8_t *scale;        /* column 11: 0,5,...,50                        src/vFlow.cpp:380         */
  uint8_t *valid;        /* 1 iff the event has valid local flow         src/vFlow.cpp:315         */
  int8_t *best_window;   /* diagnostic: winning candidate window 0..8 (i outer, j inner,
                            src/vFlow.cpp:870-910), -1 if none fits inside the sensor              */
  uint16_t *inliers;     /* diagnostic: computeGrads' return value       src/vFlow.cpp:1352-1369   */
  double *det;           /* diagnostic (FARMS_FLAG_DEBUG_DET): DET       src/vFlow.cpp:1316        */
} farms_out;

/* Stage timings of the most recent farms_process_* call, measured with CUDA events on the
 * context's own stream (milliseconds, summed over internal batches). */
typedef struct {
  float total_ms;        /* first kernel/copy enqueued -> last result resident                    */
  float h2d_ms;          /* host->device copies (host path only; overlapped time is not removed)  */
  float ingest_ms;       /* K1: rebase, range check, pixel keys, prefix-max time                  */
  float index_ms;        /* K2: stable radix sort by pixel + prev/next links                      */
  float fit_ms;          /* K3: SAE advance + local plane fit                                     */
  float bin_ms;          /* K4a: (time slab, tile) binning of flow events                         */
  float pool_ms;         /* K4b: multi-scale pooling                                              */
  float d2h_ms;          /* device->host copies (host path only)                                  */
  uint64_t events;       /* events processed                                                       */
  uint64_t valid_events; /* events with valid local flow                                           */
  uint64_t kernel_launches;
  uint64_t pool_candidates; /* candidate flow events inspected by the pooling kernel               */
  uint64_t pool_kernels;    /* FARMS_POOLK_* bits: which pooling kernels were launched              */
  uint64_t pool_events[3];  /* events pooled by [0] the first fast pass (FP32 partials, or its in-kernel FP64
                               re-pool when the scale decision was not clear-cut), [1] the flagged second pass
                               with larger slots, [2] the general exact kernel k_pool_any                */
} farms_timings;

#define FARMS_POOLK_TILE_DENSE 1u    /* k_pool_tile<8 warps, 512-record slots, 2 slabs per round, 2 CTAs/SM> */
#define FARMS_POOLK_TILE_SPARSE 2u   /* k_pool_tile<8, 320, 4, 2>: thin slabs                                 */
#define FARMS_POOLK_TILE_SECOND 4u   /* k_pool_tile<16, 768, 4, 1>, flagged second pass                      */
#define FARMS_POOLK_TILE_ONE_CTA 8u  /* k_pool_tile<16, 768, 4, 1> as the first pass (pool_variant 3)        */
#define FARMS_POOLK_BITS 16u         /* k_pool_bits (pool_variant 2)                                         */
#define FARMS_POOLK_ANY 32u          /* k_pool_any                                                           */
#define FARMS_POOLK_WARP_DENSE 64u   /* k_pool_warp<8 warps, 480-record slots, 2 slabs per round, 2 CTAs/SM>  */
#define FARMS_POOLK_WARP_SPARSE 128u /* k_pool_warp<8, 352, 4, 2>                                            */
#define FARMS_POOLK_WARP_SECOND 256u /* k_pool_warp<16, 768, 4, 1>, flagged second pass                      */
#define FARMS_POOLK_TILE16_DENSE 512u    /* k_pool_tile16 (16-byte packed staged records), dense streams      */
#define FARMS_POOLK_TILE16_SPARSE 1024u  /* k_pool_tile16<8, 416, 4, 2>, thin slabs                           */
#define FARMS_POOLK_TILE16_SECOND 2048u  /* k_pool_tile16<16, 960, 4, 1>, flagged second pass                 */
#define FARMS_POOLK_TILE16_XCULL 4096u   /* the k_pool_tile16 launches were the column-culled ones (pool_variant 8) */

/* ---- lifetime: replaces `vFlowManager vFlowM(...)` (src/main.cpp:186) ---- */
int farms_create(farms_ctx **out, const farms_config *cfg);
void farms_destroy(farms_ctx *ctx);
const char *farms_last_error(const farms_ctx *ctx); /* never NULL; "" when no error */
/* Back to the freshly constructed state (empty surfaces, no t0) while keeping device buffers: the
 * equivalent of constructing a new vFlowManager for the next recording. */
int farms_reset(farms_ctx *ctx);
int farms_abi_version(void);
/* 1 for libfarms_b200_checked.so (make checked: kernels compiled with -DFARMS_CHECKED carry their own bounds checks
 * and farms_process_* fails with FARMS_ERR_STATE when one trips), 0 for the product build */
int farms_build_is_checked(void);

/* The reference's filter-size normalisation (src/vFlow.cpp:32-38) as a pure host function: returns the
 * normalised filtersize and stores fRad and planeSize.  Needs no device. */
int farms_normalize_filtersize(int filtersize, int32_t *radius, int32_t *plane_size);
/* Normalised parameters actually in use (src/vFlow.cpp:32-38): filtersize, radius, plane size. */
int farms_get_params(const farms_ctx *ctx, int32_t *filtersize, int32_t *radius, int32_t *plane_size);

/* ---- the hot path: replaces the event loop of runFileCopy (src/vFlow.cpp:223-414) ----
 * Processes n more events in stream order.  State (surface of active events, flow surfaces, t0)
 * persists across calls, so feeding a stream in pieces is equivalent to one call.  t0 is the first
 * timestamp ever submitted (src/vFlow.cpp:194) unless farms_set_t0 was called.
 * x,y: pixel coordinates (u16); t: timestamp in microseconds (u64; t - t0 is reduced to u32 exactly
 * like the reference's `unsigned int`, include/vFlow.h:113); p: polarity (u8; has no effect on the
 * numbers -- the ON and OFF surfaces always hold identical values, src/vFlow.cpp:349-353 -- and
 * may be NULL).
 *   _host  : x,y,t,p,out are host pointers (pinned or pageable); copies are inside the call.
 *   _device: x,y,t,p,out are device pointers on cfg.device; the call returns after the work is
 *            complete (it synchronises the context's stream). */
int farms_process_host(farms_ctx *ctx, const uint16_t *x, const uint16_t *y, const uint64_t *t,
                       const uint8_t *p, uint64_t n, const farms_out *out);
int farms_process_device(farms_ctx *ctx, const uint16_t *x, const uint16_t *y, const uint64_t *t,
                         const uint8_t *p, uint64_t n, const farms_out *out);

/* replaces `getNumEvents()` (include/vFlow.h:108): events processed so far */
/* Optional: allocate the device working memory for calls of up to n events now (host_io != 0: also the staging
 * buffers of farms_process_host), so that the first farms_process_* call does not pay for it.  The reference
 * allocates its surfaces in the constructor (src/vFlow.cpp:47-93), outside its loop timer. */
int farms_reserve(farms_ctx *ctx, uint64_t n, int host_io);

uint64_t farms_num_events(const farms_ctx *ctx);
int farms_get_timings(const farms_ctx *ctx, farms_timings *out);

/* ---- state hand-over for time-sliced multi-GPU runs (no counterpart in the reference, which is
 * single-threaded; the state is the reference's cSurf surface, src/vFlow.cpp:93, 267) ----
 * All pointers are DEVICE pointers with width*height entries, flat index x*height + y like
 * EventMatrix (include/EventMatrix.h:32-34). */
int farms_set_t0(farms_ctx *ctx, uint64_t t0);
/* last event time (relative to t0) and hit flag of every pixel */
int farms_state_export(farms_ctx *ctx, uint32_t *d_last_t, uint8_t *d_hit);
/* overwrite the state where d_hit != 0 (call in stream order: later slices win) */
int farms_state_fold(farms_ctx *ctx, const uint32_t *d_last_t, const uint8_t *d_hit);
/* stateless helper: last event per pixel of a device-resident event slice (t rebased by t0) */
int farms_slice_surface(farms_ctx *ctx, const uint16_t *d_x, const uint16_t *d_y, const uint64_t *d_t,
                        uint64_t n, uint64_t t0, uint32_t *d_last_t, uint8_t *d_hit);

/* The same two steps with HOST arrays (width*height entries each), for callers without CUDA of their own -- the
 * FARMS_Flow command line uses them for its single-process multi-GPU mode (one host thread and one context per
 * GPU, surfaces handed over through host memory). */
int farms_slice_surface_host(farms_ctx *ctx, const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t n,
                             uint64_t t0, uint32_t *last_t, uint8_t *hit);
int farms_state_fold_host(farms_ctx *ctx, const uint32_t *last_t, const uint8_t *hit);

/* K5 helper for the final gather of a time-sliced run: interleave four f64 device columns (normally globalR,
 * globalTheta, localR, localTheta -- the README's 8-column contract minus the echoed x y t p) into one
 * float4-per-event device buffer that goes over NVLink in a single NCCL gather. */
int farms_pack4_f32(farms_ctx *ctx, const double *d_a, const double *d_b, const double *d_c, const double *d_d,
                    uint64_t n, float *d_out4);

/* ---- pinned host memory for callers without CUDA of their own (the FARMS_Flow command line): with pinned buffers
 * farms_process_host overlaps its host<->device copies with the kernels; pageable buffers work but serialise.
 * farms_host_register pins memory the caller already owns (e.g. the text reader's malloc'ed arrays). ---- */
void *farms_host_alloc(uint64_t bytes);            /* NULL on failure */
void farms_host_free(void *p);
int farms_host_register(void *p, uint64_t bytes);  /* FARMS_OK, or FARMS_ERR_CUDA (the buffer stays usable, pageable) */
int farms_host_unregister(void *p);

/* ---- time-sliced multi-GPU runs (no counterpart in the reference) --------------------------------
 * One farms_ctx per GPU -- one process per GPU, or one host thread per GPU -- joined into a farms_comm.  The recording
 * is cut into time slices, one per rank, in rank order.  Rank g passes the events of its slice INCLUDING the causal
 * halo in front of it:
 *     events [0, n_halo)     history only: processed (their local flow is state for the pooling of later events,
 *                            which admits |dt| < 500 us, src/vFlow.cpp:1002), no outputs
 *     events [n_halo, n)     owned: outputs are produced for these
 *     events [0, n_surface)  this rank's share of the surface exchange: the shares of ranks 0 .. g-1 must tile the
 *                            recording from its first event up to rank g's first (halo) event; 0 on the last rank
 * farms_comm_process is collective.  It all-gathers the ranks' "last event per pixel" surfaces, folds the earlier
 * ranks' surfaces into this context in rank order (the surface of active events never forgets, src/vFlow.cpp:267),
 * runs the event loop and -- if `gather` is given -- delivers the README's contract columns of every rank's owned
 * events to rank gather->root, batch by batch while the next batch computes.  The context must be fresh (just
 * created or farms_reset).  Transports: NCCL (libnccl.so.2, loaded on first use) or, with FARMS_COMM_LOCAL, direct
 * copies between host threads of one process. */
typedef struct farms_comm farms_comm;
#define FARMS_COMM_ID_BYTES 128
#define FARMS_COMM_LOCAL 1u            /* all ranks are threads of this process: no NCCL, ranks may share a device */
#define FARMS_IO_INPUT_ON_DEVICE 1u    /* x, y, t are device pointers                                           */
#define FARMS_IO_OUTPUT_ON_DEVICE 2u   /* the farms_out columns are device pointers                             */

typedef struct {
  int32_t root;      /* rank that receives                                                                       */
  float *dst;        /* root only: DEVICE buffer of 4 floats per owned event of the whole recording, rank after
                        rank: globalR, globalTheta, localR, localTheta (the README's 8-column row minus the echoed
                        x y t p; six printed digits fit a float)                                                  */
  uint64_t *counts;  /* optional, host, nranks entries, filled on every rank: owned events per rank            */
} farms_gather;

/* rank 0: 128 bytes that identify the group; hand them to every rank (file, pipe, torch.distributed ...) */
int farms_comm_unique_id(void *id128);
int farms_comm_create(farms_comm **out, farms_ctx *ctx, int nranks, int rank, const void *id128, uint32_t flags);
void farms_comm_destroy(farms_comm *comm);
/* transport: 0 none (one rank), 1 NCCL incl. the gather (ncclSend/ncclRecv), 2 in-process, 3 NCCL for the exchange
 * and the gather written into the root's buffer over peer memory (CUDA IPC mapping, copy-engine writes over NVLink:
 * no SM needed while the pooling kernel owns them; chosen when every rank can map the buffer; as of the last call) */
int farms_comm_info(const farms_comm *comm, int32_t *nranks, int32_t *rank, int32_t *transport);
/* host wall clock (ms) of the phases of the last farms_comm_process call on this rank: [0] slice made device-resident
 * + its "last event per pixel" surface, [1] surface exchange + fold, [2] the event loop, [3] drain of the output
 * transfers */
int farms_comm_phases(const farms_comm *comm, float ms[4]);
/* out: n - n_halo entries per column (any column, or out itself, may be NULL); io_flags: FARMS_IO_* */
int farms_comm_process(farms_comm *comm, const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t n,
                       uint64_t n_halo, uint64_t n_surface, uint64_t t0, uint32_t io_flags, const farms_out *out,
                       const farms_gather *gather);

#ifdef __cplusplus
}
#endif
#endif /* FARMS_B200_H */
