// farms_dev.cuh -- device-side data layout and kernel launchers shared by the .cu files.
//
// One internal batch = [halo | new] events, m = h + n entries, all arrays structure-of-arrays in HBM
// and indexed by the batch-local event index (stream order):
//   ex, ey  u16   pixel                      et   u32  t - t0   (src/vFlow.cpp:241)
//   em      u32   running max of et (makes time slabs monotone in the index even for unsorted input)
//   pix     u32   x*H + y  (EventMatrix flat index, include/EventMatrix.h:32-34)
//   prevp   int2  {index of previous event at the same pixel (or SAE_OLD / SAE_NEVER), its time}
//   nextp   i32   index of the next event at the same pixel, or INT_MAX
//   vx, vy, len, lcx, lcy  f64  local flow, |flow|, |flow|*cos(theta), |flow|*sin(theta)
// Per-pixel persistent state (the reference's cSurf, src/vFlow.cpp:93, 267):
//   sae     uint2 {time of the latest event, its batch-local index or SAE_OLD / SAE_NEVER}
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define SAE_OLD (-1)    // pixel was hit by an event of an earlier batch
#define SAE_NEVER (-2)  // pixel never hit: the reference cell is Event(0,0,0,0) (src/vFlow.cpp:80)
#define NEXT_NONE 0x7fffffff

#define FARMS_KILL_OLD_FLOW_TIME 500  // src/vFlow.cpp:961
#define FARMS_WINDOW_JUMP 5           // src/vFlow.cpp:73
#define FARMS_MAX_WINDOW 50           // src/vFlow.cpp:74
#define FARMS_NSCALES 11
// Pooling time slabs are 2^slab_shift us long (PoolGeom::slab_shift, chosen per batch from the density of flow
// events): 128 us for dense streams (64 us measured slower: 45.7 against 38.0 ms per 20 M events at 1280x720),
// longer for sparse ones so that a slab of one tile region holds ~100+ records.
#define FARMS_SLAB_SHIFT_MIN 7
#define FARMS_SLAB_SHIFT_MAX 14
// dense slabs a 500-us window can reach back from the slab of its event, at most: ceil(499 / 128)
#define FARMS_SLAB_LOOKBACK 4

// ---- self-checking build (make checked -> libfarms_b200_checked.so, -DFARMS_CHECKED) ----
// compute-sanitizer is closed on the B200 pool this was developed on, so the index arithmetic of the kernels carries
// its own bounds checks: FARMS_CHK records the first violated check (a code per site) in a per-file device word
// instead of making the access; farms_process_* then fails with FARMS_ERR_STATE naming the code.  Compiled out of
// the product build.
#ifdef FARMS_CHECKED
#define FARMS_CHK_DECL static __device__ unsigned int g_farms_chk[2];
#define FARMS_CHK(cond, code) ((cond) ? true : (atomicCAS(&g_farms_chk[0], 0u, (unsigned int)(code)), false))
#else
#define FARMS_CHK_DECL
#define FARMS_CHK(cond, code) (true)
#endif
// first failed check of each kernel file (0 = none), clearing it; only meaningful in the checked build
unsigned int farms_chk_pooling(cudaStream_t s);
unsigned int farms_chk_planefit(cudaStream_t s);
unsigned int farms_chk_index(cudaStream_t s);

struct FitParams {
  int W, H, r, P, min_inl;
};

struct FitOut {       // indexed by event
  double *vx, *vy;    // raw local result (src/vFlow.cpp:938-939)
  double *len, *theta, *lcx, *lcy;
  uint8_t *valid;
  int8_t *best_window;
  uint16_t *inliers;
  double *det;        // may be null
};

struct PoolGeom {
  int W, H;
  int tile_shift;     // tiles of (1<<tile_shift)^2 pixels
  int ntx, nty;       // tiles per axis; tile id = tx*nty + ty (y fastest, like the pixel layout)
  int slab_shift;     // time slabs of 2^slab_shift us
};

// ---- sort.cu ----
// Stable LSD radix sort of (key, value) pairs, 8 bits per pass.  Result ends in (keys_out, vals_out)
// if the number of passes is odd, else back in (keys, vals); returns which (0 = in place, 1 = out).
size_t radix_sort_temp_bytes(size_t n);
int radix_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_out, uint32_t *vals_out, size_t n,
                     int key_bits, void *temp, cudaStream_t s, uint64_t *launches);
size_t scan_temp_bytes(size_t n);
// exclusive prefix sum (in != out allowed to alias); total written to *d_total if non-null
void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, void *temp, cudaStream_t s,
                        uint64_t *launches);
// inclusive running maximum seeded with `seed`
void inclusive_max_scan_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t seed, void *temp,
                            cudaStream_t s, uint64_t *launches);

// ---- index.cu ----
// ghost: index (among the n new events) of an event that gets the dummy pixel key W*H, i.e. stays out of the
// surface of active events (the first event of a stream under FARMS_FLAG_SERIAL_SEMANTICS), or -1
void launch_ingest(const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t t0, size_t n,
                   int W, int H, uint16_t *ex, uint16_t *ey, uint32_t *et, uint32_t *pix, uint32_t *idx,
                   uint32_t idx_base, int ghost, int *err_flag, cudaStream_t s);
// serial semantics: own_ok[i] = the event's own pixel passes the age test with the time of the PREVIOUS event at
// the pixel (src/vFlow.cpp:790); the ghost event's outputs are cleared
void launch_serial_fix(const int2 *prevp, const uint32_t *et, const uint32_t *pix, size_t h, size_t m, int ghost_index,
                       uint32_t ghost_pix, uint64_t ghost_raw_t, int ghost_prev_pending, FitOut fo, uint8_t *own_ok,
                       uint32_t *ghost_consumed, cudaStream_t s);
void launch_halo_keys(const uint16_t *ex, const uint16_t *ey, size_t h, int H, uint32_t *pix, uint32_t *idx,
                      cudaStream_t s);
void launch_links(const uint32_t *skeys, const uint32_t *svals, const uint32_t *et, const uint2 *sae, size_t m,
                  int2 *prevp, int32_t *nextp, cudaStream_t s);
void launch_slab_flags(const uint32_t *em, const uint32_t *et, size_t m, size_t h, int slab_shift, uint32_t *flags,
                       uint32_t *nonmono, uint32_t *regress, cudaStream_t s);
void launch_slice_surface(const uint16_t *x, const uint16_t *y, const uint64_t *t, size_t n, uint32_t index_base,
                          uint64_t t0, int W, int H, unsigned long long *packed, int *err_flag, cudaStream_t s);
void launch_unpack_surface(const unsigned long long *packed, size_t npx, uint32_t *last_t, uint8_t *hit,
                           cudaStream_t s);

void launch_pack4(const double *a, const double *b, const double *c, const double *d, size_t n, float4 *out,
                  cudaStream_t s);

// ---- planefit.cu ----
void launch_sae_init(uint2 *sae, size_t npx, cudaStream_t s);
void launch_sae_advance(uint2 *sae, const uint32_t *pix, const uint32_t *et, const int32_t *nextp, int c0, int c1,
                        cudaStream_t s);
void launch_sae_finalize(uint2 *sae, const uint32_t *pix, const int32_t *nextp, int m, cudaStream_t s);
size_t plane_fit_scratch_bytes(int r, size_t chunk_events);
int launch_plane_fit(const uint2 *sae, const int2 *prevp, const uint16_t *ex, const uint16_t *ey,
                     const uint32_t *et, int i0, int i1, FitParams fp, FitOut fo, unsigned long long *valid_count,
                     void *scratch, cudaStream_t s);
void launch_sae_export(const uint2 *sae, size_t npx, uint32_t *last_t, uint8_t *hit, cudaStream_t s);
void launch_sae_fold(uint2 *sae, size_t npx, const uint32_t *last_t, const uint8_t *hit, cudaStream_t s);

// ---- pooling.cu ----
void launch_cell_keys(const uint16_t *ex, const uint16_t *ey, const uint32_t *em, const uint32_t *excl,
                      const double *len, size_t m, PoolGeom g, uint32_t ncells, uint32_t *keys, uint32_t *idx,
                      uint32_t *slab_ids, uint32_t *slab_first, cudaStream_t s);
void launch_build_records(const uint32_t *skeys, const uint32_t *sidx, size_t m, const uint16_t *ex,
                          const uint16_t *ey, const uint32_t *et, const int32_t *nextp, const double *len,
                          const double *lcx, const double *lcy, int monotone, uint4 *rec, double *pay,
                          uint32_t *cell_start, uint32_t ncells, uint32_t h, unsigned int *n_targets,
                          const uint32_t *time_table, uint32_t tt_base, uint32_t tt_size, cudaStream_t s);
// time_table[u] = first index with et >= tt_base + u (sorted timestamps); tt_size entries
void launch_time_table(const uint32_t *et, size_t m, uint32_t tt_base, uint32_t *time_table, uint32_t tt_size,
                       cudaStream_t s);
// returns the number of kernels launched.  work_counter: eight zeroed words (three work counters, then the per-batch
// words of PoolArgs::batch_words, whose [1] launch_build_records fills); done: m zeroed bytes; fin: m - h zeroed words.
// cand_count: four counters {candidates inspected, events pooled by the first fast pass, by the flagged second pass,
// by k_pool_any}; kernels_used: FARMS_POOLK_* bits of the kernels launched are OR-ed in.
int launch_pooling(const uint4 *rec, const double *pay, const uint32_t *cell_start, const uint32_t *slab_ids,
                   const uint32_t *slab_first, uint32_t *fin, uint32_t *item_ovf, uint8_t *done, size_t m, uint32_t ncells, int h,
                   const double *ev_len, const double *ev_lcx, const double *ev_lcy, int nslabs, PoolGeom g, int fast,
                   double flow_per_slab, double *global_r, double *global_theta, uint8_t *scale,
                   unsigned int *work_counter, unsigned long long *cand_count, int num_sms, cudaStream_t s,
                   unsigned *kernels_used, const uint8_t *own_ok);
int pool_tile_smem_bytes();
// words of the zeroed item_ovf array launch_pooling needs
size_t pool_item_words(int W, int H, int nslabs);
