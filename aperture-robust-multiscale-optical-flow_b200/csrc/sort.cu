// sort.cu -- K2 building blocks: stable LSD radix sort of (key, value) pairs and prefix scans.
//
// The history index of the FARMS path is "events grouped by pixel, in stream order" (SURVEY.md 8(a),
// north_star part 2); the pooling index is "events grouped by (time slab, tile), in stream order".
// Both are stable sorts of 32-bit keys carrying the event index, done here with a classic three-kernel
// radix pass (per-tile digit histogram -> scan of the digit-major count table -> stable scatter).
// Everything is streaming, HBM-bound work: per pass each pair is read twice (4 B + 8 B) and written
// once (8 B).  Tiles are 4096 pairs; the grid is n/4096 CTAs, many waves over the 148 SMs.
#include "farms_dev.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 pairs per CTA
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_STRIP = RS_TILE / RS_WARPS;    // 512 consecutive pairs per warp

__global__ void __launch_bounds__(RS_THREADS) rs_hist(const uint32_t *__restrict__ keys, size_t n, int shift,
                                                      uint32_t *__restrict__ counts, int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * RS_TILE;
  const int lane = threadIdx.x & 31;
  // all loads first (16 independent requests per thread in flight), then the warp-aggregated counting: with the
  // load inside the counting loop every round waited for its own key (ncu: 40 % of the stall samples)
  uint32_t kk[RS_ITEMS];
#pragma unroll
  for (int it = 0; it < RS_ITEMS; it++) {
    const size_t i = base + (size_t)it * RS_THREADS + threadIdx.x;
    kk[it] = i < n ? keys[i] : 0u;
  }
#pragma unroll
  for (int it = 0; it < RS_ITEMS; it++) {
    size_t i = base + (size_t)it * RS_THREADS + threadIdx.x;
    bool ok = i < n;
    uint32_t d = ok ? ((kk[it] >> shift) & 255u) : (256u + lane);
    // warp-aggregate: sorted-ish inputs put whole warps on one digit
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (ok && lane == __ffs(peers) - 1) atomicAdd(&h[d], (uint32_t)__popc(peers));
  }
  __syncthreads();
  counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter.  Order inside a tile is (warp strip, round, lane) == memory order, so ranks computed
// with per-warp digit counters and match_any peer masks preserve the input order of equal digits.
// The tile is first reordered by digit in shared memory, then written out: the pairs of one digit leave as
// one contiguous run (16 pairs = 64 + 64 bytes on average) instead of 4096 scattered 4-byte stores.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const uint32_t *__restrict__ keys,
                                                         const uint32_t *__restrict__ vals,
                                                         uint32_t *__restrict__ keys_out,
                                                         uint32_t *__restrict__ vals_out, size_t n, int shift,
                                                         const uint32_t *__restrict__ offsets, int nblocks) {
  __shared__ uint32_t wc[RS_WARPS][256];
  __shared__ uint32_t lstart[256], gbase[256], wtot[RS_WARPS];
  __shared__ uint2 buf[RS_TILE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
  __syncthreads();
  const size_t tbase = (size_t)blockIdx.x * RS_TILE;
  const size_t wbase = tbase + (size_t)warp * RS_STRIP;
  uint32_t k[RS_ITEMS], v[RS_ITEMS], rank[RS_ITEMS];
  // all 32 loads of the thread first, then the ranking rounds (each of which synchronises the warp): with the loads
  // inside the ranking loop every round waited for its own key (ncu: 40 % of the stall samples on the first use)
#pragma unroll
  for (int it = 0; it < RS_ITEMS; it++) {
    const size_t i = wbase + (size_t)it * 32 + lane;
    const bool ok = i < n;
    k[it] = ok ? keys[i] : 0u;
    v[it] = ok ? vals[i] : 0u;
  }
#pragma unroll
  for (int it = 0; it < RS_ITEMS; it++) {
    size_t i = wbase + (size_t)it * 32 + lane;
    bool ok = i < n;
    uint32_t d = ok ? ((k[it] >> shift) & 255u) : (256u + lane);
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (ok && lane == leader) {
      old = wc[warp][d];
      wc[warp][d] = old + (uint32_t)__popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[it] = old + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  {
    // thread d: offsets of the warps' strips inside digit d, the digit's tile-local start (exclusive scan
    // over the 256 digit totals) and its global position
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
      uint32_t c = wc[w][d];
      wc[w][d] = run;
      run += c;
    }
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    uint32_t carry = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) carry += w < warp ? wtot[w] : 0u;
    lstart[d] = carry + inc - run;
    gbase[d] = offsets[(size_t)d * nblocks + blockIdx.x];
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < RS_ITEMS; it++) {
    size_t i = wbase + (size_t)it * 32 + lane;
    if (i < n) {
      uint32_t d = (k[it] >> shift) & 255u;
      buf[lstart[d] + wc[warp][d] + rank[it]] = make_uint2(k[it], v[it]);
    }
  }
  __syncthreads();
  const uint32_t cnt = (uint32_t)(n - tbase < (size_t)RS_TILE ? n - tbase : (size_t)RS_TILE);
#pragma unroll 4
  for (int it = 0; it < RS_ITEMS; it++) {
    const uint32_t j = (uint32_t)it * RS_THREADS + threadIdx.x;
    if (j < cnt) {
      const uint2 kv = buf[j];
      const uint32_t d = (kv.x >> shift) & 255u;
      const uint32_t dst = gbase[d] + (j - lstart[d]);
      keys_out[dst] = kv.x;
      vals_out[dst] = kv.y;
    }
  }
}

// ---------------- scans ----------------
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

template <bool MAXOP>
__device__ __forceinline__ uint32_t sc_op(uint32_t a, uint32_t b) {
  return MAXOP ? (a > b ? a : b) : a + b;
}

template <bool MAXOP>
__global__ void __launch_bounds__(SC_THREADS) sc_reduce(const uint32_t *__restrict__ in, size_t n,
                                                        uint32_t *__restrict__ partial) {
  __shared__ uint32_t ws[SC_THREADS / 32];
  const size_t base = (size_t)blockIdx.x * SC_TILE;
  uint32_t acc = 0;
#pragma unroll 4
  for (int it = 0; it < SC_ITEMS; it++) {
    size_t i = base + (size_t)it * SC_THREADS + threadIdx.x;
    if (i < n) acc = sc_op<MAXOP>(acc, in[i]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc = sc_op<MAXOP>(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < SC_THREADS / 32; w++) t = sc_op<MAXOP>(t, ws[w]);
    partial[blockIdx.x] = t;
  }
}

// out[i] = (INCLUSIVE ? op(carry, in[0..i]) : op(carry, in[0..i-1])), carry = op(seed, tile_prefix[b])
template <bool MAXOP, bool INCLUSIVE>
__global__ void __launch_bounds__(SC_THREADS) sc_apply(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                       size_t n, const uint32_t *__restrict__ tile_prefix,
                                                       uint32_t seed) {
  __shared__ uint32_t ws[SC_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t base = (size_t)blockIdx.x * SC_TILE;
  uint32_t carry = tile_prefix ? sc_op<MAXOP>(seed, tile_prefix[blockIdx.x]) : seed;
  for (int it = 0; it < SC_ITEMS; it++) {
    size_t i = base + (size_t)it * SC_THREADS + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x = sc_op<MAXOP>(x, y);
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    uint32_t wpre = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; w++) {
      uint32_t s = ws[w];
      if (w < warp) wpre = sc_op<MAXOP>(wpre, s);
      total = sc_op<MAXOP>(total, s);
    }
    uint32_t incl = sc_op<MAXOP>(sc_op<MAXOP>(carry, wpre), x);
    if (i < n) out[i] = INCLUSIVE ? incl : (incl - v);  // exclusive form is only used with ADD
    carry = sc_op<MAXOP>(carry, total);
    __syncthreads();
  }
}

inline size_t tiles_of(size_t n) { return (n + SC_TILE - 1) / SC_TILE; }

template <bool MAXOP, bool INCLUSIVE>
void scan_rec(const uint32_t *in, uint32_t *out, size_t n, uint32_t seed, uint32_t *temp, cudaStream_t s,
              uint64_t *launches) {
  if (n == 0) return;
  size_t nt = tiles_of(n);
  if (nt == 1) {
    sc_apply<MAXOP, INCLUSIVE><<<1, SC_THREADS, 0, s>>>(in, out, n, nullptr, seed);
    if (launches) *launches += 1;
    return;
  }
  uint32_t *partial = temp;  // nt entries, then deeper levels
  sc_reduce<MAXOP><<<(unsigned)nt, SC_THREADS, 0, s>>>(in, n, partial);
  // exclusive scan of the tile partials (identity 0 for both ops), in place
  scan_rec<MAXOP, false>(partial, partial, nt, 0u, temp + ((nt + 63) & ~(size_t)63), s, launches);
  sc_apply<MAXOP, INCLUSIVE><<<(unsigned)nt, SC_THREADS, 0, s>>>(in, out, n, partial, seed);
  if (launches) *launches += 2;
}

// exclusive MAX scan of partials: out[i] = max(in[0..i-1]).  sc_apply's exclusive branch subtracts, which is
// only right for ADD, so the MAX variant is specialised here.
template <>
void scan_rec<true, false>(const uint32_t *in, uint32_t *out, size_t n, uint32_t seed, uint32_t *temp,
                           cudaStream_t s, uint64_t *launches);

__global__ void shift_right_max(const uint32_t *__restrict__ incl, uint32_t *__restrict__ out, size_t n,
                                uint32_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i == 0 ? seed : (incl[i - 1] > seed ? incl[i - 1] : seed);
}

template <>
void scan_rec<true, false>(const uint32_t *in, uint32_t *out, size_t n, uint32_t seed, uint32_t *temp,
                           cudaStream_t s, uint64_t *launches) {
  // inclusive max scan into a scratch copy, then shift by one
  uint32_t *incl = temp;
  size_t pad = (n + 63) & ~(size_t)63;
  scan_rec<true, true>(in, incl, n, 0u, temp + pad, s, launches);
  shift_right_max<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(incl, out, n, seed);
  if (launches) *launches += 1;
}

}  // namespace

size_t scan_temp_bytes(size_t n) {
  // generous: every level keeps <= n/4096 partials plus (max variant) a scratch copy of them
  size_t words = 0, lvl = tiles_of(n);
  while (true) {
    words += 2 * ((lvl + 63) & ~(size_t)63) + 128;
    if (lvl <= 1) break;
    lvl = tiles_of(lvl);
  }
  return words * sizeof(uint32_t) + 1024;
}

void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, void *temp, cudaStream_t s,
                        uint64_t *launches) {
  scan_rec<false, false>(in, out, n, 0u, (uint32_t *)temp, s, launches);
}

void inclusive_max_scan_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t seed, void *temp,
                            cudaStream_t s, uint64_t *launches) {
  scan_rec<true, true>(in, out, n, seed, (uint32_t *)temp, s, launches);
}

size_t radix_sort_temp_bytes(size_t n) {
  size_t nblocks = (n + RS_TILE - 1) / RS_TILE;
  size_t counts = 256 * nblocks * sizeof(uint32_t);
  counts = (counts + 255) & ~(size_t)255;
  return counts + scan_temp_bytes(256 * nblocks);
}

int radix_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_out, uint32_t *vals_out, size_t n,
                     int key_bits, void *temp, cudaStream_t s, uint64_t *launches) {
  if (n == 0) return 0;
  if (key_bits < 1) key_bits = 1;
  const int passes = (key_bits + 7) / 8;
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  uint32_t *counts = (uint32_t *)temp;
  size_t counts_bytes = ((size_t)256 * nblocks * sizeof(uint32_t) + 255) & ~(size_t)255;
  void *scan_temp = (char *)temp + counts_bytes;
  uint32_t *ki = keys, *vi = vals, *ko = keys_out, *vo = vals_out;
  for (int p = 0; p < passes; p++) {
    const int shift = 8 * p;
    rs_hist<<<nblocks, RS_THREADS, 0, s>>>(ki, n, shift, counts, nblocks);
    exclusive_scan_u32(counts, counts, (size_t)256 * nblocks, scan_temp, s, launches);
    rs_scatter<<<nblocks, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, shift, counts, nblocks);
    if (launches) *launches += 2;
    uint32_t *t;
    t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  return passes & 1;
}
