// planefit.cu -- K3: local plane fit on the surface of active events (SAE).
//
// Replaces computeLocalFlow (src/vFlow.cpp:841-949) and both computeGrads overloads
// (src/vFlow.cpp:1214-1238, 1241-1381).  One thread per event; every event is independent because the
// SAE state "as of event i" is reconstructed from a chunk-end snapshot plus a walk along the per-pixel
// prev links (sae_lookup), instead of replaying the stream.
//
// Bit-exact decisions (window choice, DET<1 gate, inlier count, validity) require the FP64 operation
// sequence of the reference as built for the oracle (oracle/shim/Eigen/Core): every product and sum
// below is an explicit round-to-nearest intrinsic so nothing is contracted into an FMA or reordered.
#include "farms_dev.cuh"

namespace {

FARMS_CHK_DECL

#define MAXSTAMP_D 4294967296.0  // include/vFlow.h:27
#define TSTOSEC_D 1e-6           // include/vFlow.h:28

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// Latest event at pixel q with index <= i: time and whether the pixel was ever hit.
__device__ __forceinline__ uint32_t sae_lookup(const uint2 *__restrict__ sae, const int2 *__restrict__ prevp,
                                               int q, int i, bool &hit) {
  uint2 c = sae[q];
  int j = (int)c.y;
  uint32_t t = c.x;
  while (j > i) {  // events of this chunk that come after i: step back along the pixel's history
    int2 pp = prevp[j];
    j = pp.x;
    t = (uint32_t)pp.y;
  }
  hit = j != SAE_NEVER;
  return t;
}

// Eigen's dynamic-size determinant = partialPivLu().determinant(); operation order of oracle/shim/Eigen/Core.
__device__ __forceinline__ double lu_det3(double a00, double a01, double a02, double a10, double a11, double a12,
                                          double a20, double a21, double a22) {
  double sign = 1.0;
  // k = 0
  {
    int piv = 0;
    double best = fabs(a00);
    if (fabs(a10) > best) { best = fabs(a10); piv = 1; }
    if (fabs(a20) > best) { best = fabs(a20); piv = 2; }
    if (best != 0.0) {
      if (piv == 1) {
        double t;
        t = a00; a00 = a10; a10 = t;
        t = a01; a01 = a11; a11 = t;
        t = a02; a02 = a12; a12 = t;
        sign = -sign;
      } else if (piv == 2) {
        double t;
        t = a00; a00 = a20; a20 = t;
        t = a01; a01 = a21; a21 = t;
        t = a02; a02 = a22; a22 = t;
        sign = -sign;
      }
      a10 = ddiv(a10, a00);
      a20 = ddiv(a20, a00);
    }
    a11 = dsub(a11, dmul(a10, a01));
    a12 = dsub(a12, dmul(a10, a02));
    a21 = dsub(a21, dmul(a20, a01));
    a22 = dsub(a22, dmul(a20, a02));
  }
  // k = 1
  {
    double best = fabs(a11);
    bool swap = fabs(a21) > best;
    if (swap) best = fabs(a21);
    if (best != 0.0) {
      if (swap) {
        double t;
        t = a10; a10 = a20; a20 = t;
        t = a11; a11 = a21; a21 = t;
        t = a12; a12 = a22; a22 = t;
        sign = -sign;
      }
      a21 = ddiv(a21, a11);
    }
    a22 = dsub(a22, dmul(a21, a12));
  }
  double prod = dmul(dmul(a00, a11), a22);
  return dmul(sign, prod);
}

struct Cell {
  double sx, sy, Y;  // stored coordinates (0,0 for a never-hit cell: src/vFlow.cpp:80, 1226-1227) and time
};

__device__ __forceinline__ Cell load_cell(const uint2 *__restrict__ sae, const int2 *__restrict__ prevp, int cx,
                                          int cy, int H, int i, uint32_t t) {
  bool hit;
  uint32_t tc = sae_lookup(sae, prevp, cx * H + cy, i, hit);
  Cell c;
  c.sx = hit ? (double)cx : 0.0;
  c.sy = hit ? (double)cy : 0.0;
  // src/vFlow.cpp:1229-1233
  c.Y = tc > t ? dmul(dsub((double)tc, MAXSTAMP_D), TSTOSEC_D) : dmul((double)tc, TSTOSEC_D);
  return c;
}

__global__ void __launch_bounds__(128) k_plane_fit(const uint2 *__restrict__ sae, const int2 *__restrict__ prevp,
                                                   const uint16_t *__restrict__ ex, const uint16_t *__restrict__ ey,
                                                   const uint32_t *__restrict__ et, int i0, int i1, FitParams fp,
                                                   FitOut fo, unsigned long long *__restrict__ valid_count) {
  const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = false;
  if (i < i1) {
    const int W = fp.W, H = fp.H, r = fp.r;
    const int x = ex[i], y = ey[i];
    const uint32_t t = et[i];

    // ---- candidate windows: sums of ages over the 9 shifted (2r+1)^2 windows (src/vFlow.cpp:870-912).
    // age = (t - tc) + (tc > t ? 2^32 : 0) == unsigned 32-bit wrap-around difference; all candidates that
    // survive have the same cell count, so arg-min of the mean == arg-min of the exact integer sum.
    unsigned long long sums[9];
#pragma unroll
    for (int w = 0; w < 9; w++) sums[w] = 0ull;
    for (int a = -2 * r; a <= 2 * r; a++) {
      const int cx = x + a;
      if (cx < 0 || cx >= W) continue;  // such a column only belongs to windows that are rejected below
      unsigned long long c0 = 0, c1 = 0, c2 = 0;
      for (int b = -2 * r; b <= 2 * r; b++) {
        const int cy = y + b;
        if (cy < 0 || cy >= H) continue;
        bool hit;
        const uint32_t tc = sae_lookup(sae, prevp, cx * H + cy, i, hit);
        const unsigned long long age = (uint32_t)(t - tc);
        if (b <= 0) c0 += age;
        if (b >= -r && b <= r) c1 += age;
        if (b >= 0) c2 += age;
      }
      if (a <= 0) { sums[0] += c0; sums[1] += c1; sums[2] += c2; }
      if (a >= -r && a <= r) { sums[3] += c0; sums[4] += c1; sums[5] += c2; }
      if (a >= 0) { sums[6] += c0; sums[7] += c1; sums[8] += c2; }
    }
    int best = -1;
    unsigned long long bestsum = ~0ull;
#pragma unroll
    for (int w = 0; w < 9; w++) {
      const int di = w / 3 - 1, dj = w % 3 - 1;
      const int wx = x + di * r, wy = y + dj * r;
      const bool inb = wx - r >= 0 && wx + r <= W - 1 && wy - r >= 0 && wy + r <= H - 1;  // :889
      if (inb && sums[w] < bestsum) {  // strict '<' keeps the first minimum (:906)
        bestsum = sums[w];
        best = w;
      }
    }

    double vx = 0.0, vy = 0.0, det = __longlong_as_double(0x7ff8000000000000ll);
    int inliers = 0;
    if (best >= 0) {
      const int bx0 = x + (best / 3 - 1) * r - r, by0 = y + (best % 3 - 1) * r - r;
      const int n1 = 2 * r + 1;
      // ---- AtA: exact integer sums (src/vFlow.cpp:1307-1311) ----
      long long Sxx = 0, Sxy = 0, Sx = 0, Syy = 0, Sy = 0, Sn = 0;
      for (int a = 0; a < n1; a++)
        for (int b = 0; b < n1; b++) {
          bool hit;
          (void)sae_lookup(sae, prevp, (bx0 + a) * H + (by0 + b), i, hit);
          const long long sx = hit ? bx0 + a : 0, sy = hit ? by0 + b : 0;
          Sxx += sx * sx; Sxy += sx * sy; Sx += sx; Syy += sy * sy; Sy += sy; Sn += 1;
        }
      const double m00 = (double)Sxx, m01 = (double)Sxy, m02 = (double)Sx, m11 = (double)Syy, m12 = (double)Sy,
                   m22 = (double)Sn;
      double DET = lu_det3(m00, m01, m02, m01, m11, m12, m02, m12, m22);  // :1316
      det = DET;
      if (!(DET < 1)) {  // :1323
        // column-major data pointer of the symmetric AtA: d[c*3+r]
        const double d0 = m00, d1 = m01, d2 = m02, d3 = m01, d4 = m11, d5 = m12, d6 = m02, d7 = m12, d8 = m22;
        DET = ddiv(1.0, DET);  // :1327-1336
        const double A0 = dmul(DET, dsub(dmul(d8, d4), dmul(d7, d5)));
        const double A1 = dmul(DET, dsub(dmul(d7, d2), dmul(d8, d1)));
        const double A3 = dmul(DET, dsub(dmul(d6, d5), dmul(d8, d3)));
        const double A4 = dmul(DET, dsub(dmul(d8, d0), dmul(d6, d2)));
        const double A6 = dmul(DET, dsub(dmul(d7, d3), dmul(d6, d4)));
        const double A7 = dmul(DET, dsub(dmul(d6, d1), dmul(d7, d0)));
        // abc = (A2*At)*Y, rows 0 and 1 (:1338); A2(i,k) = A[k*3+i]
        double abc0 = 0.0, abc1 = 0.0;
        for (int a = 0; a < n1; a++)
          for (int b = 0; b < n1; b++) {
            const Cell c = load_cell(sae, prevp, bx0 + a, by0 + b, H, i, t);
            const double mk0 = dadd(dadd(dadd(0.0, dmul(A0, c.sx)), dmul(A3, c.sy)), A6);
            const double mk1 = dadd(dadd(dadd(0.0, dmul(A1, c.sx)), dmul(A4, c.sy)), A7);
            abc0 = dadd(abc0, dmul(mk0, c.Y));
            abc1 = dadd(abc1, dmul(mk1, c.Y));
          }
        const double dtdp = __dsqrt_rn(dadd(dmul(abc0, abc0), dmul(abc1, abc1)));  // :1349
        const double half = dmul(dtdp, 0.5);
        const double cxd = (double)x, cyd = (double)y, cz = dmul((double)t, TSTOSEC_D);  // :1236-1237
        for (int a = 0; a < n1; a++)
          for (int b = 0; b < n1; b++) {  // :1352-1369
            const Cell c = load_cell(sae, prevp, bx0 + a, by0 + b, H, i, t);
            const double planedt = dadd(dmul(abc0, dsub(c.sx, cxd)), dmul(abc1, dsub(c.sy, cyd)));
            const double actualdt = dsub(c.Y, cz);
            if (fabs(dsub(planedt, actualdt)) < half && c.Y > 0) inliers++;
          }
        if (inliers >= fp.min_inl) {  // :934-939
          const double speed = ddiv(1.0, dtdp);  // :1373-1377
          const double angle = atan2(abc0, abc1);
          vx = dmul(speed, cos(angle));
          vy = dmul(speed, sin(angle));
        }
      }
    }

    valid = !isnan(vx) && !isnan(vy) && vx != 0.0 && vy != 0.0;  // src/vFlow.cpp:315
    double len = 0.0, theta = 0.0, lcx = 0.0, lcy = 0.0;
    if (valid) {
      len = __dsqrt_rn(dadd(dmul(vx, vx), dmul(vy, vy)));  // :324
      theta = atan2(vy, vx);                                 // :325
      lcx = dmul(len, cos(theta));                           // :1007
      lcy = dmul(len, sin(theta));                           // :1008
    }
    fo.vx[i] = vx;
    fo.vy[i] = vy;
    fo.len[i] = len;
    fo.theta[i] = theta;
    fo.lcx[i] = lcx;
    fo.lcy[i] = lcy;
    fo.valid[i] = valid ? 1 : 0;
    fo.best_window[i] = (int8_t)best;
    fo.inliers[i] = (uint16_t)inliers;
    if (fo.det) fo.det[i] = det;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(valid_count, (unsigned long long)__popc(bal));
}

// ---------------------------------------------------------------------------------------------------
// Specialisation for the usual radii (filtersize 3, 5, 7): two kernels per chunk.
//
// k_fit_gather: 16 lanes per event, lane = row of the (4R+1)^2 footprint.  The surface is x-major, so the cells
// of one footprint column are contiguous in memory and the lanes of a group read them with one or two
// 128-byte lines per column (a thread-per-event gather issues 81 fully scattered 8-byte loads and saturates
// the L1 wavefront pipe).  Row sums -> prefix over lanes -> the 9 window sums -> best window; the winning
// window's cells (time, hit) go to a small per-event record that stays in L2.
// k_fit_solve: one thread per event reads its record with 16-byte loads and runs the FP64 algebra.
// Same arithmetic, operation for operation, as k_plane_fit.
// ---------------------------------------------------------------------------------------------------
template <int R>
struct FitRec {
  static constexpr int N = 4 * R + 1, N1 = 2 * R + 1, P = N1 * N1;
  static constexpr int WORDS = (P + 3) <= 16 ? 16 : (P + 3) <= 32 ? 32 : 64;  // t[P], hit mask lo/hi, best
  // lanes per event in k_fit_gather (>= N footprint rows) and events per warp: 8 x 4, 10 x 3 (two lanes idle), 16 x 2
  static constexpr int G = R == 1 ? 8 : R == 2 ? 10 : 16, EPW = 32 / G;
};

#ifndef FARMS_GATHER_MINB
#define FARMS_GATHER_MINB 1
#endif
#ifndef FARMS_SOLVE_MINB
#define FARMS_SOLVE_MINB 4  // 128 registers (196 bytes of spills) instead of 168: 16 warps per SM, fit stage -6 %
#endif
template <int R>
__global__ void __launch_bounds__(256, FARMS_GATHER_MINB) k_fit_gather(const uint2 *__restrict__ sae, const int2 *__restrict__ prevp,
                                                    const uint16_t *__restrict__ ex, const uint16_t *__restrict__ ey,
                                                    const uint32_t *__restrict__ et, int i0, int i1, int W, int H,
                                                    uint32_t *__restrict__ recs) {
  constexpr int N = FitRec<R>::N, N1 = FitRec<R>::N1, P = FitRec<R>::P, WORDS = FitRec<R>::WORDS;
  constexpr int G = FitRec<R>::G, EPW = FitRec<R>::EPW;
  const int lane = threadIdx.x & 31;
  const int grp = lane / G, gbase = grp * G;
  const int b = lane - gbase;  // footprint row handled by this lane
  const int gid = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * EPW + grp;
  const int i = i0 + gid;
  const bool live = grp < EPW && i < i1;  // uniform per group of G lanes
  const int ii = live ? i : i1 - 1;
  const int x = ex[ii], y = ey[ii];
  const uint32_t t = et[ii];
  const int cy = y + b - 2 * R;
  const bool row_ok = live && b < N && cy >= 0 && cy < H;
  uint32_t tc[N];
  uint32_t hit = 0;
  unsigned long long r0 = 0, r1 = 0, r2 = 0;
  {
    uint2 cell[N];
#pragma unroll
    for (int a = 0; a < N; a++) {
      const int cx = x + a - 2 * R;
      cell[a] = (row_ok && cx >= 0 && cx < W) ? sae[cx * H + cy] : make_uint2(0u, (uint32_t)SAE_NEVER);
    }
#pragma unroll
    for (int a = 0; a < N; a++) {
      const int cx = x + a - 2 * R;
      int j = (int)cell[a].y;
      uint32_t tv = cell[a].x;
      while (j > i) {  // later events of this chunk: step back along the pixel's history
        if (!FARMS_CHK(j < i1, 201)) break;  // a surface index beyond the chunk would be a stale snapshot
        const int2 pp = prevp[j];
        j = pp.x;
        tv = (uint32_t)pp.y;
      }
      tc[a] = tv;
      if (j != SAE_NEVER) hit |= 1u << a;
      if (row_ok && cx >= 0 && cx < W) {
        const unsigned long long age = (uint32_t)(t - tv);  // src/vFlow.cpp:894-902 as a 32-bit wrap-around
        if (a <= 2 * R) r0 += age;
        if (a >= R && a <= 3 * R) r1 += age;
        if (a >= 2 * R) r2 += age;
      }
    }
  }
  // inclusive prefix over the rows (lanes) of the group
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    const unsigned long long p0 = __shfl_up_sync(0xffffffffu, r0, o), p1 = __shfl_up_sync(0xffffffffu, r1, o),
                             p2 = __shfl_up_sync(0xffffffffu, r2, o);
    if (b >= o) {
      r0 += p0;
      r1 += p1;
      r2 += p2;
    }
  }
  // window (di, dj): columns by di (r0/r1/r2), rows [dj*R, dj*R + 2R]
  unsigned long long sums[9];
#pragma unroll
  for (int dj = 0; dj < 3; dj++) {
    const int hi = gbase + dj * R + 2 * R, lo = dj * R - 1;
    const unsigned long long h0 = __shfl_sync(0xffffffffu, r0, hi), h1 = __shfl_sync(0xffffffffu, r1, hi),
                             h2 = __shfl_sync(0xffffffffu, r2, hi);
    unsigned long long l0 = 0, l1 = 0, l2 = 0;
    if (lo >= 0) {
      l0 = __shfl_sync(0xffffffffu, r0, gbase + lo);
      l1 = __shfl_sync(0xffffffffu, r1, gbase + lo);
      l2 = __shfl_sync(0xffffffffu, r2, gbase + lo);
    }
    sums[0 * 3 + dj] = h0 - l0;
    sums[1 * 3 + dj] = h1 - l1;
    sums[2 * 3 + dj] = h2 - l2;
  }
  int best = -1;
  unsigned long long bestsum = ~0ull;
#pragma unroll
  for (int w = 0; w < 9; w++) {
    const int di = w / 3 - 1, dj = w % 3 - 1;
    const int wx = x + di * R, wy = y + dj * R;
    const bool inb = wx - R >= 0 && wx + R <= W - 1 && wy - R >= 0 && wy + R <= H - 1;  // src/vFlow.cpp:889
    if (inb && sums[w] < bestsum) {  // strict '<' keeps the first minimum (:906)
      bestsum = sums[w];
      best = w;
    }
  }
  // the winning window's cells, cx-major / cy-minor like the reference's regather (:923-930)
  const int oa = best >= 0 ? (best / 3) * R : 0, ob = best >= 0 ? (best % 3) * R : 0;
  const int brel = b - ob;
  const bool mine = live && best >= 0 && brel >= 0 && brel < N1;
  uint32_t *rec = recs + (size_t)gid * WORDS;
  (void)FARMS_CHK(!live || (gid >= 0 && i0 + gid < i1), 202);
  unsigned long long hm = 0ull;
#pragma unroll
  for (int k = 0; k < N1; k++) {
    // a = oa + k with oa in {0, R, 2R}: static selects keep tc[] in registers
    const uint32_t v = oa == 0 ? tc[k] : (oa == R ? tc[k + R] : tc[k + 2 * R]);
    const uint32_t hb = oa == 0 ? (hit >> k) : (oa == R ? (hit >> (k + R)) : (hit >> (k + 2 * R)));
    if (mine) {
      rec[k * N1 + brel] = v;
      if (hb & 1u) hm |= 1ull << (k * N1 + brel);
    }
  }
  // OR of the group's hit bits: inclusive OR-scan over the lanes of the group, total from its last lane
  uint32_t hlo = (uint32_t)hm, hhi = (uint32_t)(hm >> 32);
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    const uint32_t plo = __shfl_up_sync(0xffffffffu, hlo, o), phi = __shfl_up_sync(0xffffffffu, hhi, o);
    if (b >= o) {
      hlo |= plo;
      hhi |= phi;
    }
  }
  hlo = __shfl_sync(0xffffffffu, hlo, gbase + G - 1);
  hhi = __shfl_sync(0xffffffffu, hhi, gbase + G - 1);
  if (live && b == 0) {
    rec[P] = hlo;
    rec[P + 1] = hhi;
    rec[P + 2] = (uint32_t)best;
  }
}

template <int R>
__global__ void __launch_bounds__(128, FARMS_SOLVE_MINB) k_fit_solve(const uint32_t *__restrict__ recs, const uint16_t *__restrict__ ex,
                                                   const uint16_t *__restrict__ ey, const uint32_t *__restrict__ et,
                                                   int i0, int i1, FitParams fp, FitOut fo,
                                                   unsigned long long *__restrict__ valid_count) {
  constexpr int N1 = FitRec<R>::N1, P = FitRec<R>::P, WORDS = FitRec<R>::WORDS;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = i0 + g;
  bool valid = false;
  if (i < i1) {
    uint32_t wd[WORDS];
    const uint4 *rp = reinterpret_cast<const uint4 *>(recs + (size_t)g * WORDS);
#pragma unroll
    for (int q = 0; q < (P + 3 + 3) / 4; q++) {
      const uint4 v = rp[q];
      wd[4 * q] = v.x; wd[4 * q + 1] = v.y; wd[4 * q + 2] = v.z; wd[4 * q + 3] = v.w;
    }
    const unsigned long long hm = (unsigned long long)wd[P] | ((unsigned long long)wd[P + 1] << 32);
    const int best = (int)wd[P + 2];
    const int x = ex[i], y = ey[i];
    const uint32_t t = et[i];
    double vx = 0.0, vy = 0.0, det = __longlong_as_double(0x7ff8000000000000ll);
    int inliers = 0;
    if (best >= 0) {
      const int bx0 = x + (best / 3 - 1) * R - R, by0 = y + (best % 3 - 1) * R - R;
      long long Sxx = 0, Sxy = 0, Sx = 0, Syy = 0, Sy = 0;
#pragma unroll
      for (int k = 0; k < P; k++) {
        const bool hit = (hm >> k) & 1ull;
        const long long sx = hit ? bx0 + k / N1 : 0, sy = hit ? by0 + k % N1 : 0;
        Sxx += sx * sx; Sxy += sx * sy; Sx += sx; Syy += sy * sy; Sy += sy;
      }
      const double m00 = (double)Sxx, m01 = (double)Sxy, m02 = (double)Sx, m11 = (double)Syy, m12 = (double)Sy,
                   m22 = (double)P;
      double DET = lu_det3(m00, m01, m02, m01, m11, m12, m02, m12, m22);  // src/vFlow.cpp:1316
      det = DET;
      if (!(DET < 1)) {  // :1323
        const double d0 = m00, d1 = m01, d2 = m02, d3 = m01, d4 = m11, d5 = m12, d6 = m02, d7 = m12, d8 = m22;
        DET = ddiv(1.0, DET);  // :1327-1336
        const double A0 = dmul(DET, dsub(dmul(d8, d4), dmul(d7, d5)));
        const double A1 = dmul(DET, dsub(dmul(d7, d2), dmul(d8, d1)));
        const double A3 = dmul(DET, dsub(dmul(d6, d5), dmul(d8, d3)));
        const double A4 = dmul(DET, dsub(dmul(d8, d0), dmul(d6, d2)));
        const double A6 = dmul(DET, dsub(dmul(d7, d3), dmul(d6, d4)));
        const double A7 = dmul(DET, dsub(dmul(d6, d1), dmul(d7, d0)));
        double abc0 = 0.0, abc1 = 0.0;
#pragma unroll
        for (int k = 0; k < P; k++) {  // (A2*At)*Y, left to right (:1338)
          const bool hit = (hm >> k) & 1ull;
          const uint32_t tc = wd[k];
          const double sx = hit ? (double)(bx0 + k / N1) : 0.0, sy = hit ? (double)(by0 + k % N1) : 0.0;
          const double Y = tc > t ? dmul(dsub((double)tc, MAXSTAMP_D), TSTOSEC_D) : dmul((double)tc, TSTOSEC_D);
          const double mk0 = dadd(dadd(dadd(0.0, dmul(A0, sx)), dmul(A3, sy)), A6);
          const double mk1 = dadd(dadd(dadd(0.0, dmul(A1, sx)), dmul(A4, sy)), A7);
          abc0 = dadd(abc0, dmul(mk0, Y));
          abc1 = dadd(abc1, dmul(mk1, Y));
        }
        const double dtdp = __dsqrt_rn(dadd(dmul(abc0, abc0), dmul(abc1, abc1)));  // :1349
        const double half = dmul(dtdp, 0.5);
        const double cxd = (double)x, cyd = (double)y, cz = dmul((double)t, TSTOSEC_D);  // :1236-1237
#pragma unroll
        for (int k = 0; k < P; k++) {  // :1352-1369
          const bool hit = (hm >> k) & 1ull;
          const uint32_t tc = wd[k];
          const double sx = hit ? (double)(bx0 + k / N1) : 0.0, sy = hit ? (double)(by0 + k % N1) : 0.0;
          const double Y = tc > t ? dmul(dsub((double)tc, MAXSTAMP_D), TSTOSEC_D) : dmul((double)tc, TSTOSEC_D);
          const double planedt = dadd(dmul(abc0, dsub(sx, cxd)), dmul(abc1, dsub(sy, cyd)));
          const double actualdt = dsub(Y, cz);
          if (fabs(dsub(planedt, actualdt)) < half && Y > 0) inliers++;
        }
        if (inliers >= fp.min_inl) {  // :934-939
          const double speed = ddiv(1.0, dtdp);  // :1373-1377
          const double angle = atan2(abc0, abc1);
          vx = dmul(speed, cos(angle));
          vy = dmul(speed, sin(angle));
        }
      }
    }
    valid = !isnan(vx) && !isnan(vy) && vx != 0.0 && vy != 0.0;  // src/vFlow.cpp:315
    double len = 0.0, theta = 0.0, lcx = 0.0, lcy = 0.0;
    if (valid) {
      len = __dsqrt_rn(dadd(dmul(vx, vx), dmul(vy, vy)));  // :324
      theta = atan2(vy, vx);                                 // :325
      lcx = dmul(len, cos(theta));                           // :1007
      lcy = dmul(len, sin(theta));                           // :1008
    }
    fo.vx[i] = vx;
    fo.vy[i] = vy;
    fo.len[i] = len;
    fo.theta[i] = theta;
    fo.lcx[i] = lcx;
    fo.lcy[i] = lcy;
    fo.valid[i] = valid ? 1 : 0;
    fo.best_window[i] = (int8_t)best;
    fo.inliers[i] = (uint16_t)inliers;
    if (fo.det) fo.det[i] = det;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(valid_count, (unsigned long long)__popc(bal));
}

template <int R>
void launch_fit_r(const uint2 *sae, const int2 *prevp, const uint16_t *ex, const uint16_t *ey, const uint32_t *et,
                  int i0, int i1, FitParams fp, FitOut fo, unsigned long long *valid_count, uint32_t *recs,
                  cudaStream_t s) {
  const size_t n = (size_t)(i1 - i0);
  constexpr int EPB = FitRec<R>::EPW * 8;  // events per 256-thread block
  k_fit_gather<R><<<(unsigned)((n + EPB - 1) / EPB), 256, 0, s>>>(sae, prevp, ex, ey, et, i0, i1, fp.W, fp.H, recs);
  k_fit_solve<R><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(recs, ex, ey, et, i0, i1, fp, fo, valid_count);
}

__global__ void k_sae_init(uint2 *sae, size_t npx) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < npx) sae[q] = make_uint2(0u, (uint32_t)SAE_NEVER);
}

// After this kernel the surface holds the state at the END of chunk [c0, c1): exactly one event per touched
// pixel is the last of the chunk, so no atomics are needed.
__global__ void k_sae_advance(uint2 *__restrict__ sae, const uint32_t *__restrict__ pix,
                              const uint32_t *__restrict__ et, const int32_t *__restrict__ nextp, int c0, int c1) {
  int j = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j < c1 && nextp[j] >= c1) sae[pix[j]] = make_uint2(et[j], (uint32_t)j);
}

// Batch-local indices become meaningless once the batch is done: mark touched pixels as "old".
__global__ void k_sae_finalize(uint2 *__restrict__ sae, const uint32_t *__restrict__ pix,
                               const int32_t *__restrict__ nextp, int m) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m && nextp[j] == NEXT_NONE) sae[pix[j]].y = (uint32_t)SAE_OLD;
}

__global__ void k_sae_export(const uint2 *__restrict__ sae, size_t npx, uint32_t *__restrict__ last_t,
                             uint8_t *__restrict__ hit) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npx) return;
  uint2 c = sae[q];
  last_t[q] = c.x;
  hit[q] = (int)c.y != SAE_NEVER;
}

__global__ void k_sae_fold(uint2 *__restrict__ sae, size_t npx, const uint32_t *__restrict__ last_t,
                           const uint8_t *__restrict__ hit) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < npx && hit[q]) sae[q] = make_uint2(last_t[q], (uint32_t)SAE_OLD);
}

inline unsigned nb(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void launch_sae_init(uint2 *sae, size_t npx, cudaStream_t s) { k_sae_init<<<nb(npx, 256), 256, 0, s>>>(sae, npx); }

void launch_sae_advance(uint2 *sae, const uint32_t *pix, const uint32_t *et, const int32_t *nextp, int c0, int c1,
                        cudaStream_t s) {
  if (c1 > c0) k_sae_advance<<<nb((size_t)(c1 - c0), 256), 256, 0, s>>>(sae, pix, et, nextp, c0, c1);
}

void launch_sae_finalize(uint2 *sae, const uint32_t *pix, const int32_t *nextp, int m, cudaStream_t s) {
  if (m > 0) k_sae_finalize<<<nb((size_t)m, 256), 256, 0, s>>>(sae, pix, nextp, m);
}

size_t plane_fit_scratch_bytes(int r, size_t chunk_events) {
  const size_t words = r == 1 ? FitRec<1>::WORDS : r == 2 ? FitRec<2>::WORDS : r == 3 ? FitRec<3>::WORDS : 0;
  return words * sizeof(uint32_t) * chunk_events + 64;
}

// scratch: plane_fit_scratch_bytes(fp.r, i1 - i0) bytes.  Returns the number of kernels launched.
int launch_plane_fit(const uint2 *sae, const int2 *prevp, const uint16_t *ex, const uint16_t *ey,
                     const uint32_t *et, int i0, int i1, FitParams fp, FitOut fo, unsigned long long *valid_count,
                     void *scratch, cudaStream_t s) {
  if (i1 <= i0) return 0;
  uint32_t *recs = (uint32_t *)scratch;
  if (fp.r == 1) launch_fit_r<1>(sae, prevp, ex, ey, et, i0, i1, fp, fo, valid_count, recs, s);
  else if (fp.r == 2) launch_fit_r<2>(sae, prevp, ex, ey, et, i0, i1, fp, fo, valid_count, recs, s);
  else if (fp.r == 3) launch_fit_r<3>(sae, prevp, ex, ey, et, i0, i1, fp, fo, valid_count, recs, s);
  else {
    k_plane_fit<<<nb((size_t)(i1 - i0), 128), 128, 0, s>>>(sae, prevp, ex, ey, et, i0, i1, fp, fo, valid_count);
    return 1;
  }
  return 2;
}

void launch_sae_export(const uint2 *sae, size_t npx, uint32_t *last_t, uint8_t *hit, cudaStream_t s) {
  k_sae_export<<<nb(npx, 256), 256, 0, s>>>(sae, npx, last_t, hit);
}

void launch_sae_fold(uint2 *sae, size_t npx, const uint32_t *last_t, const uint8_t *hit, cudaStream_t s) {
  k_sae_fold<<<nb(npx, 256), 256, 0, s>>>(sae, npx, last_t, hit);
}

unsigned int farms_chk_planefit(cudaStream_t s) {
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
