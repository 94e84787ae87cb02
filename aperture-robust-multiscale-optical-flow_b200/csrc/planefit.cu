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

// Specialisation for the usual radii: the (4R+1)^2 footprint is gathered ONCE (column by column, all loads of
// a column in flight together) into shared memory laid out [cell][thread]; window selection, AtA, the solve
// and the inlier loop then read shared memory.  Same arithmetic, operation for operation, as k_plane_fit.
template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS) k_plane_fit_r(const uint2 *__restrict__ sae, const int2 *__restrict__ prevp,
                                                         const uint16_t *__restrict__ ex,
                                                         const uint16_t *__restrict__ ey,
                                                         const uint32_t *__restrict__ et, int i0, int i1,
                                                         FitParams fp, FitOut fo,
                                                         unsigned long long *__restrict__ valid_count) {
  constexpr int N = 4 * R + 1, N1 = 2 * R + 1;
  extern __shared__ uint32_t s_tc[];  // times [N*N][THREADS], then hit flags (bytes) [N*N][THREADS]
  uint32_t *my = s_tc + threadIdx.x;
  uint8_t *myh = reinterpret_cast<uint8_t *>(s_tc + N * N * THREADS) + threadIdx.x;
  const int i = i0 + blockIdx.x * THREADS + threadIdx.x;
  bool valid = false;
  if (i < i1) {
    const int W = fp.W, H = fp.H;
    const int x = ex[i], y = ey[i];
    const uint32_t t = et[i];
    unsigned long long sums[9];
#pragma unroll
    for (int w = 0; w < 9; w++) sums[w] = 0ull;
#pragma unroll 1
    for (int a = 0; a < N; a++) {
      const int cx = x + a - 2 * R;
      unsigned long long c0 = 0, c1 = 0, c2 = 0;
      if (cx >= 0 && cx < W) {
        uint2 cell[N];
#pragma unroll
        for (int b = 0; b < N; b++) {
          const int cy = y + b - 2 * R;
          cell[b] = (cy >= 0 && cy < H) ? sae[cx * H + cy] : make_uint2(0u, (uint32_t)SAE_NEVER);
        }
#pragma unroll
        for (int b = 0; b < N; b++) {
          const int cy = y + b - 2 * R;
          int j = (int)cell[b].y;
          uint32_t tc = cell[b].x;
          while (j > i) {  // later events of this chunk: step back along the pixel's history
            const int2 pp = prevp[j];
            j = pp.x;
            tc = (uint32_t)pp.y;
          }
          my[(a * N + b) * THREADS] = tc;
          myh[(a * N + b) * THREADS] = j != SAE_NEVER;
          if (cy >= 0 && cy < H) {
            const unsigned long long age = (uint32_t)(t - tc);
            if (b <= 2 * R) c0 += age;
            if (b >= R && b <= 3 * R) c1 += age;
            if (b >= 2 * R) c2 += age;
          }
        }
      }
      if (a <= 2 * R) { sums[0] += c0; sums[1] += c1; sums[2] += c2; }
      if (a >= R && a <= 3 * R) { sums[3] += c0; sums[4] += c1; sums[5] += c2; }
      if (a >= 2 * R) { sums[6] += c0; sums[7] += c1; sums[8] += c2; }
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

    double vx = 0.0, vy = 0.0, det = __longlong_as_double(0x7ff8000000000000ll);
    int inliers = 0;
    if (best >= 0) {
      const int oa = (best / 3) * R, ob = (best % 3) * R;  // window origin inside the footprint
      const int bx0 = x - 2 * R + oa, by0 = y - 2 * R + ob;
      long long Sxx = 0, Sxy = 0, Sx = 0, Syy = 0, Sy = 0;
#pragma unroll 1
      for (int a = 0; a < N1; a++)
#pragma unroll
        for (int b = 0; b < N1; b++) {
          const int cidx = (oa + a) * N + ob + b;
          const bool hit = myh[cidx * THREADS] != 0;
          const long long sx = hit ? bx0 + a : 0, sy = hit ? by0 + b : 0;
          Sxx += sx * sx; Sxy += sx * sy; Sx += sx; Syy += sy * sy; Sy += sy;
        }
      const double m00 = (double)Sxx, m01 = (double)Sxy, m02 = (double)Sx, m11 = (double)Syy, m12 = (double)Sy,
                   m22 = (double)(N1 * N1);
      double DET = lu_det3(m00, m01, m02, m01, m11, m12, m02, m12, m22);  // :1316
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
#pragma unroll 1
        for (int a = 0; a < N1; a++)
#pragma unroll
          for (int b = 0; b < N1; b++) {
            const int cidx = (oa + a) * N + ob + b;
            const bool hit = myh[cidx * THREADS] != 0;
            const uint32_t tc = my[cidx * THREADS];
            const double sx = hit ? (double)(bx0 + a) : 0.0, sy = hit ? (double)(by0 + b) : 0.0;
            const double Y = tc > t ? dmul(dsub((double)tc, MAXSTAMP_D), TSTOSEC_D) : dmul((double)tc, TSTOSEC_D);
            const double mk0 = dadd(dadd(dadd(0.0, dmul(A0, sx)), dmul(A3, sy)), A6);
            const double mk1 = dadd(dadd(dadd(0.0, dmul(A1, sx)), dmul(A4, sy)), A7);
            abc0 = dadd(abc0, dmul(mk0, Y));
            abc1 = dadd(abc1, dmul(mk1, Y));
          }
        const double dtdp = __dsqrt_rn(dadd(dmul(abc0, abc0), dmul(abc1, abc1)));  // :1349
        const double half = dmul(dtdp, 0.5);
        const double cxd = (double)x, cyd = (double)y, cz = dmul((double)t, TSTOSEC_D);  // :1236-1237
#pragma unroll 1
        for (int a = 0; a < N1; a++)
#pragma unroll
          for (int b = 0; b < N1; b++) {  // :1352-1369
            const int cidx = (oa + a) * N + ob + b;
            const bool hit = myh[cidx * THREADS] != 0;
            const uint32_t tc = my[cidx * THREADS];
            const double sx = hit ? (double)(bx0 + a) : 0.0, sy = hit ? (double)(by0 + b) : 0.0;
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

template <int R, int THREADS>
void launch_fit_r(const uint2 *sae, const int2 *prevp, const uint16_t *ex, const uint16_t *ey, const uint32_t *et,
                  int i0, int i1, FitParams fp, FitOut fo, unsigned long long *valid_count, cudaStream_t s) {
  constexpr int N = 4 * R + 1;
  constexpr size_t smem = (size_t)N * N * THREADS * (sizeof(uint32_t) + 1);
  auto kern = k_plane_fit_r<R, THREADS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const unsigned grid = (unsigned)((i1 - i0 + THREADS - 1) / THREADS);
  kern<<<grid, THREADS, smem, s>>>(sae, prevp, ex, ey, et, i0, i1, fp, fo, valid_count);
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

void launch_plane_fit(const uint2 *sae, const int2 *prevp, const uint16_t *ex, const uint16_t *ey,
                      const uint32_t *et, int i0, int i1, FitParams fp, FitOut fo, unsigned long long *valid_count,
                      cudaStream_t s) {
  if (i1 <= i0) return;
  if (fp.r == 1) launch_fit_r<1, 128>(sae, prevp, ex, ey, et, i0, i1, fp, fo, valid_count, s);
  else if (fp.r == 2) launch_fit_r<2, 128>(sae, prevp, ex, ey, et, i0, i1, fp, fo, valid_count, s);
  else if (fp.r == 3) launch_fit_r<3, 64>(sae, prevp, ex, ey, et, i0, i1, fp, fo, valid_count, s);
  else k_plane_fit<<<nb((size_t)(i1 - i0), 128), 128, 0, s>>>(sae, prevp, ex, ey, et, i0, i1, fp, fo, valid_count);
}

void launch_sae_export(const uint2 *sae, size_t npx, uint32_t *last_t, uint8_t *hit, cudaStream_t s) {
  k_sae_export<<<nb(npx, 256), 256, 0, s>>>(sae, npx, last_t, hit);
}

void launch_sae_fold(uint2 *sae, size_t npx, const uint32_t *last_t, const uint8_t *hit, cudaStream_t s) {
  k_sae_fold<<<nb(npx, 256), 256, 0, s>>>(sae, npx, last_t, hit);
}
