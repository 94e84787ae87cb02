/* oracle/farms_oracle.c -- TEST INFRASTRUCTURE ONLY (see farms_oracle.h).
 *
 * Sequential CPU restatement of the reference batch path.  Every function cites the reference
 * file:line it follows.  Compile with -O2 -ffp-contract=off (no FMA contraction) so that the FP64
 * operation sequence is exactly the one written here.
 */
#include "farms_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXSTAMP 4294967296.0 /* pow(2,32)  include/vFlow.h:27 */
#define TSTOSEC 1e-6          /*            include/vFlow.h:28 */
#define WINDOW_JUMP 5         /* src/vFlow.cpp:73 */
#define MAX_WINDOW 50         /* src/vFlow.cpp:74 */
#define NSCALES (MAX_WINDOW / WINDOW_JUMP + 1)
#define KILL_OLD_FLOW_TIME 500.0 /* src/vFlow.cpp:961 */

/* SAE cell = reference `Event` {int x,y,pol; double t}; default Event(0,0,0,0)  vFlow.cpp:80 */
typedef struct {
  int x, y;
  double t;
} cell_t;

struct farms_oracle {
  int width, height;
  int frad;       /* vFlow.cpp:36 */
  int plane_size; /* vFlow.cpp:38 */
  int min_evts;   /* vFlow.cpp:40 */
  int have_t0;
  uint32_t t0; /* vFlow.cpp:194 */
  cell_t *sae; /* cSurf, flat [a*height + b]   vFlow.cpp:93, EventMatrix.h:32-34 */
  uint8_t *hit;
  double *last_time; /* lastEventTime          vFlow.cpp:66  */
  double *len;       /* flowSurfaceLengthOn/Of (always written identically, vFlow.cpp:349-353) */
  double *theta;     /* flowSurfaceThetaOn/Of  */
  /* ---- fast pooling mode (farms_oracle_set_fast): a FILTER in front of the same arithmetic ----
   * `active` is a bitmap over the flat index holding a superset of the cells that can pass the test of
   * vFlow.cpp:1002 (len > 0 and |t - lastEventTime| < 500); compute_true_flow_fast visits only those cells, in
   * the reference's own order, applies the same test and adds in the same order, so every sum is bit-identical
   * to compute_true_flow's.  A cell leaves the set when an event without flow overwrites it (vFlow.cpp:398-402)
   * or when its latest event is 500 us old, which is final only while timestamps are non-decreasing: the first
   * decreasing timestamp switches the oracle back to the plain scan for good. */
  /* ---- serial mode (farms_oracle_set_serial): the semantics of vFlowManager::run, vFlow.cpp:465-826 ---- */
  int serial;
  uint64_t events_seen;
  int fast;
  uint32_t tmax;
  uint64_t seq;
  uint64_t *active;
  double *lcos, *lsin; /* len*cos(theta), len*sin(theta) as evaluated at vFlow.cpp:1007-1008 (pure functions) */
  uint64_t *last_seq;  /* sequence number of the latest event of the pixel */
  struct fifo_ent *fifo; /* flow events in stream order, oldest first (ring buffer) */
  size_t fifo_cap, fifo_head, fifo_len;
  uint32_t *hit_i, *hit_j; /* scratch: contributing logical cells of the widest window */
  size_t *hit_f;
};

struct fifo_ent {
  size_t f;
  uint32_t t;
  uint64_t seq;
};

farms_oracle *farms_oracle_create(int width, int height, int filtersize, int inlier_check) {
  if (width <= 0 || height <= 0) return NULL;
  farms_oracle *o = (farms_oracle *)calloc(1, sizeof(*o));
  if (!o) return NULL;
  o->width = width;
  o->height = height;
  /* vFlow.cpp:32-38 */
  if (filtersize < 5) filtersize = 3;
  if (!(filtersize % 2)) filtersize--;
  o->frad = filtersize / 2;
  o->plane_size = filtersize * filtersize;
  o->min_evts = inlier_check;
  size_t npx = (size_t)width * (size_t)height;
  o->sae = (cell_t *)calloc(npx, sizeof(cell_t));
  o->hit = (uint8_t *)calloc(npx, 1);
  o->last_time = (double *)calloc(npx, sizeof(double));
  o->len = (double *)calloc(npx, sizeof(double));
  o->theta = (double *)calloc(npx, sizeof(double));
  if (!o->sae || !o->hit || !o->last_time || !o->len || !o->theta) {
    farms_oracle_destroy(o);
    return NULL;
  }
  return o;
}

void farms_oracle_destroy(farms_oracle *o) {
  if (!o) return;
  free(o->active);
  free(o->lcos);
  free(o->lsin);
  free(o->last_seq);
  free(o->fifo);
  free(o->hit_i);
  free(o->hit_j);
  free(o->hit_f);
  free(o->sae);
  free(o->hit);
  free(o->last_time);
  free(o->len);
  free(o->theta);
  free(o);
}

int farms_oracle_set_fast(farms_oracle *o, int on) {
  if (!o) return -1;
  if (!on) {
    o->fast = 0;
    return 0;
  }
  if (o->seq) return -1; /* only before the first event: the filter is built incrementally */
  const size_t npx = (size_t)o->width * (size_t)o->height;
  const size_t nhit = (size_t)(2 * MAX_WINDOW + 1) * (2 * MAX_WINDOW + 1);
  o->active = (uint64_t *)calloc((npx + 63) / 64 + 1, sizeof(uint64_t));
  o->lcos = (double *)calloc(npx, sizeof(double));
  o->lsin = (double *)calloc(npx, sizeof(double));
  o->last_seq = (uint64_t *)calloc(npx, sizeof(uint64_t));
  o->fifo_cap = 1u << 16;
  o->fifo = (struct fifo_ent *)malloc(o->fifo_cap * sizeof(struct fifo_ent));
  o->hit_i = (uint32_t *)malloc(nhit * sizeof(uint32_t));
  o->hit_j = (uint32_t *)malloc(nhit * sizeof(uint32_t));
  o->hit_f = (size_t *)malloc(nhit * sizeof(size_t));
  if (!o->active || !o->lcos || !o->lsin || !o->last_seq || !o->fifo || !o->hit_i || !o->hit_j || !o->hit_f) return -1;
  o->fast = 1;
  return 0;
}

int farms_oracle_is_fast(const farms_oracle *o) { return o ? o->fast : 0; }

int farms_oracle_set_serial(farms_oracle *o, int on) {
  if (!o || o->events_seen) return -1;
  o->serial = on ? 1 : 0;
  return 0;
}

void farms_oracle_state(const farms_oracle *o, double *last_time, uint8_t *hit, double *len,
                        double *theta) {
  size_t npx = (size_t)o->width * (size_t)o->height;
  if (last_time) memcpy(last_time, o->last_time, npx * sizeof(double));
  if (hit) memcpy(hit, o->hit, npx);
  if (len) memcpy(len, o->len, npx * sizeof(double));
  if (theta) memcpy(theta, o->theta, npx * sizeof(double));
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* Eigen dynamic-size determinant == partialPivLu().determinant(); the operation order is that of
 * oracle/shim/Eigen/Core (which the tier-1 reference build uses).  m is row/col symmetric here but
 * treated generally: m[r][c]. */
static double lu_det3(const double in[3][3]) {
  double lu[3][3];
  int sign = 1;
  memcpy(lu, in, sizeof(lu));
  for (int k = 0; k < 3; k++) {
    int piv = k;
    double best = fabs(lu[k][k]);
    for (int i = k + 1; i < 3; i++) {
      double v = fabs(lu[i][k]);
      if (v > best) {
        best = v;
        piv = i;
      }
    }
    if (best != 0.0) {
      if (piv != k) {
        for (int j = 0; j < 3; j++) {
          double tmp = lu[k][j];
          lu[k][j] = lu[piv][j];
          lu[piv][j] = tmp;
        }
        sign = -sign;
      }
      for (int i = k + 1; i < 3; i++) lu[i][k] = lu[i][k] / lu[k][k];
    }
    for (int i = k + 1; i < 3; i++)
      for (int j = k + 1; j < 3; j++) lu[i][j] = lu[i][j] - lu[i][k] * lu[k][j];
  }
  double prod = lu[0][0];
  prod = prod * lu[1][1];
  prod = prod * lu[2][2];
  return (double)sign * prod;
}

/* computeGrads, both overloads: vFlow.cpp:1214-1238 (build A, Y) and 1241-1381 (solve, inliers).
 * sub = the P gathered cells (cx-major, cy-minor), cen = current event.  Returns inlier count. */
static int compute_grads(const cell_t *sub, int P, const cell_t *cen, double *dtdy, double *dtdx,
                         double *det_out) {
  double ax[169], ay[169], yy[169]; /* P <= 13*13 handled by caller's plane_size check */
  double *AX = ax, *AY = ay, *YY = yy;
  double *heap = NULL;
  if (P > 169) {
    heap = (double *)malloc(sizeof(double) * 3 * (size_t)P);
    AX = heap;
    AY = heap + P;
    YY = heap + 2 * P;
  }
  for (int k = 0; k < P; k++) { /* vFlow.cpp:1224-1234 */
    AX[k] = (double)sub[k].x;
    AY[k] = (double)sub[k].y;
    if (sub[k].t > cen->t)
      YY[k] = (sub[k].t - MAXSTAMP) * TSTOSEC;
    else
      YY[k] = sub[k].t * TSTOSEC;
  }
  const double cx = (double)cen->x, cy = (double)cen->y, cz = cen->t * TSTOSEC; /* :1236-1237 */

  /* AtA = At*A  (vFlow.cpp:1307-1311); entry (i,j) = sum_k At(i,k)*A(k,j), k ascending from 0.0 */
  double ata[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < P; k++) {
        double a = (i == 0) ? AX[k] : (i == 1) ? AY[k] : 1.0;
        double b = (j == 0) ? AX[k] : (j == 1) ? AY[k] : 1.0;
        s = s + a * b;
      }
      ata[i][j] = s;
    }
  double DET = lu_det3(ata); /* vFlow.cpp:1316 */
  if (det_out) *det_out = DET;
  if (DET < 1) { /* vFlow.cpp:1323 */
    free(heap);
    return 0;
  }
  /* column-major data pointer: m[c*3+r] = AtA(r,c)   vFlow.cpp:1315 */
  double m[9], a2d[9];
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) m[c * 3 + r] = ata[r][c];
  DET = 1.0 / DET; /* vFlow.cpp:1327-1336 */
  a2d[0] = DET * (m[8] * m[4] - m[7] * m[5]);
  a2d[1] = DET * (m[7] * m[2] - m[8] * m[1]);
  a2d[2] = DET * (m[5] * m[1] - m[4] * m[2]);
  a2d[3] = DET * (m[6] * m[5] - m[8] * m[3]);
  a2d[4] = DET * (m[8] * m[0] - m[6] * m[2]);
  a2d[5] = DET * (m[3] * m[2] - m[5] * m[0]);
  a2d[6] = DET * (m[7] * m[3] - m[6] * m[4]);
  a2d[7] = DET * (m[6] * m[1] - m[7] * m[0]);
  a2d[8] = DET * (m[4] * m[0] - m[3] * m[1]);
  /* temp = (A2*At)*Y   vFlow.cpp:1338; A2(i,k) = a2d[k*3+i] */
  double abc[3];
  for (int i = 0; i < 3; i++) {
    double acc = 0.0;
    for (int k = 0; k < P; k++) {
      double mik = 0.0;
      mik = mik + a2d[0 * 3 + i] * AX[k];
      mik = mik + a2d[1 * 3 + i] * AY[k];
      mik = mik + a2d[2 * 3 + i] * 1.0;
      acc = acc + mik * YY[k];
    }
    abc[i] = acc;
  }
  double dtdp = sqrt(abc[0] * abc[0] + abc[1] * abc[1]); /* vFlow.cpp:1349 (pow(x,2.0) == x*x) */
  int inliers = 0;
  for (int k = 0; k < P; k++) { /* vFlow.cpp:1352-1369 */
    double planedt = (abc[0] * (AX[k] - cx) + abc[1] * (AY[k] - cy));
    double actualdt = YY[k] - cz;
    if (fabs(planedt - actualdt) < dtdp / 2 && YY[k] > 0) inliers++;
  }
  double speed = 1.0 / dtdp; /* vFlow.cpp:1373-1377 */
  double angle = atan2(abc[0], abc[1]);
  *dtdx = speed * cos(angle);
  *dtdy = speed * sin(angle);
  free(heap);
  return inliers;
}

/* computeLocalFlow  vFlow.cpp:841-949.  Returns (vx, vy); also the diagnostics. */
static void compute_local_flow(const farms_oracle *o, const cell_t *vr, double *vx, double *vy,
                               int *best_window, int *inliers_out, double *det_out) {
  const int r = o->frad, W = o->width, H = o->height;
  double dtdy = 0, dtdx = 0;
  double bestscore = MAXSTAMP + 1; /* :864 */
  int besti = 0, bestj = 0, bestw = -1;
  *vx = 0;
  *vy = 0;
  *inliers_out = 0;
  *det_out = NAN;
  int wi = 0;
  for (int i = vr->x - r; i <= vr->x + r; i += r) {     /* :870 */
    for (int j = vr->y - r; j <= vr->y + r; j += r, wi++) { /* :872 */
      double sobeltsdiff = 0;
      int cnt = 0;
      for (int cx_ = imax(0, i - r); cx_ <= imin(W - 1, i + r); cx_++)
        for (int cy_ = imax(0, j - r); cy_ <= imin(H - 1, j + r); cy_++) cnt++;
      if (cnt < o->plane_size) continue; /* :889 */
      for (int cx_ = imax(0, i - r); cx_ <= imin(W - 1, i + r); cx_++)
        for (int cy_ = imax(0, j - r); cy_ <= imin(H - 1, j + r); cy_++) {
          const cell_t *c = &o->sae[(size_t)cx_ * H + cy_];
          sobeltsdiff += vr->t - c->t;                 /* :894 */
          if (c->t > vr->t) sobeltsdiff += MAXSTAMP;   /* :897-902 */
        }
      sobeltsdiff /= cnt; /* :905 */
      if (sobeltsdiff < bestscore) {
        bestscore = sobeltsdiff;
        besti = i;
        bestj = j;
        bestw = wi;
      }
    }
  }
  *best_window = bestw;
  if (bestscore > MAXSTAMP) return; /* :915-918 */

  cell_t sub[169];
  cell_t *S = sub, *heap = NULL;
  if (o->plane_size > 169) S = heap = (cell_t *)malloc(sizeof(cell_t) * (size_t)o->plane_size);
  int P = 0;
  for (int cx_ = imax(0, besti - r); cx_ <= besti + r; cx_++) /* :923-930 */
    for (int cy_ = imax(0, bestj - r); cy_ <= bestj + r; cy_++) S[P++] = o->sae[(size_t)cx_ * H + cy_];
  int inl = compute_grads(S, P, vr, &dtdy, &dtdx, det_out);
  *inliers_out = inl;
  if (inl >= o->min_evts) { /* :934-939 */
    *vx = dtdx;
    *vy = dtdy;
  }
  free(heap);
}

/* computeTrueFlow(x, y, time, pol)  vFlow.cpp:952-1210.  The pol==1 and else branches are textual
 * mirrors over surfaces that always hold identical values, so one body serves both.
 * Flat-index rule: the reference bounds j by width-1 (vFlow.cpp:1000, 1113) and indexes
 * _data[i*height + j] unchecked (EventMatrix.h:32-34); reads past the end of the vector land in
 * zero padding => "no flow".  Restated explicitly: f = i*H + j; f >= W*H => empty. */
static void compute_true_flow(const farms_oracle *o, int x, int y, uint32_t time_, double *out_vx,
                              double *out_vy, int *out_scale) {
  const int W = o->width, H = o->height;
  const size_t npx = (size_t)W * (size_t)H;
  double pool[NSCALES], vecx[NSCALES], vecy[NSCALES];
  int nwin = 0;
  const double tev = (double)time_;
  for (int s = 0; s <= MAX_WINDOW; s += WINDOW_JUMP) { /* :987 */
    double length_sp = 0, sx = 0, sy = 0, nn = 0;
    for (int i = imax(0, x - s); i <= imin(x + s, W - 1); i++) {   /* :998 */
      for (int j = imax(0, y - s); j <= imin(y + s, W - 1); j++) { /* :1000 (width-1: sic) */
        size_t f = (size_t)i * H + j;
        if (f >= npx) continue;
        double l = o->len[f];
        if (l > 0 && (fabs(tev - o->last_time[f]) < KILL_OLD_FLOW_TIME)) { /* :1002 */
          double th = o->theta[f];
          length_sp = length_sp + l;   /* :1005 */
          sx = sx + l * cos(th);       /* :1007 */
          sy = sy + l * sin(th);       /* :1008 */
          nn++;                        /* :1012 */
        }
      }
    }
    if (nn > 0) { /* :1023-1036 */
      pool[nwin] = length_sp / nn;
      vecx[nwin] = sx / nn;
      vecy[nwin] = sy / nn;
    } else {
      pool[nwin] = 0;
      vecx[nwin] = 0;
      vecy[nwin] = 0;
    }
    nwin++;
  }
  double max_val = 0; /* :1047-1059 */
  int max_idx = 0;
  for (int k = 0; k < nwin; k++)
    if (pool[k] > max_val) {
      max_val = pool[k];
      max_idx = k;
    }
  if (max_val > 0) { /* :1067-1078 */
    *out_vx = vecx[max_idx];
    *out_vy = vecy[max_idx];
    *out_scale = max_idx * WINDOW_JUMP;
  } else { /* :1085-1094 */
    size_t f = (size_t)x * H + y;
    *out_vx = o->len[f] * cos(o->theta[f]);
    *out_vy = o->len[f] * sin(o->theta[f]);
    *out_scale = 0;
  }
}

/* The same function with the `active` filter in front (see struct farms_oracle).  The contributing cells of
 * the widest square are collected once in the reference's enumeration order (i ascending, j ascending,
 * vFlow.cpp:998-1000; logical cells, so an aliased pixel counts once per logical cell that maps to it); every
 * smaller square's cells are a subsequence of that list, so walking the list with the square's bounds performs
 * the reference's additions in the reference's order. */
static void compute_true_flow_fast(farms_oracle *o, int x, int y, uint32_t time_, double *out_vx, double *out_vy,
                                   int *out_scale) {
  const int W = o->width, H = o->height;
  const size_t npx = (size_t)W * (size_t)H;
  const double tev = (double)time_;
  size_t nh = 0;
  const int xlo = imax(0, x - MAX_WINDOW), xhi = imin(x + MAX_WINDOW, W - 1);
  const int jlo = imax(0, y - MAX_WINDOW), jhi = imin(y + MAX_WINDOW, W - 1); /* :1000 (sic) */
  for (int i = xlo; i <= xhi && jlo <= jhi; i++) {
    size_t f0 = (size_t)i * H + jlo, f1 = (size_t)i * H + jhi;
    if (f0 >= npx) break;
    if (f1 >= npx) f1 = npx - 1;
    for (size_t w = f0 >> 6; w <= (f1 >> 6); w++) {
      uint64_t bits = o->active[w];
      if (w == (f0 >> 6)) bits &= ~0ull << (f0 & 63);
      if (w == (f1 >> 6) && (f1 & 63) != 63) bits &= (1ull << ((f1 & 63) + 1)) - 1;
      while (bits) {
        const size_t f = (w << 6) + (size_t)__builtin_ctzll(bits);
        bits &= bits - 1;
        if (o->len[f] > 0 && (fabs(tev - o->last_time[f]) < KILL_OLD_FLOW_TIME)) { /* :1002 */
          o->hit_i[nh] = (uint32_t)i;
          o->hit_j[nh] = (uint32_t)(f - (size_t)i * H);
          o->hit_f[nh] = f;
          nh++;
        }
      }
    }
  }
  double pool[NSCALES], vecx[NSCALES], vecy[NSCALES];
  int nwin = 0;
  for (int s = 0; s <= MAX_WINDOW; s += WINDOW_JUMP) { /* :987 */
    double length_sp = 0, sx = 0, sy = 0, nn = 0;
    const int ia = x - s, ib = x + s, ja = y - s, jb = y + s;
    for (size_t k = 0; k < nh; k++) {
      const int i = (int)o->hit_i[k], j = (int)o->hit_j[k];
      if (i < ia || i > ib || j < ja || j > jb) continue;
      const size_t f = o->hit_f[k];
      length_sp = length_sp + o->len[f]; /* :1005 */
      sx = sx + o->lcos[f];              /* :1007 */
      sy = sy + o->lsin[f];              /* :1008 */
      nn++;                              /* :1012 */
    }
    if (nn > 0) { /* :1023-1036 */
      pool[nwin] = length_sp / nn;
      vecx[nwin] = sx / nn;
      vecy[nwin] = sy / nn;
    } else {
      pool[nwin] = 0;
      vecx[nwin] = 0;
      vecy[nwin] = 0;
    }
    nwin++;
  }
  double max_val = 0; /* :1047-1059 */
  int max_idx = 0;
  for (int k = 0; k < nwin; k++)
    if (pool[k] > max_val) {
      max_val = pool[k];
      max_idx = k;
    }
  if (max_val > 0) { /* :1067-1078 */
    *out_vx = vecx[max_idx];
    *out_vy = vecy[max_idx];
    *out_scale = max_idx * WINDOW_JUMP;
  } else { /* :1085-1094 */
    size_t f = (size_t)x * H + y;
    *out_vx = o->len[f] * cos(o->theta[f]);
    *out_vy = o->len[f] * sin(o->theta[f]);
    *out_scale = 0;
  }
}

/* fast mode bookkeeping for one event whose surfaces were just written */
static void fast_note_event(farms_oracle *o, size_t f, uint32_t time_, int valid) {
  o->seq++;
  o->last_seq[f] = o->seq;
  if (!valid) {
    o->active[f >> 6] &= ~(1ull << (f & 63));
    return;
  }
  o->active[f >> 6] |= 1ull << (f & 63);
  o->lcos[f] = o->len[f] * cos(o->theta[f]);
  o->lsin[f] = o->len[f] * sin(o->theta[f]);
  if (o->fifo_len == o->fifo_cap) { /* grow the ring buffer, oldest entry to slot 0 */
    struct fifo_ent *nf = (struct fifo_ent *)malloc(2 * o->fifo_cap * sizeof(struct fifo_ent));
    for (size_t k = 0; k < o->fifo_len; k++) nf[k] = o->fifo[(o->fifo_head + k) % o->fifo_cap];
    free(o->fifo);
    o->fifo = nf;
    o->fifo_head = 0;
    o->fifo_cap *= 2;
  }
  struct fifo_ent *e = &o->fifo[(o->fifo_head + o->fifo_len) % o->fifo_cap];
  e->f = f;
  e->t = time_;
  e->seq = o->seq;
  o->fifo_len++;
}

/* drop cells whose latest event is >= 500 us older than `now` (final for non-decreasing timestamps) */
static void fast_expire(farms_oracle *o, uint32_t now) {
  while (o->fifo_len) {
    const struct fifo_ent *e = &o->fifo[o->fifo_head];
    if ((double)now - (double)e->t < KILL_OLD_FLOW_TIME) break;
    if (o->last_seq[e->f] == e->seq) o->active[e->f >> 6] &= ~(1ull << (e->f & 63));
    o->fifo_head = (o->fifo_head + 1) % o->fifo_cap;
    o->fifo_len--;
  }
}

/* Loop body of runFileCopy  vFlow.cpp:223-414 */
int farms_oracle_process(farms_oracle *o, const int32_t *X, const int32_t *Y, const uint32_t *T,
                         const int32_t *POL, uint64_t n, const farms_oracle_out *out) {
  const int H = o->height;
  for (uint64_t e = 0; e < n; e++) {
    int x = X[e], y = Y[e];
    if (x < 0 || x >= o->width || y < 0 || y >= o->height) return -1;
    if (!o->have_t0) { /* :194 */
      o->t0 = T[e];
      o->have_t0 = 1;
    }
    if (o->serial && o->events_seen == 0) {
      /* vFlow.cpp:531-558: the first line only sets t0; its pixel's lastEventTime keeps the RAW timestamp, the
       * event enters neither the surface of active events nor the flow surfaces, and no flow is computed */
      o->events_seen = 1;
      o->last_time[(size_t)x * H + y] = (double)T[e];
      if (out) {
        if (out->t_rel) out->t_rel[e] = 0;
        if (out->pol) out->pol[e] = POL[e] < 0 ? 0 : POL[e];
        if (out->global_r) out->global_r[e] = 0;
        if (out->global_theta) out->global_theta[e] = 0;
        if (out->vx) out->vx[e] = 0;
        if (out->vy) out->vy[e] = 0;
        if (out->local_r) out->local_r[e] = 0;
        if (out->local_theta) out->local_theta[e] = 0;
        if (out->scale) out->scale[e] = 0;
        if (out->valid) out->valid[e] = 0;
        if (out->best_window) out->best_window[e] = -1;
        if (out->inliers) out->inliers[e] = 0;
        if (out->det) out->det[e] = NAN;
      }
      continue;
    }
    o->events_seen++;
    uint32_t time_ = T[e] - o->t0; /* :241 / :570 (unsigned wrap) */
    if (o->serial) o->fast = 0; /* the filter assumes lastEventTime is written before pooling */
    if (o->fast) {
      if (time_ < o->tmax) o->fast = 0; /* timestamps decreased: the filter's expiry is no longer final */
      else o->tmax = time_;
    }
    int pol = POL[e];
    if (pol < 0) pol = 0; /* :246 */
    size_t f = (size_t)x * H + y;
    cell_t cur;
    cur.x = x;
    cur.y = y;
    cur.t = (double)time_;
    if (!o->serial) o->last_time[f] = (double)time_; /* :264; the serial loop only does it after pooling (:790) */
    o->sae[f] = cur;                 /* :267 / :595-610 (cSurf = surfaceOfL) */
    o->hit[f] = 1;

    double vx, vy, det;
    int bw, inl;
    compute_local_flow(o, &cur, &vx, &vy, &bw, &inl, &det); /* :304 */

    double gr = 0, gth = 0, lr = 0, lth = 0;
    int scale = 0;
    int valid = (!isnan(fabs(vx)) && !isnan(fabs(vy)) && vx != 0 && vy != 0); /* :315 */
    if (valid) {
      double length = sqrt((vx * vx + vy * vy)); /* :324 */
      double theta = atan2(vy, vx);              /* :325 */
      o->len[f] = length;                        /* :349-353 */
      o->theta[f] = theta;
      double tvx = 0, tvy = 0;
      if (o->fast) {
        fast_note_event(o, f, time_, 1);
        fast_expire(o, time_);
        compute_true_flow_fast(o, x, y, time_, &tvx, &tvy, &scale);
      } else
        compute_true_flow(o, x, y, time_, &tvx, &tvy, &scale); /* :362 */
      gr = sqrt(tvy * tvy + tvx * tvx);                      /* :365 */
      gth = atan2(tvy, tvx);                                 /* :366 */
      lr = length;
      lth = theta;
    } else {
      o->len[f] = 0; /* :398-402 */
      o->theta[f] = 0;
      if (o->fast) fast_note_event(o, f, time_, 0);
    }
    o->last_time[f] = (double)time_; /* :407 */
    if (out) {
      if (out->t_rel) out->t_rel[e] = (int32_t)time_;
      if (out->pol) out->pol[e] = pol;
      if (out->global_r) out->global_r[e] = gr;
      if (out->global_theta) out->global_theta[e] = gth;
      if (out->vx) out->vx[e] = vx;
      if (out->vy) out->vy[e] = vy;
      if (out->local_r) out->local_r[e] = lr;
      if (out->local_theta) out->local_theta[e] = lth;
      if (out->scale) out->scale[e] = scale;
      if (out->valid) out->valid[e] = (uint8_t)valid;
      if (out->best_window) out->best_window[e] = (int8_t)bw;
      if (out->inliers) out->inliers[e] = inl;
      if (out->det) out->det[e] = det;
    }
  }
  return 0;
}
