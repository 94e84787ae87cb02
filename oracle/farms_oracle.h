/* oracle/farms_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Tier-2 oracle: a plain-C, sequential, CPU restatement of the reference's batch path
 * (vFlowManager::runFileCopy -> computeLocalFlow -> computeGrads -> computeTrueFlow;
 * reference src/vFlow.cpp:111-460, 841-949, 1214-1381, 952-1210; include/EventMatrix.h:32-34).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.  The product
 * (libfarms_b200.so) never links, loads or calls it.
 *
 * Pinning: validated byte-for-byte (11-column text) against oracle/_ref/FARMS_Flow, which is the
 * UNMODIFIED reference sources compiled against oracle/shim/ (Eigen3 and Boost headers are absent
 * from the image).  The FP64 operation order inside Eigen (At*A, determinant, A2*At*Y) is therefore
 * the shim's, not real Eigen's: that slice of parity is unpinned (see DESIGN.md "Oracle").
 * The serial mode (farms_oracle_set_serial: the reference's default driver vFlowManager::run, which writes no
 * file) is pinned against what the reference's own computeLocalFlow / computeTrueFlow returned inside run(),
 * recorded by the call probe oracle/serial_probe.cpp (tests/golden/make_golden_serial.py).
 */
#ifndef FARMS_ORACLE_H
#define FARMS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct farms_oracle farms_oracle;

/* Per-event outputs, SoA, caller-allocated, n entries each.  Any pointer may be NULL. */
typedef struct {
  int32_t *t_rel;       /* column 3: (u32)(t - t0) printed as int           vFlow.cpp:241, 373 */
  int32_t *pol;         /* column 4: polarity after clamping <0 to 0        vFlow.cpp:246      */
  double *global_r;     /* column 5                                         vFlow.cpp:365      */
  double *global_theta; /* column 6                                         vFlow.cpp:366      */
  double *vx;           /* column 7 (raw local result, also for invalid)    vFlow.cpp:378, 394 */
  double *vy;           /* column 8                                                            */
  double *local_r;      /* column 9                                         vFlow.cpp:324      */
  double *local_theta;  /* column 10                                        vFlow.cpp:325      */
  int32_t *scale;       /* column 11                                        vFlow.cpp:380      */
  uint8_t *valid;       /* the test at vFlow.cpp:315                                           */
  /* intermediates the reference never prints (for bit-exact GPU checks) */
  int8_t *best_window;  /* 0..8 = 3*di+dj index (i outer, j inner) of the winning window, -1 none */
  int32_t *inliers;     /* return value of computeGrads (0 when DET<1 or no window)            */
  double *det;          /* DET as returned by determinant(); NaN when no window                */
} farms_oracle_out;

/* width/height/filtersize/inlier_check as given on the reference CLI (filtersize is normalised
 * inside exactly like vFlow.cpp:32-36).  Returns NULL on bad arguments. */
farms_oracle *farms_oracle_create(int width, int height, int filtersize, int inlier_check);
void farms_oracle_destroy(farms_oracle *o);

/* Fast pooling mode (call before the first event): computeTrueFlow's 39,611-cell scan is replaced by a walk
 * over a bitmap of the cells that can still pass the test of vFlow.cpp:1002, visiting them in the reference's
 * order and adding in the reference's order, so every output is bit-identical to the plain mode (which stays
 * the slow witness: tests run both).  ~20x faster on dense streams, which is what makes multi-million-event
 * parity runs possible.  Falls back to the plain scan by itself when timestamps decrease.  Returns 0 or -1. */
int farms_oracle_set_fast(farms_oracle *o, int on);
int farms_oracle_is_fast(const farms_oracle *o);

/* Serial mode (call before the first event): the semantics of the reference's DEFAULT driver vFlowManager::run
 * (vFlow.cpp:465-826) instead of runFileCopy.  Differences: the first event only sets t0 -- it is
 * not inserted into the surface of active events and leaves its RAW timestamp in lastEventTime (:531-558); and
 * lastEventTime[x][y] is updated only AFTER pooling (:790), so an event's own pixel is pooled with the time of the
 * PREVIOUS event there (the fallback of :1085-1094 becomes reachable).  The reference writes no file in this
 * mode (:487-489, 727-765), so these are the outputs it computes but never emits: PARITY UNPINNED for this mode
 * (nothing of the reference's can be compared); the restatement shares every function with the batch mode.
 * The `numEvents + 1` loop bound (:565) and the filesize/18 cap (:511) are file-level and live in the CLI. */
int farms_oracle_set_serial(farms_oracle *o, int on);

/* Process n more events in order.  State persists between calls; t0 is the first timestamp ever
 * seen (vFlow.cpp:194).  Returns 0, or -1 if an event lies outside the sensor. */
int farms_oracle_process(farms_oracle *o, const int32_t *x, const int32_t *y, const uint32_t *t,
                         const int32_t *p, uint64_t n, const farms_oracle_out *out);

/* Copy out the per-pixel state (x-major flat index a*height+b like EventMatrix): last event time,
 * hit flag, flow length and theta.  Any pointer may be NULL. */
void farms_oracle_state(const farms_oracle *o, double *last_time, uint8_t *hit, double *len,
                        double *theta);

#ifdef __cplusplus
}
#endif
#endif
