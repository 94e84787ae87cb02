/* farms_textio.h -- multi-threaded reader/writer for the reference's two text formats (host only, no CUDA).
 *
 * at 10^8 events/s on the GPU the reference's getline/stringstream reader
 * (src/vFlow.cpp:173-188) and its `ofstream << ... << endl` writer (src/vFlow.cpp:436-440) would dominate the
 * FARMS_Flow command line by two orders of magnitude.  Same bytes in, same bytes out, all host cores.
 */
#ifndef FARMS_TEXTIO_H
#define FARMS_TEXTIO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t n;
  uint16_t *x, *y;      /* device-ready SoA (include/farms_b200.h)                                          */
  uint64_t *t;
  int32_t *xi, *yi;     /* echoes for the output rows (the reference prints the ints it parsed)             */
  int32_t *pol;         /* polarity after clamping negatives to 0 (src/vFlow.cpp:246-247)                   */
} farms_events;

/* Reads "<x> <y> <t> <p>" lines like `stream >> x >> y >> time_ >> pol` per getline (src/vFlow.cpp:173-188):
 * at most max_events lines; a field that fails to parse leaves that and all later variables at the previous
 * line's values (blank line => a copy of the previous event).  Returns 0, or -1 with a message in err
 * (cannot open, coordinate outside 0..65535).  Arrays are malloc'ed; release with farms_text_free.
 * nthreads <= 0: hardware concurrency. */
int farms_text_read(const char *path, uint64_t max_events, int nthreads, farms_events *out, char *err, size_t errlen);
void farms_text_free(farms_events *ev);

/* Writes the 11-column batch file `x y t p globalR globalTheta Vx Vy localR localTheta scale`
 * (src/vFlow.cpp:438; ostream default formatting == "%g"; t printed as int like the reference's vector<int>)
 * and, if path8 != NULL, the README's 8-column file (columns 1-6, 9, 10; README.md:63).  Returns 0 or -1. */
int farms_text_write(const char *path11, const char *path8, uint64_t n, const int32_t *xi, const int32_t *yi,
                     const uint32_t *t_rel, const int32_t *pol, const double *global_r, const double *global_theta,
                     const double *vx, const double *vy, const double *local_r, const double *local_theta,
                     const uint8_t *scale, int nthreads);

/* ---- binary side-format (not in the reference) ---------------------------------
 * At 10^9 events the text files are 19 GB in and ~60 GB out; these little-endian SoA files carry the same
 * columns at 13 and 58 bytes per event and load with one read per column.
 *   events  "<name>.evb":  "FARMSEV1", u64 n, then x u16[n], y u16[n], t u64[n], p u8[n]
 *   results "<name>_FARMSOut_.bin": "FARMSOU1", u64 n, then x u16[n], y u16[n], t_rel u32[n], p u8[n], scale u8[n],
 *            globalR, globalTheta, Vx, Vy, localR, localTheta f64[n] each  (the 11 columns of src/vFlow.cpp:438)
 * farms_bin_read fills the same farms_events as the text reader (polarity clamped like src/vFlow.cpp:246-247);
 * both return 0, or -1 with a message in err. */
int farms_bin_read(const char *path, uint64_t max_events, farms_events *out, char *err, size_t errlen);
int farms_bin_write_events(const char *path, uint64_t n, const uint16_t *x, const uint16_t *y, const uint64_t *t,
                           const uint8_t *p);
int farms_bin_write(const char *path, uint64_t n, const int32_t *xi, const int32_t *yi, const uint32_t *t_rel,
                    const int32_t *pol, const double *global_r, const double *global_theta, const double *vx,
                    const double *vy, const double *local_r, const double *local_theta, const uint8_t *scale);
#ifdef __cplusplus
}
#endif
#endif
