/* tools/farms_synth.h -- deterministic synthetic event streams for the five BASELINE.json configs
 * (SURVEY.md section 8(d)).  Bench/test tooling: not part of the reference's interface.
 *
 * A stream is a pure function of (config, seed): events are produced per 1024-us time bucket from
 * counter-based hashes, so any time range can be generated independently (each rank of a multi-GPU
 * run generates its own time slice) and in parallel, and the result never depends on the range
 * boundaries or the thread count.  Events are sorted by time; ties keep generation order. */
#ifndef FARMS_SYNTH_H
#define FARMS_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct farms_synth farms_synth;

/* config 1..5 = BASELINE.json configs[0..4]; seed 0 = the config's default seed. */
farms_synth *farms_synth_open(int config, uint64_t seed);
void farms_synth_close(farms_synth *s);
/* sensor shape, the CLI filtersize of the config, and the approximate event rate (events/s). */
void farms_synth_info(const farms_synth *s, int *width, int *height, int *filtersize, double *rate);

/* All events whose timestamp (microseconds since stream start) lies in [t_begin_us, t_end_us), in
 * order.  Reported timestamps are offset by +1000 (so t0 = 1000-ish like a real recording).
 * Returns the count; if it exceeds cap nothing is guaranteed about the arrays and the NEGATED
 * required count is returned.  nthreads <= 0 picks the hardware concurrency. */
int64_t farms_synth_range(farms_synth *s, uint64_t t_begin_us, uint64_t t_end_us, uint16_t *x,
                          uint16_t *y, uint64_t *t, uint8_t *p, int64_t cap, int nthreads);
#ifdef __cplusplus
}
#endif
#endif
