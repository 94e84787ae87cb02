/* oracle/farms_oracle_cli.c -- TEST INFRASTRUCTURE ONLY.
 * Runs the tier-2 oracle over "<base>.txt" and writes "<base>_FARMSOut_oracle.txt" in the reference's
 * 11-column batch format (vFlow.cpp:436-440; ostream default == "%g"), so it can be diffed against
 * the tier-1 build's "<base>_FARMSOut_batch.txt".
 * usage: farms_oracle_cli <width> <height> <filtersize> <inlierCheck> <base> [numEvents]
 * FARMS_ORACLE_FAST=1 selects the fast pooling mode (farms_oracle_set_fast), FARMS_ORACLE_SERIAL=1 the semantics of
 * the reference's default driver vFlowManager::run (farms_oracle_set_serial; row 0 is the line that only sets t0). */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "farms_oracle.h"

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage: %s W H filtersize inlierCheck base [numEvents]\n", argv[0]); return 2; }
  int W = atoi(argv[1]), H = atoi(argv[2]), fs = atoi(argv[3]), inl = atoi(argv[4]);
  unsigned long long maxn = argc > 6 ? strtoull(argv[6], 0, 10) : ~0ull;
  char path[4096];
  snprintf(path, sizeof path, "%s.txt", argv[5]);
  FILE *f = fopen(path, "r");
  if (!f) { perror(path); return 1; }
  size_t cap = 1 << 20, n = 0;
  int32_t *x = malloc(cap * 4), *y = malloc(cap * 4), *p = malloc(cap * 4);
  uint32_t *t = malloc(cap * 4);
  char line[256];
  while (n < maxn && fgets(line, sizeof line, f)) {
    if (n == cap) { cap *= 2; x = realloc(x, cap * 4); y = realloc(y, cap * 4); p = realloc(p, cap * 4); t = realloc(t, cap * 4); }
    unsigned tt; int xx, yy, pp;
    if (sscanf(line, "%d %d %u %d", &xx, &yy, &tt, &pp) != 4) continue;
    x[n] = xx; y[n] = yy; t[n] = tt; p[n] = pp; n++;
  }
  fclose(f);
  farms_oracle_out o;
  o.t_rel = malloc(n * 4); o.pol = malloc(n * 4); o.scale = malloc(n * 4); o.inliers = malloc(n * 4);
  o.global_r = malloc(n * 8); o.global_theta = malloc(n * 8); o.vx = malloc(n * 8); o.vy = malloc(n * 8);
  o.local_r = malloc(n * 8); o.local_theta = malloc(n * 8); o.det = malloc(n * 8);
  o.valid = malloc(n); o.best_window = malloc(n);
  farms_oracle *orc = farms_oracle_create(W, H, fs, inl);
  if (getenv("FARMS_ORACLE_FAST") && atoi(getenv("FARMS_ORACLE_FAST")) && farms_oracle_set_fast(orc, 1)) return 1;
  if (getenv("FARMS_ORACLE_SERIAL") && atoi(getenv("FARMS_ORACLE_SERIAL")) && farms_oracle_set_serial(orc, 1)) return 1;
  struct timespec a, b;
  clock_gettime(CLOCK_MONOTONIC, &a);
  int rc = farms_oracle_process(orc, x, y, t, p, n, &o);
  clock_gettime(CLOCK_MONOTONIC, &b);
  if (rc) { fprintf(stderr, "event outside sensor\n"); return 1; }
  double sec = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
  fprintf(stderr, "oracle: %zu events in %.3f s = %.1f events/s\n", n, sec, n / sec);
  snprintf(path, sizeof path, "%s_FARMSOut_oracle.txt", argv[5]);
  FILE *g = fopen(path, "w");
  for (size_t i = 0; i < n; i++)
    fprintf(g, "%d %d %d %d %g %g %g %g %g %g %d\n", x[i], y[i], o.t_rel[i], o.pol[i], o.global_r[i],
            o.global_theta[i], o.vx[i], o.vy[i], o.local_r[i], o.local_theta[i], o.scale[i]);
  fclose(g);
  if (getenv("FARMS_ORACLE_DIAG")) {
    snprintf(path, sizeof path, "%s_FARMSOut_oracle_diag.txt", argv[5]);
    g = fopen(path, "w");
    for (size_t i = 0; i < n; i++)
      fprintf(g, "%d %d %d %.17g\n", o.valid[i], o.best_window[i], o.inliers[i], o.det[i]);
    fclose(g);
  }
  return 0;
}
