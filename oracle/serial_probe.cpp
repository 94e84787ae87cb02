// serial_probe.cpp -- TEST INFRASTRUCTURE ONLY (oracle/): makes the reference's DEFAULT driver, and the intermediate
// results of its plane fit, observable.
//
// vFlowManager::run (src/vFlow.cpp:465-826) computes a local and a pooled flow per event and then throws both away:
// its file writes are commented out (:727-765), so the unmodified reference leaves nothing to compare the serial
// semantics with.  This probe does not touch the reference's sources.  `make -C oracle refserial` compiles them,
// where they lie, into a shared library (position-independent code: run()'s calls to the two flow functions go
// through the PLT), and links this file in front of it.  The two definitions below have the reference's own member
// signatures (include/vFlow.h:84-91), so the dynamic linker binds the reference's calls here; each forwards to the
// reference's implementation (dlsym RTLD_NEXT) and appends what it returned to the file named by
// FARMS_SERIAL_PROBE_OUT:
//     G <inliers> <window> <dtdy> <dtdx>           one line per computeGrads(subsurf, cen, ..)   (:928)
//     L <vx> <vy>                                  one line per computeLocalFlow() call          (:629, :304)
//     T <x> <y> <time> <pol> <vx> <vy> <scale>     one line per computeTrueFlow(x,y,time,pol)    (:703, :362)
// tests/golden/make_golden_serial.py turns that log into the golden vectors of the oracle's serial mode and -- run
// under the batch driver -- of the plane fit's intermediate results (inlier count, winning window), which the
// reference's output files do not carry.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>

#include "vFlow.h"

namespace {

FILE *probe_file() {
  static FILE *f = [] {
    const char *path = getenv("FARMS_SERIAL_PROBE_OUT");
    FILE *g = path ? fopen(path, "w") : nullptr;
    if (!g) {
      fprintf(stderr, "serial_probe: set FARMS_SERIAL_PROBE_OUT to a writable path\n");
      abort();
    }
    atexit([] { fclose(probe_file()); });
    return g;
  }();
  return f;
}

void *next_symbol(const char *mangled) {
  void *p = dlsym(RTLD_NEXT, mangled);
  if (!p) {
    fprintf(stderr, "serial_probe: %s not found behind the probe\n", mangled);
    abort();
  }
  return p;
}

}  // namespace

// (x86-64 Itanium ABI: a non-static member function is called like a free function whose first argument is `this`)
FlowEvent vFlowManager::computeLocalFlow() {
  typedef FlowEvent (*fn_t)(vFlowManager *);
  static fn_t next = reinterpret_cast<fn_t>(next_symbol("_ZN12vFlowManager16computeLocalFlowEv"));
  FlowEvent r = next(this);
  fprintf(probe_file(), "L %.17g %.17g\n", r.getVx(), r.getVy());
  return r;
}

FlowEvent vFlowManager::computeTrueFlow(int x, int y, unsigned int time, int pol) {
  typedef FlowEvent (*fn_t)(vFlowManager *, int, int, unsigned int, int);
  static fn_t next = reinterpret_cast<fn_t>(next_symbol("_ZN12vFlowManager15computeTrueFlowEiiji"));
  FlowEvent r = next(this, x, y, time, pol);
  fprintf(probe_file(), "T %d %d %u %d %.17g %.17g %d\n", x, y, time, pol, r.getVx(), r.getVy(), r.getScale());
  return r;
}

// computeLocalFlow hands the cells of the winning window (x-major, src/vFlow.cpp:919-927) and the current event to
// computeGrads; the return value is the inlier count (:1352-1369).  The window is recovered from the position of the
// current event's own cell in that list (every candidate window contains it): 0..8, i outer, j inner (:849-851).
int vFlowManager::computeGrads(std::vector<Event> subsurf, Event &cen, double &dtdy, double &dtdx) {
  typedef int (*fn_t)(vFlowManager *, std::vector<Event>, Event &, double &, double &);
  static fn_t next = reinterpret_cast<fn_t>(next_symbol("_ZN12vFlowManager12computeGradsESt6vectorI5EventSaIS1_EERS1_RdS5_"));
  int window = -1;
  int n1 = 1;
  while ((size_t)n1 * n1 < subsurf.size()) n1++;
  const int r = (n1 - 1) / 2;
  if ((size_t)n1 * n1 == subsurf.size() && r > 0)
    for (size_t k = 0; k < subsurf.size(); k++)
      if (subsurf[k].getX() == cen.getX() && subsurf[k].getY() == cen.getY() && subsurf[k].getStamp() == cen.getStamp()) {
        const int a = (int)k / n1, b = (int)k % n1;  // the window starts at (x - a, y - b), its centre is r further
        if ((r - a) % r == 0 && (r - b) % r == 0) window = ((r - a) / r + 1) * 3 + ((r - b) / r + 1);
        break;
      }
  const int inliers = next(this, subsurf, cen, dtdy, dtdx);
  fprintf(probe_file(), "G %d %d %.17g %.17g\n", inliers, window, dtdy, dtdx);
  return inliers;
}
