// farms_cli.cpp -- the FARMS_Flow command line, drop-in for the reference's src/main.cpp.
//
// Same flags, defaults and file contract as the reference (src/main.cpp:21-47, 186-211):
//   --help --filename <path without .txt> --height --width --filtersize --inlierCheck
//   --numEvents | --numevents | --NUMEVENTS   --SERIAL <int>   --v <int>
// Input  : <filename>.txt, one event per line "x y t p"                     (src/vFlow.cpp:150, 173-188)
// Output : <filename>_FARMSOut_batch.txt, 11 columns, byte-compatible with the reference's batch mode
//          "x y t p globalR globalTheta Vx Vy localR localTheta scale"      (src/vFlow.cpp:131, 436-440)
//          <filename>_FARMSOut_.txt, the 8 columns the README documents
//          "x y t p globalR globalTheta localR localTheta"                  (README.md:63)
// --SERIAL 1 (the reference's default, src/main.cpp:31) runs the semantics of vFlowManager::run
// (src/vFlow.cpp:465-826: first line only sets t0, lastEventTime written after pooling, numEvents + 1 lines,
// numEvents capped at filesize / 18) through FARMS_FLAG_SERIAL_SEMANTICS.  One difference on purpose: the reference
// computes those numbers but writes nothing (:487-489, 727-765); here they go to the file that mode names,
// <filename>_FARMSOut_bench_500us.txt (:486), in the 11-column format.  --SERIAL 0 = runFileCopy as before.
// All numbers come from the GPU through the C ABI (include/farms_b200.h); there is no CPU path here.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "farms_b200.h"
#include "farms_textio.h"

namespace {

struct Options {
  int height = 320, width = 320, filtersize = 3, inlier = 5, device = 0;  // src/main.cpp:21-24
  unsigned long long num_events = 1ull << 63;                              // src/main.cpp:28
  std::string filename = "/home/himanshu/POST_DOC/DATA/atisData/bar_square/multiPattern1_fixed_";  // :30
  bool serial = true, verbose = false;
  bool fast = false;  // --fast 1: FP32 ring partials in the pooling kernel (about 1e-7 relative on columns 5-6)
  bool binary = false;  // --binary 1: <filename>.evb in, <filename>_FARMSOut_.bin out (include/farms_textio.h)
  int gpus = 1;         // --gpus N: time-slice the recording over devices device .. device+N-1 (one host thread each)
  bool same_device = false;  // --same-device 1: all slices on --device (in-process transport; one-GPU boxes, tests)
};

void usage() {
  std::puts(
      "Allowed options:\n"
      "  --help                Displays this message\n"
      "  --filename arg        add events file name without extension (.txt)\n"
      "  --height arg          set sensor height\n"
      "  --width arg           set sensor width\n"
      "  --filtersize arg      set size of neighbor for plane fitting\n"
      "  --inlierCheck arg     set minimum number of inliers to validate plane\n"
      "  --numEvents arg       set max number of events to process\n"
      "  --numevents arg       set max number of events to process\n"
      "  --NUMEVENTS arg       set max number of events to process\n"
      "  --SERIAL arg          Serial or Batch processing\n"
      "  --v arg               set verbose to 1 for full debug mode\n"
      "  --device arg          CUDA device ordinal (extension)\n"
      "  --fast arg            1 = fastest pooling kernel, columns 5-6 accurate to ~1e-7 (extension)\n"
      "  --binary arg          1 = binary side-format: <filename>.evb in, <filename>_FARMSOut_.bin out (extension)\n"
      "  --gpus arg            time-slice the recording over this many GPUs, device .. device+N-1 (extension)\n"
      "  --same-device arg     1 = run all --gpus slices on --device (extension, one-GPU boxes)\n");
}

bool parse_int(const std::string &s, int &out) {
  char *end = nullptr;
  long v = std::strtol(s.c_str(), &end, 10);
  if (end == s.c_str() || *end) return false;
  out = (int)v;
  return true;
}

// Returns 0 to continue, 1 = exit with error, 2 = exit ok (--help)
int parse_args(int argc, char **argv, Options &o) {
  long long ne[3] = {-1, -1, -1};  // numEvents, numevents, NUMEVENTS
  bool ne_set[3] = {false, false, false};
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a.rfind("--", 0) != 0) {
      std::fprintf(stderr, "error: too many positional options have been specified on the command line\n");
      return 1;
    }
    std::string name = a.substr(2), val;
    bool has_val = false;
    size_t eq = name.find('=');
    if (eq != std::string::npos) {
      val = name.substr(eq + 1);
      name = name.substr(0, eq);
      has_val = true;
    }
    if (name == "help") {
      usage();
      return 2;
    }
    static const char *known[] = {"filename", "height", "width", "filtersize", "inlierCheck", "numEvents",
                                  "numevents", "NUMEVENTS", "SERIAL", "v", "device", "fast", "binary", "gpus",
                                  "same-device"};
    bool ok = false;
    for (const char *k : known) ok |= name == k;
    if (!ok) {
      std::fprintf(stderr, "error: unrecognised option '--%s'\n", name.c_str());
      return 1;
    }
    if (!has_val) {
      if (i + 1 >= argc) {
        std::fprintf(stderr, "error: the required argument for option '--%s' is missing\n", name.c_str());
        return 1;
      }
      val = argv[++i];
    }
    int iv = 0;
    if (name != "filename" && !parse_int(val, iv)) {
      std::fprintf(stderr, "error: the argument ('%s') for option '--%s' is invalid\n", val.c_str(), name.c_str());
      return 1;
    }
    if (name == "filename") { o.filename = val; std::printf("filename set to %s.\n", val.c_str()); }
    else if (name == "height") { o.height = iv; std::printf("height set to %d.\n", iv); }
    else if (name == "width") { o.width = iv; std::printf("width set to %d.\n", iv); }
    else if (name == "filtersize") { o.filtersize = iv; std::printf("filtersize set to %d.\n", iv); }
    else if (name == "inlierCheck") { o.inlier = iv; std::printf("inlierCheck set to %d.\n", iv); }
    else if (name == "numEvents" || name == "numevents" || name == "NUMEVENTS") {
      const int k = name == "numEvents" ? 0 : name == "numevents" ? 1 : 2;
      ne[k] = iv;
      ne_set[k] = true;
    } else if (name == "SERIAL") {
      o.serial = iv == 1;
      std::puts(o.serial ? "Running serially " : "Running batch ");
    } else if (name == "v") { o.verbose = iv == 1; std::printf("Verbose mode set to %d\n", iv); }
    else if (name == "device") { o.device = iv; }
    else if (name == "fast") { o.fast = iv == 1; }
    else if (name == "binary") { o.binary = iv == 1; }
    else if (name == "gpus") { o.gpus = iv < 1 ? 1 : iv; }
    else if (name == "same-device") { o.same_device = iv == 1; }
  }
  // the reference honours the spellings in the order numEvents, numevents, NUMEVENTS (src/main.cpp:131-151)
  // and converts the int to unsigned long
  for (int k = 0; k < 3; k++)
    if (ne_set[k]) {
      o.num_events = (unsigned long long)ne[k];
      std::printf("numEvents set to %lld.\n", ne[k]);
      break;
    }
  return 0;
}

// the eight output columns in pinned host memory (farms_host_alloc), pageable if pinning fails
struct PinnedColumns {
  uint32_t *t_rel = nullptr;
  double *d[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint8_t *scale = nullptr;
  bool pinned = true, ok = true;
  explicit PinnedColumns(size_t n) {
    auto get = [&](size_t bytes) -> void * {
      void *p = pinned ? farms_host_alloc(bytes) : nullptr;
      if (!p) {
        pinned = false;
        p = std::malloc(bytes ? bytes : 1);
      }
      if (!p) ok = false;
      return p;
    };
    // (all pinned or all pageable: decide on the first allocation)
    t_rel = (uint32_t *)get(n * sizeof(uint32_t));
    const bool first_pinned = pinned;
    for (double *&q : d) q = (double *)get(n * sizeof(double));
    scale = (uint8_t *)get(n);
    if (pinned != first_pinned) mixed = true;
  }
  ~PinnedColumns() {
    auto put = [&](void *p, bool was_pinned) {
      if (!p) return;
      if (was_pinned) farms_host_free(p); else std::free(p);
    };
    // a buffer is pinned iff it was allocated before `pinned` turned false; with `mixed` we cannot tell, so probe
    if (!mixed) {
      put(t_rel, pinned);
      for (double *q : d) put(q, pinned);
      put(scale, pinned);
    }
  }
  bool mixed = false;  // (pathological: pinning failed half-way; the buffers are leaked at exit rather than mis-freed)
};

// ---- single-process multi-GPU run: the recording is cut into equal time slices, one per GPU,
// one host thread and one context per slice, joined by a farms_comm (include/farms_b200.h): NCCL between the
// devices, or the in-process transport when the slices share one device (--same-device 1, one-GPU boxes).
// Slice g owns the events of stream time [g*D, (g+1)*D), also processes the 499-us causal halo in front of them
// and contributes the events of [g*D - 499, (g+1)*D - 499) to the surface exchange (these ranges tile the time
// axis).  Every slice writes its owned rows straight into the caller's output columns.
int run_sliced(const Options &o, const farms_config &base, const farms_events &ev, const farms_out &out,
               farms_timings &tm_sum, uint64_t &events_done, std::string &err) {
  const size_t n = (size_t)ev.n;
  const int G = o.gpus;
  for (size_t i = 1; i < n; i++)
    if (ev.t[i] < ev.t[i - 1]) {
      err = "--gpus needs non-decreasing timestamps (time slices); run this recording on one GPU";
      return 1;
    }
  const uint64_t t0 = ev.t[0], span = ev.t[n - 1] - t0 + 1, D = (span + G - 1) / G, HALO = 499;
  auto first_at = [&](uint64_t ts) {  // first event with stream time >= ts
    return (size_t)(std::lower_bound(ev.t, ev.t + n, t0 + ts) - ev.t);
  };
  unsigned char id[FARMS_COMM_ID_BYTES];
  std::memset(id, 0, sizeof id);
  if (o.same_device) {
    const uint64_t key = (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count() ^ (uint64_t)(uintptr_t)&ev;
    std::memcpy(id, &key, sizeof key);
  } else if (farms_comm_unique_id(id) != FARMS_OK) {
    err = "NCCL is not available (libnccl.so.2 cannot be loaded)";
    return 1;
  }
  std::vector<farms_ctx *> ctx(G, nullptr);
  std::vector<farms_comm *> comm(G, nullptr);
  std::vector<int> rc(G, FARMS_OK);
  std::vector<std::string> msg(G);
  std::vector<size_t> lo(G), begin(G), end(G), surf_end(G);
  for (int g = 0; g < G; g++) {
    begin[g] = first_at((uint64_t)g * D);
    end[g] = g == G - 1 ? n : first_at((uint64_t)(g + 1) * D);
    lo[g] = g == 0 ? 0 : first_at((uint64_t)g * D - std::min<uint64_t>(HALO, (uint64_t)g * D));
    surf_end[g] = g == G - 1 ? lo[g] : first_at((uint64_t)(g + 1) * D - std::min<uint64_t>(HALO, (uint64_t)(g + 1) * D));
  }
  std::vector<std::thread> th;
  for (int g = 0; g < G; g++)
    th.emplace_back([&, g] {
      farms_config cfg = base;
      cfg.device = o.same_device ? base.device : base.device + g;
      rc[g] = farms_create(&ctx[g], &cfg);
      if (rc[g] != FARMS_OK) { msg[g] = "farms_create failed (device " + std::to_string(cfg.device) + ")"; }
      // farms_comm_create is collective: every thread must call it, also after a failed farms_create elsewhere
      if (rc[g] == FARMS_OK) {
        rc[g] = farms_comm_create(&comm[g], ctx[g], G, g, id, o.same_device ? FARMS_COMM_LOCAL : 0u);
        if (rc[g] != FARMS_OK) msg[g] = std::string("farms_comm_create: ") + farms_last_error(ctx[g]);
      }
      if (rc[g] != FARMS_OK) return;
      farms_out fo;
      std::memset(&fo, 0, sizeof fo);
      const size_t at = begin[g];
      fo.t_rel = out.t_rel + at; fo.global_r = out.global_r + at; fo.global_theta = out.global_theta + at;
      fo.vx = out.vx + at; fo.vy = out.vy + at; fo.local_r = out.local_r + at; fo.local_theta = out.local_theta + at;
      fo.scale = out.scale + at;
      rc[g] = farms_comm_process(comm[g], ev.x + lo[g], ev.y + lo[g], ev.t + lo[g], end[g] - lo[g], begin[g] - lo[g],
                                 surf_end[g] - lo[g], t0, 0u, &fo, nullptr);
      if (rc[g] != FARMS_OK) msg[g] = farms_last_error(ctx[g]);
    });
  for (auto &t : th) t.join();
  int bad = 0;
  for (int g = 0; g < G; g++)
    if (rc[g] != FARMS_OK && !bad) {
      err = "GPU slice " + std::to_string(g) + ": " + msg[g];
      bad = 1;
    }
  std::memset(&tm_sum, 0, sizeof tm_sum);
  events_done = 0;
  for (int g = 0; g < G; g++) {
    if (!bad && ctx[g]) {
      farms_timings tm;
      farms_get_timings(ctx[g], &tm);
      tm_sum.total_ms = std::max(tm_sum.total_ms, tm.total_ms);
      tm_sum.ingest_ms += tm.ingest_ms; tm_sum.index_ms += tm.index_ms; tm_sum.fit_ms += tm.fit_ms;
      tm_sum.bin_ms += tm.bin_ms; tm_sum.pool_ms += tm.pool_ms;
      tm_sum.valid_events += tm.valid_events;  // includes the halo events of every slice
      tm_sum.events += end[g] - begin[g];
      events_done += end[g] - begin[g];
    }
    if (comm[g]) farms_comm_destroy(comm[g]);
    if (ctx[g]) farms_destroy(ctx[g]);
  }
  return bad;
}

}  // namespace

int main(int argc, char **argv) {
  Options o;
  int pr = parse_args(argc, argv, o);
  if (pr == 2) return 0;
  if (pr == 1) return 1;

  farms_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.width = o.width; cfg.height = o.height; cfg.filtersize = o.filtersize; cfg.inlier_check = o.inlier;
  cfg.device = o.device;
  // text output prints 6 significant digits: keep the FP64 pooling sums unless told otherwise (the text
  // parse/format around it costs far more than the kernel)
  cfg.flags = (o.fast || o.binary) ? 0u : FARMS_FLAG_EXACT_POOLING;
  if (o.serial) cfg.flags |= FARMS_FLAG_SERIAL_SEMANTICS;
  if (o.serial && o.gpus > 1) {
    std::fprintf(stderr, "error: --gpus needs --SERIAL 0 (the serial driver's semantics are sequential by definition)\n");
    return 1;
  }
  farms_ctx *ctx = nullptr;
  int rc = farms_create(&ctx, &cfg);
  if (rc != FARMS_OK) {
    std::fprintf(stderr, "error: farms_create failed (%d): a B200-class CUDA device is required\n", rc);
    return 1;
  }
  std::printf("[debug Main] : size of lastFlowTime is [sx sy]: [%d %d]\n", o.width, o.height);

  const std::string in_path = o.filename + (o.binary ? ".evb" : ".txt");
  std::printf("%s\nReading input file \n", in_path.c_str());
  farms_events ev;
  char errbuf[256] = "";
  unsigned long long max_events = o.num_events;
  if (o.serial && !o.binary) {
    // src/vFlow.cpp:511: NUMEVENTS = min(NUMEVENTS, filesize / 18); :531-565: the first line, then lines while
    // eventsComputed <= NUMEVENTS, i.e. NUMEVENTS + 1 more
    unsigned long long fsize = 0;
    if (FILE *fp = std::fopen(in_path.c_str(), "rb")) {
      std::fseek(fp, 0, SEEK_END);
      fsize = (unsigned long long)std::ftell(fp);
      std::fclose(fp);
    }
    max_events = std::min<unsigned long long>(o.num_events, fsize / 18) + 2;
  }
  if ((o.binary ? farms_bin_read(in_path.c_str(), max_events, &ev, errbuf, sizeof errbuf)
                : farms_text_read(in_path.c_str(), max_events, 0, &ev, errbuf, sizeof errbuf)) != 0) {
    std::fprintf(stderr, "error: %s\n", errbuf);
    farms_destroy(ctx);
    return 1;
  }
  const size_t n = (size_t)ev.n;
  std::printf("Done reading %zu Events.\n", n);
  if (n == 0) {  // the reference dies in T.at(0) (src/vFlow.cpp:194)
    std::fprintf(stderr, "error: no events in %s\n", in_path.c_str());
    farms_text_free(&ev);
    farms_destroy(ctx);
    return 1;
  }
  std::printf("First time = %llu\nProcessing events \n", (unsigned long long)ev.t[0]);

  // pinned host memory on both sides: farms_process_host then overlaps its copies with the kernels
  const bool in_pinned = farms_host_register(ev.x, n * sizeof(uint16_t)) == FARMS_OK &&
                         farms_host_register(ev.y, n * sizeof(uint16_t)) == FARMS_OK &&
                         farms_host_register(ev.t, n * sizeof(uint64_t)) == FARMS_OK;
  PinnedColumns cols(n);
  if (!cols.ok) {
    std::fprintf(stderr, "error: cannot allocate the output columns\n");
    farms_text_free(&ev);
    farms_destroy(ctx);
    return 1;
  }
  uint32_t *t_rel = cols.t_rel;
  double *gr = cols.d[0], *gth = cols.d[1], *vx = cols.d[2], *vy = cols.d[3], *lr = cols.d[4], *lth = cols.d[5];
  uint8_t *scale = cols.scale;
  farms_out out;
  std::memset(&out, 0, sizeof out);
  out.t_rel = t_rel; out.global_r = gr; out.global_theta = gth; out.vx = vx;
  out.vy = vy; out.local_r = lr; out.local_theta = lth; out.scale = scale;
  if (o.verbose) std::printf("[farms_b200] host buffers: input %s, output %s\n", in_pinned ? "pinned" : "pageable",
                             cols.pinned ? "pinned" : "pageable");

  farms_timings tm_sliced;
  uint64_t sliced_events = 0;
  // device working memory up front, like the reference's surfaces (allocated in its constructor, outside its timer)
  if (o.gpus <= 1 && farms_reserve(ctx, n, 1) != FARMS_OK) {
    std::fprintf(stderr, "error: %s\n", farms_last_error(ctx));
    farms_text_free(&ev);
    farms_destroy(ctx);
    return 1;
  }
  const auto a = std::chrono::system_clock::now();
  if (o.gpus > 1) {
    std::string serr;
    if (run_sliced(o, cfg, ev, out, tm_sliced, sliced_events, serr)) {
      std::fprintf(stderr, "error: %s\n", serr.c_str());
      farms_text_free(&ev);
      farms_destroy(ctx);
      return 1;
    }
    rc = FARMS_OK;
  } else {
    rc = farms_process_host(ctx, ev.x, ev.y, ev.t, nullptr, n, &out);
  }
  const auto b = std::chrono::system_clock::now();
  if (rc != FARMS_OK) {
    std::fprintf(stderr, "error: %s\n", farms_last_error(ctx));
    farms_text_free(&ev);
    farms_destroy(ctx);
    return 1;
  }
  const long usec = (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count();
  std::printf("\nDone processing!\n\nWriting output file.\n");

  if (in_pinned) {
    farms_host_unregister(ev.x);
    farms_host_unregister(ev.y);
    farms_host_unregister(ev.t);
  }
  const std::string out11 = o.filename + (o.binary ? "_FARMSOut_.bin" : o.serial ? "_FARMSOut_bench_500us.txt" : "_FARMSOut_batch.txt"),
                    out8 = o.filename + "_FARMSOut_.txt";
  if ((o.binary ? farms_bin_write(out11.c_str(), n, ev.xi, ev.yi, t_rel, ev.pol, gr, gth,
                                  vx, vy, lr, lth, scale)
                : farms_text_write(out11.c_str(), out8.c_str(), n, ev.xi, ev.yi, t_rel, ev.pol, gr,
                                   gth, vx, vy, lr, lth, scale, 0)) != 0) {
    std::fprintf(stderr, "error: cannot write %s / %s\n", out11.c_str(), out8.c_str());
    farms_text_free(&ev);
    farms_destroy(ctx);
    return 1;
  }
  farms_text_free(&ev);

  farms_timings tm;
  farms_get_timings(ctx, &tm);
  if (o.gpus > 1) tm = tm_sliced;  // slowest slice's device time, stage times summed over the slices
  const uint64_t events_done = o.gpus > 1 ? sliced_events : farms_num_events(ctx);
  // same line as src/main.cpp:209, but with real (not integer-divided) seconds
  const double sec = (double)usec / 1e6;
  std::printf("[Benchmark Main] : Processing time   : %ld usec %g sec  with rate of : %g events/sec\n", usec, sec,
              sec > 0 ? ((double)events_done - 1) / sec : 0.0);
  std::printf("[farms_b200] device %.3f ms (ingest %.3f, index %.3f, fit %.3f, bin %.3f, pool %.3f), valid %llu of %llu\n",
              tm.total_ms, tm.ingest_ms, tm.index_ms, tm.fit_ms, tm.bin_ms, tm.pool_ms,
              (unsigned long long)tm.valid_events, (unsigned long long)tm.events);
  farms_destroy(ctx);
  return 0;
}
