// comm.cu -- time-sliced multi-GPU runs of the FARMS path behind the C ABI (include/farms_b200.h, farms_comm_*).
//
// The reference is single-threaded; what is distributed here is its event loop
// (src/vFlow.cpp:223-414) cut into time slices, one per GPU.  Rank g owns the events of its slice, also processes
// the 499-us causal halo in front of it (pooling admits |dt| < 500 us, src/vFlow.cpp:1002) and rebuilds the surface
// of active events at its halo start from the "last event per pixel" surfaces of the earlier slices (the surface
// never forgets, src/vFlow.cpp:267): one all-gather of W*H x 5 bytes per rank, folded in rank order.  The per-event
// outputs of the README's 8-column contract (globalR, globalTheta, localR, localTheta as one float4) travel to the
// root rank batch by batch while the next batch computes.
//
// Two transports behind one interface:
//   NCCL   one process (or host thread) per GPU; ncclAllGather for the surfaces; the outputs go to the root batch
//          by batch as copy-engine writes into its buffer mapped over CUDA IPC (peer memory over NVLink), or with
//          ncclSend/ncclRecv where that mapping is not available.  libnccl.so.2 is loaded on first use (dlopen),
//          so the single-GPU library has no NCCL dependency.
//   LOCAL  all ranks are host threads of one process (the FARMS_Flow command line): the ranks publish their device
//          buffers in a shared table and copy between devices directly (cudaMemcpyAsync over NVLink peer access),
//          outputs are written straight into the root's buffer.  Also what lets a one-GPU box test the whole
//          protocol (several ranks on one device, which NCCL refuses).
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "farms_ctx.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// NCCL, loaded lazily
// ---------------------------------------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string error;
};

#ifndef FARMS_NCCL_FALLBACK
#define FARMS_NCCL_FALLBACK ""
#endif

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so", FARMS_NCCL_FALLBACK};
    for (const char *nm : names) {
      if (!nm[0]) continue;
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      const char *why = dlerror();  // (a second call would return null: dlerror clears its state)
      api.error = std::string("cannot load libnccl.so.2: ") + (why ? why : "?");
      return;
    }
#define SYM(field, name)                                                  \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); \
  if (!api.field) api.error = std::string("libnccl lacks ") + name;
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather") SYM(Broadcast, "ncclBroadcast") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  });
  return &api;
}

// ---------------------------------------------------------------------------------------------------------------
// LOCAL transport: a group of host threads
// ---------------------------------------------------------------------------------------------------------------
struct LocalGroup {
  int nranks = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t generation = 0;
  int refs = 0;
  // published per rank
  std::vector<const uint32_t *> surf_t;
  std::vector<const uint8_t *> surf_hit;
  std::vector<uint64_t> meta;  // 4 words per rank
  float *root_dst = nullptr;
  int failed = 0;

  // false if some rank has failed (it will never arrive): the others must not wait for it
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (failed) return false;
    const uint64_t gen = generation;
    if (++arrived == nranks) {
      arrived = 0;
      generation++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen || failed; });
    }
    return !failed;
  }
  void fail() {
    std::lock_guard<std::mutex> lk(mu);
    failed = 1;
    cv.notify_all();
  }
};

std::mutex g_groups_mu;
std::vector<std::pair<uint64_t, LocalGroup *>> g_groups;  // keyed by the first 8 bytes of the group id

}  // namespace

struct farms_comm {
  farms_ctx *ctx = nullptr;
  int nranks = 1, rank = 0;
  bool local = false;
  ncclComm_t nccl = nullptr;
  LocalGroup *group = nullptr;
  cudaStream_t cstream = nullptr;  // transfers of outputs, beside the compute stream
  // surface exchange
  uint32_t *surf_t = nullptr, *all_t = nullptr;
  uint8_t *surf_hit = nullptr, *all_hit = nullptr;
  uint64_t *d_meta = nullptr, *h_meta = nullptr;  // 4 words per rank: events, halo events, max_batch, unused
  // device-resident copy of a host slice (the slice is read twice: surface pass, then the event loop)
  DevBuf dev_x, dev_y, dev_t;
  // outputs on their way to the root
  float *sendbuf[2] = {nullptr, nullptr};
  size_t sendcap = 0;
  cudaEvent_t ev_packed[2]{}, ev_sent[2]{};
  bool sent_pending[2] = {false, false};
  // gather over peer memory (NCCL transport): the root's destination buffer mapped into this process with CUDA IPC
  // (or its plain address when root and this rank share a process); outputs then travel as copy-engine writes over
  // NVLink, which need no SM while the pooling kernel owns them all.  Falls back to ncclSend/ncclRecv.
  unsigned char *d_ipc = nullptr, *h_ipc = nullptr;  // 128-byte descriptor broadcast from the root
  unsigned char ipc_cached[64] = {0};                // handle of the mapping held in ipc_base
  void *ipc_base = nullptr;
  float *peer_dst = nullptr;                         // root's dst as seen from this rank, or null (fallback)
  int gather_path = 0;                               // of the last call: 0 none, 1 nccl send/recv, 2 peer memory
  // state of the call in flight (read by the batch hook)
  const farms_gather *gather = nullptr;
  std::vector<uint64_t> n_all, n_halo, first_out;  // per rank
  uint64_t maxb = 0;
  uint64_t batches_seen = 0;     // batches with outputs this rank has packed
  uint64_t recv_posted = 0;      // root: batch indices [0, recv_posted) have their receives posted
  // host wall clock of the phases of the last farms_comm_process call (each ends with a synchronisation):
  // upload + slice surface, exchange + fold, event loop, drain of the output transfers
  float phase_ms[4] = {0.f, 0.f, 0.f, 0.f};
};

namespace {

#define CUC(call)                                                                                              \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess)                                                                                     \
      return farms_fail(cm->ctx, e_ == cudaErrorMemoryAllocation ? FARMS_ERR_NOMEM : FARMS_ERR_CUDA,           \
                        "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);                  \
  } while (0)
#define NC(call)                                                                                               \
  do {                                                                                                         \
    ncclResult_t r_ = (call);                                                                                  \
    if (r_ != ncclSuccess)                                                                                     \
      return farms_fail(cm->ctx, FARMS_ERR_COMM, "%s: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_),      \
                        __FILE__, __LINE__);                                                                   \
  } while (0)

int ensure_buf(farms_comm *cm, DevBuf &b, size_t bytes) {
  if (b.bytes >= bytes) return 0;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  const size_t want = bytes + bytes / 16 + 4096;
  CUC(cudaMalloc(&b.p, want));
  b.bytes = want;
  return 0;
}

// owned events of rank r that internal batch k of its slice produces: count and position among its outputs
void batch_outputs(const farms_comm *cm, int r, uint64_t k, uint64_t *count, uint64_t *out_off) {
  const uint64_t lo = k * cm->maxb, hi = std::min(cm->n_all[r], (k + 1) * cm->maxb);
  const uint64_t a = std::max(lo, cm->n_halo[r]);
  *count = hi > a ? hi - a : 0;
  *out_off = a - cm->n_halo[r];
}
uint64_t batches_of(const farms_comm *cm, int r) { return (cm->n_all[r] + cm->maxb - 1) / cm->maxb; }

// root: post the receives of batch index k from every peer (one NCCL group: they run concurrently)
int post_receives(farms_comm *cm, uint64_t k) {
  NcclApi *N = nccl_api();
  bool any = false;
  for (int r = 0; r < cm->nranks; r++) {
    if (r == cm->rank) continue;
    uint64_t cnt, off;
    batch_outputs(cm, r, k, &cnt, &off);
    if (!cnt || k >= batches_of(cm, r)) continue;
    if (!any) NC(N->GroupStart());
    any = true;
    NC(N->Recv(cm->gather->dst + 4 * (cm->first_out[r] + off), 4 * cnt, ncclFloat, r, cm->nccl, cm->cstream));
  }
  if (any) NC(N->GroupEnd());
  return 0;
}

// Batch hook: the four contract columns of the batch as float4 per event, on their way to the root while the next
// batch computes.
int gather_hook(void *user, farms_ctx *c, const FarmsBatchView *v) {
  farms_comm *cm = (farms_comm *)user;
  const farms_gather *g = cm->gather;
  const uint64_t my_first = cm->first_out[cm->rank];
  if (cm->rank == g->root) {
    // own outputs: packed straight into the destination
    launch_pack4(v->gr, v->gth, v->lr, v->lth, v->n_out, (float4 *)(g->dst + 4 * (my_first + v->out_off)), v->stream);
    CUC(cudaGetLastError());
    if (!cm->local && cm->nranks > 1 && cm->gather_path == 1) {
      // the peers run in step with this rank: by now they have (nearly) finished the previous batch, so its
      // receive kernels do not sit on SMs waiting for data
      const uint64_t upto = cm->batches_seen;  // batches before this one
      for (; cm->recv_posted < upto; cm->recv_posted++) {
        int rc = post_receives(cm, cm->recv_posted);
        if (rc) return rc;
      }
    }
    cm->batches_seen++;
    return 0;
  }
  const int b = (int)(cm->batches_seen & 1);
  if (v->n_out > cm->sendcap) return farms_fail(c, FARMS_ERR_STATE, "gather: batch larger than the send buffer");
  if (cm->sent_pending[b]) CUC(cudaStreamWaitEvent(v->stream, cm->ev_sent[b], 0));  // buffer free again
  launch_pack4(v->gr, v->gth, v->lr, v->lth, v->n_out, (float4 *)cm->sendbuf[b], v->stream);
  CUC(cudaGetLastError());
  CUC(cudaEventRecord(cm->ev_packed[b], v->stream));
  CUC(cudaStreamWaitEvent(cm->cstream, cm->ev_packed[b], 0));
  if (cm->local) {
    // one process: the root's buffer is directly addressable (peer access / same device)
    CUC(cudaMemcpyAsync(cm->group->root_dst + 4 * (my_first + v->out_off), cm->sendbuf[b], v->n_out * 16,
                        cudaMemcpyDefault, cm->cstream));
  } else if (cm->gather_path == 2) {
    // the root's buffer is mapped here: a copy-engine write over NVLink, no SM on either side
    CUC(cudaMemcpyAsync(cm->peer_dst + 4 * (my_first + v->out_off), cm->sendbuf[b], v->n_out * 16, cudaMemcpyDefault,
                        cm->cstream));
  } else {
    NC(nccl_api()->Send(cm->sendbuf[b], 4 * v->n_out, ncclFloat, g->root, cm->nccl, cm->cstream));
  }
  CUC(cudaEventRecord(cm->ev_sent[b], cm->cstream));
  cm->sent_pending[b] = true;
  cm->batches_seen++;
  return 0;
}

// Agree on the gather path (NCCL transport).  The root describes its destination buffer -- IPC handle of the
// allocation, offset of dst inside it, a token of its process and the raw address -- and broadcasts the 128 bytes; every
// other rank maps it (cudaIpcOpenMemHandle, cached across calls; the plain address if it shares the root's
// process).  An all-gather of one flag per rank then settles it: peer memory only if EVERY rank succeeded.
struct IpcDesc {
  cudaIpcMemHandle_t handle;  // 64 bytes
  uint64_t offset, address, process, ok;
};
static_assert(sizeof(IpcDesc) <= 128, "IPC descriptor fits the broadcast buffer");

// Identifies this process among the ranks of a node.  Not the pid alone: ranks in different containers (separate
// pid namespaces) can share a pid, and taking the root's raw address then would be a wild pointer.
uint64_t process_token() {
  static const uint64_t token = [] {
    uint64_t v = ((uint64_t)getpid() << 32) ^ (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count();
    v ^= (uint64_t)(uintptr_t)&g_groups_mu * 0x9E3779B97F4A7C15ull;
    try {
      std::random_device rd;
      v ^= ((uint64_t)rd() << 32) ^ (uint64_t)rd();
    } catch (...) {
    }
    return v ? v : 1ull;
  }();
  return token;
}

int setup_gather_path(farms_comm *cm, const farms_gather *g) {
  NcclApi *N = nccl_api();
  cudaStream_t s = cm->ctx->stream;
  IpcDesc *d = reinterpret_cast<IpcDesc *>(cm->h_ipc);
  memset(d, 0, sizeof *d);
  if (cm->rank == g->root) {
    // base of the allocation dst lives in (the handle names whole allocations): driver entry point via the runtime
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    unsigned long long base = 0;
    size_t size = 0;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn &&
        reinterpret_cast<range_fn>(fn)(&base, &size, (unsigned long long)(uintptr_t)g->dst) == 0 &&
        cudaIpcGetMemHandle(&d->handle, (void *)(uintptr_t)base) == cudaSuccess) {
      d->offset = (uint64_t)((uintptr_t)g->dst - (uintptr_t)base);
      d->ok = 1;
    }
    cudaGetLastError();
    d->address = (uint64_t)(uintptr_t)g->dst;
    d->process = process_token();
  }
  CUC(cudaMemcpyAsync(cm->d_ipc, cm->h_ipc, 128, cudaMemcpyHostToDevice, s));
  NC(N->Broadcast(cm->d_ipc, cm->d_ipc, 128, ncclChar, g->root, cm->nccl, s));
  CUC(cudaMemcpyAsync(cm->h_ipc, cm->d_ipc, 128, cudaMemcpyDeviceToHost, s));
  CUC(cudaStreamSynchronize(s));
  cm->peer_dst = nullptr;
  uint64_t mine = 1;
  if (cm->rank != g->root) {
    mine = 0;
    if (d->process == process_token()) {  // root is a thread of this process: its address is ours
      cm->peer_dst = (float *)(uintptr_t)d->address;
      mine = 1;
    } else if (d->ok) {
      if (!cm->ipc_base || memcmp(cm->ipc_cached, &d->handle, 64) != 0) {
        if (cm->ipc_base) cudaIpcCloseMemHandle(cm->ipc_base);
        cm->ipc_base = nullptr;
        if (cudaIpcOpenMemHandle(&cm->ipc_base, d->handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess)
          memcpy(cm->ipc_cached, &d->handle, 64);
        else
          cm->ipc_base = nullptr;
        cudaGetLastError();
      }
      if (cm->ipc_base) {
        cm->peer_dst = (float *)((char *)cm->ipc_base + d->offset);
        mine = 1;
      }
    }
  }
  // every rank must take the same path
  cm->h_meta[4 * cm->rank + 3] = mine;
  CUC(cudaMemcpyAsync(cm->d_meta + 4 * cm->rank + 3, cm->h_meta + 4 * cm->rank + 3, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  NC(N->AllGather(cm->d_meta + 4 * cm->rank, cm->d_meta, 4, ncclUint64, cm->nccl, s));
  CUC(cudaMemcpyAsync(cm->h_meta, cm->d_meta, (size_t)cm->nranks * 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  CUC(cudaStreamSynchronize(s));
  bool all = true;
  for (int r = 0; r < cm->nranks; r++) all = all && cm->h_meta[4 * r + 3] != 0;
  cm->gather_path = all ? 2 : 1;
  return 0;
}

}  // namespace

extern "C" {

int farms_comm_unique_id(void *id128) {
  if (!id128) return FARMS_ERR_ARG;
  NcclApi *N = nccl_api();
  if (!N->error.empty()) return FARMS_ERR_COMM;
  ncclUniqueId id;
  if (N->GetUniqueId(&id) != ncclSuccess) return FARMS_ERR_COMM;
  static_assert(sizeof(id) == FARMS_COMM_ID_BYTES, "NCCL unique id size");
  memcpy(id128, &id, sizeof id);
  return FARMS_OK;
}

int farms_comm_create(farms_comm **out, farms_ctx *ctx, int nranks, int rank, const void *id128, uint32_t flags) {
  if (!out || !ctx || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) return FARMS_ERR_ARG;
  *out = nullptr;
  farms_comm *cm = new (std::nothrow) farms_comm();
  if (!cm) return FARMS_ERR_NOMEM;
  cm->ctx = ctx;
  cm->nranks = nranks;
  cm->rank = rank;
  cm->local = (flags & FARMS_COMM_LOCAL) != 0;
  auto bail = [&](int code) {
    farms_comm_destroy(cm);
    return code;
  };
  if (cudaSetDevice(ctx->cfg.device) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&cm->cstream, cudaStreamNonBlocking) != cudaSuccess) return bail(FARMS_ERR_CUDA);
  for (int b = 0; b < 2; b++)
    if (cudaEventCreateWithFlags(&cm->ev_packed[b], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&cm->ev_sent[b], cudaEventDisableTiming) != cudaSuccess)
      return bail(FARMS_ERR_CUDA);
  const size_t npx = ctx->npx;
  bool ok = cudaMalloc((void **)&cm->surf_t, npx * 4) == cudaSuccess && cudaMalloc((void **)&cm->surf_hit, npx) == cudaSuccess &&
            cudaMalloc((void **)&cm->all_t, (size_t)nranks * npx * 4) == cudaSuccess &&
            cudaMalloc((void **)&cm->all_hit, (size_t)nranks * npx) == cudaSuccess &&
            cudaMalloc((void **)&cm->d_meta, (size_t)nranks * 4 * sizeof(uint64_t)) == cudaSuccess &&
            cudaMallocHost((void **)&cm->h_meta, (size_t)nranks * 4 * sizeof(uint64_t)) == cudaSuccess &&
            cudaMalloc((void **)&cm->d_ipc, 128) == cudaSuccess && cudaMallocHost((void **)&cm->h_ipc, 128) == cudaSuccess;
  if (!ok) return bail(FARMS_ERR_NOMEM);
  cm->n_all.assign(nranks, 0);
  cm->n_halo.assign(nranks, 0);
  cm->first_out.assign(nranks, 0);
  if (nranks > 1 && cm->local) {
    uint64_t key;
    memcpy(&key, id128, sizeof key);
    bool mismatch = false;
    {
      std::lock_guard<std::mutex> lk(g_groups_mu);
      LocalGroup *found = nullptr;
      for (auto &kv : g_groups)
        if (kv.first == key) found = kv.second;
      if (!found) {
        found = new LocalGroup();
        found->nranks = nranks;
        found->surf_t.assign(nranks, nullptr);
        found->surf_hit.assign(nranks, nullptr);
        found->meta.assign((size_t)nranks * 4, 0);
        g_groups.emplace_back(key, found);
      }
      if (found->nranks == nranks) {
        cm->group = found;
        found->refs++;
      } else {
        mismatch = true;  // (bail takes the registry lock itself: not from inside this scope)
      }
    }
    if (mismatch) {
      farms_fail(ctx, FARMS_ERR_ARG, "the group with this id was created for a different number of ranks");
      return bail(FARMS_ERR_ARG);
    }
  } else if (nranks > 1) {
    NcclApi *N = nccl_api();
    if (!N->error.empty()) {
      farms_fail(ctx, FARMS_ERR_COMM, "%s", N->error.c_str());
      return bail(FARMS_ERR_COMM);
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = N->CommInitRank(&cm->nccl, nranks, id, rank);
    if (r != ncclSuccess) {
      farms_fail(ctx, FARMS_ERR_COMM, "ncclCommInitRank: %s", N->GetErrorString(r));
      return bail(FARMS_ERR_COMM);
    }
  }
  *out = cm;
  return FARMS_OK;
}

void farms_comm_destroy(farms_comm *cm) {
  if (!cm) return;
  if (cm->ctx) cudaSetDevice(cm->ctx->cfg.device);
  if (cm->cstream) cudaStreamSynchronize(cm->cstream);
  if (cm->nccl) nccl_api()->CommDestroy(cm->nccl);
  if (cm->group) {
    std::lock_guard<std::mutex> lk(g_groups_mu);
    if (--cm->group->refs == 0) {
      for (size_t i = 0; i < g_groups.size(); i++)
        if (g_groups[i].second == cm->group) {
          g_groups.erase(g_groups.begin() + (long)i);
          break;
        }
      delete cm->group;
    }
  }
  void *ps[] = {cm->surf_t, cm->surf_hit, cm->all_t, cm->all_hit, cm->d_meta, cm->dev_x.p, cm->dev_y.p, cm->dev_t.p,
                cm->sendbuf[0], cm->sendbuf[1]};
  for (void *p : ps)
    if (p) cudaFree(p);
  if (cm->h_meta) cudaFreeHost(cm->h_meta);
  if (cm->h_ipc) cudaFreeHost(cm->h_ipc);
  if (cm->d_ipc) cudaFree(cm->d_ipc);
  if (cm->ipc_base) cudaIpcCloseMemHandle(cm->ipc_base);
  for (int b = 0; b < 2; b++) {
    if (cm->ev_packed[b]) cudaEventDestroy(cm->ev_packed[b]);
    if (cm->ev_sent[b]) cudaEventDestroy(cm->ev_sent[b]);
  }
  if (cm->cstream) cudaStreamDestroy(cm->cstream);
  delete cm;
}

static int comm_process_impl(farms_comm *cm, const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t n,
                             uint64_t n_halo, uint64_t n_surface, uint64_t t0, uint32_t flags, const farms_out *out,
                             const farms_gather *gather);

int farms_comm_process(farms_comm *cm, const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t n,
                       uint64_t n_halo, uint64_t n_surface, uint64_t t0, uint32_t flags, const farms_out *out,
                       const farms_gather *gather) {
  if (!cm || !cm->ctx) return FARMS_ERR_ARG;
  const int rc = comm_process_impl(cm, x, y, t, n, n_halo, n_surface, t0, flags, out, gather);
  if (rc != FARMS_OK && cm->group) cm->group->fail();  // in-process ranks waiting at a barrier give up too
  return rc;
}

static int comm_process_impl(farms_comm *cm, const uint16_t *x, const uint16_t *y, const uint64_t *t, uint64_t n,
                             uint64_t n_halo, uint64_t n_surface, uint64_t t0, uint32_t flags, const farms_out *out,
                             const farms_gather *gather) {
  farms_ctx *c = cm->ctx;
  c->err.clear();
  if (n_halo > n || n_surface > n || (n && (!x || !y || !t))) return farms_fail(c, FARMS_ERR_ARG, "bad slice arguments");
  if (n >= (1ull << 32) - 1) return farms_fail(c, FARMS_ERR_ARG, "slice too long");
  if (gather && (gather->root < 0 || gather->root >= cm->nranks || (cm->rank == gather->root && !gather->dst)))
    return farms_fail(c, FARMS_ERR_ARG, "bad gather descriptor");
  if (c->serial && cm->nranks > 1)
    return farms_fail(c, FARMS_ERR_ARG, "FARMS_FLAG_SERIAL_SEMANTICS is sequential by definition: one rank only");
  const bool in_device = (flags & FARMS_IO_INPUT_ON_DEVICE) != 0, out_device = (flags & FARMS_IO_OUTPUT_ON_DEVICE) != 0;
  CUC(cudaSetDevice(c->cfg.device));
  cudaStream_t s = c->stream;
  const int R = cm->nranks;
  using clk = std::chrono::steady_clock;
  auto ms_since = [](clk::time_point a) { return std::chrono::duration<float, std::milli>(clk::now() - a).count(); };
  clk::time_point tp = clk::now();
  cm->phase_ms[0] = cm->phase_ms[1] = cm->phase_ms[2] = cm->phase_ms[3] = 0.f;
  int rc = farms_set_t0(c, t0);
  if (rc) return rc;
  cm->maxb = c->cfg.max_batch ? c->cfg.max_batch : DEFAULT_MAX_BATCH;

  // ---- a host slice is made device-resident once: it is read twice (surface pass, event loop) ----
  const uint16_t *dx = x, *dy = y;
  const uint64_t *dt = t;
  if (!in_device && R > 1 && n) {
    if ((rc = ensure_buf(cm, cm->dev_x, n * 2)) || (rc = ensure_buf(cm, cm->dev_y, n * 2)) || (rc = ensure_buf(cm, cm->dev_t, n * 8)))
      return rc;
    const size_t piece = 8u << 20;
    for (uint64_t off = 0; off < n; off += piece) {  // pieces keep the copy engine and the surface kernel in step
      const size_t nb = (size_t)std::min<uint64_t>(piece, n - off);
      CUC(cudaMemcpyAsync((uint16_t *)cm->dev_x.p + off, x + off, nb * 2, cudaMemcpyHostToDevice, s));
      CUC(cudaMemcpyAsync((uint16_t *)cm->dev_y.p + off, y + off, nb * 2, cudaMemcpyHostToDevice, s));
      CUC(cudaMemcpyAsync((uint64_t *)cm->dev_t.p + off, t + off, nb * 8, cudaMemcpyHostToDevice, s));
    }
    dx = (const uint16_t *)cm->dev_x.p;
    dy = (const uint16_t *)cm->dev_y.p;
    dt = (const uint64_t *)cm->dev_t.p;
  }

  // ---- exchange: slice sizes, and the surface of active events at every slice start ----
  cm->h_meta[4 * cm->rank + 0] = n;
  cm->h_meta[4 * cm->rank + 1] = n_halo;
  cm->h_meta[4 * cm->rank + 2] = cm->maxb;
  cm->h_meta[4 * cm->rank + 3] = 0;
  if (R > 1) {
    // farms_slice_surface synchronises the compute stream: the uploads above are done when it returns
    if ((rc = farms_slice_surface(c, dx, dy, dt, n_surface, t0, cm->surf_t, cm->surf_hit))) return rc;
    cm->phase_ms[0] = ms_since(tp);
    tp = clk::now();
    const size_t npx = c->npx;
    if (cm->local) {
      LocalGroup *G = cm->group;
      {
        std::lock_guard<std::mutex> lk(G->mu);
        G->surf_t[cm->rank] = cm->surf_t;
        G->surf_hit[cm->rank] = cm->surf_hit;
        memcpy(&G->meta[4 * cm->rank], &cm->h_meta[4 * cm->rank], 4 * sizeof(uint64_t));
        if (gather && cm->rank == gather->root) G->root_dst = gather->dst;
      }
      if (!G->barrier()) return farms_fail(c, FARMS_ERR_COMM, "another rank of the group failed");
      for (int r = 0; r < R; r++) {
        memcpy(&cm->h_meta[4 * r], &G->meta[4 * r], 4 * sizeof(uint64_t));
        if (r >= cm->rank) continue;  // only earlier slices are folded
        CUC(cudaMemcpyAsync(cm->all_t + (size_t)r * npx, G->surf_t[r], npx * 4, cudaMemcpyDefault, s));
        CUC(cudaMemcpyAsync(cm->all_hit + (size_t)r * npx, G->surf_hit[r], npx, cudaMemcpyDefault, s));
      }
      CUC(cudaStreamSynchronize(s));
      if (!G->barrier()) return farms_fail(c, FARMS_ERR_COMM, "another rank of the group failed");  // all have read the surfaces
    } else {
      NcclApi *N = nccl_api();
      CUC(cudaMemcpyAsync(cm->d_meta + 4 * cm->rank, cm->h_meta + 4 * cm->rank, 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
      NC(N->GroupStart());
      NC(N->AllGather(cm->surf_t, cm->all_t, npx, ncclUint32, cm->nccl, s));
      NC(N->AllGather(cm->surf_hit, cm->all_hit, npx, ncclUint8, cm->nccl, s));
      NC(N->AllGather(cm->d_meta + 4 * cm->rank, cm->d_meta, 4, ncclUint64, cm->nccl, s));
      NC(N->GroupEnd());
      CUC(cudaMemcpyAsync(cm->h_meta, cm->d_meta, (size_t)R * 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
      CUC(cudaStreamSynchronize(s));
    }
    for (int r = 0; r < cm->rank; r++)  // later slices win: fold in stream order
      launch_sae_fold(c->sae, npx, cm->all_t + (size_t)r * npx, cm->all_hit + (size_t)r * npx, s);
    CUC(cudaGetLastError());
  }
  uint64_t first = 0;
  for (int r = 0; r < R; r++) {
    cm->n_all[r] = cm->h_meta[4 * r + 0];
    cm->n_halo[r] = cm->h_meta[4 * r + 1];
    if (cm->h_meta[4 * r + 2] != cm->maxb) return farms_fail(c, FARMS_ERR_ARG, "ranks disagree on max_batch");
    cm->first_out[r] = first;
    first += cm->n_all[r] - cm->n_halo[r];
    if (gather && gather->counts) gather->counts[r] = cm->n_all[r] - cm->n_halo[r];
  }

  if (R > 1) {
    CUC(cudaStreamSynchronize(s));
    cm->phase_ms[1] = ms_since(tp);
  }
  tp = clk::now();
  // ---- the event loop of this slice, outputs leaving batch by batch ----
  FarmsBatchHook hook;
  cm->gather = gather;
  cm->batches_seen = 0;
  cm->recv_posted = 0;
  cm->sent_pending[0] = cm->sent_pending[1] = false;
  cm->gather_path = 0;
  if (gather) {
    if (!cm->local && R > 1) {
      if ((rc = setup_gather_path(cm, gather))) return rc;
    }
    hook.fn = gather_hook;
    hook.user = cm;
    if (cm->rank != gather->root) {
      const size_t want = (size_t)std::min<uint64_t>(cm->maxb, std::max<uint64_t>(n, 1));
      if (want > cm->sendcap) {
        CUC(cudaDeviceSynchronize());
        for (int b = 0; b < 2; b++) {
          if (cm->sendbuf[b]) cudaFree(cm->sendbuf[b]);
          cm->sendbuf[b] = nullptr;
        }
        cm->sendcap = 0;
        for (int b = 0; b < 2; b++) CUC(cudaMalloc((void **)&cm->sendbuf[b], want * 16));
        cm->sendcap = want;
      }
    }
  }
  const bool stays_on_device = in_device || (R > 1 && n);
  rc = farms_process_impl(c, dx, dy, dt, n, out, stays_on_device, out_device, n_halo, gather ? &hook : nullptr);
  if (rc) return rc;
  cm->phase_ms[2] = ms_since(tp);
  tp = clk::now();
  if (gather) {
    if (!cm->local && R > 1 && cm->gather_path == 2) {
      // peer-memory path: every rank's copies are done when its side stream is; one tiny collective tells the root
      CUC(cudaStreamSynchronize(cm->cstream));
      NC(nccl_api()->AllGather(cm->d_meta + 4 * cm->rank, cm->d_meta, 4, ncclUint64, cm->nccl, s));
      CUC(cudaStreamSynchronize(s));
    } else if (!cm->local && R > 1 && cm->rank == gather->root) {
      uint64_t most = 0;
      for (int r = 0; r < R; r++) most = std::max(most, batches_of(cm, r));
      for (; cm->recv_posted < most; cm->recv_posted++)
        if ((rc = post_receives(cm, cm->recv_posted))) return rc;
    }
    CUC(cudaStreamSynchronize(cm->cstream));
    // the root's buffer is complete when every rank has passed here
    if (cm->local && R > 1 && !cm->group->barrier()) return farms_fail(c, FARMS_ERR_COMM, "another rank of the group failed");
  }
  cm->gather = nullptr;
  cm->phase_ms[3] = ms_since(tp);
  return FARMS_OK;
}

int farms_comm_phases(const farms_comm *cm, float ms[4]) {
  if (!cm || !ms) return FARMS_ERR_ARG;
  for (int k = 0; k < 4; k++) ms[k] = cm->phase_ms[k];
  return FARMS_OK;
}

int farms_comm_info(const farms_comm *cm, int32_t *nranks, int32_t *rank, int32_t *transport) {
  if (!cm) return FARMS_ERR_ARG;
  if (nranks) *nranks = cm->nranks;
  if (rank) *rank = cm->rank;
  // 0 none, 1 NCCL (gather by send/recv), 2 in-process, 3 NCCL + gather over peer memory (CUDA IPC / shared address)
  if (transport) *transport = cm->nranks == 1 ? 0 : cm->local ? 2 : cm->gather_path == 2 ? 3 : 1;
  return FARMS_OK;
}

}  // extern "C"
