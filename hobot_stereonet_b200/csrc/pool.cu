// Multi-GPU front behind the C ABI (include/snb200.h, snb_pool_*): ONE process, one model replica per GPU, stereo pairs
// sharded by batch, the weight blob installed with a single NCCL broadcast at init and no collective on the per-frame
// path (BASELINE.json north_star).  The reference drives one BPU from one node process
// (stereonet_node.cpp:44,812); this is that node process driving N B200s.
//
//   init      replica 0 loads `model_file` (a normal snb_create); the others are created with SNB_FLAG_DEFER_WEIGHTS.
//             The blob goes to replica 0's HBM, ncclBroadcast (one group call over ncclCommInitAll communicators:
//             NVLink 5 / NVSwitch) puts it into every other GPU's HBM, and each replica installs it from its own
//             device buffer with snb_set_weights(is_device = 1).
//   per call  snb_pool_infer_async -> host/dispatch.h Dispatcher (least calls in flight, round-robin among equals);
//             snb_pool_infer (a whole batch in one call) -> shard_range chunks, all replicas concurrently.
// NCCL is resolved with dlopen at pool creation, so libsnb200.so itself has no link-time dependency on it; a pool of
// more than one GPU without a loadable libnccl.so.2 fails loudly (SNB_ERR_CUDA), it does not fall back to host copies.
#include <dlfcn.h>
#include <string.h>

#include <fstream>
#include <memory>

#include "../host/dispatch.h"
#include "net.h"

namespace {

// the slice of nccl.h this file uses (types kept opaque so no NCCL header is needed to build)
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
  void* h = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(char* err, size_t n) {
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) { snprintf(err, n, "snb_pool_create: cannot load libnccl.so.2 (%s)", dlerror()); return false; }
    auto sym = [&](const char* name) { return dlsym(h, name); };
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    Broadcast = (decltype(Broadcast))sym("ncclBroadcast");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !Broadcast || !GetErrorString) {
      snprintf(err, n, "snb_pool_create: libnccl.so.2 lacks a required symbol");
      return false;
    }
    return true;
  }
};
constexpr int kNcclUint8 = 1;    // ncclUint8 / ncclChar family: ncclInt8 = 0, ncclUint8 = 1 (nccl.h ncclDataType_t)

thread_local char g_pool_err[512] = {0};

}  // namespace

struct snb_pool {
  std::vector<snb_ctx*> ctx;
  std::vector<int> devices;
  std::mutex mu;
  std::unique_ptr<snb::Dispatcher> disp;
  std::vector<int64_t> calls;          // calls served per replica
  double bcast_ms = 0;
  uint64_t blob_bytes = 0;
  char err[512] = {0};
};

namespace {

struct PoolCall { snb_pool* pool; int replica; snb_done_fn done; void* user; };

void pool_done(void* user, int status, const snb_rt_stat* stat) {
  std::unique_ptr<PoolCall> pc(static_cast<PoolCall*>(user));
  {
    std::lock_guard<std::mutex> lk(pc->pool->mu);
    pc->pool->disp->done(pc->replica);
  }
  if (pc->done) pc->done(pc->user, status, stat);
}

int pool_fail(snb_pool* p, int code, const char* msg) {
  if (p) snprintf(p->err, sizeof(p->err), "%s", msg);
  snprintf(g_pool_err, sizeof(g_pool_err), "%s", msg);
  return code;
}

}  // namespace

extern "C" {

int snb_shard_range(int64_t n_pairs, int32_t world, int32_t rank, int64_t* start, int64_t* stop) {
  return snb::shard_range(n_pairs, world, rank, start, stop) ? SNB_OK : SNB_ERR_INVALID;
}

const char* snb_pool_last_error(const snb_pool* p) { return p ? p->err : g_pool_err; }

void snb_pool_destroy(snb_pool* p) {
  if (!p) return;
  for (snb_ctx* c : p->ctx) snb_destroy(c);
  delete p;
}

int snb_pool_create(snb_pool** out, const snb_config* cfg, const int32_t* devices, int32_t n_devices) {
  if (!out || !cfg || cfg->struct_size != (int32_t)sizeof(snb_config)) return pool_fail(nullptr, SNB_ERR_INVALID, "snb_pool_create: bad config struct");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return pool_fail(nullptr, SNB_ERR_CUDA, "snb_pool_create: no CUDA device (this library has no CPU fallback)");
  std::unique_ptr<snb_pool> p(new snb_pool());
  if (devices && n_devices > 0) p->devices.assign(devices, devices + n_devices);
  else for (int i = 0; i < ndev; ++i) p->devices.push_back(i);
  for (size_t i = 0; i < p->devices.size(); ++i) {
    if (p->devices[i] < 0 || p->devices[i] >= ndev) return pool_fail(nullptr, SNB_ERR_INVALID, "snb_pool_create: bad device ordinal");
    for (size_t j = 0; j < i; ++j)
      if (p->devices[j] == p->devices[i]) return pool_fail(nullptr, SNB_ERR_INVALID, "snb_pool_create: a device is listed twice");
  }
  const int n = (int)p->devices.size();
  auto bail = [&](int code, const char* msg) {
    for (snb_ctx* c : p->ctx) snb_destroy(c);
    p->ctx.clear();
    return pool_fail(nullptr, code, msg);
  };
  // replica 0: the model file is read once, by the rank that owns it
  std::vector<char> blob;
  snb_config c0 = *cfg;
  c0.device = p->devices[0];
  if (!(cfg->weights && cfg->weights_bytes)) {
    if (!cfg->model_file) return pool_fail(nullptr, SNB_ERR_MODEL, "snb_pool_create: neither model_file nor weights given");
    std::ifstream f(cfg->model_file, std::ios::binary);
    if (!f) { snprintf(g_pool_err, sizeof(g_pool_err), "File is not exist! model_file: %s", cfg->model_file); return SNB_ERR_MODEL; }
    blob.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    c0.weights = blob.data(); c0.weights_bytes = blob.size();
  }
  p->blob_bytes = c0.weights_bytes;
  snb_ctx* root = nullptr;
  int r = snb_create(&root, &c0);
  if (r != SNB_OK) return pool_fail(nullptr, r, snb_last_error(nullptr));
  p->ctx.push_back(root);
  if (n > 1) {
    NcclApi nccl;
    if (!nccl.load(g_pool_err, sizeof(g_pool_err))) return bail(SNB_ERR_CUDA, g_pool_err);
    for (int i = 1; i < n; ++i) {
      snb_config ci = *cfg;
      ci.device = p->devices[i];
      ci.flags |= SNB_FLAG_DEFER_WEIGHTS;
      ci.model_file = nullptr; ci.weights = nullptr; ci.weights_bytes = 0;
      snb_ctx* c = nullptr;
      r = snb_create(&c, &ci);
      if (r != SNB_OK) return bail(r, snb_last_error(nullptr));
      p->ctx.push_back(c);
    }
    // the single collective of this workload: blob in replica 0's HBM -> every GPU's HBM
    std::vector<void*> dbuf(n, nullptr);
    std::vector<cudaStream_t> st(n, nullptr);
    std::vector<ncclComm_t> comm(n, nullptr);
    auto cleanup = [&] {
      for (int i = 0; i < n; ++i) {
        cudaSetDevice(p->devices[i]);
        if (comm[i]) nccl.CommDestroy(comm[i]);
        if (dbuf[i]) cudaFree(dbuf[i]);
        if (st[i]) cudaStreamDestroy(st[i]);
      }
    };
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
      ok = cudaSetDevice(p->devices[i]) == cudaSuccess && cudaMalloc(&dbuf[i], p->blob_bytes) == cudaSuccess &&
           cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) == cudaSuccess;
    }
    if (ok) { cudaSetDevice(p->devices[0]); ok = cudaMemcpy(dbuf[0], c0.weights, p->blob_bytes, cudaMemcpyHostToDevice) == cudaSuccess; }
    if (!ok) { cleanup(); return bail(SNB_ERR_NOMEM, "snb_pool_create: staging the weight blob on the devices failed"); }
    int nr = nccl.CommInitAll(comm.data(), n, p->devices.data());
    if (nr != 0) { snprintf(g_pool_err, sizeof(g_pool_err), "ncclCommInitAll: %s", nccl.GetErrorString(nr)); cleanup(); return bail(SNB_ERR_CUDA, g_pool_err); }
    cudaEvent_t e0, e1;
    cudaSetDevice(p->devices[0]);
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st[0]);
    nccl.GroupStart();
    for (int i = 0; i < n && nr == 0; ++i) {
      cudaSetDevice(p->devices[i]);
      nr = nccl.Broadcast(dbuf[0], dbuf[i], p->blob_bytes, kNcclUint8, 0, comm[i], st[i]);
    }
    const int ge = nccl.GroupEnd();
    if (nr == 0) nr = ge;
    cudaSetDevice(p->devices[0]);
    cudaEventRecord(e1, st[0]);
    for (int i = 0; i < n; ++i) { cudaSetDevice(p->devices[i]); if (cudaStreamSynchronize(st[i]) != cudaSuccess) ok = false; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    p->bcast_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (nr != 0 || !ok) {
      snprintf(g_pool_err, sizeof(g_pool_err), "ncclBroadcast of the weight blob failed: %s", nr ? nccl.GetErrorString(nr) : "stream error");
      cleanup();
      return bail(SNB_ERR_CUDA, g_pool_err);
    }
    for (int i = 1; i < n && r == SNB_OK; ++i) {
      cudaSetDevice(p->devices[i]);
      r = snb_set_weights(p->ctx[i], dbuf[i], p->blob_bytes, 1);
      if (r != SNB_OK) snprintf(g_pool_err, sizeof(g_pool_err), "replica %d: %s", i, snb_last_error(p->ctx[i]));
    }
    cleanup();
    if (r != SNB_OK) return bail(r, g_pool_err);
  }
  p->disp.reset(new snb::Dispatcher(n));
  p->calls.assign(n, 0);
  *out = p.release();
  return SNB_OK;
}

int32_t snb_pool_size(const snb_pool* p) { return p ? (int32_t)p->ctx.size() : SNB_ERR_INVALID; }
snb_ctx* snb_pool_ctx(const snb_pool* p, int32_t replica) {
  return p && replica >= 0 && replica < (int32_t)p->ctx.size() ? p->ctx[replica] : nullptr;
}

static int pool_submit(snb_pool* p, const int8_t* in, const uint8_t* frames, int32_t* out, int32_t batch, snb_done_fn done, void* user,
                       int32_t timeout_ms) {
  if (!p || (!in && !frames) || !out || batch < 1) return pool_fail(p, SNB_ERR_INVALID, "snb_pool_infer_async: bad arguments");
  int rep;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    rep = p->disp->pick();
    ++p->calls[rep];
  }
  PoolCall* pc = new PoolCall{p, rep, done, user};
  const int r = frames ? snb_infer_nv12_async(p->ctx[rep], frames, out, batch, pool_done, pc, timeout_ms)
                       : snb_infer_async(p->ctx[rep], in, out, batch, pool_done, pc, timeout_ms);
  if (r != SNB_OK) {
    {
      std::lock_guard<std::mutex> lk(p->mu);
      p->disp->done(rep);
      --p->calls[rep];
    }
    delete pc;
    return pool_fail(p, r, snb_last_error(p->ctx[rep]));
  }
  return SNB_OK;
}

int snb_pool_infer_async(snb_pool* p, const int8_t* in, int32_t* out, int32_t batch, snb_done_fn done, void* user, int32_t timeout_ms) {
  return pool_submit(p, in, nullptr, out, batch, done, user, timeout_ms);
}

int snb_pool_infer_nv12_async(snb_pool* p, const uint8_t* frames, int32_t* out, int32_t batch, snb_done_fn done, void* user,
                              int32_t timeout_ms) {
  return pool_submit(p, nullptr, frames, out, batch, done, user, timeout_ms);
}

int snb_pool_wait_all(snb_pool* p) {
  if (!p) return SNB_ERR_INVALID;
  for (snb_ctx* c : p->ctx) snb_wait_all(c);
  return SNB_OK;
}

// One call, `batch` pairs: contiguous shards (shard_range), every replica works on its shard concurrently.
int snb_pool_infer(snb_pool* p, const int8_t* in, int32_t* out, int32_t batch) {
  if (!p || !in || !out || batch < 1) return pool_fail(p, SNB_ERR_INVALID, "snb_pool_infer: bad arguments");
  const int n = (int)p->ctx.size();
  snb_tensor_props pi, po;
  snb_get_io(p->ctx[0], &pi, &po);
  struct Flag { std::mutex mu; std::condition_variable cv; int pending = 0; int status = SNB_OK; } flag;
  auto done = [](void* user, int status, const snb_rt_stat*) {
    Flag* f = static_cast<Flag*>(user);
    std::lock_guard<std::mutex> lk(f->mu);
    if (status != SNB_OK) f->status = status;
    --f->pending;
    f->cv.notify_all();
  };
  int rc = SNB_OK;
  for (int i = 0; i < n; ++i) {
    int64_t a, b;
    snb::shard_range(batch, n, i, &a, &b);
    if (b <= a) continue;
    { std::lock_guard<std::mutex> lk(flag.mu); ++flag.pending; }
    const int r = snb_infer_async(p->ctx[i], in + (size_t)a * pi.mem_size, out + (size_t)a * (po.mem_size / 4), (int32_t)(b - a), done, &flag, -1);
    if (r != SNB_OK) {
      std::lock_guard<std::mutex> lk(flag.mu);
      --flag.pending;
      rc = r;
      snprintf(p->err, sizeof(p->err), "replica %d: %s", i, snb_last_error(p->ctx[i]));
    } else {
      std::lock_guard<std::mutex> lk(p->mu);
      ++p->calls[i];
    }
  }
  std::unique_lock<std::mutex> lk(flag.mu);
  flag.cv.wait(lk, [&] { return flag.pending == 0; });
  return rc != SNB_OK ? rc : flag.status;
}

int snb_pool_get_stat(const snb_pool* p, snb_pool_stat* s) {
  if (!p || !s) return SNB_ERR_INVALID;
  memset(s, 0, sizeof(*s));
  s->n_devices = (int32_t)p->ctx.size();
  s->weight_bytes = p->blob_bytes;
  s->broadcast_ms = (float)p->bcast_ms;
  for (size_t i = 0; i < p->ctx.size() && i < 16; ++i) { s->device[i] = p->devices[i]; s->calls[i] = p->calls[i]; }
  return SNB_OK;
}

}  // extern "C"
