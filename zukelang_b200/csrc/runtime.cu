// Library lifecycle: device selection, the library stream, last-error string, handles.
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <vector>
#include "runtime.cuh"

namespace zk {

struct DeviceCtx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr, aux = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};

static thread_local std::string g_error;
static thread_local int tl_ctx = 0;
static std::vector<DeviceCtx> g_devs;
static bool g_init = false;
static std::recursive_mutex g_api_mu;    // one API call at a time
static std::mutex g_mu;                  // handle table
static std::map<uint64_t, std::unique_ptr<HandleBase>> g_handles;
static uint64_t g_next_handle = 1;

void set_error(const std::string& s) { g_error = s; }
const std::string& last_error() { return g_error; }

int device_count() { return (int)g_devs.size(); }
int current_ctx() { return tl_ctx; }
void set_ctx(int ctx) {
  if (ctx < 0 || ctx >= (int)g_devs.size()) throw Error{ZK_EARG, "device context out of range"};
  ZK_CUDA(cudaSetDevice(g_devs[ctx].device));
  tl_ctx = ctx;
}
cudaStream_t stream_of(int ctx) { return g_devs[ctx].stream; }
cudaStream_t default_stream() { return g_devs[tl_ctx].stream; }
// The auxiliary stream of the current device, ordered after everything already enqueued on `st`; a
// later join_aux(st) makes `st` wait for it.  Used to finish the G2 tail next to the G1 tail.
cudaStream_t fork_aux(cudaStream_t st) {
  DeviceCtx& d = g_devs[tl_ctx];
  ZK_CUDA(cudaEventRecord(d.fork, st));
  ZK_CUDA(cudaStreamWaitEvent(d.aux, d.fork, 0));
  return d.aux;
}
void join_aux(cudaStream_t st) {
  DeviceCtx& d = g_devs[tl_ctx];
  ZK_CUDA(cudaEventRecord(d.join, d.aux));
  ZK_CUDA(cudaStreamWaitEvent(st, d.join, 0));
}
int sm_count() { return g_devs[tl_ctx].sm_count; }

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

ApiGuard::ApiGuard() {
  g_api_mu.lock();
  if (!g_init) {
    g_api_mu.unlock();
    throw Error{ZK_ECUDA, "zk_init has not been called (or found no CUDA device)"};
  }
  cudaGetDevice(&caller_device);
  try {
    set_ctx(0);
  } catch (...) {
    g_api_mu.unlock();
    throw;
  }
}
ApiGuard::~ApiGuard() {
  if (caller_device >= 0) cudaSetDevice(caller_device);
  tl_ctx = 0;
  g_api_mu.unlock();
}

uint64_t register_handle(std::unique_ptr<HandleBase> h) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t id = g_next_handle++;
  g_handles[id] = std::move(h);
  return id;
}
HandleBase* lookup_handle(uint64_t id, int kind) {
  HandleBase* h;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_handles.find(id);
    if (it == g_handles.end()) throw Error{ZK_EARG, "unknown handle"};
    if (kind && it->second->kind != kind) throw Error{ZK_EARG, "handle is of a different kind"};
    h = it->second.get();
  }
  if (h->ctx != tl_ctx) set_ctx(h->ctx);
  return h;
}
void drop_handle(uint64_t id) {
  std::unique_ptr<HandleBase> victim;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_handles.find(id);
    if (it == g_handles.end()) throw Error{ZK_EARG, "unknown handle"};
    victim = std::move(it->second);
    g_handles.erase(it);
  }
}

void sync_all_devices() {
  const int keep = tl_ctx;
  for (int i = 0; i < (int)g_devs.size(); i++) {
    set_ctx(i);
    ZK_CUDA(cudaDeviceSynchronize());
  }
  set_ctx(keep);
}

static void destroy_devices() {
  for (DeviceCtx& d : g_devs) {
    if (cudaSetDevice(d.device) != cudaSuccess) continue;
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.aux) cudaStreamDestroy(d.aux);
    if (d.fork) cudaEventDestroy(d.fork);
    if (d.join) cudaEventDestroy(d.join);
  }
  g_devs.clear();
}

static int init_devices(const int* devs, int ndev) {
  std::lock_guard<std::recursive_mutex> api(g_api_mu);
  int caller = -1;
  cudaGetDevice(&caller);
  try {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw Error{ZK_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libzkb200 has no CPU fallback)"};
    if (ndev < 1 || ndev > MAX_DEVICES) throw Error{ZK_EARG, "zk_init_devices: 1 to 8 devices"};
    std::vector<int> want(devs, devs + ndev);
    for (int i = 0; i < ndev; i++) {
      if (want[i] < 0) {       // -1 = keep the current device (zk_init(-1))
        if (ndev != 1) throw Error{ZK_EARG, "zk_init_devices: negative device ordinal"};
        ZK_CUDA(cudaGetDevice(&want[i]));
      }
      if (want[i] >= count) throw Error{ZK_EARG, "zk_init_devices: device ordinal out of range"};
      for (int j = 0; j < i; j++)
        if (want[j] == want[i]) throw Error{ZK_EARG, "zk_init_devices: device listed twice"};
    }
    // same device list as before: nothing to do (zk_init is idempotent)
    bool same = g_init && (int)g_devs.size() == ndev;
    for (int i = 0; same && i < ndev; i++) same = g_devs[i].device == want[i];
    if (same) {
      ZK_CUDA(cudaSetDevice(g_devs[0].device));
      return ZK_OK;
    }
    {
      std::lock_guard<std::mutex> lk(g_mu);
      if (!g_handles.empty()) throw Error{ZK_EARG, "zk_init: the device list cannot change while handles are alive (zk_shutdown first)"};
    }
    destroy_devices();
    g_init = false;
    for (int i = 0; i < ndev; i++) {
      DeviceCtx d;
      d.device = want[i];
      ZK_CUDA(cudaSetDevice(d.device));
      cudaDeviceProp prop;
      ZK_CUDA(cudaGetDeviceProperties(&prop, d.device));
      if (prop.major < 10)
        throw Error{ZK_ECUDA, std::string("device ") + prop.name + " is not sm_100-class; libzkb200 is built for sm_100a only"};
      d.sm_count = prop.multiProcessorCount;
      ZK_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
      ZK_CUDA(cudaStreamCreateWithFlags(&d.aux, cudaStreamNonBlocking));
      ZK_CUDA(cudaEventCreateWithFlags(&d.fork, cudaEventDisableTiming));
      ZK_CUDA(cudaEventCreateWithFlags(&d.join, cudaEventDisableTiming));
      g_devs.push_back(d);
    }
    // peer access between the primary device and every other one, both ways: the shards read their
    // scalars from the primary device's memory and store their partial sums into it
    for (int i = 1; i < ndev; i++) {
      int ab = 0, ba = 0;
      ZK_CUDA(cudaDeviceCanAccessPeer(&ab, want[0], want[i]));
      ZK_CUDA(cudaDeviceCanAccessPeer(&ba, want[i], want[0]));
      if (!ab || !ba) throw Error{ZK_ECUDA, "zk_init_devices: no peer access between device " + std::to_string(want[0]) + " and device " + std::to_string(want[i])};
      for (int dir = 0; dir < 2; dir++) {
        ZK_CUDA(cudaSetDevice(dir ? want[i] : want[0]));
        cudaError_t pe = cudaDeviceEnablePeerAccess(dir ? want[0] : want[i], 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else ZK_CUDA(pe);
      }
    }
    ZK_CUDA(cudaSetDevice(g_devs[0].device));
    tl_ctx = 0;
    g_init = true;
  } catch (const Error& e) {
    set_error(e.msg);
    if (!g_init) destroy_devices();
    if (caller >= 0) cudaSetDevice(caller);
    return e.code;
  }
  return ZK_OK;
}

}  // namespace zk

extern "C" {

int zk_init(int device) { return zk::init_devices(&device, 1); }

int zk_init_devices(const int* devs, int ndev) {
  if (!devs) {
    zk::set_error("zk_init_devices: null device list");
    return ZK_EARG;
  }
  return zk::init_devices(devs, ndev);
}

int zk_device_count(void) { return zk::g_init ? zk::device_count() : 0; }

int zk_shutdown(void) {
  std::lock_guard<std::recursive_mutex> api(zk::g_api_mu);
  int caller = -1;
  cudaGetDevice(&caller);
  for (zk::DeviceCtx& d : zk::g_devs)
    if (cudaSetDevice(d.device) == cudaSuccess) cudaDeviceSynchronize();
  {
    std::lock_guard<std::mutex> lk(zk::g_mu);
    zk::g_handles.clear();
  }
  zk::destroy_devices();
  zk::g_init = false;
  if (caller >= 0) cudaSetDevice(caller);
  return ZK_OK;
}

const char* zk_last_error(void) { return zk::last_error().c_str(); }

int zk_device_info(char* buf, size_t cap) {
  ZK_API_BEGIN
  ZK_REQUIRE(buf && cap > 0, ZK_EARG, "device_info: null buffer");
  int dev = 0;
  ZK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  ZK_CUDA(cudaGetDeviceProperties(&prop, dev));
  snprintf(buf, cap, "%s;%d;%d.%d", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  ZK_API_END
}

int zk_table_free(uint64_t handle) {
  ZK_API_BEGIN
  zk::HandleBase* h = zk::lookup_handle(handle, 0);
  ZK_REQUIRE(h->kind == 1 || h->kind == 2, ZK_EARG, "table_free: not a table handle");
  zk::sync_all_devices();
  zk::drop_handle(handle);
  ZK_API_END
}

}  // extern "C"
