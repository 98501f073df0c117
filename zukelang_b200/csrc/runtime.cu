// Library lifecycle: device selection, the library stream, last-error string, handles.
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include "runtime.cuh"

namespace zk {

static thread_local std::string g_error;
static cudaStream_t g_stream = nullptr;
static cudaStream_t g_aux_stream = nullptr;
static cudaEvent_t g_fork = nullptr, g_join = nullptr;
static bool g_init = false;
static int g_sm_count = 0;
static std::mutex g_mu;
static std::map<uint64_t, std::unique_ptr<HandleBase>> g_handles;
static uint64_t g_next_handle = 1;

void set_error(const std::string& s) { g_error = s; }
const std::string& last_error() { return g_error; }
cudaStream_t default_stream() { return g_stream; }
// Runs `on_aux(aux)` on the auxiliary stream, ordered after everything already enqueued on `st`; a
// later join_aux(st) makes `st` wait for it.  Used to finish the G2 tail next to the G1 tail.
cudaStream_t fork_aux(cudaStream_t st) {
  ZK_CUDA(cudaEventRecord(g_fork, st));
  ZK_CUDA(cudaStreamWaitEvent(g_aux_stream, g_fork, 0));
  return g_aux_stream;
}
void join_aux(cudaStream_t st) {
  ZK_CUDA(cudaEventRecord(g_join, g_aux_stream));
  ZK_CUDA(cudaStreamWaitEvent(st, g_join, 0));
}
int sm_count() { return g_sm_count; }

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

void require_init() {
  if (!g_init) throw Error{ZK_ECUDA, "zk_init has not been called (or found no CUDA device)"};
}

uint64_t register_handle(std::unique_ptr<HandleBase> h) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t id = g_next_handle++;
  g_handles[id] = std::move(h);
  return id;
}
HandleBase* lookup_handle(uint64_t id, int kind) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(id);
  if (it == g_handles.end()) throw Error{ZK_EARG, "unknown handle"};
  if (kind && it->second->kind != kind) throw Error{ZK_EARG, "handle is of a different kind"};
  return it->second.get();
}
void drop_handle(uint64_t id) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_handles.erase(id)) throw Error{ZK_EARG, "unknown handle"};
}

}  // namespace zk

extern "C" {

int zk_init(int device) {
  try {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw zk::Error{ZK_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                    " (libzkb200 has no CPU fallback)"};
    if (device >= 0) ZK_CUDA(cudaSetDevice(device));
    int dev = 0;
    ZK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    ZK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
      throw zk::Error{ZK_ECUDA, std::string("device ") + prop.name + " is not sm_100-class; libzkb200 is built for sm_100a only"};
    zk::g_sm_count = prop.multiProcessorCount;
    if (!zk::g_stream) {
      ZK_CUDA(cudaStreamCreateWithFlags(&zk::g_stream, cudaStreamNonBlocking));
      ZK_CUDA(cudaStreamCreateWithFlags(&zk::g_aux_stream, cudaStreamNonBlocking));
      ZK_CUDA(cudaEventCreateWithFlags(&zk::g_fork, cudaEventDisableTiming));
      ZK_CUDA(cudaEventCreateWithFlags(&zk::g_join, cudaEventDisableTiming));
    }
    zk::g_init = true;
  } catch (const zk::Error& e) {
    zk::set_error(e.msg);
    return e.code;
  }
  return ZK_OK;
}

int zk_shutdown(void) {
  {
    std::lock_guard<std::mutex> lk(zk::g_mu);
    zk::g_handles.clear();
  }
  if (zk::g_stream) {
    cudaStreamDestroy(zk::g_stream);
    cudaStreamDestroy(zk::g_aux_stream);
    cudaEventDestroy(zk::g_fork);
    cudaEventDestroy(zk::g_join);
    zk::g_stream = nullptr;
    zk::g_aux_stream = nullptr;
  }
  zk::g_init = false;
  return ZK_OK;
}

const char* zk_last_error(void) { return zk::last_error().c_str(); }

int zk_device_info(char* buf, size_t cap) {
  ZK_API_BEGIN
  int dev = 0;
  ZK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  ZK_CUDA(cudaGetDeviceProperties(&prop, dev));
  snprintf(buf, cap, "%s;%d;%d.%d", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  ZK_API_END
}

int zk_table_free(uint64_t handle) {
  ZK_API_BEGIN
  ZK_CUDA(cudaDeviceSynchronize());
  zk::drop_handle(handle);
  ZK_API_END
}

}  // extern "C"
