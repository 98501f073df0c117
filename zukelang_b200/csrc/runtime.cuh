// Handle registry, device contexts and API glue shared by the extern "C" translation units.
//
// Devices.  zk_init(device) drives one GPU; zk_init_devices(devs, ndev) drives up to 8 GPUs of one
// box from ONE host thread ("single process, <= 8 devices").  Every device has a
// context (its own stream, auxiliary stream and fork/join events); context 0 is the primary device:
// QAP / domain handles live there, results are combined there.  Key and table handles loaded while
// several devices are driven spread their base-point ranges over all of them (§8e); the partial
// sums travel to the primary device as peer-to-peer stores over NVLink (peer access is enabled
// between the primary and every other device at init).
#pragma once
#include <memory>
#include "common.cuh"
#include "../../include/zkb200.h"

namespace zk {

constexpr int MAX_DEVICES = 8;
int env_int(const char* name, int dflt);  // integer environment knob (runtime.cu)

struct HandleBase {
  int kind = 0;  // 1 = G1 table, 2 = G2 table, 3 = QAP, 4 = Groth16 key, 5 = Pinocchio key, 6 = evaluation domain
  int ctx = 0;   // device context the handle's (primary) buffers live on
  virtual ~HandleBase() {}
};
uint64_t register_handle(std::unique_ptr<HandleBase> h);
HandleBase* lookup_handle(uint64_t id, int kind);  // throws ZK_EARG when missing / wrong kind; switches to the handle's context
void drop_handle(uint64_t id);

// ---- device contexts ---------------------------------------------------------------
int device_count();              // devices driven by this process
int current_ctx();               // context the calling thread is on
void set_ctx(int ctx);           // cudaSetDevice + thread-local index
cudaStream_t stream_of(int ctx);
void sync_all_devices();         // cudaDeviceSynchronize on every driven device (before freeing a handle)
struct CtxScope {                // switch for a scope, restore on exit (also on unwind)
  int prev;
  explicit CtxScope(int ctx) : prev(current_ctx()) { set_ctx(ctx); }
  ~CtxScope() { set_ctx(prev); }
  CtxScope(const CtxScope&) = delete;
  CtxScope& operator=(const CtxScope&) = delete;
};

// Entered by every extern "C" entry point: serialises API calls (the library has one stream per
// device and per-handle workspaces, so concurrent callers would interleave their enqueues),
// checks zk_init, starts on the primary device and gives the caller's CUDA device back on exit.
struct ApiGuard {
  int caller_device = -1;
  ApiGuard();
  ~ApiGuard();
};

}  // namespace zk

#define ZK_API_BEGIN try { zk::ApiGuard zk_api_guard_;
#define ZK_API_END                                   \
  }                                                  \
  catch (const zk::Error& e) {                       \
    zk::set_error(e.msg);                            \
    return e.code;                                   \
  }                                                  \
  catch (const std::exception& e) {                  \
    zk::set_error(e.what());                         \
    return ZK_ECUDA;                                 \
  }                                                  \
  return ZK_OK;
