// Handle registry and API glue shared by the extern "C" translation units.
#pragma once
#include <memory>
#include "common.cuh"
#include "../../include/zkb200.h"

namespace zk {

struct HandleBase {
  int kind = 0;  // 1 = G1 table, 2 = G2 table, 3 = QAP, 4 = Groth16 key, 5 = Pinocchio key
  virtual ~HandleBase() {}
};
uint64_t register_handle(std::unique_ptr<HandleBase> h);
HandleBase* lookup_handle(uint64_t id, int kind);  // throws ZK_EARG when missing / wrong kind
void drop_handle(uint64_t id);
void require_init();

}  // namespace zk

#define ZK_API_BEGIN try { zk::require_init();
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
