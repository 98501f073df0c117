// Host-side plumbing shared by every translation unit of libzkb200: error
// reporting, stream-ordered device buffers, vector load/store and the byte
// formats of the C ABI (include/zkb200.h).
#pragma once
// Bounds checks of our own (compute-sanitizer is not available on the B200 pool): a build with
// -DZK_CHECKED (`make CHECKED=1` -> libzkb200_checked.so, selected with ZKB200_LIB) turns every
// ZK_DCHECK into a device-side assert — an out-of-range index then stops the kernel with file and
// line on stderr and the next API call fails — and costs nothing in the product build.
#ifdef ZK_CHECKED
#include <assert.h>
#define ZK_DCHECK(cond) assert(cond)
#else
#define ZK_DCHECK(cond) ((void)0)
#endif
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "ec.cuh"

// error codes of include/zkb200.h
#define ZK_OK 0
#define ZK_EARG (-1)
#define ZK_EPOINT (-2)
#define ZK_ECUDA (-3)
#define ZK_EREMAINDER (-4)

namespace zk {

void set_error(const std::string& s);
const std::string& last_error();
struct Error {
  int code;
  std::string msg;
};

#define ZK_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      throw zk::Error{ZK_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)};       \
  } while (0)
#define ZK_REQUIRE(cond, code, msg) \
  do {                              \
    if (!(cond)) throw zk::Error{code, msg}; \
  } while (0)

// library-wide stream (created by zk_init; every kernel of a call is ordered on it
// unless the caller passes its own stream to a *_dev entry point)
cudaStream_t default_stream();
cudaStream_t fork_aux(cudaStream_t st);   // auxiliary stream, ordered after the work enqueued on st so far
void join_aux(cudaStream_t st);           // st waits for the auxiliary stream
int sm_count();

// RAII device buffer
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) ZK_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
  }
  void ensure(size_t count) {
    if (count > n) alloc(count);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// RAII pinned host buffer
struct PinnedBuf {
  uint8_t* p = nullptr;
  size_t n = 0;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  void ensure(size_t bytes) {
    if (bytes <= n) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    ZK_CUDA(cudaMallocHost((void**)&p, bytes));
    n = bytes;
  }
};

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace zk

// -------------------------------------------------------------------------
// device helpers
// -------------------------------------------------------------------------
#ifndef ZK_HOST_SIM
// 128-bit vectorised load / store of a limb struct (sizeof(T) % 16 == 0, 16-B aligned)
template <class T>
__device__ __forceinline__ T load_vec(const T* p) {
  static_assert(sizeof(T) % 16 == 0, "limb structs are multiples of 16 bytes");
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(s + i);
  return r;
}
template <class T>
__device__ __forceinline__ T load_vec_rw(const T* p) {  // for buffers written in the same launch sequence
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
  return r;
}
template <class T>
__device__ __forceinline__ void store_vec(T* p, const T& v) {
  uint4* d = reinterpret_cast<uint4*>(p);
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}

// Streaming variants (ld/st.global.cs: evict-first): for data that is written once and read once per
// join — sorted entries, bucket sums, partial pieces — so that it does not push the randomly
// gathered base table out of L2.
template <class T>
__device__ __forceinline__ T load_vec_stream(const T* p) {
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldcs(s + i);
  return r;
}
template <class T>
__device__ __forceinline__ void store_vec_stream(T* p, const T& v) {
  uint4* d = reinterpret_cast<uint4*>(p);
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) __stcs(d + i, s[i]);
}

// ---- byte formats (zcash / blst, see include/zkb200.h) -----------------------
// 48 big-endian bytes -> raw (non-Montgomery) limbs; returns false if >= p
__device__ __forceinline__ bool fp_from_be(const uint8_t* b, Fp& out, uint8_t mask0 = 0xff) {
#pragma unroll
  for (int k = 0; k < 12; k++) {
    uint32_t b0 = b[4 * k] & (k == 0 ? mask0 : 0xff);
    out.v[11 - k] = (b0 << 24) | ((uint32_t)b[4 * k + 1] << 16) | ((uint32_t)b[4 * k + 2] << 8) | b[4 * k + 3];
  }
  return out.is_canonical_raw();
}
__device__ __forceinline__ void fp_to_be(const Fp& raw, uint8_t* b) {
#pragma unroll
  for (int k = 0; k < 12; k++) {
    uint32_t w = raw.v[11 - k];
    b[4 * k] = (uint8_t)(w >> 24);
    b[4 * k + 1] = (uint8_t)(w >> 16);
    b[4 * k + 2] = (uint8_t)(w >> 8);
    b[4 * k + 3] = (uint8_t)w;
  }
}
// raw value > (p-1)/2
__device__ __forceinline__ bool fp_gt_half(const Fp& raw) {
  uint32_t h[12];
#pragma unroll
  for (int i = 0; i < 12; i++) h[i] = FpParams::half(i);
  return !Fp::geq_raw(h, raw.v);
}

// Group traits: parsing / serialising / curve constant for G1 and G2.
struct G1Traits {
  typedef Fp F;
  static constexpr int RAW = 96;    // uncompressed bytes
  static constexpr int COMP = 48;   // compressed bytes
  static constexpr int ID = 1;
  static __device__ __forceinline__ Fp curve_b() {
    Fp b;
#pragma unroll
    for (int i = 0; i < 12; i++) b.v[i] = FpParams::b4(i);
    return b;
  }
  static __device__ __forceinline__ Affine<Fp> generator() {
    Affine<Fp> g;
#pragma unroll
    for (int i = 0; i < 12; i++) { g.x.v[i] = FpParams::g1x(i); g.y.v[i] = FpParams::g1y(i); }
    return g;
  }
  // returns 0 ok, 1 bad encoding / not on curve
  static __device__ int parse(const uint8_t* b, Affine<Fp>& out) {
    if (b[0] & 0x40) {  // identity: 0x40 then zeros
      bool ok = b[0] == 0x40;
      for (int i = 1; i < RAW; i++) ok = ok && b[i] == 0;
      out = Affine<Fp>::inf();
      return ok ? 0 : 1;
    }
    if (b[0] & 0xe0) return 1;
    Fp x, y;
    if (!fp_from_be(b, x) || !fp_from_be(b + 48, y)) return 1;
    out.x = x.to_mont();
    out.y = y.to_mont();
    return (out.y.sqr() == out.x.sqr() * out.x + curve_b()) ? 0 : 1;
  }
  // writes RAW uncompressed bytes then COMP compressed bytes
  static __device__ void serialize(const Affine<Fp>& p, uint8_t* out) {
    if (p.is_inf()) {
      for (int i = 0; i < RAW + COMP; i++) out[i] = 0;
      out[0] = 0x40;
      out[RAW] = 0xc0;
      return;
    }
    Fp x = p.x.from_mont(), y = p.y.from_mont();
    fp_to_be(x, out);
    fp_to_be(y, out + 48);
    fp_to_be(x, out + RAW);
    out[RAW] |= 0x80 | (fp_gt_half(y) ? 0x20 : 0);
  }
};

struct G2Traits {
  typedef Fp2 F;
  static constexpr int RAW = 192;
  static constexpr int COMP = 96;
  static constexpr int ID = 2;
  static __device__ __forceinline__ Fp2 curve_b() { return Fp2{G1Traits::curve_b(), G1Traits::curve_b()}; }
  static __device__ __forceinline__ Affine<Fp2> generator() {
    Affine<Fp2> g;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      g.x.c0.v[i] = FpParams::g2x0(i); g.x.c1.v[i] = FpParams::g2x1(i);
      g.y.c0.v[i] = FpParams::g2y0(i); g.y.c1.v[i] = FpParams::g2y1(i);
    }
    return g;
  }
  static __device__ int parse(const uint8_t* b, Affine<Fp2>& out) {
    if (b[0] & 0x40) {
      bool ok = b[0] == 0x40;
      for (int i = 1; i < RAW; i++) ok = ok && b[i] == 0;
      out = Affine<Fp2>::inf();
      return ok ? 0 : 1;
    }
    if (b[0] & 0xe0) return 1;
    Fp x1, x0, y1, y0;  // wire order: x.c1 | x.c0 | y.c1 | y.c0
    if (!fp_from_be(b, x1) || !fp_from_be(b + 48, x0) || !fp_from_be(b + 96, y1) || !fp_from_be(b + 144, y0)) return 1;
    out.x = Fp2{x0.to_mont(), x1.to_mont()};
    out.y = Fp2{y0.to_mont(), y1.to_mont()};
    return (out.y.sqr() == out.x.sqr() * out.x + curve_b()) ? 0 : 1;
  }
  static __device__ void serialize(const Affine<Fp2>& p, uint8_t* out) {
    if (p.is_inf()) {
      for (int i = 0; i < RAW + COMP; i++) out[i] = 0;
      out[0] = 0x40;
      out[RAW] = 0xc0;
      return;
    }
    Fp x0 = p.x.c0.from_mont(), x1 = p.x.c1.from_mont();
    Fp y0 = p.y.c0.from_mont(), y1 = p.y.c1.from_mont();
    fp_to_be(x1, out);
    fp_to_be(x0, out + 48);
    fp_to_be(y1, out + 96);
    fp_to_be(y0, out + 144);
    fp_to_be(x1, out + RAW);
    fp_to_be(x0, out + RAW + 48);
    bool big = y1.is_zero() ? fp_gt_half(y0) : fp_gt_half(y1);  // c1 first, then c0
    out[RAW] |= 0x80 | (big ? 0x20 : 0);
  }
};
#endif  // !ZK_HOST_SIM
