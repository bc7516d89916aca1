// Shared helpers for libpvb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/pv_b200.h"

namespace pv {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define PV_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      pv::set_error(__VA_ARGS__);          \
      return PV_ERR_INVALID;               \
    }                                      \
  } while (0)

#define PV_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      pv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,      \
                    __LINE__);                                                             \
      return PV_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define PV_LAUNCH_OK(name)                                                                 \
  do {                                                                                     \
    cudaError_t _e = cudaPeekAtLastError();                                                \
    if (_e != cudaSuccess) {                                                               \
      pv::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));              \
      (void)cudaGetLastError();                                                            \
      return PV_ERR_CUDA;                                                                  \
    }                                                                                      \
    pv::count_launch();                                                                    \
  } while (0)

static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// ---- per-device launch state -------------------------------------------------------------------
// Function attributes (opt-in dynamic shared memory) and the SM count are PER DEVICE: a process that runs a
// plan on cuda:0 and later on cuda:1 must opt in again on the second device.  One flag word per launch site,
// one bit per device ordinal (devices >= 64 simply re-set the attribute on every launch).
struct DeviceOnce {
  unsigned long long done = 0ull;
  // true exactly once per (site, current device)
  bool first(int dev) {
    if (dev < 0 || dev >= 64) return true;
    const unsigned long long bit = 1ull << dev;
    if (done & bit) return false;
    done |= bit;
    return true;
  }
};
static inline int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}
// SM count of the CURRENT device (cached per ordinal)
static inline int current_sm_count() {
  static int cache[64] = {0};
  const int dev = current_device();
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  if (dev < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 0;
  if (dev < 64) cache[dev] = n;
  return n;
}
#define PV_OPT_IN_SMEM(kernel, bytes)                                                                   \
  do {                                                                                                  \
    static pv::DeviceOnce _once;                                                                        \
    if (_once.first(pv::current_device()))                                                              \
      PV_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
  } while (0)

// ---- element helpers ------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
};
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};

// 4-element vector load / store with fp32 maths
template <typename T> __device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<__half>(const __half* p, float (&v)[4]) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&r.x), b = *reinterpret_cast<__half2*>(&r.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  float4 r = *reinterpret_cast<const float4*>(p);
  v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <typename T> __device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void st4<__half>(__half* p, const float (&v)[4]) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// 8-element vector load / store
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<__half>(const __half* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<__half>(__half* p, const float (&v)[8]) {
  uint4 r;
  __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case PV_ACT_RELU: return fmaxf(x, 0.f);
    case PV_ACT_SWISH: return x / (1.f + __expf(-x));
    case PV_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    case PV_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    default: return x;
  }
}

}  // namespace pv
