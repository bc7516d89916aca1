// Probe: does cuTensorMapEncodeTiled accept OVERLAPPING global strides (stride[1] < extent of dim0),
// and does the TMA then deliver sliding windows?  (needed for the TMA-fed stem im2col trick)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, __half* out, int c0, int c1, int c2, int nbytes) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(nbytes));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)&tm), "r"(bar_a), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.u32 %0,1,0,P;}" : "=r"(ok) : "r"(bar_a));
  for (int i = threadIdx.x; i < nbytes / 2; i += blockDim.x) out[i] = ((__half*)sm)[i];
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int W = 128, H = 10;   // rows of W pixels, 4 ch each (8 B per pixel)
  std::vector<__half> h(W * 4 * H + 64);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half((float)(i % 2048));
  __half* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  for (int swz = 0; swz < 2; ++swz) {
    CUtensorMap tm;
    cuuint64_t gdim[3] = {32, 60, (cuuint64_t)H};            // window of 32 elems, 60 windows (stride 2 px), H rows
    cuuint64_t gstr[2] = {16, (cuuint64_t)W * 8};            // window stride 16 B  (< 64 B window extent: OVERLAP)
    cuuint32_t box[3] = {32, 8, 2}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle=%d encode overlapping strides -> CUresult %d\n", swz, (int)r);
    if (r != CUDA_SUCCESS) continue;
    __half* o; cudaMalloc(&o, 32 * 8 * 2 * 2); cudaMemset(o, 0, 32 * 8 * 2 * 2);
    k<<<1, 128, 4096>>>(tm, o, 0, 5, 3, 32 * 8 * 2 * 2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<__half> r2(32 * 8 * 2); cudaMemcpy(r2.data(), o, r2.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int hh = 0; hh < 2; ++hh) for (int w = 0; w < 8; ++w) for (int e2 = 0; e2 < 32; ++e2) {
      size_t src = (size_t)(3 + hh) * W * 4 + (size_t)(5 + w) * 8 + e2;
      float want = (float)(src % 2048);
      int row = hh * 8 + w; int idx;
      if (!swz) idx = row * 32 + e2;
      else { int chunk = e2 / 8, within = e2 % 8; int sw = chunk ^ ((row >> 1) & 3); idx = row * 32 + sw * 8 + within; }
      if (__half2float(r2[idx]) != want) { if (bad < 5) printf("mismatch row %d e %d got %f want %f\n", row, e2, __half2float(r2[idx]), want); ++bad; }
    }
    printf("swizzle=%d sliding-window check: %s (%d mismatches)\n", swz, bad ? "FAIL" : "OK", bad);
  }
  return 0;
}
