// Probe: can a NO-SWIZZLE K-major UMMA shared-memory descriptor address OVERLAPPING rows?
//   canonical no-swizzle K-major layout: core matrix = 8 rows x 16 bytes (row pitch 16 B), next 8-row group at +SBO,
//   next 16-byte K chunk at +LBO.  With SBO = 128 B and LBO = 16 B the address of (row m, chunk j) is
//   start + 16*(m + j): row m of the A operand is the 32-byte WINDOW of a raw byte stream starting at 16*m -
//   an im2col of a stride-2, 4-channel f16 image row with no copy (the stem convolutions: one output pixel per 16 B).
// Test: raw[i] = i (f16, exact), B = selector so that D[m][n] = A[m][n (+8)] = raw[8*m + n (+8)].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_window umma_window.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  return d;                        // layout_type 0 = no swizzle
}
__device__ __forceinline__ uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode 0: canonical no-swizzle A (LBO 128 between K chunks... stored accordingly)   mode 1: sliding window (LBO 16, SBO 128)
__global__ void __launch_bounds__(128) probe(int mode, int bsel, float* out) {
  __shared__ __align__(1024) __half raw[4096];
  __shared__ __align__(1024) __half bmat[8 * 16];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 4096; i += 128) raw[i] = __float2half(0.f);
  __syncthreads();
  if (mode == 0) {
    // A[m][k] = 8*m + k stored canonically: chunk j = k/8 at byte j*2048 (LBO), group g = m/8 at g*128 (SBO), row r at r*16
    for (int i = tid; i < 128 * 16; i += 128) {
      const int m = i / 16, k = i % 16;
      const int off_bytes = (k / 8) * 2048 + (m / 8) * 128 + (m % 8) * 16 + (k % 8) * 2;
      raw[off_bytes / 2] = __float2half((float)((8 * m + k) % 2048));
    }
  } else {
    for (int i = tid; i < 4096; i += 128) raw[i] = __float2half((float)(i % 2048));
  }
  // B[n][k] = (k == n + bsel): canonical no-swizzle, one 8-row group, two K chunks 128 B apart
  for (int i = tid; i < 8 * 16; i += 128) {
    const int n = i / 16, k = i % 16;
    bmat[((k / 8) * 128 + n * 16 + (k % 8) * 2) / 2] = __float2half(k == n + bsel ? 1.f : 0.f);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint64_t a = mode == 0 ? desc_noswz(smem_u32(raw), 2048, 128) : desc_noswz(smem_u32(raw), 16, 128);
    const uint64_t b = desc_noswz(smem_u32(bmat), 128, 128);
    const uint32_t id = idesc_f16(128, 8);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem), "l"(a), "l"(b), "r"(id), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMA
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[8];
  const uint32_t taddr = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 8; ++n) out[tid * 8 + n] = __uint_as_float(v[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  float* d; cudaMalloc(&d, 128 * 8 * 4);
  float h[128 * 8];
  int bad_total = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int bsel = 0; bsel <= 8; bsel += 8) {
      probe<<<1, 128>>>(mode, bsel, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d bsel %d: %s\n", mode, bsel, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 8; ++n) {
          const float want = (float)((8 * m + n + bsel) % 2048);
          if (h[m * 8 + n] != want) { if (bad < 4) printf("  mode %d bsel %d: D[%d][%d] = %g, want %g\n", mode, bsel, m, n, h[m * 8 + n], want); ++bad; }
        }
      printf("%s A operand, B selects k = n + %d: %s (%d mismatches)\n", mode == 0 ? "canonical no-swizzle" : "sliding window (LBO 16 B, SBO 128 B)",
             bsel, bad ? "MISMATCH" : "exact", bad);
      bad_total += bad;
    }
  return bad_total ? 2 : 0;
}
