// Probe: MN-major B operand (b_major = 1 in the instruction descriptor) with the shared-memory layout a TMA tiled load
// produces for a [keys][d] tile (rows = K index, contiguous along N): SWIZZLE_128B (N = 64 f16 = 128-byte rows) and
// SWIZZLE_64B (N = 32).  Lets the P.V GEMM of attention read V as it lies in memory (no transposed copy).
//   D[m][n] = sum_k A[m][k] B[n][k];  A = selector (A[m][k] = (k == m), m < 16)  ->  D[m][n] = B[n][m] = V[key m][d n].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_mnmajor umma_mnmajor.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;           // 0 none, 2 = SW128, 4 = SW64, 6 = SW32
  return d;
}

// swz_bytes: 128 (N = 64) or 64 (N = 32).  variant: which (LBO, SBO) convention to try.
__global__ void __launch_bounds__(128) probe(int swz_bytes, int variant, float* out) {
  __shared__ __align__(1024) __half amat[128 * 16];     // canonical no-swizzle K-major: chunk j at j*2048, group g at g*128, row r at r*16
  __shared__ __align__(1024) __half bmat[16 * 64];      // [k rows][N] swizzled rows of swz_bytes
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  const int N = swz_bytes / 2;
  for (int i = tid; i < 128 * 16; i += 128) amat[i] = __float2half(0.f);
  for (int i = tid; i < 16 * 64; i += 128) bmat[i] = __float2half(0.f);
  __syncthreads();
  for (int i = tid; i < 128 * 16; i += 128) {
    const int m = i / 16, k = i % 16;
    amat[((k / 8) * 2048 + (m / 8) * 128 + (m % 8) * 16 + (k % 8) * 2) / 2] = __float2half(m == k ? 1.f : 0.f);
  }
  // V[k][n] = 64 * k + n, stored like TMA would: row k at k * swz_bytes, 16-byte chunk c at (c ^ f(k))
  for (int i = tid; i < 16 * N; i += 128) {
    const int k = i / N, n = i % N;
    const int c = n / 8;
    const int phys = swz_bytes == 128 ? (c ^ (k & 7)) : (c ^ ((k >> 1) & 3));
    bmat[(k * swz_bytes + phys * 16 + (n % 8) * 2) / 2] = __float2half((float)(64 * k + n));
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint64_t a = make_desc(smem_u32(amat), 2048, 128, 0);
    const uint32_t sbo8 = 8u * (uint32_t)swz_bytes;        // 8 K-rows further
    uint64_t b;
    if (variant == 0) b = make_desc(smem_u32(bmat), sbo8, sbo8, swz_bytes == 128 ? 2 : 4);
    else if (variant == 1) b = make_desc(smem_u32(bmat), 16, sbo8, swz_bytes == 128 ? 2 : 4);
    else b = make_desc(smem_u32(bmat), sbo8, 16, swz_bytes == 128 ? 2 : 4);
    const uint32_t id = (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // b_major = MN
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem), "l"(a), "l"(b), "r"(id), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 8; ++n) out[tid * 64 + c0 + n] = __uint_as_float(v[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  float* d; cudaMalloc(&d, 128 * 64 * 4);
  static float h[128 * 64];
  for (int swz = 128; swz >= 64; swz /= 2)
    for (int variant = 0; variant < 3; ++variant) {
      cudaMemset(d, 0, sizeof(h));
      probe<<<1, 128>>>(swz, variant, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("swz %d variant %d: %s\n", swz, variant, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      const int N = swz / 2;
      int bad = 0;
      for (int m = 0; m < 16; ++m)
        for (int n = 0; n < N; ++n) {
          const float want = (float)(64 * m + n);
          if (h[m * 64 + n] != want) { if (bad < 3) printf("  swz %d variant %d: D[%d][%d] = %g, want %g\n", swz, variant, m, n, h[m * 64 + n], want); ++bad; }
        }
      printf("MN-major B, SWIZZLE_%dB (N = %d), descriptor variant %d (0: LBO=SBO=8 rows, 1: LBO=16, 2: SBO=16): %s (%d mismatches)\n", swz, N, variant,
             bad ? "MISMATCH" : "exact", bad);
    }
  return 0;
}
