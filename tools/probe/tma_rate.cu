// Probe: per-SM TMA tiled-load throughput as a function of box shape (rows x row bytes, rank).
// One elected thread per CTA keeps SLOTS loads in flight; 148 CTAs; data L2-resident on the timed pass.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int SLOTS = 12;

struct Cfg {
  CUtensorMap map;
  int rank;          // 2 or 5
  int box_bytes;     // bytes landed per load
  int iters;
  int c1_step;       // coordinate step of dim1 per iteration
  int c1_mod;        // wrap of dim1 coordinate (in steps)
  int cta_c4;        // 5-D: use blockIdx % this as c4 (batch), 2-D: unused
  int cta_c1_off;    // 2-D: per-CTA offset in dim1
  int c0_chunks;     // 5-D: number of 64-channel chunks cycled in dim0
  int slots;         // loads in flight per issuing warp
  int issuers;       // issuing warps (1..4), each with its own slots
  int triv5;         // 1: rank-5 map [inner, rows, 1, 1, 1] addressed with the .5d instruction
  int store;         // 1: cp.async.bulk.tensor store (smem -> global) with bulk groups instead of loads
};

__global__ void __launch_bounds__(128) tma_rate_kernel(const __grid_constant__ Cfg P, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[SLOTS];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < P.issuers) {
    const int S = P.slots;
    const long long t0 = clock64();
    if (P.store) {
      for (int it = 0; it < P.iters; ++it) {
        const int s = warp * S + it % S;
        const uint32_t src = base + (uint32_t)s * 16384u;
        const int step = it % P.c1_mod;
        const int c1 = P.cta_c1_off * (int)blockIdx.x + (step * P.issuers + warp) * P.c1_step;
        if (P.triv5)
          asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&P.map)), "r"(src), "r"(0), "r"(c1), "r"(0), "r"(0), "r"(0) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&P.map)), "r"(src), "r"(0), "r"(c1) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
    for (int it = 0; it < P.iters + S; ++it) {
      const int sl = it % S;
      const int s = warp * S + sl;
      const uint32_t bar = smem_u32(&bars[s]);
      if (it >= S) {   // wait for the load issued S iterations ago
        const uint32_t parity = (uint32_t)(((it / S) - 1) & 1);
        uint32_t ok = 0;
        while (!ok) {
          asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0, 1, 0, P;\n}\n"
                       : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        }
      }
      if (it < P.iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)P.box_bytes) : "memory");
        const uint32_t dst = base + (uint32_t)s * 16384u;
        const int step = it % P.c1_mod;
        if (P.rank == 2) {
          const int c1 = P.cta_c1_off * (int)blockIdx.x + (step * P.issuers + warp) * P.c1_step;
          if (P.triv5)
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&P.map)), "r"(bar), "r"(0), "r"(c1), "r"(0), "r"(0), "r"(0) : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&P.map)), "r"(bar), "r"(0), "r"(c1) : "memory");
        } else {
          asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                       ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&P.map)), "r"(bar), "r"((step % P.c0_chunks) * 64), "r"(0), "r"((step >> 2) % 2 * 5),
                         "r"((step >> 3) % 8), "r"((int)blockIdx.x % P.cta_c4) : "memory");
        }
      }
    }
    }
    if (warp == 0) out[blockIdx.x] = clock64() - t0;
  }
}

static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  return (EncodeTiledFn)fn;
}

static void run(const char* name, Cfg& c, long long* d_out, int rows) {
  const size_t smem = SLOTS * 16384 + 1024;
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int pass = 0; pass < 2; ++pass) {
    tma_rate_kernel<<<148, 128, smem>>>(c, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; long long mx = 0;
  for (int i = 0; i < 148; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
  avg /= 148;
  const double per_box = avg / c.iters / c.issuers;     // SM-level clocks per box
  printf("%-40s %s%s slots %2d x%d | rows %4d box %6d B | %8.1f clk/box/SM | %6.1f B/clk/SM\n", name, c.store ? "ST" : "LD",
         c.triv5 ? "5" : " ", c.slots, c.issuers, rows, c.box_bytes, per_box, c.box_bytes / per_box);
}

int main() {
  EncodeTiledFn enc = get_encode();
  if (!enc) { printf("no encode fn\n"); return 1; }
  const size_t bytes = 256ull << 20;
  void* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const int iters = 2000;
  struct D2 { const char* name; int inner_elems; int pitch_bytes; int rows; CUtensorMapSwizzle sw; };
  D2 d2[] = {
    {"2D 128Bx128 rows, pitch 128 (contiguous)", 64, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B},
    {"2D 128Bx128 rows, pitch 512", 64, 512, 128, CU_TENSOR_MAP_SWIZZLE_128B},
    {"2D 128Bx128 rows, pitch 4096", 64, 4096, 128, CU_TENSOR_MAP_SWIZZLE_128B},
    {"2D 128Bx64 rows, pitch 512", 64, 512, 64, CU_TENSOR_MAP_SWIZZLE_128B},
    {"2D 128Bx32 rows, pitch 512", 64, 512, 32, CU_TENSOR_MAP_SWIZZLE_128B},
    {"2D 64Bx128 rows, pitch 64", 32, 64, 128, CU_TENSOR_MAP_SWIZZLE_64B},
    {"2D 64Bx128 rows, pitch 16 (overlapping window)", 32, 16, 128, CU_TENSOR_MAP_SWIZZLE_64B},
    {"2D 32Bx128 rows, pitch 32", 16, 32, 128, CU_TENSOR_MAP_SWIZZLE_32B},
    {"2D 16Bx128 rows, pitch 16", 8, 16, 128, CU_TENSOR_MAP_SWIZZLE_NONE},
    {"2D 16Bx128 rows, pitch 64", 8, 64, 128, CU_TENSOR_MAP_SWIZZLE_NONE},
  };
  for (auto& t : d2) {
    Cfg c; memset(&c, 0, sizeof(c));
    const int rows_box = t.rows > 128 ? 128 : t.rows;
    const unsigned long long total_rows = (bytes - 65536) / t.pitch_bytes;
    cuuint64_t gdim[2] = {(cuuint64_t)t.inner_elems, total_rows};
    cuuint64_t gstr[1] = {(cuuint64_t)t.pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)t.inner_elems, (cuuint32_t)rows_box};
    CUresult r = enc(&c.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, t.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", t.name, (int)r); continue; }
    c.rank = 2; c.box_bytes = rows_box * t.inner_elems * 2; c.iters = iters;
    c.c1_step = rows_box; c.c1_mod = 16;                       // each CTA cycles over 16 boxes (L2/L1-resident footprint)
    c.cta_c1_off = (int)(total_rows / 148 / rows_box) * rows_box;
    if (c.cta_c1_off < 16 * rows_box) { c.c1_mod = c.cta_c1_off / rows_box; if (c.c1_mod < 1) c.c1_mod = 1; }
    c.slots = 12; c.issuers = 1; c.c0_chunks = 1;
    run(t.name, c, d_out, rows_box);
    if (&t == &d2[0] || &t == &d2[1] || &t == &d2[8]) {
      Cfg v = c;
      v.slots = 6; run(t.name, v, d_out, rows_box);
      v.slots = 3; run(t.name, v, d_out, rows_box);
      v.slots = 3; v.issuers = 4; run(t.name, v, d_out, rows_box);
      v.slots = 6; v.issuers = 2; run(t.name, v, d_out, rows_box);
      v = c; v.store = 1; run(t.name, v, d_out, rows_box);
      v.issuers = 2; v.slots = 6; run(t.name, v, d_out, rows_box);
    }
    if (&t == &d2[1]) {   // same box through a rank-5 map with trivial outer dims
      Cfg v = c; v.triv5 = 1;
      cuuint64_t g5[5] = {(cuuint64_t)t.inner_elems, total_rows, 1, 1, 1};
      cuuint64_t s5[4] = {(cuuint64_t)t.pitch_bytes, (cuuint64_t)t.pitch_bytes * total_rows, (cuuint64_t)t.pitch_bytes * total_rows, (cuuint64_t)t.pitch_bytes * total_rows};
      cuuint32_t b5[5] = {(cuuint32_t)t.inner_elems, (cuuint32_t)rows_box, 1, 1, 1};
      CUresult r5 = enc(&v.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, buf, g5, s5, b5, es, CU_TENSOR_MAP_INTERLEAVE_NONE, t.sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r5 == CUDA_SUCCESS) { run("  ... rank-5 map, trivial outer dims", v, d_out, rows_box); v.store = 1; run("  ... rank-5 map, trivial outer dims", v, d_out, rows_box); }
      else printf("rank-5 trivial encode failed %d\n", (int)r5);
    }
  }
  {  // 5-D conv-like map: [C=256, W=14, H=14, T=8, N=64], box [64, 14, 9, 1, 1] = 126 rows of 128 B
    Cfg c; memset(&c, 0, sizeof(c));
    cuuint64_t gdim[5] = {256, 14, 14, 8, 64};
    cuuint64_t gstr[4] = {512, 512 * 14, 512 * 14 * 14, 512ull * 14 * 14 * 8};
    cuuint32_t box[5] = {64, 14, 9, 1, 1};
    CUresult r = enc(&c.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("5D encode failed %d\n", (int)r); return 1; }
    c.rank = 5; c.box_bytes = 126 * 128; c.iters = iters; c.c1_mod = 64; c.cta_c4 = 64; c.c0_chunks = 4; c.slots = 12; c.issuers = 1;
    run("5D [256,14,14,8,64] box [64,14,9,1,1]", c, d_out, 126);
    c.slots = 6; run("5D [256,14,14,8,64] box [64,14,9,1,1]", c, d_out, 126);
    c.slots = 3; c.issuers = 4; run("5D [256,14,14,8,64] box [64,14,9,1,1]", c, d_out, 126);
  }
  {  // 5-D big-spatial map: [C=64, W=56, H=56, T=8, N=8], box [64, 56, 2, 1, 1] = 112 rows
    Cfg c; memset(&c, 0, sizeof(c));
    cuuint64_t gdim[5] = {64, 56, 56, 8, 8};
    cuuint64_t gstr[4] = {128, 128 * 56, 128 * 56 * 56, 128ull * 56 * 56 * 8};
    cuuint32_t box[5] = {64, 56, 2, 1, 1};
    CUresult r = enc(&c.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("5D encode failed %d\n", (int)r); return 1; }
    c.rank = 5; c.box_bytes = 112 * 128; c.iters = iters; c.c1_mod = 64; c.cta_c4 = 8; c.c0_chunks = 1; c.slots = 12; c.issuers = 1;
    run("5D [64,56,56,8,8] box [64,56,2,1,1]", c, d_out, 112);
  }
  return 0;
}
