// Fused bottleneck block for NARROW pathways (SlowFast Fast pathway: C_inner = 8 / 16 / 32):
//
//   a = relu(bn_a(conv_a(x)))      conv_a: (kt,1,1) temporal, C_in -> C_mid            (resnet.py:1345-1365)
//   b = relu(bn_b(conv_b(a)))      conv_b: (1,3,3) spatial stride (1,s,s), C_mid -> C_mid
//   y = act(bn_c(conv_c(b)) + sc)  conv_c: 1x1x1, C_mid -> C_out; sc = x or bn_1(conv_1(x)) stride (1,s,s) (resnet.py:1179-1189)
//
// Unfused, every one of these is a launch whose tensors (a, b: 8..32 channels) round-trip through L2 and whose
// tiles are far too thin for the warp-specialised tcgen05 pipeline (measured 1 us per 128-row tile, 16 TFLOP/s:
// profiles/r01_layer_gaps.md) - 43 % of the SlowFast step for 10 % of its FLOPs.  Here ONE CTA walks the frames of
// one spatial tile of one clip: the x frames t-1, t, t+1 live in a shared-memory ring (each frame tile is read from
// L2 exactly once, zero-filled outside the clip = the padding of conv_a), a and b never leave shared memory, and
// the residual comes from the x tile that is already there.  The three GEMMs are tiny (K <= 384, N <= 128) and run
// on mma.sync m16n8k16 with ldmatrix-fed A fragments gathered straight from the position-major tiles (the im2col
// of conv_b is just a different row address per lane); the block is bound by its x read + y write.
//
// Layouts: activations NDHWC f16 (row stride >= C); weights packed [n][k] f16 with k = (tap, ci), K padded to 16;
// folded BatchNorm as fp32 (scale, bias) per output channel, applied on the fp32 accumulator like everywhere else.
#include "pv_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace pv {

struct FbParams {
  int N, T, H, W, Ho, Wo;
  int Cin, Cout;
  int kt, sb, has_sc, act;
  int TH, TW, TC;            // output tile (rows, cols) and frames per CTA
  int RH, RW;                // halo tile of x / a in input resolution: (TH-1)*sb+3, (TW-1)*sb+3
  int tiles_h, tiles_w, tchunks;
  int KA, KB, KC, KS;        // padded K extents (multiples of 16)
  int ldwa, ldwb, ldwc, ldws;  // shared-memory row pitches of the weight matrices (elements, K + 8: conflict-free B loads)
  long long xrs, yrs;        // row strides (elements)
  // shared memory carve-up (byte offsets)
  unsigned off_x, x_slot_bytes, off_a, off_b, off_y, off_wa, off_wb, off_wc, off_ws, off_sb;
};

__device__ __forceinline__ unsigned fb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// 16-byte chunk swizzle inside a position row of `row_bytes` (power of two >= 16): keeps the 8 row addresses of an
// ldmatrix phase on distinct banks for 32 / 64 / >= 128-byte rows (16-byte rows are contiguous already).
__device__ __forceinline__ unsigned fb_swz(unsigned row, unsigned chunk, unsigned row_bytes) {
  if (row_bytes >= 128u) return chunk ^ (row & 7u);
  if (row_bytes == 64u) return chunk ^ ((row >> 1) & 3u);
  if (row_bytes == 32u) return chunk ^ ((row >> 2) & 1u);
  return chunk;
}

__device__ __forceinline__ void ldmatrix_x4(unsigned addr, unsigned (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16u : 0u) : "memory");
}

constexpr int FB_WARPS = 8;
constexpr int FB_THREADS = FB_WARPS * 32;
constexpr int FB_SLOTS = 4;      // frame ring: t-1, t, t+1 in use, t+2 in flight

// One GEMM phase for one m-tile of 16 rows: acc[nt] += A(16 x K) * W^T, n-tiles nt0 .. nt0+NT-1.
// Lane l feeds ldmatrix with the address of A row fb_lane_row(l) = (l & 7) + 8 * ((l >> 3) & 1), K-half (l >> 4):
// chunk_addr(kk) = shared-memory BYTE address (u32) of the 16-byte chunk A[that row][kk .. kk+7], kk a multiple of 8.
// (The row is fixed per lane, so callers hoist everything row-dependent out of the K loop.)
__device__ __forceinline__ int fb_lane_row(int lane) { return (lane & 7) + ((lane >> 3) & 1) * 8; }

template <int NT, typename ChunkAddr>
__device__ __forceinline__ void fb_gemm(float (&acc)[NT][4], int K, const __half* __restrict__ wsm, int ldw, int n0,
                                        int lane, ChunkAddr chunk_addr) {
  const int kh = (lane >> 4) * 8;
  const int g = lane >> 2, q = lane & 3;
  for (int k0 = 0; k0 < K; k0 += 16) {
    unsigned a[4];
    ldmatrix_x4(chunk_addr(k0 + kh), a);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const __half* wrow = wsm + (size_t)(n0 + nt * 8 + g) * ldw + k0 + 2 * q;
      const unsigned b0 = *reinterpret_cast<const unsigned*>(wrow);
      const unsigned b1 = *reinterpret_cast<const unsigned*>(wrow + 8);
      mma_16816(acc[nt], a, b0, b1);
    }
  }
}

template <int CMID>
__global__ void __launch_bounds__(FB_THREADS)
bottleneck_fused_kernel(const __grid_constant__ FbParams P, const __half* __restrict__ x, const __half* __restrict__ wa,
                        const __half* __restrict__ wb, const __half* __restrict__ wc, const __half* __restrict__ wsc,
                        const float* __restrict__ sa, const float* __restrict__ ba, const float* __restrict__ sb_,
                        const float* __restrict__ bb, const float* __restrict__ sc, const float* __restrict__ bc,
                        const float* __restrict__ ssc, const float* __restrict__ bsc, __half* __restrict__ y) {
  constexpr int NTM = CMID / 8;                  // n-tiles of the inner width
  extern __shared__ __align__(128) unsigned char fb_smem[];
  const unsigned sbase = fb_smem_u32(fb_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;

  // ---- which tile
  int bid = blockIdx.x;
  const int tw = bid % P.tiles_w; bid /= P.tiles_w;
  const int th = bid % P.tiles_h; bid /= P.tiles_h;
  const int tch = bid % P.tchunks;
  const int n = bid / P.tchunks;
  const int oy0 = th * P.TH, ox0 = tw * P.TW;                 // output tile origin
  const int iy0 = oy0 * P.sb - 1, ix0 = ox0 * P.sb - 1;       // halo tile origin in the input (may be -1)
  const int t_begin = tch * P.TC, t_end = min(P.T, t_begin + P.TC);
  const int npos_in = P.RH * P.RW, npos_out = P.TH * P.TW;
  const int pad_t = P.kt >> 1;
  const unsigned xrow_bytes = (unsigned)P.Cin * 2u, arow_bytes = (unsigned)CMID * 2u;
  const unsigned xchunks = xrow_bytes >> 4;
  const int cin_log2 = 31 - __clz(P.Cin);

  // ---- weights + folded BN into shared memory (once per CTA)
  {
    __half* s_wa = reinterpret_cast<__half*>(fb_smem + P.off_wa);
    __half* s_wb = reinterpret_cast<__half*>(fb_smem + P.off_wb);
    __half* s_wc = reinterpret_cast<__half*>(fb_smem + P.off_wc);
    __half* s_ws = reinterpret_cast<__half*>(fb_smem + P.off_ws);
    for (int i = tid; i < CMID * (P.KA >> 3); i += FB_THREADS) {
      const int r = i / (P.KA >> 3), c = (i - r * (P.KA >> 3)) * 8;
      *reinterpret_cast<uint4*>(s_wa + r * P.ldwa + c) = *reinterpret_cast<const uint4*>(wa + (size_t)r * P.KA + c);
    }
    for (int i = tid; i < CMID * (P.KB >> 3); i += FB_THREADS) {
      const int r = i / (P.KB >> 3), c = (i - r * (P.KB >> 3)) * 8;
      *reinterpret_cast<uint4*>(s_wb + r * P.ldwb + c) = *reinterpret_cast<const uint4*>(wb + (size_t)r * P.KB + c);
    }
    for (int i = tid; i < P.Cout * (P.KC >> 3); i += FB_THREADS) {
      const int r = i / (P.KC >> 3), c = (i - r * (P.KC >> 3)) * 8;
      *reinterpret_cast<uint4*>(s_wc + r * P.ldwc + c) = *reinterpret_cast<const uint4*>(wc + (size_t)r * P.KC + c);
    }
    if (P.has_sc)
      for (int i = tid; i < P.Cout * (P.KS >> 3); i += FB_THREADS) {
        const int r = i / (P.KS >> 3), c = (i - r * (P.KS >> 3)) * 8;
        *reinterpret_cast<uint4*>(s_ws + r * P.ldws + c) = *reinterpret_cast<const uint4*>(wsc + (size_t)r * P.KS + c);
      }
    float* s_sb = reinterpret_cast<float*>(fb_smem + P.off_sb);     // [sa ba sb bb](CMID each) [sc bc ssc bsc](Cout each)
    for (int i = tid; i < CMID; i += FB_THREADS) {
      s_sb[i] = sa[i]; s_sb[CMID + i] = ba[i]; s_sb[2 * CMID + i] = sb_[i]; s_sb[3 * CMID + i] = bb[i];
    }
    for (int i = tid; i < P.Cout; i += FB_THREADS) {
      float* o = s_sb + 4 * CMID;
      o[i] = sc[i]; o[P.Cout + i] = bc[i];
      o[2 * P.Cout + i] = P.has_sc ? ssc[i] : 0.f; o[3 * P.Cout + i] = P.has_sc ? bsc[i] : 0.f;
    }
  }
  const __half* s_wa = reinterpret_cast<const __half*>(fb_smem + P.off_wa);
  const __half* s_wb = reinterpret_cast<const __half*>(fb_smem + P.off_wb);
  const __half* s_wc = reinterpret_cast<const __half*>(fb_smem + P.off_wc);
  const __half* s_ws = reinterpret_cast<const __half*>(fb_smem + P.off_ws);
  const float* s_sb = reinterpret_cast<const float*>(fb_smem + P.off_sb);

  // ---- frame loader: halo tile of frame f -> ring slot (f + pad_t + 4) & 3, zero-filled outside the clip
  const __half* xn = x + (size_t)n * P.T * P.H * P.W * P.xrs;
  auto load_frame = [&](int f) {
    const unsigned slot = (unsigned)(f + 8) & (FB_SLOTS - 1);
    const unsigned dst0 = sbase + P.off_x + slot * P.x_slot_bytes;
    const bool f_ok = f >= 0 && f < P.T;
    const int total = npos_in * (int)xchunks;
    for (int i = tid; i < total; i += FB_THREADS) {
      const int pos = i / (int)xchunks, ch = i - pos * (int)xchunks;
      const int py = pos / P.RW, px = pos - py * P.RW;
      const int iy = iy0 + py, ix = ix0 + px;
      const bool ok = f_ok && (unsigned)iy < (unsigned)P.H && (unsigned)ix < (unsigned)P.W;
      const __half* src = ok ? xn + ((size_t)((size_t)f * P.H + iy) * P.W + ix) * P.xrs + ch * 8 : x;
      cp_async16(dst0 + (unsigned)pos * xrow_bytes + (fb_swz((unsigned)pos, (unsigned)ch, xrow_bytes) << 4), src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // prologue: frames t_begin - pad_t .. t_begin + pad_t (pad_t = 0: just t_begin), then one frame ahead
  for (int f = t_begin - pad_t; f <= t_begin + pad_t; ++f) load_frame(f);

  const int mt_in = (npos_in + 15) >> 4, mt_out = (npos_out + 15) >> 4;
  __half* s_a = reinterpret_cast<__half*>(fb_smem + P.off_a);
  __half* s_b = reinterpret_cast<__half*>(fb_smem + P.off_b);
  __half* s_y = reinterpret_cast<__half*>(fb_smem + P.off_y);
  const unsigned a_base = sbase + P.off_a, b_base = sbase + P.off_b;
  const unsigned yrow_bytes = (unsigned)P.Cout * 2u;

  for (int t = t_begin; t < t_end; ++t) {
    // frame t + pad_t + 1 goes into the slot that frame t - pad_t - 1 ... no longer needs (4 slots >= kt + 1)
    load_frame(t + pad_t + 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");      // everything but the frame just requested has landed
    __syncthreads();                                           // (also: weights visible on the first step; s_y drained)

    // ================= phase A: a = relu(bn_a(conv_a(x))) on the halo tile, 0 outside the image =================
    for (int mt = warp; mt < mt_in; mt += FB_WARPS) {
      float acc[NTM][4];
#pragma unroll
      for (int i = 0; i < NTM; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
      const int row0 = mt * 16;
      {
        const unsigned pos = (unsigned)min(row0 + fb_lane_row(lane), npos_in - 1);
        const unsigned rowoff = sbase + P.off_x + pos * xrow_bytes;
        fb_gemm<NTM>(acc, P.KA, s_wa, P.ldwa, 0, lane, [&](int kk) -> unsigned {
          int dt = kk >> cin_log2;                                 // k = dt * Cin + ci (Cin is a power of two)
          const int ci = kk & (P.Cin - 1);
          dt = min(dt, P.kt - 1);                                  // K padding: zero weights, any finite data
          const unsigned slot = (unsigned)(t + dt - pad_t + 8) & (FB_SLOTS - 1);
          return rowoff + slot * P.x_slot_bytes + (fb_swz(pos, (unsigned)(ci >> 3), xrow_bytes) << 4);
        });
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pos = row0 + g + 8 * h;
        if (pos < npos_in) {
          const int py = pos / P.RW, px = pos - py * P.RW;
          const bool inside = (unsigned)(iy0 + py) < (unsigned)P.H && (unsigned)(ix0 + px) < (unsigned)P.W;
#pragma unroll
          for (int nt = 0; nt < NTM; ++nt) {
            const int c = nt * 8 + 2 * q;
            float v0 = fmaf(acc[nt][2 * h], s_sb[c], s_sb[CMID + c]);
            float v1 = fmaf(acc[nt][2 * h + 1], s_sb[c + 1], s_sb[CMID + c + 1]);
            v0 = inside ? fmaxf(v0, 0.f) : 0.f;
            v1 = inside ? fmaxf(v1, 0.f) : 0.f;
            const unsigned off = (unsigned)pos * arow_bytes + (fb_swz((unsigned)pos, (unsigned)(c >> 3), arow_bytes) << 4) + (unsigned)(c & 7) * 2u;
            *reinterpret_cast<__half2*>(reinterpret_cast<unsigned char*>(s_a) + off) = __floats2half2_rn(v0, v1);
          }
        }
      }
    }
    __syncthreads();

    // ================= phase B: b = relu(bn_b(conv_b(a))), 3x3 window of the a tile ============================
    for (int mt = warp; mt < mt_out; mt += FB_WARPS) {
      float acc[NTM][4];
#pragma unroll
      for (int i = 0; i < NTM; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
      const int row0 = mt * 16;
      {
        const int opos = min(row0 + fb_lane_row(lane), npos_out - 1);
        const int qy = opos / P.TW, qx = opos - qy * P.TW;
        const unsigned pos0 = (unsigned)(qy * P.sb * P.RW + qx * P.sb);       // window corner in the a tile
        fb_gemm<NTM>(acc, P.KB, s_wb, P.ldwb, 0, lane, [&](int kk) -> unsigned {
          int tap = kk / CMID;                                     // k = (dh * 3 + dw) * CMID + ci, CMID a power of two
          const int ci = kk & (CMID - 1);
          tap = min(tap, 8);                                       // K padding: zero weights
          const int dh = (tap * 11) >> 5, dw = tap - dh * 3;       // tap / 3 for tap in 0..8
          const unsigned pos = pos0 + (unsigned)(dh * P.RW + dw);
          return a_base + pos * arow_bytes + (fb_swz(pos, (unsigned)(ci >> 3), arow_bytes) << 4);
        });
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int opos = row0 + g + 8 * h;
        if (opos < npos_out) {
#pragma unroll
          for (int nt = 0; nt < NTM; ++nt) {
            const int c = nt * 8 + 2 * q;
            const float v0 = fmaxf(fmaf(acc[nt][2 * h], s_sb[2 * CMID + c], s_sb[3 * CMID + c]), 0.f);
            const float v1 = fmaxf(fmaf(acc[nt][2 * h + 1], s_sb[2 * CMID + c + 1], s_sb[3 * CMID + c + 1]), 0.f);
            const unsigned off = (unsigned)opos * arow_bytes + (fb_swz((unsigned)opos, (unsigned)(c >> 3), arow_bytes) << 4) + (unsigned)(c & 7) * 2u;
            *reinterpret_cast<__half2*>(reinterpret_cast<unsigned char*>(s_b) + off) = __floats2half2_rn(v0, v1);
          }
        }
      }
    }
    __syncthreads();

    // ================= phase C: y = act(bn_c(conv_c(b)) + shortcut), 32 output channels per pass ================
    const unsigned xslot_t = sbase + P.off_x + ((unsigned)(t + 8) & (FB_SLOTS - 1)) * P.x_slot_bytes;
    const float* s_c = s_sb + 4 * CMID;
    const int ngroups = P.Cout >> 5;
    for (int item = warp; item < mt_out * ngroups; item += FB_WARPS) {
      const int mt = item / ngroups, ng = item - mt * ngroups;
      const int row0 = mt * 16, n0 = ng * 32;
      float acc[4][4], acs[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; acs[i][0] = acs[i][1] = acs[i][2] = acs[i][3] = 0.f; }
      const unsigned lpos = (unsigned)min(row0 + fb_lane_row(lane), npos_out - 1);     // this lane's A row
      fb_gemm<4>(acc, P.KC, s_wc, P.ldwc, n0, lane, [&](int kk) -> unsigned {
        const int ci = min(kk, CMID - 8);                        // K padding (CMID = 8): zero weights
        return b_base + lpos * arow_bytes + (fb_swz(lpos, (unsigned)(ci >> 3), arow_bytes) << 4);
      });
      if (P.has_sc) {
        const int lqy = (int)lpos / P.TW, lqx = (int)lpos - lqy * P.TW;
        const unsigned xpos = (unsigned)((lqy * P.sb + 1) * P.RW + lqx * P.sb + 1);     // centre of the window in the x tile
        fb_gemm<4>(acs, P.KS, s_ws, P.ldws, n0, lane, [&](int kk) -> unsigned {
          const int ci = min(kk, P.Cin - 8);
          return xslot_t + xpos * xrow_bytes + (fb_swz(xpos, (unsigned)(ci >> 3), xrow_bytes) << 4);
        });
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int opos = row0 + g + 8 * h;
        if (opos < npos_out) {
          const int qy = opos / P.TW, qx = opos - qy * P.TW;
          const unsigned pos = (unsigned)((qy * P.sb + 1) * P.RW + qx * P.sb + 1);     // centre position in the x tile
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int c = n0 + nt * 8 + 2 * q;
            float v0 = fmaf(acc[nt][2 * h], s_c[c], s_c[P.Cout + c]);
            float v1 = fmaf(acc[nt][2 * h + 1], s_c[c + 1], s_c[P.Cout + c + 1]);
            if (P.has_sc) {
              v0 += fmaf(acs[nt][2 * h], s_c[2 * P.Cout + c], s_c[3 * P.Cout + c]);
              v1 += fmaf(acs[nt][2 * h + 1], s_c[2 * P.Cout + c + 1], s_c[3 * P.Cout + c + 1]);
            } else {
              const unsigned xo = pos * xrow_bytes + (fb_swz(pos, (unsigned)(c >> 3), xrow_bytes) << 4) + (unsigned)(c & 7) * 2u;
              __half2 rv;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(*reinterpret_cast<unsigned*>(&rv)) : "r"(xslot_t + xo));
              const float2 rf = __half22float2(rv);
              v0 += rf.x; v1 += rf.y;
            }
            if (P.act == PV_ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            *reinterpret_cast<__half2*>(reinterpret_cast<unsigned char*>(s_y) + (unsigned)opos * yrow_bytes + (unsigned)c * 2u) = __floats2half2_rn(v0, v1);
          }
        }
      }
    }
    __syncthreads();

    // ================= store the y tile: 16-byte coalesced rows ================================================
    {
      const int ychunks = (int)(yrow_bytes >> 4);
      __half* yt = y + ((size_t)n * P.T + t) * P.Ho * P.Wo * P.yrs;
      for (int i = tid; i < npos_out * ychunks; i += FB_THREADS) {
        const int opos = i / ychunks, ch = i - opos * ychunks;
        const int qy = opos / P.TW, qx = opos - qy * P.TW;
        const int oy = oy0 + qy, ox = ox0 + qx;
        if (oy < P.Ho && ox < P.Wo)
          *reinterpret_cast<uint4*>(yt + ((size_t)oy * P.Wo + ox) * P.yrs + ch * 8) =
              *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(s_y) + (unsigned)opos * yrow_bytes + (unsigned)ch * 16u);
      }
    }
    // (the next iteration's first __syncthreads orders these shared-memory reads before s_y / s_a are rewritten)
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

static int fb_pad16(int k) { return (k + 15) / 16 * 16; }

}  // namespace pv

using namespace pv;

extern "C" int pv_bottleneck_fused_supported(const pv_bottleneck_desc* d) {
  if (!d) return 0;
  if (!(d->Cmid == 8 || d->Cmid == 16 || d->Cmid == 32)) return 0;
  if (d->Cin % 8 || d->Cin < 8 || d->Cin > 256 || (d->Cin & (d->Cin - 1))) return 0;   // power-of-two row bytes (swizzle)
  if (d->Cout % 32 || d->Cout > 256) return 0;
  if (!(d->kt == 1 || d->kt == 3)) return 0;
  if (!(d->sb == 1 || d->sb == 2)) return 0;
  if (!d->has_shortcut && (d->Cin != d->Cout || d->sb != 1)) return 0;
  if (!(d->act == PV_ACT_RELU || d->act == PV_ACT_NONE)) return 0;
  if (d->x_row_stride % 8 || d->y_row_stride % 8 || d->x_row_stride < d->Cin || d->y_row_stride < d->Cout) return 0;
  return 1;
}

extern "C" int pv_bottleneck_fused_fwd(const pv_bottleneck_desc* d, const void* x, const void* wa, const void* wb,
                                       const void* wc, const void* wsc, const float* sa, const float* ba,
                                       const float* sb_, const float* bb, const float* sc, const float* bc,
                                       const float* ssc, const float* bsc, void* y, void* stream) {
  PV_CHECK_ARG(d && x && wa && wb && wc && sa && ba && sb_ && bb && sc && bc && y, "null argument");
  if (!pv_bottleneck_fused_supported(d)) { set_error("fused bottleneck: unsupported configuration"); return PV_ERR_UNSUPPORTED; }
  PV_CHECK_ARG(!d->has_shortcut || (wsc && ssc && bsc), "shortcut weights missing");
  FbParams P;
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.T = d->T; P.H = d->H; P.W = d->W;
  P.Ho = (d->H + 2 - 3) / d->sb + 1; P.Wo = (d->W + 2 - 3) / d->sb + 1;
  P.Cin = d->Cin; P.Cout = d->Cout; P.kt = d->kt; P.sb = d->sb; P.has_sc = d->has_shortcut; P.act = d->act;
  P.xrs = d->x_row_stride; P.yrs = d->y_row_stride;
  P.KA = fb_pad16(d->kt * d->Cin); P.KB = fb_pad16(9 * d->Cmid); P.KC = fb_pad16(d->Cmid); P.KS = fb_pad16(d->Cin);
  P.ldwa = P.KA + 8; P.ldwb = P.KB + 8; P.ldwc = P.KC + 8; P.ldws = P.KS + 8;
  const int sm_count = current_sm_count();
  if (sm_count <= 0) { set_error("cannot query the SM count"); return PV_ERR_CUDA; }
  // ---- tile search: the largest output tile whose shared memory fits, then enough T chunks to fill the GPU
  const size_t w_bytes = ((size_t)d->Cmid * P.ldwa + (size_t)d->Cmid * P.ldwb + (size_t)d->Cout * P.ldwc +
                          (d->has_shortcut ? (size_t)d->Cout * P.ldws : 0)) * 2;
  const size_t sb_bytes = (size_t)(4 * d->Cmid + 4 * d->Cout) * 4;
  const size_t budget = 200 * 1024;
  int best_th = 0, best_tw = 0;
  double best_score = -1;
  size_t best_smem = 0;
  for (int th = 2; th <= 16; ++th)
    for (int tw = 2; tw <= 16; ++tw) {
      if (th > P.Ho + 1 || tw > P.Wo + 1) continue;
      const int rh = (th - 1) * d->sb + 3, rw = (tw - 1) * d->sb + 3;
      const size_t xs = (size_t)rh * rw * d->Cin * 2;
      const size_t need = FB_SLOTS * ((xs + 127) & ~(size_t)127) + (((size_t)rh * rw * d->Cmid * 2 + 127) & ~(size_t)127) +
                          (((size_t)th * tw * d->Cmid * 2 + 127) & ~(size_t)127) + (((size_t)th * tw * d->Cout * 2 + 127) & ~(size_t)127) +
                          ((w_bytes + 127) & ~(size_t)127) + sb_bytes + 256;
      if (need > budget) continue;
      const double cover = (double)P.Ho * P.Wo / ((double)cdiv(P.Ho, th) * th * cdiv(P.Wo, tw) * tw);   // edge waste
      const double reuse = (double)(th * tw) * d->sb * d->sb / (double)(rh * rw);                        // halo overhead
      const double mtile = (double)(th * tw) / (double)(cdiv(th * tw, 16) * 16) * (double)(rh * rw) / (double)(cdiv(rh * rw, 16) * 16);
      const double score = cover * reuse * mtile;
      if (score > best_score) { best_score = score; best_th = th; best_tw = tw; best_smem = need; }
    }
  if (best_score < 0) { set_error("fused bottleneck: no tile fits in shared memory"); return PV_ERR_UNSUPPORTED; }
  P.TH = best_th; P.TW = best_tw;
  P.RH = (P.TH - 1) * d->sb + 3; P.RW = (P.TW - 1) * d->sb + 3;
  P.tiles_h = (int)cdiv(P.Ho, P.TH); P.tiles_w = (int)cdiv(P.Wo, P.TW);
  {
    const long long spatial = (long long)d->N * P.tiles_h * P.tiles_w;
    int tchunks = 1;
    while (spatial * tchunks < 2 * sm_count && d->T / (tchunks + 1) >= 4) ++tchunks;      // >= 4 frames per CTA: halo frames stay cheap
    P.TC = (int)cdiv(d->T, tchunks);
    P.tchunks = (int)cdiv(d->T, P.TC);
  }
  {
    unsigned off = 0;
    auto take = [&](size_t bytes) { const unsigned o = off; off += (unsigned)((bytes + 127) & ~(size_t)127); return o; };
    P.x_slot_bytes = (unsigned)(((size_t)P.RH * P.RW * d->Cin * 2 + 127) & ~(size_t)127);
    P.off_x = take((size_t)FB_SLOTS * P.x_slot_bytes);
    P.off_a = take((size_t)P.RH * P.RW * d->Cmid * 2);
    P.off_b = take((size_t)P.TH * P.TW * d->Cmid * 2);
    P.off_y = take((size_t)P.TH * P.TW * d->Cout * 2);
    P.off_wa = take((size_t)d->Cmid * P.ldwa * 2);
    P.off_wb = take((size_t)d->Cmid * P.ldwb * 2);
    P.off_wc = take((size_t)d->Cout * P.ldwc * 2);
    P.off_ws = take(d->has_shortcut ? (size_t)d->Cout * P.ldws * 2 : 16);
    P.off_sb = take(sb_bytes);
    best_smem = off;
  }
  const long long grid = (long long)d->N * P.tchunks * P.tiles_h * P.tiles_w;
  if (grid <= 0) return PV_OK;
  PV_CHECK_ARG(grid < (1ll << 31), "grid too large");
  cudaStream_t s = (cudaStream_t)stream;
#define PV_FB(CM)                                                                                                   \
  do {                                                                                                              \
    PV_OPT_IN_SMEM(bottleneck_fused_kernel<CM>, 208 * 1024);                                                        \
    bottleneck_fused_kernel<CM><<<(unsigned)grid, FB_THREADS, best_smem, s>>>(                                      \
        P, (const __half*)x, (const __half*)wa, (const __half*)wb, (const __half*)wc, (const __half*)wsc, sa, ba,   \
        sb_, bb, sc, bc, ssc, bsc, (__half*)y);                                                                     \
  } while (0)
  if (d->Cmid == 8) PV_FB(8);
  else if (d->Cmid == 16) PV_FB(16);
  else PV_FB(32);
#undef PV_FB
  PV_LAUNCH_OK("bottleneck_fused_kernel");
  return PV_OK;
}
