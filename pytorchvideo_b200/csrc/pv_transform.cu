// Fused clip transform: temporal gather + /255 + normalize + bilinear resize + crop in ONE pass.
// HBM-bound byte work: every output pixel reads exactly its 2x2 uint8 taps and writes one
// f16/f32 value; nothing full-resolution is ever materialised (the reference materialises two
// fp32 full-resolution copies, transforms/functional.py:604-615 and transforms.py:187-195).
#include "pv_common.cuh"

namespace pv {

// One thread produces PX horizontally adjacent outputs of one (c, frame, row).
// grid = (ceil(out_w / (PX*128)), out_h, C * n_t);   block = 128
template <typename T> __device__ __forceinline__ float ld_src(const T* p);
template <> __device__ __forceinline__ float ld_src<uint8_t>(const uint8_t* p) { return (float)__ldg(p); }
template <> __device__ __forceinline__ float ld_src<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_src<__half>(const __half* p) { return __half2float(__ldg(p)); }

template <typename SrcT, typename OutT, int PX>
__global__ void __launch_bounds__(128)
clip_transform_kernel(pv_clip_transform_desc d, const SrcT* __restrict__ src,
                      const int32_t* __restrict__ idx_t, const int32_t* __restrict__ y0t,
                      const int32_t* __restrict__ y1t, const float* __restrict__ lyt,
                      const int32_t* __restrict__ x0t, const int32_t* __restrict__ x1t,
                      const float* __restrict__ lxt, OutT* __restrict__ dst) {
  const int y = blockIdx.y;
  const int ct = blockIdx.z;
  const int c = ct / d.n_t, j = ct - c * d.n_t;
  const int xb = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
  if (xb >= d.out_w) return;

  const long long frame = (long long)c * d.sc + (long long)__ldg(idx_t + j) * d.st;
  const SrcT* r0 = src + frame + (long long)__ldg(y0t + y) * d.sh;
  const SrcT* r1 = src + frame + (long long)__ldg(y1t + y) * d.sh;
  const float ly1 = __ldg(lyt + y), ly0 = 1.f - ly1;
  const float mean = d.mean[c], stdv = d.stdv[c];

  float out[PX];
#pragma unroll
  for (int i = 0; i < PX; ++i) {
    const int x = min(xb + i, d.out_w - 1);
    const long long xa = (long long)__ldg(x0t + x) * d.sw, xc = (long long)__ldg(x1t + x) * d.sw;
    const float lx1 = __ldg(lxt + x), lx0 = 1.f - lx1;
    float v00 = ld_src<SrcT>(r0 + xa), v01 = ld_src<SrcT>(r0 + xc);
    float v10 = ld_src<SrcT>(r1 + xa), v11 = ld_src<SrcT>(r1 + xc);
    if (d.div255) {   // same op order as the reference: x/255.0 then (x-mean)/std, all fp32
      v00 = v00 / 255.0f; v01 = v01 / 255.0f; v10 = v10 / 255.0f; v11 = v11 / 255.0f;
    }
    v00 = (v00 - mean) / stdv; v01 = (v01 - mean) / stdv;
    v10 = (v10 - mean) / stdv; v11 = (v11 - mean) / stdv;
    // ATen upsample_bilinear2d: l_h0*(l_w0*v00 + l_w1*v01) + l_h1*(l_w0*v10 + l_w1*v11)
    out[i] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
  }
  OutT* o = dst + (((long long)ct * d.out_h + y) * d.out_w + xb);
  if (PX == 2 && xb + 1 < d.out_w && ((reinterpret_cast<uintptr_t>(o) & (2 * sizeof(OutT) - 1)) == 0)) {
    if constexpr (sizeof(OutT) == 2) {
      *reinterpret_cast<__half2*>(o) = __floats2half2_rn(out[0], out[PX - 1]);
    } else {
      *reinterpret_cast<float2*>(o) = make_float2(out[0], out[PX - 1]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < PX; ++i)
      if (xb + i < d.out_w) Elem<OutT>::st(o + i, out[i]);
  }
}

}  // namespace pv

extern "C" int pv_clip_transform_fwd(const pv_clip_transform_desc* d, const void* src,
                                     const int32_t* idx_t, const int32_t* y0, const int32_t* y1,
                                     const float* ly, const int32_t* x0, const int32_t* x1,
                                     const float* lx, void* dst, void* stream) {
  PV_CHECK_ARG(d && src && idx_t && y0 && y1 && ly && x0 && x1 && lx && dst, "null argument");
  PV_CHECK_ARG(d->C >= 1 && d->C <= 4, "C must be in 1..4 (got %d)", d->C);
  PV_CHECK_ARG(d->n_t >= 1 && d->out_h >= 1 && d->out_w >= 1, "empty output");
  PV_CHECK_ARG(d->out_h <= 65535 && (long long)d->C * d->n_t <= 65535, "grid too large");
  PV_CHECK_ARG(d->dst_dtype == PV_F16 || d->dst_dtype == PV_F32, "dst dtype must be f16|f32");
  cudaStream_t s = (cudaStream_t)stream;
  constexpr int PX = 2;
  dim3 grid((unsigned)pv::cdiv(d->out_w, PX * 128), d->out_h, d->C * d->n_t), block(128);
#define PV_TR(ST, OT)                                                                         \
  pv::clip_transform_kernel<ST, OT, PX><<<grid, block, 0, s>>>(*d, (const ST*)src, idx_t, y0, y1, ly, \
                                                              x0, x1, lx, (OT*)dst)
  const bool h = d->dst_dtype == PV_F16;
  switch (d->src_dtype) {
    case PV_U8: if (h) PV_TR(uint8_t, __half); else PV_TR(uint8_t, float); break;
    case PV_F32: if (h) PV_TR(float, __half); else PV_TR(float, float); break;
    case PV_F16: if (h) PV_TR(__half, __half); else PV_TR(__half, float); break;
    default: pv::set_error("src dtype %d unsupported", d->src_dtype); return PV_ERR_INVALID;
  }
#undef PV_TR
  PV_LAUNCH_OK("clip_transform_kernel");
  return PV_OK;
}
