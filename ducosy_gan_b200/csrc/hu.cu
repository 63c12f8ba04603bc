// HU windowing, HU threshold candidates and the de-window + dual-HU complementary composite.
// HBM-bound byte/integer work: 128-bit loads/stores, 8 pixels per thread, grid sized in multiples of
// the SM count.  All float steps use explicit round-to-nearest intrinsics (no FMA contraction) so the
// results are bit-identical to the numpy float32 arithmetic of the reference.
#include "common.cuh"

namespace ducosy {
namespace {

__device__ __forceinline__ float stored_to_hu(int16_t px, float slope, float intercept) {
  // generate.py:140-145 / preprocess.py:72-75: float32(px) * slope + intercept (two roundings)
  return __fadd_rn(__fmul_rn(float(px), slope), intercept);
}
__device__ __forceinline__ float window1(float hu, float lo, float hi, float span) {
  // preprocess.py:79-84: clip, then 2*(x-lo)/(hi-lo)-1
  const float c = fminf(fmaxf(hu, lo), hi);
  return __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(c, lo)), span), 1.0f);
}
__device__ __forceinline__ int16_t dewindow1(float y, float lo, float span, float slope, float intercept) {
  // preprocess.py:100-111: (y+1)/2*(hi-lo)+lo ; (hu-b)/m ; astype(int16) truncates toward zero
  const float hu = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(y, 1.0f), 2.0f), span), lo);
  const float v = __fdiv_rn(__fsub_rn(hu, intercept), slope);
  return static_cast<int16_t>(__float2int_rz(v));
}

struct alignas(16) Px8 { int16_t v[8]; };
struct alignas(8) U8x8 { uint8_t v[8]; };

__global__ void hu_window_kernel(const int16_t* __restrict__ px, float* __restrict__ out_soft,
                                 float* __restrict__ out_lung, long long n, float slope, float intercept, float slo,
                                 float shi, float sspan, float llo, float lhi, float lspan) {
  pdl_prologue();
  const long long groups = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const Px8 p = reinterpret_cast<const Px8*>(px)[g];
    float s[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float hu = stored_to_hu(p.v[i], slope, intercept);
      s[i] = window1(hu, slo, shi, sspan);
      l[i] = window1(hu, llo, lhi, lspan);
    }
    if (out_soft) {
      float4* d = reinterpret_cast<float4*>(out_soft + g * 8);
      d[0] = make_float4(s[0], s[1], s[2], s[3]);
      d[1] = make_float4(s[4], s[5], s[6], s[7]);
    }
    if (out_lung) {
      float4* d = reinterpret_cast<float4*>(out_lung + g * 8);
      d[0] = make_float4(l[0], l[1], l[2], l[3]);
      d[1] = make_float4(l[4], l[5], l[6], l[7]);
    }
  }
  // tail (n not a multiple of 8)
  for (long long i = (groups << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float hu = stored_to_hu(px[i], slope, intercept);
    if (out_soft) out_soft[i] = window1(hu, slo, shi, sspan);
    if (out_lung) out_lung[i] = window1(hu, llo, lhi, lspan);
  }
}

__global__ void hu_thresholds_kernel(const int16_t* __restrict__ px, uint8_t* __restrict__ body,
                                     uint8_t* __restrict__ lung, uint8_t* __restrict__ bone, long long n, float slope,
                                     float intercept) {
  pdl_prologue();
  const long long groups = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const Px8 p = reinterpret_cast<const Px8*>(px)[g];
    U8x8 b, l, o;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float hu = stored_to_hu(p.v[i], slope, intercept);
      const bool isbody = hu > -1000.0f;                                  // mask_generator.py:14,179
      b.v[i] = isbody;
      l.v[i] = (hu >= -1000.0f && hu <= -300.0f) && isbody;               // mask_generator.py:17-20
      o.v[i] = (hu >= 200.0f) && isbody;                                  // mask_generator.py:182-183
    }
    if (body) reinterpret_cast<U8x8*>(body)[g] = b;
    if (lung) reinterpret_cast<U8x8*>(lung)[g] = l;
    if (bone) reinterpret_cast<U8x8*>(bone)[g] = o;
  }
  for (long long i = (groups << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float hu = stored_to_hu(px[i], slope, intercept);
    const bool isbody = hu > -1000.0f;
    if (body) body[i] = isbody;
    if (lung) lung[i] = (hu >= -1000.0f && hu <= -300.0f) && isbody;
    if (bone) bone[i] = (hu >= 200.0f) && isbody;
  }
}

struct CompositeOne {
  int16_t merged, soft, lung;
  uint8_t mask;
};
__device__ __forceinline__ CompositeOne composite1(int16_t raw, float ys, float yl, float slope, float intercept,
                                                   float slo, float shi, float sspan, float llo, float lhi,
                                                   float lspan) {
  CompositeOne r;
  const float hu = stored_to_hu(raw, slope, intercept);
  r.soft = dewindow1(ys, slo, sspan, slope, intercept);
  r.lung = dewindow1(yl, llo, lspan, slope, intercept);
  const bool sm = (hu >= slo) && (hu <= shi);   // generate.py:224-227
  const bool lm = (hu >= llo) && (hu <= lhi);   // generate.py:230-233
  int16_t m = raw;                              // generate.py:218
  if (sm) m = r.soft;                           // generate.py:236
  if (lm) m = r.lung;                           // generate.py:237 (lung written last: wins where ranges touch)
  r.merged = m;
  r.mask = uint8_t(sm) | uint8_t(uint8_t(lm) << 1);
  return r;
}

__global__ void dewindow_composite_kernel(const int16_t* __restrict__ raw, const float* __restrict__ ys,
                                          const float* __restrict__ yl, int16_t* __restrict__ merged,
                                          int16_t* __restrict__ soft_px, int16_t* __restrict__ lung_px,
                                          uint8_t* __restrict__ masks, long long n, float slope, float intercept,
                                          float slo, float shi, float sspan, float llo, float lhi, float lspan) {
  pdl_prologue();
  const long long groups = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const Px8 p = reinterpret_cast<const Px8*>(raw)[g];
    const float4 s0 = reinterpret_cast<const float4*>(ys + g * 8)[0], s1 = reinterpret_cast<const float4*>(ys + g * 8)[1];
    const float4 l0 = reinterpret_cast<const float4*>(yl + g * 8)[0], l1 = reinterpret_cast<const float4*>(yl + g * 8)[1];
    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
    Px8 m, so, lu;
    U8x8 mk;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const CompositeOne r = composite1(p.v[i], sv[i], lv[i], slope, intercept, slo, shi, sspan, llo, lhi, lspan);
      m.v[i] = r.merged;
      so.v[i] = r.soft;
      lu.v[i] = r.lung;
      mk.v[i] = r.mask;
    }
    reinterpret_cast<Px8*>(merged)[g] = m;
    if (soft_px) reinterpret_cast<Px8*>(soft_px)[g] = so;
    if (lung_px) reinterpret_cast<Px8*>(lung_px)[g] = lu;
    if (masks) reinterpret_cast<U8x8*>(masks)[g] = mk;
  }
  for (long long i = (groups << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const CompositeOne r = composite1(raw[i], ys[i], yl[i], slope, intercept, slo, shi, sspan, llo, lhi, lspan);
    merged[i] = r.merged;
    if (soft_px) soft_px[i] = r.soft;
    if (lung_px) lung_px[i] = r.lung;
    if (masks) masks[i] = r.mask;
  }
}

// training-side windowing with soft squeezing (modules/preprocess.py:6-40,43-55; dataset.py:118-120)
__global__ void hu_window_soft_kernel(const int16_t* __restrict__ px, float* __restrict__ out, long long n, float slope,
                                      float intercept, float lo, float hi, float span, float k) {
  pdl_prologue();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float hu = stored_to_hu(px[i], slope, intercept);
    const float c = fminf(fmaxf(hu, lo), hi);
    const float nrm = __fdiv_rn(__fsub_rn(c, lo), span);                                   // preprocess.py:20
    const float sm = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fmul_rn(-k, __fsub_rn(nrm, 0.9f)))));  // preprocess.py:28
    const float r = nrm < 0.9f ? nrm : __fadd_rn(0.9f, __fmul_rn(0.1f, sm));               // preprocess.py:31-35
    out[i] = __fsub_rn(__fmul_rn(2.0f, r), 1.0f);                                          // preprocess.py:38
  }
}
// display windowing of a tanh-range tensor (modules/preprocess.py:58-65)
__global__ void apply_windowing_kernel(const float* __restrict__ y, float* __restrict__ out, long long n, float lo, float span,
                                       float wlo, float whi, float ww) {
  pdl_prologue();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float hu = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(y[i], 1.0f), 2.0f), span), lo);
    out[i] = __fdiv_rn(__fsub_rn(fminf(fmaxf(hu, wlo), whi), wlo), ww);
  }
}

int grid_for(long long work_items, int threads) {
  const int sms = num_sms();
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)(sms > 0 ? sms : 148) * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" int ducosy_hu_window(const int16_t* px, float* out_soft, float* out_lung, long long n, float slope,
                                float intercept, float soft_lo, float soft_hi, float lung_lo, float lung_hi,
                                ducosy_stream_t stream) {
  if (n == 0) return 0;
  DUCOSY_CHECK(px != nullptr && n > 0, DUCOSY_ERR_ARG, "hu_window: null input");
  DUCOSY_CHECK(aligned16(px) && aligned16(out_soft) && aligned16(out_lung), DUCOSY_ERR_ALIGN,
               "hu_window: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  const float sspan = float(double(soft_hi) - double(soft_lo)), lspan = float(double(lung_hi) - double(lung_lo));
  pdl(hu_window_kernel, grid_for((n + 7) / 8, 256), 256, 0, static_cast<cudaStream_t>(stream))(
      px, out_soft, out_lung, n, slope, intercept, soft_lo, soft_hi, sspan, lung_lo, lung_hi, lspan);
  return check_launch("hu_window_kernel");
}

extern "C" int ducosy_hu_thresholds(const int16_t* px, uint8_t* body, uint8_t* lung, uint8_t* bone, long long n,
                                    float slope, float intercept, ducosy_stream_t stream) {
  if (n == 0) return 0;
  DUCOSY_CHECK(px != nullptr && n > 0, DUCOSY_ERR_ARG, "hu_thresholds: null input");
  DUCOSY_CHECK(aligned16(px) && aligned16(body) && aligned16(lung) && aligned16(bone), DUCOSY_ERR_ALIGN,
               "hu_thresholds: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  pdl(hu_thresholds_kernel, grid_for((n + 7) / 8, 256), 256, 0, static_cast<cudaStream_t>(stream))(px, body, lung, bone,
                                                                                                   n, slope, intercept);
  return check_launch("hu_thresholds_kernel");
}

extern "C" int ducosy_dewindow_composite(const int16_t* raw_px, const float* y_soft, const float* y_lung,
                                         int16_t* merged, int16_t* soft_px, int16_t* lung_px, uint8_t* masks,
                                         long long n, float slope, float intercept, float soft_lo, float soft_hi,
                                         float lung_lo, float lung_hi, ducosy_stream_t stream) {
  if (n == 0) return 0;
  DUCOSY_CHECK(raw_px && y_soft && y_lung && merged && n > 0, DUCOSY_ERR_ARG, "dewindow_composite: null pointer");
  DUCOSY_CHECK(slope != 0.0f, DUCOSY_ERR_ARG, "dewindow_composite: RescaleSlope is 0");
  DUCOSY_CHECK(aligned16(raw_px) && aligned16(y_soft) && aligned16(y_lung) && aligned16(merged) && aligned16(soft_px) &&
                   aligned16(lung_px) && aligned16(masks),
               DUCOSY_ERR_ALIGN, "dewindow_composite: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  const float sspan = float(double(soft_hi) - double(soft_lo)), lspan = float(double(lung_hi) - double(lung_lo));
  pdl(dewindow_composite_kernel, grid_for((n + 7) / 8, 256), 256, 0, static_cast<cudaStream_t>(stream))(
      raw_px, y_soft, y_lung, merged, soft_px, lung_px, masks, n, slope, intercept, soft_lo, soft_hi, sspan, lung_lo,
      lung_hi, lspan);
  return check_launch("dewindow_composite_kernel");
}

extern "C" int ducosy_hu_window_soft(const int16_t* px, float* out, long long n, float slope, float intercept, float hu_lo,
                                     float hu_hi, float sigma, ducosy_stream_t stream) {
  if (n == 0) return 0;
  DUCOSY_CHECK(px && out && n > 0 && sigma > 0.f, DUCOSY_ERR_ARG, "hu_window_soft: bad argument");
  pdl(hu_window_soft_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream))(
      px, out, n, slope, intercept, hu_lo, hu_hi, float(double(hu_hi) - double(hu_lo)), float(10.0 / double(sigma)));
  return check_launch("hu_window_soft_kernel");
}

extern "C" int ducosy_apply_windowing(const float* y, float* out, long long n, float hu_lo, float hu_hi, float window_center,
                                      float window_width, ducosy_stream_t stream) {
  if (n == 0) return 0;
  DUCOSY_CHECK(y && out && n > 0 && window_width != 0.f, DUCOSY_ERR_ARG, "apply_windowing: bad argument");
  const double half = double(window_width) / 2.0;
  pdl(apply_windowing_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream))(
      y, out, n, hu_lo, float(double(hu_hi) - double(hu_lo)), float(double(window_center) - half),
      float(double(window_center) + half), window_width);
  return check_launch("apply_windowing_kernel");
}
