// Output layer of the generator (modules/model.py:112): ReflectionPad2d(3) + Conv2d(64, 1, 7) + Tanh.
// One output channel makes this a GEMV-like, HBM/L2-bound layer (reads the 64-channel map once), so it does
// not go through the tcgen05 implicit-GEMM kernel.  It is evaluated as
//     P[y][xp][s] = sum_{r,c} in[y+r][xp][c] * w[c][r][s]           (warp-level m16n8k16 MMAs, N = 7 taps -> 8)
//     out[y][x]   = tanh(bias + sum_s P[y][x+s][s])                 (shift-sum through shared memory)
// A CTA produces 8 output rows x 128 columns so that every input row it loads is reused by up to 7 output rows.
#include "common.cuh"

namespace ducosy {
namespace {

constexpr int kRowsOut = 8;
constexpr int kColsOut = 128;
constexpr int kPos = 144;            // 9 warps x 16 padded columns (>= 128 + 6)
constexpr int kOutThreads = 288;

template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                 uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                                        uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// packed weight layout: [r = 7][s = 8 (s == 7 is zero)][c = 64]
// planes == 2 (split-operand mode): a second [7][8][64] block holds lo = rn(w - hi)
template <typename T>
__global__ void pack_out_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int planes) {
  pdl_prologue();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 7 * 8 * 64; i += gridDim.x * blockDim.x) {
    const int c = i & 63, s = (i >> 6) & 7, r = i >> 9;
    const float v = s < 7 ? w[c * 49 + r * 7 + s] : 0.f;
    const T hi = Cvt<T>::from_f(v);
    out[i] = hi;
    if (planes == 2) out[7 * 8 * 64 + i] = Cvt<T>::from_f(v - Cvt<T>::to_f(hi));
  }
}

__device__ __forceinline__ int reflect_px(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// kFused == false: `in` is the reflect-padded, already normalised map [B][H+6][W+6][64].
// kFused == true : `in` is the RAW conv output [B][H][W][64]; InstanceNorm apply + ReLU (scale/shift [B][64]) and the
//                  reflection padding are done on the fly while loading the A fragments (saves a full read+write pass).
// kSplit (needs kFused): split-operand mode (DUCOSY_F16X2) -- `in` holds 128 channels per pixel (hi plane, lo plane), `wp` the
//                  hi and lo weight blocks; every product is A_hi*W_hi + A_lo*W_hi + A_hi*W_lo.
template <typename T, bool kFused, bool kSplit>
__global__ void __launch_bounds__(kOutThreads)
out_conv7x7_tanh_kernel(const T* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ shift,
                        const T* __restrict__ wp, const float* __restrict__ bias, float* __restrict__ out, int B, int H,
                        int W) {
  pdl_prologue();
  __shared__ float P[kRowsOut][kPos][9];  // 9-float rows: the diagonal reads below are bank-conflict free
  const int tiles_x = W / kColsOut, tiles_y = H / kRowsOut;
  const int tx = blockIdx.x % tiles_x;
  const int ty = (blockIdx.x / tiles_x) % tiles_y;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Wp = W + 6, Hp = H + 6;

  // B fragments for all 7 filter rows stay in registers: thread (g,t) holds w[r][s=g][cb*32 + 8t .. +7]
  uint4 wr[7][2], wl[kSplit ? 7 : 1][2];
#pragma unroll
  for (int r = 0; r < 7; ++r)
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      wr[r][cb] = __ldg(reinterpret_cast<const uint4*>(wp + ((r * 8 + g) * 64 + cb * 32 + 8 * t)));
      if (kSplit) wl[r][cb] = __ldg(reinterpret_cast<const uint4*>(wp + 7 * 8 * 64 + ((r * 8 + g) * 64 + cb * 32 + 8 * t)));
    }

  float acc[kRowsOut][4];
#pragma unroll
  for (int i = 0; i < kRowsOut; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int xp0 = tx * kColsOut + 16 * warp + g;
  int xpA = min(xp0, Wp - 1), xpB = min(xp0 + 8, Wp - 1);  // columns past the row end feed unused outputs
  float sc[2][8], sh[2][8];
  if (kFused) {
    xpA = reflect_px(xpA - 3, W);
    xpB = reflect_px(xpB - 3, W);
#pragma unroll
    for (int cb = 0; cb < 2; ++cb)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[cb][j] = scale[b * 64 + cb * 32 + 8 * t + j];
        sh[cb][j] = shift[b * 64 + cb * 32 + 8 * t + j];
      }
  }
  constexpr int kCp = kSplit ? 128 : 64;   // 16-bit channels per pixel
  const T* base = kFused ? in + size_t(b) * H * W * kCp + 8 * t
                         : in + (size_t(b) * Hp + size_t(ty) * kRowsOut) * Wp * 64 + 8 * t;
  const int rowlen = kFused ? W : Wp;

#pragma unroll
  for (int i = 0; i < kRowsOut + 6; ++i) {
    const T* rowp = base + size_t(kFused ? reflect_px(ty * kRowsOut + i - 3, H) : i) * rowlen * kCp;
    uint4 A[2][2], L[kSplit ? 2 : 1][2];
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      A[cb][0] = *reinterpret_cast<const uint4*>(rowp + size_t(xpA) * kCp + cb * 32);
      A[cb][1] = *reinterpret_cast<const uint4*>(rowp + size_t(xpB) * kCp + cb * 32);
      if (kSplit) {   // relu((hi + lo)*scale + shift) in fp32, split again
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 lo4 = *reinterpret_cast<const uint4*>(rowp + size_t(h == 0 ? xpA : xpB) * kCp + 64 + cb * 32);
          uint32_t w4[4] = {A[cb][h].x, A[cb][h].y, A[cb][h].z, A[cb][h].w};
          uint32_t l4[4] = {lo4.x, lo4.y, lo4.z, lo4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = Cvt<T>::unpack2(w4[j]);
            const float2 l = Cvt<T>::unpack2(l4[j]);
            f.x = fmaxf(fmaf(f.x + l.x, sc[cb][2 * j], sh[cb][2 * j]), 0.f);
            f.y = fmaxf(fmaf(f.y + l.y, sc[cb][2 * j + 1], sh[cb][2 * j + 1]), 0.f);
            w4[j] = Cvt<T>::pack2(f.x, f.y);
            const float2 hv = Cvt<T>::unpack2(w4[j]);
            l4[j] = Cvt<T>::pack2(f.x - hv.x, f.y - hv.y);
          }
          A[cb][h] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          L[cb][h] = make_uint4(l4[0], l4[1], l4[2], l4[3]);
        }
      } else if (kFused) {  // relu(y*scale + shift) in fp32, rounded back to T exactly like the stand-alone apply kernel
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t w4[4] = {A[cb][h].x, A[cb][h].y, A[cb][h].z, A[cb][h].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = Cvt<T>::unpack2(w4[j]);
            f.x = fmaxf(fmaf(f.x, sc[cb][2 * j], sh[cb][2 * j]), 0.f);
            f.y = fmaxf(fmaf(f.y, sc[cb][2 * j + 1], sh[cb][2 * j + 1]), 0.f);
            w4[j] = Cvt<T>::pack2(f.x, f.y);
          }
          A[cb][h] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
#pragma unroll
    for (int yo = 0; yo < kRowsOut; ++yo) {
      const int r = i - yo;
      if (r < 0 || r > 6) continue;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        mma16816<T>(acc[yo], A[cb][0].x, A[cb][1].x, A[cb][0].y, A[cb][1].y, wr[r][cb].x, wr[r][cb].y);
        mma16816<T>(acc[yo], A[cb][0].z, A[cb][1].z, A[cb][0].w, A[cb][1].w, wr[r][cb].z, wr[r][cb].w);
        if (kSplit) {
          mma16816<T>(acc[yo], L[cb][0].x, L[cb][1].x, L[cb][0].y, L[cb][1].y, wr[r][cb].x, wr[r][cb].y);
          mma16816<T>(acc[yo], L[cb][0].z, L[cb][1].z, L[cb][0].w, L[cb][1].w, wr[r][cb].z, wr[r][cb].w);
          mma16816<T>(acc[yo], A[cb][0].x, A[cb][1].x, A[cb][0].y, A[cb][1].y, wl[r][cb].x, wl[r][cb].y);
          mma16816<T>(acc[yo], A[cb][0].z, A[cb][1].z, A[cb][0].w, A[cb][1].w, wl[r][cb].z, wl[r][cb].w);
        }
      }
    }
  }
#pragma unroll
  for (int yo = 0; yo < kRowsOut; ++yo) {
    P[yo][16 * warp + g][2 * t] = acc[yo][0];
    P[yo][16 * warp + g][2 * t + 1] = acc[yo][1];
    P[yo][16 * warp + g + 8][2 * t] = acc[yo][2];
    P[yo][16 * warp + g + 8][2 * t + 1] = acc[yo][3];
  }
  __syncthreads();
  const float bv = __ldg(bias);
  for (int idx = threadIdx.x; idx < kRowsOut * kColsOut; idx += kOutThreads) {
    const int yo = idx / kColsOut, x = idx % kColsOut;
    float v = bv;
#pragma unroll
    for (int s = 0; s < 7; ++s) v += P[yo][x + s][s];
    out[(size_t(b) * H + ty * kRowsOut + yo) * W + tx * kColsOut + x] = tanhf(v);
  }
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" int ducosy_pack_out_weight(const float* w, void* packed, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(w && packed, DUCOSY_ERR_ARG, "pack_out_weight: null pointer");
  DUCOSY_DISPATCH_DTYPE(dtype, T,
                        (pdl(pack_out_weight_kernel<T>, 14, 256, 0, (cudaStream_t)stream)(w, static_cast<T*>(packed), dtype == DUCOSY_F16X2 ? 2 : 1)));
  return check_launch("pack_out_weight_kernel");
}

extern "C" int ducosy_out_conv7x7_tanh(const void* in_pad, const void* w_packed, const float* bias, float* out, int B,
                                       int H, int W, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(in_pad && w_packed && bias && out && B > 0, DUCOSY_ERR_ARG, "out_conv7x7_tanh: bad argument");
  DUCOSY_CHECK(H % kRowsOut == 0 && W % kColsOut == 0, DUCOSY_ERR_SHAPE,
               "out_conv7x7_tanh: H must be a multiple of 8 and W of 128 (got %dx%d)", H, W);
  DUCOSY_CHECK((reinterpret_cast<uintptr_t>(in_pad) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
               DUCOSY_ERR_ALIGN, "out_conv7x7_tanh: 16-byte alignment");
  const int grid = B * (H / kRowsOut) * (W / kColsOut);
  DUCOSY_CHECK(dtype != DUCOSY_F16X2, DUCOSY_ERR_ARG, "out_conv7x7_tanh: split-operand mode is available in the fused variant only");
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(out_conv7x7_tanh_kernel<T, false, false>, grid, kOutThreads, 0, (cudaStream_t)stream)(
                                      static_cast<const T*>(in_pad), nullptr, nullptr, static_cast<const T*>(w_packed), bias,
                                      out, B, H, W)));
  return check_launch("out_conv7x7_tanh_kernel");
}

extern "C" int ducosy_out_conv7x7_tanh_fused(const void* y_raw, const float* scale, const float* shift, const void* w_packed,
                                             const float* bias, float* out, int B, int H, int W, int dtype,
                                             ducosy_stream_t stream) {
  DUCOSY_CHECK(y_raw && scale && shift && w_packed && bias && out && B > 0, DUCOSY_ERR_ARG, "out_conv7x7_tanh_fused: bad argument");
  DUCOSY_CHECK(H % kRowsOut == 0 && W % kColsOut == 0 && H >= 4, DUCOSY_ERR_SHAPE,
               "out_conv7x7_tanh_fused: H must be a multiple of 8 and W of 128 (got %dx%d)", H, W);
  DUCOSY_CHECK((reinterpret_cast<uintptr_t>(y_raw) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
               DUCOSY_ERR_ALIGN, "out_conv7x7_tanh_fused: 16-byte alignment");
  const int grid = B * (H / kRowsOut) * (W / kColsOut);
  if (dtype == DUCOSY_F16X2)
    pdl(out_conv7x7_tanh_kernel<__half, true, true>, grid, kOutThreads, 0, (cudaStream_t)stream)(
        static_cast<const __half*>(y_raw), scale, shift, static_cast<const __half*>(w_packed), bias, out, B, H, W);
  else
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(out_conv7x7_tanh_kernel<T, true, false>, grid, kOutThreads, 0, (cudaStream_t)stream)(
                                        static_cast<const T*>(y_raw), scale, shift, static_cast<const T*>(w_packed), bias, out,
                                        B, H, W)));
  return check_launch("out_conv7x7_tanh_kernel(fused)");
}
