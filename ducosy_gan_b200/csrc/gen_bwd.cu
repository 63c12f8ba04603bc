// Generator backward pieces that are not tensor-core GEMMs (what autograd computes through modules/model.py:80-115 for
// modules/trainer.py:497): the 64 -> 1 output conv + tanh, the adjoint of the 7x7 stem's reflection-padded im2col, and the
// CBAM / residual-block element-wise adjoints.  Everything is bandwidth-bound CUDA-core work; reductions run in a
// fixed order (deterministic gradients).  16-bit gradient maps carry the power-of-two scale gs[0] (ducosy_grad_scale);
// fp32 parameter / image gradients leave with the true scale.
#include "common.cuh"
#include "input_fn.cuh"

namespace ducosy {
namespace {

int grid_items(long long items, int threads) { return int((items + threads - 1) / threads); }

// The padded coordinates u (0 <= u < n + 2*pad) that ReflectionPad2d(pad) fills from source index i; returns the count.
__device__ __forceinline__ int reflect_sources(int i, int n, int pad, int* u) {
  int k = 0;
  u[k++] = i + pad;
  if (i >= 1 && i <= pad) u[k++] = pad - i;
  if (i <= n - 2 && i >= n - 1 - pad) u[k++] = 2 * (n - 1) - i + pad;
  return k;
}

// ------------------------------------------------------------------ output conv backward (modules/model.py:112-113)
// out = tanh(v), v = conv7x7(reflect_pad3(a)) + bias.  dv = dout * (1 - out^2); db = sum dv.
__global__ void __launch_bounds__(256)
out_tanh_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dv,
                    float* __restrict__ partial, long long n) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float o = out[i], g = dout[i] * (1.f - o * o);
    dv[i] = g;
    acc += g;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < int(blockDim.x >> 5); ++i) s += red[i];
    out[0] = s;
  }
}

// da[b][y][x][c] = gs0 * sum over the padded positions (u,v) that mirror (y,x) of sum_{r,s} dv[u-r][v-s] * w[c][r][s]
// thread = (pixel, 8 channels); weights transposed to [tap][channel] in shared memory.
template <typename T>
__global__ void __launch_bounds__(256)
out_conv_dgrad_kernel(const float* __restrict__ dv, const float* __restrict__ w, T* __restrict__ da,
                      const float* __restrict__ gs, int B, int H, int W) {
  __shared__ __align__(16) float ws[49][64];
  for (int i = threadIdx.x; i < 49 * 64; i += 256) ws[i % 49][i / 49] = w[i];   // w is [c][tap]
  __syncthreads();
  const long long item = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * H * W * 8;
  if (item >= total) return;
  const int c8 = int(item & 7);
  const long long pix = item >> 3;
  const int x = int(pix % W), y = int((pix / W) % H), b = int(pix / ((long long)W * H));
  const float* dvb = dv + (long long)b * H * W;
  int us[3], vs[3];
  const int nu = reflect_sources(y, H, 3, us), nv = reflect_sources(x, W, 3, vs);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int iu = 0; iu < nu; ++iu)
    for (int iv = 0; iv < nv; ++iv) {
      const int u = us[iu], v = vs[iv];
#pragma unroll
      for (int r = 0; r < 7; ++r) {
        const int yo = u - r;
        if (yo < 0 || yo >= H) continue;
#pragma unroll
        for (int s = 0; s < 7; ++s) {
          const int xo = v - s;
          if (xo < 0 || xo >= W) continue;
          const float g = __ldg(dvb + (long long)yo * W + xo);
          const float4 w0 = *reinterpret_cast<const float4*>(&ws[r * 7 + s][c8 * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&ws[r * 7 + s][c8 * 8 + 4]);
          acc[0] += g * w0.x; acc[1] += g * w0.y; acc[2] += g * w0.z; acc[3] += g * w0.w;
          acc[4] += g * w1.x; acc[5] += g * w1.y; acc[6] += g * w1.z; acc[7] += g * w1.w;
        }
      }
    }
  const float sc = gs[0];
  uint4 o;
  o.x = Cvt<T>::pack2(acc[0] * sc, acc[1] * sc);
  o.y = Cvt<T>::pack2(acc[2] * sc, acc[3] * sc);
  o.z = Cvt<T>::pack2(acc[4] * sc, acc[5] * sc);
  o.w = Cvt<T>::pack2(acc[6] * sc, acc[7] * sc);
  *reinterpret_cast<uint4*>(da + pix * 64 + c8 * 8) = o;
}

// dw[c][r][s] = sum over padded pixels (u,v) of in_pad[u][v][c] * dv[u-r][v-s]   (dv zero outside the image)
// block = 256 threads = 64 channels x 4 tap groups; one padded row per iteration, the 7 dv rows it meets staged in
// shared memory with zero borders.  partial: [gridDim.x][64*49].
constexpr int kOutWgradTapsPerGroup = 13;
template <typename T>
__global__ void __launch_bounds__(256)
out_conv_wgrad_kernel(const T* __restrict__ in_pad, const float* __restrict__ dv, float* __restrict__ partial, int B, int H, int W) {
  extern __shared__ float rows[];                 // [7][W + 12], rows[r][j] = dv[u - r][j - 6]
  const int Wp = W + 6, Hp = H + 6, pitch = W + 12;
  const int c = threadIdx.x & 63, tg = threadIdx.x >> 6;
  const int tap0 = tg * kOutWgradTapsPerGroup;
  float acc[kOutWgradTapsPerGroup];
#pragma unroll
  for (int i = 0; i < kOutWgradTapsPerGroup; ++i) acc[i] = 0.f;
  int roff[kOutWgradTapsPerGroup];                // smem offset of tap (r,s) relative to column v: r*pitch + 6 - s
#pragma unroll
  for (int i = 0; i < kOutWgradTapsPerGroup; ++i) {
    const int tap = min(tap0 + i, 48);
    roff[i] = (tap / 7) * pitch + 6 - (tap % 7);
  }
  for (int row = blockIdx.x; row < B * Hp; row += gridDim.x) {
    const int b = row / Hp, u = row % Hp;
    __syncthreads();
    for (int i = threadIdx.x; i < 7 * pitch; i += 256) {
      const int r = i / pitch, j = i % pitch, yo = u - r, xo = j - 6;
      rows[i] = (yo >= 0 && yo < H && xo >= 0 && xo < W) ? dv[((long long)b * H + yo) * W + xo] : 0.f;
    }
    __syncthreads();
    const T* src = in_pad + ((long long)row * Wp) * 64 + c;
#pragma unroll 2
    for (int v = 0; v < Wp; ++v) {
      const float a = Cvt<T>::to_f(src[(long long)v * 64]);
#pragma unroll
      for (int i = 0; i < kOutWgradTapsPerGroup; ++i) acc[i] += a * rows[roff[i] + v];
    }
  }
  float* dst = partial + (long long)blockIdx.x * (64 * 49) + c * 49;
#pragma unroll
  for (int i = 0; i < kOutWgradTapsPerGroup; ++i)
    if (tap0 + i < 49) dst[tap0 + i] = acc[i];
}

// tap groups: 0 -> taps 0..12, 1 -> 13..25, 2 -> 26..38, 3 -> 39..48 (10 taps; the clamp above keeps its tail idle)
__global__ void out_conv_wgrad_reduce_kernel(const float* __restrict__ partial, int blocks, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 49) return;
  float s = 0.f;
  for (int k = 0; k < blocks; ++k) s += partial[(long long)k * (64 * 49) + i];
  dw[i] = s;
}

// ------------------------------------------------------------------ stem backward (modules/model.py:90-92)
// dcol [B][H][W][64] 16-bit holds, in column k = r*7 + s (k < 49), the gradient of the im2col entry (pixel, channel 0,
// tap (r,s)); the image gradient sums the entries that read the same padded pixel, then folds the reflection.
template <typename T>
__global__ void __launch_bounds__(256)
stem_col2im_kernel(const T* __restrict__ dcol, float* __restrict__ dx, const float* __restrict__ gs, int B, int H, int W) {
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= (long long)B * H * W) return;
  const int x = int(pix % W), y = int((pix / W) % H), b = int(pix / ((long long)W * H));
  const T* base = dcol + (long long)b * H * W * 64;
  int us[3], vs[3];
  const int nu = reflect_sources(y, H, 3, us), nv = reflect_sources(x, W, 3, vs);
  float acc = 0.f;
  for (int iu = 0; iu < nu; ++iu)
    for (int iv = 0; iv < nv; ++iv) {
      const int u = us[iu], v = vs[iv];
#pragma unroll
      for (int r = 0; r < 7; ++r) {
        const int yo = u - r;
        if (yo < 0 || yo >= H) continue;
#pragma unroll
        for (int s = 0; s < 7; ++s) {
          const int xo = v - s;
          if (xo < 0 || xo >= W) continue;
          acc += Cvt<T>::to_f(base[((long long)yo * W + xo) * 64 + r * 7 + s]);
        }
      }
    }
  dx[pix] = acc * gs[1];
}

// packed stem weight gradient [64][Kpad] (k = c*49 + r*7 + s, the im2col column order) -> OIHW [64][Cin][7][7], true scale
__global__ void unpack_stem_wgrad_kernel(const float* __restrict__ packed, float* __restrict__ g, int Cin, int Kpad,
                                         const float* __restrict__ gs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = Cin * 49;
  if (i >= 64 * K) return;
  g[i] = packed[(i / K) * Kpad + (i % K)] * gs[1];
}

// a += b on 16-bit maps (the skip connection of the residual blocks carries the gradient straight through)
template <typename T>
__global__ void __launch_bounds__(256)
add_inplace_kernel(T* __restrict__ a, const T* __restrict__ b, long long n8) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n8) return;
  const uint4 va = reinterpret_cast<const uint4*>(a)[i], vb = reinterpret_cast<const uint4*>(b)[i];
  const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 fa = Cvt<T>::unpack2(wa[k]), fb = Cvt<T>::unpack2(wb[k]);
    o[k] = Cvt<T>::pack2(fa.x + fb.x, fa.y + fb.y);
  }
  reinterpret_cast<uint4*>(a)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" int ducosy_add_inplace(void* a, const void* b, long long n, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(a && b && n > 0 && n % 8 == 0, DUCOSY_ERR_ARG, "add_inplace: needs non-null pointers and n %% 8 == 0");
  DUCOSY_DISPATCH_DTYPE(dtype, T, (add_inplace_kernel<T><<<grid_items(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
                                      static_cast<T*>(a), static_cast<const T*>(b), n / 8)));
  return check_launch("add_inplace_kernel");
}

namespace {
constexpr int kTanhBlocks = 592;
constexpr int kOutWgradBlocks = 592;
}

extern "C" size_t ducosy_out_conv_backward_scratch_bytes(int B, int H, int W) {
  return (size_t(B) * H * W + kTanhBlocks + size_t(kOutWgradBlocks) * 64 * 49) * sizeof(float);
}

extern "C" int ducosy_out_conv_backward(const float* dout, const float* out, const void* in_pad, const float* w, void* da,
                                        float* dw, float* db, float* scratch, const float* gs, int B, int H, int W, int dtype,
                                        ducosy_stream_t stream) {
  DUCOSY_CHECK(dout && out && in_pad && w && da && dw && db && scratch && gs, DUCOSY_ERR_ARG, "out_conv_backward: null pointer");
  DUCOSY_CHECK(B > 0 && H >= 8 && W >= 8, DUCOSY_ERR_SHAPE, "out_conv_backward: needs B > 0 and H, W >= 8 (got %d x %d x %d)", B, H, W);
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "out_conv_backward: bad dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * H * W;
  float* dv = scratch;
  float* tpart = dv + n;
  float* wpart = tpart + kTanhBlocks;
  out_tanh_bwd_kernel<<<kTanhBlocks, 256, 0, st>>>(dout, out, dv, tpart, n);
  DUCOSY_TRY(check_launch("out_tanh_bwd_kernel"));
  sum_partials_kernel<<<1, 256, 0, st>>>(tpart, kTanhBlocks, db);
  DUCOSY_TRY(check_launch("sum_partials_kernel"));
  DUCOSY_DISPATCH_DTYPE(dtype, T, (out_conv_dgrad_kernel<T><<<grid_items(n * 8, 256), 256, 0, st>>>(dv, w, static_cast<T*>(da), gs, B, H, W)));
  DUCOSY_TRY(check_launch("out_conv_dgrad_kernel"));
  const int blocks = min(kOutWgradBlocks, B * (H + 6));
  const size_t smem = size_t(7) * (W + 12) * sizeof(float);
  DUCOSY_CHECK(smem <= 48 * 1024, DUCOSY_ERR_SHAPE, "out_conv_backward: W too large (%d)", W);
  DUCOSY_DISPATCH_DTYPE(dtype, T, (out_conv_wgrad_kernel<T><<<blocks, 256, smem, st>>>(static_cast<const T*>(in_pad), dv, wpart, B, H, W)));
  DUCOSY_TRY(check_launch("out_conv_wgrad_kernel"));
  out_conv_wgrad_reduce_kernel<<<(64 * 49 + 255) / 256, 256, 0, st>>>(wpart, blocks, dw);
  return check_launch("out_conv_wgrad_reduce_kernel");
}

extern "C" int ducosy_stem_col2im(const void* dcol, float* dx, const float* gs, int B, int H, int W, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(dcol && dx && gs && B > 0 && H >= 8 && W >= 8, DUCOSY_ERR_ARG, "stem_col2im: bad argument");
  const long long n = (long long)B * H * W;
  DUCOSY_DISPATCH_DTYPE(dtype, T, (stem_col2im_kernel<T><<<grid_items(n, 256), 256, 0, (cudaStream_t)stream>>>(
                                      static_cast<const T*>(dcol), dx, gs, B, H, W)));
  return check_launch("stem_col2im_kernel");
}

extern "C" int ducosy_unpack_stem_wgrad(const float* packed, float* g_oihw, int Cin, int Kpad, const float* gs, ducosy_stream_t stream) {
  DUCOSY_CHECK(packed && g_oihw && gs && Cin >= 1 && Kpad >= Cin * 49, DUCOSY_ERR_ARG, "unpack_stem_wgrad: bad argument");
  unpack_stem_wgrad_kernel<<<grid_items(64LL * Cin * 49, 256), 256, 0, (cudaStream_t)stream>>>(packed, g_oihw, Cin, Kpad, gs);
  return check_launch("unpack_stem_wgrad_kernel");
}
