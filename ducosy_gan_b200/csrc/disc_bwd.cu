// PatchGAN discriminator backward (what autograd computes for modules/trainer.py:518-524 through modules/model.py:118-131).
// Tensor-core work reuses the forward machinery:
//   * input gradient of the 4x4 stride-2 convs = four phase-specific 2x2 convs over the zero-padded output gradient
//     (the adjoint of a stride-2 conv is "fractionally strided"): exactly the sub-pixel launch of conv_gemm.cu with a
//     transposed weight packing;
//   * weight gradients = conv_wgrad.cu (MN-major tcgen05 GEMM over the pixels).
// InstanceNorm + LeakyReLU backward, the first (Cin = 1) and last (Cout = 1) layers are bandwidth-bound kernels here.
// All reductions run in a fixed order (deterministic gradients).
#include <algorithm>

#include "common.cuh"

namespace ducosy {
namespace {

__device__ __forceinline__ float act_grad(float n, int act) {  // derivative of the activation at pre-activation n
  if (act == DUCOSY_ACT_RELU) return n > 0.f ? 1.f : 0.f;
  if (act == DUCOSY_ACT_LRELU02) return n > 0.f ? 1.f : 0.2f;
  return 1.f;
}

// ------------------------------------------------------------------ gradient scaling for the 16-bit backward maps
// Loss gradients are tiny (mean-reduced losses divide by the element count) and would sit in the fp16 subnormal range.
// The incoming gradient is multiplied by a power of two that brings its max magnitude into [1, 2); the backward is
// linear, so the final fp32 parameter / input gradients are simply multiplied by the inverse.  gs[0] = scale, gs[1] = 1/scale.
// Two kernels: the max magnitude is collected with an integer atomicMax on the bit pattern (non-negative floats order
// like their bits; max is exact and order-independent, so the result is deterministic), then one thread derives the scale.
__global__ void __launch_bounds__(256)
grad_absmax_kernel(const float* __restrict__ g, long long n, int* __restrict__ slot) {
  pdl_prologue();
  __shared__ float red[8];
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float v = fabsf(g[i]);
    m = (v > m || v != v) ? v : m;   // NaN propagates to the "not finite" branch below
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, m, o);
    m = (t > m || t != t) ? t : m;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = (red[i] > m || red[i] != red[i]) ? red[i] : m;
    atomicMax(slot, __float_as_int(m));   // NaN bit patterns (> 0x7f800000) win the max as well
  }
}
__global__ void grad_scale_finalize_kernel(float* __restrict__ gs) {
  pdl_prologue();
  const float m = __int_as_float(reinterpret_cast<int*>(gs)[1]);
  int e = 0;
  if (m > 0.f && isfinite(m)) {
    frexpf(m, &e);             // m = f * 2^e, f in [0.5, 1)
    e = 1 - e;                 // m * 2^e in [1, 2)
    e = max(-60, min(60, e));
  }
  gs[0] = ldexpf(1.f, e);
  gs[1] = ldexpf(1.f, -e);
}

// ------------------------------------------------------------------ InstanceNorm(+activation) backward
// forward: n = y*rstd + shift, a = act(n).  Given da: g = da * act'(n);
//   dy = rstd * (g - mean(g) - n * mean(g*n))          (per sample and channel, means over H*W)
// pass 1: per-block partial sums [B][blocks][2][C];  pass 2 (in_bwd_finalize): fixed-order sum -> [B][2][C];
// pass 3: dy written 16-bit into a zero-padded buffer.
// kFold: `da` is not the gradient map itself but the gradient w.r.t. the map padded by ONE pixel, [B][H+2][W+2][C] (what the
// 3x3 dgrad convolution writes): the streaming kernels read its interior directly, so the folded map never makes its own round
// trip through HBM.  For ReflectionPad2d the border rows / columns that mirror onto source rows 1, H-2 and columns 1, W-2 are
// added into those interior cells beforehand, in place, by pad1_reflect_border_kernel (a few microseconds: 2W + 2(H-2) pixels
// per sample; reads touch only the outermost ring, writes only cells of the next-but-one ring, so there is no ordering hazard).
// (A first version added the mirrored terms inside the streaming loops: its temporaries cost them 50 registers and half their
// occupancy -- ncu: reduce 33 -> 55 us, apply 42 -> 55 us per 134 MB map.)
template <typename T>
__global__ void __launch_bounds__(256)
pad1_reflect_border_kernel(T* __restrict__ dxpad, int B, int H, int W, int C) {
  pdl_prologue();
  const int cv = C / 8, per_sample = 2 * W + 2 * (H - 2);
  const long long item = (long long)blockIdx.x * 256 + threadIdx.x;
  if (item >= (long long)B * per_sample * cv) return;
  const int c8 = int(item % cv);
  const long long pid = item / cv;
  const int b = int(pid / per_sample), idx = int(pid - (long long)b * per_sample);
  int y, x;
  if (idx < W) { y = 1; x = idx; }
  else if (idx < 2 * W) { y = H - 2; x = idx - W; }
  else { const int k = idx - 2 * W, j = k >> 1; y = j == 0 ? 0 : (j == H - 3 ? H - 1 : j + 1); x = (k & 1) ? W - 2 : 1; }   // rows 0, 2..H-3, H-1
  const int ry = y == 1 ? 0 : (y == H - 2 ? H + 1 : -1), cx = x == 1 ? 0 : (x == W - 2 ? W + 1 : -1);
  uint4* base = reinterpret_cast<uint4*>(dxpad) + (size_t(b) * (H + 2) * (W + 2)) * cv + c8;
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto add = [&](int r, int c) {
    const uint4 v = base[(size_t(r) * (W + 2) + c) * cv];
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 q = Cvt<T>::unpack2(w4[k]);
      f[2 * k] += q.x;
      f[2 * k + 1] += q.y;
    }
  };
  add(y + 1, x + 1);                      // same summation order as pad_fold_kernel: interior first, then the mirrored cells
  if (ry >= 0) add(ry, x + 1);
  if (cx >= 0) {
    add(y + 1, cx);
    if (ry >= 0) add(ry, cx);
  }
  base[(size_t(y + 1) * (W + 2) + x + 1) * cv] =
      make_uint4(Cvt<T>::pack2(f[0], f[1]), Cvt<T>::pack2(f[2], f[3]), Cvt<T>::pack2(f[4], f[5]), Cvt<T>::pack2(f[6], f[7]));
}

template <typename T, bool kFold>
__device__ __forceinline__ void in_bwd_reduce_body(const T* __restrict__ da, const T* __restrict__ y, const float* __restrict__ scale,
                                                   const float* __restrict__ shift, float* __restrict__ partial, int HW, int C, int act,
                                                   int pix_per_block, int H /* kFold: log2(W) */, int W) {
  pdl_prologue();
  extern __shared__ float red[];  // [rows][2][C] with rows = 256 / (C/8)
  const int cv = C / 8, c8 = threadIdx.x % cv, prow = threadIdx.x / cv, rows = 256 / cv;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, HW);
  float sc[8], sh[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[b * C + c8 * 8 + j];
    sh[j] = shift[b * C + c8 * 8 + j];
    s1[j] = 0.f;
    s2[j] = 0.f;
  }
  const uint4* dav = reinterpret_cast<const uint4*>(da) + (kFold ? size_t(b) * (HW / W + 2) * (W + 2) : size_t(b) * HW) * cv + c8;
  const uint4* yv = reinterpret_cast<const uint4*>(y) + (size_t(b) * HW) * cv + c8;
  constexpr int kILP = 4;   // 8 independent 16-byte loads in flight per thread
  for (int pb = p0 + prow; pb < p1; pb += rows * kILP) {
    uint4 a[kILP], v[kILP];
#pragma unroll
    for (int u = 0; u < kILP; ++u) {
      const int p = pb + u * rows;
      const bool ok = p < p1;
      if (kFold) {   // W is a power of two here (checked on the host): H carries log2(W), no division in the streaming loop
        // padded index of pixel p = (py + 1) * (W + 2) + px + 1 = p + 2 * py + W + 3   (32-bit: one sample's map is < 2^31 chunks)
        const unsigned ip = unsigned(p + 2 * (p >> H) + W + 3) * unsigned(cv);
        a[u] = ok ? dav[ip] : make_uint4(0, 0, 0, 0);
      } else {
        a[u] = ok ? dav[size_t(p) * cv] : make_uint4(0, 0, 0, 0);
      }
      v[u] = ok ? yv[size_t(p) * cv] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kILP; ++u) {
      const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, vw[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fa = Cvt<T>::unpack2(aw[k]), fy = Cvt<T>::unpack2(vw[k]);
        const float n0 = fmaf(fy.x, sc[2 * k], sh[2 * k]), n1 = fmaf(fy.y, sc[2 * k + 1], sh[2 * k + 1]);
        const float g0 = fa.x * act_grad(n0, act), g1 = fa.y * act_grad(n1, act);   // masked-out loads have a = 0: g = 0
        s1[2 * k] += g0;
        s1[2 * k + 1] += g1;
        s2[2 * k] = fmaf(g0, n0, s2[2 * k]);
        s2[2 * k + 1] = fmaf(g1, n1, s2[2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(prow * 2 + 0) * C + c8 * 8 + j] = s1[j];
    red[(prow * 2 + 1) * C + c8 * 8 + j] = s2[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc += red[r * 2 * C + i];
    partial[(size_t(b) * gridDim.x + blockIdx.x) * 2 * C + i] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
in_bwd_reduce_kernel(const T* __restrict__ da, const T* __restrict__ y, const float* __restrict__ scale,
                     const float* __restrict__ shift, float* __restrict__ partial, int HW, int C, int act, int pix_per_block) {
  in_bwd_reduce_body<T, false>(da, y, scale, shift, partial, HW, C, act, pix_per_block, 0, 1);
}
// the padded addressing costs registers: three CTAs per SM are kept by capping them (20 bytes of spill)
template <typename T>
__global__ void __launch_bounds__(256, 3)
in_bwd_reduce_fold_kernel(const T* __restrict__ da_pad1, const T* __restrict__ y, const float* __restrict__ scale,
                          const float* __restrict__ shift, float* __restrict__ partial, int HW, int C, int act, int pix_per_block,
                          int log2W, int W) {
  in_bwd_reduce_body<T, true>(da_pad1, y, scale, shift, partial, HW, C, act, pix_per_block, log2W, W);
}

// grid (ceil(2C / 32), B), 256 threads = 32 columns x 8 slices of the partial rows; fixed-order double sums (deterministic).
// (One thread per column walking all rows serially was fine for 32 rows; with the finer reduction grid of small batches
// -- up to 256 rows for one 128x128 sample -- it became the longest kernel of the InstanceNorm backward.)
__global__ void __launch_bounds__(256)
in_bwd_finalize_kernel(const float* __restrict__ partial, float* __restrict__ sums, int blocks, int C, float inv_hw) {
  pdl_prologue();
  __shared__ double part[8][33];
  const int b = blockIdx.y;
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
  double acc = 0.0;
  if (col < 2 * C)
    for (int k = slice; k < blocks; k += 8) acc += double(partial[(size_t(b) * blocks + k) * 2 * C + col]);
  part[slice][threadIdx.x & 31] = acc;
  __syncthreads();
  if (slice == 0 && col < 2 * C) {
    double t = 0.0;
#pragma unroll
    for (int s = 0; s < 8; ++s) t += part[s][threadIdx.x];
    sums[size_t(b) * 2 * C + col] = float(t * inv_hw);   // means
  }
}

// grid (row CTAs, B): a CTA walks padded rows py = blockIdx.x, += gridDim.x of its sample; thread t handles the 16-byte
// chunks t, t+256, ... of a row -- 256 is a multiple of C/8, so its 8 channels (and their rstd / shift / means) stay in
// registers; four chunks are in flight per thread.
template <typename T, bool kFold>
__global__ void __launch_bounds__(256)
in_bwd_apply_pad_kernel(const T* __restrict__ da, const T* __restrict__ y, const float* __restrict__ scale,
                        const float* __restrict__ shift, const float* __restrict__ means, T* __restrict__ dy_pad, int B,
                        int H, int W, int C, int pad, int act, int fold_mode) {
  pdl_prologue();
  const int Hp = H + 2 * pad, Wp = W + 2 * pad, cv = C / 8;
  const int b = blockIdx.y, c8 = threadIdx.x % cv, px0 = threadIdx.x / cv, px_step = 256 / cv;
  float rs[8], sh[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c8 * 8 + j;
    rs[j] = scale[b * C + c];
    sh[j] = shift[b * C + c];
    m1[j] = means[size_t(b) * 2 * C + c];
    m2[j] = means[size_t(b) * 2 * C + C + c];
  }
  constexpr int kILP = 4;
  for (int py = blockIdx.x; py < Hp; py += gridDim.x) {
    const int sy = py - pad;
    uint4* dst_row = reinterpret_cast<uint4*>(dy_pad) + ((size_t(b) * Hp + py) * Wp) * cv + c8;
    if (sy < 0 || sy >= H) {
      for (int px = px0; px < Wp; px += px_step) dst_row[size_t(px) * cv] = make_uint4(0, 0, 0, 0);
      continue;
    }
    const uint4* a_pad = reinterpret_cast<const uint4*>(da) + (size_t(b) * (H + 2) * (W + 2)) * cv + c8;   // kFold: sample base
    const uint4* a_row = kFold ? a_pad + (size_t(sy + 1) * (W + 2) + 1) * cv
                               : reinterpret_cast<const uint4*>(da) + ((size_t(b) * H + sy) * W) * cv + c8;
    const uint4* y_row = reinterpret_cast<const uint4*>(y) + ((size_t(b) * H + sy) * W) * cv + c8;
    for (int pb = px0; pb < Wp; pb += px_step * kILP) {
      uint4 a[kILP], v[kILP];
      bool inside[kILP];
#pragma unroll
      for (int u = 0; u < kILP; ++u) {
        const int sx = pb + u * px_step - pad;
        inside[u] = sx >= 0 && sx < W;
        a[u] = inside[u] ? a_row[size_t(sx) * cv] : make_uint4(0, 0, 0, 0);
        v[u] = inside[u] ? y_row[size_t(sx) * cv] : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < kILP; ++u) {
        const int px = pb + u * px_step;
        if (px >= Wp) break;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (inside[u]) {
          const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, vw[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          uint32_t ow[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fa = Cvt<T>::unpack2(aw[k]), fy = Cvt<T>::unpack2(vw[k]);
            const float n0 = fmaf(fy.x, rs[2 * k], sh[2 * k]), n1 = fmaf(fy.y, rs[2 * k + 1], sh[2 * k + 1]);
            const float g0 = fa.x * act_grad(n0, act), g1 = fa.y * act_grad(n1, act);
            ow[k] = Cvt<T>::pack2(rs[2 * k] * (g0 - m1[2 * k] - n0 * m2[2 * k]),
                                  rs[2 * k + 1] * (g1 - m1[2 * k + 1] - n1 * m2[2 * k + 1]));
          }
          o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        dst_row[size_t(px) * cv] = o;
      }
    }
  }
}

// ------------------------------------------------------------------ dgrad weight packing for the 4x4 stride-2 convs
// Wd[phase*Ci + c][(a*2+b)*Co + o] = W[o][c][r(py,a)][s(px,b)]
//   4x4 kernel: r(0,.) = {3,1}, r(1,.) = {2,0};  3x3 kernel (model.py:96-98): r(0,.) = {-,1}, r(1,.) = {2,0} (- = no tap: zero)
template <typename T>
__global__ void pack_dgrad_s2_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int Co, int Ci, int ksz) {
  pdl_prologue();
  const long long total = 16LL * Co * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o = int(i % Co);
    long long r = i / Co;
    const int tap = int(r % 4);
    r /= 4;
    const int c = int(r % Ci);
    const int phase = int(r / Ci);
    const int py = phase >> 1, px = phase & 1, a = tap >> 1, b = tap & 1;
    const int rr = py == 0 ? (a == 0 ? 3 : 1) : (a == 0 ? 2 : 0);
    const int ss = px == 0 ? (b == 0 ? 3 : 1) : (b == 0 ? 2 : 0);
    out[i] = Cvt<T>::from_f((rr < ksz && ss < ksz) ? w[(((long long)o * Ci + c) * ksz + rr) * ksz + ss] : 0.f);
  }
}

// ------------------------------------------------------------------ 3x3 stride-1 input gradient
// Wd[c][(r*3+s)*Co + o] = W[o][c][r][s]
template <typename T>
__global__ void pack_dgrad_s1_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int Co, int Ci) {
  pdl_prologue();
  const long long total = 9LL * Co * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o = int(i % Co);
    long long r = i / Co;
    const int tap = int(r % 9);
    const int c = int(r / 9);
    out[i] = Cvt<T>::from_f(w[((long long)o * Ci + c) * 9 + tap]);
  }
}
// the two outer columns of the padded-input gradient (v = 0 uses filter column s = 0, v = W+1 uses s = 2):
//   dxpad[u][v][c] = sum_{r,o} dy[u-r][v-s][o] * W[o][c][r][s]     (dy_pad2 has a zero border of 2)
template <typename T>
__global__ void dgrad_s1_edge_cols_kernel(const T* __restrict__ dy_pad2, const T* __restrict__ wd, T* __restrict__ dxpad, int B, int H,
                                          int W, int Ci, int Co) {
  pdl_prologue();
  const long long total = (long long)B * (H + 2) * 2 * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % Ci);
    long long q = i / Ci;
    const int side = int(q & 1);
    q >>= 1;
    const int u = int(q % (H + 2));
    const int b = int(q / (H + 2));
    const int v = side ? W + 1 : 0, s = side ? 2 : 0;
    float acc = 0.f;
    for (int r = 0; r < 3; ++r) {
      const T* d = dy_pad2 + (((long long)b * (H + 4) + (u - r + 2)) * (W + 4) + (v - s + 2)) * Co;
      const T* wv = wd + ((long long)c * 9 + r * 3 + s) * Co;
      for (int o = 0; o < Co; o += 8) {
        const uint4 dv = *reinterpret_cast<const uint4*>(d + o), kv = *reinterpret_cast<const uint4*>(wv + o);
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, kw[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = Cvt<T>::unpack2(dw[k]), w2 = Cvt<T>::unpack2(kw[k]);
          acc = fmaf(a.x, w2.x, acc);
          acc = fmaf(a.y, w2.y, acc);
        }
      }
    }
    dxpad[(((long long)b * (H + 2) + u) * (W + 2) + v) * Ci + c] = Cvt<T>::from_f(acc);
  }
}
// Co = 256 (the residual blocks): a CTA = 8 input channels (one per warp, its 3 x 256 filter taps in registers, 8 per
// lane) x one chunk of rows of one sample; the dy rows the chunk meets are staged once in shared memory, so L2 sees each
// of them once per 8 channels instead of once per channel.
template <typename T>
__global__ void __launch_bounds__(256)
dgrad_s1_edge_cols256_kernel(const T* __restrict__ dy_pad2, const T* __restrict__ wd, T* __restrict__ dxpad, int B, int H, int W,
                             int Ci, int rows_per_chunk, int chunks_per_sample) {
  pdl_prologue();
  constexpr int Co = 256;
  extern __shared__ uint4 srows[];   // [rows_per_chunk + 2][32] 16-byte pieces: dy_pad2 rows u0 .. u1+1 at column v - s + 2
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp, side = blockIdx.y;
  const int b = blockIdx.z / chunks_per_sample, chunk = blockIdx.z % chunks_per_sample;
  const int u0 = chunk * rows_per_chunk, u1 = min(H + 2, u0 + rows_per_chunk);
  if (u0 >= u1) return;
  const int v = side ? W + 1 : 0, s = side ? 2 : 0;
  const int nrows = u1 - u0 + 2;
  for (int i = threadIdx.x; i < nrows * 32; i += 256) {
    const int j = i >> 5, piece = i & 31;
    srows[i] = *reinterpret_cast<const uint4*>(dy_pad2 + (((long long)b * (H + 4) + (u0 + j)) * (W + 4) + (v - s + 2)) * Co + piece * 8);
  }
  float wr[3][8];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint4 kv = *reinterpret_cast<const uint4*>(wd + ((long long)c * 9 + r * 3 + s) * Co + lane * 8);
    const uint32_t kw[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = Cvt<T>::unpack2(kw[k]);
      wr[r][2 * k] = f.x;
      wr[r][2 * k + 1] = f.y;
    }
  }
  __syncthreads();
  for (int u = u0; u < u1; ++u) {
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const uint4 dv = srows[(u - r + 2 - u0) * 32 + lane];   // dy_pad2 row u - r + 2
      const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = Cvt<T>::unpack2(dw[k]);
        acc = fmaf(a.x, wr[r][2 * k], acc);
        acc = fmaf(a.y, wr[r][2 * k + 1], acc);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dxpad[(((long long)b * (H + 2) + u) * (W + 2) + v) * Ci + c] = Cvt<T>::from_f(acc);
  }
}
// Ci = Co = 256 (the residual blocks): the two edge columns are a [16 positions] x [768] x [256 channels] GEMM per CTA on
// the warp-level tensor cores (mma.sync m16n8k16, fp32 accumulate): A = the dy rows the 16 positions meet (shared
// memory, rows padded by 16 bytes against bank conflicts), B = the packed dgrad weights read straight from L2.
template <typename T> struct EdgeMma;
template <> struct EdgeMma<__half> {
  static __device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
};
template <> struct EdgeMma<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
};
constexpr int kEdgePitch = 256 + 8;   // elements per staged dy row
// grid (8 channel groups x row_splits, 2 sides, B), 128 threads.  A CTA owns 32 output channels of one edge column of one
// sample: warp w holds the B fragments of its 8 channels for the WHOLE K = 3 filter rows x 256 input channels in 96 registers
// (read once), and the CTA walks the 16-row tiles of the column (tile, tile + row_splits, ...), staging the 18 dy rows a tile
// needs in shared memory.  (The first version re-read every B fragment from L2 for every MMA: 132 us per launch at batch 16
// for 1.6 GFLOP.)
template <typename T>
__global__ void __launch_bounds__(128)
dgrad_s1_edge_cols_mma_kernel(const T* __restrict__ dy_pad2, const T* __restrict__ wd, T* __restrict__ dxpad, int B, int H, int W,
                              int row_splits) {
  pdl_prologue();
  constexpr int C = 256;
  __shared__ __align__(16) T rows[18 * kEdgePitch];   // rows[j][o] = dy_pad2[b][u0 + j][v - s + 2][o], j = m - r + 2
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tid = lane & 3;
  const int cg = blockIdx.x / row_splits, rs = blockIdx.x % row_splits, side = blockIdx.y, b = blockIdx.z;
  const int v = side ? W + 1 : 0, s = side ? 2 : 0;
  const int c0 = cg * 32 + warp * 8;
  uint32_t bf[3][16][2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const T* wrow = wd + ((long long)(c0 + g) * 9 + r * 3 + s) * C + 2 * tid;
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      bf[r][kk][0] = __ldg(reinterpret_cast<const unsigned int*>(wrow + kk * 16));
      bf[r][kk][1] = __ldg(reinterpret_cast<const unsigned int*>(wrow + kk * 16 + 8));
    }
  }
  const int tiles = (H + 2 + 15) / 16;
  for (int tile = rs; tile < tiles; tile += row_splits) {
    const int u0 = tile * 16;
    __syncthreads();
    for (int i = threadIdx.x; i < 18 * 32; i += 128) {
      const int j = i >> 5, piece = i & 31;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (u0 + j < H + 4)
        val = *reinterpret_cast<const uint4*>(dy_pad2 + (((long long)b * (H + 4) + (u0 + j)) * (W + 4) + (v - s + 2)) * C + piece * 8);
      *reinterpret_cast<uint4*>(rows + j * kEdgePitch + piece * 8) = val;
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const T* ra = rows + (g - r + 2) * kEdgePitch + 2 * tid;        // position m = g
      const T* rb = ra + 8 * kEdgePitch;                               // position m = g + 8
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const int o0 = kk * 16;
        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(ra + o0), a1 = *reinterpret_cast<const uint32_t*>(rb + o0);
        const uint32_t a2 = *reinterpret_cast<const uint32_t*>(ra + o0 + 8), a3 = *reinterpret_cast<const uint32_t*>(rb + o0 + 8);
        EdgeMma<T>::mma(acc, a0, a1, a2, a3, bf[r][kk][0], bf[r][kk][1]);
      }
    }
    const int c = c0 + 2 * tid;
    const int ua = u0 + g, ub = u0 + g + 8;
    if (ua < H + 2)
      *reinterpret_cast<uint32_t*>(dxpad + (((long long)b * (H + 2) + ua) * (W + 2) + v) * C + c) = Cvt<T>::pack2(acc[0], acc[1]);
    if (ub < H + 2)
      *reinterpret_cast<uint32_t*>(dxpad + (((long long)b * (H + 2) + ub) * (W + 2) + v) * C + c) = Cvt<T>::pack2(acc[2], acc[3]);
  }
}
// adjoint of the padding: gradient w.r.t. the padded map [B][H+2p][W+2p][C] -> gradient w.r.t. the un-padded map
template <typename T>
__global__ void pad_fold_kernel(const T* __restrict__ dxpad, const T* __restrict__ add, T* __restrict__ dx, int B, int H, int W, int C,
                                int pad, int mode) {
  pdl_prologue();
  const int cv = C / 8, Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long total = (long long)B * H * W * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = int(i % cv);
    long long r = i / cv;
    const int x = int(r % W);
    r /= W;
    const int y = int(r % H);
    const int b = int(r / H);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (add != nullptr) {   // the skip connection's gradient joins here (saves a separate pass over the map)
      const uint4 v = reinterpret_cast<const uint4*>(add)[i];
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = Cvt<T>::unpack2(w4[k]);
        acc[2 * k] = f.x;
        acc[2 * k + 1] = f.y;
      }
    }
    // padded rows / cols that reflect onto (y, x): p -> |p - pad| on the low side, 2(n-1) - (p - pad) on the high side
    int rows[3], cols[3], nr = 0, nc = 0;
    rows[nr++] = y + pad;
    cols[nc++] = x + pad;
    if (mode == DUCOSY_PAD_REFLECT) {
      if (y >= 1 && y <= pad) rows[nr++] = pad - y;
      if (y <= H - 2 && y >= H - 1 - pad) rows[nr++] = pad + 2 * (H - 1) - y;
      if (x >= 1 && x <= pad) cols[nc++] = pad - x;
      if (x <= W - 2 && x >= W - 1 - pad) cols[nc++] = pad + 2 * (W - 1) - x;
    }
    for (int a = 0; a < nr; ++a)
      for (int bb = 0; bb < nc; ++bb) {
        const uint4 v = reinterpret_cast<const uint4*>(dxpad)[(((long long)b * Hp + rows[a]) * Wp + cols[bb]) * cv + c8];
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = Cvt<T>::unpack2(w4[k]);
          acc[2 * k] += f.x;
          acc[2 * k + 1] += f.y;
        }
      }
    reinterpret_cast<uint4*>(dx)[i] = make_uint4(Cvt<T>::pack2(acc[0], acc[1]), Cvt<T>::pack2(acc[2], acc[3]),
                                                 Cvt<T>::pack2(acc[4], acc[5]), Cvt<T>::pack2(acc[6], acc[7]));
  }
}

// ------------------------------------------------------------------ Upsample(x2 nearest) + Conv3x3(pad 1) backward
// forward (conv_gemm sub-pixel form): out[2i+py][2j+px] = sum_{a,b} Wp[ph][a][b] . srcpad[i+a+py][j+b+px]; its adjoint
// gathers, for every source pixel, 16 (phase, tap) entries of the output gradient on the stride-2 sub-lattices:
//   Wd[c][(ph*4 + a*2 + b)*Co + o] = Wp[ph][a][b][o][c]       (Wp = the pre-summed phase weights of pack_upconv_weight)
template <typename T>
__global__ void pack_upconv_dgrad_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int Co, int Ci) {
  pdl_prologue();
  const long long total = 16LL * Co * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o = int(i % Co);
    long long r = i / Co;
    const int t16 = int(r % 16);
    const int c = int(r / 16);
    const int phase = t16 >> 2, a = (t16 >> 1) & 1, b = t16 & 1, py = phase >> 1, px = phase & 1;
    const int r0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), r1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
    const int s0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), s1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
    const float* wk = w + ((long long)o * Ci + c) * 9;
    float acc = 0.f;
    for (int rr = r0; rr <= r1; ++rr)
      for (int ss = s0; ss <= s1; ++ss) acc += wk[rr * 3 + ss];
    out[i] = Cvt<T>::from_f(acc);
  }
}
// nearest x2 upsampling of the (un-padded interior of the) zero-padded source + zero pad 1: the operand the weight
// gradient needs; only the backward materialises it.
template <typename T>
__global__ void upsample2x_pad_kernel(const T* __restrict__ src_pad, T* __restrict__ up_pad, int B, int Hs, int Ws, int C) {
  pdl_prologue();
  const int cv = C / 8, Hu = 2 * Hs + 2, Wu = 2 * Ws + 2;
  const long long total = (long long)B * Hu * Wu * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = int(i % cv);
    long long r = i / cv;
    const int x = int(r % Wu);
    r /= Wu;
    const int y = int(r % Hu);
    const int b = int(r / Hu);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 1 && y <= 2 * Hs && x >= 1 && x <= 2 * Ws)
      v = reinterpret_cast<const uint4*>(src_pad)[(((long long)b * (Hs + 2) + (y - 1) / 2 + 1) * (Ws + 2) + (x - 1) / 2 + 1) * cv + c8];
    reinterpret_cast<uint4*>(up_pad)[i] = v;
  }
}

// packed fp32 weight gradient [Co][taps*Ci + c] -> OIHW [Co][Ci][taps]
__global__ void unpack_wgrad_kernel(const float* __restrict__ packed, float* __restrict__ g, int Co, int Ci, int taps,
                                    const float* __restrict__ gs) {
  pdl_prologue();
  const float inv = gs != nullptr ? gs[1] : 1.f;
  const long long total = (long long)Co * Ci * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = int(i % taps);
    long long r = i / taps;
    const int c = int(r % Ci);
    const int o = int(r / Ci);
    g[i] = packed[((long long)o * taps + tap) * Ci + c] * inv;
  }
}

// ------------------------------------------------------------------ last layer (512 -> 1) backward
// forward: out[yo][xo] = b + sum_{r,s,c} p4[yo+r][xo+s][c] w[c][r][s], p4 = a4 zero-padded by 2.
// da4[y][x][c] = sum_{r,s} dout[y+2-r][x+2-s] w[c][r][s]
template <typename T>
__global__ void disc_last_dgrad_kernel(const float* __restrict__ dout, const T* __restrict__ wp /*[16][512]*/,
                                       T* __restrict__ da, int B, int Hs, int Ws, const float* __restrict__ gs) {
  pdl_prologue();
  const float gscale = gs != nullptr ? gs[0] : 1.f;
  const long long total = (long long)B * Hs * Ws * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = int(i & 63);
    long long r = i >> 6;
    const int x = int(r % Ws);
    r /= Ws;
    const int y = int(r % Hs);
    const int b = int(r / Hs);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 16; ++tap) {
      const int yo = y + 2 - (tap >> 2), xo = x + 2 - (tap & 3);
      if (yo < 0 || yo >= Hs || xo < 0 || xo >= Ws) continue;
      const float d = dout[((long long)b * Hs + yo) * Ws + xo] * gscale;
      const uint4 wv = *reinterpret_cast<const uint4*>(wp + tap * 512 + c8 * 8);
      const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = Cvt<T>::unpack2(ww[k]);
        acc[2 * k] = fmaf(d, f.x, acc[2 * k]);
        acc[2 * k + 1] = fmaf(d, f.y, acc[2 * k + 1]);
      }
    }
    reinterpret_cast<uint4*>(da)[i] = make_uint4(Cvt<T>::pack2(acc[0], acc[1]), Cvt<T>::pack2(acc[2], acc[3]),
                                                 Cvt<T>::pack2(acc[4], acc[5]), Cvt<T>::pack2(acc[6], acc[7]));
  }
}
// dw[c][tap] = sum_{b,yo,xo} dout * p4[b][yo+r][xo+s][c];  db = sum dout.  One block per tap, one thread per channel.
constexpr int kLastSplits = 32;
// grid (16 taps, kLastSplits): each block reduces a slice of the output rows; partial [split][8192 + 1]
template <typename T>
__global__ void __launch_bounds__(512)
disc_last_wgrad_kernel(const float* __restrict__ dout, const T* __restrict__ p4, float* __restrict__ partial, int B, int Hs, int Ws) {
  pdl_prologue();
  const int tap = blockIdx.x, c = threadIdx.x, r = tap >> 2, s = tap & 3;
  const int Hp = Hs + 4, Wp = Ws + 4;
  const int rows = B * Hs, r0 = (rows * blockIdx.y) / kLastSplits, r1 = (rows * (blockIdx.y + 1)) / kLastSplits;
  float acc = 0.f, dsum = 0.f;
  for (int row = r0; row < r1; ++row) {
    const int b = row / Hs, yo = row - b * Hs;
    const T* src = p4 + (((long long)b * Hp + yo + r) * Wp + s) * 512 + c;
    const float* dr = dout + (long long)row * Ws;
#pragma unroll 4
    for (int xo = 0; xo < Ws; ++xo) {
      const float d = dr[xo];
      acc = fmaf(d, Cvt<T>::to_f(src[(long long)xo * 512]), acc);
      dsum += d;
    }
  }
  partial[(size_t)blockIdx.y * 8193 + c * 16 + tap] = acc;
  if (tap == 0 && c == 0) partial[(size_t)blockIdx.y * 8193 + 8192] = dsum;
}
__global__ void disc_last_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > 8192) return;
  float acc = 0.f;
  for (int k = 0; k < kLastSplits; ++k) acc += partial[(size_t)k * 8193 + i];
  if (i < 8192) dw[i] = acc;
  else *db = acc;
}

// ------------------------------------------------------------------ first layer (1 -> 64) backward
// dpre = da1 * lrelu'(a1); dw1[o][tap] = sum dpre[pix][o] * x[2y+r-1][2x+s-1]; db1[o] = sum dpre[pix][o].
// thread = (channel o, tap group of 4); each block reduces a pixel range into partial[block][64][17].
template <typename T>
__global__ void __launch_bounds__(256)
disc_first_wgrad_kernel(const T* __restrict__ da1, const T* __restrict__ p1, const float* __restrict__ x,
                        float* __restrict__ partial, int B, int H, int W, int pix_per_block) {
  pdl_prologue();
  // thread = (pixel lane pl 0..7, tap group tg 0..3 = filter row, channel group cg 0..7 = 8 channels): a warp is one pixel
  // lane, its 128-byte da1 / p1 rows are read with 16-byte loads (shared by the 4 tap groups through L1)
  __shared__ float red[8][64 * 17];
  const int Ho = H / 2, Wo = W / 2;
  const int cg = threadIdx.x & 7, tg = (threadIdx.x >> 3) & 3, pl = threadIdx.x >> 5;
  const long long P = (long long)B * Ho * Wo;
  const long long q0 = (long long)blockIdx.x * pix_per_block, q1 = min(q0 + (long long)pix_per_block, P);
  float acc[8][4], bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bsum[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
  }
  for (long long q = q0 + pl; q < q1; q += 8) {
    const int xo = int(q % Wo);
    const long long r = q / Wo;
    const int yo = int(r % Ho), b = int(r / Ho);
    const uint4 aq = *reinterpret_cast<const uint4*>(p1 + (((long long)b * (Ho + 2) + yo + 1) * (Wo + 2) + xo + 1) * 64 + cg * 8);
    const uint4 dq = *reinterpret_cast<const uint4*>(da1 + q * 64 + cg * 8);
    const uint32_t aw4[4] = {aq.x, aq.y, aq.z, aq.w}, dw4[4] = {dq.x, dq.y, dq.z, dq.w};
    float xv[4];
    const int iy = 2 * yo + tg - 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ix = 2 * xo + k - 1;
      xv[k] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(x + ((long long)b * H + iy) * W + ix) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 av = Cvt<T>::unpack2(aw4[j]), dv = Cvt<T>::unpack2(dw4[j]);
      const float d0 = dv.x * (av.x > 0.f ? 1.f : 0.2f), d1 = dv.y * (av.y > 0.f ? 1.f : 0.2f);
      bsum[2 * j] += d0;
      bsum[2 * j + 1] += d1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[2 * j][k] = fmaf(d0, xv[k], acc[2 * j][k]);
        acc[2 * j + 1][k] = fmaf(d1, xv[k], acc[2 * j + 1][k]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int o = cg * 8 + j;
#pragma unroll
    for (int k = 0; k < 4; ++k) red[pl][o * 17 + tg * 4 + k] = acc[j][k];
    if (tg == 0) red[pl][o * 17 + 16] = bsum[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 17; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) a += red[l][i];          // fixed order: deterministic
    partial[size_t(blockIdx.x) * 64 * 17 + i] = a;
  }
}

__global__ void disc_first_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db,
                                               int blocks, const float* __restrict__ gs) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 64 * 17
  if (i >= 64 * 17) return;
  double acc = 0.0;
  for (int k = 0; k < blocks; ++k) acc += double(partial[size_t(k) * 64 * 17 + i]);
  const int o = i / 17, t = i % 17;
  if (gs != nullptr) acc *= double(gs[1]);
  if (t < 16) dw[o * 16 + t] = float(acc);
  else db[o] = float(acc);
}
// input gradient of the first layer: dx[iy][ix] = sum_{o, (r,s) with matching parity} dpre[(iy+1-r)/2][(ix+1-s)/2][o] w1[o][r][s]
template <typename T>
__global__ void disc_first_dgrad_kernel(const T* __restrict__ da1, const T* __restrict__ p1, const float* __restrict__ w1,
                                        float* __restrict__ dx, int B, int H, int W, const float* __restrict__ gs) {
  pdl_prologue();
  const float inv = gs != nullptr ? gs[1] : 1.f;
  __shared__ float sw[64 * 16];
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) sw[i] = w1[i];
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ix = int(i % W);
    long long r = i / W;
    const int iy = int(r % H);
    const int b = int(r / H);
    float acc = 0.f;
    for (int rr = (iy + 1) & 1; rr < 4; rr += 2) {
      const int yo = (iy + 1 - rr) / 2;
      if (iy + 1 - rr < 0 || yo >= Ho) continue;
      for (int ss = (ix + 1) & 1; ss < 4; ss += 2) {
        const int xo = (ix + 1 - ss) / 2;
        if (ix + 1 - ss < 0 || xo >= Wo) continue;
        const T* d = da1 + (((long long)b * Ho + yo) * Wo + xo) * 64;
        const T* a = p1 + (((long long)b * (Ho + 2) + yo + 1) * (Wo + 2) + xo + 1) * 64;
#pragma unroll 2
        for (int o = 0; o < 64; o += 8) {
          const uint4 dq = *reinterpret_cast<const uint4*>(d + o), aq = *reinterpret_cast<const uint4*>(a + o);
          const uint32_t dw4[4] = {dq.x, dq.y, dq.z, dq.w}, aw4[4] = {aq.x, aq.y, aq.z, aq.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 dv = Cvt<T>::unpack2(dw4[k]), av = Cvt<T>::unpack2(aw4[k]);
            acc = fmaf(dv.x * (av.x > 0.f ? 1.f : 0.2f), sw[(o + 2 * k) * 16 + rr * 4 + ss], acc);
            acc = fmaf(dv.y * (av.y > 0.f ? 1.f : 0.2f), sw[(o + 2 * k + 1) * 16 + rr * 4 + ss], acc);
          }
        }
      }
    }
    dx[i] = acc * inv;
  }
}

int grid_for_items(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = (long long)(num_sms() > 0 ? num_sms() : 148) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

namespace {
constexpr int kInBwdPix = 512;   // pixels per CTA of the InstanceNorm-backward reduction (large batches)

// Pixels per CTA of the reduction: 512 when that already gives every SM a few CTAs; fewer (down to 64) for small batches --
// at one sample per rank a 128x128 map was reduced by 32 CTAs on 148 SMs (32 us for 16 MB, 10x off the HBM roofline).
int in_bwd_pix_per_block(int B, int HW) {
  const long long want_ctas = 4LL * (num_sms() > 0 ? num_sms() : 148);
  long long ppb = ((long long)B * HW + want_ctas - 1) / want_ctas;
  ppb = (ppb + 63) / 64 * 64;
  if (ppb < 64) ppb = 64;
  if (ppb > kInBwdPix) ppb = kInBwdPix;
  return int(ppb);
}
}

// Generic InstanceNorm(+activation) backward on NHWC 16-bit maps (also the generator's building block):
//   da, y [B][H][W][C]; scale/shift from ducosy_in_finalize of the forward; scratch: fp32 [B*(blocks+1)*2*C] with
//   blocks = ceil(H*W / 64) at most (ducosy_in_backward_scratch_bytes); dy_pad [B][H+2p][W+2p][C] (zero border).
extern "C" size_t ducosy_in_backward_scratch_bytes(int B, int H, int W, int C) {
  const int blocks = (H * W + 63) / 64;      // upper bound for any pixels-per-CTA choice
  return size_t(B) * (blocks + 1) * 2 * C * 4;
}
namespace {
int in_backward_impl(const void* da, bool folded, int fold_mode, const void* y, const float* scale, const float* shift, void* dy_pad,
                     float* scratch, int B, int H, int W, int C, int pad, int act, int dtype, cudaStream_t st) {
  DUCOSY_CHECK(da && y && scale && shift && dy_pad && scratch && B > 0, DUCOSY_ERR_ARG, "in_backward_pad: null pointer");
  DUCOSY_CHECK(C % 8 == 0 && 256 % (C / 8) == 0 && pad >= 0, DUCOSY_ERR_SHAPE, "in_backward_pad: C/8 must divide 256");
  DUCOSY_CHECK(!folded || (H >= 5 && W >= 8 && (W & (W - 1)) == 0), DUCOSY_ERR_SHAPE,
               "in_backward_pad_folded: H >= 5 and W a power of two >= 8 (got %dx%d)", H, W);
  const int HW = H * W, ppb = in_bwd_pix_per_block(B, HW), blocks = (HW + ppb - 1) / ppb;
  float* partial = scratch;
  float* means = scratch + size_t(B) * blocks * 2 * C;
  const size_t smem = size_t(256 / (C / 8)) * 2 * C * 4;
  if (folded && fold_mode == DUCOSY_PAD_REFLECT) {
    // mirrored border rows / columns are added into the interior cells they reflect onto, in place (da_pad1 is consumed here)
    const long long items = (long long)B * (2 * W + 2 * (H - 2)) * (C / 8);
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pad1_reflect_border_kernel<T>, int((items + 255) / 256), 256, 0, st)(
                                        static_cast<T*>(const_cast<void*>(da)), B, H, W, C)));
    DUCOSY_TRY(check_launch("pad1_reflect_border_kernel"));
  }
  if (folded)
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(in_bwd_reduce_fold_kernel<T>, dim3(blocks, B), 256, smem, st)(
                                        static_cast<const T*>(da), static_cast<const T*>(y), scale, shift, partial, HW, C, act, ppb,
                                        __builtin_ctz(unsigned(W)), W)));
  else
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(in_bwd_reduce_kernel<T>, dim3(blocks, B), 256, smem, st)(
                                        static_cast<const T*>(da), static_cast<const T*>(y), scale, shift, partial, HW, C, act, ppb)));
  DUCOSY_TRY(check_launch("in_bwd_reduce_kernel"));
  pdl(in_bwd_finalize_kernel, dim3((2 * C + 31) / 32, B), 256, 0, st)(partial, means, blocks, C, 1.0f / float(HW));
  DUCOSY_TRY(check_launch("in_bwd_finalize_kernel"));
  const int row_ctas = std::max(1, std::min(H + 2 * pad, (num_sms() * 4 + B - 1) / B));
  if (folded)
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(in_bwd_apply_pad_kernel<T, true>, dim3(row_ctas, B), 256, 0, st)(
                                        static_cast<const T*>(da), static_cast<const T*>(y), scale, shift, means,
                                        static_cast<T*>(dy_pad), B, H, W, C, pad, act, fold_mode)));
  else
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(in_bwd_apply_pad_kernel<T, false>, dim3(row_ctas, B), 256, 0, st)(
                                        static_cast<const T*>(da), static_cast<const T*>(y), scale, shift, means,
                                        static_cast<T*>(dy_pad), B, H, W, C, pad, act, fold_mode)));
  return check_launch("in_bwd_apply_pad_kernel");
}
}  // namespace

extern "C" int ducosy_in_backward_pad(const void* da, const void* y, const float* scale, const float* shift, void* dy_pad,
                                      float* scratch, int B, int H, int W, int C, int pad, int act, int dtype,
                                      ducosy_stream_t stream) {
  return in_backward_impl(da, false, DUCOSY_PAD_ZERO, y, scale, shift, dy_pad, scratch, B, H, W, C, pad, act, dtype,
                          static_cast<cudaStream_t>(stream));
}

// The same with the padding adjoint of a pad-1 convolution folded into the loads: da_pad1 [B][H+2][W+2][C] is the gradient
// w.r.t. the PADDED map (ducosy_conv3x3s1_dgrad_nhwc's output); fold_mode = DUCOSY_PAD_REFLECT | DUCOSY_PAD_ZERO.  Replaces
// ducosy_pad_fold + ducosy_in_backward_pad (one full read + write of the map less).  da_pad1 is CONSUMED: with
// DUCOSY_PAD_REFLECT the mirrored border terms are added into its interior cells in place.
extern "C" int ducosy_in_backward_pad_folded(void* da_pad1, int fold_mode, const void* y, const float* scale, const float* shift,
                                             void* dy_pad, float* scratch, int B, int H, int W, int C, int pad, int act, int dtype,
                                             ducosy_stream_t stream) {
  DUCOSY_CHECK(fold_mode == DUCOSY_PAD_REFLECT || fold_mode == DUCOSY_PAD_ZERO, DUCOSY_ERR_ARG, "in_backward_pad_folded: bad fold mode");
  return in_backward_impl(da_pad1, true, fold_mode, y, scale, shift, dy_pad, scratch, B, H, W, C, pad, act, dtype,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int ducosy_pack_dgrad_s2_weight(const float* w_oihw, void* packed, int Cout, int Cin, int ksize, int dtype,
                                           ducosy_stream_t stream) {
  DUCOSY_CHECK(w_oihw && packed && Cout > 0 && Cin > 0 && (ksize == 3 || ksize == 4), DUCOSY_ERR_ARG, "pack_dgrad_s2_weight: bad argument");
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pack_dgrad_s2_weight_kernel<T>, grid_for_items(16LL * Cout * Cin, 256), 256, 0,
                                                                   (cudaStream_t)stream)(w_oihw, static_cast<T*>(packed), Cout, Cin, ksize)));
  return check_launch("pack_dgrad_s2_weight_kernel");
}

// Input gradient of Conv2d(Cin, Cout, 4 or 3, stride 2, padding 1): dy_pad [B][Ho+2][Wo+2][Cout] (zero border) -> dx [B][2Ho][2Wo][Cin].
extern "C" int ducosy_convs2_dgrad_nhwc(const void* dy_pad, const void* w_dgrad, void* dx, int B, int Ho, int Wo, int Cin,
                                           int Cout, int dtype, ducosy_stream_t stream) {
  // four 2x2 phase convs over the padded gradient: same launch as the x2 up-conv with (Cin, Cout) swapped
  return ducosy_upconv2x_nhwc(dy_pad, w_dgrad, dx, nullptr, B, Ho, Wo, Cout, Cin, dtype, stream);
}

extern "C" int ducosy_grad_scale(const float* g, long long n, float* gs, ducosy_stream_t stream) {
  DUCOSY_CHECK(g && gs && n > 0, DUCOSY_ERR_ARG, "grad_scale: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DUCOSY_CHECK(cudaMemsetAsync(gs, 0, 2 * sizeof(float), st) == cudaSuccess, DUCOSY_ERR_CUDA, "grad_scale: memset failed");
  const int blocks = int(std::min<long long>((n + 255) / 256, 148 * 4));
  pdl(grad_absmax_kernel, blocks, 256, 0, st)(g, n, reinterpret_cast<int*>(gs) + 1);
  DUCOSY_TRY(check_launch("grad_absmax_kernel"));
  pdl(grad_scale_finalize_kernel, 1, 1, 0, st)(gs);
  return check_launch("grad_scale_finalize_kernel");
}

extern "C" int ducosy_unpack_wgrad(const float* packed, float* g_oihw, int Cout, int Cin, int taps, const float* gs,
                                   ducosy_stream_t stream) {
  DUCOSY_CHECK(packed && g_oihw, DUCOSY_ERR_ARG, "unpack_wgrad: null pointer");
  pdl(unpack_wgrad_kernel, grid_for_items((long long)Cout * Cin * taps, 256), 256, 0, (cudaStream_t)stream)(packed, g_oihw, Cout, Cin, taps, gs);
  return check_launch("unpack_wgrad_kernel");
}

extern "C" size_t ducosy_disc_last_backward_scratch_bytes(void) { return size_t(kLastSplits) * 8193 * 4; }
extern "C" int ducosy_disc_last_backward(const float* dout, const void* w5_packed, const void* p4, void* da4, float* dw5,
                                         float* db5, float* scratch, const float* gs, int B, int Hs, int Ws, int dtype,
                                         ducosy_stream_t stream) {
  DUCOSY_CHECK(dout && w5_packed && p4 && da4 && dw5 && db5, DUCOSY_ERR_ARG, "disc_last_backward: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(disc_last_dgrad_kernel<T>, grid_for_items((long long)B * Hs * Ws * 64, 256), 256, 0, st)(
                                      dout, static_cast<const T*>(w5_packed), static_cast<T*>(da4), B, Hs, Ws, gs)));
  DUCOSY_TRY(check_launch("disc_last_dgrad_kernel"));
  DUCOSY_CHECK(scratch != nullptr, DUCOSY_ERR_ARG, "disc_last_backward: scratch (ducosy_disc_last_backward_scratch_bytes) is null");
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(disc_last_wgrad_kernel<T>, dim3(16, kLastSplits), 512, 0, st)(dout, static_cast<const T*>(p4), scratch, B, Hs, Ws)));
  DUCOSY_TRY(check_launch("disc_last_wgrad_kernel"));
  pdl(disc_last_wgrad_reduce_kernel, (8193 + 255) / 256, 256, 0, st)(scratch, dw5, db5);
  return check_launch("disc_last_wgrad_reduce_kernel");
}

extern "C" size_t ducosy_disc_first_backward_scratch_bytes(int B, int H, int W) {
  const long long P = (long long)B * (H / 2) * (W / 2);
  return size_t((P + 255) / 256) * 64 * 17 * 4;
}
extern "C" int ducosy_disc_first_backward(const void* da1, const void* p1, const float* x, const float* w1, float* dw1,
                                          float* db1, float* dx, float* scratch, const float* gs, int B, int H, int W,
                                          int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(da1 && p1 && x && w1 && dw1 && db1 && scratch, DUCOSY_ERR_ARG, "disc_first_backward: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long P = (long long)B * (H / 2) * (W / 2);
  const int blocks = int((P + 255) / 256);
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(disc_first_wgrad_kernel<T>, blocks, 256, 0, st)(static_cast<const T*>(da1), static_cast<const T*>(p1),
                                                                                     x, scratch, B, H, W, 256)));
  DUCOSY_TRY(check_launch("disc_first_wgrad_kernel"));
  pdl(disc_first_wgrad_reduce_kernel, (64 * 17 + 255) / 256, 256, 0, st)(scratch, dw1, db1, blocks, gs);
  DUCOSY_TRY(check_launch("disc_first_wgrad_reduce_kernel"));
  if (dx != nullptr) {
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(disc_first_dgrad_kernel<T>, grid_for_items((long long)B * H * W, 256), 256, 0, st)(
                                        static_cast<const T*>(da1), static_cast<const T*>(p1), w1, dx, B, H, W, gs)));
    return check_launch("disc_first_dgrad_kernel");
  }
  return 0;
}

extern "C" int ducosy_pack_dgrad_s1_weight(const float* w_oihw, void* packed, int Cout, int Cin, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(w_oihw && packed && Cout > 0 && Cin > 0, DUCOSY_ERR_ARG, "pack_dgrad_s1_weight: bad argument");
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pack_dgrad_s1_weight_kernel<T>, grid_for_items(9LL * Cout * Cin, 256), 256, 0,
                                                                   (cudaStream_t)stream)(w_oihw, static_cast<T*>(packed), Cout, Cin)));
  return check_launch("pack_dgrad_s1_weight_kernel");
}

// Gradient w.r.t. the PADDED input of a 3x3 stride-1 conv: dy_pad2 [B][H+4][W+4][Cout] (zero border 2) ->
// dxpad [B][H+2][W+2][Cin].  Columns 1..W on the tensor cores (implicit GEMM over the H+2 rows), the two outer columns
// on CUDA cores.  Follow with ducosy_pad_fold for the gradient of the un-padded map.
extern "C" int ducosy_conv3x3s1_dgrad_nhwc(const void* dy_pad2, const void* w_dgrad, void* dxpad, int B, int H, int W, int Cin,
                                           int Cout, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(dy_pad2 && w_dgrad && dxpad && B > 0, DUCOSY_ERR_ARG, "conv3x3s1_dgrad: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "conv3x3s1_dgrad: bad dtype");
  DUCOSY_CHECK(Cout % 64 == 0 && Cin % 64 == 0, DUCOSY_ERR_SHAPE, "conv3x3s1_dgrad: channels must be multiples of 64");
  DUCOSY_CHECK(W % 128 == 0 || (W == 64 && H % 2 == 0), DUCOSY_ERR_SHAPE,
               "conv3x3s1_dgrad: W must be a multiple of 128, or 64 with an even H (the H+2 padded rows are tiled in 128-pixel tiles)");
  DUCOSY_TRY(ducosy_check_device());
  ConvPlan p{};
  p.in = dy_pad2; p.B = B; p.Hp = H + 4; p.Wp = W + 4; p.Cin = Cout; p.stride = 1;
  p.w = w_dgrad; p.Cout = Cin; p.num_phases = 1; p.num_taps = 9;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      p.tap_dy[0][r * 3 + s] = int8_t(2 - r);   // dxpad[u][v'+1] reads dy_pad2[u + 2 - r][v' + 3 - s]
      p.tap_dx[0][r * 3 + s] = int8_t(3 - s);
    }
  p.Hg = H + 2; p.Wg = W; p.out = dxpad; p.Ho = H + 2; p.Wo = W + 2; p.oy_mul = p.ox_mul = 1;
  p.out_y_off = 0; p.out_x_off = 1;
  p.partials = nullptr; p.dtype = dtype;
  DUCOSY_TRY(launch_conv_gemm(p, static_cast<cudaStream_t>(stream)));
  if (Cout == 256 && Cin == 256) {
    // enough CTAs for ~2 per SM: the row tiles of a column are split over CTAs only when the batch alone does not give that
    const int tiles = (H + 2 + 15) / 16, sms = num_sms() > 0 ? num_sms() : 148;
    const int row_splits = std::max(1, std::min(tiles, (2 * sms + 16 * B - 1) / (16 * B)));
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(dgrad_s1_edge_cols_mma_kernel<T>, dim3(8 * row_splits, 2, B), 128, 0, (cudaStream_t)stream)(
                                        static_cast<const T*>(dy_pad2), static_cast<const T*>(w_dgrad), static_cast<T*>(dxpad), B, H, W, row_splits)));
    return check_launch("dgrad_s1_edge_cols_mma_kernel");
  }
  if (Cout == 256 && Cin % 8 == 0) {
    const int cps = std::max(1, std::min(H + 2, (148 * 16) / std::max(1, Cin / 8 * 2 * B)));
    const int per = (H + 2 + cps - 1) / cps, chunks = (H + 2 + per - 1) / per;
    const size_t smem = size_t(per + 2) * 512;
    DUCOSY_CHECK(smem <= 48 * 1024, DUCOSY_ERR_SHAPE, "conv3x3s1_dgrad: H too large for the edge-column kernel");
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(dgrad_s1_edge_cols256_kernel<T>, dim3(Cin / 8, 2, B * chunks), 256, smem, (cudaStream_t)stream)(
                                        static_cast<const T*>(dy_pad2), static_cast<const T*>(w_dgrad), static_cast<T*>(dxpad), B, H, W,
                                        Cin, per, chunks)));
    return check_launch("dgrad_s1_edge_cols256_kernel");
  }
  const long long total = (long long)B * (H + 2) * 2 * Cin;
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(dgrad_s1_edge_cols_kernel<T>, grid_for_items(total, 128), 128, 0, (cudaStream_t)stream)(
                                      static_cast<const T*>(dy_pad2), static_cast<const T*>(w_dgrad), static_cast<T*>(dxpad), B, H, W,
                                      Cin, Cout)));
  return check_launch("dgrad_s1_edge_cols_kernel");
}

// Adjoint of ReflectionPad2d(pad) / zero padding: dxpad [B][H+2p][W+2p][C] -> dx [B][H][W][C] (16-bit NHWC).
extern "C" int ducosy_pad_fold_add(const void* dxpad, const void* add, void* dx, int B, int H, int W, int C, int pad, int pad_mode,
                                   int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(dxpad && dx && B > 0 && C % 8 == 0 && pad >= 0 && pad < H && pad < W, DUCOSY_ERR_ARG, "pad_fold: bad argument");
  const long long total = (long long)B * H * W * (C / 8);
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pad_fold_kernel<T>, grid_for_items(total, 256), 256, 0, (cudaStream_t)stream)(
                                      static_cast<const T*>(dxpad), static_cast<const T*>(add), static_cast<T*>(dx), B, H, W, C, pad,
                                      pad_mode)));
  return check_launch("pad_fold_kernel");
}
extern "C" int ducosy_pad_fold(const void* dxpad, void* dx, int B, int H, int W, int C, int pad, int pad_mode, int dtype,
                               ducosy_stream_t stream) {
  return ducosy_pad_fold_add(dxpad, nullptr, dx, B, H, W, C, pad, pad_mode, dtype, stream);
}

extern "C" int ducosy_pack_upconv_dgrad_weight(const float* w_oihw, void* packed, int Cout, int Cin, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(w_oihw && packed && Cout > 0 && Cin > 0, DUCOSY_ERR_ARG, "pack_upconv_dgrad_weight: bad argument");
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pack_upconv_dgrad_weight_kernel<T>, grid_for_items(16LL * Cout * Cin, 256), 256, 0,
                                                                       (cudaStream_t)stream)(w_oihw, static_cast<T*>(packed), Cout, Cin)));
  return check_launch("pack_upconv_dgrad_weight_kernel");
}

// Input gradient of Upsample(x2)+Conv3x3 (modules/model.py:108-109): dy_pad2 [B][2Hs+4][2Ws+4][Cout] (zero border 2) ->
// dsrc [B][Hs][Ws][Cin]; one implicit GEMM over 16 (phase, tap) gathers with stride-2 TMA boxes.
extern "C" int ducosy_upconv2x_dgrad_nhwc(const void* dy_pad2, const void* w_dgrad, void* dsrc, int B, int Hs, int Ws, int Cin,
                                          int Cout, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(dy_pad2 && w_dgrad && dsrc && B > 0, DUCOSY_ERR_ARG, "upconv2x_dgrad: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "upconv2x_dgrad: bad dtype");
  DUCOSY_TRY(ducosy_check_device());
  ConvPlan p{};
  p.in = dy_pad2; p.B = B; p.Hp = 2 * Hs + 4; p.Wp = 2 * Ws + 4; p.Cin = Cout; p.stride = 2;
  p.w = w_dgrad; p.Cout = Cin; p.num_phases = 1; p.num_taps = 16;
  for (int t16 = 0; t16 < 16; ++t16) {
    const int phase = t16 >> 2, a = (t16 >> 1) & 1, b = t16 & 1, py = phase >> 1, px = phase & 1;
    p.tap_dy[0][t16] = int8_t(2 * (2 - a - py) + py);   // padded-by-2 pixel row of dY[2(i+1-a-py)+py] relative to 2i
    p.tap_dx[0][t16] = int8_t(2 * (2 - b - px) + px);
  }
  p.Hg = Hs; p.Wg = Ws; p.out = dsrc; p.Ho = Hs; p.Wo = Ws; p.oy_mul = p.ox_mul = 1;
  p.partials = nullptr; p.dtype = dtype;
  return launch_conv_gemm(p, static_cast<cudaStream_t>(stream));
}

// up_pad [B][2Hs+2][2Ws+2][C] = zero-pad-1(nearest-x2(interior of src_pad [B][Hs+2][Ws+2][C])): x operand of the weight
// gradient of Upsample(x2)+Conv3x3 (then ducosy_conv2d_wgrad_nhwc(up_pad, dy, ..., 3, 3, 1)).
extern "C" int ducosy_upsample2x_pad(const void* src_pad, void* up_pad, int B, int Hs, int Ws, int C, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(src_pad && up_pad && B > 0 && C % 8 == 0, DUCOSY_ERR_ARG, "upsample2x_pad: bad argument");
  const long long total = (long long)B * (2 * Hs + 2) * (2 * Ws + 2) * (C / 8);
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(upsample2x_pad_kernel<T>, grid_for_items(total, 256), 256, 0, (cudaStream_t)stream)(
                                      static_cast<const T*>(src_pad), static_cast<T*>(up_pad), B, Hs, Ws, C)));
  return check_launch("upsample2x_pad_kernel");
}
