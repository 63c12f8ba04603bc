// Stand-alone building blocks of modules/model.py on the reference's own tensor layout (NCHW fp32 in, NCHW fp32 out):
//   ChannelAttention  modules/model.py:6-24   x * sigmoid(fc(avgpool x) + fc(maxpool x))     -- BOTH pooling branches
//   SpatialAttention  modules/model.py:27-39  x * sigmoid(conv_kxk(cat[mean_C x, max_C x]))
// plus the two layout converters that let a stand-alone ResidualBlock / ResidualBlockWithCBAM (modules/model.py:56-87)
// run on the NHWC 16-bit tensor-core path of the generator:
//   nchw_to_nhwc_pad  fp32 [B,C,H,W] -> 16-bit [B,H+2p,W+2p,C] with reflect / zero padding
//   nhwc_to_nchw      16-bit [B,H,W,C] -> fp32 [B,C,H,W]
// Inside Generator.forward the same math runs fused (the avg branch of the channel attention is identically zero behind a
// non-affine InstanceNorm and is dropped there); these kernels serve callers that use the blocks on their own, e.g.
// BASELINE config 5 `ResidualBlockWithCBAM(256)(randn(B,256,128,128))`.  All bandwidth-bound, fp32 arithmetic.
#include "common.cuh"

namespace ducosy {
namespace {

__device__ __forceinline__ int reflect_i(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- ChannelAttention ------------------------------------------------------------------------------------------------
// one CTA per (b, c) plane: mean and max over H*W (fixed-order tree: deterministic)
__global__ void __launch_bounds__(256)
ca_pool_kernel(const float* __restrict__ x, float* __restrict__ avg, float* __restrict__ mx, long long HW) {
  pdl_prologue();
  const float* p = x + (long long)blockIdx.x * HW;
  float s = 0.f, m = -INFINITY;
  if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (long long i = threadIdx.x; i < HW / 4; i += 256) {
      const float4 v = p4[i];
      s += (v.x + v.y) + (v.z + v.w);
      m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
  } else {
    for (long long i = threadIdx.x; i < HW; i += 256) {
      const float v = p[i];
      s += v;
      m = fmaxf(m, v);
    }
  }
  __shared__ float ss[8], sm[8];
  s = warp_sum(s);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) {
    ss[threadIdx.x >> 5] = s;
    sm[threadIdx.x >> 5] = m;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = -INFINITY;
    for (int i = 0; i < 8; ++i) {
      a += ss[i];
      b = fmaxf(b, sm[i]);
    }
    avg[blockIdx.x] = a / float(HW);
    mx[blockIdx.x] = b;
  }
}

// one CTA per sample: att[c] = sigmoid(fc2 . relu(fc0 . avg) + fc2 . relu(fc0 . max));  fc0 [Hd][C], fc2 [C][Hd]
__global__ void __launch_bounds__(256)
ca_mlp_kernel(const float* __restrict__ avg, const float* __restrict__ mx, const float* __restrict__ fc0,
              const float* __restrict__ fc2, float* __restrict__ att, int C, int Hd) {
  pdl_prologue();
  extern __shared__ float sh[];          // hidden_avg[Hd], hidden_max[Hd]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < Hd; h += 8) {
    float sa = 0.f, sm = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float w = fc0[(long long)h * C + c];
      sa = fmaf(w, avg[(long long)b * C + c], sa);
      sm = fmaf(w, mx[(long long)b * C + c], sm);
    }
    sa = warp_sum(sa);
    sm = warp_sum(sm);
    if (lane == 0) {
      sh[h] = fmaxf(sa, 0.f);
      sh[Hd + h] = fmaxf(sm, 0.f);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float o = 0.f;
    for (int h = 0; h < Hd; ++h) o = fmaf(fc2[(long long)c * Hd + h], sh[h] + sh[Hd + h], o);
    att[(long long)b * C + c] = 1.f / (1.f + __expf(-o));
  }
}

// out[b,c,:] = x[b,c,:] * att[b,c]
__global__ void __launch_bounds__(256)
scale_planes_kernel(const float* __restrict__ x, const float* __restrict__ att, float* __restrict__ out, long long HW,
                    long long total) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256)
    out[i] = x[i] * att[i / HW];
}

// ---- SpatialAttention ------------------------------------------------------------------------------------------------
// pooled[b][0][p] = mean_c x, pooled[b][1][p] = max_c x  (thread per pixel, coalesced along W for every channel plane)
__global__ void __launch_bounds__(256)
sa_pool_kernel(const float* __restrict__ x, float* __restrict__ pooled, int C, long long HW, long long total) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long b = i / HW, p = i - b * HW;
    const float* src = x + b * C * HW + p;
    float s = 0.f, m = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const float v = src[(long long)c * HW];
      s += v;
      m = fmaxf(m, v);
    }
    pooled[(b * 2 + 0) * HW + p] = s / float(C);
    pooled[(b * 2 + 1) * HW + p] = m;
  }
}

// att[b][y][x] = sigmoid(sum_{ch,dy,dx} w[ch][dy][dx] * pooled[b][ch][y+dy-k/2][x+dx-k/2]), zero padding, no bias
__global__ void __launch_bounds__(256)
sa_conv_kernel(const float* __restrict__ pooled, const float* __restrict__ w, float* __restrict__ att, int H, int W, int k,
               long long total) {
  pdl_prologue();
  extern __shared__ float wsh[];
  for (int i = threadIdx.x; i < 2 * k * k; i += 256) wsh[i] = w[i];
  __syncthreads();
  const int r = k / 2;
  const long long HW = (long long)H * W;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long b = i / HW;
    const int p = int(i - b * HW), y = p / W, x = p - y * W;
    float acc = 0.f;
    for (int ch = 0; ch < 2; ++ch) {
      const float* src = pooled + (b * 2 + ch) * HW;
      for (int dy = 0; dy < k; ++dy) {
        const int yy = y + dy - r;
        if (yy < 0 || yy >= H) continue;
        for (int dx = 0; dx < k; ++dx) {
          const int xx = x + dx - r;
          if (xx < 0 || xx >= W) continue;
          acc = fmaf(wsh[(ch * k + dy) * k + dx], src[(long long)yy * W + xx], acc);
        }
      }
    }
    att[i] = 1.f / (1.f + __expf(-acc));
  }
}

// out[b,c,p] = x[b,c,p] * att[b,p]
__global__ void __launch_bounds__(256)
scale_pixels_kernel(const float* __restrict__ x, const float* __restrict__ att, float* __restrict__ out, int C, long long HW,
                    long long total) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long bc = i / HW, p = i - bc * HW;
    out[i] = x[i] * att[(bc / C) * HW + p];
  }
}

// ---- layout converters -----------------------------------------------------------------------------------------------
// grid (ceil(Wp/32), Hp, B), block (32, 8): a 32-pixel x 32-channel tile goes through shared memory so that both the NCHW
// reads (along W) and the NHWC writes (along C) are coalesced
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_pad_kernel(const float* __restrict__ x, T* __restrict__ out, int C, int H, int W, int pad, int pad_mode) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int b = blockIdx.z, yp = blockIdx.y, xp0 = blockIdx.x * 32;
  int sy = yp - pad;
  const bool row_in = sy >= 0 && sy < H;
  sy = reflect_i(sy, H);
  const int xp = xp0 + threadIdx.x;
  int sx = xp - pad;
  const bool col_in = sx >= 0 && sx < W;
  sx = reflect_i(sx, W);
  const bool live = xp < Wp && ((row_in && col_in) || pad_mode == DUCOSY_PAD_REFLECT);
  for (int c0 = 0; c0 < C; c0 += 32) {
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int c = c0 + j;
      tile[j][threadIdx.x] = (live && c < C) ? x[(((long long)b * C + c) * H + sy) * W + sx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int c = c0 + threadIdx.x;
      if (xp0 + i < Wp && c < C)
        out[(((long long)b * Hp + yp) * Wp + xp0 + i) * C + c] = Cvt<T>::from_f(tile[threadIdx.x][i]);
    }
    __syncthreads();
  }
  (void)Hp;
}

template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T* __restrict__ y, float* __restrict__ out, int C, int H, int W) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, yy = blockIdx.y, x0 = blockIdx.x * 32;
  for (int c0 = 0; c0 < C; c0 += 32) {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int c = c0 + threadIdx.x;
      tile[i][threadIdx.x] = (x0 + i < W && c < C) ? Cvt<T>::to_f(y[(((long long)b * H + yy) * W + x0 + i) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int c = c0 + j;
      if (x0 + threadIdx.x < W && c < C) out[(((long long)b * C + c) * H + yy) * W + x0 + threadIdx.x] = tile[threadIdx.x][j];
    }
    __syncthreads();
  }
}

int ew_blocks(long long total) {
  const long long cap = (long long)(num_sms() > 0 ? num_sms() : 148) * 8;
  const long long n = (total + 255) / 256;
  return int(n < cap ? (n > 0 ? n : 1) : cap);
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" size_t ducosy_channel_attention_scratch_bytes(int B, int C) { return size_t(B) * C * 3 * sizeof(float); }

extern "C" int ducosy_channel_attention_nchw(const float* x, const float* fc0, const float* fc2, float* out, float* scratch,
                                             int B, int C, int hidden, int H, int W, ducosy_stream_t stream) {
  DUCOSY_CHECK(x && fc0 && fc2 && out && scratch, DUCOSY_ERR_ARG, "channel_attention_nchw: null pointer");
  DUCOSY_CHECK(B > 0 && C > 0 && hidden > 0 && hidden <= 4096 && H > 0 && W > 0, DUCOSY_ERR_SHAPE, "channel_attention_nchw: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long HW = (long long)H * W;
  float *avg = scratch, *mx = scratch + (size_t)B * C, *att = scratch + (size_t)2 * B * C;
  pdl(ca_pool_kernel, B * C, 256, 0, st)(x, avg, mx, HW);
  DUCOSY_TRY(check_launch("ca_pool_kernel"));
  pdl(ca_mlp_kernel, B, 256, 2 * hidden * sizeof(float), st)(avg, mx, fc0, fc2, att, C, hidden);
  DUCOSY_TRY(check_launch("ca_mlp_kernel"));
  const long long total = (long long)B * C * HW;
  pdl(scale_planes_kernel, ew_blocks(total), 256, 0, st)(x, att, out, HW, total);
  return check_launch("scale_planes_kernel");
}

extern "C" size_t ducosy_spatial_attention_scratch_bytes(int B, int H, int W) { return size_t(B) * H * W * 3 * sizeof(float); }

extern "C" int ducosy_spatial_attention_nchw(const float* x, const float* w, float* out, float* scratch, int B, int C, int H,
                                             int W, int ksize, ducosy_stream_t stream) {
  DUCOSY_CHECK(x && w && out && scratch, DUCOSY_ERR_ARG, "spatial_attention_nchw: null pointer");
  DUCOSY_CHECK(B > 0 && C > 0 && H > 0 && W > 0 && ksize >= 1 && ksize <= 31 && (ksize & 1), DUCOSY_ERR_SHAPE,
               "spatial_attention_nchw: bad shape (odd kernel size up to 31)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long HW = (long long)H * W, npix = (long long)B * HW;
  float *pooled = scratch, *att = scratch + (size_t)2 * B * HW;
  pdl(sa_pool_kernel, ew_blocks(npix), 256, 0, st)(x, pooled, C, HW, npix);
  DUCOSY_TRY(check_launch("sa_pool_kernel"));
  pdl(sa_conv_kernel, ew_blocks(npix), 256, 2 * ksize * ksize * sizeof(float), st)(pooled, w, att, H, W, ksize, npix);
  DUCOSY_TRY(check_launch("sa_conv_kernel"));
  const long long total = npix * C;
  pdl(scale_pixels_kernel, ew_blocks(total), 256, 0, st)(x, att, out, C, HW, total);
  return check_launch("scale_pixels_kernel");
}

extern "C" int ducosy_nchw_to_nhwc_pad(const float* x, void* out, int B, int C, int H, int W, int pad, int pad_mode, int dtype,
                                       ducosy_stream_t stream) {
  DUCOSY_CHECK(x && out && B > 0 && C > 0 && H > 0 && W > 0, DUCOSY_ERR_ARG, "nchw_to_nhwc_pad: bad argument");
  DUCOSY_CHECK(pad >= 0 && pad < H && pad < W && B <= 65535 && H + 2 * pad <= 65535, DUCOSY_ERR_SHAPE, "nchw_to_nhwc_pad: bad shape");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "nchw_to_nhwc_pad: bad dtype");
  const dim3 grid((W + 2 * pad + 31) / 32, H + 2 * pad, B), block(32, 8);
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(nchw_to_nhwc_pad_kernel<T>, grid, block, 0, (cudaStream_t)stream)(
                                      x, static_cast<T*>(out), C, H, W, pad, pad_mode)));
  return check_launch("nchw_to_nhwc_pad_kernel");
}

extern "C" int ducosy_nhwc_to_nchw(const void* y, float* out, int B, int C, int H, int W, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(y && out && B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535 && H <= 65535, DUCOSY_ERR_ARG, "nhwc_to_nchw: bad argument");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "nhwc_to_nchw: bad dtype");
  const dim3 grid((W + 31) / 32, H, B), block(32, 8);
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(nhwc_to_nchw_kernel<T>, grid, block, 0, (cudaStream_t)stream)(
                                      static_cast<const T*>(y), out, C, H, W)));
  return check_launch("nhwc_to_nchw_kernel");
}
