// Bandwidth-bound kernels around the tensor-core convolutions: weight packing, stem im2col (optionally fused
// with the HU window), InstanceNorm finalize/apply (+ReLU, + padding writer), CBAM channel MLP, CBAM spatial
// pooling / 7x7 attention conv, and the residual add.  NHWC 16-bit activations, 128-bit accesses.
#include <cstdlib>

#include "common.cuh"
#include "input_fn.cuh"

namespace ducosy {
namespace {

int ew_grid(long long work_items, int threads) {
  const int sms = num_sms();
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)(sms > 0 ? sms : 148) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}

// ------------------------------------------------------------------ 8-channel chunks (one 128-bit access)
template <typename T>
__device__ __forceinline__ void unpack8(uint4 raw, float (&f)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = Cvt<T>::unpack2(w[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(Cvt<T>::pack2(f[0], f[1]), Cvt<T>::pack2(f[2], f[3]), Cvt<T>::pack2(f[4], f[5]), Cvt<T>::pack2(f[6], f[7]));
}
// split-operand mode: hi = rn(v), lo = rn(v - hi)
template <typename T>
__device__ __forceinline__ void split8(const float (&f)[8], uint4& hi, uint4& lo) {
  hi = pack8<T>(f);
  float h[8], r[8];
  unpack8<T>(hi, h);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = f[i] - h[i];
  lo = pack8<T>(r);
}
// value of a (hi, lo) chunk pair
template <typename T>
__device__ __forceinline__ void join8(uint4 hi, uint4 lo, float (&f)[8]) {
  float l[8];
  unpack8<T>(hi, f);
  unpack8<T>(lo, l);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] += l[i];
}

// ------------------------------------------------------------------ weight packing
// Element i of a packed [rows][K] matrix.  planes == 2 (split-operand mode, DUCOSY_F16X2): the row is [K hi | K lo] with
// hi = rn(v), lo = rn(v - hi).
template <typename T>
__device__ __forceinline__ void store_packed(T* __restrict__ out, long long i, long long K, int planes, float v) {
  const T hi = Cvt<T>::from_f(v);
  if (planes == 1) {
    out[i] = hi;
  } else {
    const long long row = i / K, k = i - row * K;
    out[row * 2 * K + k] = hi;
    out[row * 2 * K + K + k] = Cvt<T>::from_f(v - Cvt<T>::to_f(hi));
  }
}

template <typename T>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int Cout, int Cin, int kh,
                                        int kw, int planes) {
  pdl_prologue();
  const long long total = (long long)Cout * Cin * kh * kw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % Cin);
    long long r = i / Cin;
    const int tap = int(r % (kh * kw));
    const int o = int(r / (kh * kw));
    store_packed(out, i, (long long)Cin * kh * kw, planes, w[((long long)o * Cin + c) * kh * kw + tap]);
  }
}

// Upsample(x2 nearest) + 3x3/pad1  ==  four 2x2 convs on the source grid:
// rows: phase 0 -> [w0, w1+w2], phase 1 -> [w0+w1, w2]; same for columns.  Sums are formed in fp32.
template <typename T>
__global__ void pack_upconv_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int Cout, int Cin, int planes) {
  pdl_prologue();
  const long long total = 4LL * Cout * 4 * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % Cin);
    long long r = i / Cin;
    const int tap = int(r % 4);
    r /= 4;
    const int o = int(r % Cout);
    const int phase = int(r / Cout);
    const int py = phase >> 1, px = phase & 1, a = tap >> 1, b = tap & 1;
    // source-row tap a of output-row phase py gathers original rows [r0, r1]
    const int r0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
    const int r1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
    const int s0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2);
    const int s1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
    const float* wk = w + ((long long)o * Cin + c) * 9;
    float acc = 0.f;
    for (int rr = r0; rr <= r1; ++rr)
      for (int ss = s0; ss <= s1; ++ss) acc += wk[rr * 3 + ss];
    store_packed(out, i, 4LL * Cin, planes, acc);
  }
}

// Merged-phase variant for Cout = 64: ONE GEMM with N = 4*Cout columns (column block f = output phase) over the nine
// source offsets (dy, dx) in {0,1,2}^2; entries of offsets a phase does not touch are zero.  2.25x the MACs of the
// four 2x2 convs, but A tiles are loaded once for all phases and the tile is N = 256 wide.
template <typename T>
__global__ void pack_upconv_merged_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int Cout, int Cin) {
  pdl_prologue();
  const long long total = 4LL * Cout * 9 * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % Cin);
    long long r = i / Cin;
    const int tap = int(r % 9);
    r /= 9;
    const int o = int(r % Cout);
    const int f = int(r / Cout);
    const int py = f >> 1, px = f & 1, a = tap / 3 - py, b = tap % 3 - px;
    float acc = 0.f;
    if (a >= 0 && a <= 1 && b >= 0 && b <= 1) {
      const int r0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
      const int r1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
      const int s0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2);
      const int s1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
      const float* wk = w + ((long long)o * Cin + c) * 9;
      for (int rr = r0; rr <= r1; ++rr)
        for (int ss = s0; ss <= s1; ++ss) acc += wk[rr * 3 + ss];
    }
    out[i] = Cvt<T>::from_f(acc);
  }
}

template <typename T>
__global__ void pack_stem_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int Cin, int Kpad, int planes) {
  pdl_prologue();
  const int total = 64 * Kpad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % Kpad, o = i / Kpad;
    store_packed(out, i, Kpad, planes, k < 49 * Cin ? w[o * 49 * Cin + k] : 0.f);
  }
}

// ------------------------------------------------------------------ stem im2col
// A[(b,y,x)][k], k = c*49 + r*7 + s  <-  in[b][c][reflect(y+r-3)][reflect(x+s-3)], zero for k >= 49*Cin.
// One thread builds 8 consecutive k (one 16-byte store); consecutive threads -> consecutive chunks of a row.
template <typename T, typename In, bool kSplit>
__global__ void stem_im2col_kernel(In in, T* __restrict__ A, int B, int Cin, int H, int W, int Kpad) {
  pdl_prologue();
  const int chunks = Kpad / 8;
  const long long total = (long long)B * H * W * chunks;
  const int K = 49 * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ck = int(i % chunks);
    long long pix = i / chunks;
    const int x = int(pix % W);
    pix /= W;
    const int y = int(pix % H);
    const int b = int(pix / H);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = ck * 8 + j;
      if (k < K) {
        const int c = k / 49, rs = k - c * 49, r = rs / 7, s = rs - r * 7;
        v[j] = in.at(b, c, reflect_idx(y + r - 3, H), reflect_idx(x + s - 3, W), Cin, H, W);
      } else {
        v[j] = 0.f;
      }
    }
    if (kSplit) {   // row = [Kpad hi | Kpad lo]
      uint4 hi, lo;
      split8<T>(v, hi, lo);
      uint4* rowp = reinterpret_cast<uint4*>(A) + (i / chunks) * (2 * chunks);
      rowp[ck] = hi;
      rowp[chunks + ck] = lo;
    } else {
      reinterpret_cast<uint4*>(A)[i] = pack8<T>(v);
    }
  }
}

// Row-staged variant for small Cin: a CTA owns one image row, stages the 7 reflected input rows (already converted
// to T) in shared memory, then every thread assembles whole 128-byte K-rows with compile-time (r, s) offsets.
template <typename T, typename In, int CIN>
__global__ void __launch_bounds__(256)
stem_im2col_rows_kernel(In in, T* __restrict__ A, int H, int W) {
  pdl_prologue();
  constexpr int KPAD = (49 * CIN + 63) / 64 * 64;
  extern __shared__ __align__(16) uint8_t im2col_smem[];
  T* tile = reinterpret_cast<T*>(im2col_smem);  // [CIN][7][W + 6]
  const int y = blockIdx.x, b = blockIdx.y, Wp = W + 6;
  for (int i = threadIdx.x; i < CIN * 7 * Wp; i += blockDim.x) {
    const int xp = i % Wp, r = (i / Wp) % 7, c = i / (7 * Wp);
    tile[i] = Cvt<T>::from_f(in.at(b, c, reflect_idx(y + r - 3, H), reflect_idx(xp - 3, W), CIN, H, W));
  }
  __syncthreads();
  const unsigned short* t16 = reinterpret_cast<const unsigned short*>(tile);
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    uint4* dst = reinterpret_cast<uint4*>(A + (((long long)b * H + y) * W + x) * KPAD);
#pragma unroll
    for (int ck = 0; ck < KPAD / 8; ++ck) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t lo = 0, hi = 0;
        const int k0 = ck * 8 + 2 * j, k1 = k0 + 1;
        if (k0 < 49 * CIN) lo = t16[((k0 / 49) * 7 + (k0 % 49) / 7) * Wp + x + (k0 % 7)];
        if (k1 < 49 * CIN) hi = t16[((k1 / 49) * 7 + (k1 % 49) / 7) * Wp + x + (k1 % 7)];
        w[j] = lo | (hi << 16);
      }
      dst[ck] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

template <typename T, typename In>
int launch_im2col(In in, T* A, int B, int Cin, int H, int W, bool split, cudaStream_t st) {
  const size_t smem = size_t(Cin) * 7 * (W + 6) * sizeof(T);
  const dim3 grid(H, B);
  if (split) {
    const int Kpad = (49 * Cin + 63) / 64 * 64;
    const long long total = (long long)B * H * W * (Kpad / 8);
    pdl(stem_im2col_kernel<T, In, true>, ew_grid(total, 256), 256, 0, st)(in, A, B, Cin, H, W, Kpad);
  } else if (Cin <= 3 && smem <= 48 * 1024) {
    if (Cin == 1) pdl(stem_im2col_rows_kernel<T, In, 1>, grid, 256, smem, st)(in, A, H, W);
    else if (Cin == 2) pdl(stem_im2col_rows_kernel<T, In, 2>, grid, 256, smem, st)(in, A, H, W);
    else pdl(stem_im2col_rows_kernel<T, In, 3>, grid, 256, smem, st)(in, A, H, W);
  } else {
    const int Kpad = (49 * Cin + 63) / 64 * 64;
    const long long total = (long long)B * H * W * (Kpad / 8);
    pdl(stem_im2col_kernel<T, In, false>, ew_grid(total, 256), 256, 0, st)(in, A, B, Cin, H, W, Kpad);
  }
  return check_launch("stem_im2col_kernel");
}

// ------------------------------------------------------------------ InstanceNorm finalize (+ CBAM channel MLP)
// grid (C/32, B), 32 warps: warp w reduces tiles w, w+32, ... for 32 consecutive channels (coalesced 128-byte rows,
// 4 tiles = 12 loads in flight), partial sums in double, combined through shared memory in a fixed order
// (deterministic).  var is the biased variance, eps = 1e-5.
constexpr int kFinWarps = 32;
__global__ void __launch_bounds__(kFinWarps * 32)
in_finalize_kernel(const float* __restrict__ partials, int tiles, int npix, float* __restrict__ scale,
                   float* __restrict__ shift, float* __restrict__ chmax, int C) {
  pdl_prologue();
  __shared__ double s1s[kFinWarps][32], s2s[kFinWarps][32];
  __shared__ float mxs[kFinWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane, b = blockIdx.y;
  const float* p = partials + (long long)b * tiles * 3 * C + c;
  double s1 = 0.0, s2 = 0.0;
  float mx = -INFINITY;
  for (int t0 = warp; t0 < tiles; t0 += kFinWarps * 4) {
    float a[4], q[4], m[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * kFinWarps;
      const bool ok = t < tiles;
      a[u] = ok ? p[(long long)(t * 3 + 0) * C] : 0.f;
      q[u] = ok ? p[(long long)(t * 3 + 1) * C] : 0.f;
      m[u] = ok ? p[(long long)(t * 3 + 2) * C] : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s1 += double(a[u]);
      s2 += double(q[u]);
      mx = fmaxf(mx, m[u]);
    }
  }
  s1s[warp][lane] = s1;
  s2s[warp][lane] = s2;
  mxs[warp][lane] = mx;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int w = 1; w < kFinWarps; ++w) {
      s1 += s1s[w][lane];
      s2 += s2s[w][lane];
      mx = fmaxf(mx, mxs[w][lane]);
    }
    const double mean = s1 / npix;
    double var = s2 / npix - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = float(1.0 / sqrt(var + 1e-5));
    const float fmean = float(mean);
    scale[b * C + c] = rstd;
    shift[b * C + c] = -fmean * rstd;
    if (chmax != nullptr) chmax[b * C + c] = (mx - fmean) * rstd;  // max over H*W of the normalised map (rstd > 0)
  }
}

// CBAM channel attention (modules/model.py:20-24), one CTA per sample: s = sigmoid(fc(avgpool) + fc(maxpool)) folded
// into the InstanceNorm affine.  The avg-pool branch sees the mean of a non-affine InstanceNorm output, which is
// exactly zero here (statistics are taken from the stored values), and fc has no bias: fc(0) = 0.
__global__ void __launch_bounds__(256)
cbam_channel_mlp_kernel(const float* __restrict__ chmax, const float* __restrict__ fc0, const float* __restrict__ fc2,
                        float* __restrict__ scale, float* __restrict__ shift, int C) {
  pdl_prologue();
  extern __shared__ float sm[];  // [C] normalised max, [C/16] hidden
  float* smax = sm;
  float* hidden = sm + C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) smax[c] = chmax[b * C + c];
  __syncthreads();
  const int Hd = C / 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < Hd; j += nwarps) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc += fc0[j * C + c] * smax[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) hidden[j] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < Hd; ++j) acc += fc2[c * Hd + j] * hidden[j];
    const float s = 1.f / (1.f + __expf(-acc));
    scale[b * C + c] *= s;
    shift[b * C + c] *= s;
  }
}

// ------------------------------------------------------------------ IN apply (+act) + padding writer
template <typename T>
__device__ __forceinline__ uint4 affine8(uint4 raw, const float* sc, const float* sh, int act) {
  uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = Cvt<T>::unpack2(w[i]);
    f.x = fmaf(f.x, sc[2 * i], sh[2 * i]);
    f.y = fmaf(f.y, sc[2 * i + 1], sh[2 * i + 1]);
    if (act == DUCOSY_ACT_RELU) {
      f.x = fmaxf(f.x, 0.f);
      f.y = fmaxf(f.y, 0.f);
    } else if (act == DUCOSY_ACT_LRELU02) {
      f.x = f.x > 0.f ? f.x : 0.2f * f.x;
      f.y = f.y > 0.f ? f.y : 0.2f * f.y;
    }
    w[i] = Cvt<T>::pack2(f.x, f.y);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Row-persistent layout shared by the two padding writers below: a CTA walks padded output rows (b, py); inside a
// row thread t handles 16-byte chunks t, t+256, ...  Because 256 is a multiple of C/8, a thread always sees the same
// 8 channels, so its scale/shift live in registers; four chunks are in flight per thread (ILP) and the grid is
// 4 CTAs per SM so that a tensor-core conv CTA of the other generator's stream can share the SM.
constexpr int kRowILP = 4;

// split-operand mode: relu(v*scale + shift) of the joined (hi, lo) value, split again
template <typename T>
__device__ __forceinline__ void affine8_split(uint4 rh, uint4 rl, const float* sc, const float* sh, int act, uint4& oh, uint4& ol) {
  float f[8];
  join8<T>(rh, rl, f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[i] = fmaf(f[i], sc[i], sh[i]);
    if (act == DUCOSY_ACT_RELU) f[i] = fmaxf(f[i], 0.f);
    else if (act == DUCOSY_ACT_LRELU02) f[i] = f[i] > 0.f ? f[i] : 0.2f * f[i];
  }
  split8<T>(f, oh, ol);
}

// kSplit (DUCOSY_F16X2): a pixel holds 2*C channels, the lo plane C channels behind the hi plane.
template <typename T, bool kSplit, bool kCap = true>
__global__ void __launch_bounds__(256, (kSplit || !kCap) ? 1 : 4)   // 64 registers: the 4-CTAs-per-SM grid (row_grid) is ONE resident wave
in_apply_pad_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                    T* __restrict__ out, int B, int H, int W, int C, int pad, int pad_mode, int act) {
  pdl_prologue();
  const int Hp = H + 2 * pad, Wp = W + 2 * pad, cv = C / 8, cvp = kSplit ? 2 * cv : cv;
  const int c8 = threadIdx.x % cv, row_chunks = Wp * cvp;
  const int px0 = threadIdx.x / cv, px_step = 256 / cv;
  int cur_b = -1;
  float s8[8], h8[8];
  for (int row = blockIdx.x; row < B * Hp; row += gridDim.x) {
    const int b = row / Hp, py = row - b * Hp;
    if (b != cur_b) {
      cur_b = b;
      const float4* sc = reinterpret_cast<const float4*>(scale + b * C + c8 * 8);
      const float4* sh = reinterpret_cast<const float4*>(shift + b * C + c8 * 8);
      const float4 a0 = __ldg(sc), a1 = __ldg(sc + 1), b0 = __ldg(sh), b1 = __ldg(sh + 1);
      s8[0] = a0.x; s8[1] = a0.y; s8[2] = a0.z; s8[3] = a0.w; s8[4] = a1.x; s8[5] = a1.y; s8[6] = a1.z; s8[7] = a1.w;
      h8[0] = b0.x; h8[1] = b0.y; h8[2] = b0.z; h8[3] = b0.w; h8[4] = b1.x; h8[5] = b1.y; h8[6] = b1.z; h8[7] = b1.w;
    }
    int sy = py - pad;
    const bool row_inside = sy >= 0 && sy < H;
    sy = reflect_idx(sy, H);
    const uint4* src_row = reinterpret_cast<const uint4*>(y) + ((long long)b * H + sy) * W * cvp + c8;
    uint4* dst_row = reinterpret_cast<uint4*>(out) + (long long)row * row_chunks + c8;
    for (int px = px0; px < Wp; px += px_step * kRowILP) {
      uint4 raw[kRowILP], rawl[kSplit ? kRowILP : 1];
      bool live[kRowILP];
#pragma unroll
      for (int u = 0; u < kRowILP; ++u) {
        const int p = px + u * px_step;
        int sx = p - pad;
        const bool inside = row_inside && sx >= 0 && sx < W;
        live[u] = p < Wp && (inside || pad_mode == DUCOSY_PAD_REFLECT);
        sx = reflect_idx(sx, W);
        if (live[u]) {
          raw[u] = src_row[(long long)sx * cvp];
          if (kSplit) rawl[u] = src_row[(long long)sx * cvp + cv];
        }
      }
#pragma unroll
      for (int u = 0; u < kRowILP; ++u) {
        const int p = px + u * px_step;
        if (p >= Wp) continue;
        if (kSplit) {
          uint4 oh = make_uint4(0, 0, 0, 0), ol = oh;
          if (live[u]) affine8_split<T>(raw[u], rawl[u], s8, h8, act, oh, ol);
          dst_row[(long long)p * cvp] = oh;
          dst_row[(long long)p * cvp + cv] = ol;
        } else {
          dst_row[(long long)p * cv] = live[u] ? affine8<T>(raw[u], s8, h8, act) : make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ CBAM spatial pooling, C = 256
// 8 lanes per pixel, each lane owns 4 x 8 channels (four independent 128-bit loads, every load instruction of a
// pixel's 8 lanes covers one full 128-byte line), local reduction then 3 shuffle steps.  grid (x, B).
// (compiling this for three CTAs per SM -- 80 registers, 32 bytes of spill -- measured 0.8 % SLOWER on the synthesis step)
template <typename T, bool kSplit>
__global__ void __launch_bounds__(256)
cbam_pool_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                 float2* __restrict__ pooled, int HW) {
  pdl_prologue();
  constexpr int C = 256;
  const int lane = threadIdx.x & 31, seg = lane & 7, sub = lane >> 3;
  const int b = blockIdx.y;
  float sc[4][8], sh[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sc[j][i] = scale[b * C + j * 64 + seg * 8 + i];
      sh[j][i] = shift[b * C + j * 64 + seg * 8 + i];
    }
  const int warps = (blockDim.x >> 5) * gridDim.x;
  constexpr int kPix = (kSplit ? 2 : 1) * (C / 8);   // 16-byte chunks per pixel
  const uint4* base = reinterpret_cast<const uint4*>(y) + (long long)b * HW * kPix;
  for (int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g * 4 < HW; g += warps) {
    const int pix = g * 4 + sub;
    uint4 raw[4], rawl[kSplit ? 4 : 1];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      raw[j] = base[(long long)pix * kPix + j * 8 + seg];
      if (kSplit) rawl[j] = base[(long long)pix * kPix + C / 8 + j * 8 + seg];
    }
    float s = 0.f, m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
      uint32_t wl[4] = {0, 0, 0, 0};
      if (kSplit) { wl[0] = rawl[j].x; wl[1] = rawl[j].y; wl[2] = rawl[j].z; wl[3] = rawl[j].w; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = Cvt<T>::unpack2(w[i]);
        if (kSplit) {
          const float2 l = Cvt<T>::unpack2(wl[i]);
          f.x += l.x;
          f.y += l.y;
        }
        const float v0 = fmaf(f.x, sc[j][2 * i], sh[j][2 * i]), v1 = fmaf(f.y, sc[j][2 * i + 1], sh[j][2 * i + 1]);
        s += v0 + v1;
        m = fmaxf(m, fmaxf(v0, v1));
      }
    }
#pragma unroll
    for (int o = 4; o; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if (seg == 0) pooled[(long long)b * HW + pix] = make_float2(s * (1.f / C), m);
  }
}

// sa[b][y][x] = sigmoid( sum_{ch,r,s} w[ch][r][s] * pooled[b][y+r-3][x+s-3][ch] ), zero padding
__global__ void cbam_spatial_conv_kernel(const float2* __restrict__ pooled, const float* __restrict__ w,
                                         float* __restrict__ sa, int B, int H, int W) {
  pdl_prologue();
  __shared__ float sw[98];
  for (int i = threadIdx.x; i < 98; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % W);
    long long r = i / W;
    const int yy = int(r % H);
    const int b = int(r / H);
    float acc = 0.f;
#pragma unroll
    for (int dr = 0; dr < 7; ++dr) {
      const int sy = yy + dr - 3;
      if (sy < 0 || sy >= H) continue;
#pragma unroll
      for (int ds = 0; ds < 7; ++ds) {
        const int sx = x + ds - 3;
        if (sx < 0 || sx >= W) continue;
        const float2 p = __ldg(&pooled[((long long)b * H + sy) * W + sx]);
        acc = fmaf(sw[dr * 7 + ds], p.x, acc);
        acc = fmaf(sw[49 + dr * 7 + ds], p.y, acc);
      }
    }
    sa[i] = 1.f / (1.f + __expf(-acc));
  }
}

// out_pad = residual + (y*scale+shift) * sa ; borders by reflection / zero.  res_pad has padding res_pw.
template <typename T, bool kSplit>
__global__ void __launch_bounds__(256)
residual_apply_pad_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ sa, const float2* __restrict__ pooled, const float* __restrict__ w_sa,
                          const T* __restrict__ res_pad, int res_pw, T* __restrict__ out,
                          int B, int H, int W, int C, int pad, int pad_mode) {
  pdl_prologue();
  // pooled != nullptr: the 7x7 spatial-attention conv + sigmoid (modules/model.py:36-38) is evaluated here, one source row at
  // a time into shared memory, instead of by a kernel of its own (98 MACs per pixel against 3 x 512 bytes moved per pixel)
  extern __shared__ float sa_smem[];   // [98] weights, [W] attention of the current source row
  float* sa_line = sa_smem + 98;
  if (pooled != nullptr) {
    for (int i = threadIdx.x; i < 98; i += 256) sa_smem[i] = w_sa[i];
  }
  const int Hp = H + 2 * pad, Wp = W + 2 * pad, cv = C / 8, cvp = kSplit ? 2 * cv : cv;
  const int Wr = W + 2 * res_pw, Hr = H + 2 * res_pw;
  const int c8 = threadIdx.x % cv, row_chunks = Wp * cvp;
  const int px0 = threadIdx.x / cv, px_step = 256 / cv;
  int cur_b = -1;
  float s8[8], h8[8];
  for (int row = blockIdx.x; row < B * Hp; row += gridDim.x) {
    const int b = row / Hp, py = row - b * Hp;
    if (b != cur_b) {
      cur_b = b;
      const float4* sc = reinterpret_cast<const float4*>(scale + b * C + c8 * 8);
      const float4* sh = reinterpret_cast<const float4*>(shift + b * C + c8 * 8);
      const float4 a0 = __ldg(sc), a1 = __ldg(sc + 1), b0 = __ldg(sh), b1 = __ldg(sh + 1);
      s8[0] = a0.x; s8[1] = a0.y; s8[2] = a0.z; s8[3] = a0.w; s8[4] = a1.x; s8[5] = a1.y; s8[6] = a1.z; s8[7] = a1.w;
      h8[0] = b0.x; h8[1] = b0.y; h8[2] = b0.z; h8[3] = b0.w; h8[4] = b1.x; h8[5] = b1.y; h8[6] = b1.z; h8[7] = b1.w;
    }
    int sy = py - pad;
    const bool row_inside = sy >= 0 && sy < H;
    sy = reflect_idx(sy, H);
    const uint4* y_row = reinterpret_cast<const uint4*>(y) + ((long long)b * H + sy) * W * cvp + c8;
    const uint4* r_row = reinterpret_cast<const uint4*>(res_pad) + (((long long)b * Hr + sy + res_pw) * Wr + res_pw) * cvp + c8;
    const float* sa_row = sa != nullptr ? sa + ((long long)b * H + sy) * W : nullptr;
    if (pooled != nullptr) {
      __syncthreads();                       // the previous row's readers are done with sa_line (and the weights are loaded)
      for (int x = threadIdx.x; x < W; x += 256) {
        float acc = 0.f;
#pragma unroll
        for (int dr = 0; dr < 7; ++dr) {
          const int qy = sy + dr - 3;
          if (qy < 0 || qy >= H) continue;
#pragma unroll
          for (int ds = 0; ds < 7; ++ds) {
            const int qx = x + ds - 3;
            if (qx < 0 || qx >= W) continue;
            const float2 pv = __ldg(&pooled[((long long)b * H + qy) * W + qx]);
            acc = fmaf(sa_smem[dr * 7 + ds], pv.x, acc);
            acc = fmaf(sa_smem[49 + dr * 7 + ds], pv.y, acc);
          }
        }
        sa_line[x] = 1.f / (1.f + __expf(-acc));
      }
      __syncthreads();
    }
    uint4* dst_row = reinterpret_cast<uint4*>(out) + (long long)row * row_chunks + c8;
    for (int px = px0; px < Wp; px += px_step * kRowILP) {
      uint4 raw[kRowILP], res[kRowILP], rawl[kSplit ? kRowILP : 1], resl[kSplit ? kRowILP : 1];
      float att[kRowILP];
      bool live[kRowILP];
#pragma unroll
      for (int u = 0; u < kRowILP; ++u) {
        const int p = px + u * px_step;
        int sx = p - pad;
        const bool inside = row_inside && sx >= 0 && sx < W;
        live[u] = p < Wp && (inside || pad_mode == DUCOSY_PAD_REFLECT);
        sx = reflect_idx(sx, W);
        if (live[u]) {
          raw[u] = y_row[(long long)sx * cvp];
          res[u] = r_row[(long long)sx * cvp];
          if (kSplit) {
            rawl[u] = y_row[(long long)sx * cvp + cv];
            resl[u] = r_row[(long long)sx * cvp + cv];
          }
          att[u] = pooled != nullptr ? sa_line[sx] : (sa_row != nullptr ? __ldg(sa_row + sx) : 1.f);
        }
      }
#pragma unroll
      for (int u = 0; u < kRowILP; ++u) {
        const int p = px + u * px_step;
        if (p >= Wp) continue;
        if (kSplit) {
          uint4 oh = make_uint4(0, 0, 0, 0), ol = oh;
          if (live[u]) {
            float f[8], g[8];
            join8<T>(raw[u], rawl[u], f);
            join8<T>(res[u], resl[u], g);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaf(fmaf(f[k], s8[k], h8[k]), att[u], g[k]);
            split8<T>(f, oh, ol);
          }
          dst_row[(long long)p * cvp] = oh;
          dst_row[(long long)p * cvp + cv] = ol;
          continue;
        }
        uint4 o = make_uint4(0, 0, 0, 0);
        if (live[u]) {
          const uint32_t yw[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
          const uint32_t rw[4] = {res[u].x, res[u].y, res[u].z, res[u].w};
          uint32_t ow[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = Cvt<T>::unpack2(yw[k]);
            const float2 g = Cvt<T>::unpack2(rw[k]);
            const float v0 = fmaf(f.x, s8[2 * k], h8[2 * k]), v1 = fmaf(f.y, s8[2 * k + 1], h8[2 * k + 1]);
            ow[k] = Cvt<T>::pack2(fmaf(v0, att[u], g.x), fmaf(v1, att[u], g.y));
          }
          o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        dst_row[(long long)p * cv] = o;
      }
    }
  }
}

int row_grid(int rows) {
  const int cap = (num_sms() > 0 ? num_sms() : 148) * 4;
  return rows < cap ? (rows > 0 ? rows : 1) : cap;
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" int ducosy_pack_conv_weight(const float* w, void* packed, int Cout, int Cin, int kh, int kw, int dtype,
                                       ducosy_stream_t stream) {
  DUCOSY_CHECK(w && packed && Cout > 0 && Cin > 0 && kh > 0 && kw > 0, DUCOSY_ERR_ARG, "pack_conv_weight: bad argument");
  const long long total = (long long)Cout * Cin * kh * kw;
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pack_conv_weight_kernel<T>, ew_grid(total, 256), 256, 0, (cudaStream_t)stream)(
                                      w, static_cast<T*>(packed), Cout, Cin, kh, kw, dtype == DUCOSY_F16X2 ? 2 : 1)));
  return check_launch("pack_conv_weight_kernel");
}

extern "C" int ducosy_pack_upconv_weight(const float* w, void* packed, int Cout, int Cin, int dtype,
                                         ducosy_stream_t stream) {
  DUCOSY_CHECK(w && packed && Cout > 0 && Cin > 0, DUCOSY_ERR_ARG, "pack_upconv_weight: bad argument");
  const long long total = 16LL * Cout * Cin;
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pack_upconv_weight_kernel<T>, ew_grid(total, 256), 256, 0, (cudaStream_t)stream)(
                                      w, static_cast<T*>(packed), Cout, Cin, dtype == DUCOSY_F16X2 ? 2 : 1)));
  return check_launch("pack_upconv_weight_kernel");
}

extern "C" int ducosy_pack_upconv_merged_weight(const float* w, void* packed, int Cout, int Cin, int dtype,
                                                ducosy_stream_t stream) {
  DUCOSY_CHECK(w && packed && Cout > 0 && Cin > 0, DUCOSY_ERR_ARG, "pack_upconv_merged_weight: bad argument");
  DUCOSY_CHECK(dtype != DUCOSY_F16X2, DUCOSY_ERR_ARG, "pack_upconv_merged_weight: not available in split-operand mode");
  const long long total = 36LL * Cout * Cin;
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pack_upconv_merged_weight_kernel<T>, ew_grid(total, 256), 256, 0, (cudaStream_t)stream)(
                                      w, static_cast<T*>(packed), Cout, Cin)));
  return check_launch("pack_upconv_merged_weight_kernel");
}

extern "C" int ducosy_pack_stem_weight(const float* w, void* packed, int Cin, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(w && packed && Cin > 0, DUCOSY_ERR_ARG, "pack_stem_weight: bad argument");
  const int Kpad = (49 * Cin + 63) / 64 * 64;
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pack_stem_weight_kernel<T>, ew_grid(64 * Kpad, 256), 256, 0, (cudaStream_t)stream)(
                                      w, static_cast<T*>(packed), Cin, Kpad, dtype == DUCOSY_F16X2 ? 2 : 1)));
  return check_launch("pack_stem_weight_kernel");
}

extern "C" int ducosy_stem_im2col(const float* x, void* a_mat, int B, int Cin, int H, int W, int dtype,
                                  ducosy_stream_t stream) {
  DUCOSY_CHECK(x && a_mat && B > 0 && Cin > 0, DUCOSY_ERR_ARG, "stem_im2col: bad argument");
  DUCOSY_CHECK(H >= 4 && W >= 4, DUCOSY_ERR_SHAPE, "stem_im2col: reflect pad 3 needs H,W >= 4");
  InF32 in{x};
  DUCOSY_DISPATCH_DTYPE(dtype, T, return (launch_im2col<T, InF32>(in, static_cast<T*>(a_mat), B, Cin, H, W, dtype == DUCOSY_F16X2, (cudaStream_t)stream)));
}

extern "C" int ducosy_stem_im2col_hu(const int16_t* px, void* a_mat, int B, int H, int W, float slope, float intercept,
                                     float lo, float hi, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(px && a_mat && B > 0, DUCOSY_ERR_ARG, "stem_im2col_hu: bad argument");
  DUCOSY_CHECK(H >= 4 && W >= 4, DUCOSY_ERR_SHAPE, "stem_im2col_hu: reflect pad 3 needs H,W >= 4");
  InHU in{px, slope, intercept, lo, hi, float(double(hi) - double(lo))};
  DUCOSY_DISPATCH_DTYPE(dtype, T, return (launch_im2col<T, InHU>(in, static_cast<T*>(a_mat), B, 1, H, W, dtype == DUCOSY_F16X2, (cudaStream_t)stream)));
}

extern "C" int ducosy_in_finalize(const float* partials, int tiles_per_sample, int npix_per_sample, float* scale,
                                  float* shift, const float* fc0, const float* fc2, float* chmax, int B, int C,
                                  ducosy_stream_t stream) {
  DUCOSY_CHECK(partials && scale && shift && B > 0 && C > 0 && tiles_per_sample > 0 && npix_per_sample > 0,
               DUCOSY_ERR_ARG, "in_finalize: bad argument");
  DUCOSY_CHECK((fc0 == nullptr) == (fc2 == nullptr), DUCOSY_ERR_ARG, "in_finalize: fc0 and fc2 go together");
  DUCOSY_CHECK(fc0 == nullptr || chmax != nullptr, DUCOSY_ERR_ARG, "in_finalize: CBAM needs the chmax scratch [B][C]");
  DUCOSY_CHECK(C % 32 == 0, DUCOSY_ERR_SHAPE, "in_finalize: C %% 32 != 0");
  pdl(in_finalize_kernel, dim3(C / 32, B), kFinWarps * 32, 0, (cudaStream_t)stream)(partials, tiles_per_sample, npix_per_sample,
                                                                        scale, shift, chmax, C);
  DUCOSY_TRY(check_launch("in_finalize_kernel"));
  if (fc0 != nullptr) {
    const size_t smem = sizeof(float) * (C + C / 16 + 1);
    pdl(cbam_channel_mlp_kernel, B, 256, smem, (cudaStream_t)stream)(chmax, fc0, fc2, scale, shift, C);
    return check_launch("cbam_channel_mlp_kernel");
  }
  return 0;
}

extern "C" int ducosy_in_apply_pad(const void* y, const float* scale, const float* shift, void* out_pad, int B, int H,
                                   int W, int C, int pad, int pad_mode, int act, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(y && scale && shift && out_pad && B > 0, DUCOSY_ERR_ARG, "in_apply_pad: bad argument");
  DUCOSY_CHECK(C % 8 == 0 && pad >= 0 && pad < H && pad < W, DUCOSY_ERR_SHAPE, "in_apply_pad: C %% 8 != 0 or pad too large");
  DUCOSY_CHECK(al16(y) && al16(out_pad) && al16(scale) && al16(shift), DUCOSY_ERR_ALIGN, "in_apply_pad: 16-byte alignment");
  DUCOSY_CHECK(256 % (C / 8) == 0, DUCOSY_ERR_SHAPE, "in_apply_pad: C must be one of 8..2048 with C/8 dividing 256");
  static const bool cap = []() { const char* e = getenv("DUCOSY_APPLY_CAP"); return e == nullptr || atoi(e) != 0; }();
  if (dtype == DUCOSY_F16X2)
    pdl(in_apply_pad_kernel<__half, true>, row_grid(B * (H + 2 * pad)), 256, 0, (cudaStream_t)stream)(
        static_cast<const __half*>(y), scale, shift, static_cast<__half*>(out_pad), B, H, W, C, pad, pad_mode, act);
  else if (cap)
      DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(in_apply_pad_kernel<T, false, true>, row_grid(B * (H + 2 * pad)), 256, 0, (cudaStream_t)stream)(
                                          static_cast<const T*>(y), scale, shift, static_cast<T*>(out_pad), B, H, W, C, pad,
                                          pad_mode, act)));
    else
      DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(in_apply_pad_kernel<T, false, false>, row_grid(B * (H + 2 * pad)), 256, 0, (cudaStream_t)stream)(
                                          static_cast<const T*>(y), scale, shift, static_cast<T*>(out_pad), B, H, W, C, pad,
                                          pad_mode, act)));
  return check_launch("in_apply_pad_kernel");
}

extern "C" int ducosy_cbam_pool(const void* y, const float* scale, const float* shift, float* pooled, int B, int H, int W,
                                int C, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(y && scale && shift && pooled && B > 0, DUCOSY_ERR_ARG, "cbam_pool: bad argument");
  DUCOSY_CHECK(C == 256, DUCOSY_ERR_SHAPE, "cbam_pool: C must be 256 (got %d)", C);
  DUCOSY_CHECK((H * W) % 4 == 0, DUCOSY_ERR_SHAPE, "cbam_pool: H*W must be a multiple of 4");
  const int groups = H * W / 4;
  int gx = (groups + 7) / 8;                     // 8 warps per CTA
  const int cap = (num_sms() > 0 ? num_sms() : 148) * 8 / (B > 0 ? B : 1) + 1;
  if (gx > cap) gx = cap;
  if (dtype == DUCOSY_F16X2)
    pdl(cbam_pool_kernel<__half, true>, dim3(gx, B), 256, 0, (cudaStream_t)stream)(static_cast<const __half*>(y), scale, shift,
                                                                                   reinterpret_cast<float2*>(pooled), H * W);
  else
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(cbam_pool_kernel<T, false>, dim3(gx, B), 256, 0, (cudaStream_t)stream)(
                                        static_cast<const T*>(y), scale, shift, reinterpret_cast<float2*>(pooled), H * W)));
  return check_launch("cbam_pool_kernel");
}

extern "C" int ducosy_cbam_spatial_conv(const float* pooled, const float* w_sa, float* sa, int B, int H, int W,
                                        ducosy_stream_t stream) {
  DUCOSY_CHECK(pooled && w_sa && sa && B > 0, DUCOSY_ERR_ARG, "cbam_spatial_conv: bad argument");
  pdl(cbam_spatial_conv_kernel, ew_grid((long long)B * H * W, 128), 128, 0, (cudaStream_t)stream)(
      reinterpret_cast<const float2*>(pooled), w_sa, sa, B, H, W);
  return check_launch("cbam_spatial_conv_kernel");
}

extern "C" int ducosy_residual_apply_pad(const void* y, const float* scale, const float* shift, const float* sa,
                                         const void* res_pad, int res_pad_width, void* out_pad, int B, int H, int W,
                                         int C, int pad, int pad_mode, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(y && scale && shift && res_pad && out_pad && B > 0, DUCOSY_ERR_ARG, "residual_apply_pad: bad argument");
  DUCOSY_CHECK(C % 8 == 0 && pad >= 0 && pad < H && pad < W && res_pad_width >= 0, DUCOSY_ERR_SHAPE,
               "residual_apply_pad: bad shape");
  DUCOSY_CHECK(res_pad != out_pad, DUCOSY_ERR_ARG, "residual_apply_pad: in-place is not supported");
  DUCOSY_CHECK(256 % (C / 8) == 0, DUCOSY_ERR_SHAPE, "residual_apply_pad: C/8 must divide 256");
  if (dtype == DUCOSY_F16X2)
    pdl(residual_apply_pad_kernel<__half, true>, row_grid(B * (H + 2 * pad)), 256, 0, (cudaStream_t)stream)(
        static_cast<const __half*>(y), scale, shift, sa, nullptr, nullptr, static_cast<const __half*>(res_pad), res_pad_width,
        static_cast<__half*>(out_pad), B, H, W, C, pad, pad_mode);
  else
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(residual_apply_pad_kernel<T, false>, row_grid(B * (H + 2 * pad)), 256, 0, (cudaStream_t)stream)(
                                        static_cast<const T*>(y), scale, shift, sa, nullptr, nullptr, static_cast<const T*>(res_pad),
                                        res_pad_width, static_cast<T*>(out_pad), B, H, W, C, pad, pad_mode)));
  return check_launch("residual_apply_pad_kernel");
}

// The same with the spatial attention evaluated inside: pooled [B][H][W][2] (channel mean, channel max of the attended map,
// ducosy_cbam_pool) and the 7x7 conv weight w_sa [1][2][7][7] replace the precomputed sa map (and the launch of
// ducosy_cbam_spatial_conv): modules/model.py:34-39 + :83-87 in one pass.
extern "C" int ducosy_residual_cbam_apply_pad(const void* y, const float* scale, const float* shift, const float* pooled,
                                              const float* w_sa, const void* res_pad, int res_pad_width, void* out_pad, int B,
                                              int H, int W, int C, int pad, int pad_mode, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(y && scale && shift && pooled && w_sa && res_pad && out_pad && B > 0, DUCOSY_ERR_ARG, "residual_cbam_apply_pad: bad argument");
  DUCOSY_CHECK(C % 8 == 0 && pad >= 0 && pad < H && pad < W && res_pad_width >= 0 && W <= 8192, DUCOSY_ERR_SHAPE,
               "residual_cbam_apply_pad: bad shape");
  DUCOSY_CHECK(res_pad != out_pad, DUCOSY_ERR_ARG, "residual_cbam_apply_pad: in-place is not supported");
  DUCOSY_CHECK(256 % (C / 8) == 0, DUCOSY_ERR_SHAPE, "residual_cbam_apply_pad: C/8 must divide 256");
  const size_t smem = (98 + size_t(W)) * sizeof(float);
  if (dtype == DUCOSY_F16X2)
    pdl(residual_apply_pad_kernel<__half, true>, row_grid(B * (H + 2 * pad)), 256, smem, (cudaStream_t)stream)(
        static_cast<const __half*>(y), scale, shift, nullptr, reinterpret_cast<const float2*>(pooled), w_sa,
        static_cast<const __half*>(res_pad), res_pad_width, static_cast<__half*>(out_pad), B, H, W, C, pad, pad_mode);
  else
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(residual_apply_pad_kernel<T, false>, row_grid(B * (H + 2 * pad)), 256, smem, (cudaStream_t)stream)(
                                        static_cast<const T*>(y), scale, shift, nullptr, reinterpret_cast<const float2*>(pooled), w_sa,
                                        static_cast<const T*>(res_pad), res_pad_width, static_cast<T*>(out_pad), B, H, W, C, pad, pad_mode)));
  return check_launch("residual_apply_pad_kernel");
}
