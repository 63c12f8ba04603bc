// Generator backward pieces that are not tensor-core GEMMs (what autograd computes through modules/model.py:80-115 for
// modules/trainer.py:497): the 64 -> 1 output conv + tanh, the adjoint of the 7x7 stem's reflection-padded im2col, and the
// CBAM / residual-block element-wise adjoints.  Everything is bandwidth-bound CUDA-core work; reductions run in a
// fixed order (deterministic gradients).  16-bit gradient maps carry the power-of-two scale gs[0] (ducosy_grad_scale);
// fp32 parameter / image gradients leave with the true scale.
#include <algorithm>
#include <cmath>
#include <cstdlib>

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
// dvh (optional): the same map times the power-of-two gradient scale gs[0], rounded to T -- the A operand of the tensor-core
// input-gradient kernel below (|dv| <= |dout| and gs[0] * max|dout| is in [1, 2), so nothing overflows or goes subnormal).
template <typename T>
__global__ void __launch_bounds__(256)
out_tanh_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dv, T* __restrict__ dvh,
                    const float* __restrict__ gs, float* __restrict__ partial, long long n) {
  pdl_prologue();
  __shared__ float red[8];
  float acc = 0.f;
  const float sc = dvh != nullptr ? gs[0] : 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float o = out[i], g = dout[i] * (1.f - o * o);
    dv[i] = g;
    if (dvh != nullptr) dvh[i] = Cvt<T>::from_f(g * sc);
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
  pdl_prologue();
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
// thread = (4 pixels along x, 8 channels); weights transposed to [tap][channel] in shared memory.  Away from the border
// (one padded position per pixel, every tap inside the image) a row of 10 dv values feeds 7 taps x 4 pixels x 8 channels
// of FMAs; the 2 % of threads that touch the reflection / the image edge take the general path.
template <typename T>
__global__ void __launch_bounds__(256)
out_conv_dgrad_kernel(const float* __restrict__ dv, const float* __restrict__ w, T* __restrict__ da,
                      const float* __restrict__ gs, int B, int H, int W) {
  pdl_prologue();
  __shared__ __align__(16) float ws[49][64];
  for (int i = threadIdx.x; i < 49 * 64; i += 256) ws[i % 49][i / 49] = w[i];   // w is [c][tap]
  __syncthreads();
  const long long item = (long long)blockIdx.x * 256 + threadIdx.x;
  const int Wg = W / 4;
  const long long total = (long long)B * H * Wg * 8;
  if (item >= total) return;
  const int c8 = int(item & 7);
  const long long grp = item >> 3;
  const int x0 = int(grp % Wg) * 4, y = int((grp / Wg) % H), b = int(grp / ((long long)Wg * H));
  const float* dvb = dv + (long long)b * H * W;
  const float sc = gs[0];
  T* dst = da + (((long long)b * H + y) * W + x0) * 64 + c8 * 8;
  if (y >= 4 && y <= H - 5 && x0 >= 4 && x0 + 3 <= W - 5) {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const float* row = dvb + (long long)(y + 3 - r) * W + (x0 - 3);
      float d[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) d[j] = __ldg(row + j);
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const float4 w0 = *reinterpret_cast<const float4*>(&ws[r * 7 + s][c8 * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&ws[r * 7 + s][c8 * 8 + 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float g = d[i + 6 - s];   // dv[y+3-r][x0+i+3-s]
          acc[i][0] += g * w0.x; acc[i][1] += g * w0.y; acc[i][2] += g * w0.z; acc[i][3] += g * w0.w;
          acc[i][4] += g * w1.x; acc[i][5] += g * w1.y; acc[i][6] += g * w1.z; acc[i][7] += g * w1.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float o8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o8[k] = acc[i][k] * sc;
      uint4 o;
      o.x = Cvt<T>::pack2(o8[0], o8[1]); o.y = Cvt<T>::pack2(o8[2], o8[3]);
      o.z = Cvt<T>::pack2(o8[4], o8[5]); o.w = Cvt<T>::pack2(o8[6], o8[7]);
      *reinterpret_cast<uint4*>(dst + i * 64) = o;
    }
    return;
  }
  for (int i = 0; i < 4; ++i) {
    const int x = x0 + i;
    int us[3], vs[3];
    const int nu = reflect_sources(y, H, 3, us), nv = reflect_sources(x, W, 3, vs);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int iu = 0; iu < nu; ++iu)
      for (int iv = 0; iv < nv; ++iv) {
        const int u = us[iu], v = vs[iv];
        for (int r = 0; r < 7; ++r) {
          const int yo = u - r;
          if (yo < 0 || yo >= H) continue;
          for (int s2 = 0; s2 < 7; ++s2) {
            const int xo = v - s2;
            if (xo < 0 || xo >= W) continue;
            const float g = __ldg(dvb + (long long)yo * W + xo);
            const float4 w0 = *reinterpret_cast<const float4*>(&ws[r * 7 + s2][c8 * 8]);
            const float4 w1 = *reinterpret_cast<const float4*>(&ws[r * 7 + s2][c8 * 8 + 4]);
            acc[0] += g * w0.x; acc[1] += g * w0.y; acc[2] += g * w0.z; acc[3] += g * w0.w;
            acc[4] += g * w1.x; acc[5] += g * w1.y; acc[6] += g * w1.z; acc[7] += g * w1.w;
          }
        }
      }
    uint4 o;
    o.x = Cvt<T>::pack2(acc[0] * sc, acc[1] * sc);
    o.y = Cvt<T>::pack2(acc[2] * sc, acc[3] * sc);
    o.z = Cvt<T>::pack2(acc[4] * sc, acc[5] * sc);
    o.w = Cvt<T>::pack2(acc[6] * sc, acc[7] * sc);
    *reinterpret_cast<uint4*>(dst + i * 64) = o;
  }
}

// The 4-pixel frame of the image (where ReflectionPad2d(3) folds mirrored positions onto a pixel, or a tap leaves the image)
// after out_conv_dgrad_mma_kernel wrote the zero-extended sums everywhere: thread = (frame pixel, 8 channels), overwrites da.
// Frame pixels per sample: rows 0..3 and H-4..H-1 in full (8 W), columns 0..3 and W-4..W-1 of the other rows (8 (H - 8)).
template <typename T>
__global__ void __launch_bounds__(256)
out_conv_dgrad_border_kernel(const float* __restrict__ dv, const float* __restrict__ w, T* __restrict__ da,
                             const float* __restrict__ gs, int B, int H, int W) {
  pdl_prologue();
  __shared__ __align__(16) float ws[49][64];
  for (int i = threadIdx.x; i < 49 * 64; i += 256) ws[i % 49][i / 49] = w[i];   // w is [c][tap]
  __syncthreads();
  const int per_sample = 8 * W + 8 * (H - 8);
  const long long item = (long long)blockIdx.x * 256 + threadIdx.x;
  if (item >= (long long)B * per_sample * 8) return;
  const int c8 = int(item & 7);
  const long long pid = item >> 3;
  const int b = int(pid / per_sample), idx = int(pid - (long long)b * per_sample);
  int y, x;
  if (idx < 4 * W) { y = idx / W; x = idx - y * W; }
  else if (idx < 8 * W) { const int k = idx - 4 * W; y = H - 4 + k / W; x = k % W; }
  else { const int k = idx - 8 * W, j = k & 7; y = 4 + (k >> 3); x = j < 4 ? j : W - 8 + j; }
  const float* dvb = dv + (long long)b * H * W;
  const float sc = gs[0];
  int us[3], vs[3];
  const int nu = reflect_sources(y, H, 3, us), nv = reflect_sources(x, W, 3, vs);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int iu = 0; iu < nu; ++iu)
    for (int iv = 0; iv < nv; ++iv) {
      const int u = us[iu], v = vs[iv];
      for (int r = 0; r < 7; ++r) {
        const int yo = u - r;
        if (yo < 0 || yo >= H) continue;
        for (int s2 = 0; s2 < 7; ++s2) {
          const int xo = v - s2;
          if (xo < 0 || xo >= W) continue;
          const float gq = __ldg(dvb + (long long)yo * W + xo);
          const float4 w0 = *reinterpret_cast<const float4*>(&ws[r * 7 + s2][c8 * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&ws[r * 7 + s2][c8 * 8 + 4]);
          acc[0] += gq * w0.x; acc[1] += gq * w0.y; acc[2] += gq * w0.z; acc[3] += gq * w0.w;
          acc[4] += gq * w1.x; acc[5] += gq * w1.y; acc[6] += gq * w1.z; acc[7] += gq * w1.w;
        }
      }
    }
  uint4 o;
  o.x = Cvt<T>::pack2(acc[0] * sc, acc[1] * sc);
  o.y = Cvt<T>::pack2(acc[2] * sc, acc[3] * sc);
  o.z = Cvt<T>::pack2(acc[4] * sc, acc[5] * sc);
  o.w = Cvt<T>::pack2(acc[6] * sc, acc[7] * sc);
  *reinterpret_cast<uint4*>(da + (((long long)b * H + y) * W + x) * 64 + c8 * 8) = o;
}

// ---- the same input gradient on warp-level tensor cores (interior + zero-extended edges; the 4-pixel frame that the reflection
// touches is then overwritten by out_conv_dgrad_border_kernel above).  GEMM view per image row:
//     da[x][c] = sum_k A[x][k] * Bw[k][c],   A[x][k = r*7 + s] = dvh[y + 3 - r][x + 3 - s]  (Toeplitz, zero outside the image),
//     Bw[k][c] = w[c][r][s],  K = 49 padded to 64
// A warp owns 16 consecutive pixels of a row per step (m16n8k16: 8 channel tiles x 4 k steps = 32 MMAs); its A fragments are 2-byte
// gathers from a shared-memory tile of dvh (rows y-3..y+3, 70 columns), the 64 B-fragment registers hold the whole filter for
// the lifetime of the (persistent) CTA; results go through a per-warp staging tile so that every pixel is one 128-byte store.
template <typename T> struct GMma;
template <> struct GMma<__half> {
  static __device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
};
template <> struct GMma<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
};

constexpr int kDgRows = 4, kDgCols = 64;                 // output tile of a 4-warp CTA: warp w <-> row y0 + w, 4 steps of 16 pixels
constexpr int kDgPitch = kDgCols + 6 + 2;                // staged dvh columns x0-3 .. x0+66, padded to 72
constexpr int kDgTileRows = kDgRows + 6;
template <typename T>
__global__ void __launch_bounds__(128)
out_conv_dgrad_mma_kernel(const T* __restrict__ dvh, const float* __restrict__ w, T* __restrict__ da, int B, int H, int W) {
  pdl_prologue();
  __shared__ __align__(16) unsigned short tile[kDgTileRows * kDgPitch + 8];   // + a zero slot for the padded taps (k >= 49)
  __shared__ __align__(16) T stage[kDgRows][16][64 + 8];                      // per warp: 16 pixels x 64 channels (+ 16-byte pad)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  constexpr int kZero = kDgTileRows * kDgPitch;
  // B fragments: b0 = (k = 16ks + 2t, +1; n = 8j + g), b1 = (k + 8, + 9)
  uint32_t bf[4][8][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k0 = 16 * ks + 2 * t + 8 * h, c = 8 * j + g;
        const float v0 = k0 < 49 ? __ldg(w + c * 49 + k0) : 0.f, v1 = k0 + 1 < 49 ? __ldg(w + c * 49 + k0 + 1) : 0.f;
        bf[ks][j][h] = Cvt<T>::pack2(v0, v1);
      }
  // tile offsets of this thread's A columns: k -> (6 - r) * pitch + (6 - s) relative to (row w, pixel m) of the staged tile
  int off[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 16 * ks + 2 * t + (q & 1) + 8 * (q >> 1);
      off[ks][q] = k < 49 ? (6 - k / 7) * kDgPitch + (6 - k % 7) : -1;
    }
  if (threadIdx.x < 8) tile[kZero + threadIdx.x] = 0;
  const int tiles_x = W / kDgCols, tiles_y = H / kDgRows;
  const int total = B * tiles_y * tiles_x;
  for (int tl = blockIdx.x; tl < total; tl += gridDim.x) {
    const int tx = tl % tiles_x, ty = (tl / tiles_x) % tiles_y, b = tl / (tiles_x * tiles_y);
    const int y0 = ty * kDgRows, x0 = tx * kDgCols;
    __syncthreads();                                   // the previous tile's readers are done
    for (int i = threadIdx.x; i < kDgTileRows * kDgPitch; i += 128) {
      const int rr = i / kDgPitch, cc = i - rr * kDgPitch;
      const int yy = y0 - 3 + rr, xx = x0 - 3 + cc;
      unsigned short v = 0;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W && cc < kDgCols + 6)
        v = reinterpret_cast<const unsigned short*>(dvh)[((long long)b * H + yy) * W + xx];
      tile[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int mt = 0; mt < kDgCols / 16; ++mt) {
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
      const int base_a = warp * kDgPitch + mt * 16 + g;   // (row w, pixel g); pixel g + 8 is 8 columns further
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {                      // h: k-half (columns 2t.. / 2t+8..)
          const int o0 = off[ks][2 * h], o1 = off[ks][2 * h + 1];
          const uint32_t lo_g = tile[o0 >= 0 ? base_a + o0 : kZero], hi_g = tile[o1 >= 0 ? base_a + o1 : kZero];
          const uint32_t lo_8 = tile[o0 >= 0 ? base_a + 8 + o0 : kZero], hi_8 = tile[o1 >= 0 ? base_a + 8 + o1 : kZero];
          a[2 * h] = lo_g | (hi_g << 16);                  // a0 / a2: row g
          a[2 * h + 1] = lo_8 | (hi_8 << 16);              // a1 / a3: row g + 8
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) GMma<T>::mma(acc[j], a[0], a[1], a[2], a[3], bf[ks][j][0], bf[ks][j][1]);
      }
      // D fragment: c0,c1 = (pixel g, channels 8j + 2t, +1), c2,c3 = (pixel g + 8, ...)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint32_t*>(&stage[warp][g][8 * j + 2 * t]) = Cvt<T>::pack2(acc[j][0], acc[j][1]);
        *reinterpret_cast<uint32_t*>(&stage[warp][g + 8][8 * j + 2 * t]) = Cvt<T>::pack2(acc[j][2], acc[j][3]);
      }
      __syncwarp();
      T* dst = da + (((long long)b * H + y0 + warp) * W + x0 + mt * 16) * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) {                        // 16 pixels x 8 chunks of 16 bytes = 128 chunks, 4 per lane
        const int ch = i * 32 + lane, px = ch >> 3, c8 = ch & 7;
        *reinterpret_cast<uint4*>(dst + px * 64 + c8 * 8) = *reinterpret_cast<const uint4*>(&stage[warp][px][c8 * 8]);
      }
      __syncwarp();
    }
  }
}

// dw[c][r][s] = sum over padded pixels (u,v) of in_pad[u][v][c] * dv[u-r][v-s]   (dv zero outside the image)
// block = 448 threads = 64 channels x 7 filter rows; one padded row per iteration, the 7 dv rows it meets staged in
// shared memory with zero borders.  A thread keeps the 7 dv values its filter row needs for the current pixel in a
// register window that slides by one per pixel: 1 shared load + 1 global load per 7 FMAs.  partial: [gridDim.x][64*49].
constexpr int kOutWgradThreads = 448;
template <typename T>
__global__ void __launch_bounds__(kOutWgradThreads)
out_conv_wgrad_kernel(const T* __restrict__ in_pad, const float* __restrict__ dv, float* __restrict__ partial, int B, int H, int W) {
  pdl_prologue();
  extern __shared__ float rows[];                 // [7][W + 12], rows[r][j] = dv[u - r][j - 6]
  const int Wp = W + 6, Hp = H + 6, pitch = W + 12;
  const int c = threadIdx.x & 63, r = threadIdx.x >> 6;
  float acc[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) acc[i] = 0.f;
  for (int row = blockIdx.x; row < B * Hp; row += gridDim.x) {
    const int b = row / Hp, u = row % Hp;
    __syncthreads();
    for (int i = threadIdx.x; i < 7 * pitch; i += kOutWgradThreads) {
      const int rr = i / pitch, j = i % pitch, yo = u - rr, xo = j - 6;
      rows[i] = (yo >= 0 && yo < H && xo >= 0 && xo < W) ? dv[((long long)b * H + yo) * W + xo] : 0.f;
    }
    __syncthreads();
    const T* src = in_pad + ((long long)row * Wp) * 64 + c;
    const float* rw = rows + r * pitch;           // tap s of pixel v reads rw[v + 6 - s]
    float win[7];                                 // win[k] = rw[v + k], k = 0..6  (tap s uses win[6 - s])
#pragma unroll
    for (int k = 0; k < 6; ++k) win[k + 1] = rw[k];
#pragma unroll 7
    for (int v = 0; v < Wp; ++v) {
#pragma unroll
      for (int k = 0; k < 6; ++k) win[k] = win[k + 1];
      win[6] = rw[v + 6];
      const float a = Cvt<T>::to_f(src[(long long)v * 64]);
#pragma unroll
      for (int s = 0; s < 7; ++s) acc[s] = fmaf(a, win[6 - s], acc[s]);
    }
  }
  float* dst = partial + (long long)blockIdx.x * (64 * 49) + c * 49 + r * 7;
#pragma unroll
  for (int s = 0; s < 7; ++s) dst[s] = acc[s];
}

// ---- the same weight gradient on warp-level tensor cores.  GEMM view per padded row u, 16 padded pixels v0..v0+15 per step:
//     D[c][n] += sum_k A[c][k] * Bm[k][n],   A[c][k] = in_pad[u][v0 + k][c],   Bm[k][n = r*7 + s] = dvh[u - r][v0 + k - s]
// (M = 64 channels = 4 m-tiles, N = 49 taps padded to 56 = 7 n-tiles, K = pixels).  in_pad is pixel-major, so the A fragments
// come out of ldmatrix.trans on a per-warp staging tile filled with cp.async (double-buffered, no CTA barrier in the pixel
// loop); the Toeplitz B fragments are aligned 32-bit reads from two copies of the 7 staged dvh rows, the second shifted by one
// element so that the (k, k+1) pair of every tap is word-aligned.  Every warp walks its own pixel steps with all 28
// accumulator tiles in registers; the four warps' partial sums are added in shared memory and each CTA writes one
// [64][49] partial, in units of gs[0] (dvh is the scaled 16-bit map), which out_conv_wgrad_reduce_kernel folds back.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int bytes = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}

constexpr int kWgA = 64 + 8;            // elements per staged pixel (16-byte pad: conflict-free ldmatrix rows)
template <typename T>
__global__ void __launch_bounds__(128)
out_conv_wgrad_mma_kernel(const T* __restrict__ in_pad, const T* __restrict__ dvh, float* __restrict__ partial, int B, int H, int W) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char wg_smem[];
  const int Wp = W + 6, Hp = H + 6, pitch = ((W + 12 + 16 + 7) / 8) * 8;   // staged dvh row: 6 zeros | W values | zeros (steps run past Wp)
  unsigned short* rows0 = reinterpret_cast<unsigned short*>(wg_smem);      // [7][pitch]      rows0[r][j] = dvh[u - r][j - 6]
  unsigned short* rows1 = rows0 + 7 * pitch;                               // [7][pitch]      rows1[r][j] = rows0[r][j + 1]
  T* abuf = reinterpret_cast<T*>(rows1 + 7 * pitch);                       // [4 warps][2][16][kWgA]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  T* my_a = abuf + warp * 2 * 16 * kWgA;
  float acc[4][7][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
  // tap of this thread's B column in n-tile j: n = 8j + g -> (r, s); pair (k, k+1) starts at rows[slot(u - r)][v0 + k - s + 6]
  int tap_r[7], tap_c[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int n = 8 * j + g;
    tap_r[j] = n < 49 ? n / 7 : -1;
    tap_c[j] = 6 - (n % 7);
  }
  const int steps = (Wp + 15) / 16;
  // a CTA walks a contiguous range of padded rows, so the 7 staged dvh rows form a ring: image row yo lives in slot (yo + 7) % 7
  // and moving down one padded row replaces exactly one slot
  const int rows_total = B * Hp, per_cta = (rows_total + gridDim.x - 1) / gridDim.x;
  const int row_lo = blockIdx.x * per_cta, row_hi = min(rows_total, row_lo + per_cta);
  int have_b = -1, have_u = -100;
  for (int row = row_lo; row < row_hi; ++row) {
    const int b = row / Hp, u = row - b * Hp;
    const bool slide = b == have_b && u == have_u + 1;
    have_b = b;
    have_u = u;
    __syncthreads();                                     // every warp is done with the previous row's use of the ring
    for (int i = threadIdx.x; i < (slide ? 1 : 7) * pitch; i += 128) {
      const int rr = slide ? 0 : i / pitch, j = i - (slide ? 0 : rr * pitch), yo = u - rr;
      auto at = [&](int jj) -> unsigned short {
        const int xo = jj - 6;
        return (yo >= 0 && yo < H && xo >= 0 && xo < W) ? reinterpret_cast<const unsigned short*>(dvh)[((long long)b * H + yo) * W + xo] : (unsigned short)0;
      };
      const int slot = (yo + 14) % 7;
      rows0[slot * pitch + j] = at(j);
      rows1[slot * pitch + j] = at(j + 1);
    }
    __syncthreads();
    int boff[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) boff[j] = tap_r[j] >= 0 ? ((u - tap_r[j] + 14) % 7) * pitch + tap_c[j] : -1;
    const T* src = in_pad + (long long)row * Wp * 64;
    auto stage = [&](int st, int buf) {                  // 16 pixels x 128 bytes = 128 chunks of 16 bytes, 4 per lane
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ch = i * 32 + lane, px = ch >> 3, c8 = ch & 7, v = st * 16 + px;
        cp_async16(my_a + (buf * 16 + px) * kWgA + c8 * 8, src + (long long)(v < Wp ? v : 0) * 64 + c8 * 8, v < Wp);
      }
      cp_async_commit();
    };
    int buf = 0;
    if (warp < steps) stage(warp, 0);
    for (int st = warp; st < steps; st += 4) {
      if (st + 4 < steps) { stage(st + 4, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      __syncwarp();
      const int v0 = st * 16;
      uint32_t bfr[7][2];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        if (boff[j] < 0) { bfr[j][0] = 0; bfr[j][1] = 0; continue; }
        const int e = boff[j] + v0 + 2 * t;              // element index of (k = 2t) in rows0; word-aligned in rows0 or rows1
        const unsigned short* base = (e & 1) ? rows1 + (e - 1) : rows0 + e;
        bfr[j][0] = *reinterpret_cast<const uint32_t*>(base);
        bfr[j][1] = *reinterpret_cast<const uint32_t*>(base + 8);
      }
      const T* a_tile = my_a + buf * 16 * kWgA;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // matrices: (px 0-7, ch 16i..+7), (px 0-7, ch 16i+8..), (px 8-15, ch 16i..+7), (px 8-15, ch 16i+8..); lane l -> row l%8 of matrix l/8
        uint32_t a[4];
        const int mtx = lane >> 3, rr = lane & 7;
        ldmatrix_x4_trans(a, a_tile + ((mtx >> 1) * 8 + rr) * kWgA + 16 * i + (mtx & 1) * 8);
#pragma unroll
        for (int j = 0; j < 7; ++j) GMma<T>::mma(acc[i][j], a[0], a[1], a[2], a[3], bfr[j][0], bfr[j][1]);
      }
      __syncwarp();
      buf ^= 1;
    }
  }
  // add the four warps' partial sums (fixed order) and write this CTA's [64][49] partial
  __syncthreads();
  float* red = reinterpret_cast<float*>(wg_smem);        // [64][56] floats = 14 KB, over the dvh tiles
  for (int wsel = 0; wsel < 4; ++wsel) {
    if (warp == wsel) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = 16 * i + g + 8 * (q >> 1), n = 8 * j + 2 * t + (q & 1);
            if (wsel == 0) red[c * 56 + n] = acc[i][j][q]; else red[c * 56 + n] += acc[i][j][q];
          }
    }
    __syncthreads();
  }
  float* dst = partial + (long long)blockIdx.x * (64 * 49);
  for (int i = threadIdx.x; i < 64 * 49; i += 128) dst[i] = red[(i / 49) * 56 + i % 49];
}

// gs != nullptr: the partials are in units of gs[0] (tensor-core path), gs[1] = 1 / gs[0] (a power of two: exact)
__global__ void out_conv_wgrad_reduce_kernel(const float* __restrict__ partial, int blocks, float* __restrict__ dw, const float* __restrict__ gs) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 49) return;
  float s = 0.f;
  for (int k = 0; k < blocks; ++k) s += partial[(long long)k * (64 * 49) + i];
  dw[i] = gs != nullptr ? s * gs[1] : s;
}

// ------------------------------------------------------------------ stem backward (modules/model.py:90-92)
// dcol [B][H][W][64] 16-bit holds, in column k = r*7 + s (k < 49), the gradient of the im2col entry (pixel, channel 0,
// tap (r,s)); the image gradient sums the entries that read the same padded pixel, then folds the reflection.
template <typename T>
__global__ void __launch_bounds__(256)
stem_col2im_kernel(const T* __restrict__ dcol, float* __restrict__ dx, const float* __restrict__ gs, int B, int H, int W) {
  pdl_prologue();
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
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = Cin * 49;
  if (i >= 64 * K) return;
  g[i] = packed[(i / K) * Kpad + (i % K)] * gs[1];
}


// ------------------------------------------------------------------ CBAM backward (modules/model.py:12-53, 68-87)
// forward (per residual block, n = InstanceNorm(conv output), C = 256):
//   mx[c] = max_{y,x} n;  h = relu(fc0 mx);  ca = sigmoid(fc2 h)      (the avg-pool branch sees exactly 0, see elementwise.cu)
//   v = n * ca;  pooled = [mean_c v, max_c v];  sa = sigmoid(conv7x7(pooled));  out = x + v * sa
// The avg-pool branch contributes nothing to the backward either: its hidden activations are relu(0) = 0 (no fc2
// gradient), relu'(0) = 0 (no fc0 gradient), and a per-channel constant added to dn is removed by the InstanceNorm
// backward that follows.
constexpr int kCbamC = 256;       // channels of the residual blocks (one warp = one pixel, 8 channels per lane)
constexpr int kCbamPix = 64;      // pixels per CTA in the warp-per-pixel passes

template <typename T>
__device__ __forceinline__ void load8(const T* p, float* f) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = Cvt<T>::unpack2(w[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float* f) {
  uint4 o;
  o.x = Cvt<T>::pack2(f[0], f[1]);
  o.y = Cvt<T>::pack2(f[2], f[3]);
  o.z = Cvt<T>::pack2(f[4], f[5]);
  o.w = Cvt<T>::pack2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// training-mode channel attention: keeps ca, the hidden layer and the un-folded InstanceNorm affine
__global__ void __launch_bounds__(256)
cbam_channel_train_kernel(const float* __restrict__ chmax, const float* __restrict__ fc0, const float* __restrict__ fc2,
                          const float* __restrict__ scale_n, const float* __restrict__ shift_n, float* __restrict__ scale_v,
                          float* __restrict__ shift_v, float* __restrict__ ca, float* __restrict__ hidden, int C) {
  pdl_prologue();
  extern __shared__ float sm[];  // [C] max, [C/16] hidden
  float* smax = sm;
  float* hid = sm + C;
  const int b = blockIdx.x, Hd = C / 16;
  for (int c = threadIdx.x; c < C; c += blockDim.x) smax[c] = chmax[b * C + c];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < Hd; j += nwarps) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc += fc0[j * C + c] * smax[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      hid[j] = fmaxf(acc, 0.f);
      hidden[b * Hd + j] = hid[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < Hd; ++j) acc += fc2[c * Hd + j] * hid[j];
    const float s = 1.f / (1.f + __expf(-acc));
    ca[b * C + c] = s;
    scale_v[b * C + c] = scale_n[b * C + c] * s;
    shift_v[b * C + c] = shift_n[b * C + c] * s;
  }
}

// pass A: dz = (sum_c dout * v) * sa * (1 - sa); per-CTA arg-max over the pixels of every channel of the raw conv output
template <typename T>
__global__ void __launch_bounds__(256)
cbam_bwd_dz_kernel(const T* __restrict__ dout, const T* __restrict__ yb, const float* __restrict__ scale_v,
                   const float* __restrict__ shift_v, const float* __restrict__ sa, float* __restrict__ dz,
                   float* __restrict__ pmax_val, int* __restrict__ pmax_idx, int HW) {
  pdl_prologue();
  __shared__ float sval[8][kCbamC];
  __shared__ int sidx[8][kCbamC];
  const int b = blockIdx.y, blk = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c0 = lane * 8;
  float sv[8], hv[8], best[8];
  int bidx[8];
  load8f(scale_v + b * kCbamC + c0, sv);
  load8f(shift_v + b * kCbamC + c0, hv);
#pragma unroll
  for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; bidx[k] = 0; }
#pragma unroll 4   // four pixels' loads in flight per warp (a warp-per-pixel loop with a shuffle reduction is latency bound otherwise)
  for (int i = 0; i < kCbamPix / 8; ++i) {
    const int pix = blk * kCbamPix + warp + i * 8;
    const long long off = ((long long)b * HW + pix) * kCbamC + c0;
    float d[8], y[8];
    load8<T>(dout + off, d);
    load8<T>(yb + off, y);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc += d[k] * fmaf(y[k], sv[k], hv[k]);
      if (y[k] > best[k]) { best[k] = y[k]; bidx[k] = pix; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float s = sa[(long long)b * HW + pix];
      dz[(long long)b * HW + pix] = acc * s * (1.f - s);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { sval[warp][c0 + k] = best[k]; sidx[warp][c0 + k] = bidx[k]; }
  __syncthreads();
  const int c = threadIdx.x;
  float bv = sval[0][c];
  int bi = sidx[0][c];
  for (int w = 1; w < 8; ++w) {
    const float v = sval[w][c];
    const int i = sidx[w][c];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
  pmax_val[((long long)b * gridDim.x + blk) * kCbamC + c] = bv;
  pmax_idx[((long long)b * gridDim.x + blk) * kCbamC + c] = bi;
}

// one warp per (sample, channel): lanes scan the CTA partials, ties resolve to the smallest pixel index (first occurrence)
__global__ void __launch_bounds__(256)
cbam_argmax_finalize_kernel(const float* __restrict__ pmax_val, const int* __restrict__ pmax_idx, int nblk,
                            int* __restrict__ amax_pix) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp, b = blockIdx.y;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int k = lane; k < nblk; k += 32) {
    const float v = pmax_val[((long long)b * nblk + k) * kCbamC + c];
    const int i = pmax_idx[((long long)b * nblk + k) * kCbamC + c];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) amax_pix[b * kCbamC + c] = bi;
}

// pass B: adjoint of the 2 -> 1 7x7 spatial-attention conv: dpooled [B][H][W][2] and per-CTA partial weight gradients
__global__ void __launch_bounds__(256)
cbam_sa_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ pooled, const float* __restrict__ wsa,
                   float* __restrict__ dpooled, float* __restrict__ pdw, int B, int H, int W) {
  pdl_prologue();
  __shared__ float ws[98];
  __shared__ float red[8][98];
  if (threadIdx.x < 98) ws[threadIdx.x] = wsa[threadIdx.x];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;   // B*H*W is a multiple of 256
  const int x = int(pix % W), y = int((pix / W) % H), b = int(pix / ((long long)W * H));
  const float* dzb = dz + (long long)b * H * W;
  const float* pb = pooled + (long long)b * H * W * 2;
  const float g = dz[pix];
  float d0 = 0.f, d1 = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < 7; ++r)
    for (int s = 0; s < 7; ++s) {
      // dpooled[y][x] += dz[y + 3 - r][x + 3 - s] * w[ch][r][s]
      const int yy = y + 3 - r, xx = x + 3 - s;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float t = dzb[(long long)yy * W + xx];
        d0 += t * ws[r * 7 + s];
        d1 += t * ws[49 + r * 7 + s];
      }
      // dw[ch][r][s] += dz[y][x] * pooled[ch][y + r - 3][x + s - 3]
      const int py = y + r - 3, px = x + s - 3;
      float p0 = 0.f, p1 = 0.f;
      if (py >= 0 && py < H && px >= 0 && px < W) {
        const float2 pv = *reinterpret_cast<const float2*>(pb + ((long long)py * W + px) * 2);
        p0 = g * pv.x;
        p1 = g * pv.y;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        p0 += __shfl_xor_sync(0xffffffffu, p0, o);
        p1 += __shfl_xor_sync(0xffffffffu, p1, o);
      }
      if (lane == 0) { red[warp][r * 7 + s] = p0; red[warp][49 + r * 7 + s] = p1; }
    }
  *reinterpret_cast<float2*>(dpooled + pix * 2) = make_float2(d0, d1);
  __syncthreads();
  if (threadIdx.x < 98) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
    pdw[(long long)blockIdx.x * 98 + threadIdx.x] = a;
  }
}

// pass C: dv = dout*sa + dpooled0/C + [c == argmax_c v] dpooled1;  dn = dv*ca (16 bit);  per-CTA partial dca = sum dv*n
template <typename T>
__global__ void __launch_bounds__(256)
cbam_bwd_dv_kernel(const T* __restrict__ dout, const T* __restrict__ yb, const float* __restrict__ scale_n,
                   const float* __restrict__ shift_n, const float* __restrict__ ca, const float* __restrict__ sa,
                   const float* __restrict__ dpooled, T* __restrict__ dn, float* __restrict__ pdca, int HW) {
  pdl_prologue();
  __shared__ float sacc[8][kCbamC];
  const int b = blockIdx.y, blk = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c0 = lane * 8;
  float sn[8], hn[8], cav[8], accd[8];
  load8f(scale_n + b * kCbamC + c0, sn);
  load8f(shift_n + b * kCbamC + c0, hn);
  load8f(ca + b * kCbamC + c0, cav);
#pragma unroll
  for (int k = 0; k < 8; ++k) accd[k] = 0.f;
#pragma unroll 4   // four pixels' loads in flight per warp (a warp-per-pixel loop with a shuffle reduction is latency bound otherwise)
  for (int i = 0; i < kCbamPix / 8; ++i) {
    const int pix = blk * kCbamPix + warp + i * 8;
    const long long off = ((long long)b * HW + pix) * kCbamC + c0;
    float d[8], y[8], n[8];
    load8<T>(dout + off, d);
    load8<T>(yb + off, y);
    float vmax = -INFINITY;
    int cmax = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      n[k] = fmaf(y[k], sn[k], hn[k]);
      const float v = n[k] * cav[k];
      if (v > vmax) { vmax = v; cmax = c0 + k; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, vmax, o);
      const int oc = __shfl_xor_sync(0xffffffffu, cmax, o);
      if (ov > vmax || (ov == vmax && oc < cmax)) { vmax = ov; cmax = oc; }
    }
    const float s = sa[(long long)b * HW + pix];
    const float2 dp = *reinterpret_cast<const float2*>(dpooled + ((long long)b * HW + pix) * 2);
    const float dmean = dp.x * (1.f / kCbamC);
    float o8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float dv = d[k] * s + dmean + ((c0 + k) == cmax ? dp.y : 0.f);
      accd[k] += dv * n[k];
      o8[k] = dv * cav[k];
    }
    store8<T>(dn + off, o8);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sacc[warp][c0 + k] = accd[k];
  __syncthreads();
  float a = 0.f;
  for (int w = 0; w < 8; ++w) a += sacc[w][threadIdx.x];
  pdca[((long long)b * gridDim.x + blk) * kCbamC + threadIdx.x] = a;
}

// pass D (one CTA per sample): channel-attention MLP backward, per-sample fc gradients, and the max-pool scatter
template <typename T>
__global__ void __launch_bounds__(kCbamC)
cbam_channel_bwd_kernel(const float* __restrict__ pdca, int nblk, const float* __restrict__ ca, const float* __restrict__ hidden,
                        const float* __restrict__ chmax, const float* __restrict__ fc0, const float* __restrict__ fc2,
                        const int* __restrict__ amax_pix, T* __restrict__ dn, float* __restrict__ pfc0,
                        float* __restrict__ pfc2, int HW) {
  pdl_prologue();
  constexpr int Hd = kCbamC / 16;
  __shared__ float ds[kCbamC], h[Hd], dh[Hd];
  const int b = blockIdx.x, c = threadIdx.x;
  // nblk partial rows per sample (one per CTA of the dv pass): eight independent chains keep the loads in flight -- with a
  // single chain this one-CTA-per-sample kernel took 19 us at one sample per rank; fixed order, deterministic
  float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* prow = pdca + (long long)b * nblk * kCbamC + c;
  int k0 = 0;
  for (; k0 + 8 <= nblk; k0 += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) part[u] += prow[(long long)(k0 + u) * kCbamC];
  }
  for (; k0 < nblk; ++k0) part[k0 & 7] += prow[(long long)k0 * kCbamC];
  const float dca = ((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]));
  const float cv = ca[b * kCbamC + c];
  ds[c] = dca * cv * (1.f - cv);
  if (c < Hd) h[c] = hidden[b * Hd + c];
  __syncthreads();
  if (c < Hd) {
    float a = 0.f;
    if (h[c] > 0.f)
      for (int k = 0; k < kCbamC; ++k) a += fc2[k * Hd + c] * ds[k];
    dh[c] = a;
  }
  __syncthreads();
  float dmx = 0.f;
  const float mxc = chmax[b * kCbamC + c];
#pragma unroll
  for (int j = 0; j < Hd; ++j) {
    dmx += fc0[j * kCbamC + c] * dh[j];
    pfc2[((long long)b * kCbamC + c) * Hd + j] = ds[c] * h[j];
    pfc0[((long long)b * Hd + j) * kCbamC + c] = dh[j] * mxc;
  }
  T* cell = dn + ((long long)b * HW + amax_pix[b * kCbamC + c]) * kCbamC + c;
  *cell = Cvt<T>::from_f(Cvt<T>::to_f(*cell) + dmx);
}

// the three CBAM parameter gradients of a block in ONE launch: fixed-order sums of the per-sample / per-CTA partials, true scale
// (gs[1]); accumulate != 0: added to what the destinations hold (the parameters' existing .grad) instead of overwriting
__global__ void cbam_param_reduce3_kernel(const float* __restrict__ pfc0, const float* __restrict__ pfc2, int parts_fc, int nfc,
                                          const float* __restrict__ pdw, int parts_dw, float* __restrict__ dfc0,
                                          float* __restrict__ dfc2, float* __restrict__ dwsa, const float* __restrict__ gs,
                                          int accumulate) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* part;
  float* out;
  int parts, n;
  if (i < nfc) { part = pfc0; out = dfc0; parts = parts_fc; n = nfc; }
  else if (i < 2 * nfc) { i -= nfc; part = pfc2; out = dfc2; parts = parts_fc; n = nfc; }
  else if (i < 2 * nfc + 98) { i -= 2 * nfc; part = pdw; out = dwsa; parts = parts_dw; n = 98; }
  else return;
  float a = 0.f;
  for (int k = 0; k < parts; ++k) a += part[(long long)k * n + i];
  a *= gs[1];
  out[i] = accumulate ? out[i] + a : a;
}

// Adam (torch.optim.Adam semantics, no weight decay / amsgrad; modules/trainer.py:360-362): one fused pass over
// (param, grad, exp_avg, exp_avg_sq), all fp32.  step_size = lr / (1 - b1^t), denom = sqrt(v) / sqrt(1 - b2^t) + eps.
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                 float b1, float b2, float eps, float step_size, float inv_sqrt_bc2) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

// Same update with the learning rate and the step count read from device memory (state = {lr, step}), so that a whole
// optimisation step can live in a CUDA graph: nothing that changes between replays is baked into the launch.
__global__ void adam_advance_kernel(float* __restrict__ state) {
  pdl_prologue(); state[1] += 1.f; }
__global__ void __launch_bounds__(256)
adam_step_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                     float b1, float b2, float eps, const float* __restrict__ state) {
  pdl_prologue();
  __shared__ float s_step_size, s_inv_sqrt_bc2;
  if (threadIdx.x == 0) {
    const double t = double(state[1]);
    const double bc1 = 1.0 - pow(double(b1), t), bc2 = 1.0 - pow(double(b2), t);
    s_step_size = float(double(state[0]) / bc1);
    s_inv_sqrt_bc2 = float(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_step_size, inv_sqrt_bc2 = s_inv_sqrt_bc2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

// a += b on 16-bit maps (the skip connection of the residual blocks carries the gradient straight through)
template <typename T>
__global__ void __launch_bounds__(256)
add_inplace_kernel(T* __restrict__ a, const T* __restrict__ b, long long n8) {
  pdl_prologue();
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
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(add_inplace_kernel<T>, grid_items(n / 8, 256), 256, 0, (cudaStream_t)stream)(
                                      static_cast<T*>(a), static_cast<const T*>(b), n / 8)));
  return check_launch("add_inplace_kernel");
}

namespace {
constexpr int kTanhBlocks = 592;
constexpr int kOutWgradBlocks = 592;
}

extern "C" size_t ducosy_out_conv_backward_scratch_bytes(int B, int H, int W) {
  // dv fp32 + partial sums + the scaled 16-bit copy of dv (rounded up to whole floats)
  return (size_t(B) * H * W + kTanhBlocks + size_t(kOutWgradBlocks) * 64 * 49 + (size_t(B) * H * W + 1) / 2 + 4) * sizeof(float);
}

extern "C" int ducosy_out_conv_backward(const float* dout, const float* out, const void* in_pad, const float* w, void* da,
                                        float* dw, float* db, float* scratch, const float* gs, int B, int H, int W, int dtype,
                                        ducosy_stream_t stream) {
  DUCOSY_CHECK(dout && out && in_pad && w && da && dw && db && scratch && gs, DUCOSY_ERR_ARG, "out_conv_backward: null pointer");
  DUCOSY_CHECK(B > 0 && H >= 8 && W >= 8, DUCOSY_ERR_SHAPE, "out_conv_backward: needs B > 0 and H, W >= 8 (got %d x %d x %d)", B, H, W);
  DUCOSY_CHECK(W % 4 == 0, DUCOSY_ERR_SHAPE, "out_conv_backward: W must be a multiple of 4 (got %d)", W);
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "out_conv_backward: bad dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * H * W;
  float* dv = scratch;
  float* tpart = dv + n;
  float* wpart = tpart + kTanhBlocks;
  // the tensor-core input gradient needs whole 4 x 64 tiles; other shapes (tests) stay on the CUDA-core kernel
  static const bool mma_env = []() { const char* e = getenv("DUCOSY_OUTCONV_DGRAD_MMA"); return e == nullptr || atoi(e) != 0; }();
  const bool use_mma = mma_env && H % kDgRows == 0 && W % kDgCols == 0;
  void* dvh = wpart + size_t(kOutWgradBlocks) * 64 * 49;
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(out_tanh_bwd_kernel<T>, kTanhBlocks, 256, 0, st)(dout, out, dv, use_mma ? static_cast<T*>(dvh) : nullptr,
                                                                                       gs, tpart, n)));
  DUCOSY_TRY(check_launch("out_tanh_bwd_kernel"));
  pdl(sum_partials_kernel, 1, 256, 0, st)(tpart, kTanhBlocks, db);
  DUCOSY_TRY(check_launch("sum_partials_kernel"));
  if (use_mma) {
    const int tiles = B * (H / kDgRows) * (W / kDgCols), cap = (num_sms() > 0 ? num_sms() : 148) * 3;   // 138 registers: 3 CTAs per SM
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(out_conv_dgrad_mma_kernel<T>, tiles < cap ? tiles : cap, 128, 0, st)(
                                        static_cast<const T*>(dvh), w, static_cast<T*>(da), B, H, W)));
    DUCOSY_TRY(check_launch("out_conv_dgrad_mma_kernel"));
  }
  if (use_mma) {
    const long long items = (long long)B * (8 * W + 8 * (H - 8)) * 8;
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(out_conv_dgrad_border_kernel<T>, grid_items(items, 256), 256, 0, st)(dv, w, static_cast<T*>(da), gs, B, H, W)));
    DUCOSY_TRY(check_launch("out_conv_dgrad_border_kernel"));
  } else {
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(out_conv_dgrad_kernel<T>, grid_items(n * 2, 256), 256, 0, st)(dv, w, static_cast<T*>(da), gs, B, H, W)));
    DUCOSY_TRY(check_launch("out_conv_dgrad_kernel"));
  }
  static const bool wmma_env = []() { const char* e = getenv("DUCOSY_OUTCONV_WGRAD_MMA"); return e == nullptr || atoi(e) != 0; }();
  if (use_mma && wmma_env) {
    const int pitch = ((W + 12 + 16 + 7) / 8) * 8;
    const size_t smem_mma = size_t(2) * 7 * pitch * 2 + size_t(4) * 2 * 16 * kWgA * 2;
    DUCOSY_CHECK(smem_mma <= 48 * 1024 && smem_mma >= size_t(64) * 56 * 4, DUCOSY_ERR_SHAPE, "out_conv_backward: W unsupported (%d)", W);
    const int blocks = min((num_sms() > 0 ? num_sms() : 148) * 3, min(kOutWgradBlocks, B * (H + 6)));
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(out_conv_wgrad_mma_kernel<T>, blocks, 128, smem_mma, st)(static_cast<const T*>(in_pad), static_cast<const T*>(dvh),
                                                                                             wpart, B, H, W)));
    DUCOSY_TRY(check_launch("out_conv_wgrad_mma_kernel"));
    pdl(out_conv_wgrad_reduce_kernel, (64 * 49 + 255) / 256, 256, 0, st)(wpart, blocks, dw, gs);
    return check_launch("out_conv_wgrad_reduce_kernel");
  }
  const int blocks = min(kOutWgradBlocks, B * (H + 6));
  const size_t smem = size_t(7) * (W + 12) * sizeof(float);
  DUCOSY_CHECK(smem <= 48 * 1024, DUCOSY_ERR_SHAPE, "out_conv_backward: W too large (%d)", W);
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(out_conv_wgrad_kernel<T>, blocks, kOutWgradThreads, smem, st)(static_cast<const T*>(in_pad), dv, wpart, B, H, W)));
  DUCOSY_TRY(check_launch("out_conv_wgrad_kernel"));
  pdl(out_conv_wgrad_reduce_kernel, (64 * 49 + 255) / 256, 256, 0, st)(wpart, blocks, dw, static_cast<const float*>(nullptr));
  return check_launch("out_conv_wgrad_reduce_kernel");
}

extern "C" int ducosy_stem_col2im(const void* dcol, float* dx, const float* gs, int B, int H, int W, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(dcol && dx && gs && B > 0 && H >= 8 && W >= 8, DUCOSY_ERR_ARG, "stem_col2im: bad argument");
  const long long n = (long long)B * H * W;
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(stem_col2im_kernel<T>, grid_items(n, 256), 256, 0, (cudaStream_t)stream)(
                                      static_cast<const T*>(dcol), dx, gs, B, H, W)));
  return check_launch("stem_col2im_kernel");
}

extern "C" int ducosy_unpack_stem_wgrad(const float* packed, float* g_oihw, int Cin, int Kpad, const float* gs, ducosy_stream_t stream) {
  DUCOSY_CHECK(packed && g_oihw && gs && Cin >= 1 && Kpad >= Cin * 49, DUCOSY_ERR_ARG, "unpack_stem_wgrad: bad argument");
  pdl(unpack_stem_wgrad_kernel, grid_items(64LL * Cin * 49, 256), 256, 0, (cudaStream_t)stream)(packed, g_oihw, Cin, Kpad, gs);
  return check_launch("unpack_stem_wgrad_kernel");
}

extern "C" int ducosy_cbam_channel_train(const float* chmax, const float* fc0, const float* fc2, const float* scale_n,
                                         const float* shift_n, float* scale_v, float* shift_v, float* ca, float* hidden, int B,
                                         int C, ducosy_stream_t stream) {
  DUCOSY_CHECK(chmax && fc0 && fc2 && scale_n && shift_n && scale_v && shift_v && ca && hidden, DUCOSY_ERR_ARG,
               "cbam_channel_train: null pointer");
  DUCOSY_CHECK(B > 0 && C % 32 == 0, DUCOSY_ERR_SHAPE, "cbam_channel_train: C %% 32 != 0");
  const size_t smem = size_t(C + C / 16) * sizeof(float);
  pdl(cbam_channel_train_kernel, B, 256, smem, (cudaStream_t)stream)(chmax, fc0, fc2, scale_n, shift_n, scale_v, shift_v, ca, hidden, C);
  return check_launch("cbam_channel_train_kernel");
}

extern "C" size_t ducosy_cbam_backward_scratch_bytes(int B, int H, int W, int C) {
  const size_t HW = size_t(H) * W, nblk = HW / kCbamPix;
  size_t words = B * HW * 3                 /* dz, dpooled */
                 + 3 * B * nblk * C        /* pmax_val, pmax_idx, pdca */
                 + size_t(B) * C           /* amax_pix */
                 + (B * HW / 256) * 98     /* spatial conv weight partials */
                 + 2 * size_t(B) * C * (C / 16);
  return words * 4;
}

namespace {
int cbam_backward_impl(const void* dout, const void* yb, const float* scale_n, const float* shift_n,
                       const float* scale_v, const float* shift_v, const float* ca, const float* hidden,
                       const float* chmax, const float* pooled, const float* sa, const float* fc0, const float* fc2,
                       const float* wsa, void* dn, float* dfc0, float* dfc2, float* dwsa, float* scratch,
                       const float* gs, int B, int H, int W, int C, int dtype, ducosy_stream_t stream, int accumulate) {
  DUCOSY_CHECK(dout && yb && scale_n && shift_n && scale_v && shift_v && ca && hidden && chmax && pooled && sa && fc0 && fc2 &&
                   wsa && dn && dfc0 && dfc2 && dwsa && scratch && gs, DUCOSY_ERR_ARG, "cbam_backward: null pointer");
  DUCOSY_CHECK(C == kCbamC, DUCOSY_ERR_SHAPE, "cbam_backward: built for the %d-channel residual blocks (got %d)", kCbamC, C);
  DUCOSY_CHECK(B > 0 && (H * W) % 256 == 0, DUCOSY_ERR_SHAPE, "cbam_backward: H*W must be a multiple of 256");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "cbam_backward: bad dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int HW = H * W, nblk = HW / kCbamPix, sablk = B * HW / 256, Hd = C / 16;
  float* dz = scratch;
  float* dpooled = dz + size_t(B) * HW;
  float* pmax_val = dpooled + size_t(B) * HW * 2;
  int* pmax_idx = reinterpret_cast<int*>(pmax_val + size_t(B) * nblk * C);
  float* pdca = reinterpret_cast<float*>(pmax_idx + size_t(B) * nblk * C);
  int* amax_pix = reinterpret_cast<int*>(pdca + size_t(B) * nblk * C);
  float* pdw = reinterpret_cast<float*>(amax_pix + size_t(B) * C);
  float* pfc0 = pdw + size_t(sablk) * 98;
  float* pfc2 = pfc0 + size_t(B) * C * Hd;
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(cbam_bwd_dz_kernel<T>, dim3(nblk, B), 256, 0, st)(
                                      static_cast<const T*>(dout), static_cast<const T*>(yb), scale_v, shift_v, sa, dz, pmax_val,
                                      pmax_idx, HW)));
  DUCOSY_TRY(check_launch("cbam_bwd_dz_kernel"));
  pdl(cbam_argmax_finalize_kernel, dim3(kCbamC / 8, B), 256, 0, st)(pmax_val, pmax_idx, nblk, amax_pix);
  DUCOSY_TRY(check_launch("cbam_argmax_finalize_kernel"));
  pdl(cbam_sa_bwd_kernel, sablk, 256, 0, st)(dz, pooled, wsa, dpooled, pdw, B, H, W);
  DUCOSY_TRY(check_launch("cbam_sa_bwd_kernel"));
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(cbam_bwd_dv_kernel<T>, dim3(nblk, B), 256, 0, st)(
                                      static_cast<const T*>(dout), static_cast<const T*>(yb), scale_n, shift_n, ca, sa, dpooled,
                                      static_cast<T*>(dn), pdca, HW)));
  DUCOSY_TRY(check_launch("cbam_bwd_dv_kernel"));
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(cbam_channel_bwd_kernel<T>, B, kCbamC, 0, st)(pdca, nblk, ca, hidden, chmax, fc0, fc2, amax_pix,
                                                                                  static_cast<T*>(dn), pfc0, pfc2, HW)));
  DUCOSY_TRY(check_launch("cbam_channel_bwd_kernel"));
  const int nfc = C * Hd;
  pdl(cbam_param_reduce3_kernel, (2 * nfc + 98 + 255) / 256, 256, 0, st)(pfc0, pfc2, B, nfc, pdw, sablk, dfc0, dfc2, dwsa, gs, accumulate);
  return check_launch("cbam_param_reduce3_kernel");
}
}  // namespace

extern "C" int ducosy_cbam_backward(const void* dout, const void* yb, const float* scale_n, const float* shift_n,
                                    const float* scale_v, const float* shift_v, const float* ca, const float* hidden,
                                    const float* chmax, const float* pooled, const float* sa, const float* fc0, const float* fc2,
                                    const float* wsa, void* dn, float* dfc0, float* dfc2, float* dwsa, float* scratch,
                                    const float* gs, int B, int H, int W, int C, int dtype, ducosy_stream_t stream) {
  return cbam_backward_impl(dout, yb, scale_n, shift_n, scale_v, shift_v, ca, hidden, chmax, pooled, sa, fc0, fc2, wsa, dn, dfc0, dfc2,
                            dwsa, scratch, gs, B, H, W, C, dtype, stream, 0);
}
// The same with the three parameter gradients ACCUMULATED into dfc0 / dfc2 / dwsa (+=): the destinations are the parameters'
// existing .grad tensors.
extern "C" int ducosy_cbam_backward_acc(const void* dout, const void* yb, const float* scale_n, const float* shift_n,
                                        const float* scale_v, const float* shift_v, const float* ca, const float* hidden,
                                        const float* chmax, const float* pooled, const float* sa, const float* fc0, const float* fc2,
                                        const float* wsa, void* dn, float* dfc0, float* dfc2, float* dwsa, float* scratch,
                                        const float* gs, int B, int H, int W, int C, int dtype, ducosy_stream_t stream) {
  return cbam_backward_impl(dout, yb, scale_n, shift_n, scale_v, shift_v, ca, hidden, chmax, pooled, sa, fc0, fc2, wsa, dn, dfc0, dfc2,
                            dwsa, scratch, gs, B, H, W, C, dtype, stream, 1);
}

extern "C" int ducosy_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                                float beta1, float beta2, float eps, int step, ducosy_stream_t stream) {
  DUCOSY_CHECK(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, DUCOSY_ERR_ARG, "adam_step: bad argument");
  const double bc1 = 1.0 - pow(double(beta1), step), bc2 = 1.0 - pow(double(beta2), step);
  const int blocks = int(std::min<long long>((n + 255) / 256, 148 * 8));
  pdl(adam_step_kernel, blocks, 256, 0, (cudaStream_t)stream)(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                            float(double(lr) / bc1), float(1.0 / sqrt(bc2)));
  return check_launch("adam_step_kernel");
}

extern "C" int ducosy_adam_advance(float* state, ducosy_stream_t stream) {
  DUCOSY_CHECK(state != nullptr, DUCOSY_ERR_ARG, "adam_advance: null pointer");
  pdl(adam_advance_kernel, 1, 1, 0, (cudaStream_t)stream)(state);
  return check_launch("adam_advance_kernel");
}

extern "C" int ducosy_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                    const float* state, float beta1, float beta2, float eps, ducosy_stream_t stream) {
  DUCOSY_CHECK(param && grad && exp_avg && exp_avg_sq && state && n > 0, DUCOSY_ERR_ARG, "adam_step_dev: bad argument");
  const int blocks = int(std::min<long long>((n + 255) / 256, 148 * 8));
  pdl(adam_step_dev_kernel, blocks, 256, 0, (cudaStream_t)stream)(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, state);
  return check_launch("adam_step_dev_kernel");
}
