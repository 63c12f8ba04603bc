// Fused stem for input_channels == 1 (generate.py:29-30 builds exactly this generator):
//   [HU window ->] ReflectionPad2d(3) -> Conv2d(1, 64, 7) -> InstanceNorm2d(64) -> ReLU -> zero pad 1 (for the next conv)
//   modules/preprocess.py:72-84, modules/model.py:94
// The 7x7x1 convolution has only K = 49 (1.6 GFLOP/slice) but a 64-channel 512x512 output (33.5 MB): moving an
// im2col matrix and the raw output through HBM costs 4-5x more than the math.  InstanceNorm needs the statistics of the
// whole map before anything can be normalised, so the convolution is simply evaluated TWICE from a 7-row shared-memory
// patch of the (windowed) input:
//   pass 1  stem_fused_kernel<.., false>: conv -> per-row per-channel (sum, sum of squares) only, nothing stored;
//   pass 2  stem_fused_kernel<.., true> : conv again -> (y*scale + shift), ReLU -> 16-bit NHWC, written once, already
//           zero-padded for the stride-2 conv that follows.
// Tensor work: tcgen05.mma with the im2col A tiles (128 pixels x K = 49 -> 64) assembled in shared memory by the CTA
// itself in the 128-byte-swizzle K-major layout the UMMA descriptor expects; the 64 x 64 weight tile stays resident.
// (A first version used warp-level mma.sync: 430 us per pass for 10 slices -- the legacy tensor path is far too slow.)
#include <type_traits>

#include "common.cuh"
#include "input_fn.cuh"
#include "ptx.cuh"

namespace ducosy {
namespace {

constexpr int kStemThreads = 160;   // warps 0-3: A builders, then epilogue (TMEM lane quarters); warp 4: TMEM + MMA issue
constexpr int kTilesPerCta = 4;     // 4 x 128 pixels of one image row per round; 4 x 64 TMEM columns
constexpr int kATile = 128 * 128;   // bytes: 128 rows x 64 x 16-bit

// transposing butterfly over a warp: lane L ends with the sum over lanes of x[L]
template <int H>
__device__ __forceinline__ void bfly_step(float (&x)[32], int lane) {
  const bool up = (lane & H) != 0;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float send = up ? x[i] : x[i + H];
    const float keep = up ? x[i + H] : x[i];
    x[i] = keep + __shfl_xor_sync(0xffffffffu, send, H);
  }
}
__device__ __forceinline__ float bfly_sum(float (&x)[32], int lane) {
  bfly_step<16>(x, lane);
  bfly_step<8>(x, lane);
  bfly_step<4>(x, lane);
  bfly_step<2>(x, lane);
  bfly_step<1>(x, lane);
  return x[0];
}

// Input preparation (once per forward): fp32 tensor or stored pixels through the HU window -> 16-bit [B][H][W].
template <typename T, typename In>
__global__ void stem_input_kernel(In in, T* __restrict__ xw, int B, int H, int W) {
  pdl_prologue();
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % W);
    const long long r = i / W;
    xw[i] = Cvt<T>::from_f(in.at(int(r / H), 0, int(r % H), x, 1, H, W));
  }
}

// grid (H, B): one output row per CTA.  xw = prepared input, wp = packed stem weight [64][64] (k = r*7 + s, 0 for k >= 49).
template <typename T, bool kApply>
__global__ void __launch_bounds__(kStemThreads, 2)
stem_fused_kernel(const T* __restrict__ xw, const T* __restrict__ wp, float* __restrict__ partials,
                  const float* __restrict__ scale, const float* __restrict__ shift, T* __restrict__ out_pad, int H, int W) {
  constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  extern __shared__ uint8_t stem_raw[];
  const uint32_t raw_addr = smem_u32(stem_raw);
  uint8_t* smem = stem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* a_tiles = smem;                                      // kTilesPerCta x 16 KiB (reused as store staging)
  uint8_t* b_tile = smem + kTilesPerCta * kATile;               // 64 rows x 128 B
  float2* coef = reinterpret_cast<float2*>(b_tile + 64 * 128);  // [64] (scale, shift)          (pass 2)
  float* red = reinterpret_cast<float*>(coef + 64);             // [4 warps][2][64]             (pass 1)
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 4 * 2 * 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  // 7 reflected input rows; row pitch W + 16 halfwords with the interior at offset 8 (16-byte aligned vector copies),
  // the 3 + 3 reflected border pixels at offsets 5..7 and 8+W..10+W
  unsigned short* patch = reinterpret_cast<unsigned short*>(smem + kTilesPerCta * kATile + 64 * 128 + 64 * 8 + 4 * 2 * 64 * 4 + 32);

  const int Wp = W + 16;
  const int y = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();   // programmatic dependent launch (common.cuh): the prepared input / statistics come from the previous kernel

  {
    const unsigned short* src = reinterpret_cast<const unsigned short*>(xw) + size_t(b) * H * W;
    const int vec_per_row = W / 8;
    for (int i = threadIdx.x; i < 7 * vec_per_row; i += kStemThreads) {
      const int r = i / vec_per_row, j = i - r * vec_per_row;
      const uint4 v = *reinterpret_cast<const uint4*>(src + size_t(reflect_idx(y + r - 3, H)) * W + j * 8);
      *reinterpret_cast<uint4*>(patch + r * Wp + 8 + j * 8) = v;
    }
    if (threadIdx.x < 42) {  // reflected borders: columns -3..-1 and W..W+2
      const int r = threadIdx.x / 6, e = threadIdx.x % 6;
      const int col = e < 3 ? e - 3 : W + e - 3;
      patch[r * Wp + 8 + col] = src[size_t(reflect_idx(y + r - 3, H)) * W + reflect_idx(col, W)];
    }
  }
  for (int ch = threadIdx.x; ch < 64 * 8; ch += kStemThreads) {  // weights -> swizzled K-major B tile
    const int n = ch >> 3, c = ch & 7;
    *reinterpret_cast<uint4*>(b_tile + n * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<const uint4*>(wp + n * 64 + c * 8);
  }
  if (kApply && threadIdx.x < 64) coef[threadIdx.x] = make_float2(scale[b * 64 + threadIdx.x], shift[b * 64 + threadIdx.x]);
  if (warp == 4) {
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
    }
    tmem_alloc(tmem_slot, kTilesPerCta * 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // only after the TMEM allocation (see common.cuh)

  float s1[64], s2[64];  // pass 1 only: this thread's pixels, all 64 channels
  if (!kApply) {
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      s1[i] = 0.f;
      s2[i] = 0.f;
    }
  }
  uint32_t phase = 0;
  for (int t0 = 0; t0 < W / 128; t0 += kTilesPerCta) {
    const int nt = min(kTilesPerCta, W / 128 - t0);
    if (warp < 4) {
      // ---- build the im2col rows: thread m owns row m of every tile (pixel x = (t0+tl)*128 + m)
      const int m = threadIdx.x;
      for (int tl = 0; tl < nt; ++tl) {
        const unsigned short* p0 = patch + 5 + (t0 + tl) * 128 + m;
        uint8_t* row = a_tiles + tl * kATile + m * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k0 = c * 8 + 2 * j, k1 = k0 + 1;
            const uint32_t lo = k0 < 49 ? p0[(k0 / 7) * Wp + (k0 % 7)] : 0u;
            const uint32_t hi = k1 < 49 ? p0[(k1 / 7) * Wp + (k1 % 7)] : 0u;
            w[j] = lo | (hi << 16);
          }
          *reinterpret_cast<uint4*>(row + ((c ^ (m & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    }
    __syncthreads();
    if (warp == 4) {
      tc_fence_after();
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(kFmt, 128, 64);
        const uint64_t db = umma_desc_k_sw128(smem_u32(b_tile));
        for (int tl = 0; tl < nt; ++tl) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(a_tiles + tl * kATile));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_f16(tmem_base + uint32_t(tl * 64), da + uint64_t(j * 2), db + uint64_t(j * 2), idesc, j != 0 ? 1u : 0u);
        }
        umma_commit(bar);
      }
      __syncwarp();
    } else {
      // ---- epilogue: thread = pixel (TMEM lane), 64 fp32 channels per tile
      mbar_wait(bar, phase);
      tc_fence_after();
      for (int tl = 0; tl < nt; ++tl) {
        uint8_t* stage_row = a_tiles + tl * kATile + threadIdx.x * 128;  // A tiles are free once the MMAs have completed
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(tl * 64 + ch * 32), v);
          tmem_ld_wait();
          if (!kApply) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float f = __uint_as_float(v[i]);
              s1[ch * 32 + i] += f;
              s2[ch * 32 + i] = fmaf(f, f, s2[ch * 32 + i]);
            }
          } else {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int cidx = ch * 32 + c4 * 8 + 2 * j;
                const float2 k0 = coef[cidx], k1 = coef[cidx + 1];
                const float a0 = fmaxf(fmaf(__uint_as_float(v[c4 * 8 + 2 * j]), k0.x, k0.y), 0.f);
                const float a1 = fmaxf(fmaf(__uint_as_float(v[c4 * 8 + 2 * j + 1]), k1.x, k1.y), 0.f);
                pk[j] = Cvt<T>::pack2(a0, a1);
              }
              *reinterpret_cast<uint4*>(stage_row + (((ch * 4 + c4) ^ (threadIdx.x & 7)) << 4)) =
                  make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
        if (kApply) {
          __syncwarp();
          // this warp's 32 pixels x 128 B are contiguous in NHWC: 256 x 16-byte chunks, 8 per lane, fully coalesced
          T* dst = out_pad + ((size_t(b) * (H + 2) + y + 1) * (W + 2) + 1 + (t0 + tl) * 128 + warp * 32) * 64;
          const uint8_t* src = a_tiles + tl * kATile + warp * 32 * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int chunk = j * 32 + lane, px = chunk >> 3, c8 = chunk & 7;
            reinterpret_cast<uint4*>(dst)[chunk] = *reinterpret_cast<const uint4*>(src + px * 128 + ((c8 ^ (px & 7)) << 4));
          }
        }
      }
      tc_fence_before();
    }
    phase ^= 1;
    __syncthreads();  // TMEM columns and A tiles are reused by the next round
  }

  if (!kApply) {
    if (warp < 4) {
      float t[32];
#pragma unroll
      for (int hsel = 0; hsel < 2; ++hsel) {
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = s1[hsel * 32 + i];
        const float a = bfly_sum(t, lane);
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = s2[hsel * 32 + i];
        const float q = bfly_sum(t, lane);
        red[(warp * 2 + 0) * 64 + hsel * 32 + lane] = a;
        red[(warp * 2 + 1) * 64 + hsel * 32 + lane] = q;
      }
    }
    __syncthreads();
    if (threadIdx.x < 128) {
      const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
      const float a = (red[(0 * 2 + which) * 64 + c] + red[(1 * 2 + which) * 64 + c]) +
                      (red[(2 * 2 + which) * 64 + c] + red[(3 * 2 + which) * 64 + c]);
      float* dst = partials + (size_t(b) * H + y) * 3 * 64;   // one "tile" per image row: [3][64]
      dst[which * 64 + c] = a;
      if (which == 0) dst[2 * 64 + c] = 0.f;                   // max slot: not used by the stem
    }
  } else {
    // zero border of the padded output: left/right pixel of this row, plus the top / bottom rows
    T* prow = out_pad + (size_t(b) * (H + 2) + y + 1) * (W + 2) * 64;
    if (threadIdx.x < 16) {
      uint4* z = reinterpret_cast<uint4*>(threadIdx.x < 8 ? prow : prow + size_t(W + 1) * 64);
      z[threadIdx.x & 7] = make_uint4(0, 0, 0, 0);
    }
    if (y == 0 || y == H - 1) {
      uint4* z = reinterpret_cast<uint4*>(out_pad + (size_t(b) * (H + 2) + (y == 0 ? 0 : H + 1)) * (W + 2) * 64);
      for (int i = threadIdx.x; i < (W + 2) * 8; i += kStemThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTilesPerCta * 64);
  }
}

template <typename T, bool kApply>
int launch_stem_pass(const void* xw, const void* wp, float* partials, const float* scale, const float* shift,
                     void* out_pad, int B, int H, int W, cudaStream_t st) {
  const size_t smem = 1024 + size_t(kTilesPerCta) * kATile + 64 * 128 + 64 * 8 + 4 * 2 * 64 * 4 + 32 +
                      size_t(7) * (W + 16) * 2 + 16;
  auto kern = stem_fused_kernel<T, kApply>;
  static PerDeviceOnce configured;
  {
    const cudaError_t e = configured.once([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); });
    if (e != cudaSuccess) return fail(DUCOSY_ERR_CUDA, "cudaFuncSetAttribute(stem_fused): %s", cudaGetErrorString(e));
  }
  DUCOSY_CHECK(smem <= 112 * 1024, DUCOSY_ERR_SHAPE, "stem_fused: image too wide (W = %d)", W);
  pdl(kern, dim3(H, B), kStemThreads, smem, st)(static_cast<const T*>(xw), static_cast<const T*>(wp), partials, scale,
                                               shift, static_cast<T*>(out_pad), H, W);
  return check_launch("stem_fused_kernel");
}

template <typename T, typename In>
int launch_stem_input(In in, void* xw, int B, int H, int W, cudaStream_t st) {
  const long long total = (long long)B * H * W;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)(num_sms() > 0 ? num_sms() : 148) * 8;
  if (blocks > cap) blocks = cap;
  pdl(stem_input_kernel<T, In>, int(blocks), 256, 0, st)(in, static_cast<T*>(xw), B, H, W);
  return check_launch("stem_input_kernel");
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

// Network input (fp32 [B][1][H][W], or stored int16 pixels through the HU window of preprocess.py:72-84) -> 16-bit [B][H][W].
extern "C" int ducosy_stem_prepare(const float* x_nchw, const int16_t* px, float slope, float intercept, float lo, float hi,
                                   void* xw, int B, int H, int W, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK((x_nchw != nullptr) != (px != nullptr), DUCOSY_ERR_ARG, "stem_prepare: give exactly one of x_nchw / px");
  DUCOSY_CHECK(xw && B > 0 && H > 0 && W > 0, DUCOSY_ERR_ARG, "stem_prepare: bad argument");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "stem_prepare: bad dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (px != nullptr) {
    InHU in{px, slope, intercept, lo, hi, float(double(hi) - double(lo))};
    DUCOSY_DISPATCH_DTYPE(dtype, T, return (launch_stem_input<T, InHU>(in, xw, B, H, W, st)));
  }
  InF32 in{x_nchw};
  DUCOSY_DISPATCH_DTYPE(dtype, T, return (launch_stem_input<T, InF32>(in, xw, B, H, W, st)));
}

// pass 1 (apply == 0): partials [B][H][3][64] (per image row: sum, sum of squares, unused); pass 2 (scale/shift from
// ducosy_in_finalize(partials, H, H*W, ...)): out_pad [B][H+2][W+2][64] = ReLU(IN(conv)), zero border.
extern "C" int ducosy_stem_fused(const void* xw, const void* w_packed, float* partials, const float* scale,
                                 const float* shift, void* out_pad, int B, int H, int W, int apply, int dtype,
                                 ducosy_stream_t stream) {
  DUCOSY_CHECK(xw && w_packed && B > 0 && H >= 4 && W >= 128 && W % 128 == 0, DUCOSY_ERR_SHAPE,
               "stem_fused: W must be a multiple of 128");
  DUCOSY_CHECK(apply ? (scale && shift && out_pad) : (partials != nullptr), DUCOSY_ERR_ARG, "stem_fused: missing buffers for this pass");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "stem_fused: bad dtype");
  DUCOSY_CHECK((reinterpret_cast<uintptr_t>(xw) & 15) == 0, DUCOSY_ERR_ALIGN, "stem_fused: input must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (apply) {
    DUCOSY_DISPATCH_DTYPE(dtype, T, return (launch_stem_pass<T, true>(xw, w_packed, partials, scale, shift, out_pad, B, H, W, st)));
  }
  DUCOSY_DISPATCH_DTYPE(dtype, T, return (launch_stem_pass<T, false>(xw, w_packed, partials, scale, shift, out_pad, B, H, W, st)));
}
