// Post-composite volume smoothing: what generate.py:254-263 + modules/postprocess.py:47-60,99-109,
// 114-160 do on the CPU with scipy, kept on the device so the merged volume never leaves HBM between the composite and
// the DICOM writer.  Pipeline on a [S][H][W] int16 volume of stored values:
//   v1 = gaussian_filter1d(float32(vol), sigma_z_pre, axis 0)                     (float32 result)
//   pp = gaussian_filter(v1, (sigma_z, ~0, ~0))                                   (float32; the xy kernels have radius 0)
//   unsharp mask in float64: blur_xy(pp), blur_xy(v1) with sigma = radius -> sharpened, clipped to [min v1, max v1]
//   voxels with v1 >= hu_threshold keep v1; result truncated to int16.
// Bit-exactness with scipy needs its exact arithmetic: every 1-D pass accumulates in double as
//   tmp = x[0]*w[0];  for j = -r..-1: tmp += (x[j] + x[-j]) * w[j]
// (ni_filters.c, symmetric-kernel branch), un-fused (x86-64 baseline builds have no FMA), boundary mode 'reflect'
// (d c b a | a b c d | d c b a), results rounded to the array dtype between passes.  __dmul_rn / __dadd_rn keep the
// compiler from contracting into FMAs.
#include "common.cuh"

namespace ducosy {
namespace {

constexpr int kMaxRZ = 4;    // z kernels: sigma < 1.125 (generate.py uses 0.8 and 0.7 -> radius 3)
constexpr int kMaxRXY = 8;   // xy kernel: sigma < 2.125 (generate.py uses 1.2 -> radius 5)

struct PostWeights {
  double wz1[2 * kMaxRZ + 1], wz2[2 * kMaxRZ + 1], wxy[2 * kMaxRXY + 1];   // centred: w[r + j], j = -r..r
  double amount, one_minus_amount;
  float threshold;
  int rz1, rz2, rxy;
};

__device__ __forceinline__ int reflect_ext(int i, int n) {   // scipy.ndimage mode='reflect', any i
  const int p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i >= n ? p - 1 - i : i;
}

__device__ __forceinline__ double to_d(short v) { return double(float(v)); }
__device__ __forceinline__ double to_d(float v) { return double(v); }

// 1-D filter along z, thread = pixel (coalesced across x), sliding register window.  Optionally the per-CTA min / max of
// the float32 result (the clip range of the unsharp mask is the range of v1).
template <typename TIn, int R, bool kMinMax>
__global__ void __launch_bounds__(256)
zfilter_kernel(const TIn* __restrict__ in, float* __restrict__ out, float* __restrict__ pmin, float* __restrict__ pmax, int S,
               long long HW, PostWeights pw, int which, int mm_z0, int mm_z1) {
  pdl_prologue();
  const double* w = which == 1 ? pw.wz1 : pw.wz2;
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  float mn = INFINITY, mx = -INFINITY;
  if (p < HW) {
    double win[2 * R + 1];
#pragma unroll
    for (int j = -R; j <= R; ++j) win[R + j] = to_d(in[(long long)reflect_ext(j, S) * HW + p]);
    for (int z = 0; z < S; ++z) {
      double tmp = __dmul_rn(win[R], w[R]);
#pragma unroll
      for (int j = -R; j < 0; ++j) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(win[R + j], win[R - j]), w[R + j]));
      const float o = float(tmp);
      out[(long long)z * HW + p] = o;
      if (kMinMax && z >= mm_z0 && z < mm_z1) { mn = fminf(mn, o); mx = fmaxf(mx, o); }
#pragma unroll
      for (int j = 0; j < 2 * R; ++j) win[j] = win[j + 1];
      win[2 * R] = to_d(in[(long long)reflect_ext(z + 1 + R, S) * HW + p]);
    }
  }
  if (kMinMax) {
    __shared__ float smn[8], smx[8];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < 8; ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
      pmin[blockIdx.x] = fminf(mn, smn[0]);
      pmax[blockIdx.x] = fmaxf(mx, smx[0]);
    }
  }
}

__global__ void minmax_finalize_kernel(const float* __restrict__ pmin, const float* __restrict__ pmax, int n, float* __restrict__ mm) {
  pdl_prologue();
  __shared__ float smn[32], smx[32];
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { mn = fminf(mn, pmin[i]); mx = fmaxf(mx, pmax[i]); }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < int(blockDim.x >> 5); ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    mm[0] = mn;
    mm[1] = mx;
  }
}

// xy unsharp mask + clip + high-density restore + int16 cast.  CTA = 32x32 output tile of one slice; per source array
// (pp, then v1): reflect-extended tile in shared memory (double), vertical pass (axis 1 first, like scipy), horizontal pass.
constexpr int kTile = 32;
template <int R>
__global__ void __launch_bounds__(256)
unsharp_finalize_kernel(const float* __restrict__ pp, const float* __restrict__ v1, const float* __restrict__ mm,
                        short* __restrict__ out, int H, int W, PostWeights pw) {
  pdl_prologue();
  constexpr int TW = kTile + 2 * R;
  __shared__ double tile[TW][TW];
  __shared__ double vert[kTile][TW];
  const int z = blockIdx.z, y0 = blockIdx.y * kTile, x0 = blockIdx.x * kTile;
  const long long base = (long long)z * H * W;
  const double* w = pw.wxy;
  double blur[2][4];
  for (int a = 0; a < 2; ++a) {
    const float* src = (a == 0 ? pp : v1) + base;
    __syncthreads();
    for (int i = threadIdx.x; i < TW * TW; i += 256) {
      const int ty = i / TW, tx = i - ty * TW;
      tile[ty][tx] = double(src[(long long)reflect_ext(y0 + ty - R, H) * W + reflect_ext(x0 + tx - R, W)]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kTile * TW; i += 256) {
      const int ty = i / TW, tx = i - ty * TW;      // output row ty of the tile, column tx of the extended tile
      double tmp = __dmul_rn(tile[ty + R][tx], w[R]);
#pragma unroll
      for (int j = -R; j < 0; ++j) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(tile[ty + R + j][tx], tile[ty + R - j][tx]), w[R + j]));
      vert[ty][tx] = tmp;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = threadIdx.x + k * 256, ty = i / kTile, tx = i % kTile;
      double tmp = __dmul_rn(vert[ty][tx + R], w[R]);
#pragma unroll
      for (int j = -R; j < 0; ++j) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(vert[ty][tx + R + j], vert[ty][tx + R - j]), w[R + j]));
      blur[a][k] = tmp;
    }
  }
  const double lo = double(mm[0]), hi = double(mm[1]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * 256, ty = i / kTile, tx = i % kTile;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const long long off = base + (long long)y * W + x;
    const float o32 = v1[off];
    const double s = double(pp[off]), o = double(o32);
    const double hf = __dadd_rn(s, -blur[0][k]), ohf = __dadd_rn(o, -blur[1][k]);
    const double comb = __dadd_rn(__dmul_rn(pw.one_minus_amount, hf), __dmul_rn(pw.amount, ohf));
    double sh = __dadd_rn(s, __dmul_rn(comb, pw.amount));
    sh = fmin(fmax(sh, lo), hi);
    const double r = o32 >= pw.threshold ? o : sh;
    out[off] = short(int(r));   // astype(np.int16): truncation toward zero
  }
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" size_t ducosy_postprocess_scratch_bytes(int S, int H, int W) {
  const size_t n = size_t(S) * H * W, blocks = (size_t(H) * W + 255) / 256;
  return (2 * n + 2 * blocks + 2) * sizeof(float);
}

extern "C" size_t ducosy_postprocess_minmax_offset_bytes(int S, int H, int W) {
  return ducosy_postprocess_scratch_bytes(S, H, W) - 2 * sizeof(float);
}

extern "C" int ducosy_postprocess_volume(const int16_t* merged, int16_t* out, float* scratch, int S, int H, int W,
                                         const double* wz1, int rz1, const double* wz2, int rz2, const double* wxy, int rxy,
                                         double sharpen_amount, float hu_threshold, int phases, int mm_z0, int mm_z1,
                                         ducosy_stream_t stream) {
  DUCOSY_CHECK(phases >= 1 && phases <= 3 && mm_z0 >= 0 && mm_z1 <= S && mm_z0 < mm_z1, DUCOSY_ERR_ARG,
               "postprocess_volume: phases must be 1, 2 or 3 and 0 <= mm_z0 < mm_z1 <= S");
  DUCOSY_CHECK(merged && out && scratch && wz1 && wz2 && wxy, DUCOSY_ERR_ARG, "postprocess_volume: null pointer");
  DUCOSY_CHECK(S > 0 && H > 0 && W > 0, DUCOSY_ERR_SHAPE, "postprocess_volume: empty volume");
  DUCOSY_CHECK(rz1 >= 1 && rz1 <= kMaxRZ && rz2 >= 1 && rz2 <= kMaxRZ, DUCOSY_ERR_ARG,
               "postprocess_volume: z kernel radius must be 1..%d (got %d, %d)", kMaxRZ, rz1, rz2);
  DUCOSY_CHECK(rxy >= 1 && rxy <= kMaxRXY, DUCOSY_ERR_ARG, "postprocess_volume: xy kernel radius must be 1..%d (got %d)", kMaxRXY, rxy);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PostWeights pw{};
  for (int i = 0; i < 2 * rz1 + 1; ++i) pw.wz1[i] = wz1[i];
  for (int i = 0; i < 2 * rz2 + 1; ++i) pw.wz2[i] = wz2[i];
  for (int i = 0; i < 2 * rxy + 1; ++i) pw.wxy[i] = wxy[i];
  pw.amount = sharpen_amount;
  pw.one_minus_amount = 1.0 - sharpen_amount;   // the Python expression (1 - amount), same IEEE subtraction
  pw.threshold = hu_threshold;
  pw.rz1 = rz1; pw.rz2 = rz2; pw.rxy = rxy;
  const long long HW = (long long)H * W;
  const size_t n = size_t(S) * HW;
  const int blocks = int((HW + 255) / 256);
  float* v1 = scratch;
  float* pp = v1 + n;
  float* pmin = pp + n;
  float* pmax = pmin + blocks;
  float* mm = pmax + blocks;
#define ZF(TIn, R, MM, src, dst, which)                                                                     \
  pdl(zfilter_kernel<TIn, R, MM>, blocks, 256, 0, st)(src, dst, pmin, pmax, S, HW, pw, which, mm_z0, mm_z1)
  if (phases & 1) {
  switch (rz1) {
    case 1: ZF(short, 1, true, merged, v1, 1); break;
    case 2: ZF(short, 2, true, merged, v1, 1); break;
    case 3: ZF(short, 3, true, merged, v1, 1); break;
    default: ZF(short, 4, true, merged, v1, 1); break;
  }
  DUCOSY_TRY(check_launch("zfilter_kernel"));
  pdl(minmax_finalize_kernel, 1, 1024, 0, st)(pmin, pmax, blocks, mm);
  DUCOSY_TRY(check_launch("minmax_finalize_kernel"));
  switch (rz2) {
    case 1: ZF(float, 1, false, v1, pp, 2); break;
    case 2: ZF(float, 2, false, v1, pp, 2); break;
    case 3: ZF(float, 3, false, v1, pp, 2); break;
    default: ZF(float, 4, false, v1, pp, 2); break;
  }
  DUCOSY_TRY(check_launch("zfilter_kernel"));
  }
#undef ZF
  if (!(phases & 2)) return 0;
  const dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, S);
  short* o = reinterpret_cast<short*>(out);
  switch (rxy) {
    case 1: pdl(unsharp_finalize_kernel<1>, grid, 256, 0, st)(pp, v1, mm, o, H, W, pw); break;
    case 2: pdl(unsharp_finalize_kernel<2>, grid, 256, 0, st)(pp, v1, mm, o, H, W, pw); break;
    case 3: pdl(unsharp_finalize_kernel<3>, grid, 256, 0, st)(pp, v1, mm, o, H, W, pw); break;
    case 4: pdl(unsharp_finalize_kernel<4>, grid, 256, 0, st)(pp, v1, mm, o, H, W, pw); break;
    case 5: pdl(unsharp_finalize_kernel<5>, grid, 256, 0, st)(pp, v1, mm, o, H, W, pw); break;
    case 6: pdl(unsharp_finalize_kernel<6>, grid, 256, 0, st)(pp, v1, mm, o, H, W, pw); break;
    case 7: pdl(unsharp_finalize_kernel<7>, grid, 256, 0, st)(pp, v1, mm, o, H, W, pw); break;
    default: pdl(unsharp_finalize_kernel<8>, grid, 256, 0, st)(pp, v1, mm, o, H, W, pw); break;
  }
  return check_launch("unsharp_finalize_kernel");
}
