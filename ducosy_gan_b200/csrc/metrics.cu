// Image-quality metrics over volumes: the arithmetic of calculate.py:232-271,360-381 -- normalize,
// calculate_mae, calculate_psnr, calculate_ssim (skimage.metrics.structural_similarity defaults), calculate_cs,
// calculate_ed -- as bandwidth-bound reductions on the device, so that the evaluation of a synthesized volume does not
// have to leave HBM.  Everything accumulates in float64 like numpy does for these calls; reductions are two-stage with a
// fixed order (deterministic).  For int16 volumes (the stored pixel arrays calculate.py:226-228 saves) `img1 - img2` and
// `(img1 - img2) ** 2` are evaluated in int16 with wrap-around exactly as numpy does for that dtype.
#include "common.cuh"

namespace ducosy {
namespace {

constexpr int kStat = 12;      // per-slice statistics, see metrics_pair_kernel
constexpr int kThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> struct IsI16 { static constexpr bool v = false; };
template <> struct IsI16<int16_t> { static constexpr bool v = true; };

// Per (slice, chunk) partials of
//   0 sum |a-b|   1 sum (a-b)^2   2 sum a*b   3 sum a*a   4 sum b*b   5 sum a   6 sum b   7 min a   8 max a   9 min b  10 max b
//   11 unused
// int16 inputs: entries 0 and 1 use numpy's int16 arithmetic (difference and square wrap modulo 2^16); all sums of
// integers are exact in double here (< 2^53).
template <typename In>
__global__ void __launch_bounds__(kThreads)
metrics_pair_kernel(const In* __restrict__ a, const In* __restrict__ b, long long n, int chunks, double* __restrict__ part) {
  pdl_prologue();
  const int s = blockIdx.y, c = blockIdx.x;
  const long long per = (n + chunks - 1) / chunks;
  const long long lo = c * per, hi = lo + per < n ? lo + per : n;
  const In* pa = a + (long long)s * n;
  const In* pb = b + (long long)s * n;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  double mna = INFINITY, mxa = -INFINITY, mnb = INFINITY, mxb = -INFINITY;
  for (long long i = lo + threadIdx.x; i < hi; i += kThreads) {
    const In va = pa[i], vb = pb[i];
    const double x = double(va), y = double(vb);
    double d, d2;
    if (IsI16<In>::v) {
      const int16_t di = int16_t(int(va) - int(vb));        // wraps like numpy int16 - int16
      const int16_t ab = int16_t(di < 0 ? -int(di) : int(di));   // np.abs(int16): abs(-32768) stays -32768
      d = double(ab);
      d2 = double(int16_t(int(di) * int(di)));               // int16 ** 2 wraps as well
    } else {
      const double t = x - y;
      d = fabs(t);
      d2 = t * t;
    }
    acc[0] += d; acc[1] += d2; acc[2] += x * y; acc[3] += x * x; acc[4] += y * y; acc[5] += x; acc[6] += y;
    mna = fmin(mna, x); mxa = fmax(mxa, x); mnb = fmin(mnb, y); mxb = fmax(mxb, y);
  }
  __shared__ double sh[kThreads / 32][kStat];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 7; ++k) acc[k] = warp_sum(acc[k]);
  mna = warp_min(mna); mxa = warp_max(mxa); mnb = warp_min(mnb); mxb = warp_max(mxb);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) sh[warp][k] = acc[k];
    sh[warp][7] = mna; sh[warp][8] = mxa; sh[warp][9] = mnb; sh[warp][10] = mxb; sh[warp][11] = 0.0;
  }
  __syncthreads();
  if (threadIdx.x < kStat) {
    const int k = threadIdx.x;
    double v = sh[0][k];
    for (int w = 1; w < kThreads / 32; ++w) v = (k == 7 || k == 9) ? fmin(v, sh[w][k]) : (k == 8 || k == 10) ? fmax(v, sh[w][k]) : v + sh[w][k];
    part[((long long)s * chunks + c) * kStat + k] = v;
  }
}

// stats[s][k] = fixed-order reduction of the chunk partials
__global__ void metrics_finalize_kernel(const double* __restrict__ part, int S, int chunks, int K, double* __restrict__ stats) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * K) return;
  const int s = i / K, k = i % K;
  const bool is_min = K == kStat && (k == 7 || k == 9), is_max = K == kStat && (k == 8 || k == 10);
  double v = part[((long long)s * chunks) * K + k];
  for (int c = 1; c < chunks; ++c) {
    const double o = part[((long long)s * chunks + c) * K + k];
    v = is_min ? fmin(v, o) : is_max ? fmax(v, o) : v + o;
  }
  stats[(long long)s * K + k] = v;
}

// calculate_ed (calculate.py:369-381): sum over the slice of ((a - min a)/(range a + 1e-8) - (b - min b)/(range b + 1e-8))^2
template <typename In>
__global__ void __launch_bounds__(kThreads)
metrics_ed_kernel(const In* __restrict__ a, const In* __restrict__ b, long long n, int chunks, const double* __restrict__ stats,
                  double* __restrict__ part) {
  pdl_prologue();
  const int s = blockIdx.y, c = blockIdx.x;
  const long long per = (n + chunks - 1) / chunks;
  const long long lo = c * per, hi = lo + per < n ? lo + per : n;
  const double mna = stats[s * kStat + 7], ra = (stats[s * kStat + 8] - mna) + 1e-8;
  const double mnb = stats[s * kStat + 9], rb = (stats[s * kStat + 10] - mnb) + 1e-8;
  double acc = 0.0;
  for (long long i = lo + threadIdx.x; i < hi; i += kThreads) {
    const double x = (double(a[(long long)s * n + i]) - mna) / ra, y = (double(b[(long long)s * n + i]) - mnb) / rb;
    const double t = x - y;
    acc += t * t;
  }
  __shared__ double sh[kThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = sh[0];
    for (int w = 1; w < kThreads / 32; ++w) v += sh[w];
    part[(long long)s * chunks + c] = v;
  }
}

// normalize (calculate.py:232-238): (data - min) / (max - min) in float64; zeros when the range is 0.  mm = {min, max}.
template <typename In>
__global__ void metrics_normalize_kernel(const In* __restrict__ in, double* __restrict__ out, long long total, const double* __restrict__ mm) {
  pdl_prologue();
  const double mn = mm[0], r = mm[1] - mm[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    out[i] = r == 0.0 ? 0.0 : (double(in[i]) - mn) / r;
}

// skimage.metrics.structural_similarity with its defaults (win_size 7, uniform window, use_sample_covariance=True,
// K1 0.01, K2 0.03): per pixel of the valid region [3, H-3) x [3, W-3)
//   ux, uy, uxx, uyy, uxy = 7x7 means;  vx = cov_norm (uxx - ux^2), vy, vxy;  cov_norm = 49/48
//   S = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2));  mssim = mean S  (float64)
// The crop equals the window radius, so the boundary mode of the uniform filter never enters.
constexpr int kSsimTile = 32, kSsimHalo = kSsimTile + 6;
template <typename In>
__global__ void __launch_bounds__(kThreads)
metrics_ssim_kernel(const In* __restrict__ a, const In* __restrict__ b, int H, int W, double data_range, double* __restrict__ part) {
  pdl_prologue();
  __shared__ double ta[kSsimHalo][kSsimHalo + 1], tb[kSsimHalo][kSsimHalo + 1];
  __shared__ double red[kThreads / 32];
  const int s = blockIdx.z;
  const int y0 = 3 + blockIdx.y * kSsimTile, x0 = 3 + blockIdx.x * kSsimTile;   // first output pixel of the tile
  const In* pa = a + (long long)s * H * W;
  const In* pb = b + (long long)s * H * W;
  for (int i = threadIdx.x; i < kSsimHalo * kSsimHalo; i += kThreads) {
    const int r = i / kSsimHalo, c = i % kSsimHalo;
    const int y = y0 - 3 + r, x = x0 - 3 + c;
    const bool ok = y < H && x < W;
    ta[r][c] = ok ? double(pa[(long long)y * W + x]) : 0.0;
    tb[r][c] = ok ? double(pb[(long long)y * W + x]) : 0.0;
  }
  __syncthreads();
  const double C1 = (0.01 * data_range) * (0.01 * data_range), C2 = (0.03 * data_range) * (0.03 * data_range);
  const double cov_norm = 49.0 / 48.0;
  double acc = 0.0;
  for (int i = threadIdx.x; i < kSsimTile * kSsimTile; i += kThreads) {
    const int r = i / kSsimTile, c = i % kSsimTile;
    if (y0 + r >= H - 3 || x0 + c >= W - 3) continue;
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
    for (int dr = 0; dr < 7; ++dr)
#pragma unroll
      for (int dc = 0; dc < 7; ++dc) {
        const double x = ta[r + dr][c + dc], y = tb[r + dr][c + dc];
        sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y;
      }
    const double ux = sx / 49.0, uy = sy / 49.0, uxx = sxx / 49.0, uyy = syy / 49.0, uxy = sxy / 49.0;
    const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
    const double A1 = 2 * ux * uy + C1, A2 = 2 * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
    acc += (A1 * A2) / (B1 * B2);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = red[0];
    for (int w = 1; w < kThreads / 32; ++w) v += red[w];
    part[((long long)s * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = v;
  }
}

// ---- calculate_emd (calculate.py:320-337) for int16 volumes.  scipy.stats.wasserstein_distance of two equal-weight samples is
// the integral of |CDF1 - CDF2|; the reference normalises both slices with the same affine map first, so on integer-valued data
//     d = sum_v |C1(v) - C2(v)| / n / (global_max - global_min + 1e-8),      C = cumulative histogram of the raw values
// -- integer arithmetic up to the final division (verified to the last bit against scipy).  hist: [S][2][R] zeroed counters.
__global__ void __launch_bounds__(kThreads)
metrics_hist_kernel(const int16_t* __restrict__ a, const int16_t* __restrict__ b, long long n, int chunks, int vmin, int R,
                    unsigned int* __restrict__ hist) {
  pdl_prologue();
  const int s = blockIdx.y, c = blockIdx.x;
  const long long per = (n + chunks - 1) / chunks;
  const long long lo = c * per, hi = lo + per < n ? lo + per : n;
  unsigned int* ha = hist + (size_t(s) * 2 + 0) * R;
  unsigned int* hb = hist + (size_t(s) * 2 + 1) * R;
  for (long long i = lo + threadIdx.x; i < hi; i += kThreads) {
    atomicAdd(ha + (int(a[(long long)s * n + i]) - vmin), 1u);   // integer counters: the order of the atomics does not matter
    atomicAdd(hb + (int(b[(long long)s * n + i]) - vmin), 1u);
  }
}

// one CTA per slice: out[s] = sum_v |C1(v) - C2(v)|  (exact in double: < 2^53)
__global__ void __launch_bounds__(kThreads)
metrics_emd_scan_kernel(const unsigned int* __restrict__ hist, int R, double* __restrict__ out) {
  pdl_prologue();
  __shared__ long long wsum[kThreads / 32];
  __shared__ long long carry_s;
  __shared__ double red[kThreads / 32];
  const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int* ha = hist + (size_t(s) * 2 + 0) * R;
  const unsigned int* hb = hist + (size_t(s) * 2 + 1) * R;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  double acc = 0.0;
  for (int v0 = 0; v0 < R; v0 += kThreads) {
    const int v = v0 + threadIdx.x;
    long long d = v < R ? (long long)ha[v] - (long long)hb[v] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {          // inclusive scan inside the warp
      const long long t = __shfl_up_sync(0xffffffffu, d, o);
      if (lane >= o) d += t;
    }
    if (lane == 31) wsum[warp] = d;
    __syncthreads();
    long long before = carry_s;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (v < R) acc += double(llabs(d + before));
    __syncthreads();
    if (threadIdx.x == kThreads - 1) carry_s = before + d;
    __syncthreads();
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = red[0];
    for (int w = 1; w < kThreads / 32; ++w) t += red[w];
    out[s] = t;
  }
}

// ---- calculate_ts (calculate.py:340-358): skimage.filters.sobel of both slices (separable: [1,2,1]/4 smoothing x [1,0,-1]
// difference, boundary mode 'reflect', magnitude sqrt((gx^2 + gy^2)/2)); per (slice, CTA) partials of sum |g1 - g2|, max g1,
// max g2.  TS is a ratio of gradient magnitudes, so skimage's int -> float rescaling of the input cancels.  PARITY UNPINNED
// (scikit-image absent): restated from its published implementation (versions >= 0.18, which no longer zero the border).
template <typename In>
__global__ void __launch_bounds__(kThreads)
metrics_ts_kernel(const In* __restrict__ a, const In* __restrict__ b, int H, int W, int rows_per_cta, double* __restrict__ part) {
  pdl_prologue();
  const int s = blockIdx.y;
  const int y0 = blockIdx.x * rows_per_cta, y1 = min(H, y0 + rows_per_cta);
  const In* pa = a + (long long)s * H * W;
  const In* pb = b + (long long)s * H * W;
  auto rf = [](int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); };   // scipy 'reflect' at distance 1: (a | a b ... y z | z)
  double acc = 0.0, m1 = 0.0, m2 = 0.0;
  for (int i = y0 * W + threadIdx.x; i < y1 * W; i += kThreads) {
    const int y = i / W, x = i - y * W;
    const int ym = rf(y - 1, H), yp = rf(y + 1, H), xm = rf(x - 1, W), xp = rf(x + 1, W);
    auto mag = [&](const In* p) {
      const double a00 = double(p[ym * W + xm]), a01 = double(p[ym * W + x]), a02 = double(p[ym * W + xp]);
      const double a10 = double(p[y * W + xm]), a12 = double(p[y * W + xp]);
      const double a20 = double(p[yp * W + xm]), a21 = double(p[yp * W + x]), a22 = double(p[yp * W + xp]);
      const double gy = ((a00 + 2.0 * a01 + a02) - (a20 + 2.0 * a21 + a22)) * 0.25;   // edge along axis 0, smoothed along axis 1
      const double gx = ((a00 + 2.0 * a10 + a20) - (a02 + 2.0 * a12 + a22)) * 0.25;
      return sqrt((gx * gx + gy * gy) * 0.5);
    };
    const double g1 = mag(pa), g2 = mag(pb);
    acc += fabs(g1 - g2);
    m1 = fmax(m1, g1);
    m2 = fmax(m2, g2);
  }
  __shared__ double sh[kThreads / 32][3];
  acc = warp_sum(acc); m1 = warp_max(m1); m2 = warp_max(m2);
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = acc; sh[threadIdx.x >> 5][1] = m1; sh[threadIdx.x >> 5][2] = m2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) { acc += sh[w][0]; m1 = fmax(m1, sh[w][1]); m2 = fmax(m2, sh[w][2]); }
    double* dst = part + ((long long)s * gridDim.x + blockIdx.x) * 3;
    dst[0] = acc;                            // warp 0's sum + the other warps' in index order (fixed order)
    dst[1] = m1;
    dst[2] = m2;
  }
}

__global__ void metrics_ts_finalize_kernel(const double* __restrict__ part, int S, int ctas, double* __restrict__ out) {
  pdl_prologue();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double acc = 0.0, m1 = 0.0, m2 = 0.0;
  for (int c = 0; c < ctas; ++c) {
    const double* p = part + ((long long)s * ctas + c) * 3;
    acc += p[0];
    m1 = fmax(m1, p[1]);
    m2 = fmax(m2, p[2]);
  }
  out[s * 3 + 0] = acc;
  out[s * 3 + 1] = m1;
  out[s * 3 + 2] = m2;
}

int pick_chunks(long long n) {
  long long c = n / 8192;
  if (c < 1) c = 1;
  if (c > 64) c = 64;
  return int(c);
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

#define DUCOSY_DISPATCH_IN(in_type, In, ...)                                            \
  do {                                                                                  \
    if ((in_type) == DUCOSY_IN_I16) { using In = int16_t; __VA_ARGS__; }                \
    else if ((in_type) == DUCOSY_IN_F32) { using In = float; __VA_ARGS__; }             \
    else { using In = double; __VA_ARGS__; }                                            \
  } while (0)

extern "C" int ducosy_metrics_chunks(long long n) { return pick_chunks(n); }

extern "C" int ducosy_metrics_slice_stats(const void* a, const void* b, int in_type, int S, long long n, double* stats,
                                          double* scratch, ducosy_stream_t stream) {
  DUCOSY_CHECK(a && b && stats && scratch && S > 0 && n > 0, DUCOSY_ERR_ARG, "metrics_slice_stats: bad argument");
  DUCOSY_CHECK(in_type >= DUCOSY_IN_I16 && in_type <= DUCOSY_IN_F64, DUCOSY_ERR_ARG, "metrics_slice_stats: bad input type");
  const int chunks = pick_chunks(n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DUCOSY_DISPATCH_IN(in_type, In, (pdl(metrics_pair_kernel<In>, dim3(chunks, S), kThreads, 0, st)(
                                      static_cast<const In*>(a), static_cast<const In*>(b), n, chunks, scratch)));
  DUCOSY_TRY(check_launch("metrics_pair_kernel"));
  pdl(metrics_finalize_kernel, (S * kStat + 255) / 256, 256, 0, st)(scratch, S, chunks, kStat, stats);
  return check_launch("metrics_finalize_kernel");
}

extern "C" int ducosy_metrics_ed(const void* a, const void* b, int in_type, int S, long long n, const double* stats, double* ed_sums,
                                 double* scratch, ducosy_stream_t stream) {
  DUCOSY_CHECK(a && b && stats && ed_sums && scratch && S > 0 && n > 0, DUCOSY_ERR_ARG, "metrics_ed: bad argument");
  DUCOSY_CHECK(in_type >= DUCOSY_IN_I16 && in_type <= DUCOSY_IN_F64, DUCOSY_ERR_ARG, "metrics_ed: bad input type");
  const int chunks = pick_chunks(n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DUCOSY_DISPATCH_IN(in_type, In, (pdl(metrics_ed_kernel<In>, dim3(chunks, S), kThreads, 0, st)(
                                      static_cast<const In*>(a), static_cast<const In*>(b), n, chunks, stats, scratch)));
  DUCOSY_TRY(check_launch("metrics_ed_kernel"));
  pdl(metrics_finalize_kernel, (S + 255) / 256, 256, 0, st)(scratch, S, chunks, 1, ed_sums);
  return check_launch("metrics_finalize_kernel");
}

extern "C" int ducosy_metrics_normalize(const void* in, int in_type, double* out, long long total, const double* minmax,
                                        ducosy_stream_t stream) {
  DUCOSY_CHECK(in && out && minmax && total > 0, DUCOSY_ERR_ARG, "metrics_normalize: bad argument");
  DUCOSY_CHECK(in_type >= DUCOSY_IN_I16 && in_type <= DUCOSY_IN_F64, DUCOSY_ERR_ARG, "metrics_normalize: bad input type");
  const int sms = num_sms() > 0 ? num_sms() : 148;
  long long blocks = (total + 255) / 256;
  if (blocks > sms * 16LL) blocks = sms * 16LL;
  DUCOSY_DISPATCH_IN(in_type, In, (pdl(metrics_normalize_kernel<In>, int(blocks), 256, 0, static_cast<cudaStream_t>(stream))(
                                      static_cast<const In*>(in), out, total, minmax)));
  return check_launch("metrics_normalize_kernel");
}

extern "C" int ducosy_metrics_ssim_tiles(int H, int W) {
  if (H < 7 || W < 7) return 0;
  return ((H - 6 + kSsimTile - 1) / kSsimTile) * ((W - 6 + kSsimTile - 1) / kSsimTile);
}

extern "C" int ducosy_metrics_ssim(const void* a, const void* b, int in_type, int S, int H, int W, double data_range,
                                   double* ssim_sums, double* scratch, ducosy_stream_t stream) {
  DUCOSY_CHECK(a && b && ssim_sums && scratch && S > 0, DUCOSY_ERR_ARG, "metrics_ssim: bad argument");
  DUCOSY_CHECK(H >= 7 && W >= 7, DUCOSY_ERR_SHAPE, "metrics_ssim: win_size 7 exceeds the image extent (%dx%d)", H, W);
  DUCOSY_CHECK(in_type >= DUCOSY_IN_I16 && in_type <= DUCOSY_IN_F64, DUCOSY_ERR_ARG, "metrics_ssim: bad input type");
  const int gx = (W - 6 + kSsimTile - 1) / kSsimTile, gy = (H - 6 + kSsimTile - 1) / kSsimTile;
  DUCOSY_CHECK(S <= 65535, DUCOSY_ERR_SHAPE, "metrics_ssim: at most 65535 slices per call");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DUCOSY_DISPATCH_IN(in_type, In, (pdl(metrics_ssim_kernel<In>, dim3(gx, gy, S), kThreads, 0, st)(
                                      static_cast<const In*>(a), static_cast<const In*>(b), H, W, data_range, scratch)));
  DUCOSY_TRY(check_launch("metrics_ssim_kernel"));
  pdl(metrics_finalize_kernel, (S + 255) / 256, 256, 0, st)(scratch, S, gx * gy, 1, ssim_sums);
  return check_launch("metrics_finalize_kernel");
}

/* calculate_emd for int16 volumes: hist = S*2*R zeroed uint32 counters (R = global max - global min + 1, vmin = global min);
 * cdf_abs_sums[S] = sum_v |C1(v) - C2(v)| per slice. */
extern "C" int ducosy_metrics_emd_i16(const int16_t* a, const int16_t* b, int S, long long n, int vmin, int R, unsigned int* hist,
                                      double* cdf_abs_sums, ducosy_stream_t stream) {
  DUCOSY_CHECK(a && b && hist && cdf_abs_sums && S > 0 && n > 0 && R > 0 && R <= 65536, DUCOSY_ERR_ARG, "metrics_emd_i16: bad argument");
  DUCOSY_CHECK(S <= 65535, DUCOSY_ERR_SHAPE, "metrics_emd_i16: at most 65535 slices per call");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunks = pick_chunks(n);
  pdl(metrics_hist_kernel, dim3(chunks, S), kThreads, 0, st)(a, b, n, chunks, vmin, R, hist);
  DUCOSY_TRY(check_launch("metrics_hist_kernel"));
  pdl(metrics_emd_scan_kernel, S, kThreads, 0, st)(hist, R, cdf_abs_sums);
  return check_launch("metrics_emd_scan_kernel");
}

/* calculate_ts: ts_stats[S][3] = per slice (sum |sobel(a) - sobel(b)|, max sobel(a), max sobel(b)); scratch: S * ceil(H/8) * 3 doubles. */
extern "C" int ducosy_metrics_ts(const void* a, const void* b, int in_type, int S, int H, int W, double* ts_stats, double* scratch,
                                 ducosy_stream_t stream) {
  DUCOSY_CHECK(a && b && ts_stats && scratch && S > 0 && H > 0 && W > 0, DUCOSY_ERR_ARG, "metrics_ts: bad argument");
  DUCOSY_CHECK(in_type >= DUCOSY_IN_I16 && in_type <= DUCOSY_IN_F64, DUCOSY_ERR_ARG, "metrics_ts: bad input type");
  DUCOSY_CHECK(S <= 65535 && (long long)H * W < (1LL << 31), DUCOSY_ERR_SHAPE, "metrics_ts: volume too large for one call");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = 8, ctas = (H + rows - 1) / rows;
  DUCOSY_DISPATCH_IN(in_type, In, (pdl(metrics_ts_kernel<In>, dim3(ctas, S), kThreads, 0, st)(
                                      static_cast<const In*>(a), static_cast<const In*>(b), H, W, rows, scratch)));
  DUCOSY_TRY(check_launch("metrics_ts_kernel"));
  pdl(metrics_ts_finalize_kernel, (S + 255) / 256, 256, 0, st)(scratch, S, ctas, ts_stats);
  return check_launch("metrics_ts_finalize_kernel");
}
