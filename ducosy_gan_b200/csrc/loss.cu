// Loss terms of the CycleGAN step (modules/trainer.py:22-184,347-358,469-512), forward and backward, as fused
// bandwidth-bound kernels over fp32 [B,1,H,W] images.  Every reduction is two-stage and fixed-order (deterministic);
// scalars stay on the device (no host sync): forward writes loss_out[0] and a small `state` block the backward reads;
// backward takes the upstream gradient as a device scalar gout[0].
//   L1  MSE-GAN      trainer.py:347,470,518,523        L5  ContrastAttentionLoss  trainer.py:43-86
//   L2  L1 (cycle / identity) trainer.py:348-349        L6  ContrastRegionLoss     trainer.py:89-130
//   L3  GradientLoss trainer.py:22-40                   L7  ContrastEdgeLoss       trainer.py:133-184
//   L4  SSIM (pytorch_msssim, parity unpinned)  trainer.py:351,485
#include "common.cuh"

namespace ducosy {
namespace {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 1184;   // 148 SMs x 8
constexpr int kMaxQ = 8;

int loss_grid(long long n) {
  long long b = (n + kLossThreads - 1) / kLossThreads;
  if (b > kLossMaxBlocks) b = kLossMaxBlocks;
  if (b < 1) b = 1;
  return int(b);
}

// block-level fixed-order sum of Q per-thread values -> partial[blockIdx.x][q]
template <int Q>
__device__ __forceinline__ void block_partials(float (&v)[Q], float* __restrict__ partial) {
  __shared__ float red[Q][kLossThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    float x = v[q];
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[q][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < Q) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) acc += red[threadIdx.x][w];
    partial[blockIdx.x * Q + threadIdx.x] = acc;
  }
}
// sums[q] = sum over blocks (double, fixed order); one block of 32*Q threads
__global__ void sum_partials_kernel(const float* __restrict__ partial, int blocks, int Q, double* __restrict__ sums) {
  pdl_prologue();
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (q >= Q) return;
  double acc = 0.0;
  for (int b = lane; b < blocks; b += 32) acc += double(partial[b * Q + q]);
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) sums[q] = acc;
}

// ------------------------------------------------------------------ L1 / L2: mean |a-b|, mean (a-t)^2
__global__ void l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ partial) {
  pdl_prologue();
  float v[1] = {0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    v[0] += fabsf(a[i] - b[i]);
  block_partials<1>(v, partial);
}
__global__ void mse_const_fwd_kernel(const float* __restrict__ a, float t, long long n, float* __restrict__ partial) {
  pdl_prologue();
  float v[1] = {0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - t;
    v[0] = fmaf(d, d, v[0]);
  }
  block_partials<1>(v, partial);
}
__global__ void mean_finish_kernel(const double* __restrict__ sums, double inv_n, float* __restrict__ out) {
  pdl_prologue(); out[0] = float(sums[0] * inv_n); }
__global__ void l1_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, const float* __restrict__ gout,
                              float inv_n, float* __restrict__ da) {
  pdl_prologue();
  const float g = gout[0] * inv_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    da[i] = d > 0.f ? g : (d < 0.f ? -g : 0.f);
  }
}
__global__ void mse_const_bwd_kernel(const float* __restrict__ a, float t, long long n, const float* __restrict__ gout, float inv_n,
                                     float* __restrict__ da) {
  pdl_prologue();
  const float g = 2.f * gout[0] * inv_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    da[i] = g * (a[i] - t);
}

// ------------------------------------------------------------------ L3: GradientLoss
__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
__global__ void gradloss_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t, int B, int H, int W,
                                    float* __restrict__ partial) {
  pdl_prologue();
  float v[2] = {0.f, 0.f};  // sum over vertical pairs, sum over horizontal pairs
  const long long n = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % W), y = int((i / W) % H);
    if (y + 1 < H) v[0] += fabsf(fabsf(p[i + W] - p[i]) - fabsf(t[i + W] - t[i]));
    if (x + 1 < W) v[1] += fabsf(fabsf(p[i + 1] - p[i]) - fabsf(t[i + 1] - t[i]));
  }
  block_partials<2>(v, partial);
}
__global__ void gradloss_finish_kernel(const double* __restrict__ sums, double inv_ny, double inv_nx, float* __restrict__ out) {
  pdl_prologue();
  out[0] = float(sums[1] * inv_nx + sums[0] * inv_ny);
}
// pair term f(dp) = | |dp| - |dt| |  ->  df/d(dp) = sgn(|dp|-|dt|) * sgn(dp); p[hi] gets +, p[lo] gets -
__device__ __forceinline__ float pair_grad(float dp, float dt) { return sgn(fabsf(dp) - fabsf(dt)) * sgn(dp); }
__global__ void gradloss_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, int B, int H, int W,
                                    const float* __restrict__ gout, float inv_ny, float inv_nx, float* __restrict__ dp) {
  pdl_prologue();
  const float gy = gout[0] * inv_ny, gx = gout[0] * inv_nx;
  const long long n = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % W), y = int((i / W) % H);
    float acc = 0.f;
    if (y + 1 < H) acc -= gy * pair_grad(p[i + W] - p[i], t[i + W] - t[i]);
    if (y > 0) acc += gy * pair_grad(p[i] - p[i - W], t[i] - t[i - W]);
    if (x + 1 < W) acc -= gx * pair_grad(p[i + 1] - p[i], t[i + 1] - t[i]);
    if (x > 0) acc += gx * pair_grad(p[i] - p[i - 1], t[i] - t[i - 1]);
    dp[i] = acc;
  }
}

// ------------------------------------------------------------------ L5: ContrastAttentionLoss (7x7 box blur, zero pad, /49)
__device__ __forceinline__ float box7(const float* __restrict__ img, int y, int x, int H, int W) {
  float acc = 0.f;
#pragma unroll
  for (int dy = -3; dy <= 3; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = -3; dx <= 3; ++dx) {
      const int xx = x + dx;
      if (xx >= 0 && xx < W) acc += img[(long long)yy * W + xx];
    }
  }
  return acc * (1.f / 49.f);
}
// forward: partial sums of w*|pb-tb|; also writes the per-pixel upstream map u = w*sgn(pb-tb) for the backward
__global__ void attn_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ s, int B, int H,
                                int W, float inv_sigma, float wmin, float wmax, float* __restrict__ umap, float* __restrict__ partial) {
  pdl_prologue();
  float v[1] = {0.f};
  const long long n = (long long)B * H * W, hw = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % W), y = int((i / W) % H);
    const long long base = (i / hw) * hw;
    const float pb = box7(p + base, y, x, H, W), tb = box7(t + base, y, x, H, W), sb = box7(s + base, y, x, H, W);
    const float w = wmin + (wmax - wmin) * (1.f - expf(-fabsf(tb - sb) * inv_sigma));
    v[0] += w * fabsf(pb - tb);
    if (umap != nullptr) umap[i] = w * sgn(pb - tb);
  }
  block_partials<1>(v, partial);
}
// dp = gout/N * box7^T(u) ; the zero-padded 7x7 mean filter is self-adjoint
__global__ void attn_bwd_kernel(const float* __restrict__ umap, int B, int H, int W, const float* __restrict__ gout, float inv_n,
                                float* __restrict__ dp) {
  pdl_prologue();
  const float g = gout[0] * inv_n;
  const long long n = (long long)B * H * W, hw = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % W), y = int((i / W) % H);
    dp[i] = g * box7(umap + (i / hw) * hw, y, x, H, W);
  }
}

// ------------------------------------------------------------------ L6: ContrastRegionLoss (8x8 patches + global mean/std)
// partial sums: [0] sum_patches m*|pbar-tbar|, [1] sum p, [2] sum p^2, [3] sum t, [4] sum t^2 ; thread = one 8x8 patch
__global__ void region_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ s, int B, int H,
                                  int W, float threshold, float* __restrict__ partial) {
  pdl_prologue();
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const int Hq = H / 8, Wq = W / 8;
  const long long np = (long long)B * Hq * Wq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (long long)gridDim.x * blockDim.x) {
    const int qx = int(i % Wq), qy = int((i / Wq) % Hq);
    const long long base = (i / ((long long)Hq * Wq)) * H * W + (long long)qy * 8 * W + qx * 8;
    float sp = 0.f, st = 0.f, ss = 0.f;
    for (int dy = 0; dy < 8; ++dy) {
      const float4* pr = reinterpret_cast<const float4*>(p + base + (long long)dy * W);
      const float4* tr = reinterpret_cast<const float4*>(t + base + (long long)dy * W);
      const float4* sr = reinterpret_cast<const float4*>(s + base + (long long)dy * W);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 a = pr[h], b = tr[h], c = sr[h];
        sp += (a.x + a.y) + (a.z + a.w);
        st += (b.x + b.y) + (b.z + b.w);
        ss += (c.x + c.y) + (c.z + c.w);
        v[2] += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        v[4] += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
      }
    }
    v[1] += sp;
    v[3] += st;
    const float pb = sp * (1.f / 64.f), tb = st * (1.f / 64.f), sb = ss * (1.f / 64.f);
    const float m = 1.f / (1.f + expf(-5.f * ((tb - sb) - threshold)));
    v[0] += m * fabsf(pb - tb);
  }
  block_partials<5>(v, partial);
}
// state: [0] sgn(mean p - mean t), [1] sgn(std p - std t), [2] mean p, [3] std p
__global__ void region_finish_kernel(const double* __restrict__ sums, double npatch, double n, float weight, float* __restrict__ out,
                                     float* __restrict__ state) {
  pdl_prologue();
  const double mp = sums[1] / n, mt = sums[3] / n;
  const double vp = fmax((sums[2] - n * mp * mp) / (n - 1.0), 0.0), vt = fmax((sums[4] - n * mt * mt) / (n - 1.0), 0.0);
  const double sp = sqrt(vp), st = sqrt(vt);
  out[0] = float(weight * (sums[0] / npatch + 0.5 * (fabs(mp - mt) + fabs(sp - st))));
  state[0] = mp > mt ? 1.f : (mp < mt ? -1.f : 0.f);
  state[1] = sp > st ? 1.f : (sp < st ? -1.f : 0.f);
  state[2] = float(mp);
  state[3] = float(sp);
}
__global__ void region_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ s, int B, int H,
                                  int W, float threshold, float weight, const float* __restrict__ state,
                                  const float* __restrict__ gout, float* __restrict__ dp) {
  pdl_prologue();
  const int Hq = H / 8, Wq = W / 8;
  const long long np = (long long)B * Hq * Wq;
  const double n = double(B) * H * W;
  const float g = gout[0] * weight;
  const float g_mean = g * 0.5f * state[0] / float(n);
  const float g_std = state[3] > 0.f ? g * 0.5f * state[1] / (float(n - 1.0) * state[3]) : 0.f;
  const float mp = state[2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (long long)gridDim.x * blockDim.x) {
    const int qx = int(i % Wq), qy = int((i / Wq) % Hq);
    const long long base = (i / ((long long)Hq * Wq)) * H * W + (long long)qy * 8 * W + qx * 8;
    float sp = 0.f, st = 0.f, ss = 0.f;
    for (int dy = 0; dy < 8; ++dy)
      for (int dx = 0; dx < 8; ++dx) {
        const long long j = base + (long long)dy * W + dx;
        sp += p[j]; st += t[j]; ss += s[j];
      }
    const float pb = sp * (1.f / 64.f), tb = st * (1.f / 64.f), sb = ss * (1.f / 64.f);
    const float m = 1.f / (1.f + expf(-5.f * ((tb - sb) - threshold)));
    const float g_patch = g * m * sgn(pb - tb) / (float(np) * 64.f);
    for (int dy = 0; dy < 8; ++dy)
      for (int dx = 0; dx < 8; ++dx) {
        const long long j = base + (long long)dy * W + dx;
        dp[j] = g_patch + g_mean + g_std * (p[j] - mp);
      }
  }
}

// ------------------------------------------------------------------ L7: ContrastEdgeLoss (Sobel magnitude statistics + top-10 % mean)
__device__ __forceinline__ float px_or0(const float* __restrict__ img, int y, int x, int H, int W) {
  return (y >= 0 && y < H && x >= 0 && x < W) ? img[(long long)y * W + x] : 0.f;
}
__device__ __forceinline__ void sobel(const float* __restrict__ img, int y, int x, int H, int W, float& gx, float& gy) {
  const float a = px_or0(img, y - 1, x - 1, H, W), b = px_or0(img, y - 1, x, H, W), c = px_or0(img, y - 1, x + 1, H, W);
  const float d = px_or0(img, y, x - 1, H, W), f = px_or0(img, y, x + 1, H, W);
  const float g = px_or0(img, y + 1, x - 1, H, W), h = px_or0(img, y + 1, x, H, W), k = px_or0(img, y + 1, x + 1, H, W);
  gx = (c - a) + 2.f * (f - d) + (k - g);
  gy = (g - a) + 2.f * (h - b) + (k - c);
}
// edge maps of pred and target + partial sums [sum e_p, sum e_p^2, sum e_t, sum e_t^2]
__global__ void edge_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t, int B, int H, int W, float* __restrict__ ep,
                                float* __restrict__ et, float* __restrict__ partial) {
  pdl_prologue();
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  const long long n = (long long)B * H * W, hw = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % W), y = int((i / W) % H);
    const long long base = (i / hw) * hw;
    float gx, gy;
    sobel(p + base, y, x, H, W, gx, gy);
    const float e1 = sqrtf(gx * gx + gy * gy + 1e-6f);
    sobel(t + base, y, x, H, W, gx, gy);
    const float e2 = sqrtf(gx * gx + gy * gy + 1e-6f);
    ep[i] = e1;
    et[i] = e2;
    v[0] += e1; v[1] = fmaf(e1, e1, v[1]); v[2] += e2; v[3] = fmaf(e2, e2, v[3]);
  }
  block_partials<4>(v, partial);
}
// exact k-th largest of non-negative floats by radix selection on the bit pattern (monotone for x >= 0): four passes
// of 8 bits.  sel[0] = prefix found so far, sel[1] = remaining rank (1-based from the top) inside the prefix bucket.
__global__ void select_hist_kernel(const float* __restrict__ e, long long n, int pass, const unsigned* __restrict__ sel,
                                   unsigned* __restrict__ hist) {
  pdl_prologue();
  __shared__ unsigned sh[256];
  sh[threadIdx.x] = 0;  // blockDim.x == 256
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const unsigned prefix = sel[0];
  const unsigned mask = pass == 0 ? 0u : (0xFFFFFFFFu << (32 - 8 * pass));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned u = __float_as_uint(e[i]);
    if ((u & mask) == prefix) atomicAdd(&sh[(u >> shift) & 255u], 1u);   // integer counts: order-independent
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void select_pick_kernel(unsigned* __restrict__ hist, int pass, unsigned* __restrict__ sel) {
  pdl_prologue();
  if (threadIdx.x != 0) return;
  unsigned rank = sel[1], acc = 0;
  int b = 255;
  for (; b > 0; --b) {
    if (acc + hist[b] >= rank) break;
    acc += hist[b];
  }
  sel[0] |= unsigned(b) << (24 - 8 * pass);
  sel[1] = rank - acc;
  for (int i = 0; i < 256; ++i) hist[i] = 0;
}
// selection state for both edge maps, set from kernel arguments (a host-to-device copy of a stack array would be read
// again -- from a dead stack frame -- every time a CUDA graph holding it is replayed)
__global__ void select_init_kernel(unsigned* __restrict__ sel, unsigned k) {
  pdl_prologue();
  const int t = threadIdx.x;
  if (t < 4) sel[t] = (t & 1) ? k : 0u;
  for (int i = 4 + t; i < 4 + 256; i += blockDim.x) sel[i] = 0u;
}
// partial sums for the top-k mean: [sum of e > tau, count of e > tau]
__global__ void topk_sum_kernel(const float* __restrict__ e, long long n, const unsigned* __restrict__ sel, float* __restrict__ partial) {
  pdl_prologue();
  const float tau = __uint_as_float(sel[0]);
  float v[2] = {0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (e[i] > tau) { v[0] += e[i]; v[1] += 1.f; }
  block_partials<2>(v, partial);
}
// state: [0] sgn(dmean), [1] sgn(dstd), [2] mean e_p, [3] std e_p, [4] tau_p, [5] sgn(dtopk), [6] topk mean p, [7] topk mean t
__global__ void edge_finish_kernel(const double* __restrict__ s4, const double* __restrict__ tp, const double* __restrict__ tt,
                                   const unsigned* __restrict__ selp, const unsigned* __restrict__ selt, double n, double k,
                                   float* __restrict__ out, float* __restrict__ state) {
  pdl_prologue();
  const double mp = s4[0] / n, mt = s4[2] / n;
  const double sp = sqrt(fmax((s4[1] - n * mp * mp) / (n - 1.0), 0.0)), st = sqrt(fmax((s4[3] - n * mt * mt) / (n - 1.0), 0.0));
  const double taup = double(__uint_as_float(selp[0])), taut = double(__uint_as_float(selt[0]));
  const double kp = (tp[0] + (k - tp[1]) * taup) / k, kt = (tt[0] + (k - tt[1]) * taut) / k;   // ties at tau all equal tau
  out[0] = float(fabs(mp - mt) + fabs(sp - st) + fabs(kp - kt));
  state[0] = mp > mt ? 1.f : (mp < mt ? -1.f : 0.f);
  state[1] = sp > st ? 1.f : (sp < st ? -1.f : 0.f);
  state[2] = float(mp);
  state[3] = float(sp);
  state[4] = float(taup);
  state[5] = kp > kt ? 1.f : (kp < kt ? -1.f : 0.f);
  state[6] = float(kp);
  state[7] = float(kt);
}
// d loss / d e_p (per pixel), then back through e = sqrt(gx^2+gy^2+eps) and the (zero-padded) Sobel pair, gathered per pixel
__device__ __forceinline__ float edge_up(const float* __restrict__ ep, long long j, float g_mean, float g_std, float mp, float g_top,
                                         float tau) {
  const float e = ep[j];
  return g_mean + g_std * (e - mp) + (e >= tau ? g_top : 0.f);
}
__global__ void edge_bwd_kernel(const float* __restrict__ p, const float* __restrict__ ep, int B, int H, int W,
                                const float* __restrict__ state, const float* __restrict__ gout, float inv_n, float inv_nm1,
                                float inv_k, float* __restrict__ dp) {
  pdl_prologue();
  const float g = gout[0];
  const float g_mean = g * state[0] * inv_n;
  const float g_std = state[3] > 0.f ? g * state[1] * inv_nm1 / state[3] : 0.f;
  const float g_top = g * state[5] * inv_k;
  const float mp = state[2], tau = state[4];
  const long long n = (long long)B * H * W, hw = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % W), y = int((i / W) % H);
    const long long base = (i / hw) * hw;
    // pixel (y,x) feeds the Sobel responses of its 8 neighbours q = (y+dy, x+dx): coefficient of p(y,x) in gx(q) is
    // kx[-dy][-dx], in gy(q) ky[-dy][-dx]
    float acc = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        if (dy == 0 && dx == 0) continue;
        const int qy = y + dy, qx = x + dx;
        if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
        float gx, gy;
        sobel(p + base, qy, qx, H, W, gx, gy);
        const long long j = base + (long long)qy * W + qx;
        const float up = edge_up(ep, j, g_mean, g_std, mp, g_top, tau) / ep[j];
        const int ry = -dy, rx = -dx;                                  // position of (y,x) inside q's 3x3 window
        const float cx = float(rx) * (ry == 0 ? 2.f : 1.f);            // sobel_x = [-1 0 1; -2 0 2; -1 0 1]
        const float cy = float(ry) * (rx == 0 ? 2.f : 1.f);            // sobel_y = [-1 -2 -1; 0 0 0; 1 2 1]
        acc += up * (gx * cx + gy * cy);
      }
    dp[i] = acc;
  }
}

// ------------------------------------------------------------------ L4: SSIM (gaussian 11 / 1.5, valid, data_range L)
__constant__ float c_gauss[11];
// horizontal valid pass over the five maps x, y, xx, yy, xy -> tmp[5][B][H][W-10]
__global__ void ssim_hpass_kernel(const float* __restrict__ x, const float* __restrict__ y, int B, int H, int W, float* __restrict__ tmp) {
  pdl_prologue();
  const int Wv = W - 10;
  const long long n = (long long)B * H * Wv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int xv = int(i % Wv);
    const long long row = i / Wv;
    const float* xr = x + row * W + xv;
    const float* yr = y + row * W + xv;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = c_gauss[k], xa = xr[k], ya = yr[k];
      a = fmaf(w, xa, a); b = fmaf(w, ya, b); aa = fmaf(w, xa * xa, aa); bb = fmaf(w, ya * ya, bb); ab = fmaf(w, xa * ya, ab);
    }
    tmp[i] = a; tmp[n + i] = b; tmp[2 * n + i] = aa; tmp[3 * n + i] = bb; tmp[4 * n + i] = ab;
  }
}
// vertical valid pass + SSIM map; partial sum of the map; optionally the three derivative maps for the backward:
//   dmu1 = dS/d(mu1), dxx = dS/d(E[xx]), dxy = dS/d(E[xy])   (mu2, E[yy] belong to the target: no gradient needed)
__global__ void ssim_vpass_kernel(const float* __restrict__ tmp, int B, int H, int W, float C1, float C2, float* __restrict__ dmaps,
                                  float* __restrict__ partial) {
  pdl_prologue();
  const int Wv = W - 10, Hv = H - 10;
  const long long nh = (long long)B * H * Wv, n = (long long)B * Hv * Wv;
  float v[1] = {0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int xv = int(i % Wv), yv = int((i / Wv) % Hv);
    const long long b = i / ((long long)Hv * Wv);
    const long long src = (b * H + yv) * Wv + xv;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = c_gauss[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = fmaf(w, tmp[q * nh + src + (long long)k * Wv], m[q]);
    }
    const float mu1 = m[0], mu2 = m[1];
    const float s11 = m[2] - mu1 * mu1, s22 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
    const float A1 = 2.f * mu1 * mu2 + C1, A2 = 2.f * s12 + C2, B1 = mu1 * mu1 + mu2 * mu2 + C1, B2 = s11 + s22 + C2;
    const float S = (A1 * A2) / (B1 * B2);
    v[0] += S;
    if (dmaps != nullptr) {
      // S = A1*A2/(B1*B2); in terms of (mu1, Exx, Exy): s11 = Exx - mu1^2, s12 = Exy - mu1*mu2
      const float dS_dA1 = A2 / (B1 * B2), dS_dA2 = A1 / (B1 * B2), dS_dB1 = -S / B1, dS_dB2 = -S / B2;
      const float dExx = dS_dB2;                       // dB2/dExx = 1
      const float dExy = 2.f * dS_dA2;                 // dA2/dExy = 2
      const float dmu = dS_dA1 * 2.f * mu2 + dS_dB1 * 2.f * mu1 + dS_dA2 * (-2.f * mu2) + dS_dB2 * (-2.f * mu1);
      dmaps[i] = dmu; dmaps[n + i] = dExx; dmaps[2 * n + i] = dExy;
    }
  }
  block_partials<1>(v, partial);
}
// adjoint of the vertical valid pass: [B][Hv][Wv] -> [B][H][Wv] for the three derivative maps
__global__ void ssim_vadj_kernel(const float* __restrict__ dmaps, int B, int H, int W, float* __restrict__ tmp) {
  pdl_prologue();
  const int Wv = W - 10, Hv = H - 10;
  const long long n = (long long)B * Hv * Wv, nh = (long long)B * H * Wv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nh; i += (long long)gridDim.x * blockDim.x) {
    const int xv = int(i % Wv), yy = int((i / Wv) % H);
    const long long b = i / ((long long)H * Wv);
    float a[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const int yv = yy - k;
      if (yv < 0 || yv >= Hv) continue;
      const long long src = (b * Hv + yv) * Wv + xv;
      const float w = c_gauss[k];
#pragma unroll
      for (int q = 0; q < 3; ++q) a[q] = fmaf(w, dmaps[q * n + src], a[q]);
    }
    tmp[i] = a[0]; tmp[nh + i] = a[1]; tmp[2 * nh + i] = a[2];
  }
}
// adjoint of the horizontal pass + chain rule: dx = g/Nmap * (Gt(dmu) + 2x*Gt(dExx) + y*Gt(dExy))
__global__ void ssim_hadj_kernel(const float* __restrict__ tmp, const float* __restrict__ x, const float* __restrict__ y, int B, int H,
                                 int W, const float* __restrict__ gout, float scale, float* __restrict__ dx) {
  pdl_prologue();
  const int Wv = W - 10;
  const long long n = (long long)B * H * W, nh = (long long)B * H * Wv;
  const float g = gout[0] * scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int xx = int(i % W);
    const long long row = i / W;
    float a[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const int xv = xx - k;
      if (xv < 0 || xv >= Wv) continue;
      const float w = c_gauss[k];
#pragma unroll
      for (int q = 0; q < 3; ++q) a[q] = fmaf(w, tmp[q * nh + row * Wv + xv], a[q]);
    }
    dx[i] = g * (a[0] + 2.f * x[i] * a[1] + y[i] * a[2]);
  }
}
__global__ void ssim_finish_kernel(const double* __restrict__ sums, double inv_n, float* __restrict__ out) {
  pdl_prologue(); out[0] = float(sums[0] * inv_n); }

int sum_partials(const float* partial, int blocks, int Q, double* sums, cudaStream_t st) {
  pdl(sum_partials_kernel, 1, 32 * Q, 0, st)(partial, blocks, Q, sums);
  return check_launch("sum_partials_kernel");
}
int ensure_gauss() {
  static bool ready[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && ready[dev]) return 0;
  // pytorch_msssim builds the 1-D window in fp32: exp(-(c^2)/(2 sigma^2)) / sum, size 11, sigma 1.5
  float gf[11], sf = 0.f;
  for (int i = 0; i < 11; ++i) { gf[i] = expf(-float((i - 5) * (i - 5)) / (2.f * 1.5f * 1.5f)); sf += gf[i]; }
  for (int i = 0; i < 11; ++i) gf[i] /= sf;
  if (cudaMemcpyToSymbol(c_gauss, gf, sizeof(gf)) != cudaSuccess) return fail(DUCOSY_ERR_CUDA, "ssim: cannot upload the window");
  if (dev >= 0 && dev < 64) ready[dev] = true;
  return 0;
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

// scratch layout (floats): [0, kMaxQ*kLossMaxBlocks) block partials | then 64 doubles of sums | 512 unsigned of select state
extern "C" size_t ducosy_loss_scratch_bytes(void) { return size_t(kMaxQ) * kLossMaxBlocks * 4 + 64 * 8 + 1024 * 4; }
static double* sums_of(float* scratch) { return reinterpret_cast<double*>(scratch + kMaxQ * kLossMaxBlocks); }
static unsigned* sel_of(float* scratch) { return reinterpret_cast<unsigned*>(sums_of(scratch) + 64); }

extern "C" int ducosy_loss_l1_forward(const float* a, const float* b, long long n, float* loss_out, float* scratch, ducosy_stream_t stream) {
  DUCOSY_CHECK(a && b && loss_out && scratch && n > 0, DUCOSY_ERR_ARG, "loss_l1_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = loss_grid(n);
  pdl(l1_fwd_kernel, g, kLossThreads, 0, st)(a, b, n, scratch);
  DUCOSY_TRY(sum_partials(scratch, g, 1, sums_of(scratch), st));
  pdl(mean_finish_kernel, 1, 1, 0, st)(sums_of(scratch), 1.0 / double(n), loss_out);
  return check_launch("loss_l1_forward");
}
extern "C" int ducosy_loss_l1_backward(const float* a, const float* b, long long n, const float* gout, float* da, ducosy_stream_t stream) {
  DUCOSY_CHECK(a && b && gout && da && n > 0, DUCOSY_ERR_ARG, "loss_l1_backward: bad argument");
  pdl(l1_bwd_kernel, loss_grid(n), kLossThreads, 0, (cudaStream_t)stream)(a, b, n, gout, float(1.0 / double(n)), da);
  return check_launch("loss_l1_backward");
}
extern "C" int ducosy_loss_mse_const_forward(const float* a, float target, long long n, float* loss_out, float* scratch, ducosy_stream_t stream) {
  DUCOSY_CHECK(a && loss_out && scratch && n > 0, DUCOSY_ERR_ARG, "loss_mse_const_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = loss_grid(n);
  pdl(mse_const_fwd_kernel, g, kLossThreads, 0, st)(a, target, n, scratch);
  DUCOSY_TRY(sum_partials(scratch, g, 1, sums_of(scratch), st));
  pdl(mean_finish_kernel, 1, 1, 0, st)(sums_of(scratch), 1.0 / double(n), loss_out);
  return check_launch("loss_mse_const_forward");
}
extern "C" int ducosy_loss_mse_const_backward(const float* a, float target, long long n, const float* gout, float* da, ducosy_stream_t stream) {
  DUCOSY_CHECK(a && gout && da && n > 0, DUCOSY_ERR_ARG, "loss_mse_const_backward: bad argument");
  pdl(mse_const_bwd_kernel, loss_grid(n), kLossThreads, 0, (cudaStream_t)stream)(a, target, n, gout, float(1.0 / double(n)), da);
  return check_launch("loss_mse_const_backward");
}

extern "C" int ducosy_loss_gradient_forward(const float* pred, const float* target, int B, int H, int W, float* loss_out, float* scratch,
                                            ducosy_stream_t stream) {
  DUCOSY_CHECK(pred && target && loss_out && scratch && B > 0 && H > 1 && W > 1, DUCOSY_ERR_ARG, "loss_gradient_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * H * W;
  const int g = loss_grid(n);
  pdl(gradloss_fwd_kernel, g, kLossThreads, 0, st)(pred, target, B, H, W, scratch);
  DUCOSY_TRY(sum_partials(scratch, g, 2, sums_of(scratch), st));
  pdl(gradloss_finish_kernel, 1, 1, 0, st)(sums_of(scratch), 1.0 / (double(B) * (H - 1) * W), 1.0 / (double(B) * H * (W - 1)), loss_out);
  return check_launch("loss_gradient_forward");
}
extern "C" int ducosy_loss_gradient_backward(const float* pred, const float* target, int B, int H, int W, const float* gout, float* dpred,
                                             ducosy_stream_t stream) {
  DUCOSY_CHECK(pred && target && gout && dpred && B > 0 && H > 1 && W > 1, DUCOSY_ERR_ARG, "loss_gradient_backward: bad argument");
  pdl(gradloss_bwd_kernel, loss_grid((long long)B * H * W), kLossThreads, 0, (cudaStream_t)stream)(
      pred, target, B, H, W, gout, float(1.0 / (double(B) * (H - 1) * W)), float(1.0 / (double(B) * H * (W - 1))), dpred);
  return check_launch("loss_gradient_backward");
}

// ContrastAttentionLoss(sigma, min_weight, max_weight, blur_kernel = 7): umap [B*H*W] is kept for the backward
extern "C" int ducosy_loss_contrast_attention_forward(const float* pred, const float* target, const float* source, int B, int H, int W,
                                                      float sigma, float min_w, float max_w, float* loss_out, float* umap,
                                                      float* scratch, ducosy_stream_t stream) {
  DUCOSY_CHECK(pred && target && source && loss_out && scratch && B > 0 && sigma > 0.f, DUCOSY_ERR_ARG, "loss_contrast_attention_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * H * W;
  const int g = loss_grid(n);
  pdl(attn_fwd_kernel, g, kLossThreads, 0, st)(pred, target, source, B, H, W, 1.f / sigma, min_w, max_w, umap, scratch);
  DUCOSY_TRY(sum_partials(scratch, g, 1, sums_of(scratch), st));
  pdl(mean_finish_kernel, 1, 1, 0, st)(sums_of(scratch), 1.0 / double(n), loss_out);
  return check_launch("loss_contrast_attention_forward");
}
extern "C" int ducosy_loss_contrast_attention_backward(const float* umap, int B, int H, int W, const float* gout, float* dpred,
                                                       ducosy_stream_t stream) {
  DUCOSY_CHECK(umap && gout && dpred && B > 0, DUCOSY_ERR_ARG, "loss_contrast_attention_backward: bad argument");
  const long long n = (long long)B * H * W;
  pdl(attn_bwd_kernel, loss_grid(n), kLossThreads, 0, (cudaStream_t)stream)(umap, B, H, W, gout, float(1.0 / double(n)), dpred);
  return check_launch("loss_contrast_attention_backward");
}

// ContrastRegionLoss(threshold, weight): state[4] is kept for the backward
extern "C" int ducosy_loss_contrast_region_forward(const float* pred, const float* target, const float* source, int B, int H, int W,
                                                   float threshold, float weight, float* loss_out, float* state, float* scratch,
                                                   ducosy_stream_t stream) {
  DUCOSY_CHECK(pred && target && source && loss_out && state && scratch && B > 0, DUCOSY_ERR_ARG, "loss_contrast_region_forward: bad argument");
  DUCOSY_CHECK(H % 8 == 0 && W % 8 == 0, DUCOSY_ERR_SHAPE, "loss_contrast_region_forward: H and W must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const long long np = (long long)B * (H / 8) * (W / 8);
  const int g = loss_grid(np);
  pdl(region_fwd_kernel, g, kLossThreads, 0, st)(pred, target, source, B, H, W, threshold, scratch);
  DUCOSY_TRY(sum_partials(scratch, g, 5, sums_of(scratch), st));
  pdl(region_finish_kernel, 1, 1, 0, st)(sums_of(scratch), double(np), double(B) * H * W, weight, loss_out, state);
  return check_launch("loss_contrast_region_forward");
}
extern "C" int ducosy_loss_contrast_region_backward(const float* pred, const float* target, const float* source, int B, int H, int W,
                                                    float threshold, float weight, const float* state, const float* gout,
                                                    float* dpred, ducosy_stream_t stream) {
  DUCOSY_CHECK(pred && target && source && state && gout && dpred && B > 0, DUCOSY_ERR_ARG, "loss_contrast_region_backward: bad argument");
  const long long np = (long long)B * (H / 8) * (W / 8);
  pdl(region_bwd_kernel, loss_grid(np), kLossThreads, 0, (cudaStream_t)stream)(pred, target, source, B, H, W, threshold, weight, state, gout, dpred);
  return check_launch("loss_contrast_region_backward");
}

// ContrastEdgeLoss: edge maps ep/et [B*H*W] and state[8] are kept for the backward
extern "C" int ducosy_loss_contrast_edge_forward(const float* pred, const float* target, int B, int H, int W, float* loss_out, float* ep,
                                                 float* et, float* state, float* scratch, ducosy_stream_t stream) {
  DUCOSY_CHECK(pred && target && loss_out && ep && et && state && scratch && B > 0, DUCOSY_ERR_ARG, "loss_contrast_edge_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * H * W;
  const long long k = (long long)(double(n) * 0.1);   // int(numel * 0.1), trainer.py:178-180
  DUCOSY_CHECK(k >= 1, DUCOSY_ERR_SHAPE, "loss_contrast_edge_forward: tensor too small for a top-10 %% set");
  const int g = loss_grid(n);
  double* sums = sums_of(scratch);
  unsigned* sel = sel_of(scratch);   // [0..1] sel_p, [2..3] sel_t, [4..259] hist
  pdl(edge_fwd_kernel, g, kLossThreads, 0, st)(pred, target, B, H, W, ep, et, scratch);
  DUCOSY_TRY(sum_partials(scratch, g, 4, sums, st));
  pdl(select_init_kernel, 1, 256, 0, st)(sel, unsigned(k));
  for (int which = 0; which < 2; ++which) {
    const float* e = which == 0 ? ep : et;
    for (int pass = 0; pass < 4; ++pass) {
      pdl(select_hist_kernel, g, 256, 0, st)(e, n, pass, sel + 2 * which, sel + 4);
      pdl(select_pick_kernel, 1, 32, 0, st)(sel + 4, pass, sel + 2 * which);
    }
    pdl(topk_sum_kernel, g, kLossThreads, 0, st)(e, n, sel + 2 * which, scratch);
    DUCOSY_TRY(sum_partials(scratch, g, 2, sums + 8 + 4 * which, st));
  }
  pdl(edge_finish_kernel, 1, 1, 0, st)(sums, sums + 8, sums + 12, sel, sel + 2, double(n), double(k), loss_out, state);
  return check_launch("loss_contrast_edge_forward");
}
extern "C" int ducosy_loss_contrast_edge_backward(const float* pred, const float* ep, int B, int H, int W, const float* state,
                                                  const float* gout, float* dpred, ducosy_stream_t stream) {
  DUCOSY_CHECK(pred && ep && state && gout && dpred && B > 0, DUCOSY_ERR_ARG, "loss_contrast_edge_backward: bad argument");
  const long long n = (long long)B * H * W;
  const long long k = (long long)(double(n) * 0.1);
  pdl(edge_bwd_kernel, loss_grid(n), kLossThreads, 0, (cudaStream_t)stream)(pred, ep, B, H, W, state, gout, float(1.0 / double(n)),
                                                                            float(1.0 / double(n - 1)), float(1.0 / double(k)), dpred);
  return check_launch("loss_contrast_edge_backward");
}

// SSIM(x, y) mean (pytorch_msssim.SSIM(data_range, size_average=True, channel=1), win 11 / sigma 1.5): returns the SSIM
// value (the reference uses 1 - SSIM).  tmp: 5*B*H*(W-10) floats; dmaps: 3*B*(H-10)*(W-10) floats (NULL = no backward).
extern "C" int ducosy_loss_ssim_forward(const float* x, const float* y, int B, int H, int W, float data_range, float* ssim_out, float* tmp,
                                        float* dmaps, float* scratch, ducosy_stream_t stream) {
  DUCOSY_CHECK(x && y && ssim_out && tmp && scratch && B > 0 && H > 10 && W > 10, DUCOSY_ERR_ARG, "loss_ssim_forward: bad argument");
  DUCOSY_TRY(ensure_gauss());
  cudaStream_t st = (cudaStream_t)stream;
  const long long nv = (long long)B * (H - 10) * (W - 10);
  pdl(ssim_hpass_kernel, loss_grid((long long)B * H * (W - 10)), kLossThreads, 0, st)(x, y, B, H, W, tmp);
  const int g = loss_grid(nv);
  const float C1 = (0.01f * data_range) * (0.01f * data_range), C2 = (0.03f * data_range) * (0.03f * data_range);
  pdl(ssim_vpass_kernel, g, kLossThreads, 0, st)(tmp, B, H, W, C1, C2, dmaps, scratch);
  DUCOSY_TRY(sum_partials(scratch, g, 1, sums_of(scratch), st));
  pdl(ssim_finish_kernel, 1, 1, 0, st)(sums_of(scratch), 1.0 / double(nv), ssim_out);
  return check_launch("loss_ssim_forward");
}
// dx = gout * d mean(SSIM) / dx  (gradient w.r.t. the first image only)
extern "C" int ducosy_loss_ssim_backward(const float* x, const float* y, const float* dmaps, int B, int H, int W, const float* gout, float* tmp,
                                         float* dx, ducosy_stream_t stream) {
  DUCOSY_CHECK(x && y && dmaps && gout && tmp && dx && B > 0, DUCOSY_ERR_ARG, "loss_ssim_backward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nv = (long long)B * (H - 10) * (W - 10);
  pdl(ssim_vadj_kernel, loss_grid((long long)B * H * (W - 10)), kLossThreads, 0, st)(dmaps, B, H, W, tmp);
  pdl(ssim_hadj_kernel, loss_grid((long long)B * H * W), kLossThreads, 0, st)(tmp, x, y, B, H, W, gout, float(1.0 / double(nv)), dx);
  return check_launch("loss_ssim_backward");
}
