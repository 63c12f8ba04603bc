// extern "C" surface: error plumbing, convolution entry points and the whole-generator forward
// (modules/model.py:92-115) expressed as a fixed sequence of the kernels in this directory.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace ducosy {

namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DUCOSY_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}
// Programmatic dependent launch is OPT-IN (DUCOSY_PDL=1).  Same-box A/B (profiles/r02_pdl_ab.json): the eager batch-1 forward
// gains 14 % (2.07 -> 1.78 ms per dual-HU slice), a CUDA-graph replay of the same chain gains nothing (1.708 -> 1.703 ms: graph
// launches already have no gap to hide), and every two-stream path LOSES -- early-resident CTAs spinning at their
// griddepcontrol.wait hold SM slots the other stream's kernels would have used: train step batch 1 19.07 -> 20.98 ms, batch 8
// 108.2 -> 111.3 ms, batch-30 synthesis 958 -> 944 slices/s.
bool pdl_enabled() {
  static const bool on = []() { const char* e = getenv("DUCOSY_PDL"); return e != nullptr && atoi(e) != 0; }();
  return on;
}
int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

namespace {

int check_device_cached() {
  static int ok[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(DUCOSY_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (dev >= 0 && dev < 64 && ok[dev]) return 0;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) return fail(DUCOSY_ERR_ARCH, "device %d is sm_%d%d; libducosy_sm100 needs sm_100 (B200)", dev, major, minor);
  if (dev >= 0 && dev < 64) ok[dev] = 1;
  return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool fused_spatial() {
  static const bool on = []() { const char* e = getenv("DUCOSY_FUSED_SPATIAL"); return e != nullptr && atoi(e) != 0; }();
  return on;
}

void fill_taps_3x3(ConvPlan& p) {
  p.num_phases = 1;
  p.num_taps = 9;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      p.tap_dy[0][r * 3 + s] = int8_t(r);
      p.tap_dx[0][r * 3 + s] = int8_t(s);
    }
}

// ------------------------------------------------------------------ generator layout
struct GenLayout {
  int Kpad = 0;
  size_t stem = 0, d1 = 0, d2 = 0, up1 = 0, up2 = 0, outw = 0, outb = 0;
  std::vector<size_t> c1, c2, fc0, fc2, saw;
  size_t total = 0;
};

GenLayout make_layout(const ducosy_gen_config& c) {
  GenLayout L;
  L.Kpad = (49 * c.input_channels + 63) / 64 * 64;
  const size_t e = c.dtype == DUCOSY_F16X2 ? 4 : 2;   // bytes per packed conv weight (split-operand mode: a hi and a lo value)
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  L.stem = take(size_t(64) * L.Kpad * e);
  L.d1 = take(size_t(128) * 9 * 64 * e);
  L.d2 = take(size_t(256) * 9 * 128 * e);
  for (int i = 0; i < c.num_residual_blocks; ++i) {
    L.c1.push_back(take(size_t(256) * 9 * 256 * e));
    L.c2.push_back(take(size_t(256) * 9 * 256 * e));
    if (c.use_cbam) {
      L.fc0.push_back(take(16 * 256 * 4));
      L.fc2.push_back(take(256 * 16 * 4));
      L.saw.push_back(take(98 * 4));
    }
  }
  L.up1 = take(size_t(4) * 128 * 4 * 256 * e);
  L.up2 = take(size_t(4) * 64 * 9 * 128 * e);  // merged-phase packing (N = 256, nine source offsets); split mode: four 2x2 phases
  L.outw = take(7 * 8 * 64 * e);
  L.outb = take(4);
  L.total = off;
  return L;
}

struct GenWorkspace {
  size_t a_stem, y0, p0, y1, p1, y2a, y2b, pa, pb, pc, p_out, partials, scale, shift, chmax, pooled, sa, tickets, total;
};

GenWorkspace make_workspace(const ducosy_gen_config& c, int B, int H, int W) {
  GenWorkspace w{};
  const GenLayout L = make_layout(c);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  const size_t HW = size_t(H) * W;
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  const size_t e = c.dtype == DUCOSY_F16X2 ? 4 : 2;   // bytes per activation (split-operand mode: hi and lo plane per pixel)
  w.a_stem = take(size_t(B) * HW * L.Kpad * e);
  w.y0 = take(size_t(B) * HW * 64 * e);
  w.p0 = take(size_t(B) * (H + 2) * (W + 2) * 64 * e);
  w.y1 = take(size_t(B) * H2 * W2 * 128 * e);
  w.p1 = take(size_t(B) * (H2 + 2) * (W2 + 2) * 128 * e);
  w.y2a = take(size_t(B) * H4 * W4 * 256 * e);
  w.y2b = take(size_t(B) * H4 * W4 * 256 * e);
  const size_t padded = size_t(B) * (H4 + 2) * (W4 + 2) * 256 * e;
  w.pa = take(padded);
  w.pb = take(padded);
  w.pc = take(padded);
  w.p_out = take(size_t(B) * (H + 6) * (W + 6) * 64 * 2);
  w.partials = take(size_t(B) * (HW / 128) * 3 * 64 * 4);
  w.scale = take(size_t(B) * 256 * 4);
  w.shift = take(size_t(B) * 256 * 4);
  w.chmax = take(size_t(B) * 256 * 4);
  w.pooled = take(size_t(B) * H4 * W4 * 2 * 4);
  w.sa = take(size_t(B) * H4 * W4 * 4);
  w.tickets = take(size_t(B) * 4);
  w.total = off;
  return w;
}

int check_gen_shape(const ducosy_gen_config& c, int B, int H, int W) {
  DUCOSY_CHECK(c.input_channels >= 1 && c.input_channels <= 16, DUCOSY_ERR_SHAPE, "generator: input_channels %d unsupported", c.input_channels);
  DUCOSY_CHECK(c.num_residual_blocks >= 0 && c.num_residual_blocks <= 64, DUCOSY_ERR_SHAPE, "generator: num_residual_blocks out of range");
  DUCOSY_CHECK(c.dtype == DUCOSY_F16 || c.dtype == DUCOSY_BF16 || c.dtype == DUCOSY_F16X2, DUCOSY_ERR_ARG,
               "generator: dtype must be DUCOSY_F16, DUCOSY_BF16 or DUCOSY_F16X2");
  DUCOSY_CHECK(B >= 1, DUCOSY_ERR_SHAPE, "generator: batch must be >= 1");
  DUCOSY_CHECK(H >= 128 && W >= 128 && W % 128 == 0 && H % 32 == 0, DUCOSY_ERR_SHAPE,
               "generator: H must be a multiple of 32 and W a multiple of 128, both >= 128 (got %dx%d)", H, W);
  const int W4 = W / 4, H4 = H / 4;
  const int Wt = W4 < 128 ? W4 : 128;
  DUCOSY_CHECK((Wt & (Wt - 1)) == 0 && W4 % Wt == 0 && H4 % (128 / Wt) == 0, DUCOSY_ERR_SHAPE,
               "generator: W/4 must be 32, 64 or a multiple of 128, and H/4 a multiple of 128/min(W/4,128) (got %dx%d)", H, W);
  return 0;
}

// conv + InstanceNorm statistics + finalize (+ CBAM channel MLP) in ONE launch: the CTA that completes a sample's last tile
// reduces the per-tile partials (conv_gemm.cu: finalize_sample)
int run_conv_in(ConvPlan& p, float* scale, float* shift, const float* fc0, const float* fc2, float* chmax, int* tickets,
                int npix, cudaStream_t st) {
  // DUCOSY_FUSED_FINALIZE=1 finalizes inside the conv launch (ConvFinalize).  Measured and NOT the default: one CTA reducing a
  // whole sample's partials (393 KB through one SM's L2 port, ~10 us) is slower than the 8-CTA finalize kernel plus its launch
  // -- 967 -> 752 slices/s at batch 30 (the finalizing CTA pair stalls and the static tile schedule turns that into a tail),
  // 713 -> 610 at batch 1 (profiles/r02_fusion_ab.json).
  static const bool fused = []() { const char* e = getenv("DUCOSY_FUSED_FINALIZE"); return e != nullptr && atoi(e) != 0; }();
  if (!fused) {
    DUCOSY_TRY(launch_conv_gemm(p, st));
    const int tiles = conv_tiles_per_sample(p.num_phases, p.Hg, p.Wg);
    return ducosy_in_finalize(p.partials, tiles, npix, scale, shift, fc0, fc2, chmax, p.B, p.Cout / (p.fold > 1 ? p.fold : 1), st);
  }
  p.fin.scale = scale; p.fin.shift = shift; p.fin.chmax = chmax; p.fin.fc0 = fc0; p.fin.fc2 = fc2;
  p.fin.counter = tickets; p.fin.npix = npix;
  return launch_conv_gemm(p, st);
}

void upconv_plan(ConvPlan& p, const void* in_pad, const void* w_packed4, void* out, float* partials, int B, int Hs, int Ws,
                 int Cin, int Cout, int dtype) {
  p.in = in_pad; p.B = B; p.Hp = Hs + 2; p.Wp = Ws + 2; p.Cin = Cin; p.stride = 1;
  p.w = w_packed4; p.Cout = Cout; p.num_phases = 4; p.num_taps = 4;
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    for (int t = 0; t < 4; ++t) {
      p.tap_dy[ph][t] = int8_t((t >> 1) + py);   // source row i + a + py - 1, +1 for the zero border
      p.tap_dx[ph][t] = int8_t((t & 1) + px);
    }
    p.oy_off[ph] = int8_t(py);
    p.ox_off[ph] = int8_t(px);
  }
  p.Hg = Hs; p.Wg = Ws; p.out = out; p.Ho = 2 * Hs; p.Wo = 2 * Ws; p.oy_mul = p.ox_mul = 2;
  p.partials = partials; p.dtype = dtype;
}

void upconv_merged_plan(ConvPlan& p, const void* in_pad, const void* w_merged, void* out, float* partials, int B, int Hs,
                        int Ws, int Cin, int Cout, int dtype) {
  p.in = in_pad; p.B = B; p.Hp = Hs + 2; p.Wp = Ws + 2; p.Cin = Cin; p.stride = 1;
  p.w = w_merged; p.Cout = 4 * Cout; p.fold = 4; fill_taps_3x3(p);
  p.Hg = Hs; p.Wg = Ws; p.out = out; p.Ho = 2 * Hs; p.Wo = 2 * Ws; p.oy_mul = p.ox_mul = 2;
  p.partials = partials; p.dtype = dtype;
}

int generator_forward_impl(const ducosy_gen_config& c, const void* packed, const float* x, const int16_t* px,
                           float slope, float intercept, float lo, float hi, float* out, int B, int H, int W, void* ws,
                           size_t ws_bytes, cudaStream_t st) {
  DUCOSY_CHECK(packed && out && ws && (x || px), DUCOSY_ERR_ARG, "generator_forward: null pointer");
  DUCOSY_TRY(check_device_cached());
  DUCOSY_TRY(check_gen_shape(c, B, H, W));
  DUCOSY_CHECK(px == nullptr || c.input_channels == 1, DUCOSY_ERR_SHAPE, "generator_forward_hu: needs input_channels == 1");
  DUCOSY_CHECK((reinterpret_cast<uintptr_t>(ws) & 1023) == 0 && (reinterpret_cast<uintptr_t>(packed) & 255) == 0,
               DUCOSY_ERR_ALIGN, "generator_forward: workspace must be 1024-byte and packed weights 256-byte aligned");
  const GenLayout L = make_layout(c);
  const GenWorkspace w = make_workspace(c, B, H, W);
  DUCOSY_CHECK(ws_bytes >= w.total, DUCOSY_ERR_WORKSPACE, "generator_forward: workspace %zu < required %zu bytes", ws_bytes, w.total);
  uint8_t* base = static_cast<uint8_t*>(ws);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  auto P = [&](size_t off) { return static_cast<void*>(base + off); };
  float* partials = reinterpret_cast<float*>(base + w.partials);
  float* scale = reinterpret_cast<float*>(base + w.scale);
  float* shift = reinterpret_cast<float*>(base + w.shift);
  float* chmax = reinterpret_cast<float*>(base + w.chmax);
  float* pooled = reinterpret_cast<float*>(base + w.pooled);
  int* tickets = reinterpret_cast<int*>(base + w.tickets);
  DUCOSY_CHECK(cudaMemsetAsync(tickets, 0, size_t(B) * 4, st) == cudaSuccess, DUCOSY_ERR_CUDA, "generator_forward: cudaMemsetAsync failed");
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  const int nb = c.num_residual_blocks;
  const int dt = c.dtype;

  // ---- stem: reflect-pad 3 + 7x7 conv, IN, ReLU   modules/model.py:94
  if (c.input_channels == 1 && dt != DUCOSY_F16X2) {
    // fused two-pass stem (stem.cu): statistics pass, then conv + normalise + ReLU written once, zero-padded
    DUCOSY_TRY(ducosy_stem_prepare(x, px, slope, intercept, lo, hi, P(w.a_stem), B, H, W, dt, st));
    DUCOSY_TRY(ducosy_stem_fused(P(w.a_stem), pk + L.stem, partials, nullptr, nullptr, nullptr, B, H, W, 0, dt, st));
    DUCOSY_TRY(ducosy_in_finalize(partials, H, H * W, scale, shift, nullptr, nullptr, nullptr, B, 64, st));
    DUCOSY_TRY(ducosy_stem_fused(P(w.a_stem), pk + L.stem, nullptr, scale, shift, P(w.p0), B, H, W, 1, dt, st));
  } else {
    // general input_channels: (im2col) x (weights) on the tensor-core GEMM
    if (px != nullptr) {
        DUCOSY_TRY(ducosy_stem_im2col_hu(px, P(w.a_stem), B, H, W, slope, intercept, lo, hi, dt, st));
    } else {
      DUCOSY_TRY(ducosy_stem_im2col(x, P(w.a_stem), B, c.input_channels, H, W, dt, st));
    }
    {
      ConvPlan p{};
      p.in = P(w.a_stem); p.B = B; p.Hp = H; p.Wp = W; p.Cin = L.Kpad; p.stride = 1;
      p.w = pk + L.stem; p.Cout = 64; p.num_phases = 1; p.num_taps = 1;
      p.Hg = H; p.Wg = W; p.out = P(w.y0); p.Ho = H; p.Wo = W; p.oy_mul = p.ox_mul = 1;
      p.partials = partials; p.dtype = dt;
      DUCOSY_TRY(run_conv_in(p, scale, shift, nullptr, nullptr, nullptr, tickets, H * W, st));
      DUCOSY_TRY(ducosy_in_apply_pad(P(w.y0), scale, shift, P(w.p0), B, H, W, 64, 1, DUCOSY_PAD_ZERO, DUCOSY_ACT_RELU, dt, st));
    }
  }
  // ---- down 1: 3x3 s2 p1 64 -> 128, IN, ReLU   modules/model.py:96-98
  {
    ConvPlan p{};
    p.in = P(w.p0); p.B = B; p.Hp = H + 2; p.Wp = W + 2; p.Cin = 64; p.stride = 2;
    p.w = pk + L.d1; p.Cout = 128; fill_taps_3x3(p);
    p.Hg = H2; p.Wg = W2; p.out = P(w.y1); p.Ho = H2; p.Wo = W2; p.oy_mul = p.ox_mul = 1;
    p.partials = partials; p.dtype = dt;
    DUCOSY_TRY(run_conv_in(p, scale, shift, nullptr, nullptr, nullptr, tickets, H2 * W2, st));
    DUCOSY_TRY(ducosy_in_apply_pad(P(w.y1), scale, shift, P(w.p1), B, H2, W2, 128, 1, DUCOSY_PAD_ZERO, DUCOSY_ACT_RELU, dt, st));
  }
  // ---- down 2: 3x3 s2 p1 128 -> 256, IN, ReLU
  size_t cur = w.pa, nxt = w.pb;
  {
    ConvPlan p{};
    p.in = P(w.p1); p.B = B; p.Hp = H2 + 2; p.Wp = W2 + 2; p.Cin = 128; p.stride = 2;
    p.w = pk + L.d2; p.Cout = 256; fill_taps_3x3(p);
    p.Hg = H4; p.Wg = W4; p.out = P(w.y2a); p.Ho = H4; p.Wo = W4; p.oy_mul = p.ox_mul = 1;
    p.partials = partials; p.dtype = dt;
    DUCOSY_TRY(run_conv_in(p, scale, shift, nullptr, nullptr, nullptr, tickets, H4 * W4, st));
    DUCOSY_TRY(ducosy_in_apply_pad(P(w.y2a), scale, shift, P(cur), B, H4, W4, 256, 1,
                                   nb > 0 ? DUCOSY_PAD_REFLECT : DUCOSY_PAD_ZERO, DUCOSY_ACT_RELU, dt, st));
  }
  // ---- residual blocks (modules/model.py:56-87): x + CBAM(IN(conv(refpad(ReLU(IN(conv(refpad(x))))))))
  for (int i = 0; i < nb; ++i) {
    ConvPlan p{};
    p.in = P(cur); p.B = B; p.Hp = H4 + 2; p.Wp = W4 + 2; p.Cin = 256; p.stride = 1;
    p.w = pk + L.c1[i]; p.Cout = 256; fill_taps_3x3(p);
    p.Hg = H4; p.Wg = W4; p.out = P(w.y2a); p.Ho = H4; p.Wo = W4; p.oy_mul = p.ox_mul = 1;
    p.partials = partials; p.dtype = dt;
    DUCOSY_TRY(run_conv_in(p, scale, shift, nullptr, nullptr, nullptr, tickets, H4 * W4, st));
    DUCOSY_TRY(ducosy_in_apply_pad(P(w.y2a), scale, shift, P(w.pc), B, H4, W4, 256, 1, DUCOSY_PAD_REFLECT, DUCOSY_ACT_RELU, dt, st));
    p.in = P(w.pc); p.w = pk + L.c2[i]; p.out = P(w.y2b);
    const float* fc0 = c.use_cbam ? reinterpret_cast<const float*>(pk + L.fc0[i]) : nullptr;
    const float* fc2 = c.use_cbam ? reinterpret_cast<const float*>(pk + L.fc2[i]) : nullptr;
    DUCOSY_TRY(run_conv_in(p, scale, shift, fc0, fc2, c.use_cbam ? chmax : nullptr, tickets, H4 * W4, st));
    const int mode = i + 1 < nb ? DUCOSY_PAD_REFLECT : DUCOSY_PAD_ZERO;  // the decoder convs zero-pad
    // DUCOSY_FUSED_SPATIAL=1 evaluates the 7x7 spatial-attention conv inside the residual pass (one launch less per block).
    // Same-box A/B, round 2: the separate kernel wins at batch 30 (1003 vs 958 slices/s; the per-row gather of the pooled map
    // stalls the streaming loop, and a producer-warp variant that overlapped it was slower still, 922) and at batch 1 once the
    // forward is replayed from a CUDA graph (682 vs 636) -- so the fused form is opt-in.
    const bool fused_sa = fused_spatial();
    if (c.use_cbam && !fused_sa) {   // default: channel pooling, the 7x7 attention conv as a kernel of its own, residual pass
      float* sa = reinterpret_cast<float*>(base + w.sa);
      DUCOSY_TRY(ducosy_cbam_pool(P(w.y2b), scale, shift, pooled, B, H4, W4, 256, dt, st));
      DUCOSY_TRY(ducosy_cbam_spatial_conv(pooled, reinterpret_cast<const float*>(pk + L.saw[i]), sa, B, H4, W4, st));
      DUCOSY_TRY(ducosy_residual_apply_pad(P(w.y2b), scale, shift, sa, P(cur), 1, P(nxt), B, H4, W4, 256, 1, mode, dt, st));
    } else if (c.use_cbam) {
      // spatial attention: channel mean / max of the attended map, then the 7x7 conv + sigmoid inside the residual pass
      DUCOSY_TRY(ducosy_cbam_pool(P(w.y2b), scale, shift, pooled, B, H4, W4, 256, dt, st));
      DUCOSY_TRY(ducosy_residual_cbam_apply_pad(P(w.y2b), scale, shift, pooled, reinterpret_cast<const float*>(pk + L.saw[i]), P(cur), 1,
                                                P(nxt), B, H4, W4, 256, 1, mode, dt, st));
    } else {
      DUCOSY_TRY(ducosy_residual_apply_pad(P(w.y2b), scale, shift, nullptr, P(cur), 1, P(nxt), B, H4, W4, 256, 1, mode, dt, st));
    }
    const size_t t = cur; cur = nxt; nxt = t;
  }
  // ---- up 1: nearest x2 + 3x3 p1 256 -> 128, IN, ReLU   modules/model.py:107-111
  {
    ConvPlan p{};
    upconv_plan(p, P(cur), pk + L.up1, P(w.y1), partials, B, H4, W4, 256, 128, dt);
    DUCOSY_TRY(run_conv_in(p, scale, shift, nullptr, nullptr, nullptr, tickets, H2 * W2, st));
  }
  DUCOSY_TRY(ducosy_in_apply_pad(P(w.y1), scale, shift, P(w.p1), B, H2, W2, 128, 1, DUCOSY_PAD_ZERO, DUCOSY_ACT_RELU, dt, st));
  // ---- up 2: nearest x2 + 3x3 p1 128 -> 64, IN, ReLU
  {
    ConvPlan p{};
    if (dt == DUCOSY_F16X2) upconv_plan(p, P(w.p1), pk + L.up2, P(w.y0), partials, B, H2, W2, 128, 64, dt);
    else upconv_merged_plan(p, P(w.p1), pk + L.up2, P(w.y0), partials, B, H2, W2, 128, 64, dt);
    DUCOSY_TRY(run_conv_in(p, scale, shift, nullptr, nullptr, nullptr, tickets, H * W, st));
  }
  // ---- output: IN apply + ReLU + reflect-pad 3 folded into the loader of the 7x7 conv 64 -> 1 + tanh   modules/model.py:110-112
  return ducosy_out_conv7x7_tanh_fused(P(w.y0), scale, shift, pk + L.outw, reinterpret_cast<const float*>(pk + L.outb), out, B,
                                       H, W, dt, st);
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" int ducosy_version(void) { return DUCOSY_VERSION; }
extern "C" const char* ducosy_last_error(void) { return g_err; }
extern "C" int ducosy_check_device(void) { return check_device_cached(); }

extern "C" int ducosy_conv2d_nhwc(const void* in, const void* w, void* out, float* partials, const float* bias, int act,
                                  int B, int Hp, int Wp, int Cin, int Cout, int kh, int kw, int stride, int dtype,
                                  ducosy_stream_t stream) {
  DUCOSY_CHECK(in && w && out && B > 0, DUCOSY_ERR_ARG, "conv2d_nhwc: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16 || dtype == DUCOSY_F16X2, DUCOSY_ERR_ARG, "conv2d_nhwc: bad dtype");
  DUCOSY_CHECK(kh == kw && (kh == 1 || kh == 3 || kh == 4) && (stride == 1 || stride == 2), DUCOSY_ERR_SHAPE,
               "conv2d_nhwc: kernel %dx%d stride %d unsupported", kh, kw, stride);
  DUCOSY_CHECK(act == DUCOSY_ACT_NONE || (act == DUCOSY_ACT_LRELU02 && bias != nullptr), DUCOSY_ERR_ARG,
               "conv2d_nhwc: epilogue must be none or bias+LeakyReLU(0.2)");
  DUCOSY_TRY(check_device_cached());
  ConvPlan p{};
  p.in = in; p.B = B; p.Hp = Hp; p.Wp = Wp; p.Cin = Cin; p.stride = stride;
  p.w = w; p.Cout = Cout; p.num_phases = 1; p.num_taps = kh * kw;
  for (int r = 0; r < kh; ++r)
    for (int s = 0; s < kw; ++s) {
      p.tap_dy[0][r * kw + s] = int8_t(r);
      p.tap_dx[0][r * kw + s] = int8_t(s);
    }
  p.Ho = (Hp - kh) / stride + 1;
  p.Wo = (Wp - kw) / stride + 1;
  p.Hg = p.Ho; p.Wg = p.Wo; p.out = out; p.oy_mul = p.ox_mul = 1;
  p.partials = partials; p.bias = bias; p.epi_mode = act == DUCOSY_ACT_LRELU02 ? 1 : 0; p.dtype = dtype;
  return launch_conv_gemm(p, static_cast<cudaStream_t>(stream));
}

extern "C" int ducosy_upconv2x_nhwc(const void* in_pad, const void* w_packed4, void* out, float* partials, int B, int Hs,
                                    int Ws, int Cin, int Cout, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(in_pad && w_packed4 && out && B > 0, DUCOSY_ERR_ARG, "upconv2x_nhwc: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16 || dtype == DUCOSY_F16X2, DUCOSY_ERR_ARG, "upconv2x_nhwc: bad dtype");
  DUCOSY_TRY(check_device_cached());
  ConvPlan p{};
  upconv_plan(p, in_pad, w_packed4, out, partials, B, Hs, Ws, Cin, Cout, dtype);
  return launch_conv_gemm(p, static_cast<cudaStream_t>(stream));
}

extern "C" int ducosy_upconv2x_merged_nhwc(const void* in_pad, const void* w_merged, void* out, float* partials, int B,
                                           int Hs, int Ws, int Cin, int Cout, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(in_pad && w_merged && out && B > 0, DUCOSY_ERR_ARG, "upconv2x_merged_nhwc: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "upconv2x_merged_nhwc: bad dtype");
  DUCOSY_CHECK(Cout == 64, DUCOSY_ERR_SHAPE, "upconv2x_merged_nhwc: Cout must be 64 (N = 4*Cout = 256)");
  DUCOSY_TRY(check_device_cached());
  ConvPlan p{};
  upconv_merged_plan(p, in_pad, w_merged, out, partials, B, Hs, Ws, Cin, Cout, dtype);
  return launch_conv_gemm(p, static_cast<cudaStream_t>(stream));
}

// The three convolution entry points with the InstanceNorm finalize fused into the launch (ConvFinalize, common.cuh): scale /
// shift [B][Cout] (and chmax [B][Cout] when non-NULL) are written by the CTA that completes a sample's last tile.  `tickets`:
// int32 [B], zero before the first use, left zero by every launch; launches that may run concurrently (different streams)
// need different ticket arrays.
namespace {
void set_fin(ConvPlan& p, float* scale, float* shift, float* chmax, int* tickets, int npix) {
  p.fin.scale = scale; p.fin.shift = shift; p.fin.chmax = chmax; p.fin.fc0 = nullptr; p.fin.fc2 = nullptr;
  p.fin.counter = tickets; p.fin.npix = npix;
}
}  // namespace

extern "C" int ducosy_conv2d_nhwc_in(const void* in, const void* w, void* out, float* partials, float* scale, float* shift,
                                     float* chmax, int* tickets, int B, int Hp, int Wp, int Cin, int Cout, int kh, int kw,
                                     int stride, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(in && w && out && partials && scale && shift && tickets && B > 0, DUCOSY_ERR_ARG, "conv2d_nhwc_in: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "conv2d_nhwc_in: bad dtype");
  DUCOSY_CHECK(kh == kw && (kh == 1 || kh == 3 || kh == 4) && (stride == 1 || stride == 2), DUCOSY_ERR_SHAPE,
               "conv2d_nhwc_in: kernel %dx%d stride %d unsupported", kh, kw, stride);
  DUCOSY_TRY(check_device_cached());
  ConvPlan p{};
  p.in = in; p.B = B; p.Hp = Hp; p.Wp = Wp; p.Cin = Cin; p.stride = stride;
  p.w = w; p.Cout = Cout; p.num_phases = 1; p.num_taps = kh * kw;
  for (int r = 0; r < kh; ++r)
    for (int s = 0; s < kw; ++s) {
      p.tap_dy[0][r * kw + s] = int8_t(r);
      p.tap_dx[0][r * kw + s] = int8_t(s);
    }
  p.Ho = (Hp - kh) / stride + 1;
  p.Wo = (Wp - kw) / stride + 1;
  p.Hg = p.Ho; p.Wg = p.Wo; p.out = out; p.oy_mul = p.ox_mul = 1;
  p.partials = partials; p.dtype = dtype;
  set_fin(p, scale, shift, chmax, tickets, p.Ho * p.Wo);
  return launch_conv_gemm(p, static_cast<cudaStream_t>(stream));
}

extern "C" int ducosy_upconv2x_nhwc_in(const void* in_pad, const void* w_packed4, void* out, float* partials, float* scale,
                                       float* shift, int* tickets, int B, int Hs, int Ws, int Cin, int Cout, int dtype,
                                       ducosy_stream_t stream) {
  DUCOSY_CHECK(in_pad && w_packed4 && out && partials && scale && shift && tickets && B > 0, DUCOSY_ERR_ARG, "upconv2x_nhwc_in: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "upconv2x_nhwc_in: bad dtype");
  DUCOSY_TRY(check_device_cached());
  ConvPlan p{};
  upconv_plan(p, in_pad, w_packed4, out, partials, B, Hs, Ws, Cin, Cout, dtype);
  set_fin(p, scale, shift, nullptr, tickets, 4 * Hs * Ws);
  return launch_conv_gemm(p, static_cast<cudaStream_t>(stream));
}

extern "C" int ducosy_upconv2x_merged_nhwc_in(const void* in_pad, const void* w_merged, void* out, float* partials, float* scale,
                                              float* shift, int* tickets, int B, int Hs, int Ws, int Cin, int Cout, int dtype,
                                              ducosy_stream_t stream) {
  DUCOSY_CHECK(in_pad && w_merged && out && partials && scale && shift && tickets && B > 0, DUCOSY_ERR_ARG, "upconv2x_merged_nhwc_in: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "upconv2x_merged_nhwc_in: bad dtype");
  DUCOSY_CHECK(Cout == 64, DUCOSY_ERR_SHAPE, "upconv2x_merged_nhwc_in: Cout must be 64 (N = 4*Cout = 256)");
  DUCOSY_TRY(check_device_cached());
  ConvPlan p{};
  upconv_merged_plan(p, in_pad, w_merged, out, partials, B, Hs, Ws, Cin, Cout, dtype);
  set_fin(p, scale, shift, nullptr, tickets, 4 * Hs * Ws);
  return launch_conv_gemm(p, static_cast<cudaStream_t>(stream));
}

extern "C" int ducosy_generator_num_params(const ducosy_gen_config* cfg) {
  if (!cfg) return fail(DUCOSY_ERR_ARG, "generator_num_params: null config");
  return 6 + cfg->num_residual_blocks * (4 + (cfg->use_cbam ? 3 : 0)) + 6;
}
extern "C" size_t ducosy_generator_packed_bytes(const ducosy_gen_config* cfg) {
  return cfg ? make_layout(*cfg).total : 0;
}
extern "C" size_t ducosy_generator_workspace_bytes(const ducosy_gen_config* cfg, int B, int H, int W) {
  if (!cfg || check_gen_shape(*cfg, B, H, W) != 0) return 0;
  return make_workspace(*cfg, B, H, W).total;
}
extern "C" int ducosy_generator_num_launches(const ducosy_gen_config* cfg) {
  if (!cfg) return fail(DUCOSY_ERR_ARG, "generator_num_launches: null config");
  // stem 4 (Cin = 1), 2 x (conv + finalize + apply) down, per block 2 x (conv + finalize) + apply + residual (+ channel MLP + pool
  // + spatial-attention conv, unless that runs inside the residual pass), 2 x (up conv + finalize) + 1 apply, output conv
  return 16 + cfg->num_residual_blocks * (6 + (cfg->use_cbam ? (fused_spatial() ? 2 : 3) : 0));
}

extern "C" int ducosy_generator_pack(const ducosy_gen_config* cfg, const float* const* params, int num_params,
                                     void* packed, ducosy_stream_t stream) {
  DUCOSY_CHECK(cfg && params && packed, DUCOSY_ERR_ARG, "generator_pack: null pointer");
  const ducosy_gen_config& c = *cfg;
  DUCOSY_CHECK(num_params == ducosy_generator_num_params(cfg), DUCOSY_ERR_ARG,
               "generator_pack: expected %d parameter tensors, got %d", ducosy_generator_num_params(cfg), num_params);
  DUCOSY_CHECK(c.dtype == DUCOSY_F16 || c.dtype == DUCOSY_BF16 || c.dtype == DUCOSY_F16X2, DUCOSY_ERR_ARG, "generator_pack: bad dtype");
  DUCOSY_CHECK((reinterpret_cast<uintptr_t>(packed) & 255) == 0, DUCOSY_ERR_ALIGN, "generator_pack: packed buffer must be 256-byte aligned");
  for (int i = 0; i < num_params; ++i) DUCOSY_CHECK(params[i] != nullptr, DUCOSY_ERR_ARG, "generator_pack: parameter %d is null", i);
  const GenLayout L = make_layout(c);
  uint8_t* pk = static_cast<uint8_t*>(packed);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int k = 0;
  // conv biases that feed a non-affine InstanceNorm cancel exactly (modules/model.py:94-98,60-62): not packed.
  DUCOSY_TRY(ducosy_pack_stem_weight(params[k], pk + L.stem, c.input_channels, c.dtype, stream)); k += 2;
  DUCOSY_TRY(ducosy_pack_conv_weight(params[k], pk + L.d1, 128, 64, 3, 3, c.dtype, stream)); k += 2;
  DUCOSY_TRY(ducosy_pack_conv_weight(params[k], pk + L.d2, 256, 128, 3, 3, c.dtype, stream)); k += 2;
  for (int i = 0; i < c.num_residual_blocks; ++i) {
    DUCOSY_TRY(ducosy_pack_conv_weight(params[k], pk + L.c1[i], 256, 256, 3, 3, c.dtype, stream)); k += 2;
    DUCOSY_TRY(ducosy_pack_conv_weight(params[k], pk + L.c2[i], 256, 256, 3, 3, c.dtype, stream)); k += 2;
    if (c.use_cbam) {
      cudaMemcpyAsync(pk + L.fc0[i], params[k++], 16 * 256 * 4, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(pk + L.fc2[i], params[k++], 256 * 16 * 4, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(pk + L.saw[i], params[k++], 98 * 4, cudaMemcpyDeviceToDevice, st);
    }
  }
  DUCOSY_TRY(ducosy_pack_upconv_weight(params[k], pk + L.up1, 128, 256, c.dtype, stream)); k += 2;
  if (c.dtype == DUCOSY_F16X2) { DUCOSY_TRY(ducosy_pack_upconv_weight(params[k], pk + L.up2, 64, 128, c.dtype, stream)); }
  else { DUCOSY_TRY(ducosy_pack_upconv_merged_weight(params[k], pk + L.up2, 64, 128, c.dtype, stream)); }
  k += 2;
  DUCOSY_TRY(ducosy_pack_out_weight(params[k++], pk + L.outw, c.dtype, stream));
  cudaMemcpyAsync(pk + L.outb, params[k++], 4, cudaMemcpyDeviceToDevice, st);
  return check_launch("generator_pack");
}

extern "C" int ducosy_generator_forward(const ducosy_gen_config* cfg, const void* packed, const float* x, float* out,
                                        int B, int H, int W, void* workspace, size_t workspace_bytes,
                                        ducosy_stream_t stream) {
  DUCOSY_CHECK(cfg && x, DUCOSY_ERR_ARG, "generator_forward: null pointer");
  return generator_forward_impl(*cfg, packed, x, nullptr, 0.f, 0.f, 0.f, 0.f, out, B, H, W, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int ducosy_generator_forward_hu(const ducosy_gen_config* cfg, const void* packed, const int16_t* px,
                                           float slope, float intercept, float hu_lo, float hu_hi, float* out, int B,
                                           int H, int W, void* workspace, size_t workspace_bytes,
                                           ducosy_stream_t stream) {
  DUCOSY_CHECK(cfg && px, DUCOSY_ERR_ARG, "generator_forward_hu: null pointer");
  return generator_forward_impl(*cfg, packed, nullptr, px, slope, intercept, hu_lo, hu_hi, out, B, H, W, workspace,
                                workspace_bytes, static_cast<cudaStream_t>(stream));
}
