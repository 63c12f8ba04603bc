// Weight gradient of the NHWC convolutions on tcgen05 (building block of the training-step rows):
//   dW[o][tap][c] = sum_{pixels p} dY[p][o] * Xpad[p + offset(tap)][c]
// Per tap this is a GEMM with M = Cout, N = Cin and K = pixels.  Both operands are "MN-major" for the tensor core
// (the M / N index is the contiguous NHWC channel, K = pixel strides by the channel count): the TMA boxes
// (64 channels x 64 pixels, 128-byte swizzle) land exactly in the canonical MN-major UMMA layout -- 128-byte rows of
// 64 channels per pixel, 8-pixel swizzle atoms 1 KiB apart (SBO), 64-channel slabs 8 KiB apart (LBO).
// Work unit = (tap, 128-row block of Cout, K split); each CTA accumulates one unit in TMEM (fp32) and writes its
// partial tile; a second kernel reduces the K splits in fixed order (deterministic).
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace ducosy {
namespace {

constexpr int kWgThreads = 256;
constexpr int kKChunk = 64;                 // pixels per pipeline stage
constexpr int kSlab = kKChunk * 128;        // bytes: 64 pixel rows x 64 channels x 16 bit
constexpr int kWgStages = 4;

struct WgradArgs {
  int num_taps, B, TY, TX, R, Wt, log2Wt;   // pixel chunks: R rows x Wt cols of the dY grid (R*Wt == 64)
  int rows_per_sample;                      // outermost TMA dim units of the input map per sample
  int dy_pad, dy_rows_per_sample;           // dY may live inside a zero-padded buffer (border width dy_pad)
  int Cin, Cout, m_blocks, n_slabs;         // n_slabs = Cin / 64 (<= 4)
  int m_per_cta, stages;                    // 128-row blocks of Cout one CTA accumulates (they share the B tiles), pipeline depth
  int splits, chunks_per_split;
  int8_t tap_xp[kMaxTaps], tap_dx[kMaxTaps], tap_yp[kMaxTaps], tap_dy[kMaxTaps];
  int8_t dyq_xp[kMaxTaps], dyq_yp[kMaxTaps];   // parity of the dY pixels a unit reads (stride-2 dY map: sub-pixel up-conv wgrad)
  float* partial;                           // [splits][taps][Cout][Cin] fp32
};

// MN-major, 128-byte swizzle operand descriptor: LBO = distance between 64-element slabs, SBO = 8-row atom pitch.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <typename T, int kMP>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradArgs a) {
  constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  extern __shared__ uint8_t wg_raw[];
  const uint32_t raw_addr = smem_u32(wg_raw);
  uint8_t* smem = wg_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int a_slabs = 2 * kMP;                           // A: 2 slabs per 128 output channels, B: n_slabs
  const int stage_bytes = (a_slabs + a.n_slabs) * kSlab;
  constexpr int stages = kMP == 2 ? 3 : kWgStages;
  constexpr uint32_t tmem_cols = kMP == 2 ? 512u : 256u;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWgStages * 6 * kSlab);
  uint64_t* empty = full + kWgStages;
  uint64_t* done = empty + kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // unit decode: blockIdx.x = ((split * taps + tap) * m_units + mu), m_units = m_blocks / m_per_cta
  const int m_units = a.m_blocks / kMP;
  const int mu = blockIdx.x % m_units;
  const int tap = (blockIdx.x / m_units) % a.num_taps;
  const int split = blockIdx.x / (m_units * a.num_taps);
  const int chunks_total = a.B * a.TY * a.TX;
  const int c_begin = split * a.chunks_per_split;
  const int c_end = min(c_begin + a.chunks_per_split, chunks_total);
  const int N = a.n_slabs * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh): only after the TMEM allocation

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();            // the operand maps come from the previous kernels
      int s = 0;
      uint32_t ph = 0;
      for (int ch = c_begin; ch < c_end; ++ch) {
        const int tx = ch % a.TX;
        int r = ch / a.TX;
        const int ty = r % a.TY;
        const int b = r / a.TY;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + size_t(s) * stage_bytes;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        // A: dY chunk, 64-channel slabs of this CTA's 128-row block(s) of Cout
        for (int h = 0; h < a_slabs; ++h)
          tma_load_5d(st + h * kSlab, &tmDY, &full[s], mu * kMP * 128 + h * 64, a.dyq_xp[tap], tx * a.Wt + a.dy_pad,
                      a.dyq_yp[tap], b * a.dy_rows_per_sample + ty * a.R + a.dy_pad);
        // B: the input pixels this tap multiplies, all Cin channels
        for (int sl = 0; sl < a.n_slabs; ++sl)
          tma_load_5d(st + (a_slabs + sl) * kSlab, &tmX, &full[s], sl * 64, a.tap_xp[tap], tx * a.Wt + a.tap_dx[tap],
                      a.tap_yp[tap], b * a.rows_per_sample + ty * a.R + a.tap_dy[tap]);
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: fp32 accumulate, A and B both MN-major (bits 15 / 16), M = 128, N = Cin
    const uint32_t idesc = umma_idesc_f16(kFmt, 128, N) | (1u << 15) | (1u << 16);
    int s = 0;
    uint32_t ph = 0;
    for (int ch = c_begin; ch < c_end; ++ch) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + size_t(s) * stage_bytes);
#pragma unroll
        for (int j = 0; j < kKChunk / 16; ++j) {  // 16 pixels = two 8-row swizzle atoms = 2 KiB per K step
          const uint64_t db = umma_desc_mn_sw128(sa + a_slabs * kSlab + j * 2048, kSlab);
#pragma unroll
          for (int mh = 0; mh < kMP; ++mh) {
            const uint64_t da = umma_desc_mn_sw128(sa + mh * 2 * kSlab + j * 2048, kSlab);
            umma_f16(tmem_base + uint32_t(mh * 256), da, db, idesc, (ch > c_begin || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
        if (ch == c_end - 1) umma_commit(done);
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    pdl_wait();              // before the first global store
    mbar_wait(done, 0);
    tc_fence_after();
    for (int mh = 0; mh < kMP; ++mh) {
      const int o = (mu * kMP + mh) * 128 + q * 32 + lane;
      float* dst = a.partial + ((size_t(split) * a.num_taps + tap) * a.Cout + o) * a.Cin;
      for (int chn = 0; chn < N / 32; ++chn) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(mh * 256 + chn * 32), v);
        tmem_ld_wait();
        if (o < a.Cout) {   // Cout = 64: the upper 64 accumulator rows come from TMA zero fill and are not stored
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<uint4*>(dst + chn * 32)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dw = sum over splits, fixed order.  oihw == 0: packed forward layout dw[o][tap*Cin + c]; oihw != 0: the parameter's own layout
// dw[o][c][tap] times gs[1] (the inverse gradient scale) -- reduce and unpack in one pass
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int taps,
                                    int Cout, int Cin, int oihw, const float* __restrict__ gs) {
  pdl_prologue();
  const long long per_split = (long long)taps * Cout * Cin;
  const float inv = (oihw != 0 && gs != nullptr) ? gs[1] : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_split; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % Cin);
    long long r = i / Cin;
    const int o = int(r % Cout);
    const int tap = int(r / Cout);
    // four independent chains (loads of four splits in flight), combined in a fixed order: deterministic
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
      a0 += partial[(s + 0) * per_split + i];
      a1 += partial[(s + 1) * per_split + i];
      a2 += partial[(s + 2) * per_split + i];
      a3 += partial[(s + 3) * per_split + i];
    }
    for (; s < splits; ++s) a0 += partial[s * per_split + i];
    const float acc = (a0 + a1) + (a2 + a3);
    if (oihw == 2) dw[((long long)o * Cin + c) * taps + tap] += acc * inv;      // accumulate into an existing gradient (p.grad)
    else if (oihw != 0) dw[((long long)o * Cin + c) * taps + tap] = acc * inv;
    else dw[((long long)o * taps + tap) * Cin + c] = acc;
  }
}

int encode_nhwc_map(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, int B, int Hp, int Wp, int C, int stride,
                    int Wt, int R) {
  const cuuint64_t C2 = cuuint64_t(C) * 2, W = cuuint64_t(Wp), H = cuuint64_t(Hp);
  cuuint64_t gdim[5], gstr[4];
  if (stride == 1) {
    gdim[0] = C; gdim[1] = 1; gdim[2] = W; gdim[3] = 1; gdim[4] = cuuint64_t(B) * H;
    gstr[0] = C2; gstr[1] = C2; gstr[2] = W * C2; gstr[3] = W * C2;
  } else {
    gdim[0] = C; gdim[1] = 2; gdim[2] = W / 2; gdim[3] = 2; gdim[4] = cuuint64_t(B) * H / 2;
    gstr[0] = C2; gstr[1] = 2 * C2; gstr[2] = W * C2; gstr[3] = 2 * W * C2;
  }
  const cuuint32_t box[5] = {64, 1, cuuint32_t(Wt), 1, cuuint32_t(R)};
  return encode_tiled_cached(tm, dt, 5, base, gdim, gstr, box, "conv_wgrad");
}

// K splits: one CTA per SM and unit.  The grid must not spill into an extra wave (297 CTAs on 148 SMs run as long as
// 444), so the count is rounded DOWN to whole waves: two waves of light CTAs, one wave of the double-M ones.
// Two 128-row blocks of Cout per CTA share the B tiles (a third less L2 -> SMEM traffic), but double the fp32 partial tile
// every CTA writes and halve the number of units; with few pixel chunks (one or two 128x128 samples) the kernel is bound by
// that epilogue and by the split-K reduction behind it, so small problems take one block per CTA and half as many splits.
int wgrad_m_per_cta(int m_blocks, int chunks) { return (m_blocks % 2 == 0 && chunks >= 512) ? 2 : 1; }

int wgrad_splits(int units, int m_per_cta, int chunks) {
  const int target = (m_per_cta == 2 ? 1 : 2) * 148;
  int splits = target / units;
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  return splits;
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" size_t ducosy_conv2d_wgrad_workspace_bytes(int B, int Ho, int Wo, int Cin, int Cout, int kh, int kw) {
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  const int chunks = B * Ho * Wo / kKChunk;
  const int m_blocks = (Cout + 127) / 128, m_per_cta = wgrad_m_per_cta(m_blocks, chunks);
  const int splits = wgrad_splits(kh * kw * (m_blocks / m_per_cta), m_per_cta, chunks);
  return size_t(splits) * kh * kw * Cout * Cin * 4;
}

// x_pad: the padded NHWC input the forward conv read ([B][Hp][Wp][Cin]); dy: NHWC output gradient
// [B][Ho+2*dy_pad][Wo+2*dy_pad][Cout] (interior used), 16-bit both; dw: fp32 [Cout][kh*kw*Cin] in the packed forward layout (k = (r*kw+s)*Cin + c).
namespace {
int conv2d_wgrad_impl(const void* x_pad, const void* dy, int dy_pad, float* dw, int oihw, const float* gs, int B, int Hp, int Wp,
                      int Cin, int Cout, int kh, int kw, int stride, void* workspace, size_t workspace_bytes, int dtype,
                      ducosy_stream_t stream) {
  DUCOSY_CHECK(x_pad && dy && dw && workspace && B > 0, DUCOSY_ERR_ARG, "conv2d_wgrad: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "conv2d_wgrad: bad dtype");
  DUCOSY_CHECK(kh == kw && (kh == 1 || kh == 3 || kh == 4) && (stride == 1 || stride == 2), DUCOSY_ERR_SHAPE,
               "conv2d_wgrad: kernel %dx%d stride %d unsupported", kh, kw, stride);
  DUCOSY_CHECK(Cin % 64 == 0 && Cin <= 256 && (Cout % 128 == 0 || Cout == 64), DUCOSY_ERR_SHAPE,
               "conv2d_wgrad: needs Cin in {64,128,192,256} and Cout 64 or a multiple of 128 (got %d -> %d)", Cin, Cout);
  DUCOSY_CHECK(stride == 1 || (Hp % 2 == 0 && Wp % 2 == 0), DUCOSY_ERR_SHAPE, "conv2d_wgrad: stride 2 needs even padded extents");
  DUCOSY_TRY(ducosy_check_device());
  const int Ho = (Hp - kh) / stride + 1, Wo = (Wp - kw) / stride + 1;
  const int Wt = Wo < kKChunk ? Wo : kKChunk;
  DUCOSY_CHECK(Wt >= 8 && (Wt & (Wt - 1)) == 0 && Wo % Wt == 0 && Ho % (kKChunk / Wt) == 0, DUCOSY_ERR_SHAPE,
               "conv2d_wgrad: output grid %dx%d unsupported", Ho, Wo);
  const int R = kKChunk / Wt;
  WgradArgs a{};
  a.num_taps = kh * kw;
  a.B = B; a.R = R; a.Wt = Wt;
  a.TY = Ho / R; a.TX = Wo / Wt;
  a.rows_per_sample = stride == 1 ? Hp : Hp / 2;
  a.dy_pad = dy_pad;
  a.dy_rows_per_sample = Ho + 2 * dy_pad;
  a.Cin = Cin; a.Cout = Cout; a.m_blocks = (Cout + 127) / 128; a.n_slabs = Cin / 64;
  for (int r = 0; r < kh; ++r)
    for (int s = 0; s < kw; ++s) {
      const int t = r * kw + s;
      if (stride == 1) { a.tap_xp[t] = 0; a.tap_dx[t] = int8_t(s); a.tap_yp[t] = 0; a.tap_dy[t] = int8_t(r); }
      else { a.tap_xp[t] = int8_t(s & 1); a.tap_dx[t] = int8_t(s >> 1); a.tap_yp[t] = int8_t(r & 1); a.tap_dy[t] = int8_t(r >> 1); }
    }
  const int chunks = B * a.TY * a.TX;
  a.m_per_cta = wgrad_m_per_cta(a.m_blocks, chunks);
  a.stages = a.m_per_cta == 2 ? 3 : kWgStages;
  const int units = a.num_taps * (a.m_blocks / a.m_per_cta);
  int splits = wgrad_splits(units, a.m_per_cta, chunks);
  a.chunks_per_split = (chunks + splits - 1) / splits;
  splits = (chunks + a.chunks_per_split - 1) / a.chunks_per_split;   // no empty split
  a.splits = splits;
  const size_t need = size_t(splits) * a.num_taps * Cout * Cin * 4;
  DUCOSY_CHECK(workspace_bytes >= need, DUCOSY_ERR_WORKSPACE, "conv2d_wgrad: workspace %zu < required %zu bytes", workspace_bytes, need);
  a.partial = static_cast<float*>(workspace);

  const CUtensorMapDataType dt = dtype == DUCOSY_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmDY, tmX;
  DUCOSY_CHECK(dy_pad >= 0 && dy_pad <= 4, DUCOSY_ERR_ARG, "conv2d_wgrad: dy_pad out of range");
  DUCOSY_TRY(encode_nhwc_map(&tmDY, dt, dy, B, Ho + 2 * dy_pad, Wo + 2 * dy_pad, Cout, 1, Wt, R));
  DUCOSY_TRY(encode_nhwc_map(&tmX, dt, x_pad, B, Hp, Wp, Cin, stride, Wt, R));

  const size_t smem = 1024 + size_t(kWgStages) * 6 * kSlab + 256;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = splits * units;
#define DUCOSY_WGRAD_LAUNCH(T, MP)                                                                              \
  do {                                                                                                          \
    static PerDeviceOnce cfg;                                                                                   \
    cfg.once([&] { return cudaFuncSetAttribute(conv_wgrad_kernel<T, MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); }); \
    pdl(conv_wgrad_kernel<T, MP>, grid, kWgThreads, smem, st)(tmDY, tmX, a);                                      \
  } while (0)
  if (dtype == DUCOSY_F16) {
    if (a.m_per_cta == 2) DUCOSY_WGRAD_LAUNCH(__half, 2); else DUCOSY_WGRAD_LAUNCH(__half, 1);
  } else {
    if (a.m_per_cta == 2) DUCOSY_WGRAD_LAUNCH(__nv_bfloat16, 2); else DUCOSY_WGRAD_LAUNCH(__nv_bfloat16, 1);
  }
#undef DUCOSY_WGRAD_LAUNCH
  DUCOSY_TRY(check_launch("conv_wgrad_kernel"));
  const long long per_split = (long long)a.num_taps * Cout * Cin;
  long long blocks = (per_split + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  pdl(wgrad_reduce_kernel, int(blocks), 256, 0, st)(a.partial, dw, splits, a.num_taps, Cout, Cin, oihw, gs);
  return check_launch("wgrad_reduce_kernel");
}
}  // namespace

extern "C" int ducosy_conv2d_wgrad_nhwc(const void* x_pad, const void* dy, int dy_pad, float* dw, int B, int Hp, int Wp,
                                        int Cin, int Cout, int kh, int kw, int stride, void* workspace,
                                        size_t workspace_bytes, int dtype, ducosy_stream_t stream) {
  return conv2d_wgrad_impl(x_pad, dy, dy_pad, dw, 0, nullptr, B, Hp, Wp, Cin, Cout, kh, kw, stride, workspace, workspace_bytes, dtype, stream);
}

// The same, reduced straight into the parameter's own layout: dw_oihw fp32 [Cout][Cin][kh][kw] times gs[1] (gs = the pair of
// ducosy_grad_scale, may be NULL) -- ducosy_conv2d_wgrad_nhwc + ducosy_unpack_wgrad in one reduction pass.
extern "C" int ducosy_conv2d_wgrad_nhwc_oihw(const void* x_pad, const void* dy, int dy_pad, float* dw_oihw, const float* gs, int B,
                                             int Hp, int Wp, int Cin, int Cout, int kh, int kw, int stride, void* workspace,
                                             size_t workspace_bytes, int dtype, ducosy_stream_t stream) {
  return conv2d_wgrad_impl(x_pad, dy, dy_pad, dw_oihw, 1, gs, B, Hp, Wp, Cin, Cout, kh, kw, stride, workspace, workspace_bytes, dtype, stream);
}

// ... and ACCUMULATED into dw_oihw (dw += gradient * gs[1]): the destination is the parameter's existing .grad (a view of the
// data-parallel gradient bucket), which saves autograd's own add kernel and the temporary per weight and generator pass.
extern "C" int ducosy_conv2d_wgrad_nhwc_oihw_acc(const void* x_pad, const void* dy, int dy_pad, float* dw_oihw, const float* gs, int B,
                                                 int Hp, int Wp, int Cin, int Cout, int kh, int kw, int stride, void* workspace,
                                                 size_t workspace_bytes, int dtype, ducosy_stream_t stream) {
  return conv2d_wgrad_impl(x_pad, dy, dy_pad, dw_oihw, 2, gs, B, Hp, Wp, Cin, Cout, kh, kw, stride, workspace, workspace_bytes, dtype, stream);
}

// Weight gradient of Upsample(x2 nearest) + Conv3x3(pad 1) (modules/model.py:108-109) without materialising the
// up-sampled map: the forward runs as four phase-specific 2x2 convs on the source grid (pack_upconv_weight_kernel), so
// the gradient of the 16 pre-summed (phase, tap) blocks is a wgrad over the SOURCE pixels -- 16 units x Hs*Ws pixels
// instead of 9 x 4*Hs*Ws (4/9 of the MACs, less than half of the operand traffic) -- with dY read at stride 2 (phase =
// pixel parity).  dwph: fp32 [Cout][16*Cin], unit index (phase*4 + a*2 + b); ducosy_unpack_upconv_wgrad folds it back
// to the 3x3 weight (adjoint of the pre-summing).  src_pad: [B][Hs+2][Ws+2][Cin] (zero border 1); dy:
// [B][2Hs+2*dy_pad][2Ws+2*dy_pad][Cout], dy_pad even.
extern "C" size_t ducosy_upconv2x_wgrad_workspace_bytes(int B, int Hs, int Ws, int Cin, int Cout) {
  if (B <= 0 || Hs <= 0 || Ws <= 0) return 0;
  const int chunks = B * Hs * Ws / kKChunk;
  const int m_blocks = (Cout + 127) / 128, m_per_cta = (m_blocks % 2 == 0) ? 2 : 1;
  const int splits = wgrad_splits(16 * (m_blocks / m_per_cta), m_per_cta, chunks);
  return size_t(splits) * 16 * Cout * Cin * 4;
}

extern "C" int ducosy_upconv2x_wgrad_nhwc(const void* src_pad, const void* dy, int dy_pad, float* dwph, int B, int Hs, int Ws,
                                          int Cin, int Cout, void* workspace, size_t workspace_bytes, int dtype,
                                          ducosy_stream_t stream) {
  DUCOSY_CHECK(src_pad && dy && dwph && workspace && B > 0, DUCOSY_ERR_ARG, "upconv2x_wgrad: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "upconv2x_wgrad: bad dtype");
  DUCOSY_CHECK(Cin % 64 == 0 && Cin <= 256 && (Cout % 128 == 0 || Cout == 64), DUCOSY_ERR_SHAPE,
               "upconv2x_wgrad: needs Cin in {64,128,192,256} and Cout 64 or a multiple of 128 (got %d -> %d)", Cin, Cout);
  DUCOSY_CHECK(dy_pad >= 0 && dy_pad <= 4 && dy_pad % 2 == 0, DUCOSY_ERR_ARG, "upconv2x_wgrad: dy_pad must be 0, 2 or 4");
  DUCOSY_TRY(ducosy_check_device());
  const int Wt = Ws < kKChunk ? Ws : kKChunk;
  DUCOSY_CHECK(Wt >= 8 && (Wt & (Wt - 1)) == 0 && Ws % Wt == 0 && Hs % (kKChunk / Wt) == 0, DUCOSY_ERR_SHAPE,
               "upconv2x_wgrad: source grid %dx%d unsupported", Hs, Ws);
  const int R = kKChunk / Wt;
  WgradArgs a{};
  a.num_taps = 16;
  a.B = B; a.R = R; a.Wt = Wt;
  a.TY = Hs / R; a.TX = Ws / Wt;
  a.rows_per_sample = Hs + 2;
  a.dy_pad = dy_pad / 2;                                   // in units of the stride-2 map
  a.dy_rows_per_sample = (2 * Hs + 2 * dy_pad) / 2;
  a.Cin = Cin; a.Cout = Cout; a.m_blocks = (Cout + 127) / 128; a.n_slabs = Cin / 64;
  for (int u = 0; u < 16; ++u) {
    const int phase = u >> 2, ta = (u >> 1) & 1, tb = u & 1, py = phase >> 1, px = phase & 1;
    a.tap_xp[u] = 0; a.tap_yp[u] = 0;
    a.tap_dy[u] = int8_t(ta + py);                         // source row i + a + py of the padded source (api.cu: upconv2x_nhwc)
    a.tap_dx[u] = int8_t(tb + px);
    a.dyq_yp[u] = int8_t(py);                              // dY pixel (2i + py, 2j + px); dy_pad is even, so parity is unchanged
    a.dyq_xp[u] = int8_t(px);
  }
  const int chunks = B * a.TY * a.TX;
  a.m_per_cta = (a.m_blocks % 2 == 0) ? 2 : 1;
  a.stages = a.m_per_cta == 2 ? 3 : kWgStages;
  const int units = a.num_taps * (a.m_blocks / a.m_per_cta);
  int splits = wgrad_splits(units, a.m_per_cta, chunks);
  a.chunks_per_split = (chunks + splits - 1) / splits;
  splits = (chunks + a.chunks_per_split - 1) / a.chunks_per_split;
  a.splits = splits;
  const size_t need = size_t(splits) * a.num_taps * Cout * Cin * 4;
  DUCOSY_CHECK(workspace_bytes >= need, DUCOSY_ERR_WORKSPACE, "upconv2x_wgrad: workspace %zu < required %zu bytes", workspace_bytes, need);
  a.partial = static_cast<float*>(workspace);
  const CUtensorMapDataType dt = dtype == DUCOSY_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmDY, tmX;
  DUCOSY_TRY(encode_nhwc_map(&tmDY, dt, dy, B, 2 * Hs + 2 * dy_pad, 2 * Ws + 2 * dy_pad, Cout, 2, Wt, R));
  DUCOSY_TRY(encode_nhwc_map(&tmX, dt, src_pad, B, Hs + 2, Ws + 2, Cin, 1, Wt, R));
  const size_t smem = 1024 + size_t(kWgStages) * 6 * kSlab + 256;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = splits * units;
#define DUCOSY_WGRAD_LAUNCH(T, MP)                                                                              \
  do {                                                                                                          \
    cudaFuncSetAttribute(conv_wgrad_kernel<T, MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));      \
    pdl(conv_wgrad_kernel<T, MP>, grid, kWgThreads, smem, st)(tmDY, tmX, a);                                      \
  } while (0)
  if (dtype == DUCOSY_F16) {
    if (a.m_per_cta == 2) DUCOSY_WGRAD_LAUNCH(__half, 2); else DUCOSY_WGRAD_LAUNCH(__half, 1);
  } else {
    if (a.m_per_cta == 2) DUCOSY_WGRAD_LAUNCH(__nv_bfloat16, 2); else DUCOSY_WGRAD_LAUNCH(__nv_bfloat16, 1);
  }
#undef DUCOSY_WGRAD_LAUNCH
  DUCOSY_TRY(check_launch("conv_wgrad_kernel"));
  const long long per_split = (long long)a.num_taps * Cout * Cin;
  long long blocks = (per_split + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  pdl(wgrad_reduce_kernel, int(blocks), 256, 0, st)(a.partial, dwph, splits, a.num_taps, Cout, Cin, 0, nullptr);
  return check_launch("wgrad_reduce_kernel");
}

namespace ducosy {
namespace {
// dW[o][c][r][s] = gs1 * sum_{py,px} dWph[o][((py*2+px)*4 + a(py,r)*2 + b(px,s)) * Cin + c],  a(0,r) = (r >= 1), a(1,r) = (r >= 2)
__global__ void unpack_upconv_wgrad_kernel(const float* __restrict__ dwph, float* __restrict__ g, int Cout, int Cin,
                                           const float* __restrict__ gs) {
  pdl_prologue();
  const float inv = gs != nullptr ? gs[1] : 1.f;
  const long long total = (long long)Cout * Cin * 9;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = int(i % 9), r = tap / 3, s = tap % 3;
    const long long oc = i / 9;
    const int c = int(oc % Cin), o = int(oc / Cin);
    float acc = 0.f;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const int ta = py == 0 ? (r >= 1) : (r >= 2), tb = px == 0 ? (s >= 1) : (s >= 2);
        acc += dwph[((long long)o * 16 + (py * 2 + px) * 4 + ta * 2 + tb) * Cin + c];
      }
    g[i] = acc * inv;
  }
}
}  // namespace
}  // namespace ducosy

extern "C" int ducosy_unpack_upconv_wgrad(const float* dwph, float* g_oihw, int Cout, int Cin, const float* gs, ducosy_stream_t stream) {
  DUCOSY_CHECK(dwph && g_oihw && Cout > 0 && Cin > 0, DUCOSY_ERR_ARG, "unpack_upconv_wgrad: bad argument");
  const long long total = (long long)Cout * Cin * 9;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  pdl(unpack_upconv_wgrad_kernel, int(blocks), 256, 0, (cudaStream_t)stream)(dwph, g_oihw, Cout, Cin, gs);
  return check_launch("unpack_upconv_wgrad_kernel");
}
