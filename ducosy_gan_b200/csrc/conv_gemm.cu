// Implicit-GEMM convolution on 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), persistent over output tiles.
//
// Replaces, for NHWC 16-bit activations, the cuDNN convolutions behind
//   modules/model.py:94   (7x7 stem, via an im2col matrix = "1 tap")
//   modules/model.py:96-98 (3x3 stride-2 down convs)
//   modules/model.py:60-62,73-79 (3x3 residual-block convs on reflect-padded input)
//   modules/model.py:108-111 (nearest x2 upsample + 3x3 conv, as four phase-specific 2x2 convs)
//   modules/model.py:122-128 (4x4 stride-2 PatchGAN convs)
//
// GEMM view:  D[m, n] = sum_{tap, c} A[pixel(m) + offset(tap), c] * W[n, tap*Cin + c]
//   m: 128 pixels of the "GEMM grid" (R rows x Wt cols, R*Wt = 128), n: kN output channels.
//   A tile for (tap, 64-channel chunk) is ONE TMA box of the padded NHWC input, described by a 5-D
//   tensor map (c, x-parity, x, y-parity, y*batch) so that stride-2 convs are plain boxes too.
//
// Warp roles (384 threads, 1 CTA/SM): warp 0 = TMA producer, warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warps 4-11 = epilogue (TMEM -> registers -> 16-bit NHWC store, plus the
// InstanceNorm / CBAM-pool statistics: per-tile per-channel sum, sum of squares and max).
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace ducosy {

namespace {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // 64 x 16-bit = one 128-byte swizzle row
constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KiB
constexpr int kThreads = 384;                   // 4 control warps + 8 epilogue warps
constexpr int kEpiThreads = 256;

// kCG = 1: one CTA per 128 x kN tile.  kCG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256 x kN
// tile; each CTA stages its own 128 A rows and HALF of the B rows, so L2->SMEM traffic per FLOP drops by a third
// and the same shared memory holds 6 instead of 4 pipeline stages.
template <int kN, int kCG>
struct Cfg {
  static constexpr int kBRows = kN / kCG;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * kN;        // two accumulator stages (power of two >= 32)
  static constexpr int kStatFloats = 4 * 3 * kN;  // per epilogue warp: sum / sumsq / max
  static constexpr size_t kSmemBytes = 1024 + size_t(kStages) * kStageBytes + kStatFloats * 4 + 256;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Transposing butterfly: on entry lane L holds x[0..31] (32 columns of its row); on exit x[0] of
// lane L is the reduction of column L over the 32 lanes.  31 shuffles instead of 160.
template <bool kMax, int H>
__device__ __forceinline__ void butterfly_step(float (&x)[32], int lane) {
  const bool up = (lane & H) != 0;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float send = up ? x[i] : x[i + H];
    const float keep = up ? x[i + H] : x[i];
    const float recv = __shfl_xor_sync(0xffffffffu, send, H);
    x[i] = kMax ? fmaxf(keep, recv) : keep + recv;
  }
}
template <bool kMax>
__device__ __forceinline__ float butterfly_reduce(float (&x)[32], int lane) {
  butterfly_step<kMax, 16>(x, lane);
  butterfly_step<kMax, 8>(x, lane);
  butterfly_step<kMax, 4>(x, lane);
  butterfly_step<kMax, 2>(x, lane);
  butterfly_step<kMax, 1>(x, lane);
  return x[0];
}

struct TileCoord {
  int b, phase, ty, tx, nb, tile_m;  // tile_m: index of the m-tile inside its sample
};
__device__ __forceinline__ TileCoord decode_tile(int tile, const ConvGemmArgs& a) {
  TileCoord t;
  t.nb = tile % a.n_blocks;
  int r = tile / a.n_blocks;
  t.tx = r % a.TX;
  r /= a.TX;
  t.ty = r % a.TY;
  r /= a.TY;
  t.phase = r % a.num_phases;
  t.b = r / a.num_phases;
  t.tile_m = (t.phase * a.TY + t.ty) * a.TX + t.tx;
  return t;
}

template <int kN, typename T, int kCG>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvGemmArgs a) {
  using C = Cfg<kN, kCG>;
  constexpr int kStages = C::kStages;
  constexpr int kFmt = sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value ? 1 : 0;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B wants 1024-B alignment
  float* stat = reinterpret_cast<float*>(smem + size_t(kStages) * C::kStageBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(stat + C::kStatFloats);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kiters = a.num_taps * a.kc_per_tap;
  const int tiles_m_per_sample = a.num_phases * a.TY * a.TX;
  // work unit = (group of kCG consecutive m-tiles) x (n-block); this CTA owns m-tile  unit_m * kCG + rank
  const int total_units = (a.B * tiles_m_per_sample / kCG) * a.n_blocks;
  const uint32_t rank = kCG == 2 ? cluster_ctarank() : 0u;
  const int unit0 = blockIdx.x / kCG, unit_step = gridDim.x / kCG;
  auto unit_tile = [&](int unit) { return ((unit / a.n_blocks) * kCG + int(rank)) * a.n_blocks + unit % a.n_blocks; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kEpiThreads * kCG);  // in a pair both CTAs' epilogues release the leader's MMA warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kCG == 2) {
      tmem_alloc_cg2(tmem_slot, C::kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (kCG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int unit = unit0; unit < total_units; unit += unit_step) {
        const TileCoord tc = decode_tile(unit_tile(unit), a);
        const int x0 = tc.tx * a.Wt;
        const int row0 = tc.b * a.rows_per_sample + tc.ty * a.R;
        const int n0 = tc.phase * a.Cout + tc.nb * kN + int(rank) * C::kBRows;
        for (int t = 0; t < a.num_taps; ++t) {
          const int xp = a.tap_xp[tc.phase][t], dx = a.tap_dx[tc.phase][t];
          const int yp = a.tap_yp[tc.phase][t], dy = a.tap_dy[tc.phase][t];
          for (int kc = 0; kc < a.kc_per_tap; ++kc) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sa = smem + size_t(s) * C::kStageBytes;
            if (kCG == 2) {
              // both CTAs' loads complete on the leader's barrier, which expects the bytes of the whole pair
              if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * C::kStageBytes);
              tma_load_5d_cg2(sa, &tmA, &full[s], kc * kBlockK, xp, x0 + dx, yp, row0 + dy);
              tma_load_2d_cg2(sa + kABytes, &tmB, &full[s], (t * a.kc_per_tap + kc) * kBlockK, n0);
            } else {
              mbar_arrive_expect_tx(&full[s], C::kStageBytes);
              tma_load_5d(sa, &tmA, &full[s], kc * kBlockK, xp, x0 + dx, yp, row0 + dy);
              tma_load_2d(sa + kABytes, &tmB, &full[s], (t * a.kc_per_tap + kc) * kBlockK, n0);
            }
            if (++s == kStages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only in a pair)
    constexpr uint32_t idesc = umma_idesc_f16(kFmt, kTileM * kCG, kN);
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int unit = unit0; unit < total_units; unit += unit_step) {
      mbar_wait(&tempty[as], aph ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + uint32_t(as * kN);
      for (int k = 0; k < kiters; ++k) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + size_t(s) * C::kStageBytes);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int j = 0; j < kBlockK / 16; ++j) {  // +32 bytes (2 x 16 B) per K=16 step inside the swizzle row
            if (kCG == 2) umma_f16_cg2(d_tmem, da + uint64_t(j * 2), db + uint64_t(j * 2), idesc, (k | j) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, da + uint64_t(j * 2), db + uint64_t(j * 2), idesc, (k | j) != 0 ? 1u : 0u);
          }
          if (kCG == 2) {
            umma_commit_mc(&empty[s], 3);           // frees the stage in BOTH CTAs once these MMAs retire
            if (k == kiters - 1) umma_commit_mc(&tfull[as], 3);
          } else {
            umma_commit(&empty[s]);                 // frees the smem stage once these MMAs retire
            if (k == kiters - 1) umma_commit(&tfull[as]);
          }
        }
        __syncwarp();
        if (++s == kStages) {
          s = 0;
          ph ^= 1;
        }
      }
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;             // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;   // two warps share a quarter: each takes half of the kN columns
    const int m = q * 32 + lane;        // accumulator row = pixel inside the tile
    const int ry = m >> a.log2Wt, rx = m & (a.Wt - 1);
    const int et = threadIdx.x - (kThreads - kEpiThreads);
    int as = 0;
    uint32_t aph = 0;
    for (int unit = unit0; unit < total_units; unit += unit_step) {
      const TileCoord tc = decode_tile(unit_tile(unit), a);
      const int gy = tc.ty * a.R + ry, gx = tc.tx * a.Wt + rx;
      T* orow = reinterpret_cast<T*>(a.out) + size_t(tc.b) * a.out_bs +
                size_t(gy * a.oy_mul + a.oy_off[tc.phase]) * a.out_rs +
                size_t(gx * a.ox_mul + a.ox_off[tc.phase]) * a.out_ps + tc.nb * kN;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
#pragma unroll 1
      for (int ch = half * (kN / 64); ch < (half + 1) * (kN / 64); ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * kN + ch * 32), v);
        tmem_ld_wait();
        float r[32];
        if (a.epi_mode == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float f = __uint_as_float(v[i]) + __ldg(a.bias + tc.nb * kN + ch * 32 + i);
            r[i] = f > 0.f ? f : 0.2f * f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(v[i]);
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = Cvt<T>::pack2(r[2 * i], r[2 * i + 1]);
        uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        if (a.partials != nullptr) {
          // statistics of the values as stored (rounded to T), so that (y - mean) is exactly centred
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 f2 = Cvt<T>::unpack2(pk[i]);
            r[2 * i] = f2.x;
            r[2 * i + 1] = f2.y;
          }
          float t[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) t[i] = r[i];
          const float s1 = butterfly_reduce<false>(t, lane);
#pragma unroll
          for (int i = 0; i < 32; ++i) t[i] = r[i] * r[i];
          const float s2 = butterfly_reduce<false>(t, lane);
          const float mx = butterfly_reduce<true>(r, lane);
          stat[(q * 3 + 0) * kN + ch * 32 + lane] = s1;
          stat[(q * 3 + 1) * kN + ch * 32 + lane] = s2;
          stat[(q * 3 + 2) * kN + ch * 32 + lane] = mx;
        }
      }
      tc_fence_before();
      if (kCG == 2 && rank != 0) mbar_arrive_remote(&tempty[as], 0);  // the leader's MMA warp owns the TMEM pipeline
      else mbar_arrive(&tempty[as]);  // accumulator stage may be overwritten by the next-but-one tile
      if (a.partials != nullptr) {
        named_bar_sync(1, kEpiThreads);
        float* pdst = a.partials + (size_t(tc.b) * tiles_m_per_sample + tc.tile_m) * 3 * a.Cout + tc.nb * kN;
        for (int idx = et; idx < 3 * kN; idx += kEpiThreads) {
          const int which = idx / kN, col = idx - which * kN;
          const float v0 = stat[(0 * 3 + which) * kN + col], v1 = stat[(1 * 3 + which) * kN + col];
          const float v2 = stat[(2 * 3 + which) * kN + col], v3 = stat[(3 * 3 + which) * kN + col];
          pdst[which * a.Cout + col] = which == 2 ? fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)) : (v0 + v1) + (v2 + v3);
        }
        named_bar_sync(1, kEpiThreads);
      }
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  }

  tc_fence_before();
  if (kCG == 2) cluster_sync_all(); else __syncthreads();  // pair: the peer may still signal barriers in this CTA
  if (warp == 2) {
    tc_fence_after();
    if (kCG == 2) tmem_dealloc_cg2(tmem_base, C::kTmemCols); else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int kN, typename T, int kCG>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmArgs& args, int grid,
                cudaStream_t stream) {
  using C = Cfg<kN, kCG>;
  static bool configured = false;
  auto kern = conv_gemm_kernel<kN, T, kCG>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(C::kSmemBytes));
    if (e != cudaSuccess) return fail(DUCOSY_ERR_CUDA, "cudaFuncSetAttribute(conv_gemm): %s", cudaGetErrorString(e));
    configured = true;
  }
  if (kCG == 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, args);
    if (e != cudaSuccess) return fail(DUCOSY_ERR_CUDA, "cudaLaunchKernelEx(conv_gemm pair): %s", cudaGetErrorString(e));
  } else {
    kern<<<grid, kThreads, C::kSmemBytes, stream>>>(tmA, tmB, args);
  }
  return check_launch("conv_gemm_kernel");
}

}  // namespace

int launch_conv_gemm(const ConvPlan& p, cudaStream_t stream) {
  DUCOSY_CHECK(p.Cin % kBlockK == 0, DUCOSY_ERR_SHAPE, "conv_gemm: Cin=%d must be a multiple of 64", p.Cin);
  DUCOSY_CHECK(p.Cout % 64 == 0, DUCOSY_ERR_SHAPE, "conv_gemm: Cout=%d must be a multiple of 64", p.Cout);
  DUCOSY_CHECK(p.num_taps >= 1 && p.num_taps <= kMaxTaps && p.num_phases >= 1 && p.num_phases <= kMaxPhases,
               DUCOSY_ERR_SHAPE, "conv_gemm: taps/phases out of range");
  DUCOSY_CHECK(p.stride == 1 || (p.stride == 2 && p.Hp % 2 == 0 && p.Wp % 2 == 0), DUCOSY_ERR_SHAPE,
               "conv_gemm: stride must be 1, or 2 with even padded extents");
  const int Wt = p.Wg < kTileM ? p.Wg : kTileM;
  DUCOSY_CHECK(Wt >= 8 && (Wt & (Wt - 1)) == 0 && p.Wg % Wt == 0, DUCOSY_ERR_SHAPE,
               "conv_gemm: GEMM-grid width %d must be 8/16/32/64 or a multiple of 128", p.Wg);
  const int R = kTileM / Wt;
  DUCOSY_CHECK(p.Hg % R == 0, DUCOSY_ERR_SHAPE, "conv_gemm: GEMM-grid height %d must be a multiple of %d", p.Hg, R);
  DUCOSY_CHECK((reinterpret_cast<uintptr_t>(p.in) & 127) == 0 && (reinterpret_cast<uintptr_t>(p.w) & 127) == 0 &&
                   (reinterpret_cast<uintptr_t>(p.out) & 15) == 0,
               DUCOSY_ERR_ALIGN, "conv_gemm: input/weight buffers must be 128-byte aligned, output 16-byte");

  const int kN = p.Cout >= 256 ? 256 : p.Cout;
  DUCOSY_CHECK(kN == 64 || kN == 128 || kN == 256, DUCOSY_ERR_SHAPE, "conv_gemm: Cout=%d unsupported", p.Cout);
  DUCOSY_CHECK(p.Cout % kN == 0, DUCOSY_ERR_SHAPE, "conv_gemm: Cout=%d unsupported", p.Cout);

  ConvGemmArgs a{};
  a.num_phases = p.num_phases;
  a.num_taps = p.num_taps;
  a.kc_per_tap = p.Cin / kBlockK;
  a.n_blocks = p.Cout / kN;
  a.B = p.B;
  a.R = R;
  a.Wt = Wt;
  a.log2Wt = 0;
  while ((1 << a.log2Wt) < Wt) ++a.log2Wt;
  a.TY = p.Hg / R;
  a.TX = p.Wg / Wt;
  a.rows_per_sample = p.stride == 1 ? p.Hp : p.Hp / 2;
  a.Cout = p.Cout;
  for (int ph = 0; ph < p.num_phases; ++ph) {
    for (int t = 0; t < p.num_taps; ++t) {
      const int dy = p.tap_dy[ph][t], dx = p.tap_dx[ph][t];
      DUCOSY_CHECK(dy >= 0 && dx >= 0, DUCOSY_ERR_SHAPE, "conv_gemm: negative tap offset");
      if (p.stride == 1) {
        a.tap_xp[ph][t] = 0; a.tap_dx[ph][t] = int8_t(dx);
        a.tap_yp[ph][t] = 0; a.tap_dy[ph][t] = int8_t(dy);
      } else {
        a.tap_xp[ph][t] = int8_t(dx & 1); a.tap_dx[ph][t] = int8_t(dx >> 1);
        a.tap_yp[ph][t] = int8_t(dy & 1); a.tap_dy[ph][t] = int8_t(dy >> 1);
      }
    }
    a.oy_off[ph] = p.oy_off[ph];
    a.ox_off[ph] = p.ox_off[ph];
  }
  a.out = p.out;
  a.out_bs = (long long)p.Ho * p.Wo * p.Cout;
  a.out_rs = p.Wo * p.Cout;
  a.out_ps = p.Cout;
  a.oy_mul = p.oy_mul;
  a.ox_mul = p.ox_mul;
  a.partials = p.partials;
  a.bias = p.bias;
  a.epi_mode = p.epi_mode;

  // CTA pairs whenever the m-tiles of one (sample, phase) pair up; DUCOSY_CONV_CTA_GROUP=1 forces single CTAs.
  static const int cg_env = []() { const char* e = getenv("DUCOSY_CONV_CTA_GROUP"); return e ? atoi(e) : 2; }();
  const int cg = (cg_env == 2 && (a.TY * a.TX) % 2 == 0) ? 2 : 1;

  EncodeTiledFn encode = get_encode_fn();
  DUCOSY_CHECK(encode != nullptr, DUCOSY_ERR_CUDA, "conv_gemm: cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  const CUtensorMapDataType dt = p.dtype == DUCOSY_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmA, tmB;
  {
    const cuuint64_t C2 = cuuint64_t(p.Cin) * 2, W = cuuint64_t(p.Wp), H = cuuint64_t(p.Hp);
    cuuint64_t gdim[5], gstr[4];
    if (p.stride == 1) {
      gdim[0] = p.Cin; gdim[1] = 1; gdim[2] = W; gdim[3] = 1; gdim[4] = cuuint64_t(p.B) * H;
      gstr[0] = C2; gstr[1] = C2; gstr[2] = W * C2; gstr[3] = W * C2;
    } else {
      gdim[0] = p.Cin; gdim[1] = 2; gdim[2] = W / 2; gdim[3] = 2; gdim[4] = cuuint64_t(p.B) * H / 2;
      gstr[0] = C2; gstr[1] = 2 * C2; gstr[2] = W * C2; gstr[3] = 2 * W * C2;
    }
    const cuuint32_t box[5] = {cuuint32_t(kBlockK), 1, cuuint32_t(Wt), 1, cuuint32_t(R)};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmA, dt, 5, const_cast<void*>(p.in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DUCOSY_CHECK(r == CUDA_SUCCESS, DUCOSY_ERR_CUDA, "conv_gemm: cuTensorMapEncodeTiled(A) failed with %d", int(r));
  }
  {
    const cuuint64_t Ktot = cuuint64_t(p.num_taps) * p.Cin;
    const cuuint64_t gdim[2] = {Ktot, cuuint64_t(p.num_phases) * p.Cout};
    const cuuint64_t gstr[1] = {Ktot * 2};
    const cuuint32_t box[2] = {cuuint32_t(kBlockK), cuuint32_t(kN / cg)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmB, dt, 2, const_cast<void*>(p.w), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DUCOSY_CHECK(r == CUDA_SUCCESS, DUCOSY_ERR_CUDA, "conv_gemm: cuTensorMapEncodeTiled(B) failed with %d", int(r));
  }

  const int total_tiles = a.B * a.num_phases * a.TY * a.TX * a.n_blocks;
  const int sms = num_sms();
  DUCOSY_CHECK(sms > 0, DUCOSY_ERR_CUDA, "conv_gemm: no CUDA device");
  int grid = total_tiles < sms ? total_tiles : sms;
  if (grid == 0) return 0;
  if (cg == 2) grid &= ~1;  // whole pairs (total_tiles is even here)

#define DUCOSY_LAUNCH_N(N)                                                                             \
  if (cg == 2) {                                                                                       \
    if (p.dtype == DUCOSY_F16) return launch_impl<N, __half, 2>(tmA, tmB, a, grid, stream);             \
    else return launch_impl<N, __nv_bfloat16, 2>(tmA, tmB, a, grid, stream);                            \
  } else {                                                                                             \
    if (p.dtype == DUCOSY_F16) return launch_impl<N, __half, 1>(tmA, tmB, a, grid, stream);             \
    else return launch_impl<N, __nv_bfloat16, 1>(tmA, tmB, a, grid, stream);                            \
  }
  if (kN == 256) { DUCOSY_LAUNCH_N(256) }
  if (kN == 128) { DUCOSY_LAUNCH_N(128) }
  DUCOSY_LAUNCH_N(64)
#undef DUCOSY_LAUNCH_N
}

}  // namespace ducosy
