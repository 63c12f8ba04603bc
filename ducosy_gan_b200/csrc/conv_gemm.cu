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
// Warp roles (384 threads, 1 CTA/SM): warps 0-7 = epilogue, warp 8 = TMA producer, warp 9 = MMA issuer (one lane),
// warp 10 = TMEM allocator.  Epilogue: TMEM -> registers -> 16-bit tile staged in shared memory (TMA swizzle) ->
// one TMA store per 64-channel slab, and the InstanceNorm / CBAM-pool statistics (per-tile per-channel sum, sum of
// squares, max) as a conflict-free column pass over the staged tile.  The first version stored straight from
// registers and reduced with warp shuffles: 4096 uncoalesced store wavefronts + ~3000 shuffles per tile on the
// LSU / shared-memory crossbar cost the UMMA operand fetch 17-22 % (measured, tools/conv_bench.py).
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace ducosy {

namespace {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // 64 x 16-bit = one 128-byte swizzle row
constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KiB
constexpr int kThreads = 384;                   // 8 epilogue warps + 4 control warps
constexpr int kWarpTma = 8, kWarpMma = 9, kWarpAlloc = 10;
constexpr int kEpiThreads = 256;

// kCG = 1: one CTA per 128 x kN tile.  kCG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256 x kN
// tile; each CTA stages its own 128 A rows and HALF of the B rows, so L2->SMEM traffic per FLOP drops by a third
// and the same shared memory holds 6 instead of 4 pipeline stages.
template <int kN, int kCG, bool kSplit = false>
struct Cfg {
  static constexpr int kPlanes = kSplit ? 2 : 1;  // split-operand mode stages a (hi, lo) pair of 16-bit tiles
  static constexpr int kBRows = kN / kCG;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * kN;        // two accumulator stages (power of two >= 32)
  // epilogue staging: the 16-bit output tile as kSlabs x [128 rows][64 channels] in the TMA 128-byte-swizzle layout
  static constexpr int kSlabs = kN / 64;
  static constexpr int kSlabBytes = kTileM * 128;
  static constexpr int kPlaneBytes = kSlabs * kSlabBytes;
  static constexpr int kOutBytes = kPlanes * kPlaneBytes;
  static constexpr int kParts = 8 / kSlabs;       // row parts per slab in the column pass (8 epilogue warps)
  static constexpr int kStatFloats = kParts * 3 * kN;
  static constexpr int kFixedBytes = 1024 + kOutBytes + kStatFloats * 4 + 256;
  static constexpr int kStagesRaw = (227 * 1024 - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr size_t kSmemBytes = size_t(kFixedBytes) + size_t(kStages) * kStageBytes;
};

__device__ __forceinline__ int atom_add_acq_rel_gpu(int* p, int v) {
  int old;
  asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
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

// InstanceNorm finalize of sample b by the 256 epilogue threads of the CTA that completed the sample's last tile (see
// ConvFinalize in common.cuh).  `scratch` is the output staging area (free between two tiles), `tiles` the number of m-tiles
// per sample.  Partials were written by other SMs: they are read through L2 (ld.global.cg).
__device__ void finalize_sample(const ConvGemmArgs& a, int b, int tiles, uint8_t* scratch, int et) {
  const int C = a.Cstore;
  const int Cw = C < kEpiThreads ? C : kEpiThreads;    // channels handled side by side
  const int G = kEpiThreads / Cw;                      // tile groups per channel (64 ch: 4, 128: 2, >= 256: 1)
  const int cl = et % Cw, g = et / Cw;
  double* sd = reinterpret_cast<double*>(scratch);                        // [G][2][Cw]
  float* sm = reinterpret_cast<float*>(scratch + size_t(G) * 2 * Cw * 8); // [G][Cw] max, then [C] smax + [C/16] hidden
  const float* base = a.partials + size_t(b) * tiles * 3 * C;
  float my_scale = 0.f, my_shift = 0.f;   // C <= 256: thread (g == 0, cl) keeps its channel's pair for the MLP fold
  for (int c0 = 0; c0 < C; c0 += Cw) {
    const int c = c0 + cl;
    const float* p = base + c;
    double s1 = 0.0, s2 = 0.0;
    float mx = -INFINITY;
    constexpr int kU = 16;      // independent L2 loads in flight per thread: one CTA reduces a whole sample, latency bound
    for (int t0 = g; t0 < tiles; t0 += G * kU) {
      float v1[kU], v2[kU], vm[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int t = t0 + u * G;
        const bool ok = t < tiles;
        v1[u] = ok ? __ldcg(p + size_t(t * 3 + 0) * C) : 0.f;
        v2[u] = ok ? __ldcg(p + size_t(t * 3 + 1) * C) : 0.f;
        vm[u] = ok ? __ldcg(p + size_t(t * 3 + 2) * C) : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        s1 += double(v1[u]);
        s2 += double(v2[u]);
        mx = fmaxf(mx, vm[u]);
      }
    }
    if (G > 1) {
      sd[(g * 2 + 0) * Cw + cl] = s1;
      sd[(g * 2 + 1) * Cw + cl] = s2;
      sm[g * Cw + cl] = mx;
      named_bar_sync(1, kEpiThreads);
      if (g == 0)
        for (int k = 1; k < G; ++k) {      // fixed order: deterministic
          s1 += sd[(k * 2 + 0) * Cw + cl];
          s2 += sd[(k * 2 + 1) * Cw + cl];
          mx = fmaxf(mx, sm[k * Cw + cl]);
        }
    }
    if (g == 0) {
      const double mean = s1 / a.fin.npix;
      double var = s2 / a.fin.npix - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = float(1.0 / sqrt(var + 1e-5));
      const float fmean = float(mean);
      my_scale = rstd;
      my_shift = -fmean * rstd;
      const float nmax = (mx - fmean) * rstd;          // max over H*W of the normalised map (rstd > 0)
      if (a.fin.chmax != nullptr) a.fin.chmax[size_t(b) * C + c] = nmax;
      if (a.fin.fc0 == nullptr) {
        a.fin.scale[size_t(b) * C + c] = my_scale;
        a.fin.shift[size_t(b) * C + c] = my_shift;
      } else {
        sm[c] = nmax;                                  // C == Cw == 256 here (checked on the host): G == 1, sm is free
      }
    }
    if (G > 1) named_bar_sync(1, kEpiThreads);         // scratch is reused by the next channel block
  }
  if (a.fin.fc0 != nullptr) {
    // CBAM channel attention behind a non-affine InstanceNorm (modules/model.py:20-24): the avg-pool branch sees a zero-mean
    // map and contributes fc(0) = 0, so att = sigmoid(fc2 . relu(fc0 . max)); folded into (scale, shift).
    const int Hd = C / 16;
    float* smax = sm;
    float* hidden = sm + C;
    named_bar_sync(1, kEpiThreads);
    const int warp = et >> 5, lane = et & 31;
    for (int j = warp; j < Hd; j += kEpiThreads / 32) {
      float acc = 0.f;
      for (int c = lane; c < C; c += 32) acc += __ldg(a.fin.fc0 + j * C + c) * smax[c];
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) hidden[j] = fmaxf(acc, 0.f);
    }
    named_bar_sync(1, kEpiThreads);
    if (et < C) {
      float acc = 0.f;
      for (int j = 0; j < Hd; ++j) acc += __ldg(a.fin.fc2 + et * Hd + j) * hidden[j];
      const float att = 1.f / (1.f + __expf(-acc));
      a.fin.scale[size_t(b) * C + et] = my_scale * att;
      a.fin.shift[size_t(b) * C + et] = my_shift * att;
    }
    named_bar_sync(1, kEpiThreads);
  }
}

template <int kN, typename T, int kCG, bool kSplit>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const ConvGemmArgs a) {
  using C = Cfg<kN, kCG, kSplit>;
  constexpr int kStages = C::kStages;
  constexpr int kFmt = sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value ? 1 : 0;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B wants 1024-B alignment
  uint8_t* otile = smem + size_t(kStages) * C::kStageBytes;  // 1024-byte aligned (stage sizes are multiples of 1 KiB)
  float* stat = reinterpret_cast<float*>(otile + C::kOutBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(stat + C::kStatFloats);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  volatile int* fin_slot = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [2] parked tickets (sample + 1, or 0)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kiters = a.num_taps * a.kc_per_tap;
  const int tiles_m_per_sample = a.num_phases * a.TY * a.TX;
  // work unit = (group of kCG consecutive m-tiles) x (n-block); this CTA owns m-tile  unit_m * kCG + rank
  const int total_units = (a.B * tiles_m_per_sample / kCG) * a.n_blocks;
  const uint32_t rank = kCG == 2 ? cluster_ctarank() : 0u;
  const int unit0 = blockIdx.x / kCG, unit_step = gridDim.x / kCG;
  auto unit_tile = [&](int unit) { return ((unit / a.n_blocks) * kCG + int(rank)) * a.n_blocks + unit % a.n_blocks; };

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == kWarpMma && lane == 0) {
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
  if (warp == kWarpAlloc) {
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
  // programmatic dependent launch (common.cuh): TMEM is allocated, so the next kernel's CTAs may come in; the barrier /
  // TMEM / descriptor set-up above overlapped the tail of the previous kernel, whose results are first touched below
  pdl_launch_dependents();

  if (warp == kWarpTma) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      pdl_wait();
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
          const int ca = a.tap_c0[tc.phase][t], kb = a.tap_bk[t];   // 64-channel chunk offsets of the A plane / the B columns
          for (int kc = 0; kc < a.kc_per_tap; ++kc) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sa = smem + size_t(s) * C::kStageBytes;
            if (kCG == 2) {
              // both CTAs' loads complete on the leader's barrier, which expects the bytes of the whole pair
              if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * C::kStageBytes);
              tma_load_5d_cg2(sa, &tmA, &full[s], (ca + kc) * kBlockK, xp, x0 + dx, yp, row0 + dy);
              tma_load_2d_cg2(sa + kABytes, &tmB, &full[s], (kb + kc) * kBlockK, n0);
            } else {
              mbar_arrive_expect_tx(&full[s], C::kStageBytes);
              tma_load_5d(sa, &tmA, &full[s], (ca + kc) * kBlockK, xp, x0 + dx, yp, row0 + dy);
              tma_load_2d(sa + kABytes, &tmB, &full[s], (kb + kc) * kBlockK, n0);
            }
            if (++s == kStages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == kWarpMma && rank == 0) {
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
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;             // TMEM lane quarter this warp may read
    const int half = warp >> 2;         // two warps share a quarter: each takes half of the kN columns
    const int m = q * 32 + lane;        // accumulator row = pixel inside the tile
    const int et = threadIdx.x;
    // column pass role: this warp reduces rows [cp_part*kRows, +kRows) of slab cp_slab; lane owns 2 adjacent columns
    constexpr int kRows = kTileM / C::kParts;
    const int cp_slab = warp / C::kParts, cp_part = warp % C::kParts;
    int as = 0;
    int fin_n = 0;                       // tiles this CTA has finished (parity selects the ticket slot)
    uint32_t aph = 0;
    pdl_wait();                          // before the first global store (output tile, partials, tickets)
    for (int unit = unit0; unit < total_units; unit += unit_step) {
      const TileCoord tc = decode_tile(unit_tile(unit), a);
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
#pragma unroll 1
      for (int ch = half * (kN / 64); ch < (half + 1) * (kN / 64); ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * kN + ch * 32), v);
        tmem_ld_wait();
        float r[32];
        if ((a.epi_mode & 3) == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float f = __uint_as_float(v[i]) + __ldg(a.bias + tc.nb * kN + ch * 32 + i);
            r[i] = f > 0.f ? f : 0.2f * f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(v[i]);
        }
        // stage the rounded values: row m of slab ch/2, 16-byte chunk c at (c ^ (m & 7)) -- conflict-free STS.128
        uint8_t* rowp = otile + (ch >> 1) * C::kSlabBytes + m * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 pk = make_uint4(Cvt<T>::pack2(r[8 * i], r[8 * i + 1]), Cvt<T>::pack2(r[8 * i + 2], r[8 * i + 3]),
                                      Cvt<T>::pack2(r[8 * i + 4], r[8 * i + 5]), Cvt<T>::pack2(r[8 * i + 6], r[8 * i + 7]));
          *reinterpret_cast<uint4*>(rowp + ((((ch & 1) * 4 + i) ^ (m & 7)) << 4)) = pk;
          if (kSplit) {   // lo plane: what the 16-bit rounding of the hi plane dropped (v - hi is exact in fp32)
            uint32_t hw[4] = {pk.x, pk.y, pk.z, pk.w}, lw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 h = Cvt<T>::unpack2(hw[j]);
              lw[j] = Cvt<T>::pack2(r[8 * i + 2 * j] - h.x, r[8 * i + 2 * j + 1] - h.y);
            }
            *reinterpret_cast<uint4*>(rowp + C::kPlaneBytes + ((((ch & 1) * 4 + i) ^ (m & 7)) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
      tc_fence_before();
      if (kCG == 2 && rank != 0) mbar_arrive_remote(&tempty[as], 0);  // the leader's MMA warp owns the TMEM pipeline
      else mbar_arrive(&tempty[as]);  // accumulator stage may be overwritten by the next-but-one tile
      fence_proxy_async_smem();        // generic-proxy writes above -> visible to the TMA (async proxy) store below
      named_bar_sync(1, kEpiThreads);
      if (et == 0) {
#pragma unroll
        for (int sb = 0; sb < C::kSlabs; ++sb) {
          const int n = tc.nb * kN + sb * 64;
          const int f = n / a.Cstore;  // merged phases: column block -> output phase; otherwise 0
#pragma unroll
          for (int pl = 0; pl < C::kPlanes; ++pl)   // split mode: the lo plane lives Cstore channels behind the hi plane
            tma_store_5d(&tmO, otile + pl * C::kPlaneBytes + sb * C::kSlabBytes, pl * a.Cstore + n - f * a.Cstore,
                         a.fold > 1 ? (f & 1) : a.ox_off[tc.phase], tc.tx * a.Wt + a.out_x_off,
                         a.fold > 1 ? (f >> 1) : a.oy_off[tc.phase], tc.b * a.out_rows + tc.ty * a.R + a.out_y_off);
        }
        tma_store_commit();
      }
      if (a.partials != nullptr) {
        // statistics of the values as stored (rounded to T), so that (y - mean) is exactly centred
        const uint8_t* sl = otile + cp_slab * C::kSlabBytes + (lane & 3) * 4;
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f, mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll 8
        for (int rr = 0; rr < kRows; ++rr) {
          const int mr = cp_part * kRows + rr;
          float2 f = Cvt<T>::unpack2(*reinterpret_cast<const uint32_t*>(sl + mr * 128 + (((lane >> 2) ^ (mr & 7)) << 4)));
          if (kSplit) {
            const float2 l = Cvt<T>::unpack2(*reinterpret_cast<const uint32_t*>(sl + C::kPlaneBytes + mr * 128 + (((lane >> 2) ^ (mr & 7)) << 4)));
            f.x += l.x;
            f.y += l.y;
          }
          s1a += f.x;
          s1b += f.y;
          s2a = fmaf(f.x, f.x, s2a);
          s2b = fmaf(f.y, f.y, s2b);
          mxa = fmaxf(mxa, f.x);
          mxb = fmaxf(mxb, f.y);
        }
        const int col = cp_slab * 64 + lane * 2;
        *reinterpret_cast<float2*>(&stat[(cp_part * 3 + 0) * kN + col]) = make_float2(s1a, s1b);
        *reinterpret_cast<float2*>(&stat[(cp_part * 3 + 1) * kN + col]) = make_float2(s2a, s2b);
        *reinterpret_cast<float2*>(&stat[(cp_part * 3 + 2) * kN + col]) = make_float2(mxa, mxb);
        named_bar_sync(1, kEpiThreads);
        // per-tile partial row [3][Cstore]; merged phases fold their `fold` column blocks into the same channels
        const int ncols = kN / a.fold;
        float* pdst = a.partials + (size_t(tc.b) * tiles_m_per_sample + tc.tile_m) * 3 * a.Cstore + (tc.nb * kN) % a.Cstore;
        for (int idx = et; idx < 3 * ncols; idx += kEpiThreads) {
          const int which = idx / ncols, col2 = idx - which * ncols;
          float acc = which == 2 ? -INFINITY : 0.f;
          for (int f = 0; f < a.fold; ++f)
#pragma unroll
            for (int pp = 0; pp < C::kParts; ++pp) {
              const float o = stat[(pp * 3 + which) * kN + f * ncols + col2];
              acc = which == 2 ? fmaxf(acc, o) : acc + o;
            }
          pdst[which * a.Cstore + col2] = acc;
        }
      }
      if (et == 0) tma_store_wait_read();  // the staged tile may be overwritten once the TMA engine has read it
      named_bar_sync(1, kEpiThreads);
      if (a.fin.scale != nullptr) {
        // Ticket: the CTA that completes the last (m-tile, n-block) of a sample finalizes that sample's InstanceNorm
        // statistics.  The barrier above orders every thread's partial-row stores before the ticket thread, whose
        // acq_rel atomic (cumulative) publishes them device-wide; it is a thread of warp 1, not the thread that issued
        // the TMA stores.  The answer is parked in shared memory and acted on ONE TILE LATER (after that tile's own
        // barrier), so the round trip of the atomic never sits on the epilogue's critical path.
        if (fin_n > 0 && fin_slot[(fin_n - 1) & 1] != 0) finalize_sample(a, fin_slot[(fin_n - 1) & 1] - 1, tiles_m_per_sample, otile, et);
        if (et == 32) {
          const int total = tiles_m_per_sample * a.n_blocks;
          const int old = atom_add_acq_rel_gpu(a.fin.counter + tc.b, 1);
          const int last = old == total - 1;
          if (last) a.fin.counter[tc.b] = 0;   // self-resetting: the next launch on this workspace starts from zero again
          fin_slot[fin_n & 1] = last ? tc.b + 1 : 0;
        }
        ++fin_n;
      }
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
    if (a.fin.scale != nullptr && fin_n > 0) {   // the ticket of this CTA's last tile
      named_bar_sync(1, kEpiThreads);
      if (fin_slot[(fin_n - 1) & 1] != 0) finalize_sample(a, fin_slot[(fin_n - 1) & 1] - 1, tiles_m_per_sample, otile, et);
    }
  }

  tc_fence_before();
  if (kCG == 2) cluster_sync_all(); else __syncthreads();  // pair: the peer may still signal barriers in this CTA
  if (warp == kWarpAlloc) {
    tc_fence_after();
    if (kCG == 2) tmem_dealloc_cg2(tmem_base, C::kTmemCols); else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
template <int kN, typename T, int kCG, bool kSplit = false>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvGemmArgs& args,
                int grid, cudaStream_t stream) {
  using C = Cfg<kN, kCG, kSplit>;
  static PerDeviceOnce configured;
  auto kern = conv_gemm_kernel<kN, T, kCG, kSplit>;
  {
    const cudaError_t e = configured.once([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(C::kSmemBytes)); });
    if (e != cudaSuccess) return fail(DUCOSY_ERR_CUDA, "cudaFuncSetAttribute(conv_gemm): %s", cudaGetErrorString(e));
  }
  if (kCG == 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, args);
    if (e != cudaSuccess) return fail(DUCOSY_ERR_CUDA, "cudaLaunchKernelEx(conv_gemm pair): %s", cudaGetErrorString(e));
  } else {
    pdl(kern, grid, kThreads, C::kSmemBytes, stream)(tmA, tmB, tmO, args);
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

  // Split-operand mode (DUCOSY_F16X2, see include/ducosy.h): every activation and weight is a (hi, lo) pair of fp16 values,
  // hi = rn(v), lo = rn(v - hi).  The input holds the lo plane Cin channels behind the hi plane (2*Cin channels per pixel), the
  // packed weight holds [all taps hi | all taps lo] along K, and every filter tap becomes three "virtual taps"
  // A_hi*W_hi + A_lo*W_hi + A_hi*W_lo accumulated in fp32 (the dropped A_lo*W_lo term is ~2^-22 relative).  The output is
  // written as a (hi, lo) pair again, so the staged tile is twice as large: N is capped at 128.
  const bool split = p.dtype == DUCOSY_F16X2;
  const int planes = split ? 2 : 1;
  const int kN = split ? (p.Cout >= 128 ? 128 : p.Cout) : (p.Cout >= 256 ? 256 : p.Cout);
  DUCOSY_CHECK(!split || (p.num_taps * 3 <= kMaxVTaps && p.fold <= 1 && p.epi_mode == 0), DUCOSY_ERR_SHAPE,
               "conv_gemm: split-operand mode supports at most %d taps, no merged phases, no bias epilogue", kMaxVTaps / 3);
  DUCOSY_CHECK(kN == 64 || kN == 128 || kN == 256, DUCOSY_ERR_SHAPE, "conv_gemm: Cout=%d unsupported", p.Cout);
  DUCOSY_CHECK(p.Cout % kN == 0, DUCOSY_ERR_SHAPE, "conv_gemm: Cout=%d unsupported", p.Cout);

  ConvGemmArgs a{};
  a.num_phases = p.num_phases;
  a.num_taps = split ? 3 * p.num_taps : p.num_taps;
  a.split = split ? 1 : 0;
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
  a.fold = p.fold > 1 ? p.fold : 1;
  a.Cstore = p.Cout / a.fold;
  DUCOSY_CHECK(a.fold == 1 || (a.fold == 4 && p.num_phases == 1 && p.Cout == kN && a.Cstore % 64 == 0 && p.oy_mul == 2),
               DUCOSY_ERR_SHAPE, "conv_gemm: merged phases need fold 4, one n-block, 64-channel slabs");
  DUCOSY_CHECK((p.num_taps * planes) * a.kc_per_tap < 32768 && a.kc_per_tap < 128, DUCOSY_ERR_SHAPE, "conv_gemm: K too large");
  for (int ph = 0; ph < p.num_phases; ++ph) {
    for (int v = 0; v < a.num_taps; ++v) {
      // virtual tap v of filter tap t: j = 0 A_hi*W_hi, 1 A_lo*W_hi, 2 A_hi*W_lo (not split: v == t, j == 0)
      const int t = split ? v / 3 : v, j = split ? v % 3 : 0;
      const int dy = p.tap_dy[ph][t], dx = p.tap_dx[ph][t];
      DUCOSY_CHECK(dy >= 0 && dx >= 0, DUCOSY_ERR_SHAPE, "conv_gemm: negative tap offset");
      if (p.stride == 1) {
        a.tap_xp[ph][v] = 0; a.tap_dx[ph][v] = int8_t(dx);
        a.tap_yp[ph][v] = 0; a.tap_dy[ph][v] = int8_t(dy);
      } else {
        a.tap_xp[ph][v] = int8_t(dx & 1); a.tap_dx[ph][v] = int8_t(dx >> 1);
        a.tap_yp[ph][v] = int8_t(dy & 1); a.tap_dy[ph][v] = int8_t(dy >> 1);
      }
      a.tap_c0[ph][v] = int8_t(j == 1 ? a.kc_per_tap : 0);
      a.tap_bk[v] = int16_t((j == 2 ? p.num_taps + t : t) * a.kc_per_tap);
    }
    a.oy_off[ph] = p.oy_off[ph];
    a.ox_off[ph] = p.ox_off[ph];
  }
  a.out = p.out;
  a.out_bs = (long long)p.Ho * p.Wo * a.Cstore;
  a.out_rs = p.Wo * a.Cstore;
  a.out_ps = a.Cstore;
  a.oy_mul = p.oy_mul;
  a.ox_mul = p.ox_mul;
  a.out_rows = p.Ho / p.oy_mul;
  a.out_y_off = p.out_y_off;
  a.out_x_off = p.out_x_off;
  DUCOSY_CHECK(p.out_y_off >= 0 && p.out_x_off >= 0 && (p.Hg + p.out_y_off) * p.oy_mul <= p.Ho &&
                   (p.Wg + p.out_x_off) * p.ox_mul <= p.Wo,
               DUCOSY_ERR_SHAPE, "conv_gemm: GEMM grid does not fit the output image");
  a.partials = p.partials;
  a.bias = p.bias;
  a.epi_mode = p.epi_mode;
  a.fin = p.fin;
  if (p.fin.scale != nullptr) {
    DUCOSY_CHECK(p.partials != nullptr && p.fin.shift != nullptr && p.fin.counter != nullptr && p.fin.npix > 0 && p.epi_mode == 0,
                 DUCOSY_ERR_ARG, "conv_gemm: fused finalize needs partials, shift, counters and npix");
    DUCOSY_CHECK((p.fin.fc0 == nullptr) == (p.fin.fc2 == nullptr) && (p.fin.fc0 == nullptr || (a.Cstore == 256 && p.fin.chmax != nullptr)),
                 DUCOSY_ERR_ARG, "conv_gemm: the fused CBAM channel MLP needs fc0 + fc2 + chmax and 256 channels");
    DUCOSY_CHECK(a.Cstore % 64 == 0 && (a.Cstore <= 256 || a.Cstore % 256 == 0), DUCOSY_ERR_SHAPE, "conv_gemm: fused finalize: bad channel count");
  }

  // CTA pairs whenever the m-tiles of one (sample, phase) pair up; DUCOSY_CONV_CTA_GROUP=1 forces single CTAs.
  static const int cg_env = []() { const char* e = getenv("DUCOSY_CONV_CTA_GROUP"); return e ? atoi(e) : 2; }();
  const int cg = (cg_env == 2 && (a.TY * a.TX) % 2 == 0) ? 2 : 1;

  const CUtensorMapDataType dt = p.dtype == DUCOSY_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB, tmO;
  {
    // output tile store: (c, x-phase, x, y-phase, y*batch) so that the sub-pixel (x2 upsampling) phases are boxes too
    const cuuint64_t Cs = cuuint64_t(a.Cstore) * planes, C2 = Cs * 2, W = cuuint64_t(p.Wo), H = cuuint64_t(p.Ho);
    cuuint64_t gdim[5], gstr[4];
    DUCOSY_CHECK(p.oy_mul == p.ox_mul && (p.oy_mul == 1 || (p.oy_mul == 2 && p.Ho % 2 == 0 && p.Wo % 2 == 0)),
                 DUCOSY_ERR_SHAPE, "conv_gemm: output stride must be 1 or 2");
    if (p.oy_mul == 1) {
      gdim[0] = Cs; gdim[1] = 1; gdim[2] = W; gdim[3] = 1; gdim[4] = cuuint64_t(p.B) * H;
      gstr[0] = C2; gstr[1] = C2; gstr[2] = W * C2; gstr[3] = W * C2;
    } else {
      gdim[0] = Cs; gdim[1] = 2; gdim[2] = W / 2; gdim[3] = 2; gdim[4] = cuuint64_t(p.B) * H / 2;
      gstr[0] = C2; gstr[1] = 2 * C2; gstr[2] = W * C2; gstr[3] = 2 * W * C2;
    }
    const cuuint32_t box[5] = {64, 1, cuuint32_t(Wt), 1, cuuint32_t(R)};
    DUCOSY_TRY(encode_tiled_cached(&tmO, dt, 5, p.out, gdim, gstr, box, "conv_gemm(out)"));
  }
  {
    const cuuint64_t Ci = cuuint64_t(p.Cin) * planes, C2 = Ci * 2, W = cuuint64_t(p.Wp), H = cuuint64_t(p.Hp);
    cuuint64_t gdim[5], gstr[4];
    if (p.stride == 1) {
      gdim[0] = Ci; gdim[1] = 1; gdim[2] = W; gdim[3] = 1; gdim[4] = cuuint64_t(p.B) * H;
      gstr[0] = C2; gstr[1] = C2; gstr[2] = W * C2; gstr[3] = W * C2;
    } else {
      gdim[0] = Ci; gdim[1] = 2; gdim[2] = W / 2; gdim[3] = 2; gdim[4] = cuuint64_t(p.B) * H / 2;
      gstr[0] = C2; gstr[1] = 2 * C2; gstr[2] = W * C2; gstr[3] = 2 * W * C2;
    }
    const cuuint32_t box[5] = {cuuint32_t(kBlockK), 1, cuuint32_t(Wt), 1, cuuint32_t(R)};
    DUCOSY_TRY(encode_tiled_cached(&tmA, dt, 5, p.in, gdim, gstr, box, "conv_gemm(A)"));
  }
  {
    const cuuint64_t Ktot = cuuint64_t(p.num_taps) * p.Cin * planes;
    const cuuint64_t gdim[2] = {Ktot, cuuint64_t(p.num_phases) * p.Cout};
    const cuuint64_t gstr[1] = {Ktot * 2};
    const cuuint32_t box[2] = {cuuint32_t(kBlockK), cuuint32_t(kN / cg)};
    DUCOSY_TRY(encode_tiled_cached(&tmB, dt, 2, p.w, gdim, gstr, box, "conv_gemm(B)"));
  }

  const int total_tiles = a.B * a.num_phases * a.TY * a.TX * a.n_blocks;
  const int sms = num_sms();
  DUCOSY_CHECK(sms > 0, DUCOSY_ERR_CUDA, "conv_gemm: no CUDA device");
  int grid = total_tiles < sms ? total_tiles : sms;
  if (grid == 0) return 0;
  if (cg == 2) grid &= ~1;  // whole pairs (total_tiles is even here)

#define DUCOSY_LAUNCH_N(N)                                                                             \
  if (cg == 2) {                                                                                       \
    if (p.dtype == DUCOSY_F16) return launch_impl<N, __half, 2>(tmA, tmB, tmO, a, grid, stream);             \
    else return launch_impl<N, __nv_bfloat16, 2>(tmA, tmB, tmO, a, grid, stream);                            \
  } else {                                                                                             \
    if (p.dtype == DUCOSY_F16) return launch_impl<N, __half, 1>(tmA, tmB, tmO, a, grid, stream);             \
    else return launch_impl<N, __nv_bfloat16, 1>(tmA, tmB, tmO, a, grid, stream);                            \
  }
  if (split) {
    if (kN == 128) return cg == 2 ? launch_impl<128, __half, 2, true>(tmA, tmB, tmO, a, grid, stream)
                                  : launch_impl<128, __half, 1, true>(tmA, tmB, tmO, a, grid, stream);
    return cg == 2 ? launch_impl<64, __half, 2, true>(tmA, tmB, tmO, a, grid, stream)
                   : launch_impl<64, __half, 1, true>(tmA, tmB, tmO, a, grid, stream);
  }
  if (kN == 256) { DUCOSY_LAUNCH_N(256) }
  if (kN == 128) { DUCOSY_LAUNCH_N(128) }
  DUCOSY_LAUNCH_N(64)
#undef DUCOSY_LAUNCH_N
}

}  // namespace ducosy
