// Shared host/device declarations for libducosy_sm100.so
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/ducosy.h"

namespace ducosy {

// ---- error plumbing: every extern "C" entry returns 0 or a negative code; message is thread-local ----
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> code

#define DUCOSY_CHECK(cond, code, ...) \
  do {                                \
    if (!(cond)) return ::ducosy::fail((code), __VA_ARGS__); \
  } while (0)

#define DUCOSY_TRY(expr)        \
  do {                          \
    int _rc = (expr);           \
    if (_rc != 0) return _rc;   \
  } while (0)

int num_sms();

// Per-function attributes (cudaFuncSetAttribute) live in the device context: a process that drives several GPUs
// (nn.DataParallel replicas on Python threads) must set them once per device, not once per process.
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  // Runs f() (a cudaFuncSetAttribute call) unless it already SUCCEEDED on the current device.  The bit is published only
  // after f() returned, so a second host thread on the same device either sees the finished configuration or repeats
  // the (idempotent) call itself -- it can never launch with the attribute still unset.
  template <class F>
  cudaError_t once(F&& f) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return f();
    const unsigned long long bit = 1ull << dev;
    if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = f();
    if (e == cudaSuccess) mask.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

// ---- programmatic dependent launch (PDL) ----
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// pdl_prologue(): `griddepcontrol.launch_dependents` lets the NEXT kernel of the stream be scheduled while this one is still
// running (its CTAs become resident, run their own prologue and stop at their wait), `griddepcontrol.wait` blocks until the
// PREVIOUS kernel has completed and its memory operations are visible.  No global memory is touched before the wait, so the
// data flow is exactly that of plain stream order; what disappears is the launch gap between dependent kernels -- the
// batch-1 forward and the one-sample-per-rank training step are chains of 100 / 3000 kernels of a few microseconds each.
// Kernels that allocate TMEM trigger only AFTER their allocation: a dependent CTA that lands on the same SM first would
// hold TMEM columns while waiting for this very grid.  Without the launch attribute both instructions are no-ops, and
// that is the DEFAULT: the attribute is set only with DUCOSY_PDL=1 (measured: it helps eager batch-1 chains and hurts the
// two-stream paths, see pdl_enabled() in api.cu).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
bool pdl_enabled();

// kern<<<grid, block, smem, stream>>>(args...)  ==  pdl(kern, grid, block, smem, stream)(args...)
template <typename K>
struct PdlLaunch {
  K kern;
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
  template <typename... A>
  void operator()(A&&... args) const {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);   // errors surface through check_launch()
  }
};
template <typename K>
PdlLaunch<K> pdl(K kern, dim3 grid, dim3 block, size_t smem = 0, cudaStream_t stream = nullptr) {
  return PdlLaunch<K>{kern, grid, block, smem, stream};
}

constexpr int kMaxTaps = 16;
constexpr int kMaxVTaps = 32;   // "virtual" taps of the split-operand mode: 3 per filter tap (A_hi*W_hi, A_lo*W_hi, A_hi*W_lo)
constexpr int kMaxPhases = 4;

// InstanceNorm finalize fused into the tail of the convolution (conv_gemm.cu: finalize_sample).  The CTA whose tile is the
// LAST one of a sample to complete (a per-sample ticket counter, self-resetting) reduces that sample's per-tile partials in
// fixed order -- deterministic, the same arithmetic as in_finalize_kernel -- and writes scale = rstd, shift = -mean*rstd
// (and, with fc0/fc2, folds the CBAM channel attention of modules/model.py:20-24 into them).  This removes one or two
// dependent launches behind every convolution: ~30 % of the kernel count of a batch-1 forward.
struct ConvFinalize {
  float* scale;            // [B][Cstore]; nullptr = not fused (the caller launches ducosy_in_finalize)
  float* shift;            // [B][Cstore]
  float* chmax;            // [B][Cstore] normalised per-channel max (optional; required with fc0/fc2)
  const float* fc0;        // [Cstore/16][Cstore] or nullptr
  const float* fc2;        // [Cstore][Cstore/16] or nullptr
  int* counter;            // [B] tickets, zero before the first launch (the finalizing CTA resets its sample's ticket)
  int npix;                // pixels per sample and channel behind the statistics
};

// Arguments of the implicit-GEMM convolution kernel (see conv_gemm.cu for the meaning).
struct ConvGemmArgs {
  int num_phases, num_taps, kc_per_tap, n_blocks;
  int B, TY, TX;           // tiles per sample along the GEMM grid rows / cols
  int R, Wt, log2Wt;       // a tile is R rows x Wt cols of the GEMM grid (R*Wt == 128)
  int rows_per_sample;     // units of the outermost TMA dimension per sample
  int Cout;                // GEMM N total (= n_blocks * kN) = fold * Cstore
  int Cstore;              // channels of the output tensor
  int fold;                // 1, or 4: the four x2-upsampling phases are column blocks of ONE tile (merged phases)
  int8_t tap_xp[kMaxPhases][kMaxVTaps], tap_dx[kMaxPhases][kMaxVTaps];
  int8_t tap_yp[kMaxPhases][kMaxVTaps], tap_dy[kMaxPhases][kMaxVTaps];
  int8_t tap_c0[kMaxPhases][kMaxVTaps];   // channel offset of the A tile in 64-channel units (split mode: 0 = hi plane, Cin/64 = lo plane)
  int16_t tap_bk[kMaxVTaps];              // column offset of the tap's weight block in 64-column units (t * kc_per_tap when not split)
  int split;               // 1: split-operand mode -- activations are (hi, lo) fp16 pairs stored as 2*C channels per pixel
  void* out;               // raw conv output, NHWC [B, Ho, Wo, Cout]
  long long out_bs;        // elements per sample
  int out_rs, out_ps;      // elements per output row / pixel
  int oy_mul, ox_mul;      // GEMM-grid (y,x) -> output (y*oy_mul+oy_off, x*ox_mul+ox_off)
  int8_t oy_off[kMaxPhases], ox_off[kMaxPhases];
  int out_rows, out_y_off, out_x_off;   // output storage: rows per sample (in units of oy_mul rows) and tile offsets
  float* partials;         // [B][tiles_per_sample][3][Cout] (sum, sum of squares, max) or nullptr
  const float* bias;       // optional per-channel bias (epilogue mode 1)
  int epi_mode;            // 0: raw output + statistics, 1: bias + LeakyReLU(0.2), no statistics
  ConvFinalize fin;        // optional: InstanceNorm finalize (+ CBAM channel MLP) by the CTA that completes a sample
};

// Host-side description of one convolution as an implicit GEMM.
struct ConvPlan {
  const void* in;          // NHWC activation buffer (already padded as the conv needs)
  int B, Hp, Wp, Cin;      // its shape
  int stride;              // 1 or 2 (2: Hp and Wp must be even)
  const void* w;           // packed weights [num_phases*Cout][num_taps*Cin], K-major
  int Cout, num_phases, num_taps;
  int fold;                // 0/1 normal; 4: Cout = 4 * (output channels), column block f = output phase (f>>1, f&1)
  int8_t tap_dy[kMaxPhases][kMaxTaps], tap_dx[kMaxPhases][kMaxTaps];  // offsets in padded input pixels
  int Hg, Wg;              // GEMM grid per sample
  void* out;
  int Ho, Wo;              // output image
  int oy_mul, ox_mul;
  int8_t oy_off[kMaxPhases], ox_off[kMaxPhases];
  int out_y_off, out_x_off;  // the GEMM grid may be written at an offset inside a larger output image (Ho x Wo)
  float* partials;
  const float* bias;
  int epi_mode;
  int dtype;               // DUCOSY_F16 / DUCOSY_BF16, or DUCOSY_F16X2: split-operand mode (see ducosy.h)
  ConvFinalize fin;
};

int launch_conv_gemm(const ConvPlan& p, cudaStream_t stream);
inline int conv_tiles_per_sample(int num_phases, int Hg, int Wg) { return num_phases * (Hg * Wg / 128); }

template <typename T> struct Cvt;
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
};

#define DUCOSY_DISPATCH_DTYPE(dtype, T, ...)                         \
  do {                                                               \
    if ((dtype) == DUCOSY_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { using T = __half; __VA_ARGS__; }  /* DUCOSY_F16, DUCOSY_F16X2 */ \
  } while (0)

}  // namespace ducosy
