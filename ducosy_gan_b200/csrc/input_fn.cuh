// Input functors of the stem (shared by the im2col path and the fused stem kernels) and reflection indexing.
#pragma once
#include "common.cuh"

namespace ducosy {

__host__ __device__ __forceinline__ int reflect_idx(int i, int n) {  // ReflectionPad2d semantics (no edge repeat)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}


// in[b][c][y][x] of the network input: fp32 NCHW tensor, or stored int16 pixels through the HU window.
struct InF32 {
  const float* x;
  __device__ __forceinline__ float at(int b, int c, int y, int x_, int Cin, int H, int W) const {
    return x[(((long long)b * Cin + c) * H + y) * W + x_];
  }
};
struct InHU {
  const int16_t* px;
  float slope, intercept, lo, hi, span;
  __device__ __forceinline__ float at(int b, int, int y, int x_, int, int H, int W) const {
    const float hu = __fadd_rn(__fmul_rn(float(px[((long long)b * H + y) * W + x_]), slope), intercept);
    const float c = fminf(fmaxf(hu, lo), hi);
    return __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(c, lo)), span), 1.0f);  // preprocess.py:79-84
  }
};

}  // namespace ducosy
