// Multi-tensor Adam with a non-finite guard: one optimizer.step() of torch.optim.Adam as the reference constructs it
// (modules/trainer.py:360-362; no weight decay, no amsgrad) over ALL tensors of a parameter group in three launches
// instead of one launch per tensor (the generator pair has 162 tensors; at one sample per rank those launches were a
// measurable part of the step).
//
//   adam_check_kernel   : OR "a gradient is Inf/NaN" into state[2]                      (skippable)
//   adam_advance_kernel2: single thread -- moves the flag to state[4] and clears it; advances state[1] (the step
//                         count) on a clean step, state[3] (skipped steps) otherwise
//   adam_multi_kernel   : the update, chunk by chunk; every block returns at once when state[4] is set
//
// Why the guard exists: the training path keeps gradient maps in 16 bit with one power-of-two scale per backward
// (ducosy_grad_scale).  InstanceNorm's backward multiplies by 1/sqrt(var + 1e-5) (up to 316 for a constant channel),
// so a pathological batch can push an fp16 map past 65504; the reference's fp32 autograd cannot overflow there.  A
// poisoned step is skipped GradScaler-style -- weights, moments and the step count stay untouched -- instead of
// writing Inf/NaN into the master weights for good.
#include <algorithm>

#include "common.cuh"

namespace ducosy {
namespace {

constexpr int kAdamThreads = 256;

__device__ __forceinline__ bool finite_f(float x) { return (__float_as_uint(x) & 0x7f800000u) != 0x7f800000u; }

__global__ void __launch_bounds__(kAdamThreads)
adam_check_kernel(const ducosy_adam_tensor* __restrict__ tensors, const ducosy_adam_chunk* __restrict__ chunks,
                  float* __restrict__ state) {
  pdl_prologue();
  const ducosy_adam_chunk c = chunks[blockIdx.x];
  const ducosy_adam_tensor t = tensors[c.tensor];
  const long long end = c.start + DUCOSY_ADAM_CHUNK < t.n ? c.start + DUCOSY_ADAM_CHUNK : t.n;
  const float* g = t.grad;
  bool bad = false;
  if ((reinterpret_cast<uintptr_t>(g + c.start) & 15) == 0) {
    long long i = c.start + threadIdx.x * 4;
    for (; i + 3 < end; i += kAdamThreads * 4) {
      const float4 q = *reinterpret_cast<const float4*>(g + i);
      bad |= !(finite_f(q.x) && finite_f(q.y) && finite_f(q.z) && finite_f(q.w));
    }
    for (; i < end; ++i) bad |= !finite_f(g[i]);   // at most 3 elements, only in the thread that owns the tail
  } else {
    for (long long i = c.start + threadIdx.x; i < end; i += kAdamThreads) bad |= !finite_f(g[i]);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) state[2] = 1.f;   // benign race: every writer stores the same value
}

__global__ void adam_advance_kernel2(float* __restrict__ state) {
  pdl_prologue();
  const float bad = state[2];
  state[4] = bad;
  state[2] = 0.f;
  if (bad != 0.f) state[3] += 1.f; else state[1] += 1.f;
}

__global__ void __launch_bounds__(kAdamThreads)
adam_multi_kernel(const ducosy_adam_tensor* __restrict__ tensors, const ducosy_adam_chunk* __restrict__ chunks,
                  const float* __restrict__ state, float b1, float b2, float eps) {
  pdl_prologue();
  __shared__ float s_step_size, s_inv_sqrt_bc2;
  if (state[4] != 0.f) return;                     // a gradient of this step is not finite: leave everything alone
  if (threadIdx.x == 0) {
    const double t = double(state[1]);
    const double bc1 = 1.0 - pow(double(b1), t), bc2 = 1.0 - pow(double(b2), t);
    s_step_size = float(double(state[0]) / bc1);
    s_inv_sqrt_bc2 = float(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_step_size, inv_sqrt_bc2 = s_inv_sqrt_bc2;
  const ducosy_adam_chunk c = chunks[blockIdx.x];
  const ducosy_adam_tensor t = tensors[c.tensor];
  const long long end = c.start + DUCOSY_ADAM_CHUNK < t.n ? c.start + DUCOSY_ADAM_CHUNK : t.n;
  float* __restrict__ p = t.param;
  const float* __restrict__ g = t.grad;
  float* __restrict__ m = t.exp_avg;
  float* __restrict__ v = t.exp_avg_sq;
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  };
  const uintptr_t al = reinterpret_cast<uintptr_t>(p + c.start) | reinterpret_cast<uintptr_t>(g + c.start) |
                       reinterpret_cast<uintptr_t>(m + c.start) | reinterpret_cast<uintptr_t>(v + c.start);
  if ((al & 15) == 0) {
    long long i = c.start + threadIdx.x * 4;
    for (; i + 3 < end; i += kAdamThreads * 4) {
      float4 pq = *reinterpret_cast<float4*>(p + i), mq = *reinterpret_cast<float4*>(m + i), vq = *reinterpret_cast<float4*>(v + i);
      const float4 gq = *reinterpret_cast<const float4*>(g + i);
      upd(pq.x, gq.x, mq.x, vq.x);
      upd(pq.y, gq.y, mq.y, vq.y);
      upd(pq.z, gq.z, mq.z, vq.z);
      upd(pq.w, gq.w, mq.w, vq.w);
      *reinterpret_cast<float4*>(p + i) = pq;
      *reinterpret_cast<float4*>(m + i) = mq;
      *reinterpret_cast<float4*>(v + i) = vq;
    }
    for (; i < end; ++i) upd(p[i], g[i], m[i], v[i]);
  } else {
    for (long long i = c.start + threadIdx.x; i < end; i += kAdamThreads) upd(p[i], g[i], m[i], v[i]);
  }
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" int ducosy_adam_multi_step(const ducosy_adam_tensor* tensors_dev, const ducosy_adam_chunk* chunks_dev,
                                      int num_chunks, float* state, float beta1, float beta2, float eps, int check_finite,
                                      ducosy_stream_t stream) {
  DUCOSY_CHECK(tensors_dev && chunks_dev && state && num_chunks > 0, DUCOSY_ERR_ARG, "adam_multi_step: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (check_finite) {
    pdl(adam_check_kernel, num_chunks, kAdamThreads, 0, st)(tensors_dev, chunks_dev, state);
    DUCOSY_TRY(check_launch("adam_check_kernel"));
  }
  pdl(adam_advance_kernel2, 1, 1, 0, st)(state);
  DUCOSY_TRY(check_launch("adam_advance_kernel"));
  pdl(adam_multi_kernel, num_chunks, kAdamThreads, 0, st)(tensors_dev, chunks_dev, state, beta1, beta2, eps);
  return check_launch("adam_multi_kernel");
}
