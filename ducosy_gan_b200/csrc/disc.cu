// PatchGAN discriminator forward (modules/model.py:118-131):
//   Conv(1,64,4,s2,p1)+LeakyReLU(0.2) | Conv(64,128)+IN+LReLU | Conv(128,256)+IN+LReLU | Conv(256,512)+IN+LReLU |
//   ZeroPad2d((1,0,1,0)) + Conv(512,1,4,p1)
// The three middle convolutions (K = 1024 / 2048 / 4096) run on the tcgen05 implicit-GEMM kernel (conv_gemm.cu); the
// first layer (Cin = 1, 16 MACs per output) and the last (Cout = 1, a 8192-long dot product per patch) are
// bandwidth-bound and get their own small kernels.  Conv biases in front of the non-affine InstanceNorm cancel.
#include <algorithm>

#include "common.cuh"

namespace ducosy {
namespace {

// layer 1: x fp32 [B][1][H][W] -> LeakyReLU(conv4x4 s2 p1 + bias), written as zero-padded NHWC [B][H/2+2][W/2+2][64].
// One thread = one output pixel x 8 channels (16-byte store); border pixels of the padded map are written as zeros.
template <typename T>
__global__ void disc_first_conv_kernel(const float* __restrict__ x, const float* __restrict__ w /*[64][16]*/,
                                       const float* __restrict__ bias, T* __restrict__ out, int B, int H, int W) {
  pdl_prologue();
  __shared__ __align__(16) float sw[16 * 64];  // [tap][channel]: a thread's 8 channels are two conflict-free float4 reads
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) sw[(i & 15) * 64 + (i >> 4)] = w[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2, Hp = Ho + 2, Wp = Wo + 2;
  const long long total = (long long)B * Hp * Wp * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = int(i & 7);
    long long r = i >> 3;
    const int px = int(r % Wp);
    r /= Wp;
    const int py = int(r % Hp);
    const int b = int(r / Hp);
    uint4 o = make_uint4(0, 0, 0, 0);
    const int oy = py - 1, ox = px - 1;
    if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) {
      float in[16];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {
          const int iy = 2 * oy + rr - 1, ix = 2 * ox + ss - 1;
          in[rr * 4 + ss] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[((long long)b * H + iy) * W + ix] : 0.f;
        }
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = sb[c8 * 8 + j];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(&sw[k * 64 + c8 * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&sw[k * 64 + c8 * 8 + 4]);
        v[0] = fmaf(in[k], w0.x, v[0]); v[1] = fmaf(in[k], w0.y, v[1]); v[2] = fmaf(in[k], w0.z, v[2]); v[3] = fmaf(in[k], w0.w, v[3]);
        v[4] = fmaf(in[k], w1.x, v[4]); v[5] = fmaf(in[k], w1.y, v[5]); v[6] = fmaf(in[k], w1.z, v[6]); v[7] = fmaf(in[k], w1.w, v[7]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
      o = make_uint4(Cvt<T>::pack2(v[0], v[1]), Cvt<T>::pack2(v[2], v[3]), Cvt<T>::pack2(v[4], v[5]), Cvt<T>::pack2(v[6], v[7]));
    }
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// last layer: in = LeakyReLU(IN(.)) padded by 2 [B][Hs+4][Ws+4][512]; out[b][y][x] = bias + sum_{r,s,c} in[y+r][x+s][c] w[c][r][s]
// (ZeroPad2d((1,0,1,0)) + padding 1 = 2 zeros left/top, 1 right/bottom).  One warp per output value.
template <typename T>
__global__ void disc_last_conv_kernel(const T* __restrict__ in, const T* __restrict__ wp /*[16][512]*/,
                                      const float* __restrict__ bias, float* __restrict__ out, int B, int Hs, int Ws) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int Wp = Ws + 4, Hp = Hs + 4;
  const long long total = (long long)B * Hs * Ws;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long o = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); o < total; o += wstride) {
    const int x = int(o % Ws);
    long long r = o / Ws;
    const int y = int(r % Hs);
    const int b = int(r / Hs);
    float acc = 0.f;
#pragma unroll 4
    for (int tap = 0; tap < 16; ++tap) {
      const T* src = in + (((long long)b * Hp + y + (tap >> 2)) * Wp + x + (tap & 3)) * 512;
      const T* wt = wp + tap * 512;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint4 a = *reinterpret_cast<const uint4*>(src + h * 256 + lane * 8);
        const uint4 k = *reinterpret_cast<const uint4*>(wt + h * 256 + lane * 8);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, kw[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fa = Cvt<T>::unpack2(aw[j]), fk = Cvt<T>::unpack2(kw[j]);
          acc = fmaf(fa.x, fk.x, acc);
          acc = fmaf(fa.y, fk.y, acc);
        }
      }
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) out[o] = acc + __ldg(bias);
  }
}

template <typename T>
__global__ void pack_disc_last_weight_kernel(const float* __restrict__ w /*[1][512][4][4]*/, T* __restrict__ out) {
  pdl_prologue();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 16 * 512; i += gridDim.x * blockDim.x) {
    const int c = i & 511, tap = i >> 9;
    out[i] = Cvt<T>::from_f(w[c * 16 + tap]);
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DiscLayout {
  size_t w1, b1, w2, w3, w4, w5, b5, wd2, wd3, wd4, total;   // wd*: transposed packings for the input gradients
};
DiscLayout make_disc_layout() {
  DiscLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.w1 = take(64 * 16 * 4);
  L.b1 = take(64 * 4);
  L.w2 = take(size_t(128) * 16 * 64 * 2);
  L.w3 = take(size_t(256) * 16 * 128 * 2);
  L.w4 = take(size_t(512) * 16 * 256 * 2);
  L.w5 = take(16 * 512 * 2);
  L.b5 = take(4);
  L.wd2 = take(size_t(16) * 128 * 64 * 2);
  L.wd3 = take(size_t(16) * 256 * 128 * 2);
  L.wd4 = take(size_t(16) * 512 * 256 * 2);
  L.total = off;
  return L;
}
struct DiscWorkspace {
  size_t p1, y2, p2, y3, p3, y4, p4, partials, scale[3], shift[3], total;   // nothing aliased: backward reads it all
};
DiscWorkspace make_disc_workspace(int B, int H, int W) {
  DiscWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 1024); return o; };
  const int H1 = H / 2, W1 = W / 2, H2 = H / 4, W2 = W / 4, H3 = H / 8, W3 = W / 8, H4 = H / 16, W4 = W / 16;
  w.p1 = take(size_t(B) * (H1 + 2) * (W1 + 2) * 64 * 2);
  w.y2 = take(size_t(B) * H2 * W2 * 128 * 2);
  w.p2 = take(size_t(B) * (H2 + 2) * (W2 + 2) * 128 * 2);
  w.y3 = take(size_t(B) * H3 * W3 * 256 * 2);
  w.p3 = take(size_t(B) * (H3 + 2) * (W3 + 2) * 256 * 2);
  w.y4 = take(size_t(B) * H4 * W4 * 512 * 2);
  w.p4 = take(size_t(B) * (H4 + 4) * (W4 + 4) * 512 * 2);
  w.partials = take(size_t(B) * (size_t(H2) * W2 / 128) * 3 * 128 * 4);
  for (int l = 0; l < 3; ++l) {
    w.scale[l] = take(size_t(B) * 512 * 4);
    w.shift[l] = take(size_t(B) * 512 * 4);
  }
  w.total = off;
  return w;
}
int check_disc_shape(int B, int H, int W) {
  DUCOSY_CHECK(B >= 1 && H >= 256 && W >= 256 && H % 256 == 0 && W % 256 == 0, DUCOSY_ERR_SHAPE,
               "discriminator: H and W must be multiples of 256 (got %dx%d)", H, W);
  return 0;
}

}  // namespace
}  // namespace ducosy

using namespace ducosy;

extern "C" size_t ducosy_discriminator_packed_bytes(void) { return make_disc_layout().total; }
extern "C" size_t ducosy_discriminator_workspace_bytes(int B, int H, int W) {
  return check_disc_shape(B, H, W) == 0 ? make_disc_workspace(B, H, W).total : 0;
}

// params: host array of the 10 DEVICE fp32 tensors model.{0,2,5,8,12}.{weight,bias} in state_dict order.
extern "C" int ducosy_discriminator_pack(const float* const* params, int num_params, void* packed, int dtype,
                                         ducosy_stream_t stream) {
  DUCOSY_CHECK(params && packed && num_params == 10, DUCOSY_ERR_ARG, "discriminator_pack: expected 10 parameter tensors");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "discriminator_pack: bad dtype");
  for (int i = 0; i < 10; ++i) DUCOSY_CHECK(params[i] != nullptr, DUCOSY_ERR_ARG, "discriminator_pack: parameter %d is null", i);
  const DiscLayout L = make_disc_layout();
  uint8_t* pk = static_cast<uint8_t*>(packed);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaMemcpyAsync(pk + L.w1, params[0], 64 * 16 * 4, cudaMemcpyDeviceToDevice, st);   // [64][1][4][4] is already [64][16]
  cudaMemcpyAsync(pk + L.b1, params[1], 64 * 4, cudaMemcpyDeviceToDevice, st);
  DUCOSY_TRY(ducosy_pack_conv_weight(params[2], pk + L.w2, 128, 64, 4, 4, dtype, stream));
  DUCOSY_TRY(ducosy_pack_conv_weight(params[4], pk + L.w3, 256, 128, 4, 4, dtype, stream));
  DUCOSY_TRY(ducosy_pack_conv_weight(params[6], pk + L.w4, 512, 256, 4, 4, dtype, stream));
  DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(pack_disc_last_weight_kernel<T>, 32, 256, 0, st)(params[8], reinterpret_cast<T*>(pk + L.w5))));
  cudaMemcpyAsync(pk + L.b5, params[9], 4, cudaMemcpyDeviceToDevice, st);
  DUCOSY_TRY(ducosy_pack_dgrad_s2_weight(params[2], pk + L.wd2, 128, 64, 4, dtype, stream));
  DUCOSY_TRY(ducosy_pack_dgrad_s2_weight(params[4], pk + L.wd3, 256, 128, 4, dtype, stream));
  DUCOSY_TRY(ducosy_pack_dgrad_s2_weight(params[6], pk + L.wd4, 512, 256, 4, dtype, stream));
  return check_launch("discriminator_pack");
}

// Discriminator.forward (modules/model.py:130-131): x fp32 [B][1][H][W] -> out fp32 [B][1][H/16][W/16].
extern "C" int ducosy_discriminator_forward(const void* packed, const float* x, float* out, int B, int H, int W,
                                            void* workspace, size_t workspace_bytes, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(packed && x && out && workspace, DUCOSY_ERR_ARG, "discriminator_forward: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "discriminator_forward: bad dtype");
  DUCOSY_TRY(ducosy_check_device());
  DUCOSY_TRY(check_disc_shape(B, H, W));
  const DiscLayout L = make_disc_layout();
  const DiscWorkspace w = make_disc_workspace(B, H, W);
  DUCOSY_CHECK(workspace_bytes >= w.total, DUCOSY_ERR_WORKSPACE, "discriminator_forward: workspace %zu < required %zu bytes",
               workspace_bytes, w.total);
  DUCOSY_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0 && (reinterpret_cast<uintptr_t>(packed) & 255) == 0,
               DUCOSY_ERR_ALIGN, "discriminator_forward: workspace must be 1024-byte and packed weights 256-byte aligned");
  uint8_t* base = static_cast<uint8_t*>(workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partials = reinterpret_cast<float*>(base + w.partials);
  const int H1 = H / 2, W1 = W / 2;
  {
    const long long total = (long long)B * (H1 + 2) * (W1 + 2) * 8;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)(num_sms() > 0 ? num_sms() : 148) * 8;
    if (blocks > cap) blocks = cap;
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(disc_first_conv_kernel<T>, int(blocks), 256, 0, st)(
                                        x, reinterpret_cast<const float*>(pk + L.w1), reinterpret_cast<const float*>(pk + L.b1),
                                        reinterpret_cast<T*>(base + w.p1), B, H, W)));
    DUCOSY_TRY(check_launch("disc_first_conv_kernel"));
  }
  // three tensor-core stages: conv4x4 s2 (zero-padded input) -> IN statistics -> IN apply + LeakyReLU + zero pad
  const size_t in_off[3] = {w.p1, w.p2, w.p3}, y_off[3] = {w.y2, w.y3, w.y4}, p_off[3] = {w.p2, w.p3, w.p4};
  const size_t w_off[3] = {L.w2, L.w3, L.w4};
  int Hi = H1, Wi = W1, Ci = 64;
  for (int l = 0; l < 3; ++l) {
    const int Ho = Hi / 2, Wo = Wi / 2, Co = Ci * 2;
    float* scale = reinterpret_cast<float*>(base + w.scale[l]);
    float* shift = reinterpret_cast<float*>(base + w.shift[l]);
    DUCOSY_TRY(ducosy_conv2d_nhwc(base + in_off[l], pk + w_off[l], base + y_off[l], partials, nullptr, DUCOSY_ACT_NONE, B,
                                  Hi + 2, Wi + 2, Ci, Co, 4, 4, 2, dtype, stream));
    DUCOSY_TRY(ducosy_in_finalize(partials, Ho * Wo / 128, Ho * Wo, scale, shift, nullptr, nullptr, nullptr, B, Co, stream));
    DUCOSY_TRY(ducosy_in_apply_pad(base + y_off[l], scale, shift, base + p_off[l], B, Ho, Wo, Co, l == 2 ? 2 : 1,
                                   DUCOSY_PAD_ZERO, DUCOSY_ACT_LRELU02, dtype, stream));
    Hi = Ho; Wi = Wo; Ci = Co;
  }
  {
    const long long outputs = (long long)B * Hi * Wi;
    long long blocks = (outputs + 7) / 8;
    const long long cap = (long long)(num_sms() > 0 ? num_sms() : 148) * 8;
    if (blocks > cap) blocks = cap;
    DUCOSY_DISPATCH_DTYPE(dtype, T, (pdl(disc_last_conv_kernel<T>, int(blocks), 256, 0, st)(
                                        reinterpret_cast<const T*>(base + w.p4), reinterpret_cast<const T*>(pk + L.w5),
                                        reinterpret_cast<const float*>(pk + L.b5), out, B, Hi, Wi)));
  }
  return check_launch("disc_last_conv_kernel");
}

// ---------------------------------------------------------------- backward
namespace ducosy {
namespace {
struct DiscBwdWorkspace {
  size_t da[4], dyp[3], in_scratch, wg_ws, dwp, first_scratch, last_scratch, gs, total;   // da[l]: grad wrt the activation after layer l+1
};
DiscBwdWorkspace make_disc_bwd_workspace(int B, int H, int W) {
  DiscBwdWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 1024); return o; };
  int Hl = H / 2, Wl = W / 2, Cl = 64;
  size_t in_max = 0, wg_max = 0;
  for (int l = 0; l < 4; ++l) {
    w.da[l] = take(size_t(B) * Hl * Wl * Cl * 2);
    if (l > 0) {
      w.dyp[l - 1] = take(size_t(B) * (Hl + 2) * (Wl + 2) * Cl * 2);
      in_max = std::max(in_max, ducosy_in_backward_scratch_bytes(B, Hl, Wl, Cl));
      wg_max = std::max(wg_max, ducosy_conv2d_wgrad_workspace_bytes(B, Hl, Wl, Cl / 2, Cl, 4, 4));
    }
    Hl /= 2; Wl /= 2; Cl *= 2;
  }
  w.in_scratch = take(in_max);
  w.wg_ws = take(wg_max);
  w.dwp = take(size_t(512) * 16 * 256 * 4);
  w.first_scratch = take(ducosy_disc_first_backward_scratch_bytes(B, H, W));
  w.last_scratch = take(ducosy_disc_last_backward_scratch_bytes());
  w.gs = take(16);
  w.total = off;
  return w;
}
}  // namespace
}  // namespace ducosy

extern "C" size_t ducosy_discriminator_backward_workspace_bytes(int B, int H, int W) {
  return check_disc_shape(B, H, W) == 0 ? make_disc_bwd_workspace(B, H, W).total : 0;
}

// Backward of ducosy_discriminator_forward.  fwd_workspace must be the (untouched) workspace of that forward call.
// grads_host: host array of 10 DEVICE fp32 buffers shaped like model.{0,2,5,8,12}.{weight,bias} (written, not accumulated);
// dx: optional gradient w.r.t. the input image [B][1][H][W] (NULL to skip).
extern "C" int ducosy_discriminator_backward(const void* packed, const float* x, const float* dout, const void* fwd_workspace,
                                             float* const* grads_host, float* dx, int B, int H, int W, void* workspace,
                                             size_t workspace_bytes, int dtype, ducosy_stream_t stream) {
  DUCOSY_CHECK(packed && x && dout && fwd_workspace && grads_host && workspace, DUCOSY_ERR_ARG, "discriminator_backward: null pointer");
  DUCOSY_CHECK(dtype == DUCOSY_F16 || dtype == DUCOSY_BF16, DUCOSY_ERR_ARG, "discriminator_backward: bad dtype");
  DUCOSY_TRY(check_disc_shape(B, H, W));
  for (int i = 0; i < 10; ++i) DUCOSY_CHECK(grads_host[i] != nullptr, DUCOSY_ERR_ARG, "discriminator_backward: gradient buffer %d is null", i);
  const DiscLayout L = make_disc_layout();
  const DiscWorkspace fw = make_disc_workspace(B, H, W);
  const DiscBwdWorkspace bw = make_disc_bwd_workspace(B, H, W);
  DUCOSY_CHECK(workspace_bytes >= bw.total, DUCOSY_ERR_WORKSPACE, "discriminator_backward: workspace %zu < required %zu bytes",
               workspace_bytes, bw.total);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  const uint8_t* fb = static_cast<const uint8_t*>(fwd_workspace);
  uint8_t* bb = static_cast<uint8_t*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int H4 = H / 16, W4 = W / 16;

  // power-of-two scale that lifts the (tiny) loss gradient out of the fp16 subnormal range; undone in fp32 at the end
  float* gs = reinterpret_cast<float*>(bb + bw.gs);
  DUCOSY_TRY(ducosy_grad_scale(dout, (long long)B * H4 * W4, gs, stream));
  // last layer: grad wrt a4 (the LeakyReLU(IN(.)) map) + its weight / bias gradients
  DUCOSY_TRY(ducosy_disc_last_backward(dout, pk + L.w5, fb + fw.p4, bb + bw.da[3], grads_host[8], grads_host[9],
                                       reinterpret_cast<float*>(bb + bw.last_scratch), gs, B, H4, W4, dtype, stream));

  const size_t y_off[3] = {fw.y2, fw.y3, fw.y4}, pin_off[3] = {fw.p1, fw.p2, fw.p3}, wd_off[3] = {L.wd2, L.wd3, L.wd4};
  int Hl = H4, Wl = W4, Cl = 512;
  for (int l = 2; l >= 0; --l) {   // layers 4, 3, 2 of the reference (index l: conv l+2)
    const float* scale = reinterpret_cast<const float*>(fb + fw.scale[l]);
    const float* shift = reinterpret_cast<const float*>(fb + fw.shift[l]);
    // InstanceNorm + LeakyReLU backward -> gradient of the raw conv output, zero-padded for the phase convs below
    DUCOSY_TRY(ducosy_in_backward_pad(bb + bw.da[l + 1], fb + y_off[l], scale, shift, bb + bw.dyp[l],
                                      reinterpret_cast<float*>(bb + bw.in_scratch), B, Hl, Wl, Cl, 1, DUCOSY_ACT_LRELU02, dtype, stream));
    // weight gradient (tensor cores) -> OIHW; the conv bias in front of a non-affine InstanceNorm has zero gradient
    DUCOSY_TRY(ducosy_conv2d_wgrad_nhwc(fb + pin_off[l], bb + bw.dyp[l], 1, reinterpret_cast<float*>(bb + bw.dwp), B, 2 * Hl + 2,
                                        2 * Wl + 2, Cl / 2, Cl, 4, 4, 2, bb + bw.wg_ws, bw.dwp - bw.wg_ws, dtype, stream));
    DUCOSY_TRY(ducosy_unpack_wgrad(reinterpret_cast<const float*>(bb + bw.dwp), grads_host[2 + 2 * l], Cl, Cl / 2, 16, gs, stream));
    cudaMemsetAsync(grads_host[3 + 2 * l], 0, size_t(Cl) * 4, st);
    // input gradient (tensor cores): grad wrt the previous activation map
    DUCOSY_TRY(ducosy_convs2_dgrad_nhwc(bb + bw.dyp[l], pk + wd_off[l], bb + bw.da[l], B, Hl, Wl, Cl / 2, Cl, dtype, stream));
    Hl *= 2; Wl *= 2; Cl /= 2;
  }
  // first layer
  return ducosy_disc_first_backward(bb + bw.da[0], fb + fw.p1, x, reinterpret_cast<const float*>(pk + L.w1), grads_host[0],
                                    grads_host[1], dx, reinterpret_cast<float*>(bb + bw.first_scratch), gs, B, H, W, dtype, stream);
}
