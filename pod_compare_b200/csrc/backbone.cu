// SIMT pieces of the ResNet-50-FPN backbone (detectron2 build_retinanet_resnet_fpn_backbone, call sites
// /root/reference/src/probabilistic_modeling/probabilistic_retinanet.py:96-101).  The 1x1 / 3x3 convolutions of the residual
// stages and of the FPN run on the tcgen05 kernel (conv_tc.cu, pod_conv_tc_general); what is left is here:
//   k_stem_conv7     preprocess (normalise, zero-pad) + 7x7 / 2 convolution 3 -> 64 + folded FrozenBN + ReLU   (K = 147: no
//                    tensor-core shape; 2 % of the backbone's FLOPs)
//   k_maxpool3s2     3x3 / 2 max-pool -> fp16 split pair (operand of res2)
//   k_upsample2_add  FPN top-down pathway: nearest x2 upsample + lateral add
//   k_split_f32      fp32 channels-last -> fp16 split pair (device-resident scale)
//   k_pack_w_k       weight packing for k x k kernels
#include "common.cuh"

namespace {

constexpr int ST = 16;                 // output tile 16 x 16 pixels, one pixel per thread
constexpr int SP = 2 * ST + 5;         // input patch edge: 37
constexpr int SK = 7 * 7 * 3;          // 147

template <typename PixT>
__global__ void __launch_bounds__(256) k_stem_conv7(const PixT* __restrict__ img, int Himg, int Wimg, int H, int W, int Hc, int Wc,
                                                    float3 mean, float3 inv_std, const float* __restrict__ w /* [147][64] */,
                                                    const float* __restrict__ bias, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem_stem[];
  float* ws = smem_stem;                                                   // [147][64] weights
  float (*patch)[SP][SP + 1] = reinterpret_cast<float (*)[SP][SP + 1]>(smem_stem + SK * 64);   // [3][37][38] input patch
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * ST, ox0 = blockIdx.x * ST;
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  const float mu[3] = {mean.x, mean.y, mean.z}, is[3] = {inv_std.x, inv_std.y, inv_std.z};
  for (int i = threadIdx.x; i < 3 * SP * SP; i += 256) {
    const int c = i / (SP * SP), r = (i / SP) % SP, q = i % SP;
    const int y = iy0 + r, x = ix0 + q;
    float v = 0.f;                     // convolution padding, and the zero padding of the NORMALISED image up to (H, W)
    if (y >= 0 && x >= 0 && y < Himg && x < Wimg) v = ((float)img[(((int64_t)n * 3 + c) * Himg + y) * Wimg + x] - mu[c]) * is[c];
    patch[c][r][q] = v;
  }
  for (int i = threadIdx.x; i < SK * 64; i += 256) ws[i] = w[i];
  __syncthreads();
  const int ty = threadIdx.x / ST, tx = threadIdx.x % ST;
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  for (int ky = 0; ky < 7; ++ky)
    for (int kx = 0; kx < 7; ++kx)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float x = patch[c][2 * ty + ky][2 * tx + kx];
        const float4* wr = reinterpret_cast<const float4*>(ws + ((ky * 7 + kx) * 3 + c) * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 wv = wr[j];
          acc[4 * j] = fmaf(x, wv.x, acc[4 * j]);
          acc[4 * j + 1] = fmaf(x, wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(x, wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(x, wv.w, acc[4 * j + 3]);
        }
      }
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy < Hc && ox < Wc) {
    float4* o = reinterpret_cast<float4*>(out + (((int64_t)n * Hc + oy) * Wc + ox) * 64);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      o[j] = make_float4(fmaxf(acc[4 * j] + bias[4 * j], 0.f), fmaxf(acc[4 * j + 1] + bias[4 * j + 1], 0.f),
                         fmaxf(acc[4 * j + 2] + bias[4 * j + 2], 0.f), fmaxf(acc[4 * j + 3] + bias[4 * j + 3], 0.f));
  }
  (void)H; (void)W;
}

// (NB, Hc, Wc, 64) fp32 -> 3x3 / stride 2 / pad 1 max -> (NB, Hp, Wp, 64) split pair; 8 channels per thread
__global__ void __launch_bounds__(256) k_maxpool3s2(const float* __restrict__ x, int Hc, int Wc, int Hp, int Wp, int64_t total, float scale,
                                                    __half* __restrict__ hi, __half* __restrict__ lo) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(t % 8);
    const int64_t p = t / 8;
    const int px = (int)(p % Wp), py = (int)((p / Wp) % Hp), n = (int)(p / ((int64_t)Wp * Hp));
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = 2 * py + dy;
      if (y < 0 || y >= Hc) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = 2 * px + dx;
        if (xx < 0 || xx >= Wc) continue;
        const float4* s = reinterpret_cast<const float4*>(x + (((int64_t)n * Hc + y) * Wc + xx) * 64 + c8 * 8);
        const float4 a = __ldg(s), b = __ldg(s + 1);
        m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], a.z); m[3] = fmaxf(m[3], a.w);
        m[4] = fmaxf(m[4], b.x); m[5] = fmaxf(m[5], b.y); m[6] = fmaxf(m[6], b.z); m[7] = fmaxf(m[7], b.w);
      }
    }
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half h0, l0, h1, l1;
      pod_split_h(m[2 * i] * scale, h0, l0);
      pod_split_h(m[2 * i + 1] * scale, h1, l1);
      ph[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      pl[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    reinterpret_cast<uint4*>(hi)[t] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    reinterpret_cast<uint4*>(lo)[t] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

__global__ void __launch_bounds__(256) k_upsample2_add(float* __restrict__ dst, const float* __restrict__ src, int H, int W, int Hs, int Ws,
                                                       int C4, int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C4);
    const int64_t p = t / C4;
    const int x = (int)(p % W), y = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
    const int ys = min(y >> 1, Hs - 1), xs = min(x >> 1, Ws - 1);       // F.interpolate(scale_factor=2, mode="nearest")
    float4 d = reinterpret_cast<float4*>(dst)[t];
    const float4 s = __ldg(reinterpret_cast<const float4*>(src) + (((int64_t)n * Hs + ys) * Ws + xs) * C4 + c);
    d.x = __fadd_rn(d.x, s.x); d.y = __fadd_rn(d.y, s.y); d.z = __fadd_rn(d.z, s.z); d.w = __fadd_rn(d.w, s.w);
    reinterpret_cast<float4*>(dst)[t] = d;
  }
}

__global__ void __launch_bounds__(256) k_split_f32(const float* __restrict__ x, int64_t total8, float scale, const float* __restrict__ scale_dev,
                                                   int relu, __half* __restrict__ hi, __half* __restrict__ lo, int* status) {
  if (scale_dev != nullptr) scale = __ldg(scale_dev);
  bool saturated = false;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total8; t += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * t), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * t + 1);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (relu) v[i] = fmaxf(v[i], 0.f);
      v[i] *= scale;
      saturated |= !(fabsf(v[i]) <= 65504.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half h0, l0, h1, l1;
      pod_split_h(v[2 * i], h0, l0);
      pod_split_h(v[2 * i + 1], h1, l1);
      ph[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      pl[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    reinterpret_cast<uint4*>(hi)[t] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    reinterpret_cast<uint4*>(lo)[t] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
  if (saturated) atomicCAS(status, 0, 103);
}

__global__ void k_pack_w_k(const float* __restrict__ w, int Cout, int Cin, int taps, int Cout_pad, float scale, __half* __restrict__ hi,
                           __half* __restrict__ lo) {
  const int64_t total = (int64_t)Cout_pad * taps * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i / ((int64_t)taps * Cin));
    const int k = (int)(i % ((int64_t)taps * Cin));
    const int tap = k / Cin, ci = k % Cin;
    float v = 0.f;
    if (co < Cout) v = w[((int64_t)co * Cin + ci) * taps + tap];
    __half h, l;
    pod_split_h(v * scale, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

__device__ int g_bb_status = 0;      // 103 = an FPN map / backbone activation left the fp16 split range in k_split_f32
}  // namespace

int pod_backbone_status_fetch(int* v) {
  *v = 0;
  POD_CUDA(cudaMemcpyFromSymbol(v, g_bb_status, sizeof(int)));
  if (*v != 0) {
    int zero = 0;
    POD_CUDA(cudaMemcpyToSymbol(g_bb_status, &zero, sizeof(int)));
  }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_pack_conv_weight_k(const float* w, int Cout, int Cin, int ksize, int Cout_pad, float scale,
                                                               void* dst_hi, void* dst_lo, void* stream) {
  POD_REQUIRE(w && dst_hi && dst_lo && Cout > 0 && Cin > 0 && Cout_pad >= Cout && (ksize == 1 || ksize == 3), "pod_pack_conv_weight_k: bad args");
  k_pack_w_k<<<512, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, ksize * ksize, Cout_pad, scale, (__half*)dst_hi, (__half*)dst_lo);
  POD_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_stem_conv7_pool(const void* images, int is_u8, int NB, int Himg, int Wimg, int H, int W,
                                                            const float* mean3_host, const float* std3_host, const float* w147x64,
                                                            const float* bias, float* scratch, void* out_hi, void* out_lo,
                                                            float out_scale, void* stream) {
  POD_REQUIRE(images && mean3_host && std3_host && w147x64 && bias && scratch && out_hi && out_lo, "pod_stem_conv7_pool: null argument");
  POD_REQUIRE(NB > 0 && NB <= 65535 && Himg > 0 && Wimg > 0 && H >= Himg && W >= Wimg && out_scale > 0.f, "pod_stem_conv7_pool: bad shape");
  const int Hc = (H + 1) / 2, Wc = (W + 1) / 2, Hp = (Hc + 1) / 2, Wp = (Wc + 1) / 2;
  const float3 mean = make_float3(mean3_host[0], mean3_host[1], mean3_host[2]);
  const float3 inv = make_float3(1.f / std3_host[0], 1.f / std3_host[1], 1.f / std3_host[2]);
  dim3 grid((Wc + ST - 1) / ST, (Hc + ST - 1) / ST, NB);
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int SMEM = (SK * 64 + 3 * SP * (SP + 1)) * 4;
  static bool configured = false;
  if (!configured) {
    POD_CUDA(cudaFuncSetAttribute(k_stem_conv7<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    POD_CUDA(cudaFuncSetAttribute(k_stem_conv7<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  if (is_u8)
    k_stem_conv7<uint8_t><<<grid, 256, SMEM, st>>>((const uint8_t*)images, Himg, Wimg, H, W, Hc, Wc, mean, inv, w147x64, bias, scratch);
  else
    k_stem_conv7<float><<<grid, 256, SMEM, st>>>((const float*)images, Himg, Wimg, H, W, Hc, Wc, mean, inv, w147x64, bias, scratch);
  POD_LAUNCH_CHECK();
  const int64_t total = (int64_t)NB * Hp * Wp * 8;
  const int64_t want = (total + 255) / 256, cap = (int64_t)pod_num_sms() * 8;
  k_maxpool3s2<<<(int)(want < cap ? want : cap), 256, 0, st>>>(scratch, Hc, Wc, Hp, Wp, total, out_scale, (__half*)out_hi, (__half*)out_lo);
  POD_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_upsample2_add(float* dst, const float* src, int NB, int H, int W, int C, void* stream) {
  POD_REQUIRE(dst && src && NB > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "pod_upsample2_add: bad args (C%%4)");
  const int Hs = (H + 1) / 2, Ws = (W + 1) / 2;
  const int64_t total = (int64_t)NB * H * W * (C / 4);
  const int64_t want = (total + 255) / 256, cap = (int64_t)pod_num_sms() * 8;
  k_upsample2_add<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(dst, src, H, W, Hs, Ws, C / 4, total);
  POD_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_split_f32(const float* src, int64_t n, float scale, const float* scale_dev, int relu,
                                                      void* dst_hi, void* dst_lo, void* stream) {
  POD_REQUIRE(src && dst_hi && dst_lo && n > 0 && n % 8 == 0 && (scale > 0.f || scale_dev), "pod_split_f32: bad args (n%%8)");
  POD_REQUIRE(((uintptr_t)src | (uintptr_t)dst_hi | (uintptr_t)dst_lo) % 16 == 0, "pod_split_f32: buffers must be 16-byte aligned");
  int* status = nullptr;
  POD_CUDA(cudaGetSymbolAddress((void**)&status, g_bb_status));
  const int64_t total8 = n / 8, want = (total8 + 255) / 256, cap = (int64_t)pod_num_sms() * 8;
  k_split_f32<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(src, total8, scale, scale_dev, relu, (__half*)dst_hi,
                                                                               (__half*)dst_lo, status);
  POD_LAUNCH_CHECK();
  return 0;
}
