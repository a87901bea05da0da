// Plain fp32 SIMT 3x3 convolution (channels-last).  Cross-check for the tcgen05 kernel in the GPU
// tests; not used by the product path.
#include "common.cuh"

namespace {
constexpr int TP = 64;   // pixels per block tile (8x8)
constexpr int TC = 64;   // output channels per block tile
constexpr int TK = 16;   // reduction chunk

__global__ void __launch_bounds__(256)
k_conv3x3_simt(const float* __restrict__ in, int H, int W, int Cin, const float* __restrict__ wk,
               const float* __restrict__ bias, int Cout, int Cout_pad, int relu, pod_dropout d, uint32_t thr,
               float dscale, PhiloxKey key, float* __restrict__ out, int64_t out_map_stride, int64_t out_pixel_stride) {
  __shared__ float As[TK][TP + 1];
  __shared__ __align__(16) float Bs[TK][TC];
  const int tiles_x = (W + 7) / 8;
  const int ty0 = (blockIdx.x / tiles_x) * 8, tx0 = (blockIdx.x % tiles_x) * 8;
  const int co0 = blockIdx.y * TC;
  const int n = blockIdx.z;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const float* inm = in + (int64_t)n * H * W * Cin;
  float acc[4][4] = {};
  const int lp = tid / 4, lc = (tid % 4) * 4;          // A-load role: pixel, channel offset
  const int bk = tid / 16, bc = (tid % 16) * 4;        // B-load role
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int py = ty0 + lp / 8 + dy, px = tx0 + lp % 8 + dx;
    const bool inb = py >= 0 && py < H && px >= 0 && px < W;
    for (int c0 = 0; c0 < Cin; c0 += TK) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inb) a = *reinterpret_cast<const float4*>(inm + ((int64_t)py * W + px) * Cin + c0 + lc);
      As[lc + 0][lp] = a.x; As[lc + 1][lp] = a.y; As[lc + 2][lp] = a.z; As[lc + 3][lp] = a.w;
      *reinterpret_cast<float4*>(&Bs[bk][bc]) =
          *reinterpret_cast<const float4*>(wk + ((int64_t)(tap * Cin + c0 + bk)) * Cout_pad + co0 + bc);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  const int reps = d.samples * d.passes;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int lpix = ty * 4 + i;
    const int py = ty0 + lpix / 8, px = tx0 + lpix % 8;
    if (py >= H || px >= W) continue;
    const int pixel = py * W + px;
    const int c = co0 + tx * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j] + bias[c + j];
      if (relu) v[j] = fmaxf(v[j], 0.f);
    }
    if (d.p > 0.0) {
      const int image = d.image0 + n / reps, sample = (n / d.passes) % d.samples, pass = d.pass0 + n % d.passes;
      // 16-bit lanes (common.cuh): one call covers 8 channels; this thread's 4 are its low or high half
      const int64_t e = (int64_t)pixel * Cout_pad + c;
      const uint32_t kb = pod_keep8(philox4x32_10((uint32_t)(e >> 3), pod_dropout_c1(d.level, d.layer, d.tower, pass),
                                                  (uint32_t)sample, (uint32_t)image, key), thr) >> ((e & 4) ? 4 : 0);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ((kb >> j) & 1u) ? v[j] * dscale : 0.f;
    }
    float* o = out + (int64_t)n * out_map_stride + (int64_t)pixel * out_pixel_stride + c;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < Cout) o[j] = v[j];
  }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_simt(const float* in, int NB, int H, int W, int Cin, const float* w_kc, const float* bias,
                                int Cout, int Cout_pad, int relu, const pod_dropout* drop, float* out,
                                int64_t out_map_stride, int64_t out_pixel_stride, void* stream) {
  POD_REQUIRE(in && w_kc && bias && out && NB > 0 && NB <= 65535 && H > 0 && W > 0, "pod_conv3x3_simt: bad args");
  POD_REQUIRE(Cin % TK == 0 && Cout_pad % TC == 0 && Cout <= Cout_pad, "pod_conv3x3_simt: Cin%%16 / Cout_pad%%64 required");
  pod_dropout d = {};
  d.samples = 1;
  d.passes = 1;
  if (drop) d = *drop;
  POD_REQUIRE(d.samples > 0 && d.passes > 0, "pod_conv3x3_simt: bad dropout spec");
  POD_REQUIRE(d.p == 0.0 || Cout == Cout_pad, "pod_conv3x3_simt: dropout needs Cout == Cout_pad");
  dim3 grid(((H + 7) / 8) * ((W + 7) / 8), Cout_pad / TC, NB);
  k_conv3x3_simt<<<grid, 256, 0, (cudaStream_t)stream>>>(in, H, W, Cin, w_kc, bias, Cout, Cout_pad, relu, d,
                                                         pod_dropout_threshold16(d.p), pod_dropout_scale(d.p > 0 ? d.p : 0.5),
                                                         pod_key(d.seed, POD_STREAM_DROPOUT), out, out_map_stride,
                                                         out_pixel_stride);
  POD_LAUNCH_CHECK();
  return 0;
}
