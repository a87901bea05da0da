// Operand preparation and streaming statistics kernels (HBM-bound elementwise / transpose work).
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC (fp16 split pair | fp32): 32x32 shared-memory transpose tiles
// ---------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void k_nchw_to_nhwc(const float* __restrict__ src, int C, int HW, float scale, __half* __restrict__ hi,
                               __half* __restrict__ lo, float* __restrict__ dst32) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (int64_t)n * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? s[(int64_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) {
      const float v = tile[threadIdx.x][i];
      const int64_t o = ((int64_t)n * HW + p) * C + c;
      if (SPLIT) {
        __half h, l;
        pod_split_h(v * scale, h, l);
        hi[o] = h;
        lo[o] = l;
      } else {
        dst32[o] = v;
      }
    }
  }
}

extern "C" __attribute__((visibility("default"))) int pod_nchw_to_nhwc_split(const float* src, int NB, int C, int H, int W, float scale, void* dst_hi,
                                      void* dst_lo, void* stream) {
  POD_REQUIRE(src && dst_hi && dst_lo && NB > 0 && C > 0 && H > 0 && W > 0, "pod_nchw_to_nhwc_split: bad args");
  POD_REQUIRE(NB <= 65535, "pod_nchw_to_nhwc_split: NB too large for one launch");
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, NB), block(32, 8);
  k_nchw_to_nhwc<true><<<grid, block, 0, (cudaStream_t)stream>>>(src, C, H * W, scale, (__half*)dst_hi, (__half*)dst_lo,
                                                                 nullptr);
  POD_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_nchw_to_nhwc_f32(const float* src, int NB, int C, int H, int W, float* dst, void* stream) {
  POD_REQUIRE(src && dst && NB > 0 && NB <= 65535 && C > 0 && H > 0 && W > 0, "pod_nchw_to_nhwc_f32: bad args");
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, NB), block(32, 8);
  k_nchw_to_nhwc<false><<<grid, block, 0, (cudaStream_t)stream>>>(src, C, H * W, 1.f, nullptr, nullptr, dst);
  POD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// weight packing: (Cout, Cin, 3, 3) -> [Cout_pad][9*Cin], k = (ky*3+kx)*Cin + ci
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_w_split(const float* __restrict__ w, int Cout, int Cin, int Cout_pad, float scale,
                               __half* __restrict__ hi, __half* __restrict__ lo) {
  const int64_t total = (int64_t)Cout_pad * 9 * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i / (9 * Cin));
    const int k = (int)(i % (9 * Cin));
    const int tap = k / Cin, ci = k % Cin;
    float v = 0.f;
    if (co < Cout) v = w[((int64_t)co * Cin + ci) * 9 + tap];
    __half h, l;
    pod_split_h(v * scale, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

extern "C" __attribute__((visibility("default"))) int pod_pack_conv_weight(const float* w, int Cout, int Cin, int Cout_pad, float scale, void* dst_hi,
                                    void* dst_lo, void* stream) {
  POD_REQUIRE(w && dst_hi && dst_lo && Cout > 0 && Cin > 0 && Cout_pad >= Cout, "pod_pack_conv_weight: bad args");
  k_pack_w_split<<<512, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, Cout_pad, scale, (__half*)dst_hi, (__half*)dst_lo);
  POD_LAUNCH_CHECK();
  return 0;
}

__global__ void k_pack_w_f32(const float* __restrict__ w, int Cout, int Cin, int Cout_pad, float* __restrict__ dst) {
  const int64_t total = (int64_t)9 * Cin * Cout_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout_pad);
    const int k = (int)(i / Cout_pad);
    const int tap = k / Cin, ci = k % Cin;
    dst[i] = co < Cout ? w[((int64_t)co * Cin + ci) * 9 + tap] : 0.f;
  }
}

extern "C" __attribute__((visibility("default"))) int pod_pack_conv_weight_f32(const float* w, int Cout, int Cin, int Cout_pad, float* dst, void* stream) {
  POD_REQUIRE(w && dst && Cout > 0 && Cin > 0 && Cout_pad >= Cout, "pod_pack_conv_weight_f32: bad args");
  k_pack_w_f32<<<512, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, Cout_pad, dst);
  POD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// MC-dropout replication of the first tower layer: one read, samples*passes masked split copies
// ---------------------------------------------------------------------------------------------
__global__ void k_mask_expand(const float* __restrict__ x, int64_t quads_per_map, int NB_in, pod_dropout d, float scale,
                              uint32_t thr, float dscale, PhiloxKey key, __half* __restrict__ hi,
                              __half* __restrict__ lo) {
  const int reps = d.samples * d.passes;
  const int64_t total = quads_per_map * NB_in;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int nb = (int)(t / quads_per_map);
    const int64_t q = t % quads_per_map;
    const float4 v = reinterpret_cast<const float4*>(x)[t];
    for (int r = 0; r < reps; ++r) {
      const int sample = r / d.passes, pass = d.pass0 + r % d.passes;
      const uint4 w = philox4x32_10((uint32_t)q, pod_dropout_c1(d.level, d.layer, d.tower, pass), (uint32_t)sample,
                                    (uint32_t)(d.image0 + nb), key);
      float o[4];
      o[0] = w.x >= thr ? v.x * dscale : 0.f;
      o[1] = w.y >= thr ? v.y * dscale : 0.f;
      o[2] = w.z >= thr ? v.z * dscale : 0.f;
      o[3] = w.w >= thr ? v.w * dscale : 0.f;
      __half h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pod_split_h(o[i] * scale, h[i], l[i]);
      const int64_t oq = ((int64_t)nb * reps + r) * quads_per_map + q;
      reinterpret_cast<uint2*>(hi)[oq] = make_uint2(
          (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
          (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
      reinterpret_cast<uint2*>(lo)[oq] = make_uint2(
          (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16),
          (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16));
    }
  }
}

extern "C" __attribute__((visibility("default"))) int pod_mask_expand_split(const float* x, int NB_in, int HW, int C, const pod_dropout* d, float scale,
                                     void* dst_hi, void* dst_lo, void* stream) {
  POD_REQUIRE(x && d && dst_hi && dst_lo && NB_in > 0 && HW > 0 && C > 0 && C % 4 == 0, "pod_mask_expand_split: bad args");
  POD_REQUIRE(d->samples > 0 && d->passes > 0 && d->p > 0.0 && d->p < 1.0, "pod_mask_expand_split: bad dropout spec");
  const int64_t qpm = (int64_t)HW * C / 4;
  const int64_t total = qpm * NB_in;
  const int grid = (int)((total + 255) / 256 < (int64_t)pod_num_sms() * 16 ? (total + 255) / 256 : (int64_t)pod_num_sms() * 16);
  k_mask_expand<<<grid, 256, 0, (cudaStream_t)stream>>>(x, qpm, NB_in, *d, scale, pod_dropout_threshold(d->p),
                                                        pod_dropout_scale(d->p), pod_key(d->seed, POD_STREAM_DROPOUT),
                                                        (__half*)dst_hi, (__half*)dst_lo);
  POD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Q1 sample "mean": ((x0 + x0) + x1 + ... + x_{S-2}) / S, fp32, in the reference's order
// ---------------------------------------------------------------------------------------------
__global__ void k_sample_mean_q1(const float* __restrict__ x, int S, int64_t n, int64_t total, float* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / n, e = t % n;
    const float* p = x + b * S * n + e;
    float acc = p[0];
    for (int i = 0; i < S - 1; ++i) acc = __fadd_rn(acc, p[(int64_t)i * n]);
    out[t] = S > 1 ? __fdiv_rn(acc, (float)S) : acc;
  }
}

extern "C" __attribute__((visibility("default"))) int pod_sample_mean_q1(const float* x, int B, int S, int64_t n, float* out, void* stream) {
  POD_REQUIRE(x && out && B > 0 && S > 0 && n > 0, "pod_sample_mean_q1: bad args");
  const int64_t total = (int64_t)B * n;
  const int grid = (int)((total + 255) / 256 < (int64_t)pod_num_sms() * 32 ? (total + 255) / 256 : (int64_t)pod_num_sms() * 32);
  k_sample_mean_q1<<<grid, 256, 0, (cudaStream_t)stream>>>(x, S, n, total, out);
  POD_LAUNCH_CHECK();
  return 0;
}
