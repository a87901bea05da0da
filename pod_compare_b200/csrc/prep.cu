// Operand preparation and streaming statistics kernels (HBM-bound elementwise / transpose work).
#include "common.cuh"

// Device-side error word of this file's kernels (read and cleared by pod_status):
//   101 = a first-layer activation left the fp16 split range in k_mask_expand, 102 = non-finite input feature.
__device__ int g_prep_status = 0;

int pod_prep_status_fetch(int* v) {
  *v = 0;
  POD_CUDA(cudaMemcpyFromSymbol(v, g_prep_status, sizeof(int)));
  if (*v != 0) {
    int zero = 0;
    POD_CUDA(cudaMemcpyToSymbol(g_prep_status, &zero, sizeof(int)));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// fp16 split scale of the input feature maps, computed on the device (no host round trip):
//   k_absmax accumulates max|x| of one tensor into a word (bit pattern of a non-negative float orders like uint),
//   k_pow2_scale turns it into  min(max_scale, largest power of two s with amax * s <= target).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ amax_bits) {
  float m = 0.f;
  bool bad = false;
  const int64_t n4 = n / 4;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + t);
    const float a = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    bad |= !(fabsf(v.x) <= 3.0e38f) | !(fabsf(v.y) <= 3.0e38f) | !(fabsf(v.z) <= 3.0e38f) | !(fabsf(v.w) <= 3.0e38f);
    m = fmaxf(m, a);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t t = n4 * 4; t < n; ++t) {
      bad |= !(fabsf(x[t]) <= 3.0e38f);
      m = fmaxf(m, fabsf(x[t]));
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
  if (bad) atomicCAS(&g_prep_status, 0, 102);
}

__global__ void k_pow2_scale(const uint32_t* __restrict__ amax_bits, float max_scale, float target, float* __restrict__ scale) {
  const float amax = __uint_as_float(*amax_bits);
  float s = max_scale;
  if (amax > 0.f && amax <= 3.0e38f) {
    int e, et;
    const float m = frexpf(amax, &e);        // amax = m * 2^e, m in [0.5, 1)
    (void)frexpf(target, &et);               // target = 2^(et-1) (power of two)
    const float p = ldexpf(1.f, (et - 1) - e + (m == 0.5f ? 1 : 0));
    s = fminf(max_scale, p);
  }
  *scale = s;
}

extern "C" __attribute__((visibility("default"))) int pod_absmax_accumulate(const float* x, int64_t n, uint32_t* amax_bits, void* stream) {
  POD_REQUIRE(x && amax_bits && n > 0 && (uintptr_t)x % 16 == 0, "pod_absmax_accumulate: bad args (16-byte aligned input)");
  const int64_t want = (n / 4 + 255) / 256 + 1, cap = (int64_t)pod_num_sms() * 8;
  k_absmax<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(x, n, amax_bits);
  POD_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_pow2_scale_from_absmax(const uint32_t* amax_bits, float max_scale, float target, float* scale_dev,
                                                                   void* stream) {
  POD_REQUIRE(amax_bits && scale_dev && max_scale > 0.f && target > 0.f, "pod_pow2_scale_from_absmax: bad args");
  k_pow2_scale<<<1, 1, 0, (cudaStream_t)stream>>>(amax_bits, max_scale, target, scale_dev);
  POD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC (fp16 split pair | fp32): 32x32 shared-memory transpose tiles
// ---------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void k_nchw_to_nhwc(const float* __restrict__ src, int C, int HW, float scale, const float* __restrict__ scale_dev,
                               __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ dst32) {
  __shared__ float tile[32][33];
  if (SPLIT && scale_dev != nullptr) scale = __ldg(scale_dev);
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
  k_nchw_to_nhwc<true><<<grid, block, 0, (cudaStream_t)stream>>>(src, C, H * W, scale, nullptr, (__half*)dst_hi, (__half*)dst_lo,
                                                                 nullptr);
  POD_LAUNCH_CHECK();
  return 0;
}

// same, with the scale read from device memory (pod_pow2_scale_from_absmax): no host synchronisation
extern "C" __attribute__((visibility("default"))) int pod_nchw_to_nhwc_split_dev(const float* src, int NB, int C, int H, int W, const float* scale_dev,
                                          void* dst_hi, void* dst_lo, void* stream) {
  POD_REQUIRE(src && scale_dev && dst_hi && dst_lo && NB > 0 && C > 0 && H > 0 && W > 0, "pod_nchw_to_nhwc_split_dev: bad args");
  POD_REQUIRE(NB <= 65535, "pod_nchw_to_nhwc_split_dev: NB too large for one launch");
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, NB), block(32, 8);
  k_nchw_to_nhwc<true><<<grid, block, 0, (cudaStream_t)stream>>>(src, C, H * W, 1.f, scale_dev, (__half*)dst_hi, (__half*)dst_lo,
                                                                 nullptr);
  POD_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_nchw_to_nhwc_f32(const float* src, int NB, int C, int H, int W, float* dst, void* stream) {
  POD_REQUIRE(src && dst && NB > 0 && NB <= 65535 && C > 0 && H > 0 && W > 0, "pod_nchw_to_nhwc_f32: bad args");
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, NB), block(32, 8);
  k_nchw_to_nhwc<false><<<grid, block, 0, (cudaStream_t)stream>>>(src, C, H * W, 1.f, nullptr, nullptr, nullptr, dst);
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
// one thread = 8 consecutive channels (one Philox call, 16-bit lanes): 32-byte loads, 16-byte stores per copy
__global__ void __launch_bounds__(256)
k_mask_expand(const float* __restrict__ x, int64_t oct_per_map, int NB_in, pod_dropout d, float scale, const float* __restrict__ scale_dev,
              uint32_t thr, float dscale, PhiloxKey key, __half* __restrict__ hi, __half* __restrict__ lo, int live_reps) {
  if (scale_dev != nullptr) scale = __ldg(scale_dev);
  const int reps = d.samples * d.passes;
  const int64_t total = oct_per_map * NB_in;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int nb = (int)(t / oct_per_map);
    const int64_t o8 = t % oct_per_map;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(x) + 2 * t);
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(x) + 2 * t + 1);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    // the kept value (and its fp16 split) does not depend on the sample: split once, select per copy
    uint32_t kh[4], kl[4];
    bool saturated = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) saturated |= !(fabsf(v[i] * dscale * scale) <= 65504.f);
    if (saturated) atomicCAS(&g_prep_status, 0, 101);    // fp16 split range exceeded (or NaN): reported, never silent
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half h0, l0, h1, l1;
      pod_split_h(v[2 * i] * dscale * scale, h0, l0);
      pod_split_h(v[2 * i + 1] * dscale * scale, h1, l1);
      kh[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      kl[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    const uint32_t image = (uint32_t)(d.image0 + nb);
    for (int r = 0; r < live_reps; ++r) {
      const int sample = r / d.passes, pass = d.pass0 + r % d.passes;
      const uint32_t c1 = pod_dropout_c1(d.level, d.layer, d.tower, pass);
      // one call serves this thread's 8 channels (16-bit lanes, common.cuh)
      const uint32_t kb = pod_keep8(philox4x32_10((uint32_t)o8, c1, (uint32_t)sample, image, key), thr);
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t m = ((kb >> (2 * i)) & 1u ? 0x0000FFFFu : 0u) | ((kb >> (2 * i + 1)) & 1u ? 0xFFFF0000u : 0u);
        ph[i] = kh[i] & m;
        pl[i] = kl[i] & m;
      }
      const int64_t oq = ((int64_t)nb * reps + r) * oct_per_map + o8;
      __stcs(reinterpret_cast<uint4*>(hi) + oq, make_uint4(ph[0], ph[1], ph[2], ph[3]));   // streaming: written once,
      __stcs(reinterpret_cast<uint4*>(lo) + oq, make_uint4(pl[0], pl[1], pl[2], pl[3]));   // read by the next conv via TMA
    }
  }
}

extern "C" __attribute__((visibility("default"))) int pod_mask_expand_split(const float* x, int NB_in, int HW, int C, const pod_dropout* d, float scale,
                                     void* dst_hi, void* dst_lo, int live_reps, const float* scale_dev, void* stream) {
  POD_REQUIRE(x && d && dst_hi && dst_lo && NB_in > 0 && HW > 0 && C > 0 && C % 8 == 0, "pod_mask_expand_split: bad args (C%%8)");
  POD_REQUIRE(d->samples > 0 && d->passes > 0 && d->p > 0.0 && d->p < 1.0, "pod_mask_expand_split: bad dropout spec");
  POD_REQUIRE(live_reps >= 0 && live_reps <= d->samples * d->passes, "pod_mask_expand_split: live_reps out of range");
  if (live_reps == 0) live_reps = d->samples * d->passes;
  const int64_t qpm = (int64_t)HW * C / 8;
  const int64_t total = qpm * NB_in;
  // 39 registers x 256 threads: six CTAs are resident per SM; a grid of exactly that many leaves no partial last wave
  const int grid = (int)((total + 255) / 256 < (int64_t)pod_num_sms() * 6 ? (total + 255) / 256 : (int64_t)pod_num_sms() * 6);
  k_mask_expand<<<grid, 256, 0, (cudaStream_t)stream>>>(x, qpm, NB_in, *d, scale, scale_dev, pod_dropout_threshold16(d->p),
                                                        pod_dropout_scale(d->p), pod_key(d->seed, POD_STREAM_DROPOUT),
                                                        (__half*)dst_hi, (__half*)dst_lo, live_reps);
  POD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Q1 sample accumulation, second half: partial sums of the last tower layer (written by the tcgen05 epilogue,
// pod_conv_args.q1_acc) -> mean activation as fp16 split pair.  8 channels per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_q1_finish(const float* __restrict__ acc, int64_t oct_per_map, int64_t total, int groups, float inv_div, float scale,
            const float* __restrict__ scale_dev, __half* __restrict__ hi, __half* __restrict__ lo) {
  if (scale_dev != nullptr) scale = __ldg(scale_dev);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / oct_per_map, o8 = t % oct_per_map;
    const float4* p = reinterpret_cast<const float4*>(acc) + (m * groups * oct_per_map + o8) * 2;
    float4 a0 = __ldcs(p), a1 = __ldcs(p + 1);
    for (int g = 1; g < groups; ++g) {
      const float4 b0 = __ldcs(p + (int64_t)g * oct_per_map * 2), b1 = __ldcs(p + (int64_t)g * oct_per_map * 2 + 1);
      a0 = make_float4(__fadd_rn(a0.x, b0.x), __fadd_rn(a0.y, b0.y), __fadd_rn(a0.z, b0.z), __fadd_rn(a0.w, b0.w));
      a1 = make_float4(__fadd_rn(a1.x, b1.x), __fadd_rn(a1.y, b1.y), __fadd_rn(a1.z, b1.z), __fadd_rn(a1.w, b1.w));
    }
    const float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    uint32_t ph[4], pl[4];
    bool saturated = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float x0 = __fdiv_rn(v[2 * i], inv_div) * scale, x1 = __fdiv_rn(v[2 * i + 1], inv_div) * scale;
      saturated |= !(fabsf(x0) <= 65504.f) | !(fabsf(x1) <= 65504.f);
      __half h0, l0, h1, l1;
      pod_split_h(x0, h0, l0);
      pod_split_h(x1, h1, l1);
      ph[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      pl[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    if (saturated) atomicCAS(&g_prep_status, 0, 101);
    reinterpret_cast<uint4*>(hi)[t] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    reinterpret_cast<uint4*>(lo)[t] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// Q1 sample mean of the last tower layer, streaming form: reads the per-sample split-pair maps the tower kernel wrote
// (images x samples x passes, sample-major / pass-minor) and writes, for every pass whose bit is set in `mask`, the
// reference's weighted mean ((x_0 + x_0) + x_1 + ... + x_{L-1}) / S  (probabilistic_inference.py:214-270) as a split
// pair: the operand of the (linear) output convolutions cls_score / cls_var / bbox_cov, which then run once per image.
// HBM-bound: S - 1 maps read per map written; 8 channels per thread, eight samples in flight.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_q1_mean_act(const __half* __restrict__ hi, const __half* __restrict__ lo, int64_t oct_per_map, int64_t total, int S, int passes, int mask,
              int live0, int live1, float scale, const float* __restrict__ scale_dev, __half* __restrict__ ohi, __half* __restrict__ olo) {
  if (scale_dev != nullptr) scale = __ldg(scale_dev);
  const float inv_scale = 1.0f / scale;                 // power of two: exact
  const int n_acc = __popc(mask);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o8 = t % oct_per_map;
    const int64_t m = t / oct_per_map;                  // output map = image * n_acc + a
    const int a = (int)(m % n_acc);
    const int64_t b = m / n_acc;
    int p = 0;
    for (int k = 0, seen = 0; k < passes; ++k)
      if ((mask >> k) & 1) { if (seen == a) p = k; ++seen; }
    const int L = p == 0 ? live0 : live1;
    const uint4* ph = reinterpret_cast<const uint4*>(hi) + (b * S * passes + p) * oct_per_map + o8;
    const uint4* pl = reinterpret_cast<const uint4*>(lo) + (b * S * passes + p) * oct_per_map + o8;
    const int64_t step = (int64_t)passes * oct_per_map;
    float acc[8];
    auto unpack = [&](const uint4 h, const uint4 l, float (&v)[8]) {
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
        const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
        v[2 * i] = __fadd_rn(hf.x, lf.x) * inv_scale;
        v[2 * i + 1] = __fadd_rn(hf.y, lf.y) * inv_scale;
      }
    };
    {
      float v[8];
      unpack(__ldcs(ph), __ldcs(pl), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __fadd_rn(v[i], v[i]);        // sample 0 enters the reference's sum twice (quirk Q1)
    }
    int s = 1;
    for (; s + 8 <= L; s += 8) {                        // eight samples (sixteen 16-byte loads) in flight per thread
      uint4 h[8], l[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { h[u] = __ldcs(ph + (s + u) * step); l[u] = __ldcs(pl + (s + u) * step); }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float v[8];
        unpack(h[u], l[u], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = __fadd_rn(acc[i], v[i]);
      }
    }
    for (; s < L; ++s) {
      float v[8];
      unpack(__ldcs(ph + s * step), __ldcs(pl + s * step), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __fadd_rn(acc[i], v[i]);
    }
    uint32_t qh[4], ql[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half h0, l0, h1, l1;
      pod_split_h(__fdiv_rn(acc[2 * i], (float)S) * scale, h0, l0);
      pod_split_h(__fdiv_rn(acc[2 * i + 1], (float)S) * scale, h1, l1);
      qh[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      ql[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    reinterpret_cast<uint4*>(ohi)[t] = make_uint4(qh[0], qh[1], qh[2], qh[3]);
    reinterpret_cast<uint4*>(olo)[t] = make_uint4(ql[0], ql[1], ql[2], ql[3]);
  }
}

extern "C" __attribute__((visibility("default"))) int pod_q1_mean_act(const void* in_hi, const void* in_lo, int images, int samples, int passes, int mask,
                                                        const int* live /* host, per pass */, int64_t n, float scale, const float* scale_dev,
                                                        void* dst_hi, void* dst_lo, void* stream) {
  POD_REQUIRE(in_hi && in_lo && dst_hi && dst_lo && live && images > 0 && samples > 1 && (passes == 1 || passes == 2), "pod_q1_mean_act: bad args");
  POD_REQUIRE(mask > 0 && mask < (1 << passes) && n > 0 && n % 8 == 0 && (scale > 0.f || scale_dev), "pod_q1_mean_act: bad mask / size (n%%8)");
  for (int p = 0; p < passes; ++p)
    POD_REQUIRE(!((mask >> p) & 1) || (live[p] >= 1 && live[p] <= samples), "pod_q1_mean_act: live[%d] out of range", p);
  POD_REQUIRE(((uintptr_t)in_hi | (uintptr_t)in_lo | (uintptr_t)dst_hi | (uintptr_t)dst_lo) % 16 == 0, "pod_q1_mean_act: 16-byte alignment");
  const int n_acc = __builtin_popcount(mask);
  const int64_t opm = n / 8, total = opm * images * n_acc;
  const int64_t want = (total + 255) / 256, cap = (int64_t)pod_num_sms() * 6;
  k_q1_mean_act<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>((const __half*)in_hi, (const __half*)in_lo, opm, total, samples,
                                                                                 passes, mask, live[0], passes > 1 ? live[1] : 0, scale, scale_dev,
                                                                                 (__half*)dst_hi, (__half*)dst_lo);
  POD_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_q1_finish(const float* acc, int n_maps, int groups, int64_t n, int samples, float scale,
                                                      const float* scale_dev, void* dst_hi, void* dst_lo, void* stream) {
  POD_REQUIRE(acc && dst_hi && dst_lo && n_maps > 0 && groups > 0 && n > 0 && n % 8 == 0 && samples > 0, "pod_q1_finish: bad args (n%%8)");
  POD_REQUIRE(((uintptr_t)acc | (uintptr_t)dst_hi | (uintptr_t)dst_lo) % 16 == 0, "pod_q1_finish: buffers must be 16-byte aligned");
  const int64_t opm = n / 8, total = opm * n_maps;
  const int64_t want = (total + 255) / 256, cap = (int64_t)pod_num_sms() * 8;
  k_q1_finish<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(acc, opm, total, groups, (float)samples, scale, scale_dev,
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

// n % 4 == 0: four independent element chains per thread, 16-byte streaming loads issued eight samples at a time
// (the adds of one element stay in the reference's order: ((x0 + x0) + x1) + ... + x_{S-2}).
__global__ void __launch_bounds__(256)
k_sample_mean_q1_v4(const float4* __restrict__ x, int S, int64_t n4, int64_t total4, float4* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total4; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / n4, e = t % n4;
    const float4* p = x + b * S * n4 + e;
    float4 acc = __ldcs(p);
    int i = 0;
    for (; i + 8 <= S - 1; i += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + (int64_t)(i + u) * n4);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc.x = __fadd_rn(acc.x, v[u].x); acc.y = __fadd_rn(acc.y, v[u].y);
        acc.z = __fadd_rn(acc.z, v[u].z); acc.w = __fadd_rn(acc.w, v[u].w);
      }
    }
    for (; i < S - 1; ++i) {
      const float4 v = __ldcs(p + (int64_t)i * n4);
      acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
      acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
    }
    if (S > 1) {
      const float fs = (float)S;
      acc = make_float4(__fdiv_rn(acc.x, fs), __fdiv_rn(acc.y, fs), __fdiv_rn(acc.z, fs), __fdiv_rn(acc.w, fs));
    }
    out[t] = acc;
  }
}

extern "C" __attribute__((visibility("default"))) int pod_sample_mean_q1(const float* x, int B, int S, int64_t n, float* out, void* stream) {
  POD_REQUIRE(x && out && B > 0 && S > 0 && n > 0, "pod_sample_mean_q1: bad args");
  const int64_t total = (int64_t)B * n;
  if (n % 4 == 0 && ((uintptr_t)x | (uintptr_t)out) % 16 == 0) {
    const int64_t total4 = total / 4;
    const int64_t want = (total4 + 255) / 256, cap = (int64_t)pod_num_sms() * 6;   // six resident CTAs per SM (38 registers)
    k_sample_mean_q1_v4<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), S, n / 4, total4, reinterpret_cast<float4*>(out));
    POD_LAUNCH_CHECK();
    return 0;
  }
  const int grid = (int)((total + 255) / 256 < (int64_t)pod_num_sms() * 32 ? (total + 255) / 256 : (int64_t)pod_num_sms() * 32);
  k_sample_mean_q1<<<grid, 256, 0, (cudaStream_t)stream>>>(x, S, n, total, out);
  POD_LAUNCH_CHECK();
  return 0;
}
