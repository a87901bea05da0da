// Shared device/host helpers of libpodb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/podb200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing: thread-local message, int return codes (no exceptions cross the ABI)
// ---------------------------------------------------------------------------------------------
void pod_set_error(const char* fmt, ...);
// device-side error words of the kernel files (each reads and clears its own); combined by pod_status (api.cu)
int pod_tc_status_fetch(int* v);
int pod_prep_status_fetch(int* v);
int pod_backbone_status_fetch(int* v);

#define POD_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      pod_set_error(__VA_ARGS__);                                \
      return -1;                                                 \
    }                                                            \
  } while (0)

#define POD_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      pod_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                           \
    }                                                                                           \
  } while (0)

#define POD_LAUNCH_CHECK()                                                                      \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess) {                                                                    \
      pod_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                           \
    }                                                                                           \
  } while (0)

static inline int pod_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 and the three streams (contract: oracle/philox.py)
// ---------------------------------------------------------------------------------------------
#define POD_STREAM_DROPOUT 0x0D120F01u
#define POD_STREAM_LOGIT 0x0D120F02u
#define POD_STREAM_BOX 0x0D120F03u

struct PhiloxKey {
  uint32_t k0, k1;
};

__host__ __device__ inline PhiloxKey pod_key(uint64_t seed, uint32_t stream) {
  PhiloxKey k;
  k.k0 = (uint32_t)(seed & 0xFFFFFFFFull);
  k.k1 = (uint32_t)(seed >> 32) ^ stream;
  return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, PhiloxKey key) {
  uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__host__ __device__ inline uint32_t pod_dropout_threshold(double p) {
  // floor(p * 2^32) for 0 <= p < 1, evaluated in double like the oracle
  double t = p * 4294967296.0;
  if (t <= 0.0) return 0u;
  if (t >= 4294967295.0) return 4294967295u;
  return (uint32_t)t;
}

// The dropout stream spends 16 random bits per Bernoulli decision: one Philox4x32-10 call serves EIGHT consecutive
// elements (lane i = bits [16*(i%2), 16*(i%2)+16) of word i/2; keep iff lane >= floor(p * 2^16)).  Half the integer work
// of a 32-bit-per-decision stream in the two places that evaluate masks for every activation (k_mask_expand, the tower
// epilogue); the keep probability is (65536 - floor(65536 p)) / 65536, within 1.5e-5 of 1 - p.
__host__ __device__ inline uint32_t pod_dropout_threshold16(double p) {
  double t = p * 65536.0;
  if (t <= 0.0) return 0u;
  if (t >= 65535.0) return 65535u;
  return (uint32_t)t;
}
// keep bits of the 8 elements served by one call: bit i <-> element 8*c0 + i
__device__ __forceinline__ uint32_t pod_keep8(const uint4 w, uint32_t thr16) {
  return ((w.x & 0xFFFFu) >= thr16 ? 1u : 0u) | ((w.x >> 16) >= thr16 ? 2u : 0u) | ((w.y & 0xFFFFu) >= thr16 ? 4u : 0u) |
         ((w.y >> 16) >= thr16 ? 8u : 0u) | ((w.z & 0xFFFFu) >= thr16 ? 16u : 0u) | ((w.z >> 16) >= thr16 ? 32u : 0u) |
         ((w.w & 0xFFFFu) >= thr16 ? 64u : 0u) | ((w.w >> 16) >= thr16 ? 128u : 0u);
}

__host__ __device__ inline uint32_t pod_dropout_c1(int level, int layer, int tower, int pass) {
  return (uint32_t)(level & 0xFF) | ((uint32_t)(layer & 0xFF) << 8) | ((uint32_t)(tower & 0xFF) << 16) |
         ((uint32_t)(pass & 0xFF) << 24);
}

// uniform in (0,1) from the top 23 bits; exact in fp32
__device__ __forceinline__ float pod_u23(uint32_t w) { return ((float)(w >> 9) + 0.5f) * 1.1920928955078125e-07f; }

// Box-Muller pair: r = sqrt(-2 ln ua); (r cos 2pi ub, r sin 2pi ub)
__device__ __forceinline__ void pod_box_muller(uint32_t wa, uint32_t wb, float& n0, float& n1) {
  const float ua = pod_u23(wa), ub = pod_u23(wb);
  const float r = sqrtf(-2.0f * logf(ua));
  float s, c;
  sincospif(2.0f * ub, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// fp16 split of x*scale: hi = rn(x*scale), lo = rn(x*scale - hi), with lo rounded to POD_LO_BITS mantissa bits.
// hi carries 11 significant bits (|x*s - hi| <= 2^-11 |x*s|) and lo continues them with POD_LO_BITS + 1 more, so the
// pair represents x*s to 2^-(12 + POD_LO_BITS): 2^-19 = 1.9e-6 worst case, ~8e-7 rms per element at the default of 7
// (2^-22 with all 10 bits) -- the size of the 1.4e-6 rms the tensor core's truncating accumulation leaves in a
// convolution anyway; measured end to end at 1280x720 the scores / boxes / covariances deviate from the oracle by
// 1.13e-5 / 1.3e-3 px / 4.9e-5 with 7 bits against 1.12e-5 / 1.4e-3 px / 5.3e-5 with 10 (profiles/r2b_lo_bits.txt).
// Why not all 10 bits: the tower kernel is power-bound and the multipliers' energy depends on the
// operand bits -- measured with tools/tower_clock_probe.py at the 1 kW cap: 34.9 ms per P3 launch with 10 lo mantissa
// bits, 33.9 with 5, 32.6 with 0, 30.8 with lo = 0; whole step, same box: 615.8 ms (10 bits), 608.1 (7), 601.2 (5).
// Low-order bits that no tolerance can see were costing clock.
#ifndef POD_LO_BITS
#define POD_LO_BITS 7
#endif
__device__ __forceinline__ __half pod_round_lo(__half lo) {
  if (POD_LO_BITS >= 10) return lo;
  constexpr unsigned DROP = 10 - (POD_LO_BITS < 10 ? POD_LO_BITS : 10);
  // round-to-nearest (ties away) on the magnitude bits; a carry out of the mantissa correctly bumps the exponent
  const unsigned short b = __half_as_ushort(lo);
  const unsigned short r = (unsigned short)(((b & 0x7FFFu) + (1u << (DROP - 1))) & ~((1u << DROP) - 1u)) | (b & 0x8000u);
  return __ushort_as_half(r);
}
__device__ __forceinline__ void pod_split_h(float xs, __half& hi, __half& lo) {
  hi = __float2half_rn(xs);
  lo = pod_round_lo(__float2half_rn(xs - __half2float(hi)));
}

// 1/(1-p) as torch computes the dropout scale: fp32(1) / fp32(1-p)
__host__ __device__ inline float pod_dropout_scale(double p) { return 1.0f / (float)(1.0 - p); }
