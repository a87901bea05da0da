// libpodb200: error state, version, device probe and the random-stream test hooks.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void pod_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" __attribute__((visibility("default"))) const char* pod_last_error(void) { return g_err; }
extern "C" __attribute__((visibility("default"))) int pod_version(void) { return POD_ABI_VERSION; }

// Combined device-side error word of the library (synchronises the device): the product path calls it once per
// inference call and raises -- a bounded barrier wait that expired or an activation outside the fp16 split range
// must never come back as a plausible-looking result.
extern "C" __attribute__((visibility("default"))) int pod_status(int* status_host) {
  POD_REQUIRE(status_host, "pod_status: null");
  int a = 0, b = 0, c = 0, rc;
  if ((rc = pod_tc_status_fetch(&a))) return rc;
  if ((rc = pod_prep_status_fetch(&b))) return rc;
  if ((rc = pod_backbone_status_fetch(&c))) return rc;
  *status_host = a != 0 ? a : (b != 0 ? b : c);
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_dropout_mask(uint8_t* keep, int64_t nocts, PhiloxKey key, uint32_t c1, uint32_t sample,
                               uint32_t image, uint32_t thr16) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nocts; q += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t b = pod_keep8(philox4x32_10((uint32_t)q, c1, sample, image, key), thr16);
    uchar4 lo4, hi4;
    lo4.x = b & 1u; lo4.y = (b >> 1) & 1u; lo4.z = (b >> 2) & 1u; lo4.w = (b >> 3) & 1u;
    hi4.x = (b >> 4) & 1u; hi4.y = (b >> 5) & 1u; hi4.z = (b >> 6) & 1u; hi4.w = (b >> 7) & 1u;
    reinterpret_cast<uchar4*>(keep)[2 * q] = lo4;
    reinterpret_cast<uchar4*>(keep)[2 * q + 1] = hi4;
  }
}

extern "C" __attribute__((visibility("default"))) int pod_philox_dropout_mask(uint8_t* keep_hwc, int H, int W, int C, uint64_t seed, int image, int sample,
                                       int pass, int tower, int layer, int level, double p, void* stream) {
  POD_REQUIRE(keep_hwc && H > 0 && W > 0 && C > 0 && C % 8 == 0, "pod_philox_dropout_mask: bad shape (C%%8)");
  const int64_t nq = (int64_t)H * W * C / 8;
  const int grid = (int)((nq + 255) / 256 < 4096 ? (nq + 255) / 256 : 4096);
  k_dropout_mask<<<grid, 256, 0, (cudaStream_t)stream>>>(keep_hwc, nq, pod_key(seed, POD_STREAM_DROPOUT),
                                                         pod_dropout_c1(level, layer, tower, pass), (uint32_t)sample,
                                                         (uint32_t)image, pod_dropout_threshold16(p));
  POD_LAUNCH_CHECK();
  return 0;
}

__global__ void k_logit_normals(float* out, int draws, int64_t n, PhiloxKey key, uint32_t level, uint32_t image) {
  const int64_t nq = (n + 3) / 4;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nq * draws; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = t % nq;
    const int j = (int)(t / nq);
    const uint4 w = philox4x32_10((uint32_t)q, level, (uint32_t)j, image, key);
    float v[4];
    pod_box_muller(w.x, w.y, v[0], v[1]);
    pod_box_muller(w.z, w.w, v[2], v[3]);
    for (int i = 0; i < 4; ++i)
      if (q * 4 + i < n) out[(int64_t)j * n + q * 4 + i] = v[i];
  }
}

extern "C" __attribute__((visibility("default"))) int pod_philox_logit_normals(float* out, int draws, int n_anchor, int K, uint64_t seed, int image, int level,
                                        void* stream) {
  POD_REQUIRE(out && draws > 0 && n_anchor > 0 && K > 0, "pod_philox_logit_normals: bad shape");
  const int64_t n = (int64_t)n_anchor * K;
  k_logit_normals<<<1024, 256, 0, (cudaStream_t)stream>>>(out, draws, n, pod_key(seed, POD_STREAM_LOGIT),
                                                          (uint32_t)level, (uint32_t)image);
  POD_LAUNCH_CHECK();
  return 0;
}

__global__ void k_box_normals(float* out, const int64_t* ids, int M, int draws, PhiloxKey key, uint32_t image) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)M * draws;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(t % M);
    const int j = (int)(t / M);
    const uint4 w = philox4x32_10((uint32_t)ids[m], 0u, (uint32_t)j, image, key);
    float4 v;
    pod_box_muller(w.x, w.y, v.x, v.y);
    pod_box_muller(w.z, w.w, v.z, v.w);
    reinterpret_cast<float4*>(out)[t] = v;
  }
}

extern "C" __attribute__((visibility("default"))) int pod_philox_box_normals(float* out, const int64_t* anchor_ids, int M, int draws, uint64_t seed, int image,
                                      void* stream) {
  POD_REQUIRE(out && anchor_ids && M > 0 && draws > 0, "pod_philox_box_normals: bad shape");
  k_box_normals<<<1024, 256, 0, (cudaStream_t)stream>>>(out, anchor_ids, M, draws, pod_key(seed, POD_STREAM_BOX),
                                                        (uint32_t)image);
  POD_LAUNCH_CHECK();
  return 0;
}
