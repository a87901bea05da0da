// Post-NMS merging of per-run detections (MC-dropout samples or ensemble members), one CTA per image.
// Replaces /root/reference/src/probabilistic_inference/inference_utils.py:165-266
// (general_black_box_ensembles_post_processing up to its final NMS, which pod_nms_fuse then performs):
//   * all runs of an image are concatenated run-major,
//   * sequential clustering: box i seeds a cluster unless it is already a member of an earlier one;
//     members = { j : IoU(i, j) >= affinity  and  class_j == class_i }  (detectron2 pairwise_iou,
//     non-strict >=; a box may belong to several clusters),
//   * per cluster: mean box, unbiased sample covariance of the member boxes + mean member covariance
//     (a single member keeps its own), mean probability vector, score/class = its max/argmax.
#include "common.cuh"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float area_of(const float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

__device__ __forceinline__ float iou_d2(const float4 a, float area_a, const float4 b, float area_b) {
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  return inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter)) : 0.f;
}

__device__ __forceinline__ double wsum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wsum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(NT) k_cluster_merge(pod_merge_args a) {
  extern __shared__ unsigned char smem_raw[];
  const int cap = a.runs * a.max_dets;
  float4* s_box = reinterpret_cast<float4*>(smem_raw);
  float* s_area = reinterpret_cast<float*>(s_box + cap);
  int* s_cls = reinterpret_cast<int*>(s_area + cap);
  int* s_row = s_cls + cap;          // row of the per-run arrays this concatenated index came from
  int* s_seed = s_row + cap;
  unsigned char* s_assigned = reinterpret_cast<unsigned char*>(s_seed + cap);
  __shared__ int s_off[65];
  __shared__ int s_nclusters;

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float aff = (float)a.affinity;
  if (tid == 0) {
    int n = 0;
    for (int r = 0; r < a.runs; ++r) {
      s_off[r] = n;
      n += a.det_count[b * a.runs + r];
    }
    s_off[a.runs] = n;
    s_nclusters = 0;
  }
  __syncthreads();
  const int n = s_off[a.runs];
  for (int r = 0; r < a.runs; ++r) {
    const int cnt = s_off[r + 1] - s_off[r];
    for (int i = tid; i < cnt; i += NT) {
      const int row = (b * a.runs + r) * a.max_dets + i;
      const float4 q = reinterpret_cast<const float4*>(a.det_boxes)[row];
      const int j = s_off[r] + i;
      s_box[j] = q;
      s_area[j] = area_of(q);
      s_cls[j] = a.det_classes[row];
      s_row[j] = row;
      s_assigned[j] = 0;
    }
  }
  __syncthreads();

  // ---- sequential clustering (inference_utils.py:202-215) ----
  for (int i = 0; i < n; ++i) {
    if (s_assigned[i]) continue;                       // uniform: read after the barrier below
    const float4 bi = s_box[i];
    const float ai = s_area[i];
    const int ci = s_cls[i];
    __syncthreads();                                   // everyone has read assigned[i] before it may change
    for (int j = tid; j < n; j += NT)
      if (s_cls[j] == ci && iou_d2(bi, ai, s_box[j], s_area[j]) >= aff) s_assigned[j] = 1;
    if (tid == 0) s_seed[s_nclusters++] = i;
    __syncthreads();
  }
  __syncthreads();
  const int nc = s_nclusters;

  // ---- cluster statistics (inference_utils.py:223-247), one warp per cluster ----
  float* o_boxes = a.out_boxes + (int64_t)b * cap * 4;
  float* o_cov = a.out_cov + (int64_t)b * cap * 16;
  float* o_scores = a.out_scores + (int64_t)b * cap;
  int* o_classes = a.out_classes + (int64_t)b * cap;
  float* o_probs = a.out_probs + (int64_t)b * cap * a.K;
  for (int c = warp; c < nc; c += NT / 32) {
    const int i = s_seed[c];
    const float4 bi = s_box[i];
    const float ai = s_area[i];
    const int ci = s_cls[i];
    int m = 0;
    double sb[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = lane; j < n; j += 32) {
      if (s_cls[j] != ci || !(iou_d2(bi, ai, s_box[j], s_area[j]) >= aff)) continue;
      ++m;
      const float4 q = s_box[j];
      sb[0] += q.x; sb[1] += q.y; sb[2] += q.z; sb[3] += q.w;
    }
    m = __reduce_add_sync(0xffffffffu, m);
    for (int e = 0; e < 4; ++e) sb[e] = wsum_d(sb[e]);
    if (m == 0) {
      // degenerate (zero-area) seed: the reference would produce NaNs from an empty cluster; emit the
      // seed itself with zero score so the final NMS ranks it last
      if (lane == 0) {
        reinterpret_cast<float4*>(o_boxes)[c] = bi;
        for (int e = 0; e < 16; ++e) o_cov[(int64_t)c * 16 + e] = a.det_cov[(int64_t)s_row[i] * 16 + e];
        o_scores[c] = 0.f;
        o_classes[c] = ci;
      }
      for (int k = lane; k < a.K; k += 32) o_probs[(int64_t)c * a.K + k] = 0.f;
      continue;
    }
    float mean[4];
    for (int e = 0; e < 4; ++e) mean[e] = (float)(sb[e] / (double)m);
    double cc[16], mc[16];
    for (int e = 0; e < 16; ++e) { cc[e] = 0.0; mc[e] = 0.0; }
    for (int j = lane; j < n; j += 32) {
      if (s_cls[j] != ci || !(iou_d2(bi, ai, s_box[j], s_area[j]) >= aff)) continue;
      const float4 q = s_box[j];
      const double r[4] = {(double)__fsub_rn(q.x, mean[0]), (double)__fsub_rn(q.y, mean[1]),
                           (double)__fsub_rn(q.z, mean[2]), (double)__fsub_rn(q.w, mean[3])};
      for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y) cc[x * 4 + y] += r[x] * r[y];
      const float* cj = a.det_cov + (int64_t)s_row[j] * 16;
      for (int e = 0; e < 16; ++e) mc[e] += (double)cj[e];
    }
    for (int e = 0; e < 16; ++e) {
      cc[e] = wsum_d(cc[e]);
      mc[e] = wsum_d(mc[e]) / (double)m;
    }
    float best = -1.f;
    int bestk = 0;
    for (int k = 0; k < a.K; ++k) {
      float acc = 0.f;
      for (int j = lane; j < n; j += 32)
        if (s_cls[j] == ci && iou_d2(bi, ai, s_box[j], s_area[j]) >= aff) acc += a.det_probs[(int64_t)s_row[j] * a.K + k];
      acc = wsum_f(acc) / (float)m;
      if (lane == 0) o_probs[(int64_t)c * a.K + k] = acc;
      if (acc > best) { best = acc; bestk = k; }
    }
    if (lane == 0) {
      reinterpret_cast<float4*>(o_boxes)[c] = make_float4(mean[0], mean[1], mean[2], mean[3]);
      for (int e = 0; e < 16; ++e) {
        const float epi = m >= 2 ? (float)(cc[e] / (double)(m - 1)) : 0.f;
        o_cov[(int64_t)c * 16 + e] = m >= 2 ? __fadd_rn(epi, (float)mc[e]) : (float)mc[e];
      }
      o_scores[c] = best;
      o_classes[c] = bestk;
    }
  }
  if (tid == 0) a.out_count[b] = nc;
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int pod_cluster_merge(const pod_merge_args* a, void* stream) {
  POD_REQUIRE(a, "pod_cluster_merge: null args");
  POD_REQUIRE(a->det_boxes && a->det_cov && a->det_probs && a->det_classes && a->det_count, "pod_cluster_merge: null input");
  POD_REQUIRE(a->out_boxes && a->out_cov && a->out_scores && a->out_classes && a->out_probs && a->out_count,
              "pod_cluster_merge: null output");
  POD_REQUIRE(a->B > 0 && a->runs > 0 && a->runs <= 64 && a->max_dets > 0 && a->K > 0, "pod_cluster_merge: bad shape (runs <= 64)");
  const int cap = a->runs * a->max_dets;
  POD_REQUIRE(cap <= 8192, "pod_cluster_merge: runs*max_dets must be <= 8192");
  const size_t smem = (size_t)cap * (16 + 4 + 4 + 4 + 4 + 1) + 16;
  POD_REQUIRE(smem <= 200 * 1024, "pod_cluster_merge: too many detections for shared memory");
  static size_t configured = 0;
  if (smem > configured) {
    POD_CUDA(cudaFuncSetAttribute(k_cluster_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  k_cluster_merge<<<a->B, NT, smem, (cudaStream_t)stream>>>(*a);
  POD_LAUNCH_CHECK();
  return 0;
}
