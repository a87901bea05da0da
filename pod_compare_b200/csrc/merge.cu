// Post-NMS merging of per-run detections (MC-dropout samples or ensemble members): a sequential seed-selection kernel (one
// CTA per image) followed by a machine-wide statistics kernel (one warp per cluster).
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

// Shared by both phases: the runs of image b concatenated run-major into shared memory.
struct Concat {
  float4* box;
  float* area;
  int* cls;
  int* row;          // row of the per-run arrays this concatenated index came from
};

__device__ __forceinline__ int load_concat(const pod_merge_args& a, int b, unsigned char* smem_raw, int* s_off, Concat& c, int nthreads) {
  const int cap = a.runs * a.max_dets;
  c.box = reinterpret_cast<float4*>(smem_raw);
  c.area = reinterpret_cast<float*>(c.box + cap);
  c.cls = reinterpret_cast<int*>(c.area + cap);
  c.row = c.cls + cap;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int n = 0;
    for (int r = 0; r < a.runs; ++r) {
      s_off[r] = n;
      n += a.det_count[b * a.runs + r];
    }
    s_off[a.runs] = n;
  }
  __syncthreads();
  for (int r = 0; r < a.runs; ++r) {
    const int cnt = s_off[r + 1] - s_off[r];
    for (int i = tid; i < cnt; i += nthreads) {
      const int row = (b * a.runs + r) * a.max_dets + i;
      const float4 q = reinterpret_cast<const float4*>(a.det_boxes)[row];
      const int j = s_off[r] + i;
      c.box[j] = q;
      c.area[j] = area_of(q);
      c.cls[j] = a.det_classes[row];
      c.row[j] = row;
    }
  }
  __syncthreads();
  return s_off[a.runs];
}

// ---- phase 1: sequential clustering (inference_utils.py:202-215), one CTA of 1024 threads per image.  The only
// sequential part of the merge: box i seeds a cluster unless an earlier seed already claimed it.  Writes the seed list.
constexpr int NT_SEED = 1024;
__global__ void __launch_bounds__(NT_SEED) k_cluster_seeds(pod_merge_args a) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ int s_off[65];
  __shared__ int s_nclusters;
  const int cap = a.runs * a.max_dets;
  const int b = blockIdx.x, tid = threadIdx.x;
  Concat c;
  const int n = load_concat(a, b, smem_raw, s_off, c, NT_SEED);
  unsigned char* s_assigned = reinterpret_cast<unsigned char*>(c.row + cap);
  for (int j = tid; j < n; j += NT_SEED) s_assigned[j] = 0;
  if (tid == 0) s_nclusters = 0;
  __syncthreads();
  const float aff = (float)a.affinity;
  int* seeds = a.seed_scratch + (int64_t)b * cap;
  for (int i = 0; i < n; ++i) {
    if (s_assigned[i]) continue;                       // uniform: read after the barrier below
    const float4 bi = c.box[i];
    const float ai = c.area[i];
    const int ci = c.cls[i];
    __syncthreads();                                   // everyone has read assigned[i] before it may change
    for (int j = tid; j < n; j += NT_SEED)
      if (c.cls[j] == ci && iou_d2(bi, ai, c.box[j], c.area[j]) >= aff) s_assigned[j] = 1;
    if (tid == 0) seeds[s_nclusters++] = i;
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) a.out_count[b] = s_nclusters;
}

// ---- phase 2: cluster statistics (inference_utils.py:223-247), one warp per cluster, gridDim.y CTAs per image ----
__global__ void __launch_bounds__(NT) k_cluster_merge(pod_merge_args a) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ int s_off[65];
  const int cap = a.runs * a.max_dets;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float aff = (float)a.affinity;
  const int nc = a.out_count[b];
  if ((int)blockIdx.y * (NT / 32) >= nc) return;       // uniform per CTA: nothing to do
  Concat cc_;
  const int n = load_concat(a, b, smem_raw, s_off, cc_, NT);
  float4* s_box = cc_.box;
  float* s_area = cc_.area;
  int* s_cls = cc_.cls;
  int* s_row = cc_.row;
  const int* s_seed = a.seed_scratch + (int64_t)b * cap;

  float* o_boxes = a.out_boxes + (int64_t)b * cap * 4;
  float* o_cov = a.out_cov + (int64_t)b * cap * 16;
  float* o_scores = a.out_scores + (int64_t)b * cap;
  int* o_classes = a.out_classes + (int64_t)b * cap;
  float* o_probs = a.out_probs + (int64_t)b * cap * a.K;
  for (int c = blockIdx.y * (NT / 32) + warp; c < nc; c += gridDim.y * (NT / 32)) {
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
  POD_REQUIRE(a->seed_scratch, "pod_cluster_merge: seed_scratch (B x runs*max_dets ints) is required");
  const size_t smem = (size_t)cap * (16 + 4 + 4 + 4 + 1) + 16;
  POD_REQUIRE(smem <= 200 * 1024, "pod_cluster_merge: too many detections for shared memory");
  static size_t configured = 0;
  if (smem > configured) {
    POD_CUDA(cudaFuncSetAttribute(k_cluster_seeds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POD_CUDA(cudaFuncSetAttribute(k_cluster_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  // phase 1: the sequential seed selection (one CTA per image); phase 2: per-cluster statistics, one warp per cluster,
  // spread over enough CTAs per image to fill the machine (the single-CTA form spent 44 ms per 3000 detections)
  k_cluster_seeds<<<a->B, NT_SEED, smem, (cudaStream_t)stream>>>(*a);
  POD_LAUNCH_CHECK();
  int per_image = (2 * pod_num_sms() + a->B - 1) / a->B;
  const int max_useful = (cap + NT / 32 - 1) / (NT / 32);
  if (per_image > max_useful) per_image = max_useful;
  if (per_image < 1) per_image = 1;
  k_cluster_merge<<<dim3(a->B, per_image), NT, smem, (cudaStream_t)stream>>>(*a);
  POD_LAUNCH_CHECK();
  return 0;
}
