// NMS, BayesOD clustering + Bayesian box fusion, anchor-statistics clustering, and the final
// rescale/clip -- one CTA per image.
//
// Replaces, in /root/reference/src:
//   probabilistic_inference/inference_utils.py:12-54      general_standard_nms_postprocessing
//   probabilistic_inference/probabilistic_inference.py:536-636   post_processing_bayes_od
//   probabilistic_inference/inference_utils.py:292-334    bounding_box_bayesian_inference
//   probabilistic_inference/inference_utils.py:57-162     general_anchor_statistics_postprocessing
//   probabilistic_inference/inference_utils.py:374-425    probabilistic_detector_postprocess
// and the third-party ops they call: torchvision.ops.batched_nms (ops/boxes.py:51-120, both the
// per-class and the coordinate-offset variants) over torchvision's CPU nms loop (strict '>',
// stable descending score order, fp32 IEEE arithmetic without contraction), and detectron2's
// pairwise_iou / Boxes.scale / clip / nonempty.
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int SORT_MAX = 8192;
constexpr int MAX_DETS = 256;
constexpr int MAX_K = 128;

struct BoxA {
  float x1, y1, x2, y2, area;
  int cls;
};

__device__ __forceinline__ float box_area(float x1, float y1, float x2, float y2) {
  return __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
}

// torchvision cpu/nms_kernel.cpp inner test
__device__ __forceinline__ bool nms_suppresses(const BoxA& i, const BoxA& j, double thr) {
  const float xx1 = fmaxf(i.x1, j.x1), yy1 = fmaxf(i.y1, j.y1);
  const float xx2 = fminf(i.x2, j.x2), yy2 = fminf(i.y2, j.y2);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(i.area, j.area), inter));
  return (double)ovr > thr;
}

// detectron2 pairwise_iou element
__device__ __forceinline__ float d2_iou(const float4 a, float area_a, const float4 b, float area_b) {
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  return inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter)) : 0.f;
}

// 4x4 inverse + determinant by Laplace expansion over 2x2 minors (fp64)
__device__ __forceinline__ double inv4(const double a[16], double b[16]) {
  const double s0 = a[0] * a[5] - a[4] * a[1], s1 = a[0] * a[6] - a[4] * a[2], s2 = a[0] * a[7] - a[4] * a[3];
  const double s3 = a[1] * a[6] - a[5] * a[2], s4 = a[1] * a[7] - a[5] * a[3], s5 = a[2] * a[7] - a[6] * a[3];
  const double c5 = a[10] * a[15] - a[14] * a[11], c4 = a[9] * a[15] - a[13] * a[11], c3 = a[9] * a[14] - a[13] * a[10];
  const double c2 = a[8] * a[15] - a[12] * a[11], c1 = a[8] * a[14] - a[12] * a[10], c0 = a[8] * a[13] - a[12] * a[9];
  const double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  const double r = 1.0 / det;
  b[0] = (a[5] * c5 - a[6] * c4 + a[7] * c3) * r;
  b[1] = (-a[1] * c5 + a[2] * c4 - a[3] * c3) * r;
  b[2] = (a[13] * s5 - a[14] * s4 + a[15] * s3) * r;
  b[3] = (-a[9] * s5 + a[10] * s4 - a[11] * s3) * r;
  b[4] = (-a[4] * c5 + a[6] * c2 - a[7] * c1) * r;
  b[5] = (a[0] * c5 - a[2] * c2 + a[3] * c1) * r;
  b[6] = (-a[12] * s5 + a[14] * s2 - a[15] * s1) * r;
  b[7] = (a[8] * s5 - a[10] * s2 + a[11] * s1) * r;
  b[8] = (a[4] * c4 - a[5] * c2 + a[7] * c0) * r;
  b[9] = (-a[0] * c4 + a[1] * c2 - a[3] * c0) * r;
  b[10] = (a[12] * s4 - a[13] * s2 + a[15] * s0) * r;
  b[11] = (-a[8] * s4 + a[9] * s2 - a[11] * s0) * r;
  b[12] = (-a[4] * c3 + a[5] * c1 - a[6] * c0) * r;
  b[13] = (a[0] * c3 - a[1] * c1 + a[2] * c0) * r;
  b[14] = (-a[12] * s3 + a[13] * s1 - a[14] * s0) * r;
  b[15] = (a[8] * s3 - a[9] * s1 + a[10] * s0) * r;
  return det;
}

__device__ __forceinline__ double det4(const double a[16]) {
  const double s0 = a[0] * a[5] - a[4] * a[1], s1 = a[0] * a[6] - a[4] * a[2], s2 = a[0] * a[7] - a[4] * a[3];
  const double s3 = a[1] * a[6] - a[5] * a[2], s4 = a[1] * a[7] - a[5] * a[3], s5 = a[2] * a[7] - a[6] * a[3];
  const double c5 = a[10] * a[15] - a[14] * a[11], c4 = a[9] * a[15] - a[13] * a[11], c3 = a[9] * a[14] - a[13] * a[10];
  const double c2 = a[8] * a[15] - a[12] * a[11], c1 = a[8] * a[14] - a[12] * a[10], c0 = a[8] * a[13] - a[12] * a[9];
  return s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(NT) k_nms_fuse(pod_nms_args a, float sx, float sy) {
  extern __shared__ unsigned long long s_keys[];          // sort buffer (npow2 entries)
  __shared__ BoxA s_kept[MAX_DETS];
  __shared__ int s_kept_idx[MAX_DETS];
  __shared__ unsigned char s_alive[NT];
  __shared__ int s_first;
  __shared__ int s_nkept;
  __shared__ float s_red[NT / 32];
  __shared__ float s_out_box[MAX_DETS][4];
  __shared__ float s_out_score[MAX_DETS];
  __shared__ int s_out_cls[MAX_DETS];
  __shared__ int s_flag[MAX_DETS];
  __shared__ int s_pos[MAX_DETS];

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = a.count[b];
  const float4* boxes = reinterpret_cast<const float4*>(a.boxes) + (int64_t)b * a.cap;
  const float* covs = a.cov + (int64_t)b * a.cap * 16;
  const float* scores = a.scores + (int64_t)b * a.cap;
  const int* classes = a.classes + (int64_t)b * a.cap;
  const float* probs = a.probs + (int64_t)b * a.cap * a.K;
  float* det_boxes = a.det_boxes + (int64_t)b * a.max_dets * 4;
  float* det_cov = a.det_cov + (int64_t)b * a.max_dets * 16;
  float* det_scores = a.det_scores + (int64_t)b * a.max_dets;
  int* det_classes = a.det_classes + (int64_t)b * a.max_dets;
  float* det_probs = a.det_probs + (int64_t)b * a.max_dets * a.K;
  int* keep_out = a.keep + (int64_t)b * a.max_dets;
  int* src_out = a.det_src + (int64_t)b * a.max_dets;

  if (M <= 0) {
    if (tid == 0) { a.det_count[b] = 0; a.keep_count[b] = 0; }
    return;
  }

  // ---- 1. order by (score desc, index asc): bitonic sort of unique 64-bit keys ---------------
  int npow2 = 1;
  while (npow2 < M) npow2 <<= 1;
  for (int i = tid; i < npow2; i += NT)
    s_keys[i] = i < M ? (((unsigned long long)__float_as_uint(scores[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i))
                      : 0ull;
  __syncthreads();
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < npow2 / 2; t += NT) {
        const int i = 2 * t - (t & (stride - 1));     // lower index of the pair
        const int j = i + stride;
        const unsigned long long x = s_keys[i], y = s_keys[j];
        const bool desc = (i & size) == 0;
        if (desc ? (x < y) : (x > y)) { s_keys[i] = y; s_keys[j] = x; }
      }
      __syncthreads();
    }
  }

  // ---- 2. torchvision batched_nms variant ----------------------------------------------------
  const bool trick = a.nms_variant == 1 || (a.nms_variant == 2 && 4 * M <= 4000);
  float offset_unit = 0.f;
  if (trick) {
    float mx = -INFINITY;
    for (int i = tid; i < M; i += NT) {
      const float4 q = boxes[i];
      mx = fmaxf(mx, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
    for (int w = 1; w < NT / 32; ++w) mx = fmaxf(mx, s_red[w]);
    offset_unit = __fadd_rn(mx, 1.0f);
  }
  if (tid == 0) s_nkept = 0;
  __syncthreads();

  // ---- 3. greedy scan in sorted order, at most max_dets survivors ----------------------------
  for (int base = 0; base < M; base += NT) {
    const int nk0 = s_nkept;
    if (nk0 >= a.max_dets) break;
    const int i = base + tid;
    bool alive = i < M;
    BoxA me = {};
    int my_idx = -1;
    if (alive) {
      my_idx = (int)(0xFFFFFFFFu - (uint32_t)(s_keys[i] & 0xFFFFFFFFull));
      const float4 q = boxes[my_idx];
      me.cls = classes[my_idx];
      const float off = trick ? __fmul_rn((float)me.cls, offset_unit) : 0.f;
      me.x1 = trick ? __fadd_rn(q.x, off) : q.x;
      me.y1 = trick ? __fadd_rn(q.y, off) : q.y;
      me.x2 = trick ? __fadd_rn(q.z, off) : q.z;
      me.y2 = trick ? __fadd_rn(q.w, off) : q.w;
      me.area = box_area(me.x1, me.y1, me.x2, me.y2);
      for (int k = 0; k < nk0 && alive; ++k)
        if ((trick || s_kept[k].cls == me.cls) && nms_suppresses(s_kept[k], me, a.nms_thresh)) alive = false;
    }
    s_alive[tid] = alive;
    int cursor = 0;
    __syncthreads();
    while (true) {
      if (tid == 0) s_first = NT;
      __syncthreads();
      if (s_alive[tid] && tid >= cursor) atomicMin(&s_first, tid);
      __syncthreads();
      const int first = *reinterpret_cast<volatile int*>(&s_first);   // scalar load: never fused with s_nkept
      if (first >= NT) break;
      if (tid == first) {
        const int slot = s_nkept;
        s_kept[slot] = me;
        s_kept_idx[slot] = my_idx;
        s_nkept = slot + 1;
      }
      __syncthreads();
      if (s_nkept >= a.max_dets) break;
      if (tid > first && s_alive[tid]) {
        const BoxA kb = s_kept[s_nkept - 1];
        if ((trick || kb.cls == me.cls) && nms_suppresses(kb, me, a.nms_thresh)) s_alive[tid] = 0;
      }
      cursor = first + 1;
      __syncthreads();
    }
    __syncthreads();
  }
  __syncthreads();
  const int nk = s_nkept;
  if (tid < nk) keep_out[tid] = s_kept_idx[tid];
  if (tid == 0) a.keep_count[b] = nk;

  // ---- 4. per survivor: gather (standard) or cluster + fuse (BayesOD) ------------------------
  // results staged in det_* at position d (pre-filter); boxes/scores/classes also in smem for step 5
  for (int d = warp; d < nk; d += NT / 32) {
    const int c = s_kept_idx[d];
    const float4 cb = boxes[c];
    float out_box[4] = {cb.x, cb.y, cb.z, cb.w};
    double out_cov[16];
    if (a.has_cov) {
      for (int e = 0; e < 16; ++e) out_cov[e] = (double)covs[(int64_t)c * 16 + e];
    } else {
      for (int e = 0; e < 16; ++e) out_cov[e] = 0.0;
    }
    float out_score = scores[c];
    int out_cls = classes[c];
    if (a.mode == 1) {
      const float area_c = box_area(cb.x, cb.y, cb.z, cb.w);
      const int cls_c = classes[c];
      const float aff = (float)a.affinity;
      double sP[16], sPm[4];
      for (int e = 0; e < 16; ++e) sP[e] = 0.0;
      for (int e = 0; e < 4; ++e) sPm[e] = 0.0;
      int n_same = 0, n_all = 0;
      for (int j = lane; j < M; j += 32) {
        const float4 q = boxes[j];
        const float iou = d2_iou(cb, area_c, q, box_area(q.x, q.y, q.z, q.w));
        if (!(iou > aff)) continue;
        ++n_all;
        if (classes[j] != cls_c) continue;
        ++n_same;
        double S[16], P[16];
        for (int e = 0; e < 16; ++e) S[e] = (double)covs[(int64_t)j * 16 + e];
        inv4(S, P);
        const double mu[4] = {(double)q.x, (double)q.y, (double)q.z, (double)q.w};
        for (int r = 0; r < 4; ++r) {
          double acc = 0.0;
          for (int e = 0; e < 4; ++e) {
            sP[r * 4 + e] += P[r * 4 + e];
            acc += P[r * 4 + e] * mu[e];
          }
          sPm[r] += acc;
        }
      }
      for (int e = 0; e < 16; ++e) sP[e] = warp_sum_d(sP[e]);
      for (int e = 0; e < 4; ++e) sPm[e] = warp_sum_d(sPm[e]);
      n_same = __reduce_add_sync(0xffffffffu, n_same);
      n_all = __reduce_add_sync(0xffffffffu, n_all);
      if (n_same > 0) {
        if (a.box_merge == 0) {
          double F[16];
          inv4(sP, F);
          for (int r = 0; r < 4; ++r) {
            double acc = 0.0;
            for (int e = 0; e < 4; ++e) acc += F[r * 4 + e] * sPm[e];
            out_box[r] = (float)acc;
          }
          for (int e = 0; e < 16; ++e) out_cov[e] = F[e];
        } else {
          // covariance intersection: omega_j = (det(T) - det(T - P_j) + det(P_j)) / (n det(T) + sum_j(det(P_j) - det(T - P_j)))
          const double detT = det4(sP);
          double wP[16], wPm[4], den = 0.0;
          for (int e = 0; e < 16; ++e) wP[e] = 0.0;
          for (int e = 0; e < 4; ++e) wPm[e] = 0.0;
          for (int j = lane; j < M; j += 32) {
            const float4 q = boxes[j];
            const float iou = d2_iou(cb, area_c, q, box_area(q.x, q.y, q.z, q.w));
            if (!(iou > aff) || classes[j] != cls_c) continue;
            double S[16], P[16], Dm[16];
            for (int e = 0; e < 16; ++e) S[e] = (double)covs[(int64_t)j * 16 + e];
            inv4(S, P);
            const double detP = det4(P);
            for (int e = 0; e < 16; ++e) Dm[e] = sP[e] - P[e];
            const double detD = det4(Dm);
            const double num = detT - detD + detP;
            den += detP - detD;
            const double mu[4] = {(double)q.x, (double)q.y, (double)q.z, (double)q.w};
            for (int r = 0; r < 4; ++r) {
              double acc = 0.0;
              for (int e = 0; e < 4; ++e) {
                wP[r * 4 + e] += num * P[r * 4 + e];
                acc += P[r * 4 + e] * mu[e];
              }
              wPm[r] += num * acc;
            }
          }
          for (int e = 0; e < 16; ++e) wP[e] = warp_sum_d(wP[e]);
          for (int e = 0; e < 4; ++e) wPm[e] = warp_sum_d(wPm[e]);
          den = warp_sum_d(den) + (double)n_same * detT;
          for (int e = 0; e < 16; ++e) wP[e] /= den;
          for (int e = 0; e < 4; ++e) wPm[e] /= den;
          double F[16];
          inv4(wP, F);
          for (int r = 0; r < 4; ++r) {
            double acc = 0.0;
            for (int e = 0; e < 4; ++e) acc += F[r * 4 + e] * wPm[e];
            out_box[r] = (float)acc;
          }
          for (int e = 0; e < 16; ++e) out_cov[e] = F[e];
        }
      }
      if (a.cls_merge == 1 && n_all > 0) {
        // mean of the probability vectors of ALL IoU members (any class, reference quirk Q4)
        float best = -1.f;
        int bestk = 0;
        for (int k = 0; k < a.K; ++k) {
          float acc = 0.f;
          for (int j = lane; j < M; j += 32) {
            const float4 q = boxes[j];
            const float iou = d2_iou(cb, area_c, q, box_area(q.x, q.y, q.z, q.w));
            if (iou > aff) acc += probs[(int64_t)j * a.K + k];
          }
          acc = warp_sum_f(acc) / (float)n_all;
          if (lane == 0) det_probs[(int64_t)d * a.K + k] = acc;
          if (acc > best) { best = acc; bestk = k; }
        }
        out_score = best;
        out_cls = bestk;
      } else {
        for (int k = lane; k < a.K; k += 32) det_probs[(int64_t)d * a.K + k] = probs[(int64_t)c * a.K + k];
      }
    } else if (a.mode == 2) {
      // anchor statistics (inference_utils.py:57-162): clusters with >= 2 IoU members take the mean box,
      // sample covariance (+ mean member covariance) and mean probability vector of the SAME-CLASS members
      const float area_c = box_area(cb.x, cb.y, cb.z, cb.w);
      const int cls_c = classes[c];
      const float aff = (float)a.affinity;
      int n_all = 0, n_same = 0;
      double sb[4] = {0.0, 0.0, 0.0, 0.0};
      for (int j = lane; j < M; j += 32) {
        const float4 q = boxes[j];
        if (!(d2_iou(cb, area_c, q, box_area(q.x, q.y, q.z, q.w)) > aff)) continue;
        ++n_all;
        if (classes[j] != cls_c) continue;
        ++n_same;
        sb[0] += q.x; sb[1] += q.y; sb[2] += q.z; sb[3] += q.w;
      }
      n_all = __reduce_add_sync(0xffffffffu, n_all);
      n_same = __reduce_add_sync(0xffffffffu, n_same);
      for (int e = 0; e < 4; ++e) sb[e] = warp_sum_d(sb[e]);
      if (n_all >= 2 && n_same >= 1) {
        float mean[4];
        for (int e = 0; e < 4; ++e) mean[e] = (float)(sb[e] / (double)n_same);
        double cc[16], mc[16];
        for (int e = 0; e < 16; ++e) { cc[e] = 0.0; mc[e] = 0.0; }
        for (int j = lane; j < M; j += 32) {
          const float4 q = boxes[j];
          if (!(d2_iou(cb, area_c, q, box_area(q.x, q.y, q.z, q.w)) > aff) || classes[j] != cls_c) continue;
          const double r[4] = {(double)__fsub_rn(q.x, mean[0]), (double)__fsub_rn(q.y, mean[1]),
                               (double)__fsub_rn(q.z, mean[2]), (double)__fsub_rn(q.w, mean[3])};
          for (int x = 0; x < 4; ++x)
            for (int y = 0; y < 4; ++y) cc[x * 4 + y] += r[x] * r[y];
          if (a.has_cov)
            for (int e = 0; e < 16; ++e) mc[e] += (double)covs[(int64_t)j * 16 + e];
        }
        const double den = n_same - 1 > 1 ? (double)(n_same - 1) : 1.0;
        for (int e = 0; e < 16; ++e) {
          cc[e] = warp_sum_d(cc[e]) / den;
          mc[e] = warp_sum_d(mc[e]) / (double)n_same;
          out_cov[e] = (double)((float)cc[e]) + (a.has_cov ? (double)((float)mc[e]) : 0.0);
        }
        for (int e = 0; e < 4; ++e) out_box[e] = mean[e];
        float best = -1.f;
        int bestk = 0;
        for (int k = 0; k < a.K; ++k) {
          float acc = 0.f;
          for (int j = lane; j < M; j += 32) {
            const float4 q = boxes[j];
            if ((d2_iou(cb, area_c, q, box_area(q.x, q.y, q.z, q.w)) > aff) && classes[j] == cls_c)
              acc += probs[(int64_t)j * a.K + k];
          }
          acc = warp_sum_f(acc) / (float)n_same;
          if (lane == 0) det_probs[(int64_t)d * a.K + k] = acc;
          if (acc > best) { best = acc; bestk = k; }
        }
        out_score = best;
        out_cls = bestk;
      } else {
        if (!a.has_cov)
          for (int e = 0; e < 16; ++e) out_cov[e] = (e % 5 == 0) ? (double)1e-4f : 0.0;
        float best = -1.f;
        int bestk = 0;
        for (int k = 0; k < a.K; ++k) {
          const float v = probs[(int64_t)c * a.K + k];
          if (lane == 0) det_probs[(int64_t)d * a.K + k] = v;
          if (v > best) { best = v; bestk = k; }
        }
        out_score = best;
        out_cls = bestk;
      }
    } else {
      for (int k = lane; k < a.K; k += 32) det_probs[(int64_t)d * a.K + k] = probs[(int64_t)c * a.K + k];
    }
    if (lane == 0) {
      for (int e = 0; e < 4; ++e) s_out_box[d][e] = out_box[e];
      s_out_score[d] = out_score;
      s_out_cls[d] = out_cls;
      for (int e = 0; e < 16; ++e) det_cov[(int64_t)d * 16 + e] = (float)out_cov[e];
    }
  }
  __syncthreads();

  // ---- 5. rescale, clip, drop empty boxes, scale covariance; compact in place ----------------
  // (probabilistic_detector_postprocess; in-place compaction is safe: position p <= d and each
  //  thread reads its own record before any thread writes -- two phases separated by a barrier)
  float bx[4] = {0.f, 0.f, 0.f, 0.f}, cv[16], pr_score = 0.f;
  int pr_cls = 0, flag = 0;
  if (tid < nk && a.skip_post) {
    for (int e = 0; e < 4; ++e) bx[e] = s_out_box[tid][e];
    for (int e = 0; e < 16; ++e) cv[e] = det_cov[(int64_t)tid * 16 + e];
    flag = 1;
    pr_score = s_out_score[tid];
    pr_cls = s_out_cls[tid];
  } else if (tid < nk) {
    bx[0] = __fmul_rn(s_out_box[tid][0], sx);
    bx[1] = __fmul_rn(s_out_box[tid][1], sy);
    bx[2] = __fmul_rn(s_out_box[tid][2], sx);
    bx[3] = __fmul_rn(s_out_box[tid][3], sy);
    bx[0] = fminf(fmaxf(bx[0], 0.f), (float)a.out_w);
    bx[1] = fminf(fmaxf(bx[1], 0.f), (float)a.out_h);
    bx[2] = fminf(fmaxf(bx[2], 0.f), (float)a.out_w);
    bx[3] = fminf(fmaxf(bx[3], 0.f), (float)a.out_h);
    flag = (__fsub_rn(bx[2], bx[0]) > 0.f) && (__fsub_rn(bx[3], bx[1]) > 0.f);
    const float sc[4] = {sx, sy, sx, sy};
    for (int r = 0; r < 4; ++r)
      for (int e = 0; e < 4; ++e) {
        float v = det_cov[(int64_t)tid * 16 + r * 4 + e];
        if (r == e) v = __fadd_rn(v, 1e-4f);
        cv[r * 4 + e] = __fmul_rn(__fmul_rn(sc[r], v), sc[e]);
      }
    pr_score = s_out_score[tid];
    pr_cls = s_out_cls[tid];
  }
  if (tid < MAX_DETS) s_flag[tid] = (tid < nk) ? flag : 0;
  __syncthreads();
  if (tid == 0) {
    int p = 0;
    for (int d = 0; d < nk; ++d) {
      s_pos[d] = p;
      p += s_flag[d];
    }
    a.det_count[b] = p;
  }
  __syncthreads();
  // probability vectors move up at most to an earlier row; process rows in increasing order per k
  for (int k = 0; k < a.K; ++k) {
    float v = 0.f;
    if (tid < nk && flag) v = det_probs[(int64_t)tid * a.K + k];
    __syncthreads();
    if (tid < nk && flag) det_probs[(int64_t)s_pos[tid] * a.K + k] = v;
    __syncthreads();
  }
  if (tid < nk && flag) {
    const int p = s_pos[tid];
    for (int e = 0; e < 4; ++e) det_boxes[p * 4 + e] = bx[e];
    for (int e = 0; e < 16; ++e) det_cov[p * 16 + e] = cv[e];
    det_scores[p] = pr_score;
    det_classes[p] = pr_cls;
    src_out[p] = s_kept_idx[tid];
  }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int pod_nms_fuse(const pod_nms_args* a, void* stream) {
  POD_REQUIRE(a, "pod_nms_fuse: null args");
  POD_REQUIRE(a->boxes && a->scores && a->classes && a->probs && a->count, "pod_nms_fuse: null input");
  POD_REQUIRE(a->has_cov == 0 || a->cov, "pod_nms_fuse: has_cov without cov");
  POD_REQUIRE(a->det_boxes && a->det_cov && a->det_scores && a->det_classes && a->det_probs && a->det_count && a->keep &&
                  a->keep_count && a->det_src, "pod_nms_fuse: null output");
  POD_REQUIRE(a->B > 0 && a->cap > 0 && a->cap <= SORT_MAX, "pod_nms_fuse: cap must be in 1..%d", SORT_MAX);
  POD_REQUIRE(a->K > 0 && a->K <= MAX_K, "pod_nms_fuse: K must be in 1..%d", MAX_K);
  POD_REQUIRE(a->max_dets > 0 && a->max_dets <= MAX_DETS, "pod_nms_fuse: max_dets must be in 1..%d", MAX_DETS);
  POD_REQUIRE(a->mode == 0 || a->mode == 2 || (a->mode == 1 && a->has_cov), "pod_nms_fuse: mode must be 0, 1 (needs covariances) or 2");
  POD_REQUIRE(a->nms_variant >= 0 && a->nms_variant <= 2 && a->box_merge >= 0 && a->box_merge <= 1 && a->cls_merge >= 0 &&
                  a->cls_merge <= 1, "pod_nms_fuse: bad mode flags");
  POD_REQUIRE(a->in_h > 0 && a->in_w > 0 && a->out_h > 0 && a->out_w > 0, "pod_nms_fuse: bad image sizes");
  int npow2 = 1;
  while (npow2 < a->cap) npow2 <<= 1;
  const size_t smem = (size_t)npow2 * sizeof(unsigned long long);
  static bool configured = false;
  if (!configured) {
    POD_CUDA(cudaFuncSetAttribute(k_nms_fuse, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_MAX * 8));
    configured = true;
  }
  const float sx = (float)((double)a->out_w / (double)a->in_w);
  const float sy = (float)((double)a->out_h / (double)a->in_h);
  k_nms_fuse<<<a->B, NT, smem, (cudaStream_t)stream>>>(*a, sx, sy);
  POD_LAUNCH_CHECK();
  return 0;
}
