// Wire format of the path: one kernel turns the detections of a whole batch (output of k_nms_fuse) into fixed-size
// records, either as they are (xyxy boxes + covariance: the record the multi-GPU all-gather moves) or in the layout of
// the reference's coco_instances_results.json (XYWH boxes + T Sigma T^T, dataset category ids), so that the host
// writes the JSON from ONE device->host copy instead of five `.cpu().tolist()` synchronisations per image.
// Replaces /root/reference/src/probabilistic_inference/inference_utils.py:428-451 (covar_xyxy_to_xywh) and the
// tensor part of :454-502 (instances_to_json); category mapping semantics of src/apply_net.py:53-79.
#include "common.cuh"

namespace {

// T Sigma T^T with T = [[1,0,0,0],[0,1,0,0],[-1,0,1,0],[0,-1,0,1]], evaluated like the reference's two matmuls:
// A = T Sigma (one rounding per entry: the other products are exact zeros), then A T^T.
__device__ __forceinline__ float xywh_cov_entry(const float* S, int i, int j) {
  auto A = [&](int r, int c) { return r >= 2 ? __fsub_rn(S[r * 4 + c], S[(r - 2) * 4 + c]) : S[r * 4 + c]; };
  return j >= 2 ? __fsub_rn(A(i, j), A(i, j - 2)) : A(i, j);
}

__global__ void __launch_bounds__(256) k_wire_records(pod_wire_args a, int64_t total, int row_w, int rec_w) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / rec_w), j = (int)(t % rec_w);
    const int n = a.det_count[b];
    float v = 0.f;
    if (j == 0) {
      v = (float)n;
    } else {
      const int r = (j - 1) / row_w, c = (j - 1) % row_w;
      if (r < n) {
        const int64_t o = (int64_t)b * a.max_dets + r;
        if (c < 4) {
          const float* bx = a.det_boxes + o * 4;
          v = (a.xywh && c >= 2) ? __fsub_rn(bx[c], bx[c - 2]) : bx[c];       // BoxMode XYXY_ABS -> XYWH_ABS
        } else if (c == 4) {
          v = a.det_scores[o];
        } else if (c == 5) {
          const int cls = a.det_classes[o];
          v = (float)(a.cat_map != nullptr ? ((cls >= 0 && cls < a.K) ? a.cat_map[cls] : -1) : cls);
        } else if (c < 6 + a.K) {
          v = a.det_probs[o * a.K + (c - 6)];
        } else {
          const int e = c - 6 - a.K;
          const float* S = a.det_cov + o * 16;
          v = a.xywh ? xywh_cov_entry(S, e / 4, e % 4) : S[e];
        }
      }
    }
    a.records[t] = v;
  }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int pod_wire_records(const pod_wire_args* a, void* stream) {
  POD_REQUIRE(a, "pod_wire_records: null args");
  POD_REQUIRE(a->det_boxes && a->det_cov && a->det_scores && a->det_classes && a->det_probs && a->det_count && a->records,
              "pod_wire_records: null buffer");
  POD_REQUIRE(a->B > 0 && a->max_dets > 0 && a->K > 0, "pod_wire_records: bad shape");
  const int row_w = 4 + 1 + 1 + a->K + 16;
  const int64_t rec_w = 1 + (int64_t)a->max_dets * row_w;
  POD_REQUIRE(rec_w < (1ll << 31), "pod_wire_records: record too wide");
  const int64_t total = rec_w * a->B;
  const int64_t want = (total + 255) / 256, cap = (int64_t)pod_num_sms() * 8;
  k_wire_records<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(*a, total, row_w, (int)rec_w);
  POD_LAUNCH_CHECK();
  return 0;
}
