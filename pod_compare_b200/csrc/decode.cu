// Candidate decode: anchor deltas -> boxes with aleatoric (1000-draw Monte-Carlo) and epistemic
// (over MC-dropout samples / ensemble members) covariance.  One warp per candidate; draws are
// generated in-register from the Philox stream (one pass, shifted moments), nothing of size M x 1000 is
// ever materialised.
// Replaces /root/reference/src/probabilistic_inference/probabilistic_inference.py:310-388,
// inference_utils.py:337-371 (compute_mean_covariance_torch), :510-547 (apply_samples_deltas) and
// probabilistic_modeling/modeling_utils.py:4-22 (covariance_output_to_cholesky).
#include "common.cuh"

namespace {

constexpr int MAX_LEVELS = 8;
struct SegTable {
  int n_levels;
  int seg[MAX_LEVELS + 1];
};

struct Anchor {
  float w, h, cx, cy;
};

__device__ __forceinline__ Anchor anchor_of(const float4 a) {
  Anchor r;
  r.w = __fsub_rn(a.z, a.x);
  r.h = __fsub_rn(a.w, a.y);
  r.cx = __fadd_rn(a.x, __fmul_rn(0.5f, r.w));
  r.cy = __fadd_rn(a.y, __fmul_rn(0.5f, r.h));
  return r;
}

// detectron2 Box2BoxTransform.apply_deltas, operation for operation (no FMA contraction)
__device__ __forceinline__ void decode_box(const float d[4], const Anchor& A, float wx, float wy, float ww, float wh,
                                           float clampv, float out[4]) {
  const float dx = __fdiv_rn(d[0], wx), dy = __fdiv_rn(d[1], wy);
  const float dw = fminf(__fdiv_rn(d[2], ww), clampv), dh = fminf(__fdiv_rn(d[3], wh), clampv);
  const float pcx = __fadd_rn(__fmul_rn(dx, A.w), A.cx);
  const float pcy = __fadd_rn(__fmul_rn(dy, A.h), A.cy);
  const float pw = __fmul_rn(expf(dw), A.w);
  const float ph = __fmul_rn(expf(dh), A.h);
  out[0] = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  out[1] = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  out[2] = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
  out[3] = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// mean and unbiased covariance of `count` boxes produced by gen(j, box) for j = lane, lane+32, ...
// One pass over the draws with moments taken about `ref` (the analytic decode of the mean delta, within
// a fraction of a standard deviation of the sample mean), so the fp32 sums stay small and the final
// correction  cov = (S2 - S1 S1^T / n) / (n - 1)  does not cancel.
template <class Gen>
__device__ __forceinline__ void mean_cov(int count, int lane, const float ref[4], Gen gen, float mean[4], float cov[10]) {
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  float c[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = lane; j < count; j += 32) {
    float x[4], r[4];
    gen(j, x);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      r[d] = x[d] - ref[d];
      s[d] += r[d];
    }
    int q = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int bb = a; bb < 4; ++bb) c[q++] += r[a] * r[bb];
  }
  const float inv_n = 1.0f / (float)count;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    s[d] = warp_sum(s[d]);
    mean[d] = ref[d] + s[d] * inv_n;
  }
  int q = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int bb = a; bb < 4; ++bb) {
      const float s2 = warp_sum(c[q]);
      cov[q] = (s2 - s[a] * s[bb] * inv_n) / (float)(count - 1);
      ++q;
    }
}

__global__ void __launch_bounds__(256) k_decode(pod_decode_args a, SegTable st, PhiloxKey key, float clampv) {
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= a.B * a.cap) return;
  const int b = warp / a.cap, slot = warp % a.cap;
  int level = 0;
  while (level + 1 < st.n_levels && slot >= st.seg[level + 1]) ++level;
  const int rank = slot - st.seg[level];
  const int* cnt = a.cand_cnt + b * st.n_levels;
  int before = 0, total = 0;
  for (int l = 0; l < st.n_levels; ++l) {
    if (l < level) before += cnt[l];
    total += cnt[l];
  }
  if (slot == 0 && lane == 0) a.out_count[b] = total;
  if (rank >= cnt[level]) return;
  const int m = before + rank;
  const int gid = a.cand_idx[(int64_t)b * a.cap + slot];
  const int64_t row = (int64_t)b * a.R + gid;

  const Anchor A = anchor_of(reinterpret_cast<const float4*>(a.anchors)[gid]);
  const float4 md4 = reinterpret_cast<const float4*>(a.mean_delta)[row];
  const float md[4] = {md4.x, md4.y, md4.z, md4.w};
  float ref[4];
  decode_box(md, A, a.wx, a.wy, a.ww, a.wh, clampv, ref);

  float box[4] = {ref[0], ref[1], ref[2], ref[3]};
  float cov[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  if (a.mean_regvar) {
    // Cholesky factor: diag sqrt(exp(logvar)); strictly-lower entries (full covariance) in
    // torch.tril_indices(4,4,-1) order (1,0),(2,0),(2,1),(3,0),(3,1),(3,2)
    const float* rv = a.mean_regvar + row * a.cov_dims;
    float L[4][4] = {};
#pragma unroll
    for (int d = 0; d < 4; ++d) L[d][d] = sqrtf(expf(rv[d]));
    if (a.cov_dims > 4) {
      L[1][0] = rv[4]; L[2][0] = rv[5]; L[2][1] = rv[6]; L[3][0] = rv[7]; L[3][1] = rv[8]; L[3][2] = rv[9];
    }
    const bool diag = a.cov_dims <= 4;
    const uint32_t image = (uint32_t)(a.image0 + b / a.runs);
    const uint32_t run = (uint32_t)(b % a.runs);
    auto gen = [&](int j, float x[4]) {
      const uint4 w = philox4x32_10((uint32_t)gid, run, (uint32_t)j, image, key);
      float z[4], d[4];
      pod_box_muller(w.x, w.y, z[0], z[1]);
      pod_box_muller(w.z, w.w, z[2], z[3]);
      if (diag) {
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = __fadd_rn(md[i], __fmul_rn(L[i][i], z[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float acc = 0.f;
#pragma unroll
          for (int e = 0; e <= i; ++e) acc = fmaf(L[i][e], z[e], acc);
          d[i] = __fadd_rn(md[i], acc);
        }
      }
      decode_box(d, A, a.swx, a.swy, a.sww, a.swh, clampv, x);      // SampleBox2BoxTransform weights (RPN)
    };
    mean_cov(a.box_draws, lane, ref, gen, box, cov);
  }
  if (a.sample_delta && a.S > 1) {
    auto gen = [&](int s, float x[4]) {
      const float4 d4 = reinterpret_cast<const float4*>(a.sample_delta)[((int64_t)b * a.S + s) * a.R + gid];
      const float d[4] = {d4.x, d4.y, d4.z, d4.w};
      decode_box(d, A, a.wx, a.wy, a.ww, a.wh, clampv, x);
    };
    float emean[4], ecov[10];
    mean_cov(a.S, lane, ref, gen, emean, ecov);
#pragma unroll
    for (int q = 0; q < 10; ++q) cov[q] = __fadd_rn(cov[q], ecov[q]);
  }

  const int64_t o = (int64_t)b * a.cap + m;
  if (lane == 0) {
    reinterpret_cast<float4*>(a.out_boxes)[o] = make_float4(box[0], box[1], box[2], box[3]);
    float* c = a.out_cov + o * 16;
    int q = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i; j < 4; ++j) {
        c[i * 4 + j] = cov[q];
        c[j * 4 + i] = cov[q];
        ++q;
      }
    a.out_scores[o] = a.score[row];
    a.out_classes[o] = a.cls[row];
    a.out_anchor[o] = gid;
  }
  for (int k = lane; k < a.K; k += 32) a.out_probs[o * a.K + k] = a.probs[row * a.K + k];
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int pod_decode_cov(const pod_decode_args* a, void* stream) {
  POD_REQUIRE(a, "pod_decode_cov: null args");
  POD_REQUIRE(a->mean_delta && a->anchors && a->probs && a->score && a->cls && a->cand_idx && a->cand_cnt &&
                  a->seg_off_host, "pod_decode_cov: null input");
  POD_REQUIRE(a->out_boxes && a->out_cov && a->out_scores && a->out_classes && a->out_probs && a->out_count &&
                  a->out_anchor, "pod_decode_cov: null output");
  POD_REQUIRE(a->B > 0 && a->R > 0 && a->K > 0 && a->n_levels > 0 && a->n_levels <= MAX_LEVELS, "pod_decode_cov: bad shape");
  POD_REQUIRE(!a->mean_regvar || a->cov_dims == 4 || a->cov_dims == 10, "pod_decode_cov: cov_dims must be 4 or 10");
  POD_REQUIRE(!a->mean_regvar || a->box_draws > 1, "pod_decode_cov: box_draws must be > 1");
  POD_REQUIRE(a->runs >= 1, "pod_decode_cov: runs must be >= 1");
  POD_REQUIRE(a->wx > 0 && a->wy > 0 && a->ww > 0 && a->wh > 0, "pod_decode_cov: regression weights must be positive");
  pod_decode_args args = *a;
  if (args.swx == 0.f && args.swy == 0.f && args.sww == 0.f && args.swh == 0.f) {
    args.swx = a->wx; args.swy = a->wy; args.sww = a->ww; args.swh = a->wh;
  }
  POD_REQUIRE(args.swx > 0 && args.swy > 0 && args.sww > 0 && args.swh > 0, "pod_decode_cov: sampled-decode weights must be positive");
  SegTable st;
  st.n_levels = a->n_levels;
  for (int l = 0; l <= a->n_levels; ++l) st.seg[l] = a->seg_off_host[l];
  POD_REQUIRE(st.seg[a->n_levels] == a->cap, "pod_decode_cov: cap must equal seg_off[n_levels]");
  const int64_t warps = (int64_t)a->B * a->cap;
  const int64_t blocks = (warps * 32 + 255) / 256;
  POD_REQUIRE(blocks < (1ll << 31), "pod_decode_cov: launch too large");
  const float clampv = (float)log(1000.0 / 16.0);
  k_decode<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(args, st, pod_key(a->seed, POD_STREAM_BOX), clampv);
  POD_LAUNCH_CHECK();
  return 0;
}
