// Class scores (logit sampling -> sigmoid -> mean -> max over classes) and per-level top-k.
// Replaces /root/reference/src/probabilistic_inference/probabilistic_inference.py:283-308.
#include "common.cuh"

namespace {

constexpr int MAX_LEVELS = 8;
struct LevelTable {
  int n_levels;
  int off[MAX_LEVELS + 1];
  int seg[MAX_LEVELS + 1];
  int grp_off[MAX_LEVELS + 1];   // prefix sum of ceil(n_l / 4)
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// one thread = 4 consecutive anchors of one level = K Philox quads per draw
__global__ void k_scores(const float* __restrict__ logits, const float* __restrict__ logvar, int B, int R, int K,
                         LevelTable lt, int draws, PhiloxKey key, int image0, int runs, float* __restrict__ probs,
                         float* __restrict__ score, int* __restrict__ cls) {
  const int groups = lt.grp_off[lt.n_levels];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)B * groups;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / groups);
    const int g = (int)(t % groups);
    int level = 0;
    while (level + 1 < lt.n_levels && g >= lt.grp_off[level + 1]) ++level;
    const int a0 = (g - lt.grp_off[level]) * 4;                  // first anchor (within level)
    const int n_l = lt.off[level + 1] - lt.off[level];
    const int64_t n_elem = (int64_t)n_l * K;
    const int64_t base = ((int64_t)b * R + lt.off[level]) * K;   // element 0 of this level
    const int64_t e0 = (int64_t)a0 * K;
    float best[4] = {-1.f, -1.f, -1.f, -1.f};
    int bestk[4] = {0, 0, 0, 0};
    for (int qi = 0; qi < K; ++qi) {
      const int64_t e = e0 + (int64_t)qi * 4;
      if (e >= n_elem) break;
      float mu[4], sg[4], acc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool in = e + i < n_elem;
        mu[i] = in ? logits[base + e + i] : 0.f;
        sg[i] = (in && logvar) ? sqrtf(expf(logvar[base + e + i])) : 0.f;
        acc[i] = 0.f;
      }
      if (logvar) {
        for (int j = 0; j < draws; ++j) {
          const uint4 w = philox4x32_10((uint32_t)(e >> 2), (uint32_t)level | ((uint32_t)(b % runs) << 8), (uint32_t)j,
                                        (uint32_t)(image0 + b / runs), key);
          float z[4];
          pod_box_muller(w.x, w.y, z[0], z[1]);
          pod_box_muller(w.z, w.w, z[2], z[3]);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            acc[i] = __fadd_rn(acc[i], sigmoidf_ref(__fadd_rn(mu[i], __fmul_rn(z[i], sg[i]))));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = __fdiv_rn(acc[i], (float)draws);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = sigmoidf_ref(mu[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (e + i < n_elem) {
          probs[base + e + i] = acc[i];
          const int rel = (int)(e + i - e0);
          const int al = rel / K, k = rel % K;
#pragma unroll
          for (int s = 0; s < 4; ++s)
            if (al == s && acc[i] > best[s]) { best[s] = acc[i]; bestk[s] = k; }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (a0 + s < n_l) {
        const int64_t o = (int64_t)b * R + lt.off[level] + a0 + s;
        score[o] = best[s];
        cls[o] = bestk[s];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per (image, level) top-k: radix select on the unique 64-bit key (score bits << 32 | ~index),
// then a bitonic sort of the <= 1024 selected keys (descending score, ties -> lower index).
// ---------------------------------------------------------------------------------------------
constexpr int TOPK_THREADS = 1024;
constexpr int TOPK_MAX = 1024;

__device__ __forceinline__ unsigned long long make_key(float s, int idx) {
  return ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}

__global__ void __launch_bounds__(TOPK_THREADS)
k_topk(const float* __restrict__ score, int R, LevelTable lt, int topk, float thresh, int cap, int* __restrict__ cand_idx,
       int* __restrict__ cand_cnt) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long sel[TOPK_MAX];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need, s_done, s_count;
  const int level = blockIdx.x, b = blockIdx.y;
  const int off = lt.off[level], n = lt.off[level + 1] - off;
  const float* s = score + (int64_t)b * R + off;
  const int k = topk < n ? topk : n;
  const int tid = threadIdx.x;

  // how many pass the score threshold at all?
  if (tid == 0) s_count = 0;
  __syncthreads();
  int local = 0;
  for (int i = tid; i < n; i += TOPK_THREADS) local += s[i] > thresh;
  local = __reduce_add_sync(0xffffffffu, local);
  if ((tid & 31) == 0 && local) atomicAdd(&s_count, local);
  __syncthreads();
  const int n_gt = s_count;
  __syncthreads();

  // selection predicate: key >> shift >= T >> shift   (or simply score > thresh when few enough pass)
  unsigned long long T = 0;
  int shift = 0;
  const bool by_thresh = n_gt <= k;
  if (!by_thresh) {
    if (tid == 0) { s_prefix = 0ull; s_need = k; s_done = 0; }
    __syncthreads();
    for (int byte = 7; byte >= 0; --byte) {
      for (int i = tid; i < 256; i += TOPK_THREADS) hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const int hs = (byte + 1) * 8;
      for (int i = tid; i < n; i += TOPK_THREADS) {
        const unsigned long long key = make_key(s[i], i);
        if (byte == 7 || (key >> hs) == (prefix >> hs)) atomicAdd(&hist[(unsigned)(key >> (byte * 8)) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        int need = s_need;
        int bkt = 255;
        for (; bkt > 0; --bkt) {
          if ((int)hist[bkt] >= need) break;
          need -= (int)hist[bkt];
        }
        s_prefix = prefix | ((unsigned long long)bkt << (byte * 8));
        s_need = need;
        if ((int)hist[bkt] == need || byte == 0) s_done = byte + 1;   // whole bucket taken: stop refining
      }
      __syncthreads();
      if (s_done) break;
    }
    shift = (s_done - 1) * 8;
    T = s_prefix;
  }
  __syncthreads();
  if (tid == 0) s_count = 0;
  for (int i = tid; i < TOPK_MAX; i += TOPK_THREADS) sel[i] = 0ull;
  __syncthreads();
  for (int i = tid; i < n; i += TOPK_THREADS) {
    const float v = s[i];
    const unsigned long long key = make_key(v, i);
    const bool take = by_thresh ? (v > thresh) : ((key >> shift) >= (T >> shift));
    if (take) {
      const int slot = atomicAdd(&s_count, 1);
      if (slot < TOPK_MAX) sel[slot] = key;
    }
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= TOPK_MAX; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int i = tid, j = i ^ stride;
      if (j > i) {
        const unsigned long long a = sel[i], c = sel[j];
        const bool desc = (i & size) == 0;
        if (desc ? (a < c) : (a > c)) { sel[i] = c; sel[j] = a; }
      }
      __syncthreads();
    }
  }
  const int total = s_count < TOPK_MAX ? s_count : TOPK_MAX;
  // keep those above the score threshold (a prefix of the sorted list)
  int keep = 0;
  if (tid < total) keep = __uint_as_float((uint32_t)(sel[tid] >> 32)) > thresh;
  if (tid == 0) s_need = 0;
  __syncthreads();
  keep = __reduce_add_sync(0xffffffffu, keep);
  if ((tid & 31) == 0 && keep) atomicAdd(&s_need, keep);
  __syncthreads();
  const int cnt = s_need;
  if (tid < cnt) cand_idx[(int64_t)b * cap + lt.seg[level] + tid] = off + (int)(0xFFFFFFFFu - (uint32_t)(sel[tid] & 0xFFFFFFFFull));
  if (tid == 0) cand_cnt[b * lt.n_levels + level] = cnt;
}

int fill_table(LevelTable& lt, int n_levels, const int* level_off, const int* seg_off) {
  POD_REQUIRE(n_levels > 0 && n_levels <= MAX_LEVELS, "level count must be in 1..%d", MAX_LEVELS);
  lt.n_levels = n_levels;
  lt.grp_off[0] = 0;
  for (int l = 0; l <= n_levels; ++l) {
    lt.off[l] = level_off[l];
    lt.seg[l] = seg_off ? seg_off[l] : 0;
    if (l > 0) {
      POD_REQUIRE(level_off[l] > level_off[l - 1], "level offsets must increase");
      lt.grp_off[l] = lt.grp_off[l - 1] + (level_off[l] - level_off[l - 1] + 3) / 4;
    }
  }
  return 0;
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int pod_scores(const float* logits, const float* logvar, int B, int R, int K, int n_levels,
                          const int* level_off, int draws, uint64_t seed, int image0, int runs, float* probs, float* score, int* cls,
                          void* stream) {
  POD_REQUIRE(runs >= 1, "pod_scores: runs must be >= 1");
  POD_REQUIRE(logits && probs && score && cls && level_off && B > 0 && R > 0 && K > 0, "pod_scores: bad args");
  POD_REQUIRE(!logvar || draws > 0, "pod_scores: draws must be positive with logvar");
  LevelTable lt;
  int rc = fill_table(lt, n_levels, level_off, nullptr);
  if (rc) return rc;
  POD_REQUIRE(level_off[0] == 0 && level_off[n_levels] == R, "pod_scores: level offsets must cover [0, R)");
  const int64_t total = (int64_t)B * lt.grp_off[n_levels];
  const int grid = (int)((total + 127) / 128 < (int64_t)pod_num_sms() * 16 ? (total + 127) / 128 : (int64_t)pod_num_sms() * 16);
  k_scores<<<grid, 128, 0, (cudaStream_t)stream>>>(logits, logvar, B, R, K, lt, draws, pod_key(seed, POD_STREAM_LOGIT),
                                                   image0, runs, probs, score, cls);
  POD_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_topk_levels(const float* score, int B, int R, int n_levels, const int* level_off_host,
                               const int* seg_off_host, int topk, float thresh, int* cand_idx, int* cand_cnt,
                               void* stream) {
  POD_REQUIRE(score && cand_idx && cand_cnt && level_off_host && seg_off_host && B > 0 && R > 0, "pod_topk_levels: bad args");
  POD_REQUIRE(topk > 0 && topk <= TOPK_MAX, "pod_topk_levels: topk must be in 1..%d", TOPK_MAX);
  POD_REQUIRE(B <= 65535, "pod_topk_levels: B too large for one launch");
  LevelTable lt;
  int rc = fill_table(lt, n_levels, level_off_host, seg_off_host);
  if (rc) return rc;
  const int cap = seg_off_host[n_levels];
  dim3 grid(n_levels, B);
  k_topk<<<grid, TOPK_THREADS, 0, (cudaStream_t)stream>>>(score, R, lt, topk, thresh, cap, cand_idx, cand_cnt);
  POD_LAUNCH_CHECK();
  return 0;
}
