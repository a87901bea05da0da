// 3x3 head convolution as an implicit GEMM on the Blackwell tensor cores (sm_100a).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * Wt[cout, tap*Cin + cin]
//
// Three kernels share the TMA / mbarrier / tcgen05 plumbing of this file:
//   k_conv3x3_tc2   256 -> 256 tower layers on CTA PAIRS (cta_group::2, M = 2 x 128 pixels, N = 256): 90 % of the step;
//                   pixel tiles of 16 x 8 or, per map shape, 32 x 4 (template TW); optional in-kernel input dropout (MASKA)
//   k_conv3x3_wt    output convolutions of <= 64 channels, WEIGHTS as the A operand ([w_hi; w_lo] stacked along M = 128),
//                   16 x 16 pixels as N = 256
//   k_conv3x3_tc    everything else on a single CTA, pixels as M = 128, N = Cout_pad (hi|lo weights N-stacked for <= 128)
//
// * pixel operand: channels-last activation tiles fetched by TMA from a 4-D tensor map (C, W, H, maps); the zero
//   padding is TMA's out-of-bounds fill.  Row-halo staging (default): one box two rows taller than the tile per column
//   shift serves the three row-shifted taps through shared-memory descriptor offsets of whole 2048-byte (16-pixel) or
//   4096-byte (32-pixel) rows.
// * weight operand: packed rows [Cout_pad][9*Cin] (hi rows, then lo rows, one buffer) fetched by TMA (2-D map).
// * fp32 fidelity on fp16 tensor cores: both operands arrive as (hi, lo) fp16 pairs; hi*hi', hi*lo' and lo*hi' are
//   accumulated in fp32 TMEM.  The K loop is cut into chunks (default 12 K-blocks of 64 channels) summed in alternating
//   TMEM buffers and added in fp32 round-to-nearest by the epilogue warps; the accumulator's truncation bias
//   (0.27 ulp per MMA, measured) is compensated in the epilogue scale.
// * warp roles: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (whole warp converged, one elected lane issues:
//   straight-line UTCHMMA) + TMEM allocator, warps 4-11 = epilogue (tcgen05.ld -> chunk sums -> bias / ReLU / Philox
//   dropout -> fp16 split or fp32 store); setmaxnreg moves registers from warpgroup 0 to the epilogue warpgroups.
// * persistent: grid = min(#tiles, #SMs), static round-robin over (live map, tile_y, tile_x); every mbarrier wait is
//   bounded and reports through pod_status / pod_conv3x3_tc_status instead of hanging the device; activations that leave
//   the fp16 split range are reported the same way.
//
// Replaces nn.Conv2d(+ReLU+Dropout) of the reference head,
// /root/reference/src/probabilistic_modeling/probabilistic_retinanet.py:401-441,458-484,517-523.
#include "common.cuh"
#include <cuda.h>
#include <cstring>
#include <cstdlib>
#include <type_traits>

namespace tc {

constexpr int BM = 128;
constexpr int TILE_H = 8;
constexpr int TILE_W = 16;
constexpr int UMMA_K = 16;
constexpr int ACC_COLS = 256;           // TMEM columns reserved per accumulator buffer
constexpr int NUM_THREADS = 384;         // warpgroup 0: TMA producer (warp 0), MMA issuer + TMEM alloc (warp 1);
                                         // warpgroups 1-2 (warps 4-11): epilogue.  setmaxnreg moves registers from
                                         // warpgroup 0 to the epilogue warps so their 128 running sums stay in registers
constexpr long long WAIT_LIMIT_CYCLES = 4000000000LL;   // ~2 s: bounded waits, never hang the box

__device__ int g_status = 0;            // 0 ok; 1..99 = code of the barrier wait that expired; 100 = a hidden activation
                                        // left the fp16 split range (|x * out_scale| > 65504, would become inf / NaN)
__device__ long long g_wait_limit = WAIT_LIMIT_CYCLES;   // pod_conv3x3_tc_set_wait_limit (tests shorten it)
constexpr int STATUS_SATURATED = 100;

struct Params {
  CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
  int NB, H, W, Cin;
  int map_group, map_live;   // logical map m -> physical map (m / map_live) * map_group + m % map_live
  int tiles_x, tiles_y, num_tiles;
  int Cout, Cout_pad;
  int relu;
  int dbg_skip_ld;      // debug: skip the TMEM drains (results invalid) to attribute chunk overhead
  long long* dbg_clock; // debug (POD_TC_DEBUG_CLOCK): CTA 0 of the pair kernel records {clock64, globaltimer} at start and end
  int dbg_no_rmw;       // debug: Q1 accumulation stores without reading back (results invalid) to attribute its cost
  int dbg_fault;        // fault injection (tests): CTA 0's producer never issues its first load -> bounded waits expire
  const float* in_scale_dev;   // optional device-resident input scale (overrides the host value folded into acc_scale)
  const float* out_scale_dev;  // optional device-resident split scale of the HIDDEN output (overrides out_scale)
  int kb_per_chunk;     // K-blocks summed in the tensor core before an fp32 RN add in registers
  float acc_scale;      // 1 / (in_scale * w_scale)
  float out_scale;
  const float* bias;
  __half* out_hi;
  __half* out_lo;
  float* out_f32;
  long long out_map_stride, out_pixel_stride;
  float* out2_f32;      // optional second RAW destination for channels >= split_c
  int split_c;
  long long out2_map_stride, out2_pixel_stride;
  pod_dropout drop;
  uint32_t drop_thr;
  float drop_scale;
  PhiloxKey key;
  // generalised geometry (pod_conv_tc_general: backbone convolutions on the single-CTA kernel); the head's 3x3 / stride 1
  // launches use taps = 9, ksz = 3, stride = 1, pad = 1, w_row0 = 0, out_ch_stride = Cout_pad, out_ch_off = 0
  int taps, ksz, stride, pad;
  int w_row0;                 // first row of this output-channel block in the packed weight matrix
  int out_ch_stride, out_ch_off;
  const __half* res_hi;       // optional residual (split pair, layout of the HIDDEN output) added before the ReLU
  const __half* res_lo;
  float res_inv_scale;
  // Input masking (CTA-pair kernel, FIRST masked tower layer; template MASKA): the A operand is the mask-independent
  // first-layer activation c1 (one map per IMAGE, already multiplied by 1/(1-p)); the dropout mask of (sample, pass) is
  // applied to the staged tile in shared memory by warps 2-3, so the N x passes masked copies never exist in HBM.
  int mask_in;                // 1: MASKA launch
  int mask_in_layer;          // layer index of the mask stream applied to the input (0)
  int drop_scale_only;        // HIDDEN epilogue: multiply by 1/(1-p) but keep every element (producer of c1)
  // Q1 sample accumulation (CTA-pair kernel, last tower layer): see TileRef / tile_ref2 below
  int q1_mode;          // 0 = plain tile order
  int q1_samples;       // S: MC samples per image; maps of an image are ordered sample-major, pass-minor
  int q1_passes;        // tower passes per sample (1 or 2)
  int q1_live[2];       // samples of pass p that are evaluated at all (S or S - 1, reference quirk Q1)
  int q1_acc_mask;      // bit p: pass p is ACCUMULATED (weighted sum over its samples) instead of written per sample
  int q1_group;         // samples per accumulation group (fixed, independent of the batch: bit-reproducible sums)
  int q1_groups;        // ceil(S / q1_group)
  int q1_num_pairs;     // unit pairs of the launch
  float* q1_acc;        // [image][accumulated pass][group][H*W][Cout] fp32 partial sums
};

// Live-map indirection: the maps of a launch come in groups of `map_group` (one group per image: its samples x
// passes) of which only the first `map_live` are evaluated.  Lets the caller leave out the tower passes of the
// last MC sample / ensemble member whose outputs the reference computes but never reads.
__device__ __forceinline__ int physical_map(const Params& P, int m) {
  return P.map_live == P.map_group ? m : (m / P.map_live) * P.map_group + m % P.map_live;
}

// One unit of work of the CTA-pair kernel.  Plain mode: tile t = 2 * pair + rank.  Q1 mode: the launch is cut into UNITS
// (image, pass, sample group, pixel tile); a CTA walks the samples of its unit one after the other, so that the epilogue
// can keep a running weighted sum of the unit's output tile in the (L2-resident) accumulation buffer -- the reference's
// sample "mean" (2 x0 + x1 + ... + x_{S-2}) / S of probabilistic_inference.py:214-270 commutes with the linear output
// convolutions cls_score / cls_var / bbox_cov, which are then evaluated ONCE per image on the mean activation
// instead of once per sample.
struct TileRef {
  int n;        // physical input / output map (P.NB: none -- the TMA box is entirely out of bounds and reads zeros)
  int r;        // pixel tile within the map
  int ok;       // this CTA has real work in this slot
  int acc;      // >= 0: accumulate into partial-sum map `acc` (Q1 mode); -1: write the split pair per sample
  int first;    // first sample of its accumulation group: store instead of add
  int twice;    // sample 0 enters the reference's sum twice (quirk Q1)
};

__device__ __forceinline__ int q1_unit_len(const Params& P, int u, int units, int tiles_per_map) {
  if (u >= units) return 0;
  const int x = u / tiles_per_map;
  const int grp = x % P.q1_groups, p = (x / P.q1_groups) % P.q1_passes;
  const int left = P.q1_live[p] - grp * P.q1_group;
  return left < 0 ? 0 : (left < P.q1_group ? left : P.q1_group);
}

// slots (sample steps) the pair spends on unit pair `up`
__device__ __forceinline__ int pair_len(const Params& P, int up, int tiles_per_map) {
  if (!P.q1_mode) return 1;
  const int units = (P.NB / (P.q1_samples * P.q1_passes)) * P.q1_passes * P.q1_groups * tiles_per_map;
  const int a = q1_unit_len(P, 2 * up, units, tiles_per_map), b = q1_unit_len(P, 2 * up + 1, units, tiles_per_map);
  return a > b ? a : b;
}

__device__ __forceinline__ TileRef tile_ref2(const Params& P, int up, int rank, int slot, int tiles_per_map) {
  TileRef t;
  t.acc = -1; t.first = 0; t.twice = 0;
  if (!P.q1_mode) {
    const int tile = 2 * up + rank;
    t.ok = tile < P.num_tiles;
    t.n = t.ok ? physical_map(P, tile / tiles_per_map) : P.NB;
    t.r = t.ok ? tile % tiles_per_map : 0;
    return t;
  }
  const int images = P.NB / (P.q1_samples * P.q1_passes);
  const int units = images * P.q1_passes * P.q1_groups * tiles_per_map;
  const int u = 2 * up + rank;
  t.ok = slot < q1_unit_len(P, u, units, tiles_per_map);
  t.n = P.NB; t.r = 0;
  if (!t.ok) return t;
  t.r = u % tiles_per_map;
  int x = u / tiles_per_map;
  const int grp = x % P.q1_groups; x /= P.q1_groups;
  const int p = x % P.q1_passes, b = x / P.q1_passes;
  const int sample = grp * P.q1_group + slot;
  t.n = (b * P.q1_samples + sample) * P.q1_passes + p;
  if ((P.q1_acc_mask >> p) & 1) {
    const int n_acc = __popc(P.q1_acc_mask), a_idx = __popc(P.q1_acc_mask & ((1 << p) - 1));
    t.acc = (b * n_acc + a_idx) * P.q1_groups + grp;
    t.first = slot == 0;
    t.twice = sample == 0;
  }
  return t;
}


// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: returns false (and records `code`) if the barrier did not flip within ~2 s
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int code) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  const long long limit = g_wait_limit;
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > limit) {
      atomicCAS(&g_status, 0, code);
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// One lane of a CONVERGED warp (elect.sync).  The MMA-issuing warp runs its whole loop with warp-uniform control
// flow and predicates every tcgen05.mma / tcgen05.commit on this flag: ptxas then emits straight-line UTCHMMA
// from uniform registers.  Issued from divergent code (`if (lane == 0)`) each tcgen05 instruction is instead
// wrapped in an ELECT / BRA.U.ANY serialisation loop (~100 cycles per MMA, measured: the narrow output
// convolutions were bound by that issue rate, not by the tensor pipe or L2).
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                         uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// one lane polls the mbarrier for the warp (256 spinning threads saturate the barrier unit)
__device__ __forceinline__ bool warp_mbar_wait(uint64_t* bar, uint32_t parity, int code, int lane) {
  int ok = 1;
  if (lane == 0) ok = mbar_wait(bar, parity, code) ? 1 : 0;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
// every lane of a converged warp waits (each observes the phase flip); warp-uniform result
__device__ __forceinline__ bool mbar_wait_all(uint64_t* bar, uint32_t parity, int code) {
  const bool ok = mbar_wait(bar, parity, code);
  return __all_sync(0xffffffffu, ok) != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): rows of ROW_BYTES
// (= one swizzle span), 8-row atoms SBO apart, Blackwell descriptor version 1.
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  static_assert(ROW_BYTES == 128 || ROW_BYTES == 64, "swizzle span");
  constexpr uint64_t SBO = (uint64_t)(8 * ROW_BYTES) >> 4;
  constexpr uint64_t LAYOUT = ROW_BYTES == 128 ? 2 : 4;   // SWIZZLE_128B : SWIZZLE_64B
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (SBO << 32) | (1ull << 46) | (LAYOUT << 61);
}

template <int BN, int BK>
struct Cfg {
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int A_BYTES = BM * ROW_BYTES;
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static_assert(STAGES >= 2, "pipeline too shallow");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 512 == 0, "operand alignment");
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, K-major both, N>>3, M>>4
  static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

// ---------------------------------------------------------------------------------------------------
// Row-halo staging (HALO kernels, K-block 64 only).  The nine shifted 8x16 boxes of a pixel tile overlap:
// for a fixed column shift dx the three row shifts dy read rows y0-1+dy .. y0+6+dy of the same 16 columns.
// One TMA box of 10 rows x 16 columns x 64 channels therefore serves three taps, and because a row of the
// staged box is 16 pixels x 128 B = 2048 B (two whole SWIZZLE_128B atoms) the tap dy is simply the shared-
// memory descriptor advanced by dy*2048 B -- still 1024-B aligned, so the swizzle pattern is untouched.
// Activation bytes into the SM drop from 9 x 128 to 3 x 160 pixel rows per K-block (2.4x less); the
// L2->SM path (~43 B/clk/SM chip-wide cap) is what bounds the narrow output convolutions and sits at ~80 %
// for the paired tower kernel.  Activations and weights get separate rings: an activation stage is consumed
// by three weight stages (dy = 0,1,2).
// ---------------------------------------------------------------------------------------------------
constexpr int HALO_ROWS = TILE_H + 2;
constexpr int HALO_A_HALF = HALO_ROWS * TILE_W * 128;   // one of (hi, lo): 20 KB
constexpr int HALO_A_STAGE = 2 * HALO_A_HALF;
constexpr int HALO_A_STAGES_MAX = 3;
constexpr int HALO_ROW_BYTES = TILE_W * 128;            // 2048 B per staged pixel row
constexpr int HALO_SMEM_BUDGET = 226 * 1024;

template <int BN>
struct CfgH {
  static constexpr int B_HALF = BN * 128;
  static constexpr int B_STAGE = 2 * B_HALF;
  // An activation stage is held for three taps.  Narrow convolutions retire a tap in a few hundred cycles, less
  // than the L2 round trip, so they need a third activation stage; the 256-wide ones (1.5k cycles per tap) do not.
  static constexpr int A_STAGES = BN <= 128 ? 3 : 2;
  static constexpr int B_STAGES_RAW = (HALO_SMEM_BUDGET - 1024 - A_STAGES * HALO_A_STAGE) / B_STAGE;
  static constexpr int B_STAGES = B_STAGES_RAW > 8 ? 8 : B_STAGES_RAW;
  static constexpr int SMEM_BYTES = 1024 + A_STAGES * HALO_A_STAGE + B_STAGES * B_STAGE;
  static_assert(B_STAGES >= 2, "weight ring too shallow");
  static_assert(B_HALF % 1024 == 0, "weight tile must keep 1024-B (swizzle atom) alignment");
};

// CTA pairs (N = 256 across the pair): every CTA stages 128 weight rows per tap.  TW = tile width in pixels (the tile is
// TW x 128/TW): 16 x 8 is the default; 32 x 4 covers maps whose height is badly divisible by 8 with fewer tiles (160 x 92:
// 115 instead of 120 -- every tile is 27 648 MMA-cycles at 1 kW) at the price of a taller halo share (6 staged rows per 4
// instead of 10 per 8).  A staged row of 32 pixels is 4096 B, so the dy advance stays a whole number of swizzle atoms.
// A_ST = activation stages (3 for the in-kernel input masking, see MASKA).
template <int TW, int A_ST>
struct CfgH2G {
  static_assert(TW == 16 || TW == 32, "tile width");
  static constexpr int TH = 128 / TW;
  static constexpr int ROW_BYTES = TW * 128;
  static constexpr int A_HALF = (TH + 2) * ROW_BYTES;       // one of (hi, lo): 20 KB (16 x 8) or 24 KB (32 x 4)
  static constexpr int A_STAGE = 2 * A_HALF;
  static constexpr int B_HALF = 128 * 128;
  static constexpr int B_STAGE = 2 * B_HALF;
  static constexpr int A_STAGES = A_ST;
  static constexpr int B_STAGES = (HALO_SMEM_BUDGET - 1024 - A_STAGES * A_STAGE) / B_STAGE;
  static constexpr int SMEM_BYTES = 1024 + A_STAGES * A_STAGE + B_STAGES * B_STAGE;
  static_assert(B_STAGES >= 3, "weight ring too shallow");
  static_assert(ROW_BYTES % 1024 == 0 && A_HALF % 1024 == 0, "tap advance must keep the swizzle phase");
};

// bias / ReLU / Philox dropout / store of one pixel row: NG groups of 16 accumulator columns from `sum`
// Dropout keep-bits of one pixel row: bit j of keep[] <-> accumulator column col0 + j.  The Philox calls
// are independent of the accumulators, so the epilogue warps evaluate them in slices while they wait for
// the next accumulation chunk (`slice` of `n_slices`); only a cheap select remains for the tile end.
template <int NG>
__device__ __forceinline__ void dropout_bits_slice(const Params& P, int n, int pixel, int col0, int slice, int n_slices,
                                                   uint32_t (&keep)[(NG * 16 + 31) / 32]) {
  constexpr int CALLS = NG * 2;                        // one Philox call per 8 columns (16-bit lanes, common.cuh)
  const int per = (CALLS + n_slices - 1) / n_slices;
  const int reps = P.drop.samples * P.drop.passes;
  const uint32_t image = (uint32_t)(P.drop.image0 + n / reps);
  const uint32_t sample = (uint32_t)((n / P.drop.passes) % P.drop.samples);
  const uint32_t c1 = pod_dropout_c1(P.drop.level, P.drop.layer, P.drop.tower, P.drop.pass0 + n % P.drop.passes);
#pragma unroll
  for (int q = 0; q < CALLS; ++q) {
    if (q / per == slice) {
      const uint32_t ctr = (uint32_t)(((long long)pixel * P.Cout_pad + col0 + q * 8) >> 3);
      keep[q / 4] |= pod_keep8(philox4x32_10(ctr, c1, sample, image, P.key), P.drop_thr) << ((q % 4) * 8);
    }
  }
}

// bias / ReLU / dropout select / store of one pixel row: NG groups of 16 accumulator columns from `sum`
// Q1 mode (see TileRef): instead of writing this sample's activation, add it to the running weighted sum of the unit's
// output tile (fp32, true units).  Two phases so that the L2 round trips overlap: (1) bias / ReLU / dropout in place in
// the accumulator registers, (2) the 128 partial sums are read back in two batches of sixteen 16-byte loads, added and
// stored (one batch = one L2 latency; a load issued after a store to the same array cannot be hoisted by the compiler,
// which made the one-group-at-a-time form pay eight dependent round trips per tile: +12 % on these launches).  The same
// thread re-reads what it wrote one tile earlier (program order); __ldcg keeps the reads at L2.
template <int NG>
__device__ __forceinline__ void tile_epilogue_acc(const Params& P, float (&sum)[NG * 16], int pixel, int col0,
                                                  const uint32_t (&keep)[(NG * 16 + 31) / 32], int acc_map, int acc_first,
                                                  int acc_twice) {
  static_assert(NG == 8, "accumulating epilogue is written for 128 columns per thread");
  const float acc_scale = P.in_scale_dev != nullptr ? P.acc_scale / __ldg(P.in_scale_dev) : P.acc_scale;
  const float wgt = acc_twice ? 2.f * P.drop_scale : P.drop_scale;         // dropout scale and the Q1 weight of sample 0
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const uint32_t bits = P.drop_thr != 0u ? keep[g / 2] >> ((g % 2) * 16) : 0xFFFFu;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float v = fmaf(sum[g * 16 + i], acc_scale, __ldg(P.bias + col0 + g * 16 + i));
      if (P.relu) v = fmaxf(v, 0.f);
      sum[g * 16 + i] = ((bits >> i) & 1u) ? v * wgt : 0.f;
    }
  }
  float4* q = reinterpret_cast<float4*>(P.q1_acc + ((long long)acc_map * P.H * P.W + pixel) * P.Cout_pad + col0);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4 o[16];
    if (!acc_first && !P.dbg_no_rmw) {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = __ldcg(q + h * 16 + j);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int b = (h * 16 + j) * 4;
      q[h * 16 + j] = make_float4(__fadd_rn(o[j].x, sum[b]), __fadd_rn(o[j].y, sum[b + 1]), __fadd_rn(o[j].z, sum[b + 2]),
                                  __fadd_rn(o[j].w, sum[b + 3]));
    }
  }
}

template <int MODE, int NG>
__device__ __forceinline__ void tile_epilogue(const Params& P, const float (&sum)[NG * 16], int n, int pixel, int col0,
                                              const uint32_t (&keep)[(NG * 16 + 31) / 32]) {
    // a device-resident input scale (power of two, written by pod_feature_scale) divides exactly
    const float acc_scale = P.in_scale_dev != nullptr ? P.acc_scale / __ldg(P.in_scale_dev) : P.acc_scale;
    const float out_scale = P.out_scale_dev != nullptr ? __ldg(P.out_scale_dev) : P.out_scale;
    const float sat_limit = 65504.f / out_scale;
    bool saturated = false;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int ch = col0 + g * 16;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(sum[g * 16 + i], acc_scale, __ldg(P.bias + ch + i));
      if (MODE == POD_OUT_HIDDEN && P.res_hi != nullptr) {
        // residual connection of a bottleneck block (detectron2 BottleneckBlock.forward: out += shortcut; relu)
        const long long ro = ((long long)n * P.H * P.W + pixel) * P.out_ch_stride + P.out_ch_off + ch;
        const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(P.res_hi + ro)), h1 = __ldg(reinterpret_cast<const uint4*>(P.res_hi + ro) + 1);
        const uint4 l0 = __ldg(reinterpret_cast<const uint4*>(P.res_lo + ro)), l1 = __ldg(reinterpret_cast<const uint4*>(P.res_lo + ro) + 1);
        const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
          const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
          v[2 * i] = __fadd_rn(v[2 * i], __fadd_rn(hf.x, lf.x) * P.res_inv_scale);
          v[2 * i + 1] = __fadd_rn(v[2 * i + 1], __fadd_rn(hf.y, lf.y) * P.res_inv_scale);
        }
      }
      if (P.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (MODE == POD_OUT_HIDDEN) {
        if (P.drop_thr != 0u) {
          const uint32_t bits = keep[g / 2] >> ((g % 2) * 16);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = ((bits >> i) & 1u) ? v[i] * P.drop_scale : 0.f;
        } else if (P.drop_scale_only) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= P.drop_scale;
        }
        // the fp16 split pair holds |x * out_scale| <= 65504; beyond it hi becomes inf and every later layer NaN.
        // The reference computes in fp32 and has no such limit, so this is reported, never silent (also catches NaN).
#pragma unroll
        for (int i = 0; i < 16; ++i) saturated |= !(fabsf(v[i]) <= sat_limit);
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __half h0, l0, h1, l1;
          pod_split_h(v[2 * i] * out_scale, h0, l0);
          pod_split_h(v[2 * i + 1] * out_scale, h1, l1);
          ph[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          pl[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        const long long o = ((long long)n * P.H * P.W + pixel) * P.out_ch_stride + P.out_ch_off + ch;
        uint4* dh = reinterpret_cast<uint4*>(P.out_hi + o);
        uint4* dl = reinterpret_cast<uint4*>(P.out_lo + o);
        dh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        dh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
        dl[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        dl[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
      } else {
        float* o = P.out_f32 + (long long)n * P.out_map_stride + (long long)pixel * P.out_pixel_stride + ch;
        if (P.out2_f32 != nullptr) {
          float* o2 = P.out2_f32 + (long long)n * P.out2_map_stride + (long long)pixel * P.out2_pixel_stride + (ch - P.split_c);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (ch + i < P.split_c) o[i] = v[i];
            else if (ch + i < P.Cout) o2[i] = v[i];
          }
        } else if (ch + 16 <= P.Cout && ((P.out_map_stride | P.out_pixel_stride) & 3) == 0) {
          float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) o4[q4] = make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (ch + i < P.Cout) o[i] = v[i];
        }
      }
    }
    if (MODE == POD_OUT_HIDDEN && saturated) atomicCAS(&g_status, 0, STATUS_SATURATED);
}

template <int BN, int BK, int MODE, bool HALO>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv3x3_tc(const __grid_constant__ Params P) {
  using C = Cfg<BN, BK>;
  using CH = CfgH<BN>;
  static_assert(!HALO || BK == 64, "row-halo staging needs 128-byte operand rows");
  constexpr int STAGES = HALO ? CH::B_STAGES : C::STAGES;   // HALO: full/empty_bar track the WEIGHT ring
  // Accumulation is CHUNKED: the tensor core sums one chunk of the K loop (default: one 3x3 tap =
  // Cin channels) into a fresh TMEM accumulator; the epilogue warps add the chunk results in fp32
  // round-to-nearest in registers.  tcgen05 accumulates with truncation, so a single 2304-long chain
  // drifts by ~2e-5 relative; per-tap chains keep the head within ~2e-6 of the fp32 reference.
  // N-stacking (BN <= 128): the hi and lo weight tiles of a stage are adjacent in shared memory, so ONE MMA with
  // N = 2*BN computes hi*hi' into columns [0,BN) and hi*lo' into [BN,2BN); a second MMA (N = BN) adds lo*hi' into
  // [0,BN).  Two instructions and two reads of the 128-row activation operand per K-step instead of three; the
  // narrow convolutions are bound by exactly those shared-memory operand reads.  The epilogue adds the two halves.
  constexpr bool STACK = BN <= 128;
  constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  constexpr int COLS = BN == 256 ? 128 : BN;          // accumulator columns owned by one epilogue thread
  constexpr int EPI_THREADS = BN == 256 ? 256 : 128;  // BN=256: 8 epilogue warps (2 column halves), else 4
  constexpr int NG = COLS / 16;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t afull_bar[HALO_A_STAGES_MAX];   // HALO: activation ring
  __shared__ __align__(8) uint64_t aempty_bar[HALO_A_STAGES_MAX];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem = smem_dyn + (smem_base - smem_u32(smem_dyn));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tm_a_hi);
    tma_prefetch_desc(&P.tm_a_lo);
    tma_prefetch_desc(&P.tm_b_hi);
    tma_prefetch_desc(&P.tm_b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < HALO_A_STAGES_MAX; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_THREADS / 32);        // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int kb_per_tap = P.Cin / BK;
  const int kb_total = P.taps * kb_per_tap;
  const int kb_per_chunk = P.kb_per_chunk;
  const int n_chunks = kb_total / kb_per_chunk;
  const int tiles_per_map = P.tiles_x * P.tiles_y;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ================================ TMA producer ================================
      uint32_t stage = 0, phase = 0;
      bool ok = !(P.dbg_fault && blockIdx.x == 0);     // fault injection: CTA 0 never loads -> its waits expire
      if constexpr (HALO) {
        uint32_t as = 0, aphase = 0;
        for (int tile = blockIdx.x; tile < P.num_tiles && ok; tile += gridDim.x) {
          const int n = physical_map(P, tile / tiles_per_map), r = tile % tiles_per_map;
          const int y0 = (r / P.tiles_x) * TILE_H, x0 = (r % P.tiles_x) * TILE_W;
          for (int cb = 0; cb < kb_per_tap && ok; ++cb) {
            for (int dx = 0; dx < 3 && ok; ++dx) {
              if (!mbar_wait(&aempty_bar[as], aphase ^ 1u, 5)) { ok = false; break; }
              mbar_arrive_expect_tx(&afull_bar[as], (uint32_t)HALO_A_STAGE);
              uint8_t* sa = smem + (size_t)as * HALO_A_STAGE;
              tma_load_4d(&P.tm_a_hi, &afull_bar[as], sa, cb * 64, x0 + dx - 1, y0 - 1, n);
              tma_load_4d(&P.tm_a_lo, &afull_bar[as], sa + HALO_A_HALF, cb * 64, x0 + dx - 1, y0 - 1, n);
              if (++as == CH::A_STAGES) { as = 0; aphase ^= 1u; }
              for (int dy = 0; dy < 3; ++dy) {
                if (!mbar_wait(&empty_bar[stage], phase ^ 1u, 1)) { ok = false; break; }
                mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)CH::B_STAGE);
                uint8_t* sb = smem + CH::A_STAGES * HALO_A_STAGE + (size_t)stage * CH::B_STAGE;
                const int kcol = (dy * 3 + dx) * P.Cin + cb * 64;
                tma_load_2d(&P.tm_b_hi, &full_bar[stage], sb, kcol, 0);
                tma_load_2d(&P.tm_b_lo, &full_bar[stage], sb + CH::B_HALF, kcol, 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
              }
            }
          }
        }
      } else
      for (int tile = blockIdx.x; tile < P.num_tiles && ok; tile += gridDim.x) {
        const int n = physical_map(P, tile / tiles_per_map), r = tile % tiles_per_map;
        const int y0 = (r / P.tiles_x) * TILE_H, x0 = (r % P.tiles_x) * TILE_W;
        for (int tap = 0; tap < P.taps && ok; ++tap) {
          // input coordinates of the tile's first output pixel for this tap; a strided convolution reads every
          // stride-th input pixel through the tensor map's element strides (the box is 16 x 8 loaded pixels either way)
          const int yy = y0 * P.stride + tap / P.ksz - P.pad, xx = x0 * P.stride + tap % P.ksz - P.pad;
          for (int cb = 0; cb < kb_per_tap; ++cb) {
            if (!mbar_wait(&empty_bar[stage], phase ^ 1u, 1)) { ok = false; break; }
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)C::STAGE_BYTES);
            uint8_t* s = smem + (size_t)stage * C::STAGE_BYTES;
            tma_load_4d(&P.tm_a_hi, &full_bar[stage], s, cb * BK, xx, yy, n);
            tma_load_4d(&P.tm_a_lo, &full_bar[stage], s + C::A_BYTES, cb * BK, xx, yy, n);
            tma_load_2d(&P.tm_b_hi, &full_bar[stage], s + 2 * C::A_BYTES, tap * P.Cin + cb * BK, P.w_row0);
            tma_load_2d(&P.tm_b_lo, &full_bar[stage], s + 2 * C::A_BYTES + C::B_BYTES, tap * P.Cin + cb * BK, P.w_row0);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer (whole warp, one elected lane issues) ==============
      const uint32_t leader = elect_one_sync();
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      bool ok = true;
      if constexpr (HALO) {
        uint32_t as = 0, aphase = 0;
        for (int tile = blockIdx.x; tile < P.num_tiles && ok; tile += gridDim.x) {
          int dy = 0;                                   // unit order within a tile: (K-block, dx, dy), dy fastest
          for (int c = 0; c < n_chunks && ok; ++c) {
            if (!mbar_wait_all(&tempty_bar[acc], acc_phase ^ 1u, 2)) { ok = false; break; }
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
            for (int j = 0; j < kb_per_chunk; ++j) {
              if (dy == 0 && !mbar_wait_all(&afull_bar[as], aphase, 6)) { ok = false; break; }
              if (!mbar_wait_all(&full_bar[stage], phase, 3)) { ok = false; break; }
              tcgen05_fence_after();
              const uint32_t sa_hi = smem_base + as * HALO_A_STAGE + dy * HALO_ROW_BYTES;
              const uint32_t sa_lo = sa_hi + HALO_A_HALF;
              const uint32_t sb_hi = smem_base + CH::A_STAGES * HALO_A_STAGE + stage * CH::B_STAGE;
              const uint32_t sb_lo = sb_hi + CH::B_HALF;
              const uint64_t ah0 = make_smem_desc<128>(sa_hi), al0 = make_smem_desc<128>(sa_lo);
              const uint64_t bh0 = make_smem_desc<128>(sb_hi), bl0 = make_smem_desc<128>(sb_lo);
#pragma unroll
              for (int k = 0; k < 64 / UMMA_K; ++k) {
                const uint64_t koff = (uint64_t)(k * UMMA_K * 2) >> 4;   // descriptor address field is in 16-B units
                const uint64_t ah = ah0 + koff, al = al0 + koff, bh = bh0 + koff, bl = bl0 + koff;
                if constexpr (STACK) {
                  umma_f16(d_tmem, ah, bh, IDESC2, (j | k) != 0 ? 1u : 0u, leader);   // [hi*hi' | hi*lo'], N = 2*BN
                  umma_f16(d_tmem, al, bh, C::IDESC, 1u, leader);                     // + lo*hi' into [0,BN)
                } else {
                  umma_f16(d_tmem, al, bh, C::IDESC, (j | k) != 0 ? 1u : 0u, leader);   // small terms first
                  umma_f16(d_tmem, ah, bl, C::IDESC, 1u, leader);
                  umma_f16(d_tmem, ah, bh, C::IDESC, 1u, leader);
                }
              }
              umma_commit(&empty_bar[stage], leader);                          // weight slot free once these MMAs retire
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
              if (++dy == 3) {                                         // third row shift done: activation slot free
                dy = 0;
                umma_commit(&aempty_bar[as], leader);
                if (++as == CH::A_STAGES) { as = 0; aphase ^= 1u; }
              }
              if (j == kb_per_chunk - 1) umma_commit(&tfull_bar[acc], leader);  // chunk complete -> epilogue warps
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
          }
        }
      } else
      for (int tile = blockIdx.x; tile < P.num_tiles && ok; tile += gridDim.x) {
        for (int c = 0; c < n_chunks && ok; ++c) {
          if (!mbar_wait_all(&tempty_bar[acc], acc_phase ^ 1u, 2)) { ok = false; break; }
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
          for (int j = 0; j < kb_per_chunk; ++j) {
            if (!mbar_wait_all(&full_bar[stage], phase, 3)) { ok = false; break; }
            tcgen05_fence_after();
            const uint32_t sa_hi = smem_base + stage * C::STAGE_BYTES;
            const uint32_t sa_lo = sa_hi + C::A_BYTES;
            const uint32_t sb_hi = sa_hi + 2 * C::A_BYTES;
            const uint32_t sb_lo = sb_hi + C::B_BYTES;
            const uint64_t ah0 = make_smem_desc<C::ROW_BYTES>(sa_hi), al0 = make_smem_desc<C::ROW_BYTES>(sa_lo);
            const uint64_t bh0 = make_smem_desc<C::ROW_BYTES>(sb_hi), bl0 = make_smem_desc<C::ROW_BYTES>(sb_lo);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)(k * UMMA_K * 2) >> 4;   // descriptor address field is in 16-B units
              const uint64_t ah = ah0 + koff, al = al0 + koff, bh = bh0 + koff, bl = bl0 + koff;
              if constexpr (STACK) {
                umma_f16(d_tmem, ah, bh, IDESC2, (j | k) != 0 ? 1u : 0u, leader);   // [hi*hi' | hi*lo'], N = 2*BN
                umma_f16(d_tmem, al, bh, C::IDESC, 1u, leader);                     // + lo*hi' into [0,BN)
              } else {
                umma_f16(d_tmem, al, bh, C::IDESC, (j | k) != 0 ? 1u : 0u, leader);   // small terms first
                umma_f16(d_tmem, ah, bl, C::IDESC, 1u, leader);
                umma_f16(d_tmem, ah, bh, C::IDESC, 1u, leader);
              }
            }
            umma_commit(&empty_bar[stage], leader);                          // frees the smem slot when the MMAs retire
            if (j == kb_per_chunk - 1) umma_commit(&tfull_bar[acc], leader);  // chunk complete -> epilogue warps
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          acc ^= 1u;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ================================ epilogue ====================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int quad = warp & 3;               // the TMEM lane quadrant a warp may read is warp % 4
    const int half = (warp - 4) >> 2;        // column half (BN == 256 only)
    if (half == 0 || BN == 256) {
      const int m = quad * 32 + lane;        // GEMM row == TMEM lane == pixel of the tile
      const int col0 = half * COLS;
      uint32_t acc = 0, acc_phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < P.num_tiles && ok; tile += gridDim.x) {
        const int n = physical_map(P, tile / tiles_per_map), r = tile % tiles_per_map;
        const int py = (r / P.tiles_x) * TILE_H + m / TILE_W, px = (r % P.tiles_x) * TILE_W + m % TILE_W;
        const bool valid = py < P.H && px < P.W;
        const int pixel = py * P.W + px;
        float sum[COLS];
        uint32_t keep[(COLS + 31) / 32];
#pragma unroll
        for (int i = 0; i < COLS; ++i) sum[i] = 0.f;
#pragma unroll
        for (int i = 0; i < (COLS + 31) / 32; ++i) keep[i] = 0u;
        for (int c = 0; c < n_chunks; ++c) {
          if (!warp_mbar_wait(&tfull_bar[acc], acc_phase, 4, lane)) { ok = false; break; }
          tcgen05_fence_after();
          const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * ACC_COLS + col0;
          if (!P.dbg_skip_ld) {
            if constexpr (STACK) {
              // columns [0,BN): hi*hi' + lo*hi'; columns [BN,2BN): hi*lo'
#pragma unroll
              for (int g = 0; g < NG; ++g) {
                uint32_t r0[16], r1[16];
                tmem_ld16(trow + g * 16, r0);
                tmem_ld16(trow + BN + g * 16, r1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  sum[g * 16 + i] = __fadd_rn(sum[g * 16 + i], __fadd_rn(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
              }
            } else if (NG % 4 == 0) {
#pragma unroll
              for (int g = 0; g < NG; g += 4) {
                uint32_t r0[32], r1[32];
                tmem_ld32(trow + g * 16, r0);
                tmem_ld32(trow + (g + 2) * 16, r1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[g * 16 + i] = __fadd_rn(sum[g * 16 + i], __uint_as_float(r0[i]));
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[(g + 2) * 16 + i] = __fadd_rn(sum[(g + 2) * 16 + i], __uint_as_float(r1[i]));
              }
            } else {
#pragma unroll
              for (int g = 0; g < NG; ++g) {
                uint32_t r0[16];
                tmem_ld16(trow + g * 16, r0);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) sum[g * 16 + i] = __fadd_rn(sum[g * 16 + i], __uint_as_float(r0[i]));
              }
            }
          }
          // all TMEM reads of this warp are done: one arrival per warp hands the accumulator back
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          acc ^= 1u;
          if (acc == 0) acc_phase ^= 1u;
          if (MODE == POD_OUT_HIDDEN && P.drop_thr != 0u && valid) dropout_bits_slice<NG>(P, n, pixel, col0, c, n_chunks, keep);
        }
        if (!ok || !valid) continue;
        tile_epilogue<MODE, NG>(P, sum, n, pixel, col0, keep);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================
// cta_group::2 variant for the 256->256 convolutions: a cluster of two CTAs (one SM pair) computes two
// adjacent 128-pixel tiles against ONE copy of the weight K-block -- each CTA stages its own activation
// tile and HALF of the weight rows, the leader CTA issues tcgen05.mma.cta_group::2 (M=256, N=256) which
// reads both halves, and each CTA's TMEM receives its own 128 accumulator rows.  Compared with the
// single-CTA kernel this halves the weight traffic into each SM and the shared-memory operand reads
// per MMA, which is what bounds the 1-CTA kernel.
// =====================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t nclusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// same, releasing this CTA's prior writes at cluster scope (the waiter may sit in the other CTA of the pair)
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* tm, uint32_t mbar_cluster_addr, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* tm, uint32_t mbar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                          uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// arrives (once the issued MMAs retire) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3), "r"(leader) : "memory");
}

// MASKA: a third activation stage hides the extra hop (TMA -> mask warps -> MMA) of the in-kernel input masking
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int BK>
struct Cfg2 {
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int A_BYTES = BM * ROW_BYTES;            // this CTA's 128 pixel rows
  static constexpr int B_BYTES = 128 * ROW_BYTES;           // this CTA's half of the 256 weight rows
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  // D=f32, A=B=f16, K-major, N=256, M=256 (pair)
  static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

template <int BK, int MODE, bool HALO, bool MASKA = false, int TW = TILE_W>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) k_conv3x3_tc2(const __grid_constant__ Params P) {
  using C = Cfg2<BK>;
  using CH = CfgH2G<TW, MASKA ? 3 : 2>;                 // each CTA stages 128 of the 256 weight rows
  constexpr int TH = 128 / TW;
  static_assert(TW == TILE_W || (HALO && !MASKA), "the 32 x 4 tile exists for the plain row-halo kernel");
  static_assert(!HALO || BK == 64, "row-halo staging needs 128-byte operand rows");
  static_assert(!MASKA || (HALO && MODE == POD_OUT_HIDDEN), "input masking exists for the row-halo hidden-layer kernel");
  constexpr int STAGES = HALO ? CH::B_STAGES : C::STAGES;   // HALO: full/empty_bar track the WEIGHT ring
  constexpr int COLS = 128, NG = 8, EPI_THREADS = 256;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];    // used in the leader CTA only
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t afull_bar[HALO_A_STAGES_MAX];   // HALO: activation ring (full: leader only)
  __shared__ __align__(8) uint64_t aempty_bar[HALO_A_STAGES_MAX];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];       // used in the leader CTA only
  __shared__ __align__(8) uint64_t amask_bar[HALO_A_STAGES_MAX];   // MASKA: stage masked in BOTH CTAs (leader only)
  __shared__ uint8_t s_keep[MASKA ? 2 * 5 * (TILE_W + 2) * 8 : 8];  // MASKA: keep bits of a K-block's halo region, per mask warp
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem = smem_dyn + (smem_base - smem_u32(smem_dyn));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tm_a_hi);
    tma_prefetch_desc(&P.tm_a_lo);
    tma_prefetch_desc(&P.tm_b_hi);
    tma_prefetch_desc(&P.tm_b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < HALO_A_STAGES_MAX; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EPI_THREADS / 32);  // one arrival per epilogue warp of BOTH CTAs
    }
    for (int i = 0; i < HALO_A_STAGES_MAX; ++i) mbar_init(&amask_bar[i], 4);   // two mask warps in each CTA of the pair
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(&tmem_base_s, 512);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // peer barriers initialised, both TMEM halves allocated
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int kb_per_tap = P.Cin / BK;
  const int kb_total = 9 * kb_per_tap;
  const int kb_per_chunk = P.kb_per_chunk;
  const int n_chunks = kb_total / kb_per_chunk;
  const int tiles_per_map = P.tiles_x * P.tiles_y;
  const int num_pairs = P.q1_mode ? P.q1_num_pairs : (P.num_tiles + 1) >> 1;
  const int pair0 = (int)cluster_id_x(), pair_step = (int)nclusters_x();
  if (P.dbg_clock != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    P.dbg_clock[0] = clock64();
    P.dbg_clock[1] = (long long)gt;
  }

  if (warp < 4) {
  // warpgroup 0 keeps 40 registers (producer, MMA issuer) -- or 56 when its warps 2-3 run the input-mask loops with four
  // Philox chains / four 16-byte chunks in flight.  56 is the ceiling: the kernel launches with 168 registers per thread,
  // the epilogue warps' setmaxnreg.inc to 224 needs (224 - 168) x 256 = 14336 registers from the CTA's pool, and the pool
  // only holds what warpgroup 0 released: (168 - 56) x 128 = 14336.  (64 leaves the pool 1024 short and the inc waits
  // forever -- an unbounded hang, found the hard way.)
  if constexpr (MASKA) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0 && lane == 0) {
    // ================================ TMA producer (both CTAs) ================================
    uint32_t stage = 0, phase = 0;
    [[maybe_unused]] uint32_t as = 0, aphase = 0;
    bool ok = !(P.dbg_fault && pair0 == 0);            // fault injection: pair 0 never loads -> its waits expire
    for (int tp = pair0; tp < num_pairs && ok; tp += pair_step) {
     const int slots = pair_len(P, tp, tiles_per_map);
     for (int slot = 0; slot < slots && ok; ++slot) {
      // a CTA without work in this slot (odd tail, shorter unit) loads map index NB: entirely out of bounds -> zero fill
      const TileRef tr = tile_ref2(P, tp, (int)rank, slot, tiles_per_map);
      const int n = tr.n, r = tr.r;
      const int y0 = (r / P.tiles_x) * TH, x0 = (r % P.tiles_x) * TW;
      if constexpr (HALO) {
        for (int cb = 0; cb < kb_per_tap && ok; ++cb) {
          for (int dx = 0; dx < 3 && ok; ++dx) {
            if (!mbar_wait(&aempty_bar[as], aphase ^ 1u, 15)) { ok = false; break; }
            uint8_t* sa = smem + (size_t)as * CH::A_STAGE;
            if constexpr (MASKA) {
              // the tile of c1 (one map per image) lands under THIS CTA's barrier; its mask warps take it from there
              const int reps = P.drop.samples * P.drop.passes;
              const int img = tr.ok ? n / reps : P.NB / reps;           // no work: map index past the end -> zero fill
              mbar_arrive_expect_tx(&afull_bar[as], (uint32_t)CH::A_STAGE);
              tma_load_4d(&P.tm_a_hi, &afull_bar[as], sa, cb * 64, x0 + dx - 1, y0 - 1, img);
              tma_load_4d(&P.tm_a_lo, &afull_bar[as], sa + CH::A_HALF, cb * 64, x0 + dx - 1, y0 - 1, img);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&afull_bar[as], 2u * (uint32_t)CH::A_STAGE);
              const uint32_t fa = mapa_cluster(smem_u32(&afull_bar[as]), 0);    // the leader's barriers
              tma2_load_4d(&P.tm_a_hi, fa, sa, cb * 64, x0 + dx - 1, y0 - 1, n);
              tma2_load_4d(&P.tm_a_lo, fa, sa + CH::A_HALF, cb * 64, x0 + dx - 1, y0 - 1, n);
            }
            if (++as == CH::A_STAGES) { as = 0; aphase ^= 1u; }
            for (int dy = 0; dy < 3; ++dy) {
              if (!mbar_wait(&empty_bar[stage], phase ^ 1u, 11)) { ok = false; break; }
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (uint32_t)CH::B_STAGE);
              const uint32_t fb = mapa_cluster(smem_u32(&full_bar[stage]), 0);
              uint8_t* sb = smem + CH::A_STAGES * CH::A_STAGE + (size_t)stage * CH::B_STAGE;
              const int kcol = (dy * 3 + dx) * P.Cin + cb * 64;
              tma2_load_2d(&P.tm_b_hi, fb, sb, kcol, (int)rank * 128);
              tma2_load_2d(&P.tm_b_lo, fb, sb + CH::B_HALF, kcol, (int)rank * 128);
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      } else
      for (int tap = 0; tap < 9 && ok; ++tap) {
        const int yy = y0 + tap / 3 - 1, xx = x0 + tap % 3 - 1;
        for (int cb = 0; cb < kb_per_tap; ++cb) {
          if (!mbar_wait(&empty_bar[stage], phase ^ 1u, 11)) { ok = false; break; }
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (uint32_t)C::STAGE_BYTES);
          const uint32_t fb = mapa_cluster(smem_u32(&full_bar[stage]), 0);   // the leader's full barrier
          uint8_t* s = smem + (size_t)stage * C::STAGE_BYTES;
          tma2_load_4d(&P.tm_a_hi, fb, s, cb * BK, xx, yy, n);
          tma2_load_4d(&P.tm_a_lo, fb, s + C::A_BYTES, cb * BK, xx, yy, n);
          tma2_load_2d(&P.tm_b_hi, fb, s + 2 * C::A_BYTES, tap * P.Cin + cb * BK, (int)rank * 128);
          tma2_load_2d(&P.tm_b_lo, fb, s + 2 * C::A_BYTES + C::B_BYTES, tap * P.Cin + cb * BK, (int)rank * 128);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
     }
    }
  } else if (MASKA && (warp == 2 || warp == 3)) {
    // ================================ input-mask warps (both CTAs) =============================
    // Warp 2 owns rows 0-4 of every staged 10-row box, warp 3 rows 5-9.  Per K-block: the keep bits of the halo region
    // (own 5 rows x 18 columns x 8 channel octets; one Philox call each) go into a small table while the TMA is in
    // flight; per column shift dx: wait for the box, zero the dropped fp16 elements of its hi and lo halves in place
    // (the kept ones already carry the 1/(1-p) factor), make the generic-proxy stores visible to the tensor core's
    // async proxy, and tell the leader's MMA warp.
    if constexpr (MASKA) {
      uint32_t as = 0, aphase = 0;
      bool ok = true;
      const int row0 = (warp - 2) * 5;
      uint8_t* tab = s_keep + (warp - 2) * (5 * (TILE_W + 2) * 8);
      const int reps = P.drop.samples * P.drop.passes;
      for (int tp = pair0; tp < num_pairs && ok; tp += pair_step) {
       const int slots = pair_len(P, tp, tiles_per_map);
       for (int slot = 0; slot < slots && ok; ++slot) {
        const TileRef tr = tile_ref2(P, tp, (int)rank, slot, tiles_per_map);
        const int n = tr.n, r = tr.r;
        const int y0 = (r / P.tiles_x) * TH, x0 = (r % P.tiles_x) * TW;
        const uint32_t image = (uint32_t)(P.drop.image0 + n / reps);
        const uint32_t sample = (uint32_t)((n / P.drop.passes) % P.drop.samples);
        const uint32_t c1w = pod_dropout_c1(P.drop.level, P.mask_in_layer, P.drop.tower, P.drop.pass0 + n % P.drop.passes);
        for (int cb = 0; cb < kb_per_tap && ok; ++cb) {
          if (tr.ok) {
            __syncwarp();                                 // the previous K-block's table is no longer read
            // 720 table entries per warp = 22.5 per lane; four independent Philox chains in flight per lane (a single
            // warp cannot hide the ten dependent rounds of one call otherwise)
            constexpr int TAB = 5 * (TILE_W + 2) * 8;
#pragma unroll 1
            for (int it0 = lane; it0 < TAB; it0 += 128) {
              uint32_t ctr[4];
              bool in[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int it = it0 + 32 * u;
                const int g = it & 7, q = (it >> 3) % (TILE_W + 2), rr = (it >> 3) / (TILE_W + 2);
                const int y = y0 - 1 + row0 + rr, x = x0 - 1 + q;
                in[u] = it < TAB && y >= 0 && y < P.H && x >= 0 && x < P.W;
                ctr[u] = (uint32_t)((((long long)y * P.W + x) * P.Cin + cb * 64 + g * 8) >> 3);
              }
              uint4 w[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) w[u] = philox4x32_10(ctr[u], c1w, sample, image, P.key);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (it0 + 32 * u < TAB) tab[it0 + 32 * u] = in[u] ? (uint8_t)pod_keep8(w[u], P.drop_thr) : (uint8_t)0xFF;
            }
            __syncwarp();
          }
          for (int dx = 0; dx < 3 && ok; ++dx) {
            int okw = 1;
            if (lane == 0) okw = mbar_wait(&afull_bar[as], aphase, 17) ? 1 : 0;
            if (!__shfl_sync(0xffffffffu, okw, 0)) { ok = false; break; }
            if (tr.ok) {
              uint8_t* sa = smem + (size_t)as * CH::A_STAGE;
              // 640 16-byte chunks (8 channels) per warp and stage = 20 per lane: AND the hi and the lo chunk with the
              // element mask (16-byte read-modify-write, no branches; four chunks in flight per lane)
#pragma unroll 1
              for (int it0 = lane; it0 < 5 * TILE_W * 8; it0 += 128) {
                uint4 vh[4], vl[4];
                uint32_t kb[4];
                uint4* ph[4];
                uint4* pl[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int it = it0 + 32 * u;                     // 640 = 5 x 128: every lane has exactly five rounds
                  const int g = it & 7, q = (it >> 3) % TILE_W, rr = (it >> 3) / TILE_W;
                  kb[u] = tab[(rr * (TILE_W + 2) + q + dx) * 8 + g];
                  const int pix = (row0 + rr) * TILE_W + q;
                  // SWIZZLE_128B: 16-byte chunk g of pixel row `pix` lives at chunk g ^ (pix % 8)
                  ph[u] = reinterpret_cast<uint4*>(sa + pix * 128 + ((g ^ (pix & 7)) << 4));
                  pl[u] = reinterpret_cast<uint4*>(sa + CH::A_HALF + pix * 128 + ((g ^ (pix & 7)) << 4));
                  vh[u] = *ph[u];
                  vl[u] = *pl[u];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint32_t k = kb[u];
                  const uint32_t m0 = ((k & 1u) ? 0x0000FFFFu : 0u) | ((k & 2u) ? 0xFFFF0000u : 0u);
                  const uint32_t m1 = ((k & 4u) ? 0x0000FFFFu : 0u) | ((k & 8u) ? 0xFFFF0000u : 0u);
                  const uint32_t m2 = ((k & 16u) ? 0x0000FFFFu : 0u) | ((k & 32u) ? 0xFFFF0000u : 0u);
                  const uint32_t m3 = ((k & 64u) ? 0x0000FFFFu : 0u) | ((k & 128u) ? 0xFFFF0000u : 0u);
                  if (k != 0xFFu) {
                    *ph[u] = make_uint4(vh[u].x & m0, vh[u].y & m1, vh[u].z & m2, vh[u].w & m3);
                    *pl[u] = make_uint4(vl[u].x & m0, vl[u].y & m1, vl[u].z & m2, vl[u].w & m3);
                  }
                }
              }
              fence_proxy_async_smem();                   // generic-proxy stores -> visible to tcgen05.mma's operand reads
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_release(mapa_cluster(smem_u32(&amask_bar[as]), 0));
            if (++as == CH::A_STAGES) { as = 0; aphase ^= 1u; }
          }
        }
       }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================================ MMA issuer (leader CTA; whole warp, one elected lane issues) ====
    const uint32_t leader = elect_one_sync();
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    bool ok = true;
    if constexpr (HALO) {
      uint32_t as = 0, aphase = 0;
      for (int tp = pair0; tp < num_pairs && ok; tp += pair_step) {
       const int slots = pair_len(P, tp, tiles_per_map);
       for (int slot = 0; slot < slots && ok; ++slot) {
        int dy = 0;                                     // unit order within a tile: (K-block, dx, dy), dy fastest
        for (int c = 0; c < n_chunks && ok; ++c) {
          if (!mbar_wait_all(&tempty_bar[acc], acc_phase ^ 1u, 12)) { ok = false; break; }
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
          for (int j = 0; j < kb_per_chunk; ++j) {
            if (dy == 0 && !mbar_wait_all(MASKA ? &amask_bar[as] : &afull_bar[as], aphase, 16)) { ok = false; break; }
            if (MASKA && dy == 0) fence_acq_rel_cluster();           // the peer's mask warps released at cluster scope
            if (!mbar_wait_all(&full_bar[stage], phase, 13)) { ok = false; break; }
            tcgen05_fence_after();
            const uint32_t sa_hi = smem_base + as * CH::A_STAGE + dy * CH::ROW_BYTES;
            const uint32_t sa_lo = sa_hi + CH::A_HALF;
            const uint32_t sb_hi = smem_base + CH::A_STAGES * CH::A_STAGE + stage * CH::B_STAGE;
            const uint32_t sb_lo = sb_hi + CH::B_HALF;
            const uint64_t ah0 = make_smem_desc<128>(sa_hi), al0 = make_smem_desc<128>(sa_lo);
            const uint64_t bh0 = make_smem_desc<128>(sb_hi), bl0 = make_smem_desc<128>(sb_lo);
#pragma unroll
            for (int k = 0; k < 64 / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)(k * UMMA_K * 2) >> 4;   // descriptor address field is in 16-B units
              const uint64_t ah = ah0 + koff, al = al0 + koff, bh = bh0 + koff, bl = bl0 + koff;
              umma2_f16(d_tmem, al, bh, C::IDESC, (j | k) != 0 ? 1u : 0u, leader);
              umma2_f16(d_tmem, ah, bl, C::IDESC, 1u, leader);
              umma2_f16(d_tmem, ah, bh, C::IDESC, 1u, leader);
            }
            umma2_commit_mc(&empty_bar[stage], leader);                          // weight slot free in both CTAs
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            if (++dy == 3) {                                             // activation slot free in both CTAs
              dy = 0;
              umma2_commit_mc(&aempty_bar[as], leader);
              if (++as == CH::A_STAGES) { as = 0; aphase ^= 1u; }
            }
            if (j == kb_per_chunk - 1) umma2_commit_mc(&tfull_bar[acc], leader);  // chunk complete in both CTAs
          }
          acc ^= 1u;
          if (acc == 0) acc_phase ^= 1u;
        }
       }
      }
    } else
    for (int tp = pair0; tp < num_pairs && ok; tp += pair_step) {
     const int slots = pair_len(P, tp, tiles_per_map);
     for (int slot = 0; slot < slots && ok; ++slot) {
      for (int c = 0; c < n_chunks && ok; ++c) {
        if (!mbar_wait_all(&tempty_bar[acc], acc_phase ^ 1u, 12)) { ok = false; break; }
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int j = 0; j < kb_per_chunk; ++j) {
          if (!mbar_wait_all(&full_bar[stage], phase, 13)) { ok = false; break; }
          tcgen05_fence_after();
          const uint32_t sa_hi = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sa_lo = sa_hi + C::A_BYTES;
          const uint32_t sb_hi = sa_hi + 2 * C::A_BYTES;
          const uint32_t sb_lo = sb_hi + C::B_BYTES;
          const uint64_t ah0 = make_smem_desc<C::ROW_BYTES>(sa_hi), al0 = make_smem_desc<C::ROW_BYTES>(sa_lo);
          const uint64_t bh0 = make_smem_desc<C::ROW_BYTES>(sb_hi), bl0 = make_smem_desc<C::ROW_BYTES>(sb_lo);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)(k * UMMA_K * 2) >> 4;   // descriptor address field is in 16-B units
            const uint64_t ah = ah0 + koff, al = al0 + koff, bh = bh0 + koff, bl = bl0 + koff;
            umma2_f16(d_tmem, al, bh, C::IDESC, (j | k) != 0 ? 1u : 0u, leader);
            umma2_f16(d_tmem, ah, bl, C::IDESC, 1u, leader);
            umma2_f16(d_tmem, ah, bh, C::IDESC, 1u, leader);
          }
          umma2_commit_mc(&empty_bar[stage], leader);                          // frees the slot in both CTAs
          if (j == kb_per_chunk - 1) umma2_commit_mc(&tfull_bar[acc], leader);  // chunk complete in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
     }
    }
  }
  } else {
    // ================================ epilogue (both CTAs) ====================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int m = quad * 32 + lane;
    const int col0 = half * COLS;
    uint32_t acc = 0, acc_phase = 0;
    bool ok = true;
    const uint32_t te0 = mapa_cluster(smem_u32(&tempty_bar[0]), 0), te1 = mapa_cluster(smem_u32(&tempty_bar[1]), 0);
    for (int tp = pair0; tp < num_pairs && ok; tp += pair_step) {
     const int slots = pair_len(P, tp, tiles_per_map);
     for (int slot = 0; slot < slots && ok; ++slot) {
      const TileRef tr = tile_ref2(P, tp, (int)rank, slot, tiles_per_map);
      const bool tile_ok = tr.ok != 0;
      const int n = tile_ok ? tr.n : 0, r = tr.r;
      const int py = (r / P.tiles_x) * TH + m / TW, px = (r % P.tiles_x) * TW + m % TW;
      const bool valid = tile_ok && py < P.H && px < P.W;
      const int pixel = py * P.W + px;
      float sum[COLS];
      uint32_t keep[(COLS + 31) / 32];
#pragma unroll
      for (int i = 0; i < COLS; ++i) sum[i] = 0.f;
#pragma unroll
      for (int i = 0; i < (COLS + 31) / 32; ++i) keep[i] = 0u;
      for (int c = 0; c < n_chunks; ++c) {
        if (!warp_mbar_wait(&tfull_bar[acc], acc_phase, 14, lane)) { ok = false; break; }
        tcgen05_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * ACC_COLS + col0;
        if (!P.dbg_skip_ld) {
#pragma unroll
          for (int g = 0; g < NG; g += 4) {
            uint32_t r0[32], r1[32];
            tmem_ld32(trow + g * 16, r0);
            tmem_ld32(trow + (g + 2) * 16, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) sum[g * 16 + i] = __fadd_rn(sum[g * 16 + i], __uint_as_float(r0[i]));
#pragma unroll
            for (int i = 0; i < 32; ++i) sum[(g + 2) * 16 + i] = __fadd_rn(sum[(g + 2) * 16 + i], __uint_as_float(r1[i]));
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? te1 : te0);   // the leader's MMA thread tracks both CTAs
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
        if (MODE == POD_OUT_HIDDEN && P.drop_thr != 0u && valid) dropout_bits_slice<NG>(P, n, pixel, col0, c, n_chunks, keep);
      }
      if (!ok || !valid) continue;
      if (MODE == POD_OUT_HIDDEN && tr.acc >= 0) tile_epilogue_acc<NG>(P, sum, pixel, col0, keep, tr.acc, tr.first, tr.twice);
      else tile_epilogue<MODE, NG>(P, sum, n, pixel, col0, keep);
     }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (P.dbg_clock != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    P.dbg_clock[2] = clock64();
    P.dbg_clock[3] = (long long)gt;
  }
  cluster_sync_all();                                   // nobody exits while the peer may still signal it
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

static int encode_act(CUtensorMap* tm, const void* base, int Cin, int W, int H, int NB, long long map_stride_elems, int BK,
                      int box_rows, int box_cols = TILE_W, int stride = 1) {
  PFN_tmapEncodeTiled enc = get_encode();
  POD_REQUIRE(enc, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
  cuuint64_t gstr[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)map_stride_elems * 2};
  // a strided convolution loads every stride-th pixel: the box spans box * stride source pixels (traversal stride)
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(box_cols * stride), (cuuint32_t)(box_rows * stride), 1};
  cuuint32_t est[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, est,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  POD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: CUresult %d", (int)r);
  return 0;
}

static int encode_wt(CUtensorMap* tm, const void* base, int Ktot, int rows, int BK, int BN) {
  PFN_tmapEncodeTiled enc = get_encode();
  POD_REQUIRE(enc, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)Ktot, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
  cuuint32_t est[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, est,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  POD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: CUresult %d", (int)r);
  return 0;
}

template <int BN, int BK, int MODE, bool HALO>
static int launch(const Params& P, cudaStream_t st) {
  constexpr int SMEM = HALO ? CfgH<BN>::SMEM_BYTES : Cfg<BN, BK>::SMEM_BYTES;
  auto kern = k_conv3x3_tc<BN, BK, MODE, HALO>;
  static bool configured = false;
  if (!configured) {
    POD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const int grid = P.num_tiles < pod_num_sms() ? P.num_tiles : pod_num_sms();
  kern<<<grid, NUM_THREADS, SMEM, st>>>(P);
  POD_LAUNCH_CHECK();
  return 0;
}

template <int BK, int MODE, bool HALO, bool MASKA = false, int TW = TILE_W>
static int launch2(const Params& P, cudaStream_t st) {
  constexpr int SMEM = HALO ? CfgH2G<TW, MASKA ? 3 : 2>::SMEM_BYTES : Cfg2<BK>::SMEM_BYTES;
  auto kern = k_conv3x3_tc2<BK, MODE, HALO, MASKA, TW>;
  static bool configured = false;
  if (!configured) {
    POD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const int pairs = P.q1_mode ? P.q1_num_pairs : (P.num_tiles + 1) / 2;
  const int max_pairs = pod_num_sms() / 2;
  const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
  kern<<<grid, NUM_THREADS, SMEM, st>>>(P);
  POD_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// Weights-as-A variant for the narrow output convolutions (Cout <= 64, RAW mode).
//
// With pixels as the M operand a 63-channel head is an N=64 GEMM: every MMA re-reads the 4 KB activation operand
// for 32 cycles of math and the tensor pipe idles on shared-memory operand fetch.  Here the roles
// are swapped: A = the weight tile with its hi and lo parts STACKED along M (rows 0-63 = w_hi, 64-127 = w_lo; the
// two packed halves are adjacent in global memory, so one TMA box fetches both), B = 256 pixels (16 x 16) of one of
// (x_hi, x_lo).  Two full-size MMAs (M=128, N=256) per K-step
//     D[0:64]   += w_hi . x_hi^T + w_hi . x_lo^T
//     D[64:128] += w_lo . x_hi^T + w_lo . x_lo^T          (the lo.lo term is 2^-22, kept for free)
// evaluate 256 pixels; the epilogue adds the two lane halves through shared memory once per tile.  D rows are
// output channels and columns are pixels, so a warp stores 32 consecutive channels of one pixel: coalesced in the
// permuted (pixel, A*K) output layout.
// Rings: weight stages (16 KB) are consumed by two pixel stages (x_hi, x_lo; 32 KB each).
// =====================================================================================================
constexpr int WT_TILE = 16;                        // 16 x 16 pixel tile
constexpr int WT_X_STAGE = 256 * 128;              // one of (hi, lo): 256 pixels x 64 channels
constexpr int WT_W_STAGE = 128 * 128;              // [w_hi (64 rows); w_lo (64 rows)] x 64 channels
constexpr int WT_XH_STAGE = (WT_TILE + 2) * WT_TILE * 128;   // row-halo staging: 18 rows x 16 pixels serve three taps
constexpr int WT_X_STAGES = 4;
constexpr int WT_W_STAGES = 3;
constexpr int WT_XCHG = 2 * 64 * 64 * 4;           // lane-half exchange: 2 column halves x 64 columns x 64 rows fp32
constexpr int WT_SMEM = 1024 + WT_X_STAGES * WT_XH_STAGE + WT_W_STAGES * WT_W_STAGE + WT_XCHG;
static_assert(WT_SMEM <= 227 * 1024, "weights-as-A kernel exceeds shared memory");
static_assert(WT_XH_STAGE % 1024 == 0 && WT_X_STAGES % 2 == 0, "pixel stages keep swizzle-atom alignment, hi/lo pairs");

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// HALO: one 18-row x 16-column box per column shift dx (and per hi / lo part) serves the three row shifts dy through
// descriptor offsets of dy * 2048 B (as in the pixels-as-M HALO kernels): the pixel bytes into the SM drop 2.7x, which
// matters here because this kernel pulls 65 B/clk/SM through L2 (ncu: lts 61 %, a quarter of the issue time waiting
// for operands).  Pixel stages then come in (hi, lo) pairs that are held for three weight stages.
template <bool HALO>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_conv3x3_wt(const __grid_constant__ Params P) {
  constexpr int XS = HALO ? WT_XH_STAGE : WT_X_STAGE;          // bytes of one pixel stage
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr int COLS = 128, NG = 8, EPI_THREADS = 256;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t xfull_bar[WT_X_STAGES];
  __shared__ __align__(8) uint64_t xempty_bar[WT_X_STAGES];
  __shared__ __align__(8) uint64_t wfull_bar[WT_W_STAGES];
  __shared__ __align__(8) uint64_t wempty_bar[WT_W_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem = smem_dyn + (smem_base - smem_u32(smem_dyn));
  constexpr uint32_t W_OFF = WT_X_STAGES * WT_XH_STAGE;
  constexpr uint32_t XCHG_OFF = W_OFF + WT_W_STAGES * WT_W_STAGE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tm_a_hi);
    tma_prefetch_desc(&P.tm_a_lo);
    tma_prefetch_desc(&P.tm_b_hi);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < WT_X_STAGES; ++i) {
      mbar_init(&xfull_bar[i], 1);
      mbar_init(&xempty_bar[i], 1);
    }
    for (int i = 0; i < WT_W_STAGES; ++i) {
      mbar_init(&wfull_bar[i], 1);
      mbar_init(&wempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_THREADS / 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int kb_per_tap = P.Cin / 64;
  const int kb_total = 9 * kb_per_tap;
  const int kb_per_chunk = P.kb_per_chunk;
  const int n_chunks = kb_total / kb_per_chunk;
  const int tiles_per_map = P.tiles_x * P.tiles_y;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ================================ TMA producer ================================
      uint32_t xs = 0, xphase = 0, ws = 0, wphase = 0;
      bool ok = !(P.dbg_fault && blockIdx.x == 0);     // fault injection: CTA 0 never loads -> its waits expire
      for (int tile = blockIdx.x; tile < P.num_tiles && ok; tile += gridDim.x) {
        const int n = physical_map(P, tile / tiles_per_map), r = tile % tiles_per_map;
        const int y0 = (r / P.tiles_x) * WT_TILE, x0 = (r % P.tiles_x) * WT_TILE;
        if constexpr (HALO) {
          for (int cb = 0; cb < kb_per_tap && ok; ++cb) {
            for (int dx = 0; dx < 3 && ok; ++dx) {
              for (int part = 0; part < 2; ++part) {
                if (!mbar_wait(&xempty_bar[xs], xphase ^ 1u, 22)) { ok = false; break; }
                mbar_arrive_expect_tx(&xfull_bar[xs], (uint32_t)XS);
                tma_load_4d(part == 0 ? &P.tm_a_hi : &P.tm_a_lo, &xfull_bar[xs], smem + (size_t)xs * XS, cb * 64, x0 + dx - 1, y0 - 1, n);
                if (++xs == WT_X_STAGES) { xs = 0; xphase ^= 1u; }
              }
              for (int dy = 0; dy < 3 && ok; ++dy) {
                if (!mbar_wait(&wempty_bar[ws], wphase ^ 1u, 21)) { ok = false; break; }
                mbar_arrive_expect_tx(&wfull_bar[ws], (uint32_t)WT_W_STAGE);
                tma_load_2d(&P.tm_b_hi, &wfull_bar[ws], smem + W_OFF + (size_t)ws * WT_W_STAGE, (dy * 3 + dx) * P.Cin + cb * 64, 0);
                if (++ws == WT_W_STAGES) { ws = 0; wphase ^= 1u; }
              }
            }
          }
        } else
        for (int tap = 0; tap < 9 && ok; ++tap) {
          const int yy = y0 + tap / 3 - 1, xx = x0 + tap % 3 - 1;
          for (int cb = 0; cb < kb_per_tap && ok; ++cb) {
            if (!mbar_wait(&wempty_bar[ws], wphase ^ 1u, 21)) { ok = false; break; }
            mbar_arrive_expect_tx(&wfull_bar[ws], (uint32_t)WT_W_STAGE);
            tma_load_2d(&P.tm_b_hi, &wfull_bar[ws], smem + W_OFF + (size_t)ws * WT_W_STAGE, tap * P.Cin + cb * 64, 0);
            if (++ws == WT_W_STAGES) { ws = 0; wphase ^= 1u; }
            for (int part = 0; part < 2; ++part) {
              if (!mbar_wait(&xempty_bar[xs], xphase ^ 1u, 22)) { ok = false; break; }
              mbar_arrive_expect_tx(&xfull_bar[xs], (uint32_t)XS);
              tma_load_4d(part == 0 ? &P.tm_a_hi : &P.tm_a_lo, &xfull_bar[xs], smem + (size_t)xs * XS, cb * 64, xx, yy, n);
              if (++xs == WT_X_STAGES) { xs = 0; xphase ^= 1u; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer (whole warp, one elected lane issues) ==============
      const uint32_t leader = elect_one_sync();
      uint32_t xs = 0, xphase = 0, ws = 0, wphase = 0, acc = 0, acc_phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < P.num_tiles && ok; tile += gridDim.x) {
        int dy = 0;                                       // HALO: unit order (K-block, dx, dy), dy fastest
        for (int c = 0; c < n_chunks && ok; ++c) {
          if (!mbar_wait_all(&tempty_bar[acc], acc_phase ^ 1u, 23)) { ok = false; break; }
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
          for (int j = 0; j < kb_per_chunk && ok; ++j) {
            if (!mbar_wait_all(&wfull_bar[ws], wphase, 24)) { ok = false; break; }
            const uint64_t w0 = make_smem_desc<128>(smem_base + W_OFF + ws * WT_W_STAGE);
            if constexpr (HALO) {
              if (dy == 0) {                              // a fresh (hi, lo) pair of pixel stages: xs, xs + 1
                if (!mbar_wait_all(&xfull_bar[xs], xphase, 25)) { ok = false; break; }
                if (!mbar_wait_all(&xfull_bar[xs + 1], xphase, 25)) { ok = false; break; }
              }
              tcgen05_fence_after();
#pragma unroll
              for (int part = 0; part < 2; ++part) {
                const uint64_t x0d = make_smem_desc<128>(smem_base + (xs + part) * XS + dy * (WT_TILE * 128));
#pragma unroll
                for (int k = 0; k < 64 / UMMA_K; ++k) {
                  const uint64_t koff = (uint64_t)(k * UMMA_K * 2) >> 4;
                  umma_f16(d_tmem, w0 + koff, x0d + koff, IDESC, (j | part | k) != 0 ? 1u : 0u, leader);
                }
              }
              if (++dy == 3) {                            // third row shift done: both pixel stages are free
                dy = 0;
                umma_commit(&xempty_bar[xs], leader);
                umma_commit(&xempty_bar[xs + 1], leader);
                xs += 2;
                if (xs == WT_X_STAGES) { xs = 0; xphase ^= 1u; }
              }
            } else {
              for (int part = 0; part < 2; ++part) {
                if (!mbar_wait_all(&xfull_bar[xs], xphase, 25)) { ok = false; break; }
                tcgen05_fence_after();
                const uint64_t x0d = make_smem_desc<128>(smem_base + xs * XS);
#pragma unroll
                for (int k = 0; k < 64 / UMMA_K; ++k) {
                  const uint64_t koff = (uint64_t)(k * UMMA_K * 2) >> 4;
                  umma_f16(d_tmem, w0 + koff, x0d + koff, IDESC, (j | part | k) != 0 ? 1u : 0u, leader);
                }
                umma_commit(&xempty_bar[xs], leader);
                if (++xs == WT_X_STAGES) { xs = 0; xphase ^= 1u; }
              }
            }
            umma_commit(&wempty_bar[ws], leader);
            if (++ws == WT_W_STAGES) { ws = 0; wphase ^= 1u; }
            if (j == kb_per_chunk - 1) umma_commit(&tfull_bar[acc], leader);
          }
          acc ^= 1u;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ================================ epilogue ====================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int quad = warp & 3;               // TMEM lane quadrant: rows quad*32 .. +31 of D
    const int half = (warp - 4) >> 2;        // column (pixel) half
    const int col0 = half * COLS;
    float* xchg = reinterpret_cast<float*>(smem + XCHG_OFF) + half * (64 * 64);
    uint32_t acc = 0, acc_phase = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < P.num_tiles && ok; tile += gridDim.x) {
      const int n = physical_map(P, tile / tiles_per_map), r = tile % tiles_per_map;
      const int y0 = (r / P.tiles_x) * WT_TILE, x0 = (r % P.tiles_x) * WT_TILE;
      float sum[COLS];
#pragma unroll
      for (int i = 0; i < COLS; ++i) sum[i] = 0.f;
      for (int c = 0; c < n_chunks; ++c) {
        if (!warp_mbar_wait(&tfull_bar[acc], acc_phase, 26, lane)) { ok = false; break; }
        tcgen05_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * ACC_COLS + col0;
#pragma unroll
        for (int g = 0; g < NG; g += 4) {
          uint32_t r0[32], r1[32];
          tmem_ld32(trow + g * 16, r0);
          tmem_ld32(trow + (g + 2) * 16, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) sum[g * 16 + i] = __fadd_rn(sum[g * 16 + i], __uint_as_float(r0[i]));
#pragma unroll
          for (int i = 0; i < 32; ++i) sum[(g + 2) * 16 + i] = __fadd_rn(sum[(g + 2) * 16 + i], __uint_as_float(r1[i]));
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (!ok) break;
      // rows 64-127 hold the w_lo products of the same output channels: hand them to rows 0-63, 64 columns a round
#pragma unroll
      for (int rd = 0; rd < 2; ++rd) {
        if (quad >= 2) {
#pragma unroll
          for (int i = 0; i < 64; ++i) xchg[i * 64 + (quad - 2) * 32 + lane] = sum[rd * 64 + i];
        }
        named_bar_sync(1 + half, 128);
        if (quad < 2) {
#pragma unroll
          for (int i = 0; i < 64; ++i) sum[rd * 64 + i] = __fadd_rn(sum[rd * 64 + i], xchg[i * 64 + quad * 32 + lane]);
        }
        named_bar_sync(1 + half, 128);
      }
      const int ch = quad * 32 + lane;
      if (quad < 2 && ch < P.Cout) {
        const float b = __ldg(P.bias + ch);
        const float acc_scale = P.in_scale_dev != nullptr ? P.acc_scale / __ldg(P.in_scale_dev) : P.acc_scale;
        const long long pstride = P.out_pixel_stride;
        // this thread's 128 columns are 8 rows x 16 columns of the pixel tile
        float* o = P.out_f32 + (long long)n * P.out_map_stride + ch +
                   ((long long)(y0 + half * (WT_TILE / 2)) * P.W + x0) * pstride;
        const int rows_ok = P.H - (y0 + half * (WT_TILE / 2)), cols_ok = P.W - x0;
#pragma unroll
        for (int ty = 0; ty < WT_TILE / 2; ++ty) {
          if (ty < rows_ok) {
#pragma unroll
            for (int tx = 0; tx < WT_TILE; ++tx) {
              float v = fmaf(sum[ty * WT_TILE + tx], acc_scale, b);
              if (P.relu) v = fmaxf(v, 0.f);
              if (tx < cols_ok) o[tx * pstride] = v;
            }
          }
          o += (long long)P.W * pstride;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool HALO>
static int launch_wt(const Params& P, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    POD_CUDA(cudaFuncSetAttribute(k_conv3x3_wt<HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
    configured = true;
  }
  const int grid = P.num_tiles < pod_num_sms() ? P.num_tiles : pod_num_sms();
  k_conv3x3_wt<HALO><<<grid, NUM_THREADS, WT_SMEM, st>>>(P);
  POD_LAUNCH_CHECK();
  return 0;
}

static int g_tc_pair = 1;   // 1: 256-channel convs run on CTA pairs (cta_group::2); 0: single-CTA kernel

static int g_tc_tile_w = 0;   // CTA-pair row-halo kernel: 0 = pick per map shape, 16 / 32 = force the tile width

// Tile width of the CTA-pair row-halo kernel for an H x W map: 32 x 4 only where it covers the map with at least 2 % fewer
// tiles -- about 60 % of the saved MMAs show up as time (measured), so small savings are not worth leaving
// the reference geometry.  In practice: P3 of a 1280 x 720 frame as detectron2 pads it (160 x 92: 115 tiles instead of
// 120) and P6 (20 x 12: 3 instead of 4).  An 8 x 16 tile (smallest halo share) was built and measured too: no gain.
static int pick_tile_width(int H, int W) {
  if (g_tc_tile_w == 16 || g_tc_tile_w == 32) return g_tc_tile_w;
  const long long t16 = (long long)((W + 15) / 16) * ((H + 7) / 8);
  const long long t32 = (long long)((W + 31) / 32) * ((H + 3) / 4);
  return t32 * 1000 < t16 * 980 ? 32 : 16;
}

template <int BK, int MODE, bool HALO>
static int dispatch_bn(const Params& P, cudaStream_t st, int tw = TILE_W) {
  if (P.Cout_pad == 256 && g_tc_pair) {
    if constexpr (HALO) {
      if (tw == 32) return launch2<BK, MODE, HALO, false, 32>(P, st);
    }
    return launch2<BK, MODE, HALO>(P, st);
  }
  switch (P.Cout_pad) {
    case 256: return launch<256, BK, MODE, HALO>(P, st);
    case 128: return launch<128, BK, MODE, HALO>(P, st);
    case 96: return launch<96, BK, MODE, HALO>(P, st);
    case 80: return launch<80, BK, MODE, HALO>(P, st);
    case 64: return launch<64, BK, MODE, HALO>(P, st);
    case 48: return launch<48, BK, MODE, HALO>(P, st);
    default: break;
  }
  pod_set_error("pod_conv3x3_tc: unsupported Cout_pad %d (supported: 48,64,80,96,128,256)", P.Cout_pad);
  return -1;
}

}  // namespace tc

static float g_tc_trunc_comp = 0.27f; // expected truncation loss per MMA accumulation, in fp32 ulps of the running sum (see
                                      // pod_conv3x3_tc_set_trunc_comp); measured 0.27 (profiles/r1e_trunc_comp_*.txt); 0 = off
static int g_tc_wt = 1;       // 1: output convolutions of <= 64 channels run weights-as-A (k_conv3x3_wt)
static int g_tc_halo = 3;     // row-halo staging (K-block 64 only), bit 0: pixels-as-M kernels (one 10-row box serves three taps),
                              // bit 1: weights-as-A kernel (18-row boxes).  Both on: with 3 drains per tile the tower is
                              // purely power-bound and the L2->SM bytes saved buy clock (+3 % on the step).  DESIGN.md 3.1a
static int g_tc_bk = 64;      // K-block (channels per pipeline stage): 64 -> SWIZZLE_128B (default), 32 -> SWIZZLE_64B
static int g_tc_taps = 1;     // taps per accumulation chunk (1, 3 or 9)
static int g_tc_chunk_kb = 12; // if > 0: K-blocks per accumulation chunk (must divide 9*Cin/K-block); overrides taps.
                               // default 12 x 64 channels = 3 taps: 3 TMEM drains per tile; with the truncation compensation
                               // this is more accurate than 6-K-block chunks without it and 10 % faster

static int g_tc_fault = 0;    // pod_conv3x3_tc_debug_fault
static long long* g_tc_dbg_clock = nullptr;   // pod_conv3x3_tc_debug_clock: device buffer of 4 words

// Debug / measurement hook: after enabling (out == NULL, enable = 1) CTA 0 of every CTA-pair launch records clock64 and
// globaltimer at its start and end; a later call with out != NULL copies {cycles, nanoseconds} of the LAST such launch
// (synchronises).  cycles / ns = the SM clock the kernel really ran at (NVML's sampled clock hides fine-grained throttling).
extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_debug_clock(int enable, long long* out_cycles_ns_host) {
  if (out_cycles_ns_host != nullptr) {
    POD_REQUIRE(g_tc_dbg_clock != nullptr, "pod_conv3x3_tc_debug_clock: not enabled");
    long long h[4];
    POD_CUDA(cudaMemcpy(h, g_tc_dbg_clock, sizeof(h), cudaMemcpyDeviceToHost));
    out_cycles_ns_host[0] = h[2] - h[0];
    out_cycles_ns_host[1] = h[3] - h[1];
    return 0;
  }
  if (enable && g_tc_dbg_clock == nullptr) {
    POD_CUDA(cudaMalloc(&g_tc_dbg_clock, 4 * sizeof(long long)));
    POD_CUDA(cudaMemset(g_tc_dbg_clock, 0, 4 * sizeof(long long)));
  } else if (!enable && g_tc_dbg_clock != nullptr) {
    POD_CUDA(cudaFree(g_tc_dbg_clock));
    g_tc_dbg_clock = nullptr;
  }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_set_wait_limit(long long cycles) {
  if (cycles <= 0) cycles = tc::WAIT_LIMIT_CYCLES;
  POD_CUDA(cudaMemcpyToSymbol(tc::g_wait_limit, &cycles, sizeof(cycles)));
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_debug_fault(int on) {
  g_tc_fault = on ? 1 : 0;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_set_chunk_kblocks(int kb) {
  POD_REQUIRE(kb >= 0, "pod_conv3x3_tc_set_chunk_kblocks: must be >= 0");
  g_tc_chunk_kb = kb;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_set_trunc_comp(float ulps_per_mma) {
  POD_REQUIRE(ulps_per_mma >= 0.f && ulps_per_mma < 4.f, "pod_conv3x3_tc_set_trunc_comp: 0 <= ulps < 4");
  g_tc_trunc_comp = ulps_per_mma;
  return 0;
}

// tcgen05 adds each MMA into the fp32 TMEM accumulator with truncation toward zero: a chain of n accumulations
// comes out smaller by ~c*n*2^-24 relative (c measured, profiles/).  The factor is folded into the epilogue scale.
static inline float trunc_comp_factor(int kb_per_chunk, int bk, int mmas_per_kstep) {
  const double n = (double)kb_per_chunk * (bk / 16) * mmas_per_kstep;
  return (float)(1.0 + (double)g_tc_trunc_comp * n * 5.9604644775390625e-08);
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_set_wt(int on) {
  g_tc_wt = on ? 1 : 0;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_set_halo(int mode) {
  POD_REQUIRE(mode >= 0 && mode <= 3, "pod_conv3x3_tc_set_halo: bit 0 = pixels-as-M kernels, bit 1 = weights-as-A kernel");
  g_tc_halo = mode;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_set_tile_width(int tw) {
  POD_REQUIRE(tw == 0 || tw == 16 || tw == 32, "pod_conv3x3_tc_set_tile_width: 0 (per map shape), 16 or 32");
  tc::g_tc_tile_w = tw;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_set_pair(int on) {
  tc::g_tc_pair = on ? 1 : 0;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_set_chunk_taps(int taps) {
  POD_REQUIRE(taps == 1 || taps == 3 || taps == 9, "pod_conv3x3_tc_set_chunk_taps: 1, 3 or 9");
  g_tc_taps = taps;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_set_kblock(int bk) {
  POD_REQUIRE(bk == 32 || bk == 64, "pod_conv3x3_tc_set_kblock: 32 or 64");
  g_tc_bk = bk;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc(const pod_conv_args* a, void* stream) {
  using namespace tc;
  POD_REQUIRE(a, "pod_conv3x3_tc: null args");
  POD_REQUIRE(a->in_hi && a->in_lo && a->w_hi && a->w_lo && a->bias, "pod_conv3x3_tc: null operand");
  POD_REQUIRE(a->NB > 0 && a->H > 0 && a->W > 0 && a->Cin >= 64 && a->Cin % 64 == 0, "pod_conv3x3_tc: bad shape (Cin%%64)");
  POD_REQUIRE(a->Cout > 0 && a->Cout <= a->Cout_pad, "pod_conv3x3_tc: Cout > Cout_pad");
  POD_REQUIRE(a->in_map_stride >= (int64_t)a->H * a->W * a->Cin && a->in_map_stride % 8 == 0,
              "pod_conv3x3_tc: in_map_stride too small or not 16-byte aligned");
  POD_REQUIRE(((uintptr_t)a->in_hi | (uintptr_t)a->in_lo | (uintptr_t)a->w_hi | (uintptr_t)a->w_lo) % 16 == 0,
              "pod_conv3x3_tc: operands must be 16-byte aligned");
  POD_REQUIRE((a->in_scale > 0.f || a->in_scale_dev) && a->w_scale > 0.f, "pod_conv3x3_tc: scales must be positive");
  Params P;
  memset(&P, 0, sizeof(P));
  const int BK = g_tc_bk;
  // weights-as-A: RAW mode, <= 64 output channels, single destination, hi|lo weight halves adjacent (stacked along M)
  const bool wt = g_tc_wt && BK == 64 && a->mode == POD_OUT_RAW && a->Cout_pad == 64 && a->out2_f32 == nullptr &&
                  (const char*)a->w_lo == (const char*)a->w_hi + (size_t)64 * 9 * a->Cin * 2;
  if (wt) {
    POD_REQUIRE(a->out_f32 && a->out_pixel_stride >= a->Cout, "pod_conv3x3_tc: raw mode needs out_f32 / pixel stride >= Cout");
    int rc;
    const bool wt_halo = (g_tc_halo & 2) != 0;
    const int wt_rows = wt_halo ? WT_TILE + 2 : WT_TILE;
    if ((rc = encode_act(&P.tm_a_hi, a->in_hi, a->Cin, a->W, a->H, a->NB, a->in_map_stride, 64, wt_rows, WT_TILE))) return rc;
    if ((rc = encode_act(&P.tm_a_lo, a->in_lo, a->Cin, a->W, a->H, a->NB, a->in_map_stride, 64, wt_rows, WT_TILE))) return rc;
    if ((rc = encode_wt(&P.tm_b_hi, a->w_hi, 9 * a->Cin, 128, 64, 128))) return rc;
    P.tm_b_lo = P.tm_b_hi;
    P.NB = a->NB; P.H = a->H; P.W = a->W; P.Cin = a->Cin;
    P.tiles_x = (a->W + WT_TILE - 1) / WT_TILE;
    P.tiles_y = (a->H + WT_TILE - 1) / WT_TILE;
    P.map_group = P.map_live = 1;
    long long logical_maps = a->NB;
    if (a->map_group > 0) {
      POD_REQUIRE(a->NB % a->map_group == 0 && a->map_live > 0 && a->map_live <= a->map_group,
                  "pod_conv3x3_tc: map_group must divide NB and 0 < map_live <= map_group");
      P.map_group = a->map_group;
      P.map_live = a->map_live;
      logical_maps = (long long)(a->NB / a->map_group) * a->map_live;
    }
    const long long nt = (long long)P.tiles_x * P.tiles_y * logical_maps;
    POD_REQUIRE(nt < (1ll << 31), "pod_conv3x3_tc: too many tiles");
    P.num_tiles = (int)nt;
    P.Cout = a->Cout; P.Cout_pad = a->Cout_pad;
    P.relu = a->relu;
    P.kb_per_chunk = g_tc_taps * (a->Cin / 64);
    if (g_tc_chunk_kb > 0 && (9 * (a->Cin / 64)) % g_tc_chunk_kb == 0) P.kb_per_chunk = g_tc_chunk_kb;
    P.in_scale_dev = a->in_scale_dev;
    P.acc_scale = 1.0f / ((a->in_scale_dev ? 1.0f : a->in_scale) * a->w_scale) * trunc_comp_factor(P.kb_per_chunk, 64, 2);
    P.bias = a->bias;
    P.out_f32 = a->out_f32;
    P.out_map_stride = a->out_map_stride; P.out_pixel_stride = a->out_pixel_stride;
    P.dbg_fault = g_tc_fault;
    return wt_halo ? launch_wt<true>(P, (cudaStream_t)stream) : launch_wt<false>(P, (cudaStream_t)stream);
  }
  const bool halo = (g_tc_halo & 1) && BK == 64;
  // tile geometry: 16 x 8 pixels everywhere except where the CTA-pair row-halo kernel does better with 32 x 4
  const int tw = (halo && a->Cout_pad == 256 && tc::g_tc_pair && !a->mask_in) ? pick_tile_width(a->H, a->W) : TILE_W;
  const int th = 128 / tw;
  const int box_rows = halo ? th + 2 : th;
  int rc;
  if ((rc = encode_act(&P.tm_a_hi, a->in_hi, a->Cin, a->W, a->H, a->NB, a->in_map_stride, BK, box_rows, tw))) return rc;
  if ((rc = encode_act(&P.tm_a_lo, a->in_lo, a->Cin, a->W, a->H, a->NB, a->in_map_stride, BK, box_rows, tw))) return rc;
  // CTA pairs stage half of the 256 weight rows each
  const int b_box_rows = (a->Cout_pad == 256 && tc::g_tc_pair) ? 128 : a->Cout_pad;
  if ((rc = encode_wt(&P.tm_b_hi, a->w_hi, 9 * a->Cin, a->Cout_pad, BK, b_box_rows))) return rc;
  if ((rc = encode_wt(&P.tm_b_lo, a->w_lo, 9 * a->Cin, a->Cout_pad, BK, b_box_rows))) return rc;
  P.NB = a->NB; P.H = a->H; P.W = a->W; P.Cin = a->Cin;
  P.taps = 9; P.ksz = 3; P.stride = 1; P.pad = 1; P.w_row0 = 0;
  P.out_ch_stride = a->Cout_pad; P.out_ch_off = 0;
  P.tiles_x = (a->W + tw - 1) / tw;
  P.tiles_y = (a->H + th - 1) / th;
  P.map_group = P.map_live = 1;
  long long logical_maps = a->NB;
  if (a->map_group > 0) {
    POD_REQUIRE(a->NB % a->map_group == 0 && a->map_live > 0 && a->map_live <= a->map_group,
                "pod_conv3x3_tc: map_group must divide NB and 0 < map_live <= map_group");
    P.map_group = a->map_group;
    P.map_live = a->map_live;
    logical_maps = (long long)(a->NB / a->map_group) * a->map_live;
  }
  const long long nt = (long long)P.tiles_x * P.tiles_y * logical_maps;
  POD_REQUIRE(nt < (1ll << 31), "pod_conv3x3_tc: too many tiles");
  P.num_tiles = (int)nt;
  P.Cout = a->Cout; P.Cout_pad = a->Cout_pad;
  P.relu = a->relu;
  P.kb_per_chunk = g_tc_taps * (a->Cin / BK);
  if (g_tc_chunk_kb > 0 && (9 * (a->Cin / BK)) % g_tc_chunk_kb == 0) P.kb_per_chunk = g_tc_chunk_kb;
  P.dbg_skip_ld = getenv("POD_TC_DEBUG_SKIP_LD") ? 1 : 0;
  P.dbg_no_rmw = getenv("POD_TC_DEBUG_NO_RMW") ? 1 : 0;
  P.dbg_clock = g_tc_dbg_clock;
  P.dbg_fault = g_tc_fault;
  P.in_scale_dev = a->in_scale_dev;
  P.acc_scale = 1.0f / ((a->in_scale_dev ? 1.0f : a->in_scale) * a->w_scale) *
                trunc_comp_factor(P.kb_per_chunk, BK, a->Cout_pad <= 128 ? 2 : 3);
  P.out_scale = a->out_scale;
  P.out_scale_dev = a->out_scale_dev;
  P.bias = a->bias;
  P.out_hi = (__half*)a->out_hi; P.out_lo = (__half*)a->out_lo; P.out_f32 = a->out_f32;
  P.out_map_stride = a->out_map_stride; P.out_pixel_stride = a->out_pixel_stride;
  if (a->mode == POD_OUT_RAW && a->out2_f32 != nullptr) {
    P.out2_f32 = a->out2_f32; P.split_c = a->split_col;
    P.out2_map_stride = a->out2_map_stride; P.out2_pixel_stride = a->out2_pixel_stride;
  }
  P.drop = a->drop;
  if (P.drop.samples <= 0) P.drop.samples = 1;
  if (P.drop.passes <= 0) P.drop.passes = 1;
  P.drop_thr = 0;
  P.drop_scale = 1.f;
  if (a->mode == POD_OUT_HIDDEN) {
    POD_REQUIRE(a->out_hi && a->out_lo && (a->out_scale > 0.f || a->out_scale_dev), "pod_conv3x3_tc: hidden mode needs out_hi/out_lo/out_scale");
    POD_REQUIRE(a->Cout == a->Cout_pad, "pod_conv3x3_tc: hidden mode needs Cout == Cout_pad");
    POD_REQUIRE(((uintptr_t)a->out_hi | (uintptr_t)a->out_lo) % 16 == 0, "pod_conv3x3_tc: outputs must be 16-byte aligned");
    if (a->drop.p > 0.0) {
      POD_REQUIRE(a->drop.p < 1.0, "pod_conv3x3_tc: dropout p must be < 1");
      P.drop_thr = pod_dropout_threshold16(a->drop.p);
      if (P.drop_thr == 0u) P.drop_thr = 1u;           // p < 2^-16: keep (almost) everything, but stay in dropout mode
      P.drop_scale = pod_dropout_scale(a->drop.p);
      P.key = pod_key(a->drop.seed, POD_STREAM_DROPOUT);
    }
  } else if (a->mode == POD_OUT_RAW) {
    POD_REQUIRE(a->out_f32 && a->out_pixel_stride > 0, "pod_conv3x3_tc: raw mode needs out_f32 / pixel stride");
    if (a->out2_f32 == nullptr) {
      POD_REQUIRE(a->out_pixel_stride >= a->Cout, "pod_conv3x3_tc: raw mode pixel stride smaller than Cout");
    } else {
      const int n1 = a->split_col < 0 ? 0 : (a->split_col < a->Cout ? a->split_col : a->Cout);
      POD_REQUIRE(a->out_pixel_stride >= n1 && a->out2_pixel_stride >= a->Cout - n1 && a->out2_pixel_stride > 0,
                  "pod_conv3x3_tc: dual-destination pixel strides too small");
    }
  } else {
    POD_REQUIRE(false, "pod_conv3x3_tc: unknown mode %d", a->mode);
  }
  if (a->q1_acc != nullptr) {
    // Q1 sample accumulation (see TileRef): last tower layer of an MC-dropout head, CTA-pair kernel only
    POD_REQUIRE(a->mode == POD_OUT_HIDDEN && a->Cout_pad == 256 && tc::g_tc_pair, "pod_conv3x3_tc: q1_acc needs a 256-channel hidden "
                "convolution on the CTA-pair kernel");
    POD_REQUIRE(a->q1_samples > 1 && (a->q1_passes == 1 || a->q1_passes == 2) && a->NB % (a->q1_samples * a->q1_passes) == 0,
                "pod_conv3x3_tc: q1 maps must come as images x samples x passes");
    POD_REQUIRE(a->q1_group > 0 && a->q1_acc_mask >= 0 && a->q1_acc_mask < (1 << a->q1_passes), "pod_conv3x3_tc: bad q1 group / pass mask");
    POD_REQUIRE(a->q1_live[0] >= 0 && a->q1_live[0] <= a->q1_samples && a->q1_live[1] >= 0 && a->q1_live[1] <= a->q1_samples,
                "pod_conv3x3_tc: q1_live out of range");
    POD_REQUIRE((uintptr_t)a->q1_acc % 16 == 0, "pod_conv3x3_tc: q1_acc must be 16-byte aligned");
    POD_REQUIRE(a->map_group == 0 || a->map_group == a->q1_samples * a->q1_passes, "pod_conv3x3_tc: q1 and map_group disagree");
    P.q1_mode = 1;
    P.q1_samples = a->q1_samples; P.q1_passes = a->q1_passes;
    P.q1_live[0] = a->q1_live[0]; P.q1_live[1] = a->q1_passes > 1 ? a->q1_live[1] : 0;
    P.q1_acc_mask = a->q1_acc_mask;
    P.q1_group = a->q1_group;
    P.q1_groups = (a->q1_samples + a->q1_group - 1) / a->q1_group;
    P.q1_acc = a->q1_acc;
    const long long units = (long long)(a->NB / (a->q1_samples * a->q1_passes)) * a->q1_passes * P.q1_groups * P.tiles_x * P.tiles_y;
    POD_REQUIRE(units < (1ll << 30), "pod_conv3x3_tc: too many q1 units");
    P.q1_num_pairs = (int)((units + 1) / 2);
  }
  cudaStream_t st = (cudaStream_t)stream;
  P.drop_scale_only = 0;
  if (a->drop_scale_only && a->mode == POD_OUT_HIDDEN && a->drop.p > 0.0) {
    // producer of c1 for a mask_in consumer: every element kept, multiplied by 1/(1-p)
    P.drop_thr = 0;
    P.drop_scale_only = 1;
  }
  if (a->mask_in) {
    // first masked tower layer: the dropout mask of its INPUT is applied in shared memory (template MASKA)
    POD_REQUIRE(halo && tc::g_tc_pair && a->mode == POD_OUT_HIDDEN && a->Cout_pad == 256 && a->Cin == 256 && a->drop.p > 0.0 &&
                    a->q1_acc == nullptr, "pod_conv3x3_tc: mask_in needs the row-halo CTA-pair hidden-layer kernel with dropout");
    const int reps = P.drop.samples * P.drop.passes;
    POD_REQUIRE(a->NB % reps == 0, "pod_conv3x3_tc: mask_in maps must be images x samples x passes");
    const int images = a->NB / reps;
    if ((rc = encode_act(&P.tm_a_hi, a->in_hi, a->Cin, a->W, a->H, images, (long long)a->H * a->W * a->Cin, 64, HALO_ROWS))) return rc;
    if ((rc = encode_act(&P.tm_a_lo, a->in_lo, a->Cin, a->W, a->H, images, (long long)a->H * a->W * a->Cin, 64, HALO_ROWS))) return rc;
    P.mask_in = 1;
    P.mask_in_layer = a->mask_in_layer;
    return launch2<64, POD_OUT_HIDDEN, true, true>(P, st);
  }
  if (halo) {
    return a->mode == POD_OUT_HIDDEN ? dispatch_bn<64, POD_OUT_HIDDEN, true>(P, st, tw) : dispatch_bn<64, POD_OUT_RAW, true>(P, st, tw);
  }
  if (BK == 64) {
    return a->mode == POD_OUT_HIDDEN ? dispatch_bn<64, POD_OUT_HIDDEN, false>(P, st)
                                     : dispatch_bn<64, POD_OUT_RAW, false>(P, st);
  }
  return a->mode == POD_OUT_HIDDEN ? dispatch_bn<32, POD_OUT_HIDDEN, false>(P, st) : dispatch_bn<32, POD_OUT_RAW, false>(P, st);
}

// =====================================================================================================
// General convolution entry (ResNet-50-FPN backbone): 1x1 or 3x3, stride 1 or 2, any Cin % 64 == 0,
// output channels in column blocks of 64 / 128 / 256 out of one packed weight matrix, optional residual + ReLU.
// Runs the single-CTA pixels-as-M kernel above (k_conv3x3_tc, per-tap operand staging) with the geometry in Params.
// =====================================================================================================
static int largest_chunk(int kb_total) {
  for (int c = 12; c >= 1; --c)
    if (kb_total % c == 0) return c;
  return 1;
}

extern "C" __attribute__((visibility("default"))) int pod_conv_tc_general(const pod_convg_args* a, void* stream) {
  using namespace tc;
  POD_REQUIRE(a, "pod_conv_tc_general: null args");
  POD_REQUIRE(a->in_hi && a->in_lo && a->w_hi && a->w_lo && a->bias, "pod_conv_tc_general: null operand");
  POD_REQUIRE(a->ksize == 1 || a->ksize == 3, "pod_conv_tc_general: kernel size 1 or 3");
  POD_REQUIRE(a->stride == 1 || a->stride == 2, "pod_conv_tc_general: stride 1 or 2");
  POD_REQUIRE(a->NB > 0 && a->Hin > 0 && a->Win > 0 && a->Cin >= 64 && a->Cin % 64 == 0, "pod_conv_tc_general: bad input shape (Cin%%64)");
  POD_REQUIRE(a->Cout > 0 && a->block_cols > 0 && a->col0 >= 0 && a->col0 + a->block_cols <= a->Cout_rows,
              "pod_conv_tc_general: column block outside the packed weight matrix");
  POD_REQUIRE(a->block_cols == 64 || a->block_cols == 128 || a->block_cols == 256, "pod_conv_tc_general: block of 64, 128 or 256 columns");
  POD_REQUIRE(((uintptr_t)a->in_hi | (uintptr_t)a->in_lo | (uintptr_t)a->w_hi | (uintptr_t)a->w_lo) % 16 == 0,
              "pod_conv_tc_general: operands must be 16-byte aligned");
  POD_REQUIRE(a->in_scale > 0.f && a->w_scale > 0.f, "pod_conv_tc_general: scales must be positive");
  const int pad = a->ksize / 2;
  const int Hout = (a->Hin + 2 * pad - a->ksize) / a->stride + 1, Wout = (a->Win + 2 * pad - a->ksize) / a->stride + 1;
  POD_REQUIRE(Hout == a->Hout && Wout == a->Wout, "pod_conv_tc_general: output size %dx%d expected, got %dx%d", Hout, Wout, a->Hout, a->Wout);
  Params P;
  memset(&P, 0, sizeof(P));
  int rc;
  if ((rc = encode_act(&P.tm_a_hi, a->in_hi, a->Cin, a->Win, a->Hin, a->NB, (long long)a->Hin * a->Win * a->Cin, 64, TILE_H, TILE_W, a->stride))) return rc;
  if ((rc = encode_act(&P.tm_a_lo, a->in_lo, a->Cin, a->Win, a->Hin, a->NB, (long long)a->Hin * a->Win * a->Cin, 64, TILE_H, TILE_W, a->stride))) return rc;
  const int taps = a->ksize * a->ksize;
  if ((rc = encode_wt(&P.tm_b_hi, a->w_hi, taps * a->Cin, a->Cout_rows, 64, a->block_cols))) return rc;
  if ((rc = encode_wt(&P.tm_b_lo, a->w_lo, taps * a->Cin, a->Cout_rows, 64, a->block_cols))) return rc;
  P.NB = a->NB; P.H = Hout; P.W = Wout; P.Cin = a->Cin;
  P.taps = taps; P.ksz = a->ksize; P.stride = a->stride; P.pad = pad; P.w_row0 = a->col0;
  P.tiles_x = (Wout + TILE_W - 1) / TILE_W;
  P.tiles_y = (Hout + TILE_H - 1) / TILE_H;
  P.map_group = P.map_live = 1;
  const long long nt = (long long)P.tiles_x * P.tiles_y * a->NB;
  POD_REQUIRE(nt < (1ll << 31), "pod_conv_tc_general: too many tiles");
  P.num_tiles = (int)nt;
  const int live = a->Cout - a->col0 < a->block_cols ? a->Cout - a->col0 : a->block_cols;   // real channels of this block
  P.Cout = live; P.Cout_pad = a->block_cols;
  P.relu = a->relu;
  P.kb_per_chunk = largest_chunk(taps * (a->Cin / 64));
  P.acc_scale = 1.0f / (a->in_scale * a->w_scale) * trunc_comp_factor(P.kb_per_chunk, 64, a->block_cols <= 128 ? 2 : 3);
  P.bias = a->bias + a->col0;
  P.out_ch_stride = a->out_ch_stride; P.out_ch_off = a->col0;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->out_f32 != nullptr) {
    // fp32 channels-last output (FPN maps handed to the head): element (n, pixel, c) at n*Hout*Wout*stride + pixel*stride + c
    P.out_f32 = a->out_f32 + a->col0;
    P.out_map_stride = (long long)Hout * Wout * a->out_ch_stride; P.out_pixel_stride = a->out_ch_stride;
    switch (a->block_cols) {
      case 256: return launch<256, 64, POD_OUT_RAW, false>(P, st);
      case 128: return launch<128, 64, POD_OUT_RAW, false>(P, st);
      default: return launch<64, 64, POD_OUT_RAW, false>(P, st);
    }
  }
  POD_REQUIRE(a->out_hi && a->out_lo && a->out_scale > 0.f && live == a->block_cols, "pod_conv_tc_general: split output needs out_hi / out_lo / "
              "out_scale and a full column block");
  POD_REQUIRE(((uintptr_t)a->out_hi | (uintptr_t)a->out_lo) % 16 == 0 && a->out_ch_stride % 8 == 0, "pod_conv_tc_general: outputs must be 16-byte aligned");
  P.out_hi = (__half*)a->out_hi; P.out_lo = (__half*)a->out_lo; P.out_scale = a->out_scale;
  if (a->res_hi != nullptr) {
    POD_REQUIRE(a->res_lo && a->res_scale > 0.f && ((uintptr_t)a->res_hi | (uintptr_t)a->res_lo) % 16 == 0, "pod_conv_tc_general: bad residual");
    P.res_hi = (const __half*)a->res_hi; P.res_lo = (const __half*)a->res_lo; P.res_inv_scale = 1.0f / a->res_scale;
  }
  switch (a->block_cols) {
    case 256: return launch<256, 64, POD_OUT_HIDDEN, false>(P, st);
    case 128: return launch<128, 64, POD_OUT_HIDDEN, false>(P, st);
    default: return launch<64, 64, POD_OUT_HIDDEN, false>(P, st);
  }
}

int pod_tc_status_fetch(int* v) {
  *v = 0;
  POD_CUDA(cudaMemcpyFromSymbol(v, tc::g_status, sizeof(int)));
  if (*v != 0) {
    int zero = 0;
    POD_CUDA(cudaMemcpyToSymbol(tc::g_status, &zero, sizeof(int)));
  }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int pod_conv3x3_tc_status(int* status_host) {
  POD_REQUIRE(status_host, "pod_conv3x3_tc_status: null");
  return pod_tc_status_fetch(status_host);
}
