#!/usr/bin/env python
"""Headline benchmark: images/sec of the probabilistic-inference hot path at MC-dropout N=30 on
1280x720 inputs (BASELINE.json configs[2]: reg_cls_var_dropout + mc_dropout_ensembles pre_nms,
batch 32 per GPU).

A step = one pass of the hot path over one batch: FPN features -> N=30 MC-dropout head loop ->
Q1 means -> scores/top-k -> decode + aleatoric/epistemic covariance -> NMS -> rescale
(reference src/probabilistic_inference/probabilistic_inference.py:178-407).  The ResNet-FPN backbone
is upstream of the rebuilt path and is not part of the step.

  python bench.py --gpus N --steps K --warmup W            (N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference ...                      CPU arm: the reference's own predictor on the host cores
"""
import argparse
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if "reference" in sys.argv and "--impl" in sys.argv:
    # the reference arm is the reference's CPU path: its post-processing picks "cuda if available" at import time
    # (inference_utils.py:9), so the GPUs are hidden from this process before torch initialises
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import torch

METRIC = "images/sec @ MC-dropout N=30, 1280x720"
UNIT = "images/s"
HEIGHT, WIDTH = 720, 1280
N_MC = 30
TOWER_FLOP_PER_LOC = 2 * 9 * 256 * 256          # 1,179,648 FLOP per location per tower conv


_VAR = ["MODEL.PROBABILISTIC_MODELING.CLS_VAR_LOSS.NAME", "loss_attenuation",
        "MODEL.PROBABILISTIC_MODELING.CLS_VAR_LOSS.NUM_SAMPLES", 10,
        "MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS.NAME", "negative_log_likelihood",
        "MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS.COVARIANCE_TYPE", "diagonal"]
# BASELINE.json configs (index) -> (description, cfg options, uses MC-dropout, ensemble members)
WORKLOADS = {
    "mc_pre": ("BASELINE configs[2]: reg_cls_var_dropout head, mc_dropout_ensembles pre_nms",
               _VAR + ["MODEL.PROBABILISTIC_MODELING.DROPOUT_RATE", 0.2,
                       "PROBABILISTIC_INFERENCE.INFERENCE_MODE", "mc_dropout_ensembles",
                       "PROBABILISTIC_INFERENCE.ENSEMBLES_DROPOUT.BOX_MERGE_MODE", "pre_nms"], True, 1),
    "mc_post": ("reg_cls_var_dropout head, mc_dropout_ensembles post_nms",
                _VAR + ["MODEL.PROBABILISTIC_MODELING.DROPOUT_RATE", 0.2,
                        "PROBABILISTIC_INFERENCE.INFERENCE_MODE", "mc_dropout_ensembles",
                        "PROBABILISTIC_INFERENCE.ENSEMBLES_DROPOUT.BOX_MERGE_MODE", "post_nms"], True, 1),
    "bayes_od_mc": ("BASELINE configs[3]: reg_cls_var_dropout head, bayes_od (max_score / bayesian_inference)",
                    _VAR + ["MODEL.PROBABILISTIC_MODELING.DROPOUT_RATE", 0.2,
                            "PROBABILISTIC_INFERENCE.INFERENCE_MODE", "bayes_od",
                            "PROBABILISTIC_INFERENCE.BAYES_OD.CLS_MERGE_MODE", "max_score",
                            "PROBABILISTIC_INFERENCE.BAYES_OD.BOX_MERGE_MODE", "bayesian_inference"], True, 1),
    "loss_att": ("BASELINE configs[1]: reg_cls_var head (loss attenuation), standard_nms, single forward",
                 _VAR + ["PROBABILISTIC_INFERENCE.INFERENCE_MODE", "standard_nms"], False, 1),
    "baseline": ("BASELINE configs[0]: baseline RetinaNet head, standard_nms, single forward",
                 ["PROBABILISTIC_INFERENCE.INFERENCE_MODE", "standard_nms"], False, 1),
    "ensembles5": ("BASELINE configs[4]: reg_cls_var head, 5-member ensembles pre_nms",
                   _VAR + ["PROBABILISTIC_INFERENCE.INFERENCE_MODE", "ensembles",
                           "PROBABILISTIC_INFERENCE.ENSEMBLES.BOX_MERGE_MODE", "pre_nms"], False, 5),
}


def build_cfg(n_mc, workload="mc_pre"):
    from pod_compare_b200.config import get_cfg
    desc, opts, mc, members = WORKLOADS[workload]
    cfg = get_cfg()
    cfg.MODEL.RETINANET.NUM_CLASSES = 7
    cfg.merge_from_list(list(opts) + ["PROBABILISTIC_INFERENCE.AFFINITY_THRESHOLD", 0.9])
    if mc:
        cfg.merge_from_list(["PROBABILISTIC_INFERENCE.MC_DROPOUT.ENABLE", True,
                             "PROBABILISTIC_INFERENCE.MC_DROPOUT.NUM_RUNS", n_mc])
    cfg.SEED = 0
    cfg.freeze()
    return cfg


def head_state_dicts(members, use_dropout, cls_var, bbox_cov):
    """Synthetic weight sets of a workload: one random head, or E correlated members of an ensemble (both arms)."""
    from pod_compare_b200 import synthetic as S
    kw = dict(num_classes=7, use_dropout=use_dropout, cls_var=cls_var, bbox_cov=bbox_cov)
    return S.make_member_state_dicts(members, **kw) if members > 1 else [S.make_head_state_dict(0, **kw)]


def workload_config(args, world):
    desc, _, mc, members = WORKLOADS[args.workload]
    io = ("raw uint8 frames in (ResNet-50-FPN backbone inside the step: %s; frames padded to 736x1280), detections out"
          % ("this repository's kernels" if getattr(args, "backbone", "tc") == "tc" else "torch library fp32 convolutions")
          ) if getattr(args, "from_images", False) else "FPN features in, detections out"
    return {"workload": "%s, %s, batch %d per GPU, 1280x720 (%s)"
                        % (desc, ("N=%d" % args.n_mc) if mc else ("E=%d" % members if members > 1 else "N=1"), args.batch, io),
            "global_batch": args.batch * world, "batch_per_gpu": args.batch, "mc_samples": args.n_mc,
            "image": "%dx%d" % (WIDTH, HEIGHT), "chunk_images": args.chunk, "cuda_graph": bool(getattr(args, "cuda_graph", False)),
            "parallelism": "image-sharded dp%d + NCCL all-gather of detections" % world,
            "l2": "working set per step (tens of GB of activations) far exceeds the 126 MB L2; no flush needed",
            "sample_mean": ("per-sample output convolutions, outputs averaged (as the reference evaluates it)"
                            if (getattr(args, "no_fuse_q1", False) or getattr(args, "keep_unread", False) or not mc) else
                            "sample mean taken of the last tower layer (%s); cls_score / cls_var / bbox_cov run once per image "
                            "(mean of a linear head = head of the mean; fp32 round-off apart, same result)"
                            % ("tcgen05 epilogue accumulation" if getattr(args, "q1_epilogue", False) else "one streaming pass")),
            "unread_outputs": ("evaluated" if getattr(args, "keep_unread", False) else
                               "left out: box_cls / box_cls_var / box_reg_var of the last sample are never read by the "
                               "reference (probabilistic_inference.py:216-267); detections are bit-identical either way")}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
REF_TIME_BUDGET_S = 240.0      # the reference arm stops adding timed steps once this much wall time is spent


def reference_steps(n_mc, workload, steps, warmup, threads=None, budget_s=REF_TIME_BUDGET_S):
    """The reference arm: the reference's OWN `build_predictor(cfg)` -> `predictor(input_im)`
    (src/apply_net.py:82-91, probabilistic_inference.py:86-111) on the host cores, one FULL image per step at the
    workload's sample count -- no extrapolation.  The reference modules come unmodified from /root/reference or
    from the byte-for-byte staged copy under oracle/_ref (oracle/make_ref.py); its un-vendored detectron2 dependency
    is the stand-in of oracle/ref_shim, whose backbone hands the synthetic FPN maps to the unmodified forward()
    (the ResNet-FPN backbone is outside the step in both arms).  torch's own generator drives dropout and sampling
    (stock code path).  Falls back to the oracle port (oracle/podref.py, kind "port") only if no reference sources
    are on the machine.  Returns (per-step seconds list, cores, kind, description)."""
    from pod_compare_b200 import synthetic as S
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = build_cfg(n_mc, workload)
    desc_w, _, mc, members = WORKLOADS[workload]
    m = cfg.MODEL.PROBABILISTIC_MODELING
    use_dropout, cls_var, bbox_cov = m.DROPOUT_RATE != 0.0, m.CLS_VAR_LOSS.NAME != "none", m.BBOX_COV_LOSS.NAME != "none"
    sds = head_state_dicts(members, use_dropout, cls_var, bbox_cov)
    import torchvision          # first use pulls in seconds of lazy imports: keep them out of the timing
    torchvision.ops.nms(torch.tensor([[0.0, 0.0, 1.0, 1.0]]), torch.tensor([1.0]), 0.5)
    from oracle import ref_runner as R
    if R.reference_available():
        kind = "reference"
        pred = R.build_reference_predictor(cfg, sds if members > 1 else sds[0])

        def one_image(i):
            feats = S.make_member_features(members, i, HEIGHT, WIDTH) if members > 1 else S.make_features(0, i, HEIGHT, WIDTH)
            R._set_features(pred, feats)
            input_im = [{"image": torch.zeros((3, HEIGHT, WIDTH), dtype=torch.uint8), "height": HEIGHT, "width": WIDTH,
                         "image_id": i}]
            t0 = time.perf_counter()
            with torch.no_grad():
                out = pred(input_im)
            return time.perf_counter() - t0, len(out)
        src = "unmodified reference modules from %s via oracle/ref_shim" % R.REFERENCE_ROOT
    else:
        kind = "port"
        from oracle import podref as O
        pp = O.PathParams.from_cfg(cfg)
        hws = [O.unpack_head(sd, pp) for sd in sds]
        mode = cfg.PROBABILISTIC_INFERENCE.INFERENCE_MODE

        def one_image(i):
            feats = S.make_member_features(members, i, HEIGHT, WIDTH) if members > 1 else S.make_features(0, i, HEIGHT, WIDTH)
            t0 = time.perf_counter()
            with torch.no_grad():
                out = O.predict(feats, hws, pp, mode, (HEIGHT, WIDTH), n_mc=n_mc if mc else 1, seed=0, image=i,
                                dropout_mode="torch")
            return time.perf_counter() - t0, int(out.boxes.shape[0])
        src = "oracle/podref.py (no reference sources on this machine)"
    for w in range(min(warmup, 1)):
        one_image(1000 + w)                     # untimed; one image pages in every code path
    times, dets, t_start = [], [], time.perf_counter()
    for k in range(steps):
        t, n = one_image(k)
        times.append(t)
        dets.append(n)
        if time.perf_counter() - t_start > budget_s:
            break
    desc = ("%s; one full 1280x720 image per step through build_predictor(cfg) -> predictor(input_im), %s, %d torch threads; "
            "%d of %d requested steps timed (wall-time budget %.0f s), %.2f s per image, no extrapolation; detections per "
            "image %s" % (src, ("N=%d MC-dropout samples" % n_mc) if mc else ("E=%d members" % members if members > 1 else "single forward"),
                          threads, len(times), steps, budget_s, sum(times) / len(times), dets[:4]))
    return times, threads, kind, desc


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    times, cores, kind, desc = reference_steps(args.n_mc, args.workload, args.steps, args.warmup)
    per_image = sum(times) / len(times)
    value = 1.0 / per_image
    line = {"impl": "reference", "metric": METRIC if args.n_mc == N_MC else METRIC.replace("N=30", "N=%d" % args.n_mc),
            "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * per_image, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.workload != "mc_pre":
        line["metric"] = "images/sec, side workload (not the BASELINE headline)"
    print(json.dumps(line))


def cpu_baseline_subprocess(args):
    """cpu_baseline leg of the GPU arm: one image through the reference arm in a child process (the GPUs are hidden
    from it, see the top of this file)."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--n-mc", str(args.n_mc), "--workload", args.workload, "--batch", str(args.batch)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=1800)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    raise RuntimeError("reference arm failed: %s" % out.stderr[-2000:])


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--chunk", type=int, default=16, help="images processed per head pass (memory bound)")
    ap.add_argument("--n-mc", type=int, default=N_MC)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="mc_pre", choices=sorted(WORKLOADS),
                    help="mc_pre = the headline metric (BASELINE configs[2]); the others are side measurements")
    ap.add_argument("--kblock", type=int, default=0)
    ap.add_argument("--pair", type=int, default=-1, help="1/0: force CTA-pair (cta_group::2) / single-CTA tower convs")
    ap.add_argument("--keep-unread", action="store_true",
                    help="also evaluate the last sample's class / variance tower passes, whose outputs the reference "
                         "computes but never reads (default: left out, results identical)")
    ap.add_argument("--q1-epilogue", action="store_true",
                    help="form the sample mean of the last tower layer inside the tcgen05 epilogue instead of by the streaming pass")
    ap.add_argument("--no-fuse-q1", action="store_true",
                    help="evaluate cls_score / cls_var / bbox_cov for every MC sample and average the outputs (as the reference "
                         "does) instead of accumulating the last tower layer over the samples (default: fused)")
    ap.add_argument("--from-images", action="store_true",
                    help="side measurement: a step starts from raw uint8 frames and includes the ResNet-50-FPN backbone on this "
                         "repository's kernels (frames padded to 736x1280 as detectron2 does); default: FPN features in")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="single-forward workloads: capture the whole step (backbone + head + post-processing) in a CUDA graph and "
                         "replay it (launch-bound small batches); the batch must fit one chunk")
    ap.add_argument("--backbone", default="tc", choices=["tc", "torch"],
                    help="with --from-images: tc = this repository's kernels (default), torch = library fp32 convolutions (the "
                         "oracle of backbone_tc.py, for comparison)")
    ap.add_argument("--profile-layers", action="store_true", help="report tower time per layer in ms_per_step_by_kernel")
    ap.add_argument("--halo", type=int, default=-1, help="row-halo activation staging: bit 0 = pixels-as-M kernels, bit 1 = weights-as-A kernel (default 3)")
    ap.add_argument("--tile-width", type=int, default=0, help="pixel-tile width of the CTA-pair tower kernel: 16, 32, 0 = per map shape (default)")
    ap.add_argument("--chunk-taps", type=int, default=0, help="taps per accumulation chunk (1, 3, 9)")
    ap.add_argument("--chunk-kblocks", type=int, default=0, help="K-blocks per accumulation chunk (overrides --chunk-taps)")
    ap.add_argument("--trunc-comp", type=float, default=-1.0, help="truncation compensation, ulps per MMA accumulation")
    args = ap.parse_args()

    from pod_compare_b200 import distributed as D
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    D.init_from_env("nccl")
    import torch.distributed as dist
    from pod_compare_b200 import ops, synthetic as S
    from pod_compare_b200.predictor import build_predictor
    if args.kblock:
        ops.set_conv_kblock(args.kblock)
    if args.pair >= 0:
        ops.set_conv_pair(args.pair)
    if args.halo >= 0:
        ops.set_conv_halo(args.halo)
    if args.tile_width:
        ops.set_conv_tile_width(args.tile_width)
    if args.chunk_taps:
        ops.set_conv_chunk_taps(args.chunk_taps)
    if args.chunk_kblocks:
        ops.set_conv_chunk_kblocks(args.chunk_kblocks)
    if args.trunc_comp >= 0:
        ops.set_conv_trunc_comp(args.trunc_comp)

    cfg = build_cfg(args.n_mc, args.workload)
    pred = build_predictor(cfg)
    m = pred.model
    members = WORKLOADS[args.workload][3]
    mc_workload = WORKLOADS[args.workload][2]
    sds = head_state_dicts(members, m.use_dropout, m.compute_cls_var, m.compute_bbox_cov)
    pred.load_weight_sets(sds if members > 1 else sds[0])
    pred.skip_unread_outputs = not args.keep_unread
    pred.fuse_sample_mean = False if args.no_fuse_q1 else ("epilogue" if args.q1_epilogue else "stream")
    pred._engine.profile_layers = args.profile_layers
    B = args.batch
    # synthetic FPN features of this rank's images (weak scaling: B images per GPU)
    img0 = rank * B
    host_feats = None
    # every image of the batch is distinct (seeded by its global image id)
    if args.from_images:
        from pod_compare_b200 import backbone as BB
        bsd = [BB.random_state_dict(e) for e in range(members)]
        pred.load_backbone(bsd if members > 1 else bsd[0], impl=args.backbone)
        host_feats = torch.stack([S.make_image(0, img0 + i, HEIGHT, WIDTH) for i in range(B)]).pin_memory()
        dev_feats = host_feats.cuda(non_blocking=True)
        h2d_bytes = host_feats.numel()
    elif members > 1:
        # ensembles: every member is a full model with its own backbone, hence its own feature maps (feats[e][l])
        per = [S.make_member_features(members, img0 + i, HEIGHT, WIDTH) for i in range(B)]
        host_feats = [[torch.cat([per[i][e][l] for i in range(B)], 0).pin_memory() for l in range(5)] for e in range(members)]
        dev_feats = [[f.cuda(non_blocking=True) for f in fs] for fs in host_feats]
        h2d_bytes = sum(f.numel() * 4 for fs in host_feats for f in fs)
    else:
        per = [S.make_features(0, img0 + i, HEIGHT, WIDTH) for i in range(B)]
        host_feats = [torch.cat([per[i][l] for i in range(B)], 0).pin_memory() for l in range(5)]
        dev_feats = [f.cuda(non_blocking=True) for f in host_feats]
        h2d_bytes = sum(f.numel() * 4 for f in host_feats)
    torch.cuda.synchronize()

    runner = None
    if args.cuda_graph:
        if mc_workload or args.chunk < B:
            raise SystemExit("--cuda-graph is for the single-forward workloads with --chunk >= --batch")
        runner = pred.capture(dev_feats, out_hw=(HEIGHT, WIDTH), image0=img0)

    def step(feats):
        if runner is not None:
            _, det = runner(feats)           # copies the (host or device) inputs into the captured buffers, replays
            return D.all_gather_records(D.pack_records(det))
        # one public-API call per step: the predictor evaluates the batch in chunks of args.chunk images and, for
        # host-resident features, uploads chunk i+1 on a copy stream while chunk i computes
        if args.from_images:
            _, det = pred.infer_from_images(feats, (HEIGHT, WIDTH), image0=img0, chunk_images=args.chunk, return_det=True)
        else:
            _, _, cand, det = pred.infer_from_features(feats, (HEIGHT, WIDTH), (HEIGHT, WIDTH), image0=img0,
                                                       return_candidates=True, chunk_images=args.chunk)
        return D.all_gather_records(D.pack_records(det))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        # the cyclic garbage collector is kept out of the timed region (as timeit does): a generation-2 pass over the
        # tensor / ctypes objects of earlier steps costs 50-90 ms of host time at a random launch
        gc.collect()
        gc.disable()
        try:
            return _timed(fn, steps)
        finally:
            gc.enable()

    def _timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # ---- warm-up, then the device-resident measurement (value) ----
    for _ in range(args.warmup):
        step(dev_feats)
    launches0 = ops.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ops.PROFILE = []                                  # per-launch CUDA events of the dominant kernel
    ms, rec = timed(lambda: step(dev_feats), args.steps)
    prof = ops.PROFILE
    ops.PROFILE = None
    launches = ops.launch_count() - launches0
    # ---- end-to-end through the plugin call with HOST buffers (e2e) ----
    d2h_bytes = rec.numel() * 4

    def e2e_step():
        r = step(host_feats)                          # infer_from_features uploads the pinned host maps
        return r.cpu()

    e2e_step()
    ms_e2e, _ = timed(e2e_step, args.steps)
    clocks = sampler.stop()

    images = B * world * args.steps
    value = images / (ms / 1000.0)
    e2e_value = images / (ms_e2e / 1000.0)

    # ---- roofline of the dominant kernel: the 256->256 tower convolution (tcgen05) ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured bf16_tflops_sustained (MEASURED_PEAKS.json)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    roof = None
    if prof:
        torch.cuda.synchronize()
        tot_ms, tot_flop, n = 0.0, 0.0, 0
        all_ms = 0.0
        hbm = {}
        by_tag = {}
        for (s, e, flop, tag) in prof:
            d = s.elapsed_time(e)
            by_tag[tag] = by_tag.get(tag, 0.0) + d / args.steps
            if tag in ("mask_expand", "sample_mean", "act_mean"):
                h = hbm.setdefault(tag, [0.0, 0.0, 0])
                h[0] += d; h[1] += flop; h[2] += 1
                continue
            all_ms += d
            if tag.startswith("tower256"):
                tot_ms += d
                tot_flop += flop
                n += 1
        if n:
            ach = tot_flop / (tot_ms / 1000.0) / 1e12
            traffic, traffic_note = None, None
            try:
                with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
                    tj = json.load(f)
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_note = ("dram read+write of ONE ncu-captured launch of this kernel (%d maps of %dx%dx256; algorithmic "
                                "bytes of that launch: %.3e) -- %s" % (tj["maps"], tj["H"], tj["W"], tj["algorithmic_bytes"], tj["source"]))
            except Exception:  # noqa: BLE001
                pass
            roof = {"bound": "tensor", "kernel": "tc::k_conv3x3_tc2<64,HIDDEN> (256->256 tower conv on CTA pairs, fp16x3 split)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_note": traffic_note,
                    "peak_source": peak_src, "launches": n, "avg_launch_ms": tot_ms / n,
                    "algorithmic_flop_per_launch": tot_flop / n,
                    "mma_tflops": 3.0 * ach, "mma_frac": 3.0 * ach / peak,
                    "conv_share_of_step": all_ms / ms,
                    "ms_per_step_by_kernel": {k: round(v, 3) for k, v in sorted(by_tag.items())},
                    "note": "achieved counts ALGORITHMIC fp32 FLOPs (2*9*256*256 per location); each is issued as 3 fp16 "
                            "tensor-core MMAs (hi*hi, hi*lo, lo*hi), so frac <= 1/3 by construction and mma_frac is the "
                            "tensor-pipe utilisation"}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_kernels = {}
    if prof:
        names = {"mask_expand": "k_mask_expand (first-layer MC replication: 1 fp32 read, N x passes split-pair writes)",
                 "sample_mean": "k_sample_mean_q1 (per-anchor statistics: Q1 mean over the N samples)",
                 "act_mean": "k_q1_mean_act (Q1 mean of the last tower layer over the N samples: the cov-head statistics stream)"}
        for tag, (t_ms, nbytes, cnt) in hbm.items():
            gbs = nbytes / (t_ms / 1000.0) / 1e9
            hbm_kernels[tag] = {"kernel": names[tag], "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                "frac": gbs / hbm_peak, "launches": cnt, "share_of_step": t_ms / ms}
    line = {"metric": METRIC if args.n_mc == N_MC else METRIC.replace("N=30", "N=%d" % args.n_mc), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16x3 split operands on tcgen05, fp32 accumulate)", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches, "roofline": roof, "hbm_kernels": hbm_kernels}
    if args.workload != "mc_pre":
        line["metric"] = "images/sec, side workload (not the BASELINE headline)"
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_subprocess(args)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
