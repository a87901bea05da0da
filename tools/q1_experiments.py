#!/usr/bin/env python
"""Attribution of the cost of the Q1-accumulating tower layer (pod_conv_args.q1_acc): one P3 launch of the last class-tower
layer at the benchmark shape (16 images x 29 live samples x 2 passes of 96x160x256) in several variants, CUDA-event timed.
    python tools/q1_experiments.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pod_compare_b200 import engine, ops  # noqa: E402
from pod_compare_b200._cabi import POD_OUT_HIDDEN  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    S, P, H, W = 30, 2, 96, 160
    HW = H * W
    NB = B * S * P
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    x_hi = (torch.randn((NB, HW, 256), generator=g, device=dev, dtype=torch.float32).clamp_min_(0) * 16).to(torch.float16)
    x_lo = torch.zeros_like(x_hi)
    w = torch.randn((256, 256, 3, 3), generator=g, device=dev) * (2.0 / 2304) ** 0.5
    pcv = engine.pack_conv(w, torch.zeros(256), dev)
    o_hi = torch.empty_like(x_hi); o_lo = torch.empty_like(x_hi)
    scale = torch.full((1,), 16.0, device=dev)
    acc = torch.empty((B * 2 * 8 * HW * 256,), device=dev)
    drop = ops.make_dropout(0.2, 1, 0, S, P, 0, 0, 3, 0)
    live = (S - 1) * P

    def run(q1=None, env=None):
        if env:
            os.environ[env] = "1"
        try:
            ts = []
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.conv3x3_tc(x_hi, x_lo, 1.0, NB, H, W, 256, pcv.w_hi, pcv.w_lo, pcv.w_scale, pcv.bias, 256, 256, POD_OUT_HIDDEN, True,
                               out_hi=o_hi, out_lo=o_lo, out_scale=1.0, drop=drop, in_scale_dev=scale, out_scale_dev=scale,
                               map_group=0 if q1 else S * P, map_live=0 if q1 else live, q1=q1)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return min(ts[1:]), ts
        finally:
            if env:
                del os.environ[env]

    def q(mask, group):
        return {"acc": acc, "samples": S, "passes": P, "live": [S - 1, S - 1], "mask": mask, "group": group}

    print("maps evaluated per launch:", B * live, "status", ops.status())
    print("plain (map_live)            %.2f ms" % run()[0])
    for grp in (8, 4, 15, 30):
        print("q1 schedule only  group %2d  %.2f ms" % (grp, run(q(0, grp))[0]))
    for grp in (8, 4, 15, 30):
        print("q1 accumulate     group %2d  %.2f ms" % (grp, run(q(3, grp))[0]))
    print("q1 accumulate, no read-back %.2f ms" % run(q(3, 8), env="POD_TC_DEBUG_NO_RMW")[0])
    print("plain again                 %.2f ms" % run()[0])
    print("status", ops.status())


if __name__ == "__main__":
    main()
