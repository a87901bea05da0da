#!/usr/bin/env python
"""What clock does the tower kernel really run at?  NVML samples the SM clock every 100 ms; under the 1 kW power cap the
hardware throttles at a much finer grain.  This probe reads clock64 and globaltimer inside CTA 0 of the CTA-pair tower
kernel (pod_conv3x3_tc_debug_clock) for a steady stream of back-to-back P3 launches (16 images x 58 live maps, the
benchmark's class-tower launch) and prints: real SM MHz, launch ms, MMA cycles the launch needs at 128 cycles per
M=256 x N=256 x K=16 MMA, and the resulting tensor-pipe utilisation AT THE REAL CLOCK.
    python tools/tower_clock_probe.py [launches]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pod_compare_b200 import engine, ops  # noqa: E402
from pod_compare_b200._cabi import POD_OUT_HIDDEN  # noqa: E402


def main():
    n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    lo_bits = int(sys.argv[2]) if len(sys.argv) > 2 else 10      # mantissa bits kept in the lo operands (10 = all; -1 = lo operands zero)
    B, S, P, H, W = 16, 30, 2, 96, 160
    HW, NB, live = H * W, B * S * P, (S - 1) * P
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    x_hi = (torch.randn((NB, HW, 256), generator=g, device=dev).clamp_min_(0) * 16).to(torch.float16)
    x_lo = (torch.randn((NB, HW, 256), generator=g, device=dev) * 0.004).to(torch.float16)
    w = torch.randn((256, 256, 3, 3), generator=g, device=dev) * (2.0 / 2304) ** 0.5
    pcv = engine.pack_conv(w, torch.zeros(256), dev)
    o_hi = torch.empty_like(x_hi); o_lo = torch.empty_like(x_hi)
    scale = torch.full((1,), 16.0, device=dev)
    drop = ops.make_dropout(0.2, 1, 0, S, P, 0, 0, 2, 0)
    if lo_bits < 10:
        # data-dependent power experiment: does the tensor pipe draw less with fewer significant bits in the lo operands?
        mask = 0 if lo_bits < 0 else (0xFFFF << (10 - lo_bits)) & 0xFFFF
        mask = mask - 65536 if mask >= 32768 else mask
        x_lo.view(torch.int16).bitwise_and_(mask)
        pcv.w_lo.view(torch.int16).bitwise_and_(mask)
        print("# lo operands: %s" % ("zero" if lo_bits < 0 else "%d mantissa bits kept" % lo_bits))
    ops.conv_debug_clock(True)
    tiles = B * live * (H // 8) * (W // 16)
    mma_cycles = tiles / 148.0 * 432 * 128          # per SM: tiles x (36 K-blocks x 4 k-steps x 3 products) x 128 cycles
    import pynvml as nv
    nv.nvmlInit()
    hnd = nv.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
    print("# launch  ms   in-kernel MHz   NVML MHz   power W   tensor-pipe utilisation at the in-kernel clock")
    for i in range(n_launch):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv3x3_tc(x_hi, x_lo, 1.0, NB, H, W, 256, pcv.w_hi, pcv.w_lo, pcv.w_scale, pcv.bias, 256, 256, POD_OUT_HIDDEN, True,
                       out_hi=o_hi, out_lo=o_lo, out_scale=1.0, drop=drop, in_scale_dev=scale, out_scale_dev=scale,
                       map_group=S * P, map_live=live)
        e1.record()
        nvml_mhz = nv.nvmlDeviceGetClockInfo(hnd, nv.NVML_CLOCK_SM)
        watts = nv.nvmlDeviceGetPowerUsage(hnd) / 1000.0
        torch.cuda.synchronize()
        cyc, ns, mhz = ops.conv_debug_clock()
        if i % 4 == 3 or i < 2:
            print("%4d  %7.2f  %9.0f  %9d  %7.0f   %.3f" % (i, e0.elapsed_time(e1), mhz, nvml_mhz, watts, mma_cycles / cyc))
    print("status", ops.status())


if __name__ == "__main__":
    main()
