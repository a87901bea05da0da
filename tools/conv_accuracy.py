"""Accuracy of the tcgen05 fp16x3 convolution against an fp64 convolution, for the three
accumulation-chunk settings, next to the fp32 CPU convolution the reference runs.
    python tools/conv_accuracy.py          (needs a B200)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pod_compare_b200 import ops
from tests import gpu_util as G

torch.manual_seed(0)
g = torch.Generator().manual_seed(11)
NB, C, H, W = 2, 256, 48, 80
x = torch.relu(torch.randn((NB, C, H, W), generator=g)) * 1.25
w = torch.randn((256, C, 3, 3), generator=g) * (2.0 / (9 * C)) ** 0.5
b = torch.randn((256,), generator=g) * 0.05
ref = G.conv_ref64(x, w, b, False)
cpu = torch.nn.functional.conv2d(x, w, b, padding=1)


def stats(got):
    d = (got.double() - ref)
    return "max|d|/max|ref| %.3e   rms(d)/rms(ref) %.3e   mean(d*sign(ref))/mean|ref| %+.3e" % (
        float(d.abs().max() / ref.abs().max()), float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
        float((d * ref.sign()).mean() / ref.abs().mean()))


print("shape", (NB, C, H, W), "-> 256 channels, K = 2304")
print("fp32 CPU conv (reference arithmetic):", stats(cpu))
ops.set_conv_chunk_kblocks(0)
for taps in (9, 3, 1):
    ops.set_conv_chunk_taps(taps)
    for kb in (32, 64):
        ops.set_conv_kblock(kb)
        print("tcgen05 fp16x3, chunk = %d tap(s), kblock %d:" % (taps, kb), stats(G.tc_conv_raw(x, w, b, False)))
ops.set_conv_chunk_taps(1)
ops.set_conv_chunk_kblocks(6)
ops.set_conv_kblock(64)
