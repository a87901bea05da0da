"""Accuracy of the tcgen05 fp16x3 convolution against an fp64 convolution for the accumulation-chunk settings
(K-blocks of 64 channels per TMEM chain), with and without the truncation compensation, next to the fp32 CPU
convolution the reference runs.  Three kernel variants: 256-channel tower conv (3 MMAs per K-step), the
weights-as-A output conv (63 channels, 2 MMAs per K-step) and the N-stacked pixels-as-M output conv.
    python tools/conv_accuracy.py [ulps_per_mma ...]         (needs a B200)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pod_compare_b200 import ops
from tests import gpu_util as G

comps = [float(a) for a in sys.argv[1:]] or [0.0, 0.27]
g = torch.Generator().manual_seed(11)
NB, C, H, W = 2, 256, 48, 80
x = torch.relu(torch.randn((NB, C, H, W), generator=g)) * 1.25
cpu_threads = os.cpu_count()
torch.set_num_threads(cpu_threads)


def stats(got, ref):
    d = (got.double() - ref)
    return "max|d|/max|ref| %.3e   rms(d)/rms(ref) %.3e   mean(d*sign(ref))/mean|ref| %+.3e" % (
        float(d.abs().max() / ref.abs().max()), float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
        float((d * ref.sign()).mean() / ref.abs().mean()))


for name, cout, wt in (("tower 256ch (pair kernel)", 256, 1), ("output 63ch weights-as-A", 63, 1), ("output 63ch pixels-as-M stacked", 63, 0)):
    w = torch.randn((cout, C, 3, 3), generator=g) * (2.0 / (9 * C)) ** 0.5
    b = torch.randn((cout,), generator=g) * 0.05
    ref = G.conv_ref64(x, w, b, False)
    print("==", name, "shape", (NB, C, H, W), "K = 2304")
    print("   fp32 CPU conv (reference arithmetic):", stats(torch.nn.functional.conv2d(x, w, b, padding=1), ref))
    ops.set_conv_wt(wt)
    for kb in (36, 18, 12, 9, 6, 4):
        ops.set_conv_chunk_kblocks(kb)
        for c in comps:
            ops.set_conv_trunc_comp(c)
            print("   chunk %2d K-blocks, comp %.2f ulp/MMA:" % (kb, c), stats(G.tc_conv_raw(x, w, b, False, cout_pad=64 if cout < 64 else None), ref))
ops.set_conv_trunc_comp(0.27)
ops.set_conv_chunk_kblocks(12)
ops.set_conv_wt(1)
