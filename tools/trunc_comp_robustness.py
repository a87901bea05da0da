"""How data-dependent is the truncation-compensation coefficient of the tcgen05 convolution?  Signed bias of the
256-channel tower convolution against fp64 (single 2304-long chain and the default chunk) for several input / weight
distributions, with compensation off and on.     python tools/trunc_comp_robustness.py [ulps]     (needs a B200)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pod_compare_b200 import ops
from tests import gpu_util as G

comp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.27
g = torch.Generator().manual_seed(5)
NB, C, H, W = 1, 256, 48, 80
torch.set_num_threads(os.cpu_count())


def bias_rms(got, ref):
    d = got.double() - ref
    return float((d * ref.sign()).mean() / ref.abs().mean()), float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())


he = (2.0 / (9 * C)) ** 0.5
cases = {
    "relu(N(0,1)) x N(0,he)  [tower layers]": (torch.relu(torch.randn((NB, C, H, W), generator=g)) * 1.25, torch.randn((256, C, 3, 3), generator=g) * he),
    "N(0,1) signed x N(0,he)  [layer 0: FPN features]": (torch.randn((NB, C, H, W), generator=g), torch.randn((256, C, 3, 3), generator=g) * he),
    "relu x |N(0,he)|  [monotone accumulation]": (torch.relu(torch.randn((NB, C, H, W), generator=g)), torch.randn((256, C, 3, 3), generator=g).abs() * he),
    "dropout(relu) x N(0,he)  [60% zeros]": (torch.relu(torch.randn((NB, C, H, W), generator=g)) * (torch.rand((NB, C, H, W), generator=g) > 0.2) * 1.25,
                                            torch.randn((256, C, 3, 3), generator=g) * he),
    "relu, heavy tails (x^3) x N(0,he)": (torch.relu(torch.randn((NB, C, H, W), generator=g)) ** 3, torch.randn((256, C, 3, 3), generator=g) * he),
    "relu x N(0,he) with 10x per-channel scales": (torch.relu(torch.randn((NB, C, H, W), generator=g)) * torch.logspace(-1, 0, C).view(1, C, 1, 1),
                                                   torch.randn((256, C, 3, 3), generator=g) * he),
    "relu x N(0,0.01)  [reference init]": (torch.relu(torch.randn((NB, C, H, W), generator=g)), torch.randn((256, C, 3, 3), generator=g) * 0.01),
}
for name, (x, w) in cases.items():
    b = torch.zeros(256)
    ref = G.conv_ref64(x, w, b, False)
    line = []
    for kb in (36, 12, 6):
        ops.set_conv_chunk_kblocks(kb)
        for c in (0.0, comp):
            ops.set_conv_trunc_comp(c)
            bs, rms = bias_rms(G.tc_conv_raw(x, w, b, False), ref)
            line.append("kb%d/c%.2f: bias %+.2e rms %.2e" % (kb, c, bs, rms))
    print("%-52s %s" % (name, " | ".join(line)))
ops.set_conv_trunc_comp(0.27)
ops.set_conv_chunk_kblocks(12)
