"""A/B timing of the L1 layer backward for a few (T, L, K) geometries with the library given by IGN_B200_LIB.
Usage: IGN_B200_LIB=... python tools/ab_bwd.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))
import torch  # noqa: E402
from layers.shapelet_ops import instance_norm, shapelet_transform  # noqa: E402

B, M = 256, 125
torch.manual_seed(0)
for (T, L, K) in [(1000, 500, 5), (1000, 500, 10), (1000, 500, 100), (2000, 1000, 10), (1000, 100, 10), (1000, 300, 100)]:
    x = torch.randn(B, T, M, device="cuda")
    pack = instance_norm(x)
    W = torch.randn(K, M, L, device="cuda", requires_grad=True)
    ts = []
    for it in range(4):
        p, d, _ = shapelet_transform(pack, W, 1, 1.0, "l1")
        g = torch.ones_like(p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); p.backward(g); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        W.grad = None
    print("T=%d L=%d K=%d  bwd %.3f ms" % (T, L, K, min(ts[1:])))
