"""One forward + backward of every length group of BASELINE config 2 (B=256), for `ncu --set full` captures without
warm-up repeats: python tools/profile_step_kernels.py l1|cosine"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))
import torch  # noqa: E402
from layers.shapelet_ops import instance_norm, shapelet_transform  # noqa: E402

dist = sys.argv[1] if len(sys.argv) > 1 else "l1"
prec = "fp32" if dist == "l1" else "3xtf32"
B, T, M, K = 256, 1000, 125, 5
torch.manual_seed(0)
x = torch.randn(B, T, M, device="cuda")
pack = instance_norm(x)
Ls = [100, 200, 300, 500]
if dist != "l1":
    pack.prepare_stats(dist, [(L, 1) for L in Ls])
for L in Ls:
    W = torch.randn(K, M, L, device="cuda", requires_grad=True)
    p, d, _ = shapelet_transform(pack, W, 1, 1.0, dist, precision=prec)
    p.backward(torch.ones_like(p))
torch.cuda.synchronize()
print("ok")
