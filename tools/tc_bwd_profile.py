"""Debug: role-level wait accounting of the tcgen05 backward contraction (library built with -DIGN_TC_PROFILE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))
import torch
from layers import ign_cabi as C
from layers.shapelet_ops import instance_norm, shapelet_transform
NAMES = ["prod wait row", "prod wait emptyA", "build wait emptyB", "mma wait fullB", "mma wait fullA", "mma wait accempty",
         "drain wait accfull", "loader wait rowempty", "prod warp0 total", "builder warp0 total", "mma total", "loader total"]
B, M, T, K = 256, 125, 1000, 5
x = torch.randn(B, T, M, device="cuda"); pack = instance_norm(x)
for L in (100, 200, 500):
    W = torch.randn(K, M, L, device="cuda", requires_grad=True)
    p, _, _ = shapelet_transform(pack, W, 1, 1.0, "cosine", precision="3xtf32"); g = torch.randn_like(p)
    p.backward(g, retain_graph=True); torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * 16)(); C.lib.ign_debug_tc_profile(buf, 3)
    W.grad = None; p.backward(g, retain_graph=True); torch.cuda.synchronize()
    C.check(C.lib.ign_debug_tc_profile(buf, 3), "profile")
    print("L=%d (per CTA, Mcycles; 148 CTAs)" % L)
    for i, n in enumerate(NAMES):
        print("   %-22s %9.3f" % (n, buf[i] / 148 / 1e6))
