"""Debug: event timeline of CTA 0 of the tcgen05 forward (library built with -DIGN_TC_PROFILE; see tc_profile.py).
Prints, per local tile, cycles relative to the first MMA start: when the MMA thread got the accumulator, got its
first A stage, issued its last MMAs; when each epilogue warp saw tfull and released the accumulator."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))
import torch  # noqa: E402
from layers import ign_cabi as C  # noqa: E402
from layers.shapelet_ops import instance_norm, shapelet_transform  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 100
train = len(sys.argv) > 2 and sys.argv[2] == "train"
B, M, T, K = 256, 125, 1000, 5
x = torch.randn(B, T, M, device="cuda")
pack = instance_norm(x)
W = torch.randn(K, M, L, device="cuda", requires_grad=train)
for _ in range(2):
    if train:
        shapelet_transform(pack, W, 1, 1.0, "cosine", precision="3xtf32")
    else:
        with torch.no_grad():
            shapelet_transform(pack, W, 1, 1.0, "cosine", precision="3xtf32")
torch.cuda.synchronize()
buf = (ctypes.c_int64 * (12 * 32 * 4))()
C.check(C.lib.ign_debug_tc_trace(buf, len(buf)), "trace")
ev = lambda r, t, s: buf[(r * 32 + t) * 4 + s]
t00 = ev(8, 0, 0)
print("tile | mma: acc_ok firstA lastA commit | prod g0: top first last | prod g1: top first last | epi warps: tfull_seen/released ...")
for t in range(2, 20):
    row = "%3d | %7d %7d %7d %7d |" % (t, *(ev(8, t, s) - t00 for s in range(4)))
    for g in (9, 10):
        row += " %7d %7d %7d |" % tuple(ev(g, t, s) - t00 for s in range(3))
    for w in range(8):
        row += " %7d/%-7d" % (ev(w, t, 0) - t00, ev(w, t, 1) - t00)
    print(row)
