"""Per-group timing of the backward's phases (pooling backward, tie check, contraction, finalize) at config 2 through
ign_debug_bwd_phase_timing:  python tools/bench_pool_bwd.py [l1|cosine] [precision]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))
import torch
from layers import ign_cabi as C
from layers.shapelet_ops import instance_norm, shapelet_transform
dist = sys.argv[1] if len(sys.argv) > 1 else "l1"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
B, M, T, K = 256, 125, 1000, 5
x = torch.randn(B, T, M, device="cuda"); pack = instance_norm(x)
for L in (100, 200, 300, 500):
    W = torch.randn(K, M, L, device="cuda", requires_grad=True)
    p, _, _ = shapelet_transform(pack, W, 1, 1.0, dist, precision=prec); g = torch.randn_like(p)
    for _ in range(3):
        W.grad = None; p.backward(g, retain_graph=True)
    torch.cuda.synchronize()
    C.lib.ign_debug_bwd_phase_timing(1)
    n = 10
    for _ in range(n):
        W.grad = None; p.backward(g, retain_graph=True)
    torch.cuda.synchronize()
    ms, cnt = (ctypes.c_float * 4)(), (ctypes.c_int32 * 4)()
    C.lib.ign_debug_bwd_phase_read(ms, cnt); C.lib.ign_debug_bwd_phase_timing(0)
    Ts = (T - L + 1 + 3) // 4 * 4
    gb = 8.0 * B * M * K * Ts / 1e9
    print("%s L=%d  pool_bwd %.3f ms (%.0f GB/s algorithmic)  tie %.3f  contraction %.3f  finalize %.3f" % (
        dist, L, ms[0] / n, gb / (ms[0] / n) * 1e3, ms[1] / max(1, cnt[1]), ms[2] / n, ms[3] / n))
