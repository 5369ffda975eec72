"""Debug: role-level cycle accounting of the tcgen05 forward kernel (library must be built with
`make -C speech-imagery-eeg_b200/csrc EXTRA=-DIGN_TC_PROFILE`)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))
import torch  # noqa: E402
from layers import ign_cabi as C  # noqa: E402
from layers.shapelet_ops import instance_norm, shapelet_transform  # noqa: E402

NAMES = ["prod wait emptyA", "prod build+arrive", "mma wait fullB", "prod wait rows", "mma wait tempty", "mma wait fullA",
         "mma issue+commit", "epi wait tfull", "epi drain+math", "epi finalize", "setup: bars+alloc", "-",
         "producer loop end", "epilogue end", "mma end", "CTA lifetime"]
# slots 0,1,3 are timed by producer thread 0 (half of the stages), 7,8 by epilogue thread 0, 9 by every finaliser
B, M, T, K = 256, 125, 1000, 5
x = torch.randn(B, T, M, device="cuda")
pack = instance_norm(x)
for dist in sys.argv[1].split(","):
    for prec in ("3xtf32", "tf32"):
        for L in (100, 500):
            W = torch.randn(K, M, L, device="cuda")
            with torch.no_grad():
                shapelet_transform(pack, W, 1, 1.0, dist, precision=prec)
                torch.cuda.synchronize()
                buf = (ctypes.c_uint64 * 16)()
                C.lib.ign_debug_tc_profile(buf, 1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                shapelet_transform(pack, W, 1, 1.0, dist, precision=prec)
                e1.record()
                torch.cuda.synchronize()
                C.check(C.lib.ign_debug_tc_profile(buf, 1), "profile")
            Tw = T - L + 1
            RI = (Tw + 15) // 16
            RB = min(128 // RI, 32)
            ntiles = M * ((B + RB - 1) // RB)
            nstage = ntiles * ((L + 15 + 31) // 32)
            nctas = 148
            print(f"{dist} {prec} L={L}: {e0.elapsed_time(e1):.3f} ms, tiles {ntiles} ({ntiles / nctas:.1f} per CTA), stages {nstage}")
            for i, n in enumerate(NAMES):
                if i >= 10:
                    print(f"   {n:20s} {buf[i] / 1e6:9.2f} Mcyc total  {buf[i] / nctas:11.1f} cyc per CTA")
                    continue
                div = nstage / 2 if i in (0, 1) else nstage if i in (2, 5, 6) else ntiles
                print(f"   {n:20s} {buf[i] / 1e6:9.2f} Mcyc total  {buf[i] / div:9.1f} cyc per {'stage' if i in (0, 1, 2, 5, 6) else 'tile'}")
