"""BASELINE config 4: shapelet-distance layer sweep (K shapelets per length, L in {.1,.2,.3,.5}*T, T up to 4000,
batch 256, 125 channels).  Forward and backward for every point: the backward keeps the window distances while they
fit the store budget (layers/shapelet_ops.STORE_BUDGET_BYTES) and recomputes them chunk by chunk beyond it (K = 1000).
Every row is labelled with the engine the library reports for it (ign_shapelet_engine), not with the one asked for.
Writes a JSON list and a markdown table.   python tools/sweep.py --out profiles/r2_sweep"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))
import torch  # noqa: E402
from ctypes import byref  # noqa: E402
from layers import ign_cabi as C  # noqa: E402
from layers import shapelet_ops  # noqa: E402
from layers.shapelet_ops import instance_norm, shapelet_transform  # noqa: E402


def timed(fn, iters=3):
    fn()                                   # warm-up (also the first-use cudaFuncSetAttribute calls)
    if iters > 1:
        fn()                               # second warm-up: the few-ms rows were noisy run to run (allocator, clocks)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep")
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--M", type=int, default=125)
    ap.add_argument("--Ts", default="1000,2000,4000")
    ap.add_argument("--Ks", default="10,100,1000")
    ap.add_argument("--dists", default="l1,cosine")
    ap.add_argument("--precisions", default="fp32,3xtf32")
    ap.add_argument("--max_ms", type=float, default=4000.0)
    a = ap.parse_args()
    rows = []
    for T in [int(v) for v in a.Ts.split(",")]:
        x = torch.randn(a.B, T, a.M, device="cuda")
        pack = instance_norm(x)
        del x
        for frac in (0.1, 0.2, 0.3, 0.5):
            L = max(3, math.ceil(frac * T))
            stride = 1 if T < 3000 else max(1, int(math.log2(L)))
            Tw = (T - L) // stride + 1
            for K in [int(v) for v in a.Ks.split(",")]:
                E = a.B * a.M * K * Tw * L
                for dist in a.dists.split(","):
                    for prec in (a.precisions.split(",") if dist != "l1" else ["fp32"]):
                        est_ms = 2 * E / 20e12 * 1e3
                        if est_ms > a.max_ms:
                            continue
                        W = torch.randn(K, a.M, L, device="cuda", requires_grad=True)
                        iters = 7 if est_ms < 20 else 3 if est_ms < 100 else 1    # best of; the K = 1000 rows take seconds each
                        with torch.no_grad():
                            t_f = timed(lambda: shapelet_transform(pack, W, stride, 1.0, dist, precision=prec), iters)
                        desc = C.ShapeletDesc(a.B, a.M, T, C.padded_len(T), K, L, stride, 1.0, C.DIST[dist], 0,
                                              C.PRECISION[prec])
                        eng = [C.ENGINE[C.lib.ign_shapelet_engine(byref(desc), b)] for b in (0, 1)]
                        stored = 2 * C.lib.ign_shapelet_dstore_bytes(byref(desc)) <= shapelet_ops.STORE_BUDGET_BYTES
                        row = dict(T=T, L=L, stride=stride, windows=Tw, K=K, dist=dist, precision=prec, E=E,
                                   fwd_engine=eng[0], bwd_engine=eng[1], bwd_mode="stored" if stored else "recompute",
                                   fwd_ms=t_f, fwd_tflops=2 * E / t_f / 1e9)
                        if 4 * est_ms < a.max_ms:
                            p, _, _ = shapelet_transform(pack, W, stride, 1.0, dist, precision=prec)
                            g = torch.randn_like(p)

                            def bwd():
                                W.grad = None
                                p.backward(g, retain_graph=True)
                            t_b = timed(bwd, iters)
                            row.update(bwd_ms=t_b, bwd_tflops=4 * E / t_b / 1e9)
                            del p, g
                        rows.append(row)
                        print(row, flush=True)
                        del W
                        torch.cuda.empty_cache()
        del pack
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(rows, open(a.out + ".json", "w"), indent=1)
    with open(a.out + ".md", "w") as f:
        f.write("| T | L | stride | T' | K | dist | precision asked | engine fwd / bwd | bwd mode | fwd ms | fwd TFLOP/s | bwd ms | bwd TFLOP/s (4E) |\n"
                "|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| %d | %d | %d | %d | %d | %s | %s | %s / %s | %s | %.3f | %.1f | %s | %s |\n" % (
                r["T"], r["L"], r["stride"], r["windows"], r["K"], r["dist"], r["precision"], r["fwd_engine"], r["bwd_engine"],
                r["bwd_mode"], r["fwd_ms"], r["fwd_tflops"],
                "%.3f" % r["bwd_ms"] if "bwd_ms" in r else "-", "%.1f" % r["bwd_tflops"] if "bwd_tflops" in r else "-"))


if __name__ == "__main__":
    main()
