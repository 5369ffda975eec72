"""Per-kernel timing of the shapelet layer at BASELINE config 2 geometry (not the contract bench: see
bench.py).  CUDA events on torch's current stream (the stream the C ABI launches on), L2 flushed between
iterations.  Usage: python tools/bench_layer.py [--B 256] [--dists l1,cosine] [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))

import torch  # noqa: E402

from layers.shapelet_ops import SeriesPack, instance_norm, shapelet_transform  # noqa: E402


def timeit(fn, iters, flush):
    fn(); fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--M", type=int, default=125)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--K", type=int, default=5)
    ap.add_argument("--Ls", default="100,200,300,500")
    ap.add_argument("--dists", default="l1,sql2,cosine,pearson")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    x = torch.randn(a.B, a.T, a.M, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    rows = []
    t_norm = timeit(lambda: instance_norm(x), a.iters, flush)
    pack = instance_norm(x)
    nbytes = 2 * x.numel() * 4
    print(f"instnorm  {t_norm:8.3f} ms  {nbytes / t_norm / 1e6:8.1f} GB/s (algorithmic read+write)")
    rows.append(dict(kernel="instnorm", ms=t_norm, gbs=nbytes / t_norm / 1e6))

    Ls = [int(v) for v in a.Ls.split(",")]

    def wstats():
        pack._stats = {}
        pack.prepare_stats("cosine", [(L, 1) for L in Ls])
    t_pre = timeit(wstats, a.iters, flush)
    pb = x.numel() * 4 + sum(a.B * a.M * ((a.T - L + 1 + 15) // 16 * 16) * 4 for L in Ls)
    print(f"winstats  {t_pre:8.3f} ms  {pb / t_pre / 1e6:8.1f} GB/s (all {len(Ls)} groups, cosine)")
    rows.append(dict(kernel="window_stats", ms=t_pre, gbs=pb / t_pre / 1e6))
    for dist in a.dists.split(","):
        tot_f = tot_b = 0.0
        for L in [int(v) for v in a.Ls.split(",")]:
            W = torch.randn(a.K, a.M, L, device=dev, requires_grad=True)
            Tw = a.T - L + 1
            E = a.B * a.M * a.K * Tw * L
            with torch.no_grad():
                tf = timeit(lambda: shapelet_transform(pack, W, 1, 1.0, dist, precision=a.precision), a.iters, flush)
            tft = timeit(lambda: shapelet_transform(pack, W, 1, 1.0, dist, precision=a.precision), a.iters, flush)
            p, dmin, idx = shapelet_transform(pack, W, 1, 1.0, dist, precision=a.precision)
            g = torch.randn_like(p)

            def bwd():
                W.grad = None
                p.backward(g, retain_graph=True)
            tb = timeit(bwd, a.iters, flush)
            tot_f += tft; tot_b += tb
            print(f"{dist:8s} L={L:4d} fwd(infer) {tf:8.3f} ms {2 * E / tf / 1e9:8.1f} TFLOP/s | fwd(train) {tft:8.3f} ms | "
                  f"bwd {tb:8.3f} ms {4 * E / tb / 1e9:8.1f} TFLOP/s")
            rows.append(dict(kernel="shapelet", dist=dist, L=L, fwd_infer_ms=tf, fwd_train_ms=tft, bwd_ms=tb,
                             fwd_tflops=2 * E / tf / 1e9, bwd_tflops=4 * E / tb / 1e9, E=E))
        print(f"{dist:8s} total fwd(train) {tot_f:8.3f} ms  bwd {tot_b:8.3f} ms  -> {a.B / (tot_f + tot_b) * 1e3:9.1f} samples/s (layer only)")
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
