"""Probe: cost of the host->device batch path of one training step (resident / sequential copy / DevicePrefetcher /
hand-rolled double buffer) and whether a concurrent H2D copy disturbs the L1 forward kernel.  Run on a GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))
import torch
import bench
from exp.experiment_classification import Experiment
from exp.parallel import DevicePrefetcher
sys.argv = ["bench.py"]
a = bench.parse(); cfg = bench.model_args(a)
exp = Experiment(cfg, load_data=False); exp.model.train(); dev = exp.device
B, T, M, C = 256, 1000, 125, 3
xh = torch.randn(B, T, M).pin_memory(); yh = torch.randint(0, C, (B, 1)).pin_memory(); mh = torch.ones(B, T, dtype=torch.bool).pin_memory()
mf = torch.ones(B, T).pin_memory()
exp.grads.zero_grad()
def sync(): torch.cuda.synchronize()
step = 0
for _ in range(3):
    step += 1; float(exp.train_step(*exp._to_device(xh, yh, mh), 0, step))
sync()
def run(name, fn, n=8):
    global step
    sync(); t0 = time.perf_counter()
    fn(n)
    sync(); print("%-34s %7.2f ms/step" % (name, (time.perf_counter() - t0) / n * 1e3))
def seq(n):
    global step
    for _ in range(n):
        step += 1; float(exp.train_step(*exp._to_device(xh, yh, mh), 0, step))
def pre(n, mask=mh):
    global step
    for xd, yd, md in DevicePrefetcher(((xh, yh, mask) for _ in range(n)), dev):
        step += 1; float(exp.train_step(xd, yd, md, 0, step))
def pre_nosync(n):
    global step
    for xd, yd, md in DevicePrefetcher(((xh, yh, mf) for _ in range(n)), dev):
        step += 1; l = exp.train_step(xd, yd, md, 0, step)
    float(l)
def copy_only(n):
    for _ in range(n):
        x = xh.to(dev, non_blocking=True)
xd, yd, md = exp._to_device(xh, yh, mh)
def resident(n):
    global step
    for _ in range(n):
        step += 1; float(exp.train_step(xd, yd, md, 0, step))
run("resident", resident); run("sequential copy+step", seq); run("prefetch (bool mask)", pre); run("prefetch (float pinned mask)", lambda n: pre(n, mf)); run("prefetch no per-step sync", pre_nosync); run("copy only", copy_only)
run("resident", resident); run("prefetch (float pinned mask)", lambda n: pre(n, mf))

# ---- variant: persistent double buffers, explicit events, no allocator involvement
side = torch.cuda.Stream(device=dev)
xb = [torch.empty(B, T, M, device=dev) for _ in range(2)]
yb = [torch.empty(B, dtype=torch.long, device=dev) for _ in range(2)]
mb = [torch.empty(B, T, device=dev) for _ in range(2)]
ready = [torch.cuda.Event() for _ in range(2)]
done = [torch.cuda.Event() for _ in range(2)]
yflat = yh.squeeze(-1)
def issue(i):
    k = i & 1
    with torch.cuda.stream(side):
        side.wait_event(done[k])
        xb[k].copy_(xh, non_blocking=True); yb[k].copy_(yflat, non_blocking=True); mb[k].copy_(mf, non_blocking=True)
        ready[k].record(side)
def dbuf(n, persync=True):
    global step
    cur = torch.cuda.current_stream(dev)
    done[0].record(cur); done[1].record(cur)
    issue(0)
    for i in range(n):
        k = i & 1
        cur.wait_event(ready[k])
        if i + 1 < n: issue(i + 1)
        step += 1; l = exp.train_step(xb[k], yb[k], mb[k], 0, step)
        done[k].record(cur)
        if persync: float(l)
    float(l)
run("double buffer + events (sync)", dbuf); run("double buffer + events (sync)", dbuf); run("double buffer (no per-step sync)", lambda n: dbuf(n, False)); run("resident", resident)
# does a concurrent copy slow a lone kernel?  time the forward of one group with and without a copy in flight
from layers.shapelet_ops import instance_norm, shapelet_transform
pack = instance_norm(xd); W = torch.randn(5, M, 300, device=dev)
def fwd_alone(n):
    for _ in range(n):
        with torch.no_grad(): shapelet_transform(pack, W, 1, 1.0, "l1")
def fwd_with_copy(n):
    for _ in range(n):
        with torch.cuda.stream(side): xb[0].copy_(xh, non_blocking=True)
        with torch.no_grad(): shapelet_transform(pack, W, 1, 1.0, "l1")
run("L1 fwd L=300 alone", fwd_alone); run("L1 fwd L=300 + concurrent H2D", fwd_with_copy)
