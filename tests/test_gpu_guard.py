"""Out-of-bounds writes: every buffer the host layer hands to the C ABI (outputs, saved distances, workspaces, gradients)
is carved out of a larger allocation with 4 KB of sentinel bytes on either side, and the sentinels must survive a
forward + backward.  (compute-sanitizer is not available on the GPU pool; this is the check that is.)  The kernels' own
bounds logic is what is under test: ragged last tiles, masked float4 tails, padded rows, lag blocks, residues."""
import pytest
import torch

DEV = "cuda"

pytestmark = pytest.mark.gpu

GUARD = 4096
FILL = 0xA5


class _GuardedTorch:
    """Stands in for the `torch` module inside layers.shapelet_ops: allocation calls return views into guarded buffers."""

    def __init__(self):
        self.allocs = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def _alloc(self, shape, dtype, device, zero):
        if isinstance(shape, int):
            shape = (shape,)
        n = 1
        for v in shape:
            n *= int(v)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        body = (nbytes + 255) // 256 * 256
        raw = torch.full((GUARD + body + GUARD,), FILL, dtype=torch.uint8, device=device)
        view = raw[GUARD:GUARD + nbytes]
        if zero:
            view.zero_()
        self.allocs.append((raw, nbytes))
        return view.view(dtype).view(*shape) if n else torch.empty(shape, dtype=dtype, device=device)

    def empty(self, *shape, dtype=torch.float32, device=None, **kw):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        return self._alloc(tuple(shape), dtype, device, False)

    def zeros(self, *shape, dtype=torch.float32, device=None, **kw):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        return self._alloc(tuple(shape), dtype, device, True)

    def empty_like(self, t, **kw):
        return self._alloc(tuple(t.shape), t.dtype, t.device, False)

    def zeros_like(self, t, **kw):
        return self._alloc(tuple(t.shape), t.dtype, t.device, True)

    def check(self, what):
        torch.cuda.synchronize()
        assert self.allocs, "no guarded allocation was made: the proxy is not in the path"
        for raw, nbytes in self.allocs:
            front = raw[:GUARD]
            back = raw[GUARD + nbytes:]
            assert bool((front == FILL).all()), "%s: write BEFORE a %d-byte buffer" % (what, nbytes)
            assert bool((back == FILL).all()), "%s: write PAST a %d-byte buffer (first bad byte at +%d)" % (
                what, nbytes, int((back != FILL).nonzero()[0]))


@pytest.fixture
def guarded(monkeypatch):
    from layers import shapelet_ops
    proxy = _GuardedTorch()
    monkeypatch.setattr(shapelet_ops, "torch", proxy)
    return proxy


CASES = [
    # B, M, T, K, L, stride, dist, precision, pool
    (3, 3, 300, 5, 100, 1, "l1", "fp32", "rbf_max"),          # 10-lag tiles
    (3, 3, 300, 5, 95, 1, "l1", "fp32", "rbf_max"),           # padded lags
    (2, 5, 333, 7, 37, 1, "l1", "fp32", "lts_min"),
    (2, 2, 1030, 3, 5, 1, "cosine", "fp32", "rbf_max"),       # generic pooling backward, 2 valid windows in the last chunk
    (2, 2, 1031, 3, 5, 1, "pearson", "fp32", "rbf_max"),
    (2, 3, 600, 5, 120, 7, "l1", "fp32", "rbf_max"),          # strided, FP32 engine
    (3, 5, 250, 5, 25, 1, "cosine", "3xtf32", "rbf_max"),     # tcgen05, ragged tile
    (2, 3, 700, 6, 520, 1, "pearson", "3xtf32", "rbf_max"),   # lag blocks in the tensor-core backward
    (5, 3, 600, 5, 120, 7, "cosine", "3xtf32", "rbf_max"),    # strided on tcgen05: residue row units
    (2, 2, 2300, 3, 50, 1, "sql2", "3xtf32", "rbf_max"),      # 2251 windows: two tiles per sample, cross-tile merge
    (2, 2, 2300, 3, 50, 1, "cosine", "tf32", "rbf_max"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(str(v) for v in c))
def test_forward_backward_write_only_inside_their_buffers(guarded, case):
    from layers.shapelet_ops import instance_norm, shapelet_transform
    B, M, T, K, L, stride, dist, precision, pool = case
    torch.manual_seed(11)
    x = torch.randn(B, T, M, device=DEV)
    W = torch.randn(K, M, L, device=DEV, requires_grad=True)
    thr = torch.rand(1, K, M, device=DEV, requires_grad=True) if pool == "lts_min" else None
    pack = instance_norm(x)
    p, dmin, idx = shapelet_transform(pack, W, stride, 0.9, dist, pool, thr, precision)
    (p * torch.randn_like(p)).sum().backward()
    assert torch.isfinite(W.grad).all()
    guarded.check("%s" % (case,))
    assert len(guarded.allocs) >= 5          # xn, stats, outputs, saved distances, backward workspace, dW


def test_recompute_backward_writes_only_inside_its_buffers(guarded, monkeypatch):
    from layers import shapelet_ops
    from layers.shapelet_ops import instance_norm, shapelet_transform
    monkeypatch.setattr(shapelet_ops, "STORE_BUDGET_BYTES", 40_000)      # forces the chunked recompute backward
    torch.manual_seed(12)
    for dist, precision in (("l1", "fp32"), ("cosine", "3xtf32")):
        x = torch.randn(3, 200, 4, device=DEV)
        W = torch.randn(45, 4, 20, device=DEV, requires_grad=True)
        p, _, _ = shapelet_transform(instance_norm(x), W, 1, 1.0, dist, precision=precision)
        (p * torch.randn_like(p)).sum().backward()
        assert torch.isfinite(W.grad).all()
    guarded.check("recompute")


def test_input_gradient_writes_only_inside_its_buffers(guarded):
    from layers.shapelet_ops import instance_norm, shapelet_transform
    torch.manual_seed(13)
    for dist, precision, stride in (("l1", "fp32", 1), ("cosine", "3xtf32", 1), ("pearson", "fp32", 3)):
        x = torch.randn(3, 203, 4, device=DEV, requires_grad=True)
        W = torch.randn(6, 4, 21, device=DEV, requires_grad=True)
        p, _, _ = shapelet_transform(instance_norm(x), W, stride, 1.0, dist, precision=precision)
        (p * torch.randn_like(p)).sum().backward()
        assert torch.isfinite(W.grad).all() and torch.isfinite(x.grad).all()
    guarded.check("dx")
