"""GPU checks of the host-side step plumbing: the H2D prefetcher and the channels-last FCN expert."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_device_prefetcher_delivers_every_batch_in_order():
    from exp.parallel import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    batches = []
    for n in (8, 8, 8, 5):                                   # ragged last batch, as a DataLoader with drop_last=False
        x = torch.randn(n, 29, 12, generator=g).pin_memory()
        y = torch.randint(0, 9, (n, 1), generator=g).pin_memory()
        m = torch.ones(n, 29, dtype=torch.bool).pin_memory()
        batches.append((x, y, m))
    for _ in range(2):                                       # second pass reuses the pooled device buffers
        got = list(DevicePrefetcher(iter(batches), torch.device(DEV)))
        assert len(got) == len(batches)
        for (x, y, m), (xd, yd, md) in zip(batches, got):
            assert xd.is_cuda and xd.dtype == torch.float32 and yd.dtype == torch.int64 and md.dtype == torch.float32
            assert yd.shape == (x.shape[0],) and md.shape == (x.shape[0], 29)
            # the views are only valid until the next batch is taken: compare copies taken right away
        # values: re-run and compare batch by batch while each batch is current
        for (x, y, m), (xd, yd, md) in zip(batches, DevicePrefetcher(iter(batches), torch.device(DEV))):
            torch.cuda.synchronize()
            assert torch.equal(xd.cpu(), x) and torch.equal(yd.cpu(), y.squeeze(-1)) and torch.equal(md.cpu(), m.float())


def test_prefetcher_overlaps_with_consumer_work_without_corruption():
    """The copy of batch i+1 must not overwrite batch i while the consumer's kernels still read it."""
    from exp.parallel import DevicePrefetcher
    host = [(torch.full((64, 1000, 125), float(i)).pin_memory(), torch.zeros(64, 1, dtype=torch.long).pin_memory(),
             torch.ones(64, 1000, dtype=torch.bool).pin_memory()) for i in range(6)]
    sums = []
    for xd, yd, md in DevicePrefetcher(iter(host), torch.device(DEV)):
        acc = xd
        for _ in range(20):                                  # enough queued work for the next copy to be in flight
            acc = acc * 1.0 + 0.0
        sums.append(acc.mean())
    torch.cuda.synchronize()
    assert [round(float(s)) for s in sums] == [0, 1, 2, 3, 4, 5]


def test_fcn_channels_last_path_matches_conv1d_path():
    """models/FullyConvNet.py runs 1 x k conv2d in channels_last on CUDA; same arithmetic as the Conv1d stack."""
    from types import SimpleNamespace
    import torch.nn.functional as F
    from models.FullyConvNet import FullyConvNetwork
    torch.manual_seed(0)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        net = FullyConvNetwork(SimpleNamespace(seq_len=100, enc_in=12, num_class=5)).to(DEV).eval()
        x = torch.randn(4, 100, 12, device=DEV)
        out = net(x)
        h = x.transpose(1, 2)
        for blk in (net.block1, net.block2, net.block3):
            conv, bn = blk[0], blk[1]
            h = F.relu(F.batch_norm(F.conv1d(h, conv.weight, conv.bias), bn.running_mean, bn.running_var, bn.weight, bn.bias,
                                    False, 0.0, bn.eps))
        ref = net.fc(h.mean(dim=2))
        torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
        # parameter names and shapes are the reference's (state dicts interchange)
        sd = net.state_dict()
        assert sd["block1.0.weight"].shape == (128, 12, 8) and "block1.1.running_mean" in sd and "fc.weight" in sd
    finally:
        torch.backends.cudnn.allow_tf32 = old
