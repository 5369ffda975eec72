"""CPU, world_size 2, gloo: the data-parallel plumbing (flat gradient buffer, sliced overlapped all-reduce,
parameter broadcast) — the N>1 host logic, without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class Toy(nn.Module):
    """Has an `sbm.` prefixed part and a rest, like InterpGN (slice 0 / slice 1 of the flat buffer)."""

    def __init__(self):
        super().__init__()
        self.sbm = nn.Linear(6, 4, bias=False)
        self.deep_model = nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 4))
        self.unused = nn.Parameter(torch.zeros(3))      # never receives a gradient

    def forward(self, x):
        return self.sbm(x) + self.deep_model(x)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, overlap, ret):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speech-imagery-eeg_b200"))
    from exp.parallel import FlatGradAllReduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)               # different init per rank: broadcast must fix it
    model = Toy()
    plumbing = FlatGradAllReduce(model, world, overlap=overlap)
    torch.manual_seed(7)
    X = torch.randn(8, 6)
    Y = torch.randn(8, 4)
    shard = slice(rank * 4, rank * 4 + 4)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    for _ in range(3):
        plumbing.arm()
        loss = ((model(X[shard]) - Y[shard]) ** 2).mean()
        loss.backward()
        plumbing.finish()
        opt.step()
        plumbing.zero_grad()
    ret[rank] = {k: v.detach().clone() for k, v in model.state_dict().items()}
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_ranks_match_single_process_on_the_full_batch(overlap):
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, overlap, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    for k in r0:                                  # replicas stay bit-identical
        assert torch.equal(r0[k], r1[k]), k
    # single process, same initial weights (rank 0's), full batch: mean over 8 == average of the two shard means
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speech-imagery-eeg_b200"))
    torch.manual_seed(100)
    ref = Toy()
    torch.manual_seed(7)
    X, Y = torch.randn(8, 6), torch.randn(8, 4)
    opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for _ in range(3):
        opt.zero_grad()
        ((ref(X) - Y) ** 2).mean().backward()
        opt.step()
    for k, v in ref.state_dict().items():
        assert torch.allclose(v, r0[k], atol=1e-6), k


def test_flat_buffer_views_and_single_rank_noop():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speech-imagery-eeg_b200"))
    from exp.parallel import FlatGradAllReduce
    m = Toy()
    pl = FlatGradAllReduce(m, 1)
    assert pl.flat.numel() == sum(p.numel() for p in m.parameters())
    assert pl.bounds[0] == (0, 24)                        # the `sbm.` slice comes first
    m(torch.randn(2, 6)).sum().backward()
    assert m.sbm.weight.grad.data_ptr() == pl.flat.data_ptr()     # grads accumulate inside the flat buffer
    assert float(pl.flat.abs().sum()) > 0
    pl.finish()                                           # world 1: nothing to reduce
    pl.zero_grad()
    assert float(m.sbm.weight.grad.abs().sum()) == 0.0


def test_synthetic_provider_contract_and_rank_sharding():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speech-imagery-eeg_b200"))
    import run
    from data_provider.data_factory import data_provider
    a = run.get_args(["--dataset", "JapaneseVowels", "--batch_size", "32", "--seed", "0"])
    ds, dl = data_provider(a, "train")
    x, y, m = next(iter(dl))
    assert x.shape == (32, 29, 12) and y.shape == (32, 1) and m.shape == (32, 29)
    assert (ds.enc_in, ds.max_seq_len, ds.num_class) == (12, 29, 9)
    a.world_size, a.rank = 2, 0
    d0, _ = data_provider(a, "train")
    a.rank = 1
    d1, _ = data_provider(a, "train")
    assert len(d0) == len(d1) == 256 and not torch.equal(d0.x, d1.x)
    b = run.get_args(["--data", "synthetic", "--dataset", "EEG3"])     # CHISCO 3-class shape, synthetic series
    b.seed = 0; b.syn_train = 4
    ds3, _ = data_provider(b, "train")
    assert (ds3.enc_in, ds3.max_seq_len, ds3.num_class) == (125, 1000, 3)
    assert run.get_args(["--amp"]).amp is False and run.get_args([]).amp is True      # reference run.py:100


UEA_TS = """@problemName Odd
@timeStamps false
@missing false
@univariate false
@dimensions 2
@equalLength true
@seriesLength 4
@classLabel true a b
@data
"""


def _epoch_worker(rank, world, port, root, ret):
    """One epoch over a real (generated) UEA archive whose train split does not divide by world * batch: every step
    does the one all-reduce of the training loop; ranks with different step counts would pair all-reduces of different
    steps and hang at the final barrier."""
    import sys
    from types import SimpleNamespace
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speech-imagery-eeg_b200"))
    from data_provider.data_factory import data_provider
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    args = SimpleNamespace(data="UEA", dataset="Odd", root_path=root, batch_size=2, num_workers=0, world_size=world,
                           rank=rank, seq_len=0)
    ds, dl = data_provider(args, "train")
    steps, seen = 0, 0
    for X, y, m in dl:
        t = torch.tensor([float(X.shape[0])])
        dist.all_reduce(t)                       # the per-step gradient exchange
        steps += 1
        seen += X.shape[0]
    dist.barrier()
    ret[rank] = (len(ds), steps, seen)
    dist.destroy_process_group()


def test_uneven_real_archive_gives_every_rank_the_same_number_of_steps(tmp_path):
    d = tmp_path / "Odd"
    d.mkdir()
    rows = "".join("%d,%d,%d,%d:%d,%d,%d,%d:%s\n" % (i, i + 1, i + 2, i + 3, i, i, i, i, "ab"[i % 2]) for i in range(9))
    (d / "Odd_TRAIN.ts").write_text(UEA_TS + rows)            # 9 cases, world 2, batch 2: 5 and 4 before the fix
    (d / "Odd_TEST.ts").write_text(UEA_TS + rows)
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_epoch_worker, args=(world, port, str(d), ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert r0[0] == r1[0] == 5 and r0[1] == r1[1] == 3          # equal shards (one sample wraps around), equal step counts
