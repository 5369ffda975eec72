"""Leave-one-subject-out fold batching (exp/loso.py): host logic on CPU, and the rank hand-off under gloo."""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))


def test_folds_hold_out_each_subject_once_without_leakage():
    from exp.loso import loso_folds
    g = torch.Generator().manual_seed(0)
    subj = torch.randint(0, 5, (203,), generator=g)
    folds = loso_folds(subj, val_fraction=0.1, seed=3)
    assert [f[0] for f in folds] == sorted(int(v) for v in subj.unique())
    for s, tr, va, te in folds:
        assert bool((subj[te] == s).all()) and te.numel() == int((subj == s).sum())      # whole subject held out
        assert not bool((subj[tr] == s).any()) and not bool((subj[va] == s).any())        # never seen in training
        allidx = torch.cat([tr, va, te])
        assert allidx.numel() == 203 and allidx.unique().numel() == 203                    # a partition
        assert va.numel() >= 1
    again = loso_folds(subj, val_fraction=0.1, seed=3)
    assert all(torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) for a, b in zip(folds, again))   # seeded


def test_fold_to_rank_map_covers_every_fold_once():
    from exp.loso import folds_of_rank
    for n_folds in (1, 5, 8, 13):
        for world in (1, 2, 4, 8):
            owned = sorted(f for r in range(world) for f in folds_of_rank(n_folds, world, r))
            assert owned == list(range(n_folds))
            sizes = [len(folds_of_rank(n_folds, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1                                            # balanced


def test_fold_loaders_keep_the_batch_contract():
    from types import SimpleNamespace
    from data_provider.data_factory import collate_fn, loso_dataset
    from exp.loso import fold_loaders, loso_folds
    args = SimpleNamespace(data="synthetic", dataset="JapaneseVowels", syn_shape=None, syn_train=60, syn_val=10,
                           syn_test=10, syn_subjects=4, seed=0)
    ds = loso_dataset(args)
    assert ds.subject.unique().numel() == 4
    folds = loso_folds(ds.subject)
    tr, va, te = fold_loaders(ds, folds[1], 16, collate_fn)
    x, y, m = next(iter(tr))
    assert x.shape[1:] == (29, 12) and y.shape == (x.shape[0], 1) and m.shape == x.shape[:2] and m.dtype == torch.bool
    assert sum(b[0].shape[0] for b in te) == len(folds[1][3])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from exp.loso import folds_of_rank, gather_results, summarize
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = {f: (10 + f, 0.5 * f, 0.1 * f) for f in folds_of_rank(5, world, rank)}        # (subject, loss, acc)
    merged = gather_results(local, world)
    mean_acc, rows = summarize(merged)
    if rank == 0:
        torch.save({"keys": sorted(merged), "mean": mean_acc, "rows": rows}, out)
    dist.destroy_process_group()


def test_results_are_gathered_across_ranks_gloo(tmp_path):
    out = str(tmp_path / "loso.pt")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["keys"] == [0, 1, 2, 3, 4]
    assert abs(r["mean"] - 0.2) < 1e-12 and [row[0] for row in r["rows"]] == [10, 11, 12, 13, 14]
