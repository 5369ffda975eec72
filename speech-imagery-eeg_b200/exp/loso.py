"""Leave-one-subject-out fold batching.

The reference's README promises LOSO evaluation for CHISCO but its loader only draws a random 70/10/20 split
(reference data_factory/eeg.py:412-471).  Here the S folds of an S-subject dataset are whole, independent
training runs; they are spread over the GPUs of the box — fold f runs on rank f % world — so that the small
per-fold batches of a 39-class CHISCO fold still keep every GPU busy.  No gradient exchange happens in this
mode (each rank owns a different model); ranks only meet at the end to gather the per-fold results.

Pure host logic: everything here runs on CPU tensors and under gloo (tests/test_loso.py).
"""
import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Subset


def loso_folds(subject_ids, val_fraction=0.1, seed=0):
    """[(held_out_subject, train_idx, val_idx, test_idx)] — one fold per distinct subject.  The held-out subject is
    the test set in full; the remaining samples are split train/val by a seeded permutation (val ≥ 1 sample)."""
    subj = torch.as_tensor(subject_ids).flatten()
    folds = []
    for s in sorted(int(v) for v in subj.unique()):
        test = torch.nonzero(subj == s).flatten()
        rest = torch.nonzero(subj != s).flatten()
        if rest.numel() < 2:
            raise ValueError("leave-one-subject-out needs at least two subjects with data")
        g = torch.Generator().manual_seed(seed * 7919 + s)
        rest = rest[torch.randperm(rest.numel(), generator=g)]
        n_val = max(1, int(round(val_fraction * rest.numel())))
        folds.append((s, rest[n_val:].sort().values, rest[:n_val].sort().values, test))
    return folds


def folds_of_rank(n_folds, world, rank):
    """Round-robin fold -> rank map: rank r trains folds r, r+world, ...  Every fold is owned exactly once."""
    return list(range(rank, n_folds, world))


def fold_loaders(dataset, fold, batch_size, collate_fn, num_workers=0, pin_memory=False):
    """(train, val, test) DataLoaders over index subsets of one dataset (the reference batch contract)."""
    _, tr, va, te = fold
    mk = lambda idx, shuffle: DataLoader(Subset(dataset, idx.tolist()), batch_size=batch_size, shuffle=shuffle,
                                         num_workers=num_workers, drop_last=False, collate_fn=collate_fn,
                                         pin_memory=pin_memory)
    return mk(tr, True), mk(va, False), mk(te, False)


def gather_results(local, world):
    """Every rank contributes {fold: (subject, loss, accuracy)}; all ranks get the merged dict."""
    if world <= 1 or not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    boxes = [None] * world
    dist.all_gather_object(boxes, dict(local))
    merged = {}
    for b in boxes:
        merged.update(b)
    return merged


def summarize(results):
    """(mean accuracy, per-subject list) from the merged dict of gather_results."""
    rows = [results[k] for k in sorted(results)]
    accs = [r[2] for r in rows if r[2] is not None]
    return (sum(accs) / len(accs) if accs else float("nan")), rows
