"""Batch source with the reference's contract: data_provider(args, flag) -> (Dataset, DataLoader) whose
batches are (X[B,T,C] float, y[B,1], padding_mask[B,T]) (reference data_factory/uea.py:42, eeg.py:93).

Real data: UEA `.ts` archives (data_provider/uea.py) and dumped CHISCO epochs (data_provider/chisco.py) when the
dataset directory exists.  `--data synthetic` generates seeded class-conditional series of the named shape, so
run.py / run_uea.sh exercise the full training path without the archives; asking for a real archive whose directory
is missing raises (unless --allow_synthetic) instead of silently training on synthetic series.
"""
import math
import os

import torch
from torch.utils.data import DataLoader, Dataset

# (channels, length, classes) of the UEA archives named in run_uea.sh — public archive metadata
UEA_SHAPES = {
    "ArticularyWordRecognition": (9, 144, 25), "AtrialFibrillation": (2, 640, 3), "BasicMotions": (6, 100, 4),
    "CharacterTrajectories": (3, 182, 20), "LSST": (6, 36, 14), "ERing": (4, 65, 6), "Epilepsy": (3, 206, 4),
    "EthanolConcentration": (3, 1751, 4), "FaceDetection": (144, 62, 2), "FingerMovements": (28, 50, 2),
    "Handwriting": (3, 152, 26), "Heartbeat": (61, 405, 2), "InsectWingbeat": (200, 30, 10),
    "JapaneseVowels": (12, 29, 9), "Libras": (2, 45, 15), "NATOPS": (24, 51, 6), "PenDigits": (2, 8, 10),
    "RacketSports": (6, 30, 4), "SpokenArabicDigits": (13, 93, 10), "UWaveGestureLibrary": (3, 315, 8),
    "Cricket": (6, 1197, 12), "PhonemeSpectra": (11, 217, 39), "HandMovementDirection": (10, 400, 4),
    "SelfRegulationSCP1": (6, 896, 2), "SelfRegulationSCP2": (7, 1152, 2), "StandWalkJump": (4, 2500, 3),
    "PEMS-SF": (963, 144, 7), "DuckDuckGeese": (1345, 270, 5), "MotorImagery": (64, 3000, 2),
    "EigenWorms": (6, 17984, 5),
}
CHISCO_SHAPE = {"EEG3": (125, 1000, 3), "EEG": (125, 1000, 39)}   # BASELINE.json configs 2 and 5


class SyntheticSeries(Dataset):
    """Seeded synthetic classification set: each class owns a few localised waveforms (so shapelets have
    something to find) on a noise floor.  Exposes the attributes Experiment reads from a dataset."""

    def __init__(self, channels, seq_len, num_class, n, seed, subjects=1):
        gen = torch.Generator().manual_seed(seed)
        proto_gen = torch.Generator().manual_seed(1234567)          # class prototypes shared by all splits
        self.max_seq_len, self.enc_in, self.num_class = seq_len, channels, num_class
        self.seq_len = seq_len
        self.class_names = list(range(num_class))
        width = max(3, seq_len // 8)
        tpl = torch.randn(num_class, channels, width, generator=proto_gen).cumsum(-1)
        tpl = tpl / tpl.std(dim=-1, keepdim=True).clamp_min(1e-6)
        pos = torch.randint(0, max(1, seq_len - width), (num_class,), generator=proto_gen)
        self.y = torch.randint(0, num_class, (n,), generator=gen)
        self.subject = torch.randint(0, subjects, (n,), generator=gen)
        self.x = torch.randn(n, seq_len, channels, generator=gen)
        jitter = torch.randint(-2, 3, (n,), generator=gen)
        for i in range(n):
            c = int(self.y[i])
            p = int(min(max(int(pos[c]) + int(jitter[i]), 0), seq_len - width))
            self.x[i, p:p + width, :] += 1.5 * tpl[c].t()

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i:i + 1]


def collate_fn(batch, max_len=None):
    """Stack (x[T,C], y[1]) samples; padding_mask is all ones for fixed-length series (uea.py:7-55)."""
    xs, ys = zip(*batch)
    X = torch.stack(xs, dim=0)
    return X, torch.stack(ys, dim=0), torch.ones(X.shape[0], X.shape[1], dtype=torch.bool)


def _shape_for(args):
    if getattr(args, "syn_shape", None):
        c, t, k = (int(v) for v in args.syn_shape.split(","))
        return c, t, k
    if getattr(args, "dataset", "") in CHISCO_SHAPE:      # --data synthetic --dataset EEG3 | EEG
        return CHISCO_SHAPE[args.dataset]
    if args.data in CHISCO_SHAPE:
        return CHISCO_SHAPE[args.data]
    return UEA_SHAPES.get(getattr(args, "dataset", ""), (6, 100, 4))


def loso_dataset(args):
    """One synthetic multi-subject set for leave-one-subject-out evaluation (exp/loso.py): every sample carries a
    subject id (`.subject`), and each subject adds its own per-channel gain and offset so that held-out subjects
    are a genuine distribution shift, as between the CHISCO participants."""
    channels, seq_len, num_class = _shape_for(args)
    subjects = max(2, int(getattr(args, "syn_subjects", 1)))
    n = int(getattr(args, "syn_train", 512)) + int(getattr(args, "syn_val", 128)) + int(getattr(args, "syn_test", 128))
    ds = SyntheticSeries(channels, seq_len, num_class, n, 4242 + 1000 * max(0, int(getattr(args, "seed", 0))),
                         subjects=subjects)
    g = torch.Generator().manual_seed(97)
    gain = 1.0 + 0.2 * torch.randn(subjects, 1, channels, generator=g)
    offset = 0.3 * torch.randn(subjects, 1, channels, generator=g)
    ds.x = ds.x * gain[ds.subject] + offset[ds.subject]
    return ds


def equal_shard(n, world, rank):
    """Indices of rank `rank`'s shard of n samples: every rank gets ceil(n / world) of them (the tail wraps around to
    the first samples, DistributedSampler-style), so all ranks run the same number of optimisation steps per epoch —
    the gradient all-reduce is one collective per step and ranks with different step counts would desynchronise."""
    per = math.ceil(n / world)
    return [(rank + i * world) % n for i in range(per)]


def _require_real_root(args, root):
    """A real archive was asked for (`--data UEA|EEG|EEG3`) but its directory is missing: refuse, unless the caller
    explicitly allows synthetic series of that shape (`--allow_synthetic`) — silently training on synthetic data
    would file its accuracy under the real dataset's name."""
    if getattr(args, "allow_synthetic", False):
        print(f"WARNING: --data {args.data}: {root!r} does not exist; using SYNTHETIC series of that shape "
              f"(--allow_synthetic)", flush=True)
        args.data_source = "synthetic"
        return
    raise FileNotFoundError(
        f"--data {args.data}: dataset directory {root!r} does not exist. Pass --data_root, or --data synthetic "
        f"(or --allow_synthetic) to train on synthetic series of the same shape.")


def data_provider(args, flag):
    """flag in {'train','val','test'} (reference data_factory.py:29).  Under data-parallel training the
    train split is sharded by rank (every rank gets an equally sized shard)."""
    real_root = getattr(args, "root_path", None)
    world, rank = getattr(args, "world_size", 1), getattr(args, "rank", 0)
    have_root = bool(real_root) and os.path.isdir(real_root)
    if args.data != "synthetic" and not have_root:
        _require_real_root(args, real_root)
    if args.data == "UEA" and have_root:
        # real archive on disk: the sktime-free .ts reader with the reference's preprocessing (data_provider/uea.py)
        from data_provider import uea
        ds = uea.UEADataset(real_root, flag=flag)
        if flag == "train" and world > 1:                      # data parallel: equally sized shard per rank
            keep = equal_shard(len(ds), world, rank)
            ds.x, ds.y = [ds.x[i] for i in keep], ds.y[keep]

        def collate(batch):
            # the padded length is read when a batch is built, as the reference's collate lambda reads args.seq_len
            # (data_factory.py:107,136): Experiment sets it once to the longest series over ALL splits, so train,
            # val and test batches share one length (and the deep expert's Linear(d_model * seq_len) fits them all)
            return uea.collate_fn(batch, max_len=int(getattr(args, "seq_len", 0) or ds.max_seq_len))
        loader = DataLoader(ds, batch_size=args.batch_size, shuffle=(flag == "train"), num_workers=args.num_workers,
                            drop_last=False, collate_fn=collate, pin_memory=torch.cuda.is_available())
        return ds, loader
    if args.data in ("EEG", "EEG3") and have_root:
        # dumped CHISCO epochs with the reference's per-epoch preprocessing and 39 -> 3 map (data_provider/chisco.py)
        from data_provider import chisco
        ds = chisco.ChiscoEpochs(real_root, flag=flag, three_class=(args.data == "EEG3"),
                                 max_files=getattr(args, "max_files", None))
        if flag == "train" and world > 1:
            keep = equal_shard(len(ds), world, rank)
            ds.x, ds.y, ds.subject = ds.x[keep], ds.y[keep], ds.subject[keep]
        loader = DataLoader(ds, batch_size=args.batch_size, shuffle=(flag == "train"), num_workers=args.num_workers,
                            drop_last=False, collate_fn=collate_fn, pin_memory=torch.cuda.is_available())
        return ds, loader
    channels, seq_len, num_class = _shape_for(args)
    n = {"train": getattr(args, "syn_train", 512), "val": getattr(args, "syn_val", 128),
         "test": getattr(args, "syn_test", 128)}[flag]
    seed = {"train": 11, "val": 22, "test": 33}[flag] + 1000 * max(0, int(getattr(args, "seed", 0)))
    if flag == "train" and world > 1:
        n = math.ceil(n / world)
        seed += 7919 * rank
    ds = SyntheticSeries(channels, seq_len, num_class, n, seed, subjects=getattr(args, "syn_subjects", 1))
    loader = DataLoader(ds, batch_size=args.batch_size, shuffle=(flag == "train"), num_workers=args.num_workers,
                        drop_last=False, collate_fn=collate_fn, pin_memory=torch.cuda.is_available())
    return ds, loader
