"""The sktime-free UEA `.ts` reader and the reference loader's preprocessing (data_provider/uea.py), CPU only.
Expected values are computed with pandas by the reference's own formulas (uea.py:78-116: Normalizer
'standardization' over all rows of the split, interpolate(method='linear', limit_direction='both'))."""
import os
import sys

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))

TS = """# a tiny archive in the published .ts format
@problemName Toy
@timeStamps false
@missing true
@univariate false
@dimensions 2
@equalLength false
@classLabel true up down flat
@data
1.0,2.0,?,4.0:0.5,0.5,0.5,0.5:up
3.0,2.0,1.0:1.5,?,2.5:down
?,1.0,1.0,1.0,1.0:0.0,0.0,0.0,0.0,?:flat
5.0,4.0:2.0,1.0:down
"""


def _write(tmp_path):
    d = tmp_path / "Toy"
    d.mkdir()
    (d / "Toy_TRAIN.ts").write_text(TS)
    (d / "Toy_TEST.ts").write_text(TS.replace("5.0,4.0:2.0,1.0:down\n", ""))
    return str(d)


def test_reader_parses_header_cases_and_missing_values(tmp_path):
    from data_provider.uea import read_ts
    cases, labels, meta = read_ts(os.path.join(_write(tmp_path), "Toy_TRAIN.ts"))
    assert labels == ["up", "down", "flat", "down"] and meta["dimensions"] == ["2"] and meta["classlabel"][0] == "true"
    assert [len(c) for c in cases] == [2, 2, 2, 2] and [len(c[0]) for c in cases] == [4, 3, 5, 2]
    assert np.isnan(cases[0][0][2]) and np.isnan(cases[1][1][1]) and cases[3][1].tolist() == [2.0, 1.0]


def test_dataset_matches_the_reference_preprocessing_formulas(tmp_path):
    from data_provider.uea import UEADataset
    ds_root = _write(tmp_path)
    ds = UEADataset(ds_root, flag="train")
    assert ds.class_names == ["down", "flat", "up"] and ds.y.tolist() == [2, 0, 1, 0]
    assert ds.max_seq_len == 5 and ds.enc_in == 2 and ds.num_class == 3 and len(ds) == 4
    # reference pipeline in pandas: per-sample interpolation, then standardisation over all rows of the split
    raw = [np.array([[1, .5], [2, .5], [np.nan, .5], [4, .5]]), np.array([[3, 1.5], [2, np.nan], [1, 2.5]]),
           np.array([[np.nan, 0], [1, 0], [1, 0], [1, 0], [1, np.nan]]), np.array([[5, 2.], [4, 1.]])]
    frames = [pd.DataFrame(r).interpolate(method="linear", limit_direction="both") for r in raw]
    df = pd.concat(frames, axis=0)
    mean, std = df.mean(), df.std()
    for i, fr in enumerate(frames):
        want = ((fr - mean) / (std + np.finfo(float).eps)).values.astype(np.float32)
        np.testing.assert_allclose(ds[i][0].numpy(), want, rtol=1e-6, atol=1e-6)
        assert ds[i][1].shape == (1,) and ds[i][1].dtype == torch.int64
    assert len(UEADataset(os.path.dirname(os.path.join(ds_root, "x")), flag="test")) == 3      # TEST file for val / test


def test_collate_pads_clips_and_masks(tmp_path):
    from data_provider.uea import UEADataset, collate_fn
    ds = UEADataset(_write(tmp_path), flag="train")
    X, y, m = collate_fn([ds[i] for i in range(4)], max_len=4)
    assert X.shape == (4, 4, 2) and y.shape == (4, 1) and m.dtype == torch.bool
    assert m.sum(1).tolist() == [4, 3, 4, 2]                     # the 5-step case is clipped to 4
    assert float(X[1, 3].abs().sum()) == 0.0 and float(X[3, 2:].abs().sum()) == 0.0
    assert torch.equal(X[2], ds[2][0][:4])


def test_data_provider_uses_the_archive_when_it_is_on_disk(tmp_path):
    from types import SimpleNamespace
    from data_provider.data_factory import data_provider
    args = SimpleNamespace(data="UEA", dataset="Toy", root_path=_write(tmp_path), batch_size=3, num_workers=0, seq_len=0,
                           world_size=1, rank=0)
    ds, dl = data_provider(args, "train")
    X, y, m = next(iter(dl))
    assert X.shape[1:] == (5, 2) and X.shape[0] == 3 and m.shape == (3, 5) and y.shape == (3, 1)
    ds_te, _ = data_provider(args, "test")
    assert len(ds_te) == 3 and ds_te.max_seq_len == 5


def test_all_splits_share_one_padded_length_when_test_series_are_longer(tmp_path):
    """JapaneseVowels-like archives have longer TEST series than TRAIN series.  The collate length is read per batch
    from args.seq_len (as the reference's collate lambda does, data_factory.py:107,136), which the experiment sets once
    to the longest series over all splits — so every split pads to the same T and the model built for it fits."""
    from types import SimpleNamespace
    from data_provider.data_factory import data_provider
    d = tmp_path / "Toy2"
    d.mkdir()
    (d / "Toy2_TRAIN.ts").write_text(TS.replace("?,1.0,1.0,1.0,1.0:0.0,0.0,0.0,0.0,?:flat\n", "1.0,1.0:0.0,0.0:flat\n"))
    (d / "Toy2_TEST.ts").write_text(TS)                       # longest series: 5 (train: 4)
    args = SimpleNamespace(data="UEA", dataset="Toy2", root_path=str(d), batch_size=8, num_workers=0, world_size=1, rank=0)
    splits = {f: data_provider(args, f) for f in ("train", "val", "test")}
    assert splits["train"][0].max_seq_len == 4 and splits["test"][0].max_seq_len == 5
    # what Experiment._get_params_from_data does
    args.seq_len = max(ds.max_seq_len for ds, _ in splits.values())
    for f, (_, dl) in splits.items():
        X, _, m = next(iter(dl))
        assert X.shape[1] == 5 and m.shape[1] == 5, f
    # a second seed in the same process rebuilds the loaders with args.seq_len already set: same lengths again
    X, _, _ = next(iter(data_provider(args, "train")[1]))
    assert X.shape[1] == 5


def test_rank_shards_are_equally_sized_so_every_rank_runs_the_same_number_of_steps(tmp_path):
    import math
    from types import SimpleNamespace
    from data_provider.data_factory import data_provider, equal_shard
    for n, world in ((129, 2), (5, 4), (7, 8), (64, 2)):
        shards = [equal_shard(n, world, r) for r in range(world)]
        assert {len(s) for s in shards} == {math.ceil(n / world)}
        assert set(i for s in shards for i in s) == set(range(n))          # every sample is seen
    d = tmp_path / "Toy3"
    d.mkdir()
    rows = TS.split("@data\n")[1].strip().split("\n")
    body = "\n".join(rows[i % 4] for i in range(9)) + "\n"              # 9 cases, world 2, batch 2: 3 steps on both ranks
    (d / "Toy3_TRAIN.ts").write_text(TS.split("@data\n")[0] + "@data\n" + body)
    (d / "Toy3_TEST.ts").write_text(TS)
    steps = []
    for rank in range(2):
        args = SimpleNamespace(data="UEA", dataset="Toy3", root_path=str(d), batch_size=2, num_workers=0, world_size=2,
                               rank=rank, seq_len=0)
        ds, dl = data_provider(args, "train")
        steps.append(len(dl))
    assert steps == [3, 3]


def test_missing_archive_raises_instead_of_silently_training_on_synthetic_series(tmp_path):
    import pytest
    import run
    from data_provider.data_factory import data_provider
    a = run.get_args(["--data", "UEA", "--data_root", str(tmp_path / "nowhere"), "--dataset", "JapaneseVowels"])
    assert a.data_source == "synthetic"
    with pytest.raises(FileNotFoundError, match="does not exist"):
        data_provider(a, "train")
    b = run.get_args(["--data", "UEA", "--data_root", str(tmp_path / "nowhere"), "--dataset", "JapaneseVowels",
                      "--allow_synthetic"])
    b.seed = 0
    ds, _ = data_provider(b, "train")                       # explicit opt-in: synthetic series of the archive's shape
    assert (ds.enc_in, ds.max_seq_len, ds.num_class) == (12, 29, 9)
    assert run.get_args(["--dataset", "JapaneseVowels"]).data_source == "synthetic"        # --data synthetic is the default
