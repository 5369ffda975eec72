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
