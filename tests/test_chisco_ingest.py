"""CPU: dumped CHISCO epochs -> the reference's post-loader contract (data_provider/chisco.py; reference
data_factory/eeg.py:63-69 39->3 map, :412-471 split, :501-513 item layout; eeg_processor.py:258-376 per-epoch
preprocessing).  The archive is generated here (no recordings in this environment)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))


def _dump(tmp_path, n=40, c=125, t=300, files=2):
    rng = np.random.RandomState(0)
    d = tmp_path / "chisco"
    d.mkdir()
    for f in range(files):
        np.savez(d / f"sub-0{f + 1}_imagine.npz", epochs=(rng.randn(n, c, t) * 2e-5).astype(np.float64),
                 labels=rng.randint(0, 39, n), subjects=np.full(n, f))
    return str(d)


def test_three_category_map_is_the_reference_table():
    from data_provider.chisco import THREE_CATEGORY_MAP as M
    assert len(M) == 39 and sorted(set(M.values())) == [0, 1, 2]
    assert [k for k, v in M.items() if v == 0] == [0, 13, 14, 18, 22, 23, 26, 35, 37]                      # eeg.py:66
    assert [k for k, v in M.items() if v == 1] == [1, 2, 6, 7, 9, 12, 15, 17, 24, 29, 34, 36, 38]          # eeg.py:67
    assert sum(1 for v in M.values() if v == 2) == 17                                                    # eeg.py:68


def test_preprocess_epoch_follows_the_reference_steps():
    from data_provider.chisco import preprocess_epoch
    rng = np.random.RandomState(1)
    e = rng.randn(125, 1651) * 1e-5
    out = preprocess_epoch(e)
    # 500 -> 256 Hz: q = int(1.95) = 1, scipy cannot design that FIR filter, the reference's fallback keeps every sample
    assert out.shape == (122, 1651) and out.dtype == np.float32
    np.testing.assert_allclose(out, (e[:122] * 1e6).astype(np.float32), rtol=1e-6)
    # fewer channels than 122: zero padding; an integer factor really decimates (FIR, zero phase)
    out2 = preprocess_epoch(e[:64], original_fs=512, target_fs=256)
    assert out2.shape == (122, 826) and float(np.abs(out2[64:]).sum()) == 0.0
    from scipy import signal
    np.testing.assert_allclose(out2[:64], (signal.decimate(e[:64], 2, axis=1, ftype="fir", zero_phase=True) * 1e6),
                               rtol=1e-4, atol=1e-4)
    assert preprocess_epoch(e, target_timepoints=1000).shape == (122, 1000)


@pytest.mark.parametrize("data,ncls", [("EEG", 39), ("EEG3", 3)])
def test_provider_contract_on_a_dumped_archive(tmp_path, data, ncls):
    from types import SimpleNamespace
    from data_provider.data_factory import data_provider
    root = _dump(tmp_path)
    args = SimpleNamespace(data=data, dataset="CHISCO", root_path=root, batch_size=8, num_workers=0, world_size=1, rank=0,
                           max_files=None)
    sizes = {}
    for flag in ("train", "val", "test"):
        ds, dl = data_provider(args, flag)
        sizes[flag] = len(ds)
        assert (ds.enc_in, ds.max_seq_len, ds.num_class) == (122, 300, ncls)
        X, y, m = next(iter(dl))
        assert X.shape[1:] == (300, 122) and X.dtype == torch.float32 and y.shape[1:] == (1,) and m.shape[1:] == (300,)
        assert int(y.min()) >= 0 and int(y.max()) < ncls and bool(m.all())
    assert sizes == {"train": 56, "val": 8, "test": 16}                      # 70 / 10 / 20 of 80 epochs
    # data-parallel shards of the train split are equally sized
    args.world_size = 2
    lens = []
    for r in range(2):
        args.rank = r
        lens.append(len(data_provider(args, "train")[0]))
    assert lens == [28, 28]
