"""Every instantiation of the FP32 backward contraction (lag tiles of 8, 10 and 12 lags) with shapelet blocks of several
sizes, whatever the planner would have picked for the test shapes: the library reads IGN_BWD_LT / IGN_BWD_KB once per
process, so each forced plan runs in its own interpreter and is compared with the oracle's closed-form gradient."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "speech-imagery-eeg_b200"))
sys.path.insert(0, os.path.join(%(root)r, "oracle"))
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch
import ign_oracle as O
from helpers import MODES
from layers.shapelet_ops import SeriesPack, shapelet_transform
torch.manual_seed(5)
worst = 0.0
for (B, M, T, K, L, stride) in [(3, 3, 260, 7, 37, 1), (2, 2, 300, 5, 100, 1), (2, 3, 301, 10, 60, 3), (2, 2, 120, 3, 120, 1)]:
    xn, W, g = torch.randn(B, M, T), torch.randn(K, M, L), torch.randn(B, K, M)
    for flag in ("euclidean", "cosine"):
        mode, dist = MODES[flag]
        Wd = W.clone().cuda().requires_grad_(True)
        p, _, _ = shapelet_transform(SeriesPack.from_channel_major(xn.cuda()), Wd, stride, 0.9, dist)
        (p * g.cuda()).sum().backward()
        ref, _ = O.shapelet_backward_formula(xn.double(), W.double(), g.double(), stride, 0.9, mode)
        err = float((Wd.grad.cpu().double() - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
        assert err < 1e-4, (B, M, T, K, L, stride, flag, err)
print("OK %%.2e" %% worst)
"""


@pytest.mark.parametrize("lt,kb", [(8, 1), (8, 3), (8, 5), (10, 1), (10, 5), (12, 1), (12, 5), (12, 8)])
def test_forced_backward_plan_matches_oracle(lt, kb):
    env = dict(os.environ, IGN_BWD_LT=str(lt), IGN_BWD_KB=str(kb))
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and out.stdout.strip().startswith("OK"), out.stdout[-2000:] + out.stderr[-2000:]
