"""The measured alternatives kept behind A/B switches must stay correct: a slice of the parity
suite is re-run in a child process under each switch (the switches are read once per process)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
SLICE = "test_3d_single_view or test_3d_batched_views_two_frames or test_3d_equal_depths or ref2d_random_96x80 or test_2d_projected_views_batched " \
        "or test_3d_adversarial_randn_eye_viewmat or test_empty_input_is_background or test_autograd_function_matches_raw_backward"

VARIANTS = {
    "bin_bytes": {"PS_BIN_MODE": "bytes"},                    # block masks from the partition kernel, carried as bytes
    "bin_split": {"PS_BIN_MODE": "split"},                    # sort + block split from bitmaps in one kernel
    "bin_split_exact": {"PS_BIN_MODE": "split", "PS_EXACT_BLOCK_MASKS": "1"},
    "bwd_v5": {"PS_BWD_V5": "1"},                             # round-1 two-phase backward (culls and re-tests every pair)
    "bwd_bulk": {"PS_BWD_BULK": "1"},                         # records staged by cp.async.bulk + mbarrier (TMA 1-D)
    "fwd_v4": {"PS_FWD_V4": "1"},                             # all-lanes walk in 3D as well
    "force_sync": {"PS_FORCE_SYNC": "1"},                     # small calls sized exactly from the mailbox
    "project_per_view": {"PS_PROJECT_PER_VIEW": "1"},         # one thread per (view, Gaussian) projection
    "rank_radix": {"PS_RANK_RADIX": "1"},                     # depth ranking by the radix kernel only
    "no_fill_fork": {"PS_NO_FILL_FORK": "1"},                 # background fill on the caller's stream (no side stream)
    "fill_fork_late": {"PS_FILL_FORK_LATE": "1"},             # background fill beside the forward rasterizer, not the binning
    "loss_tiled": {"PS_LOSS_TILED": "1"},                     # round-1 32 x 32 tile loss kernels
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_parity_slice_under_switch(name):
    env = dict(os.environ)
    env.update(VARIANTS[name])
    target, pick = ("test_gpu_loss.py", "view_loss") if name.startswith("loss_") else ("test_gpu_parity.py", SLICE)
    if "fill_fork" in name:  # only calls too large for the sync-free path fork the fill: the 24-view batch at full N
        pick = SLICE + " or test_batch_invariance_full_size_c2"
    r = subprocess.run([sys.executable, "-m", "pytest", str(ROOT / "tests" / target), "-m", "gpu", "-q", "-x",
                        "-k", pick, "-p", "no:cacheprovider"], cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
