"""Schedule A/B checks that need a process of their own (switches the library reads once per process)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_fm_tiled_kernel_matches_lane_per_voice_kernel():
    """kb_fm_tiled_kernel (envelopes run-length in one warp, every sample's operator chain in parallel) against the lane-per-voice
    kernel, bit for bit: 63 voices (ragged last CTA, half of them idle), ragged blocks, releases, a control change and
    re-triggers between blocks, per-voice output, the instance mix and the note stages (tools/fm_tiled_probe.py)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fm_tiled_probe.py")], capture_output=True, text=True, timeout=240)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-400:] + out.stderr[-400:]
    d = json.loads(lines[-1])
    assert d["equal"] and d["stages_equal"] and d["ended"] > 0 and d["peak"] > 0.01


def test_c2_voice_kernel_variants_are_bit_identical_and_the_hand_over_checks_itself(tmp_path):
    """tools/c2_ab.py in child mode renders the same 6 blocks (re-triggers, releases, a ragged block) of per-voice streams with the lock-step
    kernel (KB_TILE_LAYOUT=2), the polled-counter kernel (3), the mbarrier kernel (4, the default) and the mbarrier kernel's self-checking
    instantiation (KB_C2_VARIANT=128: every ring buffer carries a tile stamp that consumers compare before and after they read; a mismatch
    poisons the output with NaNs).  All four are the same bits, and the checked run contains no NaN."""
    import numpy as np
    outs = {}
    for name, env in (("lockstep", {"KB_TILE_LAYOUT": "2"}), ("polled", {"KB_TILE_LAYOUT": "3", "KB_TILE_G": "7"}),
                      ("mbarrier", {"KB_TILE_LAYOUT": "4", "KB_TILE_G": "7"}), ("mbarrier_checked", {"KB_TILE_LAYOUT": "4", "KB_TILE_G": "7", "KB_C2_VARIANT": "128"}),
                      ("mbarrier_checked_g8", {"KB_TILE_LAYOUT": "4", "KB_TILE_G": "8", "KB_C2_VARIANT": "128"})):
        p = str(tmp_path / f"{name}.npy")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "c2_ab.py"), "child", p], env={**os.environ, **env}, capture_output=True, text=True, timeout=240)
        assert r.returncode == 0, r.stderr[-600:]
        outs[name] = np.load(p)
    ref = outs["lockstep"]
    assert np.count_nonzero(ref) > 1e6
    for name, x in outs.items():
        assert not np.isnan(x).any(), f"{name}: the hand-over protocol reported a mismatch"
        assert x.shape == ref.shape and np.array_equal(x.view(np.uint32), ref.view(np.uint32)), name
