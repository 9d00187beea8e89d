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
