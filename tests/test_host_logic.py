"""CPU tests of the host-visible logic of the product sources: the run-length envelope and the closed-form
oscillator (klang_b200/csrc/kb_prims.cuh, compiled here as plain C++) are bit-identical to the per-tick forms."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_run_length_envelope_and_closed_form_oscillator():
    exe = os.path.join(tempfile.mkdtemp(prefix="kb_host_"), "env_run_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-x", "c++",
                           os.path.join(ROOT, "tests", "host", "env_run_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches" in out.stdout
