"""CPU checks of the packed (two int16 per register) arithmetic of the kernels: the lifting steps of
csrc/ric_swar.cuh against a transcription of the reference's 1-D 9/7 lifting, and the packed encode quantiser of
csrc/ric_quant_pk.cuh against the oracle's buildTree restatement.  The headers are __host__ __device__, so the very
code the kernels run is executed here on the CPU (tests/cpp/*.cu, built by nvcc as host programs)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BUILD = os.path.join(ROOT, "build")


def _build_and_run(name, extra=()):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    cmd = ["nvcc", "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
           os.path.join(ROOT, "tests", "cpp", name + ".cu"), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1000:])
    return r.stdout


def test_packed_lifting_matches_reference_formulas():
    out = _build_and_run("swar_test")
    assert "swar_test: ok" in out


def test_packed_quantiser_matches_oracle():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    orc = os.path.join(ROOT, "oracle")
    out = _build_and_run("quant_pk_test", ["-L" + orc, "-lric_oracle", "-Xlinker", "-rpath=" + orc])
    assert "quant_pk_test: ok" in out
