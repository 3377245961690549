"""The polygon-free overlap area (csrc/dm2_aa_fast.cuh) against the oracle's transcription of the reference's polygon
clip (cuda_impl/aa.h:96-504), on the CPU: tests/native/aa_fast_check.cpp compiles the SAME header the CUDA kernels
include (host build of its float arithmetic) and checks its contract on random and adversarial (triangle, pixel) pairs -
whenever it answers, the reference answers too (no error code), areas equal to 2e-6, gradients equal to the reference's
Jacobian chain or, where that chain is dominated by the rounding of its polygon vertices, to the fp64 derivative.
The GPU side (both settings of dm2_set_exact_clip agree, parity with the oracle / the live reference) is in
tests/test_gpu_parity.py and tests/test_gpu_fullsize_parity.py."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    from oracle import oracle as orc
    orc.build()
    lib_dir = os.path.join(ROOT, "oracle", "_build")
    exe = str(tmp_path_factory.mktemp("aaf") / "aa_fast_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math",
                           os.path.join(ROOT, "tests", "native", "aa_fast_check.cpp"), "-o", exe,
                           "-L" + lib_dir, "-ldm2_oracle", "-Wl,-rpath," + lib_dir, "-pthread"])
    return exe


@pytest.mark.parametrize("seed", [1, 2])
def test_fast_overlap_area_contract(checker, seed):
    out = subprocess.run([checker, "150000", str(seed)], capture_output=True, text=True)
    rep = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and rep["violations"] == 0, (rep, out.stderr[-2000:])
    assert rep["max_area_diff"] <= 2e-6, rep
    # it must actually answer: > 99 % of the partially covered pairs of the benchmark-like scenes, and a fair share
    # of the adversarial ones (where declining is the point)
    fr = rep["answered_frac_by_mode"]
    assert fr[8] > 0.99 and fr[9] > 0.985 and fr[0] > 0.95, rep
    assert rep["answered"] > 300000, rep


def test_face_safety_word(checker):
    """Per-face thresholds and flags of the fast path (aaf_face_safety): bf16 packing rounds up, parallel / non-finite /
    huge / needle faces and inconsistent edge data are never answered for, a near-integer coordinate switches the
    per-pair vertex test on."""
    out = subprocess.run([checker, "--unit"], capture_output=True, text=True)
    assert out.returncode == 0 and json.loads(out.stdout.strip().splitlines()[-1])["unit_failures"] == 0, out.stdout
