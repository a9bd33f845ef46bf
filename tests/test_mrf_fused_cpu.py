"""CPU check of the fused MRF kernel's geometry (no GPU): tests/cpu/mrf_fused_emul.cpp re-enacts the
kernel's shared-memory / tensor-memory data movement with the library's real host-side packing,
scatter tables and launch plan (zerovox.cpp_b200/csrc/mrf_fused_host.h) and compares with a direct
evaluation of HiFiGANResidualBlock (/root/reference/src/hifigan.cpp:97-183)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emul") / "mrf_fused_emul")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpu", "mrf_fused_emul.cpp")], check=True)
    return exe


# (CH, k, T, min_eff, ncol): window interior / edges / utterance shorter than a window / split chains
CASES = [
    (32, 3, 1100, 0.0, 256), (32, 11, 1300, 0.0, 128), (32, 7, 37, 0.0, 128), (32, 3, 481, 0.0, 128),
    (64, 7, 700, 0.0, 256), (64, 11, 600, 0.8, 256), (64, 3, 5, 0.0, 256),
    (128, 3, 300, 0.8, 256), (128, 11, 280, 0.8, 256), (128, 7, 1, 0.8, 256),
]


@pytest.mark.parametrize("ch,k,T,min_eff,ncol", CASES)
def test_fused_geometry_matches_direct_conv(emul, ch, k, T, min_eff, ncol):
    r = subprocess.run([emul, str(ch), str(k), str(T), str(min_eff), str(ncol)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
