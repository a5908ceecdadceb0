"""Runs the blind-rotation kernel's per-lane phase functions on the CPU (they are
__host__ __device__) against the oracle's exact-integer path: validates the twisted FFT,
the key layout, the rotation/decomposition indexing and the warp-pair split without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="nvcc not available")
def test_lane_emulation_matches_exact_arithmetic(oracle, tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "host_emul")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-w", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "host_emul.cu"), "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                    "-Xlinker", "-rpath," + os.path.join(ROOT, "oracle")], check=True, cwd=ROOT)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "HOST EMULATION OK" in r.stdout
