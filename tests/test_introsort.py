"""CPU-only: the std::sort emulation used by the device oct-tree (visual_sgraphs_b200/csrc/introsort.cuh)
produces exactly libstdc++'s permutation, including the order of equivalent elements and the heapsort
fallback."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_introsort_emulation_matches_std_sort(tmp_path):
    exe = str(tmp_path / "introsort_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "introsort_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK")
