"""GPU test of the C++ drop-in classes (visual_sgraphs_b200/shim): compiles tests/cpp/shim_test.cpp, which calls
VS_GRAPHS::ORBextractor / ORBmatcher the way the reference's Frame and Tracking do and checks every result
against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from visual_sgraphs_b200.synth import synth_frame

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_against_oracle(tmp_path, oracle):
    pkg = os.path.join(ROOT, "visual_sgraphs_b200")
    subprocess.check_call(["make", "-C", os.path.join(pkg, "shim")], stdout=subprocess.DEVNULL)
    exe = str(tmp_path / "shim_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "shim_test.cpp"),
                           "-L" + pkg, "-lvsg_orb_shim", "-lvsg_cuda", "-L" + os.path.join(ROOT, "oracle"), "-lorb_oracle",
                           "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    a = synth_frame(11)
    rng = np.random.default_rng(12)
    b = np.roll(a, (5, 9), (0, 1))
    b = np.clip(b.astype(np.int16) + rng.integers(-3, 4, b.shape), 0, 255).astype(np.uint8)
    a.tofile(str(tmp_path / "a.raw"))
    b.tofile(str(tmp_path / "b.raw"))
    out = subprocess.run([exe, str(tmp_path / "a.raw"), str(tmp_path / "b.raw")], capture_output=True, text=True)
    print(out.stdout[:3000])
    assert out.returncode == 0, out.stdout[:3000] + out.stderr[-2000:]
    assert "SHIM TEST OK" in out.stdout
