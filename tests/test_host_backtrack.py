"""The host traceback utilities of the reference's backtrack.h (dpx_gpu_genomics_project_b200/host/backtrack.{h,cpp}): driven the
way the reference's CUDA mains drive them (tests/backtrack_driver.cpp fills plain direction matrices), their stdout must equal the
committed golden text of the reference classes; where the reference tree is present the same driver is also linked against the
reference's own c++/backtrack.cpp and the two outputs compared."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dpx_gpu_genomics_project_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/c++"
W = {"LNW": ["3", "-1", "-2", "0"], "ANW": ["3", "-1", "-3", "-1"], "LSW": ["3", "-1", "-2", "0"]}


@pytest.fixture(scope="module")
def drivers(tmp_path_factory):
    d = tmp_path_factory.mktemp("bt")
    ours = str(d / "ours")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", HOST, os.path.join(ROOT, "tests", "backtrack_driver.cpp"), os.path.join(HOST, "backtrack.cpp"), "-o", ours], check=True)
    ref = None
    if os.path.exists(os.path.join(REF, "backtrack.cpp")):
        ref = str(d / "ref")
        subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-I", REF, os.path.join(ROOT, "tests", "backtrack_driver.cpp"),
                        os.path.join(REF, "backtrack.cpp"), os.path.join(REF, "printLock.cpp"), "-o", ref], check=True)
    return ours, ref


@pytest.mark.parametrize("name", ["adversarial", "shapes", "cfg1_small"])
@pytest.mark.parametrize("algo", ["LNW", "ANW", "LSW"])
def test_host_backtrack_functions_print_the_reference_bytes(drivers, name, algo):
    ours, ref = drivers
    path = os.path.join(GOLD, f"{name}.in.txt")
    out = subprocess.run([ours, algo, path] + W[algo], check=True, capture_output=True).stdout
    assert out == open(os.path.join(GOLD, f"{name}.{algo}.out.txt"), "rb").read()
    if ref:
        assert out == subprocess.run([ref, algo, path] + W[algo], check=True, capture_output=True).stdout
