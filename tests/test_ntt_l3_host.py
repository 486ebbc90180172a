"""CPU check of the NTT's lazily reduced 3-limb arithmetic (vectorx_b200/csrc/ntt_l3.cuh): the header compiles for the host
(plain C++ restatement of the same limb arithmetic, no PTX), so the fold formulas through phi = 2^32, the products by 2^s
and the multiplication-free 2^R-point DIF network are compared with a naive Goldilocks DFT without a GPU.  The device
path's PTX carry chains implement the same 96-bit operations and are covered by the -m gpu parity tests."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_l3_dft_network_on_the_host(tmp_path):
    exe = str(tmp_path / "l3_dft_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "vectorx_b200", "csrc"), "-x", "c++",
                    os.path.join(ROOT, "tests", "cpu", "l3_dft_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
