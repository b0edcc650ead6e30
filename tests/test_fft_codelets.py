"""The in-register FFT codelets and the column pass bodies against a naive double-precision DFT, on the CPU (tests/host/fft_emu_test.cpp):
forward (with the no-reorder frequency map of Cooley-Tukey and prime-factor stages), inverse, and the fused z convolution."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# plans that cover every stage radix in use: 4/8/16 (prime powers), 5, 6, 9, 10, 12, 15, 18, 20 (prime-factor splits) and 2- and 3-stage plans
SUBSET = {36, 50, 72, 120, 128, 180, 216, 400, 432, 540, 900}


def test_codelets_against_naive_dft(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "multiview-reconstruction_b200", "csrc"))
    import gen_lengths
    plans = [p for p in gen_lengths.PLANS if p[0] in SUBSET]
    assert len(plans) == len(SUBSET)
    define = "-DMVD_TEST_PLANS(X)=" + " ".join("X(%d,%d,%d,%d,%d,%d)" % p for p in plans)
    exe = str(tmp_path / "fft_emu_test")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-DMVD_HOST_EMU", "-ffp-contract=off", define, "-I", os.path.join(ROOT, "multiview-reconstruction_b200", "csrc"),
                        os.path.join(ROOT, "tests", "host", "fft_emu_test.cpp"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:]
    assert r.stdout.count(" ok\n") == len(plans)
