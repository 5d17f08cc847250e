"""Host build of the kernels' constitutive header (dumux_b200/csrc/physics.cuh is host+device code): the fused evaluation
law_eval3 / table_interp2 / PowBase / div_by used by the tile assembly kernel must return the bits of the one-at-a-time
functions (material/fluidmatrixinteractions/2p/materiallaw.hh:104-243 call order) and of the oracle's det_pow."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_law_eval3_bits_match_separate_curves(tmp_path):
    exe = str(tmp_path / "law_eval_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-mfma",
                           os.path.join(ROOT, "tests", "cpp", "law_eval_check.cpp"), "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 mismatches" in p.stdout
