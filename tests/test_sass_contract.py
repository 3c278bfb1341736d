"""CPU: arithmetic contract of the DP kernels, checked on the SASS that ships (no GPU needed).

The reference rounds every multiply and every add separately (its build has no FMA,
CMakeLists.txt:192-193), and bit-identical paths depend on it.  nvcc -fmad=false keeps scalar
code uncontracted, but ptxas does contract the PACKED forms (mul.rn.f32x2 + add.rn.f32x2 -> FFMA2)
whatever the flag says -- so the sweep kernels must contain no fused multiply-add at all, and the
thread-per-box kernel (which inlines the meet-up's IEEE division, an FFMA sequence) no packed one."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "kalign_b200", "libkalign_b200.so")


def _functions():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", SO], stdout=subprocess.PIPE, text=True, check=True).stdout
    funcs = {}
    for chunk in re.split(r"\n\s*Function : ", out)[1:]:
        name, body = chunk.split("\n", 1)
        funcs[name.strip()] = body
    return funcs


def test_sweep_kernels_have_no_fused_multiply_add():
    funcs = _functions()
    sweep = {n: b for n, b in funcs.items() if "kb_sweep_kernel" in n}
    assert len(sweep) == 3, sorted(funcs)
    for name, body in sweep.items():
        assert "FADD2" in body, "packed additions expected in " + name
        assert not re.search(r"\bFFMA2?\b", body), "fused multiply-add in " + name


def test_no_packed_fma_anywhere_in_dp():
    funcs = _functions()
    for name, body in funcs.items():
        if "kb_dp_cu" in name or "kb_small_kernel" in name:
            assert "FFMA2" not in body, "packed fused multiply-add in " + name


def test_profile_column_records_are_staged_by_bulk_async_copies():
    """north star: "TMA-staged profile columns in shared memory".  The 5-letter profile-profile strips
    stage their packed column records with cp.async.bulk (the TMA engine's linear form: one elected lane
    per 16-column block) signalled on an mbarrier -- UBLKCP / SYNCS.ARRIVE.TRANS64 / a phase-checking
    wait must be in the SASS of every sweep kernel family."""
    funcs = _functions()
    sweeps = {n: b for n, b in funcs.items() if "kb_sweep_kernel" in n}
    assert len(sweeps) == 3
    for name, body in sweeps.items():
        assert "UBLKCP" in body, name
        assert "SYNCS.ARRIVE.TRANS64" in body, name
        assert "SYNCS.PHASECHK.TRANS64.TRYWAIT" in body, name
