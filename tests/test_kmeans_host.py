"""CPU: the host side of the guide tree (kalign_b200/csrc/kb_kmeans.h, plain C++).  The product builds
it -O3 -mavx2 (kalign_b200/csrc/Makefile); the bisection tree it produces must be the one the plain
-O2 build produces, for every thread count (no float operation may be reordered: the reference's
summation orders, lib/src/bisectingKmeans.c and euclidean_dist.c, decide the guide tree)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAG_SETS = {"O2": ["-O2"], "O3_avx2": ["-O3", "-mavx2"]}


@pytest.fixture(scope="module")
def drivers(tmp_path_factory):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    d = tmp_path_factory.mktemp("kmeans")
    out = {}
    for tag, flags in FLAG_SETS.items():
        exe = str(d / ("kmeans_" + tag))
        cmd = [cxx] + flags + ["-ffp-contract=off", "-fopenmp", "-std=c++17", "-I", os.path.join(ROOT, "kalign_b200", "csrc"),
                               os.path.join(ROOT, "tests", "kmeans_driver.cpp"), "-o", exe]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert p.returncode == 0, p.stdout
        out[tag] = exe
    return out


def _tree(exe, n, seed, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    p = subprocess.run([exe, str(n), str(seed)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stdout
    m = re.search(r"clusters=(\d+) covered=(\d+) tree=([0-9a-f]{16})", p.stdout)
    assert m and int(m.group(2)) == n, p.stdout
    return m.group(3), int(m.group(1))


@pytest.mark.parametrize("n,seed", [(60, 1), (700, 2), (5000, 3), (30000, 4)])
def test_tree_independent_of_flags_and_threads(drivers, n, seed):
    want, ncl = _tree(drivers["O2"], n, seed, 1)
    assert ncl >= 1
    for tag in FLAG_SETS:
        for threads in (1, 4, 16):
            got, _ = _tree(drivers[tag], n, seed, threads)
            assert got == want, (tag, threads)


def test_product_makefile_builds_the_tree_code_with_these_flags():
    mk = open(os.path.join(ROOT, "kalign_b200", "csrc", "Makefile")).read()
    assert re.search(r"kb_msa\.o:\s*HOSTOPT\s*:=\s*-O3,-mavx2", mk)
    assert "-ffp-contract=off" in mk and "-ffast-math" not in mk
