"""The CMake drop-in surface (integration/CMakeLists.txt, built by __graft_entry__.build() into
integration/_cmake): library targets kalign (SOVERSION 3) / kalign_static, alias kalign::kalign,
kalignConfig.cmake, CLI `kalign` -- what lib/CMakeLists.txt:78-103,135 and src/CMakeLists.txt:25-40
give a consumer of the reference.  integration/consumer is a separate CMake project that does
find_package(kalign 3 CONFIG) + target_link_libraries(app kalign::kalign) and calls kalign()."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CM = os.path.join(ROOT, "integration", "_cmake")
INST = os.path.join(CM, "install")
CONSUMER = os.path.join(CM, "consumer", "kalign_consumer")

pytestmark = pytest.mark.skipif(not os.path.exists(CONSUMER), reason="integration/_cmake not built (needs the reference sources)")

SEQS = ["GKGDPKKPRGKMSSYAFFVQTSREEHKKKHPDASVNFSEFSKKCSERWKTMSAKEKGKFEDMAKADKARYEREMKTYIPPKGE",
        "MQDRVKRPMNAFIVWSRDQRRKMALENPRMRNSEISKQLGYQWKMLTEAEKWPFFQEAQKLQAMHREKYPNYKYRPRRKAKMLPK",
        "MKKLKKHPDFPKKPLTPYFRFFMEKRAKYAKLHPEMSNLDLTKILSKKYKELPEKKKMKYIQDFQREKQEFERNLARFREDHPDLIQNAKK",
        "MHIKKPLNAFMLYMKEMRANVVAESTLKESAAINQILGRRWHALSREEQAKYYELARKERQLHMQLYPGWSARDNYGKKKKRKREK"]


def test_installed_package_layout():
    lib = os.path.join(INST, "lib")
    assert os.path.islink(os.path.join(lib, "libkalign.so"))
    assert os.path.realpath(os.path.join(lib, "libkalign.so.3")).endswith("libkalign.so.3.5.1")
    assert os.path.exists(os.path.join(lib, "libkalign_static.a"))
    assert os.path.exists(os.path.join(INST, "bin", "kalign"))
    assert os.path.exists(os.path.join(INST, "include", "kalign", "kalign.h"))
    cfg = os.path.join(lib, "cmake", "kalign")
    for f in ("kalignConfig.cmake", "kalignConfigVersion.cmake", "kalignTargets.cmake"):
        assert os.path.exists(os.path.join(cfg, f)), f
    assert "kalign::kalign" in open(os.path.join(cfg, "kalignTargets.cmake")).read()
    soname = subprocess.run(["objdump", "-p", os.path.join(lib, "libkalign.so.3")], stdout=subprocess.PIPE, text=True).stdout
    assert "SONAME" in soname and "libkalign.so.3" in soname
    # the installed header is the reference's, byte for byte
    ref_h = "/root/reference/lib/include/kalign/kalign.h"
    if os.path.exists(ref_h):
        assert open(ref_h, "rb").read() == open(os.path.join(INST, "include", "kalign", "kalign.h"), "rb").read()


def test_consumer_links_kalign_kalign():
    out = subprocess.run(["ldd", CONSUMER], stdout=subprocess.PIPE, text=True).stdout
    assert "libkalign.so.3" in out and "libkalign_b200.so" in out
    src = open(os.path.join(ROOT, "integration", "consumer", "CMakeLists.txt")).read()
    assert "find_package(kalign" in src and "kalign::kalign" in src


def test_consumer_fails_loudly_without_gpu():
    from kalign_b200 import _lib
    if _lib.load().kb200_device_count() > 0:
        pytest.skip("GPU present")
    p = subprocess.run([CONSUMER], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 3 and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_consumer_alignment_equals_reference():
    import kbind
    if not kbind.have_ref():
        pytest.skip("oracle/_ref missing")
    assert len(set(len(s) for s in SEQS)) == len(SEQS)      # no equal-length ties (array-mode names are undefined)
    p = subprocess.run([CONSUMER], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    got = p.stdout.split()
    run = kbind.RefRun(SEQS, n_threads=2, type_=3, consistency=0)      # kalign() == kalign_run(..., refine none), no consistency
    want = run.aligned()
    run.close()
    assert got == want


@pytest.mark.gpu
def test_installed_cli_equals_reference_cli(tmp_path):
    import kbind
    from kalign_b200 import synth
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "kalign_ref")
    if not os.path.exists(ref_cli):
        pytest.skip("oracle/_ref missing")
    fa = tmp_path / "in.fa"
    synth.write_fasta(str(fa), synth.family(40, 120, synth.PROTEIN, seed=77))
    a, b = tmp_path / "gpu.afa", tmp_path / "ref.afa"
    p = subprocess.run([os.path.join(INST, "bin", "kalign"), "-i", str(fa), "-o", str(a)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
    q = subprocess.run([ref_cli, "-i", str(fa), "-o", str(b)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert q.returncode == 0, q.stdout
    assert a.read_bytes() == b.read_bytes()
