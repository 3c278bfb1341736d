"""CPU: the drop-in build (integration/): the reference's own host code + one added file, linked
against libkalign_b200.so.  Checks the public C API of lib/include/kalign/kalign.h is exported
unchanged, that the three seam calls were redirected at link time, and that without a GPU the
alignment path fails loudly (no CPU fallback)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "integration", "_out")
LIB = os.path.join(OUT, "libkalign.so.3")
CLI = os.path.join(OUT, "kalign")

pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="integration/_out not built (needs the reference sources)")

# lib/include/kalign/kalign.h:36-109
KALIGN_H_API = ["kalign", "kalign_run", "kalign_run_seeded", "kalign_read_input", "kalign_write_msa",
                "kalign_free_msa", "kalign_arr_to_msa", "kalign_msa_to_arr", "kalign_msa_compare",
                "kalign_check_msa", "kalign_sort_msa", "kalign_ensemble", "kalign_essential_input_check"]


def _nm():
    out = subprocess.run(["nm", "-D", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    defined = {l.split()[-1] for l in out.splitlines() if " T " in l}
    undefined = {l.split()[-1] for l in out.splitlines() if " U " in l}
    return defined, undefined


def test_public_api_and_seams():
    defined, undefined = _nm()
    for f in KALIGN_H_API:
        assert f in defined, f
    for w in ("__wrap_d_estimation", "__wrap_anchor_consistency_build", "__wrap_create_msa_tree", "__wrap_compute_aln_pairwise_dist",
              "__wrap_build_tree_kmeans", "__wrap_build_tree_kmeans_noisy"):
        assert w in defined, w
    for f in ("kb200_guide_tree", "kb200_fasta_read", "kb200_fasta_write", "kb200_aln_pairwise_dist", "kb200_distances_on", "kb200_seqs_upload", "kb200_anchor_posmaps", "kb200_select_anchors", "kb200_align_tree_conf", "kb200_ctx_create"):
        assert f in undefined, f       # resolved by libkalign_b200.so at load time


def test_run_seeded_calls_the_seams():
    if shutil.which("objdump") is None:
        pytest.skip("objdump not on PATH")
    dis = subprocess.run(["objdump", "-d", "--no-show-raw-insn", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    body = dis.split("<kalign_run_seeded>:")[1].split("\n\n")[0]
    assert "__wrap_create_msa_tree" in body and "__wrap_anchor_consistency_build" in body
    assert "__wrap_build_tree_kmeans" in body and "__wrap_build_tree_kmeans_noisy" in body
    assert "<build_tree_kmeans@plt>" not in body
    assert "<create_msa_tree@plt>" not in body and "<anchor_consistency_build@plt>" not in body
    assert "call" in dis and "<d_estimation@plt>" not in dis
    # the realign loop reaches the GPU for its N x N identity distances as well
    body = dis.split("<kalign_run_realign>:")[1].split("\n\n")[0]
    assert "__wrap_compute_aln_pairwise_dist" in body and "__wrap_create_msa_tree" in body
    assert "<compute_aln_pairwise_dist@plt>" not in dis


def test_cli_fails_loudly_without_gpu(tmp_path):
    from kalign_b200 import _lib, synth
    if _lib.load().kb200_device_count() > 0:
        pytest.skip("GPU present")
    fa = tmp_path / "in.fa"
    fa.write_text("".join(">s%d\n%s\n" % (i, s) for i, s in enumerate(synth.family(6, 50, synth.PROTEIN, seed=3))))
    p = subprocess.run([CLI, "-i", str(fa), "-o", str(tmp_path / "out.afa")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode != 0
    assert "no CPU fallback" in p.stdout
    assert not (tmp_path / "out.afa").exists()
