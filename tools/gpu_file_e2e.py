"""file -> file timing on a GPU box: kb200_kalign_file (the product's FASTA reader, alignment, writer in one C call)
and the drop-in CLI (integration/_out/kalign: the reference's main() over the seams) on C4 / C3 sized FASTA files."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kalign_b200 import _lib, synth  # noqa: E402

out = {}
ctx = _lib.Context(0)
with tempfile.TemporaryDirectory() as d:
    for wl, type_, K, flags in (("C4", 8, 0, ["--fast"]), ("C3", 2, 5, ["--type", "rna"])):
        seqs = synth.config(wl)
        fa = os.path.join(d, wl + ".fa")
        with open(fa, "w") as f:
            for i, s in enumerate(seqs):
                f.write(">s%d\n" % i)
                for j in range(0, len(s), 60):
                    f.write(s[j:j + 60] + "\n")
        o1 = os.path.join(d, wl + ".gpu.afa")
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            ctx.kalign_file(fa, o1, n_threads=16, type_=type_, consistency=K, weight=2.0)
            ts.append(time.perf_counter() - t0)
        o2 = os.path.join(d, wl + ".cli.afa")
        t0 = time.perf_counter()
        p = subprocess.run([os.path.join(ROOT, "integration", "_out", "kalign"), "-i", fa, "-o", o2, "-n", "16"] + flags,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        tcli = time.perf_counter() - t0
        same = p.returncode == 0 and open(o1, "rb").read() == open(o2, "rb").read()
        out[wl] = {"nseq": len(seqs), "input_bytes": os.path.getsize(fa), "output_bytes": os.path.getsize(o1),
                   "kalign_file_seconds": ts, "dropin_cli_wall_seconds_incl_process_start_and_cuda_init": tcli,
                   "cli_rc": p.returncode, "cli_output_identical_to_kalign_file": same}
ctx.close()
print(json.dumps(out))
