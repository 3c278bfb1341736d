#!/bin/bash
# guide-tree seam of the drop-in: CLI parity cases, seeded / golden tests, CLI wall time on C3 / C4 files
O=gpurun_out/r2q; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_dropin.py tests/test_seeded.py tests/test_golden.py tests/test_cmake_package.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -12 $O/pytest.log
timeout 120 python tools/gpu_file_e2e.py > $O/file_e2e.json 2> $O/file_e2e.err; echo "file e2e exit $?"; cat $O/file_e2e.json
