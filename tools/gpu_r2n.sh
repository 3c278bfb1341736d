#!/bin/bash
# seeded / ensemble entry points: the GPU parity suite again (new: tests/test_seeded.py) + smoke
O=gpurun_out/r2n; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -25 $O/pytest.log
