#!/bin/bash
# round 2, call AB: the full GPU suite + smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2ab_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ab_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2ab_smoke.log
