#!/bin/bash
# last GPU call of the round (3.7 GPU-minutes left): the full GPU suite + smoke on the final tree
mkdir -p gpurun_out
timeout 175 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/last_pytest.log 2>&1; echo "pytest rc=$?"
grep -n "^FAILED\|passed\|failed\|Error" gpurun_out/last_pytest.log | tail -8
timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
