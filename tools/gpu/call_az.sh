#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) > gpurun_out/r2az_pytest.log; cat gpurun_out/r2az_pytest.log
