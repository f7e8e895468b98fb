#!/bin/bash
# armed launches (FSB_OPT_SPECULATE): the solver tests, then iteration times with and without
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2_spec_pytest.log 2>&1; tail -6 $O/r2_spec_pytest.log | cut -c1-300
for s in 1 0; do echo "FSB_SPECULATE=$s"; FSB_SPECULATE=$s timeout 120 python scripts/gpu/host_overheads.py 64; FSB_SPECULATE=$s timeout 120 python scripts/gpu/host_overheads.py 256; done 2>&1 | tee $O/r2_spec.txt
