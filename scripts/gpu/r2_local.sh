#!/bin/bash
# in-process rank groups on one GPU: which solver stalls, and the test file
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift; echo "=== $tag"; ( "$@" ) > $O/r2_local_$tag.log 2>&1; echo "rc=$?"; grep -v "^\[fsb" $O/r2_local_$tag.log | tail -${TAIL:-8} | cut -c1-400; }
run pytest timeout 600 python -m pytest tests/test_zz_local_ranks_gpu.py -q
run pytest2 timeout 600 python -m pytest tests/test_zz_local_ranks_gpu.py -q
run smoke timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
