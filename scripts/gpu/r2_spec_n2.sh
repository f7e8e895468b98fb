#!/bin/bash
cd "$(dirname "$0")/../.."
for nn in 128 160; do for s in 1 0; do
FSB_SPECULATE=$s python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2981$s scripts/gpu/host_overheads.py $nn 2>&1 | grep " cg:" | cut -c1-150 | sed "s/^/spec=$s /"
done; done
