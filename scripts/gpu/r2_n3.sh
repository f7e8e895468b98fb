#!/bin/bash
# narray: FVM assembler and slabs over in-process ranks; example programs
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_examples.py tests/test_zz_narray_gpu.py tests/test_zz_local_ranks_gpu.py -q > $O/r2_n3_pytest.log 2>&1; tail -25 $O/r2_n3_pytest.log | cut -c1-300
examples/_build/diffusion 64 examples/equilibrium_diffusion/diffusion.cfg
examples/_build/diffusion 256 examples/equilibrium_diffusion/diffusion.cfg
