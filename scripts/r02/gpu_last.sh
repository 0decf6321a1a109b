#!/bin/bash
# last seconds of the round's GPU budget: the search tests that take the finely split fp16 second chance, then
# whatever else of the parity file fits
mkdir -p gpurun_out
timeout 19 python -m pytest tests/test_gpu_parity.py -q -x --timeout=18 -k "config2_full or aligned or tables_in_one_pass_bit_exact or affine or smoke or reference_own" > gpurun_out/last_pytest2.log 2>&1
echo "pytest exit $?" >> gpurun_out/last_pytest2.log; tail -3 gpurun_out/last_pytest2.log
