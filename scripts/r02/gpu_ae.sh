#!/bin/bash
# finalize_kernel with several keys per selection round; second chance on the fp16 collect kernel.  Parity of every
# search path, timing with the second-chance knob both ways, ncu of finalize.
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=500 -k "search or multi_batch or config2_full or candidates or adversarial or config4 or config1 or fixed_mode or query_mode or golden or reference_own" > $O/ae_pytest.log 2>&1
echo "pytest exit $?" >> $O/ae_pytest.log; tail -3 $O/ae_pytest.log
for r in 1 0; do FWAV_UMMA_RETRY16=$r timeout 200 python scripts/time_topk.py 1.0 umma 4 2>/dev/null | cut -c1-330; done | tee $O/ae_time.txt
FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 0 2> $O/ae_verbose.txt > /dev/null; grep fwav $O/ae_verbose.txt | cut -c1-250
FWAV_UMMA_VERBOSE=1 timeout 300 python scripts/time_topk.py 5.0 umma 1 2> $O/ae_verbose5.txt | cut -c1-330; grep fwav $O/ae_verbose5.txt | cut -c1-250
timeout 600 ncu --set full --clock-control none -k regex:"finalize_kernel" -c 1 -f -o $O/ae_finalize python scripts/time_topk.py 1.0 umma 0 > $O/ae_ncu.log 2>&1
tail -1 $O/ae_ncu.log | cut -c1-200
