#!/bin/bash
# finalize_kernel: candidate indices first, then scores (no part-by-part walk).  Parity of the search paths, timing, ncu.
set +e
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_parity.py -q -x --timeout=400 -k "search or multi_batch or config2_full or candidates or adversarial or config4" > $O/ad_pytest.log 2>&1
echo "pytest exit $?" >> $O/ad_pytest.log; tail -3 $O/ad_pytest.log
timeout 200 python scripts/time_topk.py 1.0 umma 4 2>/dev/null | cut -c1-330 | tee $O/ad_time.txt
timeout 600 ncu --set full --clock-control none -k regex:"finalize_kernel" -c 1 -f -o $O/ad_finalize python scripts/time_topk.py 1.0 umma 0 > $O/ad_ncu.log 2>&1
tail -1 $O/ad_ncu.log | cut -c1-200
