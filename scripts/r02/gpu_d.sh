#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "collect_layouts and (sets or thin2) or multi_batch or search_paths or odd_shapes or sharded_driver or decode" > $O/r02d_tests.txt 2>&1
echo "tests: rc=$? $(tail -1 $O/r02d_tests.txt)"; grep -E "^(FAILED|ERROR)|Error" $O/r02d_tests.txt | head
rm -f $O/r02d_timing.txt
for lib in libfwav_b200.so libfwav_b200_mask.so; do
 for lay in sets thin2; do
  echo "== $lib FWAV_UMMA_COLLECT=$lay (config 2)" >> $O/r02d_timing.txt
  FWAV_LIB=$PWD/audio-compression_b200/fwav_b200/$lib FWAV_UMMA_COLLECT=$lay timeout 200 python scripts/time_topk.py 1.0 umma 3 2>/dev/null | cut -c1-330 >> $O/r02d_timing.txt
 done
done
echo "== default lib, full split (precise)" >> $O/r02d_timing.txt
FWAV_UMMA_MODE=precise timeout 200 python scripts/time_topk.py 1.0 umma 2 2>/dev/null | cut -c1-330 >> $O/r02d_timing.txt
cat $O/r02d_timing.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel --launch-skip 3 --launch-count 1 \
  -o $O/r02d_prof_collect_sets -f python scripts/time_topk.py 1.0 umma 1 > $O/r02d_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/r02d_ncu.log
