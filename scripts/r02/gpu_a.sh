#!/bin/bash
# Round 2, GPU call A: parity of every collect layout and of the compact split (run-or-delete), the new parity-gap
# tests, then timings of the layouts on config 2, the microbenchmark (f16 / tf32 instruction peaks) and one ncu
# capture of the fat collect kernel.  Everything lands in gpurun_out/r02a_*.
set +e
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > $O/r02a_smi.txt
for lay in thin4 thin2 fat4 fat2 sets; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "collect_layouts and $lay" > $O/r02a_layout_$lay.txt 2>&1
  echo "layout $lay: rc=$? $(tail -1 $O/r02a_layout_$lay.txt)"
done
FWAV_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "compact" > $O/r02a_compact.txt 2>&1
echo "compact: rc=$? $(tail -1 $O/r02a_compact.txt)"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_batch" > $O/r02a_multibatch.txt 2>&1
echo "multi_batch: rc=$? $(tail -1 $O/r02a_multibatch.txt)"
FWAV_TEST_EXPERIMENTAL=1 timeout 1200 python -m pytest tests -m gpu -q --durations=12 > $O/r02a_pytest.txt 2>&1
echo "full suite: rc=$? $(tail -1 $O/r02a_pytest.txt)"
# ---- timings: the search alone on config 2, three repetitions each ----
rm -f $O/r02a_timing.txt
for lay in sets thin4 thin2 fat4 fat2; do
  echo "== FWAV_UMMA_COLLECT=$lay (config 2)" >> $O/r02a_timing.txt
  FWAV_UMMA_COLLECT=$lay FWAV_UMMA_VERBOSE=1 timeout 200 python scripts/time_topk.py 1.0 umma 3 > $O/r02a_t.out 2> $O/r02a_t.err
  grep "fwav\]" $O/r02a_t.err | tail -3 | cut -c1-220 >> $O/r02a_timing.txt
  cut -c1-400 $O/r02a_t.out >> $O/r02a_timing.txt
done
for lay in sets quad; do
  echo "== FWAV_UMMA_COLLECT=$lay FWAV_UMMA_MODE=precise (config 2, full split)" >> $O/r02a_timing.txt
  FWAV_UMMA_COLLECT=$lay FWAV_UMMA_MODE=precise timeout 200 python scripts/time_topk.py 1.0 umma 2 2>/dev/null | cut -c1-400 >> $O/r02a_timing.txt
done
for compact in 0 1; do
  echo "== FWAV_UMMA_COMPACT=$compact (config-4 shape at 1/10 length)" >> $O/r02a_timing.txt
  FWAV_UMMA_COMPACT=$compact FWAV_UMMA_VERBOSE=1 timeout 300 python bench.py --workload c4 --scale 0.1 --steps 1 --warmup 1 --no-decode --no-cpu > $O/r02a_c4_$compact.json 2> $O/r02a_c4_$compact.err
  grep "live embedding" $O/r02a_c4_$compact.err | tail -1 | cut -c1-200 >> $O/r02a_timing.txt
  python - $compact >> $O/r02a_timing.txt <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/r02a_c4_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['roofline'].get('search_phases_ms'))
except Exception as e:
    print("no result:", e)
PY
done
cat $O/r02a_timing.txt
# ---- bench lines (default layout, then fat) ----
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r02a_bench_default.json 2> $O/r02a_bench_default.err
echo "bench default rc=$?"; cut -c1-300 $O/r02a_bench_default.json
FWAV_UMMA_COLLECT=fat timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r02a_bench_fat.json 2> $O/r02a_bench_fat.err
echo "bench fat rc=$?"; cut -c1-300 $O/r02a_bench_fat.json
# ---- microbenchmark (instruction peaks incl. tf32) ----
out=$O/r02a_umma_microbench.jsonl; rm -f $out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader >> $out
for v in $(scripts/umma_microbench list); do
  timeout 60 scripts/umma_microbench $v >> $out 2>&1 || echo "{\"variant\": \"$v\", \"exit\": $?}" >> $out
done
grep -E "tf32|ss_m128n256_sw32\"|ss_m128n128_sw32\"|cg2_ss_m256n256_sw32" $out
# ---- ncu: the fat collect kernel, one launch, full set ----
FWAV_UMMA_COLLECT=fat timeout 600 ncu --set full --clock-control none --import-source on -k regex:collect_fat_kernel --launch-skip 2 --launch-count 1 \
  -o $O/r02a_prof_collect_fat -f python scripts/time_topk.py 1.0 umma 1 > $O/r02a_ncu_fat.log 2>&1
echo "ncu fat rc=$?"; tail -3 $O/r02a_ncu_fat.log
ls -la $O | grep r02a
