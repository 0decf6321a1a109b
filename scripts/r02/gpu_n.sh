#!/bin/bash
# Round 2, GPU call N (third run: 48 ms collect kernel, fp16 threshold pass, cost-model route): the state to be judged -- full GPU suite, smoke, bench N=1 (+ reference arm), ncu launch list
# and full capture of the new dominant kernel
set +e
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --durations=6 > $O/r02n3_pytest.txt 2>&1
echo "full suite: rc=$? $(tail -1 $O/r02n3_pytest.txt)"; grep -E "^(FAILED|ERROR)" $O/r02n3_pytest.txt | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02n3_smoke.txt 2>&1; echo "smoke rc=$? $(tail -1 $O/r02n3_smoke.txt)"
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02n3_bench.json 2> $O/r02n3_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02n3_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value']); print('e2e', d['e2e']); print('roof', {k:v for k,v in d['roofline'].items() if k not in ('peak_source','traffic')})
print('decode', d['decode']['damped_0.5']); print('cpu', d['cpu_baseline']['value']); print('clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02n3_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-decode > $O/r02n3_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:collect_hi_kernel --launch-skip 2 --launch-count 1 \
  -o $O/r02n3_prof_collect_hi -f python scripts/time_topk.py 1.0 umma 1 > $O/r02n3_ncu_collect.log 2>&1
echo "ncu collect_hi rc=$?"; tail -2 $O/r02n3_ncu_collect.log
