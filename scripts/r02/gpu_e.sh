#!/bin/bash
# Round 2, GPU call E: full GPU suite, smoke, the bench line, ncu launch list of the bench + full capture of the collect kernel
set +e
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $O/r02e_pytest.txt 2>&1
echo "full suite: rc=$? $(tail -1 $O/r02e_pytest.txt)"; grep -E "^(FAILED|ERROR)" $O/r02e_pytest.txt | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02e_smoke.txt 2>&1; echo "smoke rc=$? $(tail -1 $O/r02e_smoke.txt)"
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02e_bench.json 2> $O/r02e_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value']); print('e2e', d['e2e']); print('roof', {k:v for k,v in d['roofline'].items() if k!='peak_source'})
print('kern', d['kernels']); print('decode', d['decode']); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02e_bench_reference.json 2> $O/r02e_bench_reference.err; echo "reference arm rc=$?"; cut -c1-400 $O/r02e_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02e_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-decode > $O/r02e_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel --launch-skip 5 --launch-count 1 \
  -o $O/r02e_prof_collect -f python scripts/time_topk.py 1.0 umma 1 > $O/r02e_ncu_collect.log 2>&1
echo "ncu collect rc=$?"; tail -2 $O/r02e_ncu_collect.log
timeout 600 ncu --set full --clock-control none -k regex:"decode_stream|affine_kernel|embed_static|domains_from|half_sums|apply_gate|frame_energy" --launch-count 12 \
  -o $O/r02e_prof_small -f python bench.py --steps 1 --warmup 3 --no-cpu --decode-scale 1.0 > $O/r02e_ncu_small.log 2>&1
echo "ncu small rc=$?"
