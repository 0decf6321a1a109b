#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
FWAV_UMMA_VERBOSE=1 timeout 600 python bench.py --workload c4 --scale 0.1 --steps 1 --warmup 1 --no-decode --no-cpu > $O/r02i_c4.json 2> $O/r02i_c4.err
grep "fwav\]" $O/r02i_c4.err | tail -12 | cut -c1-260
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_c4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline'].get('search_phases_ms'))
PY
true
grep "fwav\]" $O/r02i_c4_nc.err | tail -6 | cut -c1-260
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_c4_nc.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline'].get('search_phases_ms'))
PY
