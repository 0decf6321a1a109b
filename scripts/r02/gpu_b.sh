#!/bin/bash
# Round 2, GPU call B: collect4 layouts (thin/fat x 2/4 issuer threads) parity + timing, device pre-step tests,
# bench with the Python-API e2e.
set +e
O=gpurun_out; mkdir -p $O
for lay in thin4 thin2 fat4 fat2; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "collect_layouts and $lay" > $O/r02b_layout_$lay.txt 2>&1
  echo "layout $lay: rc=$? $(tail -1 $O/r02b_layout_$lay.txt)"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prestep or raw_signal or multi_batch or forged" > $O/r02b_new.txt 2>&1
echo "new tests: rc=$? $(tail -1 $O/r02b_new.txt)"; grep -E "Error|assert|FAILED" $O/r02b_new.txt | head -20
rm -f $O/r02b_timing.txt
for lay in sets thin4 thin2 fat4 fat2; do
  echo "== FWAV_UMMA_COLLECT=$lay (config 2)" >> $O/r02b_timing.txt
  FWAV_UMMA_COLLECT=$lay timeout 200 python scripts/time_topk.py 1.0 umma 3 2>/dev/null | cut -c1-400 >> $O/r02b_timing.txt
done
cat $O/r02b_timing.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $O/r02b_pytest.txt 2>&1
echo "full suite: rc=$? $(tail -1 $O/r02b_pytest.txt)"; grep -E "FAILED|Error" $O/r02b_pytest.txt | head
python - <<'PY' > $O/r02b_api.txt 2>&1
import os, sys, time
sys.path.insert(0, "audio-compression_b200")
import numpy as np, fractal
from fwav_b200 import synth
sig, rate, tile, K = synth.make("c2", 1.0)
for mode in ("1", "0"):
    os.environ["FWAV_PINNED"] = mode
    ts = []
    for i in range(6):
        t0 = time.perf_counter()
        out = fractal.compress_audio_arrays(sig, tile_size=tile)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("FWAV_PINNED=%s compress_audio_arrays ms:" % mode, [round(t, 1) for t in ts])
t0 = time.perf_counter(); out = fractal.compress_audio(sig, rate, 2, tile_size=tile); print("compress_audio (tuple list) ms", (time.perf_counter() - t0) * 1e3)
PY
cat $O/r02b_api.txt
ls -la $O | grep r02b
