"""Time fwav_topk alone (CUDA events on the launching stream) on a config-2-shaped
table.  usage: time_topk.py [scale] [impl] [reps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "audio-compression_b200"))
import numpy as np, torch
from fwav_b200 import _lib, synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
impl = sys.argv[2] if len(sys.argv) > 2 else "umma"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
sig = synth.music_like(seconds=180.0 * scale, rate=44100, seed=2)
ctx = _lib.Context(0)
ctx.set_search_impl({"ffma": 1, "umma": 2}[impl])
N, ds, K, ED, tile = 16, 4, 32, 16, 4096
n_d = _lib.count_domains(len(sig), tile, ds); n_q = (len(sig) + N - 1) // N
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
d_sig = torch.from_numpy(sig).to(dev)
d_dom = torch.empty((n_d, N), device=dev); d_emb = torch.empty((n_d, ED), device=dev)
d_cand = torch.empty((n_q, K), dtype=torch.int32, device=dev)
ctx.build_tables(d_sig.data_ptr(), len(sig), tile, N, ds, ED, d_dom.data_ptr(), d_emb.data_ptr(), st.cuda_stream)
ms = []
for r in range(reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.topk(d_emb.data_ptr(), n_q, d_emb.data_ptr(), n_d, ED, K, None, d_cand.data_ptr(), None, st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    if r: ms.append(e0.elapsed_time(e1))
pairs = float(n_q) * n_d
print(json.dumps(dict(fallbacks=ctx.search_fallbacks(), phases={k: round(v, 3) for k, v in ctx.search_timings().items()}, rank=os.environ.get('FWAV_UMMA_RANK', 'default'), front=os.environ.get('FWAV_UMMA_FRONT', 'default'), reps=reps + 1, impl=impl, dbg=os.environ.get("FWAV_UMMA_DEBUG", "0"), scale=scale, n_q=n_q, n_d=n_d,
                      ms=float(np.mean(ms)), gpairs_per_s=pairs / np.mean(ms) / 1e6,
                      cycles_per_tilestep_per_sm=np.mean(ms) * 1e-3 * 1.965e9 / (((n_q + 255) // 256) * ((n_d + 127) // 128) / 148))))
