"""Fused decode + broadcast (fwav_decode_iter_bcast) against the NCCL all-gather form, on real GPUs:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_fused_decode.py

Every rank decodes the same synthetic matches three ways -- stores through the NVSwitch multicast address, stores
through peer pointers (FWAV_DECODE_MULTIMEM=0), NCCL all-gather every iteration -- and the full reconstructions,
iteration counts and deltas must be bit-identical on every rank, with and without early convergence.  Not collected
by pytest (needs >= 2 GPUs and torchrun); run by scripts/r02/gpu_h.sh."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-compression_b200"))

import numpy as np
import torch
import torch.distributed as dist

from fwav_b200 import distributed as D


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = D.CudaEngine(local)
    ok = True
    for N, n_r, n_d in ((16, 200_003, 90_000), (4, 77_777, 50_000), (8, 130_000, 64_000)):
        g = torch.Generator(device=dev)
        g.manual_seed(11)
        domains = torch.randn((n_d, N), generator=g, device=dev) * 300
        idx = torch.randint(0, n_d, (n_r,), generator=g, device=dev, dtype=torch.int32)
        idx[::97] = -1                                              # sentinels
        s = (torch.rand(n_r, generator=g, device=dev) * 2 - 1).float()
        o = (torch.rand(n_r, generator=g, device=dev) * 2000 - 1000).float()
        sym = torch.randint(0, 2, (n_r,), generator=g, device=dev, dtype=torch.uint8)
        for kw in (dict(iterations=9, convergence_eps=0.0, s_damping=0.5), dict(iterations=12, convergence_eps=3e-2, s_damping=0.5),
                   dict(iterations=8, convergence_eps=1e-3)):
            ref, it_ref, d_ref = D.decode_sharded(eng, domains, idx, s, o, sym, N, fused=False, **kw)
            for mm in ("1", "0"):
                os.environ["FWAV_DECODE_MULTIMEM"] = mm
                out, it, d = D.decode_sharded(eng, domains, idx, s, o, sym, N, fused=True, **kw)
                same = bool(torch.equal(out.view(torch.int32), ref.view(torch.int32))) and it == it_ref and d == d_ref
                ok = ok and same
                if rank == 0:
                    print(f"N={N} n_r={n_r} {kw} multimem={mm}: iters {it} (ref {it_ref}) delta {d:.6g} "
                          f"{'bit-identical' if same else 'DIFFERENT'}", flush=True)
    t = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if not int(t[0]):
        sys.exit(1)
    if rank == 0:
        print("fused decode: all variants bit-identical to the all-gather form on every rank")


if __name__ == "__main__":
    main()
