"""fractal — drop-in for xavenordu/Audio-Compression's `fractal.py`, with the
range/domain matching and the iterative reconstruction running on a B200.

Same public names, signatures, return tuples, .fwav bytes and CLI as the
reference; the bodies of the hot path are calls into
libfwav_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/fwav_b200.h).  There is no CuPy, Triton, hnswlib or CPU fallback: if the
library or a CUDA device is missing, compress_audio / decompress_audio raise.

Put this directory on sys.path (or copy it next to your script) and
`import fractal` keeps working:

    from fractal import compress_audio, save_compressed, load_compressed, decompress_audio

Behaviours of the reference that are kept on purpose (fractal.py = reference):
  * range_size = max(4, tile_size // 256), domain_step = max(1, range_size // 4)   (:1070-1071)
  * the query of range i is row i of the DOMAIN embedding table             (:1190-1195)
    -> ValueError when there are more ranges than domains; set
       FWAV_QUERY_MODE=range for true range embeddings (not reference-faithful)
  * the `top_k` ARGUMENT is ignored; the module global `top_k` is used       (:77, :611-622)
  * energy-pruned ranges are stored as (0, s, o, 0, inf)                     (:602-603, :816-822)
  * `use_gpu` is accepted and ignored: the path always runs on the GPU.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import time
from multiprocessing import Pool

import numpy as np

from fwav_b200 import _lib
from fwav_b200.container import FWAV_VERSION, MatchArrays, load_compressed, save_compressed  # noqa: F401
from fwav_b200.prestep import frame_ranges, voiced_detection  # noqa: F401
from fwav_b200.wavio import read_wav_mono, write_wav  # noqa: F401

logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s: %(message)s")
logger = logging.getLogger("fwavc")

# module globals callers of the reference read or set
top_k = 32            # candidates per range; config 4 sets fractal.top_k = 64
EMBED_K = 32
GPU_AVAILABLE = True  # kept for callers that branch on them; the path is GPU-only
GPU_WORKING = True
HNSW_AVAILABLE = False
FAISS_AVAILABLE = False


def _device():
    return int(os.environ.get("FWAV_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def _empty_result(range_size, tile_size, domain_step, energy_thresh, original_len):
    return ([], np.zeros((0, range_size), dtype=np.float32), 0, range_size, tile_size,
            domain_step, energy_thresh, original_len)


_in_tonal = [False]     # compress_audio_arrays re-enters itself once with the context switched to the tonal embedding


def compress_audio_arrays(signal, tile_size=1024, emb_dim=16, energy_thresh=1e-4, fast_mode=True,
                          k=None, query_mode=None, ctx=None, embedding=None):
    """compress_audio without the Python tuple list: returns
    (MatchArrays, domains, n_ranges, range_size, tile_size, domain_step,
    energy_thresh, original_len)."""
    signal = np.ascontiguousarray(signal, dtype=np.float32)
    range_size, domain_step = _lib.geometry(tile_size)
    n_domains = _lib.count_domains(len(signal), tile_size, domain_step)
    original_len = len(signal)
    empty = (MatchArrays.from_any([]),) + _empty_result(range_size, tile_size, domain_step, energy_thresh, original_len)[1:]
    if query_mode is None:
        query_mode = 1 if os.environ.get("FWAV_QUERY_MODE", "reference") == "range" else 0
    kk = top_k if k is None else k
    if embedding is None:
        embedding = os.environ.get("FWAV_EMBEDDING", "two_head")
    if embedding not in ("two_head", "tonal"):
        raise ValueError(f"embedding must be 'two_head' (the reference's live path) or 'tonal' (tile_embedding), got {embedding!r}")
    if embedding == "tonal" and not _in_tonal[0]:
        # the "fixed" mode of the docs: tile_embedding(k = emb_dim) (fractal.py:178-208; EMBED_K = 32 there)
        ctx = ctx or _lib.default_context(_device())
        ctx.set_embedding(_lib.EMBED_TONAL)
        _in_tonal[0] = True
        try:
            return compress_audio_arrays(signal, tile_size, emb_dim, energy_thresh, fast_mode, k, query_mode, ctx, "tonal")
        finally:
            _in_tonal[0] = False
            ctx.set_embedding(_lib.EMBED_TWO_HEAD)
    if not (signal[:4096].any() or signal.any()):
        return empty       # digital silence: every frame energy is 0, the gate never opens (fractal.py:1083-1093)
    if n_domains > 0 and original_len >= 10 * range_size:
        # the usual case: the raw signal goes to the device once; voiced gate, masking, reflect padding and framing
        # (fractal.py:880-909, :1074-1112) run there, bit-identical to the host pre-step
        ctx = ctx or _lib.default_context(_device())
        res = ctx.compress_signal_host(signal, tile_size, emb_dim, kk, energy_thresh, fast_mode, query_mode)
        if res is None:
            return empty
        n_ranges = len(res["idx"])
    else:
        # inputs of a few frames (np.convolve's 'same' mode changes shape below five frames) or without a single
        # domain: the host pre-step decides what the reference would return
        ranges, original_len = frame_ranges(signal, range_size, energy_thresh)
        if ranges is None or n_domains == 0:
            return empty
        ctx = ctx or _lib.default_context(_device())
        res = ctx.compress_host(signal, ranges, tile_size, emb_dim, kk, energy_thresh, fast_mode, query_mode)
        n_ranges = len(ranges)
    m = MatchArrays(res["idx"], res["s"], res["o"], res["sym"], res["err"])
    return (m, res["domains"], n_ranges, range_size, tile_size, domain_step, energy_thresh, original_len)


def compress_audio(signal, framerate, sampwidth, tile_size=1024, emb_dim=16, top_k=top_k, ef_search=50,
                   use_gpu=False, energy_thresh=1e-4, domains_tmpdir=None, batch_size_gpu=512,
                   batch_size_cpu=128, fast_mode=True, transient_weight=1.0, n_mels=40, cpu_workers=None):
    """Fractal compression of a mono signal.  Returns the reference's 8-tuple
    (matches, domains, n_ranges, range_size, tile_size, domain_step,
    energy_thresh, original_len); matches is a list of
    (domain index, s, o, mirrored flag, residual L2)."""
    out = compress_audio_arrays(signal, tile_size=tile_size, emb_dim=emb_dim,
                                energy_thresh=energy_thresh, fast_mode=fast_mode,
                                k=globals()["top_k"])
    return (out[0].tolist(),) + out[1:]


def decompress_audio(matches, domains_array, n_ranges, range_size, iterations=8, convergence_eps=1e-3,
                     use_gpu=False, original_len=None, s_clip=16.0, s_damping=0.0):
    """Iterative reconstruction; returns a float32 array trimmed to original_len."""
    m = MatchArrays.from_any(matches)
    if n_ranges == 0 or len(m) == 0:
        recon = np.zeros(n_ranges * range_size, dtype=np.float32)
    else:
        ctx = _lib.default_context(_device())
        recon, iters, delta = ctx.decode_host(domains_array, m.idx, m.s, m.o, m.sym, range_size,
                                              iterations, convergence_eps, s_clip, s_damping)
        if iters < iterations or delta < convergence_eps:
            logger.info(f"Converged after {iters} iterations (delta={delta:.3e})")
    if original_len is not None:
        recon = recon[:original_len]
    return recon


def compute_snr(original, reconstructed):
    n = min(len(original), len(reconstructed))
    ref = np.asarray(original[:n], dtype=np.float64)
    err = ref - np.asarray(reconstructed[:n], dtype=np.float64)
    noise = float(np.sum(err * err))
    if noise <= 0:
        return float("inf")
    return 10.0 * np.log10(float(np.sum(ref * ref)) / noise)


# ----------------------------- file drivers (unchanged behaviour) -----------------------------
# COMPATIBILITY TRANSLITERATION: everything from here to the end of the file restates the reference's file drivers and
# CLI (fractal.py:1491-1669: same argparse flags and help strings, log messages, result dicts, output-path quirks)
# because SURVEY 8(b) requires the CLI and the drivers to stay unchanged.  It is host glue, not part of the hot path.

def process_file_compress(path, outdir=None, tile=1024, energy_thresh=1e-4, use_gpu=False):
    try:
        t0 = time.time()
        signal, framerate, sampwidth = read_wav_mono(path)
        if sampwidth == 4:
            signal = np.clip(signal.astype(np.float32), -1.0, 1.0)
        (matches, domains, n_ranges, range_size, tile_size, domain_step, energy_threshold,
         original_len) = compress_audio_arrays(signal, tile_size=tile, energy_thresh=energy_thresh,
                                               k=globals()["top_k"])
        logger.info(f"Processed {len(matches)} ranges, domain matrix shape {domains.shape}")
        if outdir and not os.path.exists(outdir):
            os.makedirs(outdir)
        # the reference treats its OUTPUT argument as a directory (:1509)
        outpath = (os.path.splitext(path)[0] + ".fwav") if outdir is None \
            else os.path.join(outdir, os.path.basename(path) + ".fwav")
        save_compressed(outpath, matches, domains, range_size, framerate, sampwidth, tile_size,
                        domain_step, energy_threshold, original_len)
        elapsed = time.time() - t0
        out_size = os.path.getsize(outpath)
        ratio = os.path.getsize(path) / out_size if out_size > 0 else 0
        logger.info(f"Compressed {path} -> {outpath}  time={elapsed:.2f}s  ratio={ratio:.2f}")
        return {"input": path, "output": outpath, "time_s": elapsed, "ratio": ratio}
    except Exception as e:
        logger.exception("Compression failed for %s", path)
        return {"input": path, "error": str(e)}


def process_file_decompress(path, outdir=None, iterations=8, eps=1e-3, use_gpu=False):
    try:
        t0 = time.time()
        (matches, domains, n_ranges, range_size, framerate, sampwidth, tile_size, domain_step,
         energy_threshold, original_len) = load_compressed(path, as_arrays=True)
        recon = decompress_audio(matches, domains, n_ranges, range_size, iterations=iterations,
                                 convergence_eps=eps, original_len=original_len)
        if outdir and not os.path.exists(outdir):
            os.makedirs(outdir)
        if sampwidth == 4:
            recon = np.clip(recon, -1.0, 1.0)
        outpath = (os.path.splitext(path)[0] + "_recon.wav") if outdir is None \
            else os.path.join(outdir, os.path.basename(path) + "_recon.wav")
        write_wav(outpath, np.asarray(recon), framerate, sampwidth)
        elapsed = time.time() - t0
        logger.info(f"Decompressed {path} -> {outpath}  time={elapsed:.2f}s")
        return {"input": path, "output": outpath, "time_s": elapsed}
    except Exception as e:
        logger.exception("Decompression failed for %s", path)
        return {"input": path, "error": str(e)}


def _run_batch(files, worker, args_of, workers, metrics_path):
    pool = Pool(processes=min(workers, len(files)))
    try:
        jobs = [pool.apply_async(worker, args_of(f)) for f in files]
        results = [j.get() for j in jobs]
    finally:
        pool.close()
        pool.join()
    os.makedirs(os.path.dirname(metrics_path), exist_ok=True)
    with open(metrics_path, "w") as mf:
        json.dump(results, mf, indent=2)
    logger.info(f"Wrote metrics to {metrics_path}")


def main():
    parser = argparse.ArgumentParser(description="Fractal WAV compressor with GPU, batch processing, and metrics")
    sub = parser.add_subparsers(dest="cmd")
    pc = sub.add_parser("compress")
    pc.add_argument("input", help="input WAV file or directory")
    pc.add_argument("output", nargs="?", default=None, help="output FWAV file (required unless --batch)")
    pc.add_argument("--tile", type=int, default=1024)
    pc.add_argument("--out", default=None, help="output directory (batch mode)")
    pc.add_argument("--energy-thresh", type=float, default=1e-4)
    pc.add_argument("--gpu", action="store_true")
    pc.add_argument("--batch", action="store_true", help="treat input as directory and compress all WAV inside")
    pc.add_argument("--workers", type=int, default=4, help="parallel file-level workers for batch")
    pd = sub.add_parser("decompress")
    pd.add_argument("input", help="input file or directory")
    pd.add_argument("--out", default=None, help="output file or directory")
    pd.add_argument("--iter", type=int, default=8)
    pd.add_argument("--eps", type=float, default=1e-3)
    pd.add_argument("--gpu", action="store_true")
    pd.add_argument("--batch", action="store_true", help="treat input as directory and decompress all FWAV inside")
    pd.add_argument("--workers", type=int, default=4, help="parallel file-level workers for batch")
    args = parser.parse_args()

    if args.cmd == "compress":
        if not args.batch:
            if args.output is None:
                parser.error("compress requires OUTPUT unless --batch is used")
            process_file_compress(args.input, args.output, args.tile, args.energy_thresh, args.gpu)
            return
        if args.output is not None:
            parser.error("Do not provide positional OUTPUT when using --batch; use --out instead")
        out_dir = args.out or args.input
        files = [os.path.join(args.input, f) for f in os.listdir(args.input) if f.lower().endswith(".wav")]
        todo = [f for f in files if not os.path.exists(os.path.join(out_dir, os.path.basename(f) + ".fwav"))]
        logger.info(f"Batch compressing {len(todo)}/{len(files)} files using {args.workers} workers")
        if not todo:
            logger.info("No files to compress — all already exist.")
            return
        _run_batch(todo, process_file_compress,
                   lambda f: (f, os.path.join(out_dir, os.path.basename(f) + ".fwav"), args.tile,
                              args.energy_thresh, args.gpu),
                   args.workers, os.path.join(out_dir, "compression_metrics.json"))
    elif args.cmd == "decompress":
        if not args.batch:
            out_file = args.out or (os.path.splitext(args.input)[0] + "_recon.wav")
            process_file_decompress(args.input, out_file, args.iter, args.eps, args.gpu)
            return
        out_dir = args.out or args.input
        files = [os.path.join(args.input, f) for f in os.listdir(args.input) if f.lower().endswith(".fwav")]
        todo = [f for f in files
                if not os.path.exists(os.path.join(out_dir, os.path.basename(f).replace(".fwav", "_recon.wav")))]
        logger.info(f"Batch decompressing {len(todo)}/{len(files)} files using {args.workers} workers")
        if not todo:
            logger.info("No files to decompress — all already exist.")
            return
        _run_batch(todo, process_file_decompress,
                   lambda f: (f, os.path.join(out_dir, os.path.basename(f).replace(".fwav", "_recon.wav")),
                              args.iter, args.eps, args.gpu),
                   args.workers, os.path.join(out_dir, "decompression_metrics.json"))
    else:
        parser.print_help()


if __name__ == "__main__":
    main()
