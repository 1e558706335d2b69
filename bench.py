#!/usr/bin/env python
"""bench.py -- decoded frames/s of the token-passing search on synthetic HLG.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic input:
InitDecoding + AdvanceDecoding over all frames + GetBestPath for every
utterance lane of the workload (BASELINE.json configs[2]: HLG with a synthetic
3-gram LM, ~5M arcs, 1024 utterances x T=1000 x V=500, beam 20, max_active
7000; one graph replica and 1024 lanes per GPU -> weak scaling).

  value  : whole-job frames/s with the log-probs already resident in HBM.
  e2e    : the same through the C ABI with HOST (pinned) log-prob buffers:
           host->device copies and the device->host read of the best paths are
           inside the timed region.
  roofline: algorithmic bytes (counted by the kernel) / the
           search kernel's CUDA-event duration, against MEASURED_PEAKS.json.
  cpu_baseline: the reference's own faster-decoder.cc (oracle/_ref), one
           utterance per host thread, on a bounded sample of the same workload.

`--impl reference` times that CPU reference alone (rank 0 only).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "kaldi-decoder_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "decoded_frames_per_sec"
UNIT = "frames/s"
OPTS = dict(beam=20.0, max_active=7000, min_active=20, beam_delta=0.5, hash_ratio=2.0)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4"])
    ap.add_argument("--lanes", type=int, default=0, help="utterances per GPU (default: config)")
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--peak", type=float, default=12.0)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--threads-per-lane", type=int, default=0)
    ap.add_argument("--hash-capacity", type=int, default=0,
                    help="per-lane recombination table entries (default: per config)")
    ap.add_argument("--arena-records", type=int, default=0)
    ap.add_argument("--chunk-frames", type=int, default=0)
    ap.add_argument("--cpu-sample-utts", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--check", action="store_true", help="compare a sample with the reference")
    return ap.parse_args()


CONFIG_LANES = {"C1": 64, "C2": 256, "C3": 1024, "C4": 1024}
# tokens alive in one frame must stay below half of this (measured maxima: C1 500, C2 146k,
# C3 49k, C4 142k tokens)
CONFIG_HASH = {"C1": 1 << 14, "C2": 1 << 20, "C3": 1 << 18, "C4": 1 << 20}
CONFIG_NAME = {
    "C1": "H-500 CTC topology, 64 utts x T=1000 x V=500, beam 20, max_active 7000",
    "C2": "HL 200k-word lexicon trie, 256 utts x T=1000 x V=500, beam 20, max_active 7000",
    "C3": "HLG synthetic 3-gram (~5M arcs), 1024 utts x T=1000 x V=500, beam 20, max_active 7000",
    "C4": "HLG synthetic 4-gram (~150M arcs), 1024 utts x T=1000 x V=500, beam 20, max_active 7000",
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_device_logprobs(g, n_utts, T, seed, peak, device):
    """float32 [n_utts, T, V] on `device`: N(0,1) + peak*onehot(alignment), log-softmax.
    Alignments come from the graph's own LM (synth.make_alignment); the noise is
    torch's seeded generator on the device."""
    import torch
    from kaldi_decoder_b200 import synth
    V = int(g.lm["vocab"])
    ali = np.empty((n_utts, T), dtype=np.int64)
    for u in range(n_utts):
        rng = np.random.default_rng(np.random.PCG64(seed * 1_000_003 + u))
        ali[u] = synth.make_alignment(g, T, rng)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty((n_utts, T, V), dtype=torch.float32, device=device)
    ali_d = torch.from_numpy(ali).to(device)
    chunk = 64
    for a in range(0, n_utts, chunk):
        b = min(n_utts, a + chunk)
        x = torch.randn((b - a, T, V), generator=gen, device=device, dtype=torch.float32)
        x.scatter_add_(2, ali_d[a:b].unsqueeze(-1),
                       torch.full((b - a, T, 1), float(peak), device=device))
        out[a:b] = torch.log_softmax(x, dim=-1)
    return out


# dram__bytes_read.sum + dram__bytes_write.sum of kd_advance_kernel from the committed
# `ncu --set full` capture (profiles/r1_final_ncu_summary.txt: 9.58 + 6.72 GB for a launch of
# 1024 lanes x 100 frames of this workload), per lane-frame.  Only valid for config C3.
NCU_DRAM_BYTES_PER_LANE_FRAME_C3 = (9.580444e9 + 6.723989e9) / (1024 * 100)


def algorithmic_bytes(st: dict, cols: int) -> float:
    # SURVEY.md §8(d)
    return (16.0 * (st["emit_arcs"] + st["eps_arcs"]) + 16.0 * st["tokens_in"]
            + 24.0 * st["tokens_out"] + 4.0 * cols * st["frames"])


def run_reference_cpu(g, mats: np.ndarray, threads: int):
    from oracle import kd_ref
    rg = kd_ref.RefGraph(g)
    secs, _, _ = kd_ref.decode_batch(rg, mats, kd_ref.Options(**OPTS), threads, want_paths=False)
    return secs


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lanes = args.lanes or CONFIG_LANES[args.config]
    T = args.frames
    cores = len(os.sched_getaffinity(0))

    from kaldi_decoder_b200 import synth

    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import kd_ref
        if not kd_ref.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libkd_ref.so not built"}))
            return 0
        g = synth.make_config_graph(args.config)
        n_s = args.cpu_sample_utts or min(lanes, max(2 * cores, 16))
        mats = synth.make_batch(g, n_s, T, seed=args.seed, peak=args.peak)
        for _ in range(min(args.warmup, 1)):
            run_reference_cpu(g, mats[: max(1, min(n_s, cores))], cores)
        times = [run_reference_cpu(g, mats, cores) for _ in range(args.steps)]
        sec = sum(times) / len(times)
        value = n_s * T / sec
        sample = f"{n_s} utterances x {T} frames of the workload per step, {cores} threads"
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": CONFIG_NAME[args.config], "peak": args.peak, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return 0

    import torch
    import torch.distributed as dist
    from kaldi_decoder_b200 import capi

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_setup = time.time()
    g = synth.make_config_graph(args.config)
    V = int(g.lm["vocab"])
    dg = capi.DeviceGraph.from_graph(g, device=local_rank)
    dec = capi.LaneDecoder(dg, capi.make_options(**OPTS), max_lanes=lanes,
                           hash_capacity=args.hash_capacity or CONFIG_HASH[args.config],
                           arena_records=args.arena_records,
                           threads_per_lane=args.threads_per_lane,
                           chunk_frames=args.chunk_frames)
    # every rank decodes its own utterances (seed differs per rank): weak scaling
    logp = make_device_logprobs(g, lanes, T, args.seed + 7919 * rank, args.peak, dev)
    torch.cuda.synchronize()
    lane_ids = list(range(lanes))
    rows = [T] * lanes
    dptrs = [logp[u].data_ptr() for u in range(lanes)]
    host = None
    if not args.no_e2e:
        host = torch.empty((lanes, T, V), dtype=torch.float32, pin_memory=True)
        host.copy_(logp)
        torch.cuda.synchronize()
        hptrs = [host[u].data_ptr() for u in range(lanes)]
    setup_s = time.time() - t_setup

    kernel_ms = []
    result = {}

    def step_device():
        dec.init(lane_ids)
        dec.advance_ptrs(lane_ids, dptrs, rows, V, None, -1, capi.KD_MEM_DEVICE)
        kernel_ms.append(dec.last_advance_info()[0])
        # zero-copy result (views of the decoder's pinned buffer, consumed before the next call)
        result["paths"] = dec.best_paths(lane_ids, True, copy=False)

    def step_host():
        dec.init(lane_ids)
        dec.advance_ptrs(lane_ids, hptrs, rows, V, None, -1, capi.KD_MEM_HOST)
        result["paths"] = dec.best_paths(lane_ids, True, copy=False)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    sampler = ClockSampler(local_rank)
    kernel_ms.clear()
    for _ in range(args.warmup):
        step_device()
    kernel_ms.clear()
    sampler.start()
    dt = timed(step_device, args.steps, 0)
    clocks = sampler.stop()
    st = dec.stats()  # counters of the last step (kd_decoder_init resets them)
    ms_per_step = dt / args.steps * 1e3
    frames_per_step = lanes * T * world
    value = frames_per_step / (dt / args.steps)
    k_ms = sum(kernel_ms) / max(1, len(kernel_ms))
    peak_gbs, peak_src = measured_peaks()
    alg_bytes = algorithmic_bytes(st, V)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    paths = result["paths"]
    d2h = int(paths.ilabels.nbytes * 4) + 8 * lanes
    n_final = int(np.count_nonzero(paths.reached_final))
    # label sequences of the timed device-input step, kept for the parity sample below (the
    # views are overwritten by the next best_paths call)
    n_keep = min(lanes, args.cpu_sample_utts or max(2 * cores, 16))
    kept = [(paths[u].isyms.copy(), paths[u].osyms.copy()) for u in range(n_keep)]

    e2e = None
    if not args.no_e2e:
        dt_h = timed(step_host, args.steps, max(1, args.warmup))
        e2e = {"value": frames_per_step / (dt_h / args.steps), "unit": UNIT,
               "h2d_bytes_per_step": int(lanes * T * V * 4), "d2h_bytes_per_step": d2h,
               "ms_per_step": dt_h / args.steps * 1e3}

    cpu = None
    check = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import kd_ref
        if kd_ref.available():
            n_s = n_keep
            sample_mats = logp[:n_s].cpu().numpy()
            rg = kd_ref.RefGraph(g)
            secs, rpaths, rrf = kd_ref.decode_batch(rg, sample_mats, kd_ref.Options(**OPTS), cores,
                                                    want_paths=True)
            cpu = {"value": n_s * T / secs, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"first {n_s} utterances x {T} frames of the same batch, "
                             f"{cores} threads, oracle/_ref (unmodified faster-decoder.cc)"}
            # parity of the sample (not timed): label sequences vs the reference
            same = 0
            for u in range(n_s):
                if (np.array_equal(kept[u][0], rpaths[u].isyms)
                        and np.array_equal(kept[u][1], rpaths[u].osyms)):
                    same += 1
            check = {"utterances": n_s, "identical_label_sequences": same}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": CONFIG_NAME[args.config], "graph": g.stats(),
                       "lanes_per_gpu": lanes, "frames": T, "vocab": V, "peak": args.peak,
                       "options": OPTS, "parallelism": f"replicas x{world} (utterances sharded)",
                       "l2_policy": "inputs larger than L2 (log-probs %.2f GB per GPU)" % (lanes * T * V * 4 / 1e9),
                       "threads_per_lane": dec.info()["threads_per_lane"],
                       "reached_final": n_final},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": 4 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs,
                         "traffic": (NCU_DRAM_BYTES_PER_LANE_FRAME_C3 * lanes * T
                                     if args.config == "C3" and args.peak == 12.0 else None),
                         "traffic_note": "DRAM bytes per launch scaled from the ncu capture in "
                                         "profiles/r1_final_ncu_summary.txt (per lane-frame x lanes x frames)",
                         "kernel": "kd_advance_kernel", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "counters": st},
            "cpu_baseline": cpu,
            "parity_sample": check,
            "setup_s": setup_s,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
