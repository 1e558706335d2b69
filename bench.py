#!/usr/bin/env python
"""bench.py -- decoded frames/s of the token-passing search on synthetic HLG.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C1..C4]

One "step" = one pass of the hot path over one batch of synthetic input: InitDecoding +
AdvanceDecoding over all frames + GetBestPath for every utterance lane of the workload
(default: BASELINE.json configs[2], HLG with a synthetic 3-gram LM, ~5M arcs, 1024 utterances
x T=1000 x V=500, beam 20, max_active 7000; one graph replica and 1024 utterances per step
per GPU -> weak scaling).  A step is ONE kernel launch (kd_decoder_advance_async with
KD_ADVANCE_INIT | KD_ADVANCE_FINALIZE) plus the copies around it.  Steps are enqueued one
ahead on alternating lane groups: the lanes of step k+1 take over the SMs as the slowest
lanes of step k finish, and the upload of step k+1 runs under the search of step k.  Every
step's work lies inside the timed region; `--groups 1` gives strictly sequential steps.

  value  : whole-job frames/s with the log-probs already resident in HBM (K steps, wall time
           between two barriers, max over ranks).
  e2e    : the same through the C ABI with HOST (pinned) log-prob buffers: the host->device
           copies and the device->host read of the best paths are inside the timed region.
  roofline: .achieved = algorithmic bytes (counted by the kernel: the
           out-degree of every expanded token, whether or not a label table let the kernel
           skip the arc) / device time per launch (CUDA events on the launching streams,
           first launch's start to last launch's end over the launches, timed region);
           .touched_* = bytes the kernel actually requests, from its own counters;
           .traffic = DRAM bytes of this build measured with ncu (profiles/r2_dram_bytes.json).
  cpu_baseline: the reference's own faster-decoder.cc (oracle/_ref), one utterance per host
           thread, on a bounded sample of the same batch; parity_sample classifies that
           sample by SURVEY 8(d)'s protocol (identical / exact tie / real, never-binding /
           binding utterances).

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

# many calls in flight need more hardware queues than the default 8 (see kd_capi.cu); the
# variable must be set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
    ap.add_argument("--total-utts", type=int, default=0,
                    help="strong scaling (BASELINE config 5): this many utterances in total, "
                         "sharded evenly over the ranks")
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
    ap.add_argument("--groups", type=int, default=0,
                    help="lane groups taking turns (default per config; 1: sequential steps)")
    return ap.parse_args()


CONFIG_LANES = {"C1": 64, "C2": 256, "C3": 1024, "C4": 1024}
# Lane groups taking turns (steps in flight) and threads per lane.  A step of C1 / C2 is 64 /
# 256 utterances: one such batch cannot fill 148 SMs, so more steps are kept in flight (the
# lanes of later steps start as soon as a CTA slot is free) and lanes run at the width that is
# best when the chip is full (160 threads, 7 lanes per SM).
CONFIG_GROUPS = {"C1": 16, "C2": 8, "C3": 2, "C4": 2}
CONFIG_THREADS = {"C1": 160, "C2": 160, "C3": 0, "C4": 0}
# tokens alive in one frame must stay below half of this (measured maxima: C1 500, C2 146k,
# C3 49k, C4 142k tokens)
CONFIG_HASH = {"C1": 1 << 14, "C2": 1 << 19, "C3": 1 << 18, "C4": 1 << 19}
# backpointer-store records per lane (0 = library default: a share of the free memory).  The
# store is garbage-collected when it fills up, which costs time: C2 and C4 keep ~1.6 M / ~1.1 M
# records per utterance and get room for the whole utterance.
CONFIG_ARENA = {"C1": 0, "C2": 2_200_000, "C3": 0, "C4": 1_800_000}
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
        if os.environ.get("KD_BENCH_NO_SAMPLER"):  # (diagnosis: is nvidia-smi perturbing the run?)
            return
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

    def mark(self):
        """The timed region starts here: earlier samples are dropped."""
        self.first = len(self.lines)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[getattr(self, "first", 0):]:
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


def measured_traffic(config: str, peak: float, lanes: int, T: int):
    """DRAM bytes per launch of kd_advance_kernel, scaled from the ncu capture of this build
    kept in profiles/r2_dram_bytes.json ({config: {"peak": p, "bytes_per_lane_frame": b}})."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_dram_bytes.json")) as f:
            rec = json.load(f).get(config)
        if rec and float(rec.get("peak", 12.0)) == float(peak):
            return float(rec["bytes_per_lane_frame"]) * lanes * T
    except Exception:
        pass
    return None


def touched_bytes(st: dict, cols: int) -> float:
    """Bytes the kernel actually requests (not the out-degree formula of SURVEY 8(d)): arcs
    that label tables skip are not counted, table traffic is."""
    return (8.0 * st["arcs_evaluated"]
            + st["candidates"] * (16.0 + 16.0 + 8.0 + 32.0)
            + st["slots_claimed"] * (32.0 + 4.0 + 4.0)
            + st["tokens_in"] * (12.0 + 32.0)
            + st["tokens_out"] * 20.0
            + st["eps_arcs"] * (16.0 + 32.0)
            + 4.0 * cols * st["frames"])


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
    if args.total_utts:
        if args.total_utts % world:
            raise SystemExit("--total-utts must be a multiple of the number of ranks")
        share = args.total_utts // world
        lanes = min(share, lanes)
        if share % lanes:
            raise SystemExit("--total-utts: the per-rank share must be a multiple of --lanes")
    # a step = `calls` launches of `lanes` utterances each (strong scaling: the rank's share of
    # the job goes through the decoder's lane groups call after call, as a server would feed it)
    calls = (args.total_utts // world) // lanes if args.total_utts else 1
    scaling = "strong" if args.total_utts else "weak"
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
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
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
    # Two lane groups take turns: step k decodes its utterances in group k % 2.  The steps
    # are enqueued one ahead (kd_decoder_advance_async), so the lanes of step k+1 take over
    # the SMs as the slowest lanes of step k finish, and (end to end) the upload of step k+1
    # runs under the search of step k.  --groups 1 gives strictly sequential steps.
    n_groups = args.groups if args.groups > 0 else CONFIG_GROUPS[args.config]
    dec = capi.LaneDecoder(dg, capi.make_options(**OPTS), max_lanes=lanes * n_groups,
                           hash_capacity=args.hash_capacity or CONFIG_HASH[args.config],
                           arena_records=args.arena_records or CONFIG_ARENA[args.config],
                           threads_per_lane=args.threads_per_lane or CONFIG_THREADS[args.config],
                           chunk_frames=args.chunk_frames)
    # every rank decodes its own utterances (seed differs per rank): weak scaling
    logp = make_device_logprobs(g, lanes * calls, T, args.seed + 7919 * rank, args.peak, dev)
    torch.cuda.synchronize()
    groups = [list(range(k * lanes, (k + 1) * lanes)) for k in range(n_groups)]
    rows = [T] * lanes
    dptrs = [[logp[c * lanes + u].data_ptr() for u in range(lanes)] for c in range(calls)]
    host = None
    if not args.no_e2e:
        host = torch.empty((lanes * calls, T, V), dtype=torch.float32, pin_memory=True)
        host.copy_(logp)
        torch.cuda.synchronize()
        hptrs = [[host[c * lanes + u].data_ptr() for u in range(lanes)] for c in range(calls)]
    setup_s = time.time() - t_setup

    result = {}

    def run_steps(ptrs, mem_kind, steps):
        """`steps` steps, each = InitDecoding + AdvanceDecoding over all frames + GetBestPath of
        `lanes` utterances in ONE kernel launch; step k+1 is enqueued before step k's paths
        are read (when there is more than one lane group)."""
        pending = []
        for k in range(steps * calls):
            t = dec.advance_async(groups[k % n_groups], ptrs[k % calls], rows, V, None, -1, mem_kind,
                                  init=True, finalize=True)
            pending.append(t)
            if len(pending) >= n_groups:
                # zero-copy result (views of the decoder's pinned buffer)
                result["paths"] = dec.results(pending.pop(0), True, copy=False)
        while pending:
            result["paths"] = dec.results(pending.pop(0), True, copy=False)

    def timed(ptrs, mem_kind, steps, warmup):
        if warmup:
            run_steps(ptrs, mem_kind, warmup)
        barrier()
        dec.span_begin()
        t0 = time.perf_counter()
        run_steps(ptrs, mem_kind, steps)
        span_ms, n_launch = dec.span_end()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, span_ms, n_launch

    # one launch alone (nothing else on the GPU): its CUDA-event duration
    run_steps(dptrs, capi.KD_MEM_DEVICE, 1)
    single_ms = dec.last_advance_info()[0]
    # (nvidia-smi is started before the warm-up: its start-up holds the driver for tens of ms;
    # only the samples taken during the timed region are kept)
    sampler = ClockSampler(local_rank)
    sampler.start()
    timed(dptrs, capi.KD_MEM_DEVICE, 0, args.warmup)
    sampler.mark()
    dt, span_ms, n_launch = timed(dptrs, capi.KD_MEM_DEVICE, args.steps, 0)
    clocks = sampler.stop()
    st = dec.stats()  # counters of the last step of every lane group (InitDecoding resets them)
    for k in st:
        if k != "max_tokens":
            st[k] //= min(n_groups, (args.steps + args.warmup + 1) * calls)
    ms_per_step = dt / args.steps * 1e3
    frames_per_step = lanes * calls * T * world
    value = frames_per_step / (dt / args.steps)
    # launches of consecutive steps overlap: the device time per launch is the span from the
    # first launch's start to the last one's end, over the number of launches
    k_ms = span_ms / max(1, n_launch)
    peak_gbs, peak_src = measured_peaks()
    alg_bytes = algorithmic_bytes(st, V)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    touched = touched_bytes(st, V)
    traffic_now = measured_traffic(args.config, args.peak, lanes, T)
    paths = result["paths"]
    d2h = int(paths.d2h_bytes) + int(lanes * 256)  # parked paths + lane states
    n_final = int(np.count_nonzero(paths.reached_final))
    # label sequences of the timed device-input step, kept for the parity sample below (the
    # views are overwritten by later calls)
    n_keep = min(lanes, args.cpu_sample_utts or max(2 * cores, 16))
    kept = [(paths[u].isyms.copy(), paths[u].osyms.copy(), paths[u].total_cost,
             bool(paths[u].reached_final)) for u in range(n_keep)]
    gpu_launches = n_launch

    e2e = None
    if not args.no_e2e:
        dt_h, span_h, n_launch_h = timed(hptrs, capi.KD_MEM_HOST, args.steps, max(1, args.warmup))
        st_h = dec.stats()
        busy_h = sum(st_h[k] for k in ("cycles_cutoff", "cycles_expand", "cycles_closure", "cycles_commit"))
        e2e = {"value": frames_per_step / (dt_h / args.steps), "unit": UNIT,
               # share of the lanes' time spent waiting for rows still on their way from the host
               "input_wait_frac": st_h["cycles_input_wait"] / max(1, busy_h + st_h["cycles_input_wait"]),
               "h2d_bytes_per_step": int(lanes * calls * T * V * 4), "d2h_bytes_per_step": d2h * calls,
               "ms_per_step": dt_h / args.steps * 1e3,
               # what every rank's host link has to sustain; with several ranks uploading at
               # once this, not the search, bounds the end-to-end figure
               "h2d_gb_per_s_per_rank": lanes * calls * T * V * 4 / (dt_h / args.steps) / 1e9,
               "kernel_span_ms_per_step": span_h / max(1, n_launch_h)}

    cpu = None
    check = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import kd_ref
        if kd_ref.available():
            n_s = n_keep
            # (the paths kept are those of the step's last call)
            sample_mats = logp[(calls - 1) * lanes:(calls - 1) * lanes + n_s].cpu().numpy()
            rg = kd_ref.RefGraph(g)
            secs, rpaths, rrf = kd_ref.decode_batch(rg, sample_mats, kd_ref.Options(**OPTS), cores,
                                                    want_paths=True)
            cpu = {"value": n_s * T / secs, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"first {n_s} utterances x {T} frames of the same batch, "
                             f"{cores} threads, oracle/_ref (unmodified faster-decoder.cc)"}
            # parity of the sample (not timed), SURVEY.md section 8(d): label sequences vs the
            # reference; a divergent utterance is an exact tie if its total cost equals the
            # reference's (1e-4 relative is the stated tolerance; ties agree to fp32 rounding)
            # and real otherwise; split by whether GetCutoff ever returned through
            # max_active / min_active on that utterance (oracle counters, reference order)
            from oracle import kd_oracle
            og = kd_oracle.OracleGraph(g)
            _, _, _, _, per = kd_oracle.decode_batch(og, sample_mats, kd_ref.Options(**OPTS), cores,
                                                     mode=kd_oracle.REFERENCE_ORDER, want_paths=False)
            names = list(kd_oracle.STAT_NAMES)
            bmax = per[:, names.index("binding_max")]
            check = {"utterances": n_s, "identical_label_sequences": 0,
                     "max_active_binding_utts": int((bmax > 0).sum()),
                     "max_active_binding_frames": int(bmax.sum()),
                     "min_active_binding_frames": int(per[:, names.index("binding_min")].sum()),
                     "never_binding": {"utts": 0, "identical": 0, "ties": 0, "real": 0},
                     "binding": {"utts": 0, "identical": 0, "ties": 0, "real": 0},
                     "max_rel_cost_diff": 0.0, "reached_final_mismatch": 0}
            for u in range(n_s):
                cls = check["binding" if bmax[u] > 0 else "never_binding"]
                cls["utts"] += 1
                same = (np.array_equal(kept[u][0], rpaths[u].isyms)
                        and np.array_equal(kept[u][1], rpaths[u].osyms))
                rc_ = rpaths[u].total_cost
                rel = abs(kept[u][2] - rc_) / max(1.0, abs(rc_))
                check["max_rel_cost_diff"] = max(check["max_rel_cost_diff"], rel)
                if kept[u][3] != bool(rrf[u]):
                    check["reached_final_mismatch"] += 1
                if same:
                    cls["identical"] += 1
                    check["identical_label_sequences"] += 1
                elif rel <= 1e-6:
                    cls["ties"] += 1
                else:
                    cls["real"] += 1

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": CONFIG_NAME[args.config], "graph": g.stats(),
                       "lanes_per_gpu": lanes, "calls_per_step": calls,
                       "total_utts": lanes * calls * world, "frames": T, "vocab": V, "peak": args.peak,
                       "options": OPTS, "parallelism": f"replicas x{world} (utterances sharded)",
                       "l2_policy": "inputs larger than L2 (log-probs %.2f GB per GPU)" % (lanes * calls * T * V * 4 / 1e9),
                       "threads_per_lane": dec.info()["threads_per_lane"],
                       "lane_groups": n_groups,
                       "pipelining": ("step k+1 enqueued before step k's paths are read; one "
                                      "kernel launch per step (InitDecoding + all frames + "
                                      "GetBestPath)") if n_groups > 1 else "none",
                       "reached_final": n_final},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs,
                         "traffic": traffic_now,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu "
                                         "--set full capture of this build "
                                         "(profiles/r2_dram_bytes.json, per lane-frame x lanes x "
                                         "frames); null = no capture for this config",
                         "traffic_gb_per_s": (traffic_now / (k_ms * 1e-3) / 1e9
                                              if traffic_now and k_ms > 0 else None),
                         "random_sector_peak": {"read": 1409.0, "write": 1128.0, "mixed": 1105.0,
                                                "unit": "GB/s",
                                                "note": "independent random 32-byte sectors over 8 GB "
                                                        "on this GPU type (tools/hbm_random.cu, "
                                                        "profiles/r2_hbm_random_sectors.txt): about "
                                                        "half of `traffic` is of this kind"},
                         "kernel": "kd_advance_kernel", "kernel_ms": k_ms,
                         "kernel_ms_note": "launches of consecutive steps overlap: (end of the "
                                           "last launch - start of the first) / launches, CUDA "
                                           "events on the launching streams, timed region",
                         "kernel_ms_alone": single_ms,
                         "frac_alone": (alg_bytes / (single_ms * 1e-3) / 1e9 / peak_gbs
                                        if single_ms > 0 else None),
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "touched_bytes_per_launch": touched,
                         "touched_frac": touched / (k_ms * 1e-3) / 1e9 / peak_gbs if k_ms > 0 else None,
                         "touched_note": "bytes the kernel requests, from its own counters: arcs "
                                         "evaluated (8 B scanned, 8 B looked up), candidates "
                                         "(16 B written + read, 8 B arc record, 32 B table sector), "
                                         "slots claimed (32 B wipe, 4 B list), tokens in "
                                         "(12 B + 32 B state record) / out (20 B), epsilon arcs "
                                         "(16 B + 32 B sector), the row",
                         "peak_source": peak_src, "counters": st},
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
