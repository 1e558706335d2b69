"""Shared helpers of the test-suite (graphs, option sets, comparisons)."""
import functools

import numpy as np

from kaldi_decoder_b200 import synth
from oracle import kd_oracle, kd_ref

OPTION_SETS = [
    dict(beam=20.0, max_active=7000, min_active=20),
    dict(beam=8.0, max_active=50, min_active=5),
    dict(beam=16.0, max_active=2**31 - 1, min_active=0),
    dict(beam=12.0, max_active=200, min_active=20),
    dict(beam=20.0, max_active=30, min_active=29),
]


@functools.lru_cache(maxsize=None)
def small_graph(name: str):
    if name == "H":
        return synth.make_h(50)
    if name == "HL":
        return synth.make_hl(2000, 50, seed=1)
    if name == "HLG":
        return synth.make_hlg(2000, (100, 200), 50, seed=2)
    raise KeyError(name)


def sorted_tokens(states, costs):
    o = np.argsort(states, kind="stable")
    return np.asarray(states)[o], np.asarray(costs)[o]


def strip(x):
    x = np.asarray(x)
    return x[x != 0]


def same_labels(a, b) -> bool:
    return np.array_equal(strip(a.ilabels), strip(b.ilabels)) and \
        np.array_equal(strip(a.olabels), strip(b.olabels))


def rel_close(a: float, b: float, tol: float = 1e-4) -> bool:
    return abs(a - b) <= tol * max(1.0, abs(a), abs(b))


GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")
GOLDEN_CASES = ["h20_bind", "h20_default", "hl300", "hlg300_peaky", "hlg300_bind", "hlg300_nobeam"]


class GoldenCase:
    """One tests/golden/*.npz fixture (written by tests/golden/make_golden.py from oracle/_ref)."""

    def __init__(self, name):
        import os
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.name = name
        self.graph = synth.Graph(int(z["num_states"]), int(z["start"]), z["row_off"], z["ilabel"],
                                 z["olabel"], z["weight"], z["nextstate"], z["final"],
                                 lm={"vocab": int(z["vocab"])}, name=name)
        o = z["opts"]
        self.opts = dict(beam=float(o[0]), max_active=int(o[1]), min_active=int(o[2]),
                         beam_delta=float(o[3]), hash_ratio=float(o[4]))
        self.n_utts, self.T = int(z["n_utts"]), int(z["T"])

    def logp(self, u):
        return self.z[f"logp{u}"]

    def tokens(self, u):
        """list over frames 0..T of (states, costs) in the reference's list order"""
        n = self.z[f"tok_n{u}"]
        st, co = self.z[f"tok_states{u}"], self.z[f"tok_costs{u}"]
        out, p = [], 0
        for k in n:
            out.append((st[p:p + k], co[p:p + k]))
            p += int(k)
        return out

    def best(self, u, use_final_probs=True):
        t = "t" if use_final_probs else "f"
        z = self.z
        return kd_ref.BestPath(bool(z[f"ok_{t}{u}"]), z[f"il_{t}{u}"], z[f"ol_{t}{u}"],
                               z[f"gw_{t}{u}"], z[f"aw_{t}{u}"], z[f"fin_{t}{u}"])

    def reached_final(self, u):
        return bool(self.z[f"reached_final{u}"])


def parity_table(g, mats, opts, gpu_paths, threads=None):
    """SURVEY.md section 8(d) parity protocol for one batch: the device's best paths
    (`gpu_paths[u]`: RawPath) against the compiled reference (oracle/_ref).  Every utterance
    is classified identical / exact tie / real (label sequences differ; a tie if the total
    costs agree to 1e-6 relative, the float32 rounding of the summed path weights), split by
    whether GetCutoff ever returned through max_active on it (the reference's pruning is
    order dependent from the first such frame on)."""
    import os
    threads = threads or len(os.sched_getaffinity(0))
    mats = np.ascontiguousarray(np.stack(mats), dtype=np.float32)
    ropts = kd_ref.Options(**opts)
    _, rpaths, rrf = kd_ref.decode_batch(kd_ref.RefGraph(g), mats, ropts, threads, want_paths=True)
    _, _, _, _, per = kd_oracle.decode_batch(kd_oracle.OracleGraph(g), mats, ropts, threads,
                                             mode=kd_oracle.REFERENCE_ORDER, want_paths=False)
    names = list(kd_oracle.STAT_NAMES)
    bmax = per[:, names.index("binding_max")]
    out = {"utterances": len(mats), "identical": 0,
           "max_active_binding_utts": int((bmax > 0).sum()),
           "max_active_binding_frames": int(bmax.sum()),
           "min_active_binding_frames": int(per[:, names.index("binding_min")].sum()),
           "never_binding": {"utts": 0, "identical": 0, "ties": 0, "real": 0},
           "binding": {"utts": 0, "identical": 0, "ties": 0, "real": 0},
           "max_rel_cost_diff": 0.0, "reached_final_mismatch": 0, "ok_mismatch": 0}
    for u in range(len(mats)):
        cls = out["binding" if bmax[u] > 0 else "never_binding"]
        cls["utts"] += 1
        p, r = gpu_paths[u], rpaths[u]
        if p.ok != r.ok:
            out["ok_mismatch"] += 1
        if bool(p.reached_final) != bool(rrf[u]):
            out["reached_final_mismatch"] += 1
        rel = abs(p.total_cost - r.total_cost) / max(1.0, abs(r.total_cost)) if r.ok and p.ok else 0.0
        out["max_rel_cost_diff"] = max(out["max_rel_cost_diff"], rel)
        if np.array_equal(p.isyms, r.isyms) and np.array_equal(p.osyms, r.osyms):
            cls["identical"] += 1
            out["identical"] += 1
        elif rel <= 1e-6:
            cls["ties"] += 1
        else:
            cls["real"] += 1
    return out


def build_cpp_drop_in(out_dir):
    """Compiles tests/cpp/drop_in.cc -- a caller written against the reference's include paths
    -- with the B200 sources and links it against libkd_b200.so.  Returns the binary's path."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cxx = shutil.which("g++")
    if cxx is None:
        return None
    pkg = os.path.join(root, "kaldi-decoder_b200")
    out = os.path.join(str(out_dir), "drop_in")
    cmd = [cxx, "-std=c++17", "-O1", "-I", os.path.join(pkg, "compat"), "-I", root,
           "-I", os.path.join(pkg, "csrc", "minifst"), "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "cpp", "drop_in.cc"),
           os.path.join(pkg, "csrc", "faster-decoder.cc"), os.path.join(pkg, "csrc", "decodable-ctc.cc"),
           os.path.join(pkg, "csrc", "fst-io.cc"), "-L", os.path.join(pkg, "lib"), "-lkd_b200",
           "-Wl,-rpath," + os.path.join(pkg, "lib"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out
