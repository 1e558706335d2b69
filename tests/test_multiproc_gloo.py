"""world_size-2 gloo run of the multi-GPU host logic (replicas only: shard the utterances,
no collective on the search path; MAX-reduce the time; gather the results)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from common import small_graph
from kaldi_decoder_b200 import synth
from oracle import kd_oracle, kd_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "kaldi-decoder_b200", "python"))
    sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np
    import torch.distributed as dist
    from kaldi_decoder_b200 import parallel, synth
    from oracle import kd_oracle, kd_ref
    from common import small_graph
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank, world = dist.get_rank(), dist.get_world_size()
    g = small_graph("HLG")
    n_utts = 7
    lo, hi = parallel.shard_range(n_utts, world, rank)
    og = kd_oracle.OracleGraph(g)           # every rank holds its own graph replica
    opts = kd_ref.Options(beam=20.0, max_active=7000)
    res = {{}}
    for u in range(lo, hi):
        d = kd_oracle.OracleDecoder(og, opts, kd_oracle.CANONICAL)
        d.decode(synth.make_logprobs(g, 60, seed=u, peak=8))
        res[u] = [int(x) for x in d.get_best_path().osyms]
    dist.barrier()
    t = parallel.max_over_ranks(1.0 + rank)   # the slowest rank defines the job time
    allres = parallel.gather_objects(res, dst=0)
    if rank == 0:
        merged = {{}}
        for r in allres: merged.update(r)
        print(json.dumps({{"t": t, "res": {{str(k): v for k, v in merged.items()}}}}))
    dist.destroy_process_group()
""")


def test_two_rank_sharded_decode_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    import json
    got = json.loads(outs[0][0].strip().splitlines()[-1])
    assert got["t"] == 2.0
    g = small_graph("HLG")
    og = kd_oracle.OracleGraph(g)
    opts = kd_ref.Options(beam=20.0, max_active=7000)
    for u in range(7):
        d = kd_oracle.OracleDecoder(og, opts, kd_oracle.CANONICAL)
        d.decode(synth.make_logprobs(g, 60, seed=u, peak=8))
        assert got["res"][str(u)] == [int(x) for x in d.get_best_path().osyms]
