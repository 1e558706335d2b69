#!/bin/bash
# usage: tools/sweep_env.sh "label:ENV=V,ENV2=V2 ..." [bench args]  -- one bench line per environment
cfgs="$1"; shift
for c in $cfgs; do
  label=${c%%:*}; envs=${c#*:}
  line=$(env $(echo "$envs" | tr ',' ' ') python bench.py --steps 8 --warmup 2 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1)
  echo "$line" >> gpurun_out/sweep.jsonl
  python - "$label" "$line" <<'PY'
import json,sys
try:
    d=json.loads(sys.argv[2]); r=d["roofline"]; c=r["counters"]; fr=max(1,c["frames"])
    print("%-14s value %.2fM ms/step %.2f span %.2f alone %.2f | cyc/frame cutoff %.0f scan %.0f recomb %.0f closure %.0f commit %.0f"%(sys.argv[1],d["value"]/1e6,d["ms_per_step"],r["kernel_ms"],r["kernel_ms_alone"],c["cycles_cutoff"]/fr,c["cycles_scan"]/fr,(c["cycles_expand"]-c["cycles_scan"])/fr,c["cycles_closure"]/fr,c["cycles_commit"]/fr))
except Exception as e:
    print(sys.argv[1],"FAILED",e, sys.argv[2][:300])
PY
done
