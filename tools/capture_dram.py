#!/usr/bin/env python
"""DRAM bytes per lane-frame of kd_advance_kernel for the bench configs, measured with ncu
(dram__bytes_read.sum + dram__bytes_write.sum of one launch of FRAMES frames x the config's
lanes).  Writes profiles/r2_dram_bytes.json, which bench.py reads for roofline.traffic.

    python tools/capture_dram.py [C1 C2 C3 C4]      (on a GPU box; ~1 min per config)
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRAMES = 200
LANES = {"C1": 64, "C2": 256, "C3": 1024, "C4": 1024}
out_path = os.path.join(ROOT, "profiles", "r2_dram_bytes.json")
try:
    rec = json.load(open(out_path))
except Exception:
    rec = {}
for cfg in (sys.argv[1:] or ["C1", "C2", "C3", "C4"]):
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
           "--clock-control", "none", "-k", "regex:kd_advance", "-s", "1", "-c", "1", "--csv",
           sys.executable, os.path.join(ROOT, "bench.py"), "--config", cfg, "--frames", str(FRAMES),
           "--groups", "1", "--steps", "1", "--warmup", "1", "--no-e2e", "--no-cpu-baseline"]
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    vals = {}
    for row in csv.reader(io.StringIO(p.stdout)):
        if len(row) > 3 and row[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum",
                                        "gpu__time_duration.sum"):
            v = float(row[-1].replace(",", ""))
            unit = row[-2]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
                     "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "nsecond": 1e-9, "usecond": 1e-6,
                     "msecond": 1e-3, "second": 1}.get(unit, 1)
            vals[row[-3]] = v * scale
    if "dram__bytes_read.sum" not in vals:
        print(cfg, "capture failed", p.stdout[-800:], p.stderr[-800:])
        continue
    total = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
    rec[cfg] = {"peak": 12.0, "lanes": LANES[cfg], "frames": FRAMES,
                "dram_read_bytes": vals["dram__bytes_read.sum"],
                "dram_write_bytes": vals["dram__bytes_write.sum"],
                "bytes_per_lane_frame": total / (LANES[cfg] * FRAMES),
                "launch_seconds_under_ncu": vals.get("gpu__time_duration.sum"),
                "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                       "-k regex:kd_advance -s 1 -c 1 python bench.py --config %s --frames %d --groups 1 "
                       "--steps 1 --warmup 1" % (cfg, FRAMES)}
    print(cfg, rec[cfg])
    json.dump(rec, open(out_path, "w"), indent=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "r2_dram_bytes.json"), "w"), indent=1)
# (the file is written under profiles/ of the box's copy: also leave it where gpurun merges it back)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "r2_dram_bytes.json"), "w"), indent=1)
