#!/bin/bash
# tuning sweep of the staged-bound schedule (profiles/r02_gate_schedule.md); run on the GPU box from the repo root
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-north-star "$@" > gpurun_out/gs_$tag.json 2> gpurun_out/gs_$tag.err
  python - "$tag" <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/gs_{tag}.json").read().strip().splitlines()[-1])
    g = d["extra"]["gated_variance"]; r = d["roofline"]
    print(tag, "ms", round(d["ms_per_step"], 2), "full_rows", round(g["fraction_full"], 4), "vg_launches", r["launches"], "vg_share", round(r["share_of_step"], 3), "frac", round(r["frac"], 3), "kb_share", round(d["extra"]["kernel_build_share"], 3), "launches", d["gpu_launches"], flush=True)
except Exception as e:
    print(tag, "FAILED", e, flush=True)
P
}
