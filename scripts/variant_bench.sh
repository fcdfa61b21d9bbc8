#!/bin/bash
# Development aid: runs bench.py against pre-built variants of libslpb.so
# (sleipnir_b200/lib/libslpb_<tag>.so) and prints the numbers that matter.
# Usage (on a GPU box): bash scripts/variant_bench.sh w4 w6 w8
for tag in "$@"; do
  cp sleipnir_b200/lib/libslpb_$tag.so sleipnir_b200/lib/libslpb.so
  python bench.py --multistart 0 --batch 0 --no-cpu-baseline > gpurun_out/variant_$tag.json 2> gpurun_out/variant_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.loads(open(f"gpurun_out/variant_{tag}.json").read().strip().splitlines()[-1])
g = d["config"]["device_ms_per_launch_group"]
print(tag, "value %.1f warm %.1f e2e %.1f iters %d | factor %.4f solve %.4f eval %.4f" % (
    d["value"], d["config"]["value_warm_l2"], d["e2e"]["value"], d["e2e"]["iterations"],
    g["factor"], g["solve"], g["eval_full"]), flush=True)
PY
done
