#!/bin/bash
# One-GPU measurement set of the BASELINE configurations (run under gpurun): JSON lines into gpurun_out/round/.
mkdir -p gpurun_out/round
python bench.py > gpurun_out/round/c3_default.json 2>gpurun_out/round/c3_default.err
python bench.py --replicas 8 --steps 200 --no-cpu-baseline > gpurun_out/round/c3_8replicas.json 2>/dev/null
python bench.py --replicas 1 --steps 521 --no-cpu-baseline > gpurun_out/round/c2_1x1M.json 2>/dev/null
python bench.py --replicas 1 --agents 111166 --steps 521 --no-cpu-baseline > gpurun_out/round/c1shape_1x111k.json 2>/dev/null
python bench.py --replicas 8 --agents 111166 --steps 521 --no-cpu-baseline > gpurun_out/round/c1shape_8x111k.json 2>/dev/null
python bench.py --sweep --replicas 256 --agents 100000 --steps 200 --no-cpu-baseline > gpurun_out/round/c5_sweep_256x100k.json 2>/dev/null
python bench.py --partitioned --agents 8000000 --steps 30 --warmup 3 2>/dev/null | tail -1 > gpurun_out/round/c4_8M_1gpu.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/round/reference_arm.json 2>/dev/null
for f in gpurun_out/round/*.json; do python - "$f" <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d.get("e2e") or {}
print(sys.argv[1].split("/")[-1], "value %.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"],
      "frac %.3f" % d["roofline"]["frac"] if "roofline" in d else "", "e2e %.3e" % e["value"] if e.get("value") else "")
PY
done
