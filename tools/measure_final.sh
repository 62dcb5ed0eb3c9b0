#!/bin/bash
# Final one-GPU numbers of the round after the last kernel change (the SOLO instantiation): both bench lines with their extra
# legs, the few-replica configurations, the large single populations, and the ncu captures of push mode on 8M agents.
O=gpurun_out/round2
mkdir -p $O
python bench.py > $O/bench_default_521.json 2> $O/bench_default_521.err
python bench.py --steps 20 --warmup 5 > $O/bench_driver_20.json 2> $O/bench_driver_20.err
python tools/ab_pdl.py > $O/pdl_ab.txt 2>&1
python tools/time_big_population.py > $O/big_population.txt 2>&1
for p in 0 1; do MVB_PUSH=$p python tools/time_big_population.py 2000000:10 8000000:10 50000000:64 50000000:10 | sed "s/^/MVB_PUSH=$p /"; done >> $O/big_population.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"day_kernel|push_kernel" --launch-skip 8 --launch-count 2 -o gpurun_out/r02_push_8M -f python tools/ncu_one_population.py 8000000 10 > $O/ncu_push.log 2>&1
for f in $O/bench_*.json; do python - "$f" <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d.get("e2e") or {}
print(sys.argv[1].split("/")[-1], "value %.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
      "e2e %.3e" % e["value"] if e.get("value") else "", d["clocks"], d.get("cpu_baseline", {}).get("value"))
for k, v in (d.get("extra") or {}).items():
    print("   ", k, "%.3e" % v["agent_days_per_sec"], "us/weekday %.1f" % (1e3 * v["ms_per_weekday"]), "frac %.3f" % v["roofline_frac"])
PY
done
cat $O/pdl_ab.txt $O/big_population.txt
