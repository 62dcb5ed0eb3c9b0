#!/bin/bash
# One-GPU measurement set of round 2 (run under gpurun): everything BASELINE.md §4 quotes for one GPU comes from ONE call
# of this script on one box.  Results: gpurun_out/round2/.
#   1  default bench line (521 weekdays: the sustained, power-capped number) with its extra legs (c1, c2, c5)
#   2  the driver's setting (20 weekdays)
#   3  few-replica configurations with and without programmatic dependent launch
#   4  one large population, 2M .. 50M agents: the library's own choices, then the cold links pulled (MVB_PUSH=0) and pushed (=1)
#   5  set-up: mvb_create_batch next to the plain copy of the same pinned buffers
#   6  ncu launch list of the 8-replica bench command (plain run first, as the profiling rules ask)
O=gpurun_out/round2
mkdir -p $O
python bench.py > $O/bench_default_521.json 2> $O/bench_default_521.err
python bench.py --steps 20 --warmup 5 > $O/bench_driver_20.json 2> $O/bench_driver_20.err
python tools/ab_pdl.py > $O/pdl_ab.txt 2>&1
python tools/time_big_population.py > $O/big_population.txt 2>&1
for p in 0 1; do MVB_PUSH=$p python tools/time_big_population.py 2000000:10 8000000:10 50000000:64 50000000:10 | sed "s/^/MVB_PUSH=$p /"; done >> $O/big_population.txt 2>&1
MVB_TIMING=1 python tools/time_create.py > $O/create.txt 2>&1
MVB_TIMING=1 python tools/time_create.py --replicas 8 --repeat 3 >> $O/create.txt 2>&1
CMD="python bench.py --replicas 8 --steps 6 --warmup 3 --no-cpu-baseline --no-extra"
$CMD > $O/prof_plain.json 2> $O/prof_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
for f in $O/bench_*.json; do python - "$f" <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d.get("e2e") or {}
print(sys.argv[1].split("/")[-1], "value %.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
      "e2e %.3e" % e["value"] if e.get("value") else "", d["clocks"])
for k, v in (d.get("extra") or {}).items():
    print("   ", k, "%.3e" % v["agent_days_per_sec"], "us/weekday %.1f" % (1e3 * v["ms_per_weekday"]), "frac %.3f" % v["roofline_frac"])
PY
done
cat $O/pdl_ab.txt $O/big_population.txt; grep -E "plain|mvb_create" $O/create.txt
