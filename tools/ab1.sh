#!/bin/bash
# like ab.sh, for the single-replica and the 8M-agent single-population cases
for round in 1 2; do
  for v in "$@"; do
    cp tools/ab/lib$v.so motivate_b200/libmotivate_b200.so
    python bench.py --replicas 1 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v R=1', '%.4f' % d['roofline']['frac'], '%.4f' % d['ms_per_step'])"
  done
done
for v in "$@"; do
  cp tools/ab/lib$v.so motivate_b200/libmotivate_b200.so
  python bench.py --partitioned --agents 8000000 --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v 8M', d['value'], d['ms_per_step'])"
done
