#!/bin/bash
# A/B of several builds of libmotivate_b200.so ON THE SAME BOX (boxes differ by a few percent): tools/ab/lib*.so are
# swapped in turn, two rounds, bench.py kernel-only numbers.  usage (under gpurun): tools/ab.sh "8 64" A B C
reps="$1"; shift
for round in 1 2; do
  for v in "$@"; do
    cp tools/ab/lib$v.so motivate_b200/libmotivate_b200.so
    for R in $reps; do
      python bench.py --replicas $R --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | \
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v R=$R', '%.4f' % d['roofline']['frac'], '%.4f' % d['ms_per_step'])"
    done
  done
done
