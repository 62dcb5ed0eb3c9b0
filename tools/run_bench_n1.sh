set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo rc=$?
tail -c 3500 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
