set -x
cd /root/repo
timeout 600 python -m pytest tests/test_partitioned.py -m gpu -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for H in 10 64; do for P in 0 1; do
MVB_PUSH=$P timeout 300 $TR tools/run_partitioned.py 12000000 $H 20 2>&1 | grep -v "^\*\|OMP_NUM" | tail -2
done; done
MVB_PUSH=1 python /tmp/x.py 2>/dev/null
