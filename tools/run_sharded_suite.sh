#!/bin/bash
# usage: run_sharded_suite.sh N "lattice:u:nops[:check]" ...   (under gpurun --gpus N)
N=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read lat u nops chk <<< "$spec"
  extra=""; [ "$chk" = "check" ] && extra="--check-single"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    tools/bench_sharded.py --lattice $lat --u $u --n-ops $nops --steps 2 $extra --json gpurun_out/shard_${lat}_n${N}.json \
    > gpurun_out/shard_${lat}_n${N}.log 2>&1
  echo "== $lat N=$N rc=$?"; tail -2 gpurun_out/shard_${lat}_n${N}.log | cut -c1-1800
done
