#!/bin/bash
# r02 multi-GPU validation at N=2: bench.py under torchrun (replicas + pool-sharded 3x4 + sharded 4x4 through fh_comm)
O=gpurun_out; mkdir -p $O
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 200 --warmup 5 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err
echo "rc=$?"; tail -5 $O/r02_bench_n$N.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r02_bench_n$N.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"])
    print("pool_sharded", d.get("pool_sharded"))
    s=d.get("sharded_4x4")
    if s: print("sharded", {k:s[k] for k in ("communicator","energy","best","phase_seconds","hf_screening","gradients_per_s","pool_scan_hbm_frac_per_gpu") if k in s})
except Exception as e: print("parse failed", e)
PY
