#!/bin/bash
# C4 at N = 8 and N = 4 with the final build (torchrun, one rank per GPU), on one 8-GPU box
mkdir -p gpurun_out
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r2af_bench_n$n.json 2> gpurun_out/r2af_bench_n$n.err; echo "bench n$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2af_bench_n$n.json').read().strip().splitlines()[-1])
    print('n_gpus', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'scaling', d['scaling'], d.get('exchange'))
    print(json.dumps(d['parity'].get('sharded_sweep_vs_unsharded')))
except Exception as e:
    print('n$n failed', e)
PY
done
