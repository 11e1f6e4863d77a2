#!/bin/bash
# C4 at N = 2 with the final build (torchrun, one rank per GPU)
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2ae_bench_n2.json 2> gpurun_out/r2ae_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2ae_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ae_bench_n2.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'scaling', d['scaling'])
print('parity', json.dumps(d['parity'])[:600])
print({k: v for k, v in d.items() if k in ('exchange', 'gpu_launches')})
PY
