#!/bin/bash
# multi-GPU: C4 strong scaling through bench.py (parity incl. sharded sweep vs unsharded asserted inside)
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2i_topo_$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2i_bench_n$N.json 2> gpurun_out/r2i_bench_n$N.err; echo "bench N=$N rc=$?"
tail -5 gpurun_out/r2i_bench_n$N.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2i_bench_n$N.json'))
    print('N', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'])
    print('parity', d['parity'].get('sharded_sweep_vs_unsharded'), d['parity']['cost_rel_max'], d['parity']['first_lm_step_rel_max'])
    print('exchange', d.get('exchange'))
    r=d['roofline']; print({k:round(v['us'],1) for k,v in r.get('passes',{}).items()})
except Exception as e: print('ERR', e)
PY
