#!/bin/bash
# Multi-GPU measurement pass (gpurun --gpus N): weak scaling (64 utterances per GPU), strong scaling on BASELINE
# configs[2] (512 utterances in total) and configs[4] (FakeQuantize model, 256 utterances in total).
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/multi_gpu_pass.sh 8 r02'
N=${1:-2}; tag=${2:-pass}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
mkdir -p gpurun_out
run --steps 20 --warmup 5 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
run --steps 20 --warmup 5 --global-batch 512 --no-extras > gpurun_out/${tag}_bench_${N}gpu_strong512.json 2> gpurun_out/${tag}_bench_${N}gpu_strong512.err
run --steps 20 --warmup 5 --quantized --global-batch 256 --no-extras > gpurun_out/${tag}_bench_${N}gpu_quant256.json 2> gpurun_out/${tag}_bench_${N}gpu_quant256.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_1gpu_same_box_${N}.json 2>/dev/null
for f in gpurun_out/${tag}_bench_${N}gpu.json gpurun_out/${tag}_bench_${N}gpu_strong512.json gpurun_out/${tag}_bench_${N}gpu_quant256.json gpurun_out/${tag}_bench_1gpu_same_box_${N}.json; do
  python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d['n_gpus'], d['scaling'], d['config']['global_batch'], round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],3), round(d['e2e']['value']))
except Exception as e: print('$f', 'FAILED', e)
"; done
tail -2 gpurun_out/${tag}_bench_${N}gpu.err
