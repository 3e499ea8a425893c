#!/bin/bash
# One measurement pass on a GPU box (run through gpurun): GPU tests, both bench arms, and the ncu launch list of
# one bench step.  Outputs land in gpurun_out/<tag>_*; copy what should be judged into profiles/.
#   gpurun --timeout 900 -- 'bash tools/measure_pass.sh r01g'
tag=${1:-pass}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
cat gpurun_out/${tag}_bench.json
