set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/f_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu.log 2>&1
tail -3 gpurun_out/f_pytest.log; cat gpurun_out/f_beam.log; cat gpurun_out/f_bench.json
