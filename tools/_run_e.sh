set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/e_pytest.log
python tools/beam_bench.py > gpurun_out/e_beam.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/e_ref.json 2> gpurun_out/e_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/e_ncu.log 2>&1
tail -3 gpurun_out/e_pytest.log; cat gpurun_out/e_beam.log; cat gpurun_out/e_bench.json
