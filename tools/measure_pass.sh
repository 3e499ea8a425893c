#!/bin/bash
# One measurement pass on a GPU box (run through gpurun): GPU tests, both bench arms, the ncu launch list of one
# bench step and one `ncu --set full` capture of the scan kernel.  Outputs land in gpurun_out/<tag>_*; copy what
# should be judged into profiles/ (scan_ncu.json -> profiles/scan_ncu.json is what bench.py's roofline.traffic reads).
#   gpurun --timeout 1500 -- 'bash tools/measure_pass.sh r02a'
tag=${1:-pass}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu.log 2>&1
python tools/scan_bench.py --quick > gpurun_out/${tag}_scan_bench.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:scan_rp|scan_seq" -c 1 -f -o gpurun_out/${tag}_scan_full \
    python tools/scan_bench.py --quick > gpurun_out/${tag}_scan_ncu.log 2>&1 && \
python tools/ncu_scan_json.py gpurun_out/${tag}_scan_full.ncu-rep 64 751 384 64 > gpurun_out/${tag}_scan_ncu.json
tail -3 gpurun_out/${tag}_pytest.log
cat gpurun_out/${tag}_bench.json
