set -x
mkdir -p gpurun_out
python tools/scan_bench.py --quick > gpurun_out/f_scan_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:scan_seq -c 1 -s 3 -o gpurun_out/scan_full5 -f python tools/scan_bench.py --quick > gpurun_out/f_ncu.log 2>&1
cat gpurun_out/f_scan_plain.log; tail -3 gpurun_out/f_ncu.log
