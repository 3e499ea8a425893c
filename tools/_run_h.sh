mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/gemm_bench.py ffn1
python tools/step_profile.py 2>&1 | tail -4
