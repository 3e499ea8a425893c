mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in 0 1; do echo "== VASR_TC_WRES=$w"; for s in in_proj ffn1 ctc; do VASR_TC_WRES=$w python tools/gemm_bench.py $s; done; done
VASR_TC_WRES=0 python tools/step_profile.py 2>&1 | tail -1
python tools/step_profile.py 2>&1 | tail -1
