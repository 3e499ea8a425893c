for d in 0 4 2 6; do echo "== VASR_TC_DBG=$d"; ROWS=0 VASR_TC_DBG=$d python tools/gemm_trace.py 48064 192 768 2>&1 | grep -A14 "^producer"; done
for d in 0 4; do echo "== VASR_TC_DBG=$d"; for s in in_proj out_proj; do VASR_TC_DBG=$d python tools/gemm_bench.py $s; done; done
