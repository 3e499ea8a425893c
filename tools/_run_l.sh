for d in 3 1; do echo "== VASR_TC_DBG=$d"; ROWS=0 VASR_TC_DBG=$d python tools/gemm_trace.py 48064 192 768 2>&1 | grep -A7 "^producer"; done
