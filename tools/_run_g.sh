mkdir -p gpurun_out
for d in 0 1 2; do
  echo "== VASR_TC_DBG=$d"; for s in in_proj x_dt_proj out_proj ffn1; do VASR_TC_DBG=$d python tools/gemm_bench.py $s; done
done > gpurun_out/g_dbg.log 2>&1
cat gpurun_out/g_dbg.log
