mkdir -p gpurun_out
ROWS=60 python tools/gemm_trace.py 48064 192 768 > gpurun_out/j_trace_inproj.log 2>&1
ROWS=60 VASR_TC_DBG=2 python tools/gemm_trace.py 48064 192 768 > gpurun_out/j_trace_inproj_dbg2.log 2>&1
ROWS=60 python tools/gemm_trace.py 48064 384 192 > gpurun_out/j_trace_outproj.log 2>&1
head -75 gpurun_out/j_trace_inproj.log
