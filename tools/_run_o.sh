for rep in 1 2; do
for lib in tools/_old_libvasr.so tools/_peeled_libvasr.so ""; do
  echo "== lib=${lib:-HEAD(unroll1)}"
  for s in x_dt_proj in_proj out_proj ffn1; do VASR_LIB=$lib python tools/gemm_bench.py $s | cut -c1-100; done
done; done
