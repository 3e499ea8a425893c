python -m pytest tests -m gpu -x -q -k "scan or forward or golden" 2>&1 | tail -2
for lib in tools/_old_libvasr.so ""; do echo "== ${lib:-HEAD}"; VASR_LIB=$lib python tools/scan_bench.py --quirk; done
