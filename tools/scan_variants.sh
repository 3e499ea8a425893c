#!/bin/bash
# A/B of the scan kernels on one GPU box with the -DVASR_DEBUG build (python velocity-asr_b200/build.py --debug).
export VASR_LIB=velocity-asr_b200/velocity_asr/libvasr_dbg.so
B="${@:-49 64 74 128}"
echo "old (scan_seq_kernel)";           VASR_SCAN_OLD=1 python tools/scan_bench.py --B $B
echo "rp=1 unsplit";                    VASR_SCAN_RP=1 VASR_SCAN_SPLIT=0 python tools/scan_bench.py --B $B
echo "rp=1 split (rule)";               VASR_SCAN_RP=1 python tools/scan_bench.py --B $B
echo "rp=1 split 296 slots";            VASR_SCAN_RP=1 VASR_SCAN_SPLIT=296 python tools/scan_bench.py --B $B
echo "rp=2 unsplit";                    VASR_SCAN_RP=2 VASR_SCAN_SPLIT=0 python tools/scan_bench.py --B $B
echo "rp=2 split (rule)";               VASR_SCAN_RP=2 python tools/scan_bench.py --B $B
echo "rp=2 split 148 slots";            VASR_SCAN_RP=2 VASR_SCAN_SPLIT=148 python tools/scan_bench.py --B $B
