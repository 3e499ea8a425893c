#!/bin/bash
# scan kernel variants at config-2 shape; run on the GPU box
for w in 3 4; do echo "WARPS=$w"; VASR_SCAN_WARPS=$w python tools/scan_bench.py --quick; done
echo "RPL=1"; VASR_SCAN_RPL=1 python tools/scan_bench.py --quick
