#!/bin/bash
# scan kernel variants at config-2 shape; run on the GPU box
for r in 2 3; do echo "RPL=$r"; VASR_SCAN_RPL=$r python tools/scan_bench.py --quick; done
