#!/bin/bash
# Round 2, call a (1 GPU): (1) parity + A/B of the two opt-in tunings round 1 never ran; (2) the reference's own driver adjudicates
# Base_B200 for all kernels at default AND BASELINE sizes, with same-size incumbents (tools/ref_adjudicate.py);
# (3) the exchange incumbents on one rank.
TAG=${TAG:-r02_a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
free -g | head -2; nproc
timeout 150 python tools/time_quick.py scan_line sort_hist > gpurun_out/${TAG}_optin.log 2>&1; echo "opt-in A/B rc=$?"
grep -E "parity|scan |sort " gpurun_out/${TAG}_optin.log | head -30
timeout 1500 python tools/ref_adjudicate.py --phases default checksum timing --out gpurun_out/${TAG}_adjudicate \
    > gpurun_out/${TAG}_adjudicate.log 2>&1; echo "adjudicate rc=$?"
tail -50 gpurun_out/${TAG}_adjudicate.log
timeout 90 python tools/incumbent_suite.py --fast --groups exchange --timeout 40 --out gpurun_out/${TAG}_incumbent_exchange \
    > gpurun_out/${TAG}_incumbent_exchange.log 2>&1; echo "incumbent exchange rc=$?"; tail -6 gpurun_out/${TAG}_incumbent_exchange.log
