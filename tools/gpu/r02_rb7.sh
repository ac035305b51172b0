#!/bin/bash
# Round 2: row-band tests with window 7 (one GPU: ranks emulated on one device)
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rowband and not nvlink" > gpurun_out/pytest_rb7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_rb7.log
