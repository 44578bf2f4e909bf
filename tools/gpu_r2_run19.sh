#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ivf_client.py tests/test_gpu_parity.py -m gpu -q -x -k "ivf or IVF" 2>&1 | tail -3
for c in 1 0; do echo "== contiguous=$c"; LYNSE_B200_IVF_CONTIGUOUS=$c LYNSE_B200_IVF_TRACE=1 timeout 600 python tools/ivf_probe.py 2>&1 | tail -8; done
