#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -6 gpurun_out/r2d_pytest.log
bash tools/gpu_exp.sh r2d c3 5 default: nohits:LYNSE_B200_TC_HITS=0 epi1:LYNSE_B200_TC_EPI=1 bias:LYNSE_B200_TC_L2_BIAS=1 noscan:LYNSE_B200_TC_DEBUG=4
bash tools/gpu_exp.sh r2d c4t 3 default:
bash tools/gpu_exp.sh r2d c4 3 default:
bash tools/gpu_exp.sh r2d c2 10 default:
