#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; tail -6 gpurun_out/r2c_pytest.log
bash tools/gpu_exp.sh r2c c3 5 default: bias:LYNSE_B200_TC_L2_BIAS=1 epi1:LYNSE_B200_TC_EPI=1 nohits:LYNSE_B200_TC_HITS=0 bias_epi1:LYNSE_B200_TC_L2_BIAS=1,LYNSE_B200_TC_EPI=1 noscan:LYNSE_B200_TC_DEBUG=4 noread:LYNSE_B200_TC_DEBUG=2
bash tools/gpu_exp.sh r2c c4t 3 default:
