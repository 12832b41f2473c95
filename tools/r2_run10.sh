cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv3x3" 2>&1 | tail -12
PIR_CONV3=0 timeout 120 python tools/time_conv3.py fp16
timeout 120 python tools/time_conv3.py fp16
