cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py -x -q -k "pwdw" 2>&1 | tail -3
for a in "96 288 0" "48 144 0"; do
  PIR_PWDW_T=1 python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed "s/^/old bf16 /"
  PIR_PWDW_T=3 python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed "s/^/new bf16 /"
  PIR_PWDW_T=3 PIR_TIME_DTYPE=fp16 python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed "s/^/new fp16 /"
done | tee gpurun_out/r2_pwdw_ab6.txt
PIR_PWDW_T=3 python -m pytest tests/test_gpu_kernels.py -x -q -k "pwdw" 2>&1 | tail -3
