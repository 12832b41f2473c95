cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py -x -q -k "pwdw" 2>&1 | tail -3
for a in "96 256 1" "48 128 1"; do
  PIR_TIME_DTYPE=fp16 python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed "s/^/fp16 /"
  python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed "s/^/bf16 /"
done | tee gpurun_out/r2_pwdw_ab7.txt
