cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py -x -q -k "pwdw" 2>&1 | tail -3
for a in "96 288 0" "48 144 0" "192 576 0"; do
  PROMPTIR_B200_LIB=$GRAFT_REPO_ROOT/promptir_b200/ab/libpromptir_b200_r1.so python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed "s/^/r1  /"
  python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed "s/^/new /"
done | tee gpurun_out/r2_pwdw_ab5.txt
