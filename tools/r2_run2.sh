set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py -x -q -k "pwdw" 2>&1 | tail -5
for a in "96 256 1" "96 288 0" "48 128 1" "48 144 0" "192 512 1" "192 576 0" "160 432 1"; do
  PROMPTIR_B200_LIB=$GRAFT_REPO_ROOT/promptir_b200/ab/libpromptir_b200_r1.so python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed 's/^/r1   /'
  python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed 's/^/new  /'
  PIR_PWDW_NKB0=1 python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed 's/^/nkb0 /'
done | tee gpurun_out/r2_pwdw_ab1.txt
