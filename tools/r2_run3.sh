cd $GRAFT_REPO_ROOT
for a in "96 256 1" "96 288 0" "48 128 1" "192 512 1"; do
  for v in r1 V1 V2 V3 V4; do
    lib=$GRAFT_REPO_ROOT/promptir_b200/ab/lib_$v.so; [ $v = r1 ] && lib=$GRAFT_REPO_ROOT/promptir_b200/ab/libpromptir_b200_r1.so
    PROMPTIR_B200_LIB=$lib python tools/time_pwdw.py 16 256 256 $a 20 2>&1 | grep "pwdw B" | sed "s/^/$v /"
  done
done | tee gpurun_out/r2_pwdw_ab2.txt
