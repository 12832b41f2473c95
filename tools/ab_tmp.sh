export PIR_TIME_DTYPE=fp16
PIR_PWDW_T=3 python -m pytest tests/test_gpu_kernels.py -x -q -k "pwdw" 2>&1 | tail -15
for T in 1 3; do
  export PIR_PWDW_T=$T
  echo "== PIR_PWDW_T=$T" >> gpurun_out/r2_ab9.txt
  python tools/time_pwdw.py 16 256 256 96 288 0 20 >> gpurun_out/r2_ab9.txt
  python tools/time_pwdw.py 16 256 256 48 144 0 20 >> gpurun_out/r2_ab9.txt
  python tools/time_pwdw.py 16 128 128 96 288 0 20 >> gpurun_out/r2_ab9.txt
  PIR_TIME_DTYPE=bf16 python tools/time_pwdw.py 16 256 256 96 288 0 20 >> gpurun_out/r2_ab9.txt
  python tools/time_forward.py 16 256 256 fp16 > gpurun_out/r2_fwd9_T$T.txt 2>&1
  tail -2 gpurun_out/r2_fwd9_T$T.txt >> gpurun_out/r2_ab9.txt
done
cat gpurun_out/r2_ab9.txt
