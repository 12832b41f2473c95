export PIR_TIME_DTYPE=fp16
python -m pytest tests/test_gpu_kernels.py -x -q -k "pwdw" 2>&1 | tail -3
for L in libbase_r2c.so libpromptir_b200.so; do
  export PROMPTIR_B200_LIB=$PWD/promptir_b200/$L
  echo "== $L" >> gpurun_out/r2_ab11.txt
  python tools/time_pwdw.py 16 256 256 96 256 1 20 >> gpurun_out/r2_ab11.txt
  python tools/time_pwdw.py 16 256 256 96 288 0 20 >> gpurun_out/r2_ab11.txt
  python tools/time_pwdw.py 16 256 256 48 128 1 20 >> gpurun_out/r2_ab11.txt
  python tools/time_pwdw.py 16 256 256 48 144 0 20 >> gpurun_out/r2_ab11.txt
  python tools/time_pwdw.py 16 128 128 96 256 1 20 >> gpurun_out/r2_ab11.txt
  python tools/time_forward.py 16 256 256 fp16 > gpurun_out/r2_fwd11_$L.txt 2>&1
  tail -2 gpurun_out/r2_fwd11_$L.txt >> gpurun_out/r2_ab11.txt
done
cat gpurun_out/r2_ab11.txt
