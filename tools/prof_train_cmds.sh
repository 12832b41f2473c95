set -x
T='python tools/bench_train.py --steps 1 --warmup 1 --no-kernels'
X='python tools/bench_xrestormer.py --steps 1 --warmup 1'
$T > gpurun_out/plain_train.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'wgrad_kernel' -s 6 -c 6 -f -o gpurun_out/prof_train_wgrad_r1 $T > gpurun_out/ncu_train.log 2>&1
$X > gpurun_out/plain_x.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'ocab_kernel' -s 1 -c 2 -f -o gpurun_out/prof_ocab_r1 $X > gpurun_out/ncu_x.log 2>&1
$T > gpurun_out/plain_train2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train_r1.csv $T > gpurun_out/ncu_train2.log 2>&1
echo done
