set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for a in "96 256 1" "96 288 0" "48 128 1" "48 144 0" "192 512 1" "192 576 0"; do python tools/time_pwdw.py 16 256 256 $a 20; done 2>&1 | grep "pwdw B" | tee gpurun_out/r2_pwdw_base_times.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_base_bf16.json 2> gpurun_out/r2_bench_base_bf16.err
python bench.py --steps 10 --warmup 3 --dtype fp16 > gpurun_out/r2_bench_base_fp16.json 2> gpurun_out/r2_bench_base_fp16.err
ncu --set full --import-source on --clock-control none -k regex:pwdw -c 1 -o gpurun_out/r2_pwdw_base -f python tools/time_pwdw.py 16 256 256 96 256 1 3 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
