cd $GRAFT_REPO_ROOT
mkdir -p /tmp/prof
for cfg in "fp16 43 C96_fp16" "bf16 43 C96_bf16" "fp16 0 C48_fp16"; do
  set -- $cfg
  ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/prof/r2_block_$3 -f python tools/profile_block.py 16 256 256 $1 $2 > gpurun_out/ncu_block_$3.log 2>&1
  ncu -i /tmp/prof/r2_block_$3.ncu-rep --page raw --csv > gpurun_out/r2_block_$3.raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/prof/r2_block_$3.ncu-rep > gpurun_out/r2_block_$3.md 2>/dev/null
done
ncu -i /tmp/prof/r2_block_C96_fp16.ncu-rep --page source --csv > gpurun_out/r2_block_C96_fp16.source.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-configs > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_bench_under_ncu.err
du -sh gpurun_out
