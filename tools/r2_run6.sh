cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v3_bf16.json 2> gpurun_out/r2_bench_v3_bf16.err
PROMPTIR_B200_SPLITQKV=0 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v3_bf16_nosplit.json 2> gpurun_out/r2_bench_v3_bf16_nosplit.err
python bench.py --steps 10 --warmup 3 --dtype fp16 > gpurun_out/r2_bench_v3_fp16.json 2> gpurun_out/r2_bench_v3_fp16.err
