cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "prompt" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/r2_bench_v8.json 2> gpurun_out/r2_bench_v8.err
