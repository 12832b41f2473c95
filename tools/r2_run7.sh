cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/r2_bench_v7.json 2> gpurun_out/r2_bench_v7.err
tail -3 gpurun_out/r2_bench_v7.err
