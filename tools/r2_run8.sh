cd $GRAFT_REPO_ROOT
( time python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v5.json 2> gpurun_out/r2_bench_v5.err ) 2>&1 | grep real
tail -5 gpurun_out/r2_bench_v5.err
( time python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2_bench_v5_ref.json 2> gpurun_out/r2_bench_v5_ref.err ) 2>&1 | grep real
