cd $GRAFT_REPO_ROOT
( time timeout 90 python tools/bench_reference_eager.py > gpurun_out/r2_reference_eager_b200.json 2> gpurun_out/r2_reference_eager_b200.err ) 2>&1 | grep real
( time timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v14.json 2> gpurun_out/r2_bench_v14.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r2_reference_eager_b200.json
tail -3 gpurun_out/r2_bench_v14.err
