python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gputests_v9.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v9.json 2> gpurun_out/r2_bench_v9.err
tail -2 gpurun_out/r2_gputests_v9.txt; head -c 600 gpurun_out/r2_bench_v9.json; tail -3 gpurun_out/r2_bench_v9.err
