cd $GRAFT_REPO_ROOT
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_v5_2gpu.json 2> gpurun_out/r2_bench_v5_2gpu.err ) 2>&1 | grep real
tail -5 gpurun_out/r2_bench_v5_2gpu.err
