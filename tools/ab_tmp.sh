python -m pytest tests/test_gpu_xrestormer.py -x -q 2>&1 | tail -4
echo "== ocab_bwd 3 CTAs/SM" >> gpurun_out/r2_ab20.txt
python tools/bench_train.py --model xrestormer --steps 5 --warmup 3 > "gpurun_out/tmp_xtrain.json" 2> gpurun_out/tmp_xtrain.err
python - "gpurun_out/tmp_xtrain.json" >> gpurun_out/r2_ab20.txt <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','unit','ms_forward','ms_backward')})
k=d.get('kernels') or {}
for n,v in sorted(k.items(), key=lambda t:-t[1].get('ms',0))[:10]: print('  ',n,v.get('launches'),v.get('ms'))
P
cat gpurun_out/r2_ab20.txt; tail -3 gpurun_out/tmp_xtrain.err
