python -m pytest tests/test_gpu_xrestormer.py -x -q 2>&1 | tail -5
for V in "PIR_OCAB_BWD_SIMT=1" "PIR_X=1"; do
  echo "== $V" >> gpurun_out/r2_ab15.txt
  env $V python tools/bench_train.py --model xrestormer --steps 5 --warmup 3 > gpurun_out/tmp_xtrain_$V.json 2> gpurun_out/tmp_xtrain.err
  python - "gpurun_out/tmp_xtrain_$V.json" >> gpurun_out/r2_ab15.txt <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','unit')})
k=d.get('kernels') or {}
for n,v in sorted(k.items(), key=lambda t:-t[1].get('ms',0))[:8]: print('  ',n,v.get('launches'),v.get('ms'))
P
done
cat gpurun_out/r2_ab15.txt; tail -3 gpurun_out/tmp_xtrain.err
