python -m pytest tests/test_gpu_train.py tests/test_gpu_xrestormer.py -x -q 2>&1 | tail -4
for V in "PIR_SGEMM_SIMT=1" "PIR_X=1"; do
  echo "== $V" >> gpurun_out/r2_ab18.txt
  env $V python tools/bench_train.py --steps 5 --warmup 3 > "gpurun_out/tmp_train_$V.json" 2> gpurun_out/tmp_train.err
  python - "gpurun_out/tmp_train_$V.json" >> gpurun_out/r2_ab18.txt <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','unit','ms_forward','ms_backward')})
k=d.get('kernels') or {}
for n,v in sorted(k.items(), key=lambda t:-t[1].get('ms',0))[:8]: print('  ',n,v.get('launches'),v.get('ms'))
P
done
cat gpurun_out/r2_ab18.txt; tail -3 gpurun_out/tmp_train.err
