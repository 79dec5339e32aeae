# A/B of the closed-form DDA jump (XN_SKIP_BARE=2) against counted bare trips (=1) and of its threshold
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "dda or skip or grid or full" > gpurun_out/abj_pytest.log 2>&1; tail -3 gpurun_out/abj_pytest.log
for wl in cfg4 cfg3 cfg1; do
  python bench.py --workload $wl --steps 30 --warmup 3 --no-extras > gpurun_out/abj_${wl}_default.json 2> gpurun_out/abj_${wl}_default.err
  python -c "import json;d=json.load(open('gpurun_out/abj_${wl}_default.json'));print('$wl default', d['value'], d['ms_per_step'], d.get('parity_check'))"
  for v in bare1 j12 j48; do
    XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_$v.so python bench.py --workload $wl --steps 30 --warmup 3 --no-extras > gpurun_out/abj_${wl}_$v.json 2> gpurun_out/abj_${wl}_$v.err
    python -c "import json;d=json.load(open('gpurun_out/abj_${wl}_$v.json'));print('$wl $v', d['value'], d['ms_per_step'], d.get('parity_check'))"
  done
done
