# one-addition-per-binade jump loop, auto brick size, always-fetch trips; step counters; ncu capture
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "dda or skip or grid or full" > gpurun_out/abn_pytest.log 2>&1; tail -3 gpurun_out/abn_pytest.log
run() { # tag env...
  tag=$1; shift
  env "$@" python bench.py --workload $WL $TRAV --steps 30 --warmup 3 --no-extras > gpurun_out/abn_${WL}_$tag.json 2> gpurun_out/abn_${WL}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abn_${WL}_$tag.json'));r=d['roofline'];print('$WL $tag', d['value'], d['ms_per_step'], 'steps', r['steps_per_launch'], 'bytes', r['algorithmic_bytes_per_launch'])"
}
TRAV=
for WL in cfg4 cfg3 cfg1; do
  run default XN_X=0
  run shift3 XN_SKIP_SHIFT=3
  for v in tripfetch dbg1 dbg2 dbg3 dbg4; do run $v XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_$v.so; done
done
ncu --set full --clock-control none --import-source on -k regex:dda_ --launch-skip 19 --launch-count 1 -f \
    -o gpurun_out/r02c_dda_cfg4_f120 python bench.py --workload cfg4 --traversal dda --no-extras --steps 20 --warmup 3 > gpurun_out/r02c_dda_cfg4_f120.log 2>&1
ncu -i gpurun_out/r02c_dda_cfg4_f120.ncu-rep --page raw --csv > gpurun_out/r02c_dda_cfg4_f120_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02c_dda_cfg4_f120.ncu-rep --page source --csv --print-source sass > gpurun_out/r02c_dda_cfg4_f120_sass.csv 2>/dev/null
