# plen: jump right after a fetch trip; look-up spacing; thresholds; brick size
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "dda or skip or grid or full" > gpurun_out/abm_pytest.log 2>&1; tail -3 gpurun_out/abm_pytest.log
run() { # tag env...
  tag=$1; shift
  env "$@" python bench.py --workload $WL $TRAV --steps 30 --warmup 3 --no-extras > gpurun_out/abm_${WL}_$tag.json 2> gpurun_out/abm_${WL}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abm_${WL}_$tag.json'));r=d['roofline'];print('$WL $tag', d['value'], d['ms_per_step'])"
}
TRAV=
for WL in cfg4 cfg3 cfg1; do
  run default XN_X=0
  run shift2 XN_SKIP_SHIFT=2
  for v in plen0 look1 j3 j12; do run $v XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_$v.so; done
  run look1shift2 XN_SKIP_SHIFT=2 XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_look1.so
done
