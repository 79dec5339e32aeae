# jump to the end of the promise + hop rounds; persistent ray pool for the ESVO
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "dda or skip or grid or full or pool" > gpurun_out/abl_pytest.log 2>&1; tail -3 gpurun_out/abl_pytest.log
run() { # tag env...
  tag=$1; shift
  env "$@" python bench.py --workload $WL $TRAV --steps 30 --warmup 3 --no-extras > gpurun_out/abl_${WL}_$tag.json 2> gpurun_out/abl_${WL}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abl_${WL}_$tag.json'));r=d['roofline'];print('$WL $tag', d['value'], d['ms_per_step'])"
}
TRAV=
for WL in cfg4 cfg3 cfg1; do
  run default XN_X=0
  run shift2 XN_SKIP_SHIFT=2
  for v in full0 hops1 hops3 minb3 minb5; do run $v XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_$v.so; done
  run hops3shift2 XN_SKIP_SHIFT=2 XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_hops3.so
done
WL=cfg2
run static XN_X=0
run pool24 XN_RAY_POOL=1
run pool16 XN_RAY_POOL=1 XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_pool16.so
run pool30 XN_RAY_POOL=1 XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_pool30.so
