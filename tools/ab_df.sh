set -x
run() { tag=$1; shift
  env "$@" python bench.py --workload cfg2 --traversal svo-df --steps 30 --warmup 3 --no-extras > gpurun_out/abt_$tag.json 2> gpurun_out/abt_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abt_$tag.json'));print('cfg2 svo-df $tag', d['value'], d['ms_per_step'])"; }
V=$PWD/xenodon_b200/variants/libxenodon_b200
run default XN_X=0
run dfbf XN_LIBRARY=${V}_dfbf.so
XN_LIBRARY=${V}_dfbf.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "svo-df or df" 2>&1 | tail -2
