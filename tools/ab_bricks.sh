# A/B: ESVO leaf bricks (run-time XN_ESVO_BRICKS), lanes needed for the closed form (XN_ESVO_BRICK_LANES)
set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "esvo or svo_matches or single_leaf or anisotropic" > gpurun_out/abr_pytest.log 2>&1; tail -3 gpurun_out/abr_pytest.log
run() { # tag env...
  tag=$1; shift
  env "$@" python bench.py --workload $WL $TRAV --steps 30 --warmup 3 --no-extras > gpurun_out/abr_${WL}_$tag.json 2> gpurun_out/abr_${WL}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abr_${WL}_$tag.json'));print('$WL $TRAV $tag', d['value'], d['ms_per_step'])"
}
V=$PWD/xenodon_b200/variants/libxenodon_b200
TRAV="--traversal esvo"
for WL in cfg2 cfg4e; do
  run bricks0 XN_ESVO_BRICKS=0
  run bricks1 XN_ESVO_BRICKS=1
  for v in bl6 bl16 bl20 bl24 bl12mb5; do run $v XN_ESVO_BRICKS=1 XN_LIBRARY=${V}_$v.so; done
done
