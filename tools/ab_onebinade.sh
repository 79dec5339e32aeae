# A/B: jumps capped at the binade top of every axis (XN_SKIP_ONE_BINADE) vs continuing across binades
set -x
python -m pytest tests -m gpu -x -q -k "dda or grid or fullsize or skip" > gpurun_out/abs_pytest.log 2>&1; tail -3 gpurun_out/abs_pytest.log
run() { # tag env...
  tag=$1; shift
  env "$@" python bench.py --workload $WL $TRAV --steps 30 --warmup 3 --no-extras > gpurun_out/abs_${WL}_$tag.json 2> gpurun_out/abs_${WL}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abs_${WL}_$tag.json'));print('$WL $TRAV $tag', d['value'], d['ms_per_step'])"
}
V=$PWD/xenodon_b200/variants/libxenodon_b200
TRAV=
for WL in cfg4 cfg3 cfg1 cfg5; do
  run ob1 XN_X=0
  run ob0 XN_LIBRARY=${V}_ob0.so
done
