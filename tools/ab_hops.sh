# A/B: extra look-up + jump rounds before a trip when most lanes of the warp can jump (XN_SKIP_HOPS /
# XN_SKIP_HOP_LANES), 256- vs 128-bit record loads (svo_rope32, svo_df fast); GPU suite on the default first
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/abq_pytest.log 2>&1; tail -3 gpurun_out/abq_pytest.log
run() { # tag env...
  tag=$1; shift
  env "$@" python bench.py --workload $WL $TRAV --steps 30 --warmup 3 --no-extras > gpurun_out/abq_${WL}_$tag.json 2> gpurun_out/abq_${WL}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abq_${WL}_$tag.json'));print('$WL $TRAV $tag', d['value'], d['ms_per_step'])"
}
V=$PWD/xenodon_b200/variants/libxenodon_b200
TRAV=
for WL in cfg4 cfg3; do
  run default XN_X=0
  for v in hop2_16 hop2_24 hop2_28 hop3_20 hop3_26 hop4_24; do run $v XN_LIBRARY=${V}_$v.so; done
done
WL=cfg1; run default XN_X=0; run hop2_24 XN_LIBRARY=${V}_hop2_24.so
WL=cfg2
for TRAV in "--traversal svo-rope" "--traversal svo-df"; do
  run default$(echo $TRAV | tr -d ' -') XN_X=0
  run ldg128$(echo $TRAV | tr -d ' -') XN_LIBRARY=${V}_ldg128.so
done
WL=cfg3r; TRAV=
run default XN_X=0
run ldg128 XN_LIBRARY=${V}_ldg128.so
