# closed-form jump: threshold, brick size, radius cap; where the steps go (debug counters); one ncu capture
set -x
run() { # tag env... -- args
  tag=$1; shift
  env "$@" python bench.py --workload $WL --steps 30 --warmup 3 --no-extras > gpurun_out/abk_${WL}_$tag.json 2> gpurun_out/abk_${WL}_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/abk_${WL}_$tag.json'));r=d['roofline'];print('$WL $tag', d['value'], d['ms_per_step'], 'steps', r['steps_per_launch'], 'bytes', r['algorithmic_bytes_per_launch'])"
}
for WL in cfg4 cfg3; do
  run default XN_X=0
  for v in j6 j8 j12; do run $v XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_$v.so; done
  for c in 64 128 255; do run cap$c XN_SKIP_CAP=$c; done
  for s in 2 4; do run shift$s XN_SKIP_SHIFT=$s; done
  run shift4cap128 XN_SKIP_SHIFT=4 XN_SKIP_CAP=128
  for v in dbg1 dbg2 dbg3 dbg4; do run $v XN_LIBRARY=$PWD/xenodon_b200/variants/libxenodon_b200_$v.so; done
done
WL=cfg1; run default XN_X=0; run cap128 XN_SKIP_CAP=128; run shift2 XN_SKIP_SHIFT=2
ncu --set full --clock-control none --import-source on -k regex:dda_ --launch-skip 19 --launch-count 1 -f \
    -o gpurun_out/r02b_dda_cfg4_f120 python bench.py --workload cfg4 --traversal dda --no-extras --steps 20 --warmup 3 > gpurun_out/r02b_dda_cfg4_f120.log 2>&1
ncu -i gpurun_out/r02b_dda_cfg4_f120.ncu-rep --page raw --csv > gpurun_out/r02b_dda_cfg4_f120_ncu_raw.csv 2>/dev/null
