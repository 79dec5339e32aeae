# Round-end evidence: launch list, ncu --set full captures of the dominant kernels, bench lines.
# Numbers printed under ncu are never bench values; the bench lines come from the un-profiled runs.
set -x
TAG=${1:-r01b}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
cap() { # workload traversal kernel-regex name
  ncu --set full --clock-control none --import-source on -k regex:$3 --launch-skip 8 --launch-count 1 -f \
    -o gpurun_out/${TAG}_$4 python bench.py --workload $1 --traversal $2 --no-extras --steps 10 --warmup 3 > gpurun_out/${TAG}_$4.log 2>&1
  ncu -i gpurun_out/${TAG}_$4.ncu-rep --page raw --csv > gpurun_out/${TAG}_$4_ncu_raw.csv 2>/dev/null
}
cap cfg2 esvo esvo_kernel esvo
cap cfg2 svo-rope svo_rope_kernel rope
cap cfg2 svo-naive svo_naive_kernel naive
cap cfg2 svo-df svo_df_kernel df
cap cfg1 dda dda dda_cfg1
cap cfg3 dda dda dda_cfg3
cap cfg4 dda dda dda_cfg4
python bench.py > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
for wl in cfg1 cfg3 cfg3r cfg4 cfg4e cfg5; do
  python bench.py --workload $wl --no-extras > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
done
python bench.py --impl reference --steps 12 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for f in gpurun_out/${TAG}_bench_*.json; do python -c "import json,sys;d=json.load(open('$f'));print('$f', d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'))"; done
