# State check + first evidence of a round: GPU suite, default bench, reference arm, launch list, four ncu captures.
# usage: bash tools/state_sweep.sh <tag>        (under gpurun, one GPU)
set -x
TAG=${1:-r02a}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 6 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
cap() { # workload traversal kernel-regex name launch-skip
  ncu --set full --clock-control none --import-source on -k regex:$3 --launch-skip $5 --launch-count 1 -f \
    -o gpurun_out/${TAG}_$4 python bench.py --workload $1 --traversal $2 --no-extras --steps 20 --warmup 3 > gpurun_out/${TAG}_$4.log 2>&1
  ncu -i gpurun_out/${TAG}_$4.ncu-rep --page raw --csv > gpurun_out/${TAG}_$4_ncu_raw.csv 2>/dev/null
}
cap cfg4 dda dda_ dda_cfg4_f120 19
cap cfg2 esvo esvo_kernel esvo_f15 5
cap cfg2 esvo esvo_kernel esvo_f120 19
cap cfg2 svo-df svo_df dfr_f15 5
cap cfg2 svo-rope svo_rope rope_f15 5
for t in esvo svo-rope svo-df svo-naive; do
  python bench.py --workload cfg2 --traversal $t --steps 150 --warmup 5 --no-extras > gpurun_out/${TAG}_bench_cfg2_$t.json 2> gpurun_out/${TAG}_bench_cfg2_$t.err
done
for f in gpurun_out/${TAG}_bench_*.json; do python -c "import json,sys;d=json.load(open('$f'));print('$f', d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'))"; done
