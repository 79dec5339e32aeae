# Round-end evidence: launch list, ncu --set full captures of the dominant kernels, bench lines.
# Numbers printed under ncu are never bench values; the bench lines come from the un-profiled runs.
# usage: bash tools/final_sweep.sh <tag>        (under gpurun, one GPU)
set -x
TAG=${1:-r02}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
# every launch of the default bench (cfg4 headline + per_config): cold-cache, serialised -> compare shares
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 6 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
cap() { # workload traversal kernel-regex name launch-skip
  ncu --set full --clock-control none --import-source on -k regex:$3 --launch-skip $5 --launch-count 1 -f \
    -o gpurun_out/${TAG}_$4 python bench.py --workload $1 --traversal $2 --no-extras --steps 20 --warmup 3 > gpurun_out/${TAG}_$4.log 2>&1
  ncu -i gpurun_out/${TAG}_$4.ncu-rep --page raw --csv > gpurun_out/${TAG}_$4_ncu_raw.csv 2>/dev/null
  # per-instruction execution and lane counts (tools/sass_segments.py reads these)
  ncu -i gpurun_out/${TAG}_$4.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_$4_sass.csv 2>/dev/null
  # the loops themselves, small enough to commit: instructions executed >= 2 % as often as the hottest one
  python tools/sass_excerpt.py gpurun_out/${TAG}_$4_sass.csv gpurun_out/${TAG}_$4_sass.txt
  # gpurun brings back at most 64 MiB: keep the report of the headline kernel only
  case $4 in dda_cfg4_f120) ;; *) rm -f gpurun_out/${TAG}_$4.ncu-rep ;; esac
}
# launch-skip = 3 warm-up launches + step index; step s of 20 renders script frame floor(s * 150 / 20):
# step 2 = frame 15 (outside, distance 2), step 16 = frame 120 (inside the volume)
cap cfg4 dda dda_ dda_cfg4_f120 19
cap cfg3 dda dda_ dda_cfg3_f120 19
cap cfg1 dda dda_ dda_cfg1 5
cap cfg2 dda dda_ dda_cfg2 19
cap cfg2 esvo esvo_kernel esvo_f120 19
cap cfg4e esvo esvo_kernel esvo_cfg4e_f120 19
cap cfg2 svo-df svo_df df_f120 19
cap cfg2 svo-rope svo_rope rope_f120 19
cap cfg3r svo-rope svo_rope rope_cfg3r_f120 19
cap cfg2 svo-naive svo_naive_kernel naive_f120 19
# memory and shared-memory race checks of the GPU suites (small inputs; the full-size tests are left out)
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ingest.py \
  tests/test_gpu_convert.py tests/test_gpu_cli.py -m gpu -q -x > gpurun_out/${TAG}_memcheck.log 2>&1; tail -3 gpurun_out/${TAG}_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "svo_matches and (esvo or svo-df) and blobby" > gpurun_out/${TAG}_racecheck.log 2>&1; tail -3 gpurun_out/${TAG}_racecheck.log
# un-profiled bench lines: the driver's command, longer runs of every workload, the reference arm
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for wl in cfg1 cfg2 cfg3 cfg3r cfg4 cfg4e cfg5; do
  python bench.py --workload $wl --steps 150 --warmup 5 --no-extras > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
done
for t in svo-rope svo-df svo-naive dda; do
  python bench.py --workload cfg2 --traversal $t --steps 150 --warmup 5 --no-extras > gpurun_out/${TAG}_bench_cfg2_$t.json 2> gpurun_out/${TAG}_bench_cfg2_$t.err
done
for f in gpurun_out/${TAG}_bench_*.json; do python -c "import json,sys;d=json.load(open('$f'));print('$f', d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'))"; done
