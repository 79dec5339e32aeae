TAG=r02d
cap() { # workload traversal kernel-regex name launch-skip
  ncu --set full --clock-control none --import-source on -k regex:$3 --launch-skip $5 --launch-count 1 -f \
    -o gpurun_out/${TAG}_$4 python bench.py --workload $1 --traversal $2 --no-extras --steps 20 --warmup 3 > gpurun_out/${TAG}_$4.log 2>&1
  ncu -i gpurun_out/${TAG}_$4.ncu-rep --page raw --csv > gpurun_out/${TAG}_$4_ncu_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$4.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_$4_sass.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_$4.ncu-rep
}
cap cfg2 esvo esvo_kernel esvo_bricks_f120 19
XN_ESVO_BRICKS=0 cap cfg2 esvo esvo_kernel esvo_nobricks_f120 19
