#!/bin/bash
# round-1d profiling pass (run under gpurun): full captures of hybrd_res / hybrd_jac after the
# sequential-phase rewrite; reports exported to csv/text on the box (the .ncu-rep files are 20+ MB).
O=gpurun_out/r1d
mkdir -p $O
CMD="python bench.py --batch 20000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1"
export_rep() {
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page details > $O/$1_details.txt 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/$1_source.csv.gz
  rm -f $O/$1.ncu-rep
}
$CMD > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hybrd_res -s 100 -c 1 -o $O/prof_res $CMD > $O/ncu2.log 2>&1; export_rep prof_res
ncu --set full --clock-control none --import-source on -k regex:hybrd_jac -s 30 -c 1 -o $O/prof_jac $CMD > $O/ncu3.log 2>&1; export_rep prof_jac
ls -la $O
