#!/bin/bash
# round-1c profiling pass (run under gpurun): launch list + full captures of the three hot kernels.
# The .ncu-rep files (20+ MB each with sources) are exported to csv/text on the box and dropped.
O=gpurun_out/r1c
mkdir -p $O
CMD="python bench.py --batch 20000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1"
export_rep() {   # $1 = report base name
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page details > $O/$1_details.txt 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/$1_source.csv.gz
  rm -f $O/$1.ncu-rep
}
$CMD > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hybrd_res -s 100 -c 1 -o $O/prof_res $CMD > $O/ncu2.log 2>&1; export_rep prof_res
ncu --set full --clock-control none --import-source on -k regex:integrate_worklist -s 100 -c 1 -o $O/prof_int $CMD > $O/ncu3.log 2>&1; export_rep prof_int
python tools/probe_traj.py 1048576 > $O/probe_traj.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:traj_kernel -s 2 -c 1 -o $O/prof_traj python tools/probe_traj.py 1048576 > $O/ncu4.log 2>&1; export_rep prof_traj
tail -6 $O/probe_traj.log; ls -la $O; du -sh gpurun_out
