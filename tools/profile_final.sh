#!/bin/bash
# final profiling pass of a round (run under gpurun): bench line, launch list, full captures of the
# three hot kernels; the .ncu-rep files are exported to csv/text on the box (they are 20+ MB each)
T=${1:-r1_final}; O=gpurun_out/$T; mkdir -p $O
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
CMD="python bench.py --batch 20000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1"
export_rep() {
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page details > $O/$1_details.txt 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/$1_source.csv.gz
  rm -f $O/$1.ncu-rep
}
$CMD > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hybrd_res -s 100 -c 1 -o $O/ncu_res $CMD > $O/ncu2.log 2>&1; export_rep ncu_res
ncu --set full --clock-control none --import-source on -k regex:hybrd_jac -s 30 -c 1 -o $O/ncu_jac $CMD > $O/ncu3.log 2>&1; export_rep ncu_jac
python tools/probe_traj.py 1048576 > $O/probe_traj.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:traj_kernel -s 2 -c 1 -o $O/ncu_traj python tools/probe_traj.py 1048576 > $O/ncu4.log 2>&1; export_rep ncu_traj
ls -la $O; du -sh gpurun_out
