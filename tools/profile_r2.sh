#!/bin/bash
# Final profiling pass of round 2 (run under gpurun, ONE GPU): the bench line of both arms, the ncu launch list and one
# `--set full` capture of every kernel of the solver round + the RK4 trajectory kernel; the .ncu-rep files are
# exported to csv / text on the box (they are 20+ MB each).  Every ncu run follows a plain run of the same command.
T=${1:-r2_final}; O=gpurun_out/$T; mkdir -p $O
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
CMD="python bench.py --batch 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-stage2 --no-profile-pass --e2e-steps 1"
export_rep() {
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page details > $O/$1_details.txt 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/$1_source.csv.gz
  rm -f $O/$1.ncu-rep
}
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 480 --csv --log-file $O/launches.csv $CMD > $O/ncu1.log 2>&1
for K in chain:hybrd_chain:100 qpass:hybrd_qpass:100 jac:hybrd_jac:30 int:integrate_worklist:100 asm:assemble_kernel:31; do
  IFS=: read name pat skip <<< "$K"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -o $O/ncu_$name $CMD > $O/ncu_$name.log 2>&1; export_rep ncu_$name
done
python tools/probe_traj.py 1048576 > $O/probe_traj.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:traj_kernel -s 2 -c 1 -o $O/ncu_traj python tools/probe_traj.py 1048576 0 > $O/ncu_traj.log 2>&1; export_rep ncu_traj
ls -la $O; du -sh gpurun_out
