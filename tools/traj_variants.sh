#!/bin/bash
# RK4 trajectory kernel: occupancy variants (threads per CTA x register cap), measured with tools/probe_traj.py
O=gpurun_out/trajvar; mkdir -p $O
for v in "128 1" "128 3" "64 5" "64 6" "96 4" "128 4"; do
  set -- $v
  SOCP_NVCC_EXTRA="-DSOCP_TRAJ_THREADS=$1 -DSOCP_TRAJ_MINB=$2" python -m socp_b200.build --force > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  python tools/probe_traj.py 1048576 > $O/t$1_b$2.log 2>&1
  echo "threads $1 minblocks $2: $(grep goddard $O/t$1_b$2.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("goddard %.3f ms  %.2f Gsteps/s  frac %.3f"%(d["ms"],d["rk4_steps_per_s"]/1e9,d["frac_of_peak"]))')  $(grep vtolUAV $O/t$1_b$2.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("vtol frac %.3f"%d["frac_of_peak"])') $(grep interceptor $O/t$1_b$2.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("icp frac %.3f"%d["frac_of_peak"])') $(grep covid $O/t$1_b$2.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("covid frac %.3f"%d["frac_of_peak"])') $(grep doubleInt $O/t$1_b$2.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("di frac %.3f"%d["frac_of_peak"])')"
done
