#!/bin/bash
# A/B runs of bench.py (1 timed step) under environment overrides; prints one summary line per variant
O=gpurun_out/ab; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 1 --warmup 1 --e2e-steps 1 > $O/$name.json 2> $O/$name.err; python - $O/$name.json $name <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); k=d['roofline']['kernels']
print(sys.argv[2], "solves/s %.0f ms/step %.0f rounds %d | res %.0f jac %.0f int %.0f asm %.0f ms | conv %.3f nfev %.0f"%(d['value'],d['ms_per_step'],d['solver_rounds_per_step'],k['hybrd_res_kernel']['ms'],k['hybrd_jac_kernel']['ms'],k['integrate_worklist<goddard>']['ms'],k['assemble_kernel<goddard>']['ms'],d['converged_fraction'],d['mean_nfev']))
PY
}
for v in "$@"; do
  case $v in
    base) run base X=1;;
    clocks) run clocks SOCP_PHASE_CLOCKS=1; grep -h 'phase clocks' $O/clocks.err | tail -1;;
    seed*) run $v SOCP_BENCH_SEED_OFFSET=${v#seed};;
  esac
done
