#!/bin/bash
# full ncu capture of one hybrd_res_kernel launch (exported to csv/text on the box)
O=gpurun_out/${1:-prof}; mkdir -p $O
CMD="python bench.py --batch 20000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$CMD > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:${2:-hybrd_res} -s ${3:-100} -c 1 -o $O/prof $CMD > $O/ncu.log 2>&1
ncu -i $O/prof.ncu-rep --page raw --csv > $O/prof_raw.csv 2>/dev/null
ncu -i $O/prof.ncu-rep --page details > $O/prof_details.txt 2>/dev/null
ncu -i $O/prof.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/prof_source.csv.gz
rm -f $O/prof.ncu-rep; ls -la $O
