#!/bin/bash
# Per-source-line ncu views of kernels (development aid; one gpurun call):
#   gpurun --timeout 1500 -- 'bash tools/lines_round.sh'
# The reports stay on the box; only tools/ncu_lines.py's text output and the raw metric pages come back (gpurun_out/lines_*).
set -u
O=gpurun_out; mkdir -p $O; T=/tmp/lines; mkdir -p $T
cap() {  # cap <name> <kernel regex> <skip> <count> <prof_run args...>
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o $T/$name \
      python tools/prof_run.py "$@" > $O/lines_ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  ncu -i $T/$name.ncu-rep --page raw --csv > $O/lines_${name}_raw.csv 2>/dev/null
}
A2="--scene=heightfield --integrator=one_sample_mis --spp=32"
A5="--scene=instanced --integrator=mis --spp=4"
export TAKE_PROVISIONAL=0
cap refill2 k_extend_refill 0 2 $A2
python tools/ncu_lines.py $T/refill2.ncu-rep k_extend_refillILb0ELb0 0 60 > $O/lines_refill2_pass1.txt 2>&1
cap refill5 k_extend_refill 0 1 $A5
python tools/ncu_lines.py $T/refill5.ncu-rep k_extend_refillILb0ELb0 0 40 > $O/lines_refill5_pass1.txt 2>&1
ls -la $O | tail; du -sh $O
