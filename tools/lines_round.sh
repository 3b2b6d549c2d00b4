#!/bin/bash
# Per-source-line ncu views (tools/ncu_lines.py: stall samples and active lanes per source line) of the first bounce pass of
# config 2 -- with one ray per thread (TAKE_REFILL=0) and with lane refill -- and of the multi-sample shade kernel on the
# config-4 scene.  One gpurun call:  gpurun --timeout 1500 -- 'bash tools/lines_round.sh'
# The reports stay on the box; the text views come back as gpurun_out/lines_*.txt (copied to profiles/<tag>_lines_*.txt).
set -u
O=gpurun_out; mkdir -p $O; T=/tmp/lines; mkdir -p $T
cap() {  # cap <name> <kernel regex> <skip> <count> <prof_run args...>
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  TAKE_PROVISIONAL=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o $T/$name \
      python tools/prof_run.py "$@" > $O/lines_ncu_$name.log 2>&1
  echo "ncu $name exit $?"
}
A2="--scene=heightfield --integrator=one_sample_mis --spp=32"
A4="--scene=multi_light --integrator=mis --spp=16"
TAKE_REFILL=0 cap one_ray_per_thread 'k_extend$' 1 1 $A2          # (k_extend launches are passes 1..6 here: this captures pass 2)
python tools/ncu_lines.py $T/one_ray_per_thread.ncu-rep k_extendILb0ELb1ELb0 0 40 > $O/lines_extend_pass2_one_ray_per_thread.txt 2>&1
cap refill k_extend_refill 0 1 $A2
python tools/ncu_lines.py $T/refill.ncu-rep k_extend_refillILb0ELb0 0 40 > $O/lines_extend_pass1_lane_refill.txt 2>&1
cap shade4 k_shade 1 1 $A4
python tools/ncu_lines.py $T/shade4.ncu-rep k_shadeILi0ELb0ELb0 0 45 > $O/lines_shade_mis_config4_pass1.txt 2>&1
head -3 $O/lines_*.txt
cap primary k_extend_primary 1 1 $A2                              # the heavy half of the frame (wave 1)
python tools/ncu_lines.py $T/primary.ncu-rep k_extend_primaryILb0ELb0 0 40 > $O/lines_extend_primary.txt 2>&1
