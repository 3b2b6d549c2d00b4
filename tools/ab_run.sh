#!/bin/bash
# A/B of kernel variants on a GPU box (development aid): parity suite first (unless SKIP_TESTS=1), then tools/tune.py per scene.
# usage: V="tag:ENV=VAL,LIB=libtake_gpu_x.so ..." bash tools/ab_run.sh
O=gpurun_out; mkdir -p $O
if [ "${SKIP_TESTS:-0}" != "1" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/ab_pytest.log; fi
V=${V:-"new:"}
CFGS=${CFGS:-"heightfield,one_sample_mis,32 multi_light,mis,16 cornell,mis,64 ibl,one_sample_mis,32"}
for cfg in $CFGS; do
  IFS=, read sc integ spp <<< "$cfg"
  echo "== $sc $integ spp=$spp" | tee -a $O/ab.log
  TUNE_SCENE=$sc TUNE_INTEGRATOR=$integ TUNE_SPP=$spp timeout 600 python tools/tune.py $V 2>&1 | tee -a $O/ab.log
done
