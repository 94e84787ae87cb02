#!/bin/bash
# Memory-safety evidence for the kernels, on a B200 box:  bash tools/sanitize.sh gpurun_out/selfcheck
#  1. compute-sanitizer memcheck + racecheck over one small stream per pooling-kernel instantiation
#     (tools/sanitize_cases.py).  On the pool this was developed on the tool answers that it is closed
#     (profiles/r2a_compute_sanitizer_closed.txt); the attempt and its answer are logged.
#  2. The same cases, plus long dense streams, through the SELF-CHECKING build libfarms_b200_checked.so
#     (make checked: every staged-slot / run-table / history-link / output index is bounds-checked by the kernel
#     itself, csrc/farms_dev.cuh FARMS_CHK) with the oracle comparison on.
out=${1:-gpurun_out/selfcheck}
cs=/usr/local/cuda/bin/compute-sanitizer
: > ${out}_summary.txt
timeout 300 $cs --tool memcheck --print-limit 20 python tools/sanitize_cases.py sparse > ${out}_memcheck_sparse.log 2>&1
echo "compute-sanitizer memcheck sparse: exit=$? $(head -c 300 ${out}_memcheck_sparse.log | tr '\n' ' ')" >> ${out}_summary.txt
export FARMS_B200_LIB=$PWD/aperture-robust-multiscale-optical-flow_b200/libfarms_b200_checked.so
for case in ${CASES:-dense sparse bits tile1 tile warp tile16 tile16x4 tile16c tall aliased aliased_c exact serial long4 long3 long4_c}; do
  python tools/sanitize_cases.py $case --check > ${out}_checked_${case}.log 2>&1
  echo "checked-build $case exit=$? $(grep -E 'sanitize_case|Error|error' ${out}_checked_${case}.log | tail -2 | tr '\n' ' ')" >> ${out}_summary.txt
done
cat ${out}_summary.txt
