#!/bin/bash
# compute-sanitizer over one small stream per pooling-kernel instantiation (memcheck + racecheck; synccheck and
# initcheck on the dense case).  Run on a B200 box:  bash tools/sanitize.sh gpurun_out/sanitize
# Logs: <out>_<tool>_<case>.log ; a summary line per run goes to <out>_summary.txt
out=${1:-gpurun_out/sanitize}
cs=/usr/local/cuda/bin/compute-sanitizer
: > ${out}_summary.txt
for case in dense sparse bits tile1 tall aliased exact; do
  python tools/sanitize_cases.py $case --check > ${out}_plain_${case}.log 2>&1
  echo "plain $case exit=$? $(grep sanitize_case ${out}_plain_${case}.log)" >> ${out}_summary.txt
  for tool in memcheck racecheck; do
    timeout 900 $cs --tool $tool --print-limit 20 python tools/sanitize_cases.py $case > ${out}_${tool}_${case}.log 2>&1
    echo "$tool $case exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' ${out}_${tool}_${case}.log | tr '\n' ' ') $(grep sanitize_case ${out}_${tool}_${case}.log)" >> ${out}_summary.txt
  done
done
for tool in synccheck initcheck; do
  timeout 900 $cs --tool $tool --print-limit 20 python tools/sanitize_cases.py dense > ${out}_${tool}_dense.log 2>&1
  echo "$tool dense exit=$? $(grep -E 'ERROR SUMMARY' ${out}_${tool}_dense.log | tr '\n' ' ')" >> ${out}_summary.txt
done
cat ${out}_summary.txt
