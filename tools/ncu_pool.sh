# ncu --set full of the first pooling pass of the timed step (B200_PROFILING.md recipe: plain run first)
CMD="python bench.py --events 3000000 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-extras"
$CMD > gpurun_out/r2p_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_pool_tile16 -s 4 -c 1 -o gpurun_out/r2p_pooltile16 $CMD > gpurun_out/r2p_ncu.log 2>&1
tail -2 gpurun_out/r2p_ncu.log
for k in k_links k_build_records; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2p_$k $CMD > gpurun_out/r2p_ncu_$k.log 2>&1
  tail -1 gpurun_out/r2p_ncu_$k.log
done
