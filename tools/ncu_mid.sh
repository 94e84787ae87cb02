CMD="python bench.py --events 3000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
$CMD > gpurun_out/r2m_plain.log 2>&1 || exit 1
for k in rs_scatter rs_hist k_links k_build_records k_fit_gather k_fit_solve k_pool_tile16; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -o gpurun_out/r2m_$k $CMD > gpurun_out/r2m_ncu_$k.log 2>&1
  tail -1 gpurun_out/r2m_ncu_$k.log
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2m_launches.csv python bench.py --events 20000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2m_launches.log 2>&1
ls -la gpurun_out/r2m_*
