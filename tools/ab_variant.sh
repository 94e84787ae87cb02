# A/B of a pooling variant against the library's default on a B200 box, with its parity tests and self-check cases:
#   bash tools/ab_variant.sh            (the run recorded in profiles/r4_xcull_ab.txt: pool_variant "tile16c")
B="python bench.py --events 60000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
for v in "" tile16c; do
  $B --pool-variant "$v" > gpurun_out/r4b_ab_$v.json 2> gpurun_out/r4b_ab_$v.err
  python -c "
import json;d=json.loads([l for l in open('gpurun_out/r4b_ab_$v.json') if l.startswith('{')][-1]);print('variant [$v]', round(d['value'],1), {k:round(v,1) for k,v in d['stages_ms_per_step'].items()}, d['pool_candidates_per_step'], d['pool_paths_events_per_step'])"
done
timeout 420 python -m pytest tests/test_gpu_parity.py -q -k tile16c 2>&1 | tail -8 | tee gpurun_out/r4b_pytest.log
CASES="tile16c aliased_c long4_c" timeout 300 bash tools/sanitize.sh gpurun_out/r4b_selfcheck 2>&1 | tail -6
