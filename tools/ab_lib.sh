# A/B of library builds: bash tools/ab_lib.sh tag1 tag2 ...   (libfarms_b200_<tag>.so next to the product library)
B="python bench.py --events 60000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
P=$PWD/aperture-robust-multiscale-optical-flow_b200
for tag in base "$@"; do
  lib=$P/libfarms_b200.so; [ $tag != base ] && lib=$P/libfarms_b200_$tag.so
  FARMS_B200_LIB=$lib $B > gpurun_out/ab_lib.json 2> gpurun_out/ab_lib.err
  python -c "
import json;d=json.load(open('gpurun_out/ab_lib.json'));print('$tag', round(d['value'],1), {k:round(v,1) for k,v in d['stages_ms_per_step'].items()})"
done
