# in-step A/B of library variants (scripts/build_variant.py): bash scripts/ab_bench.sh "" A B ...   -> decode kernel ms inside the bench step
for v in "$@"; do
  if [ -n "$v" ]; then export SOCIALWAYS_B200_LIB=$PWD/socialways_b200/build/ab/libsw_$v.so; else unset SOCIALWAYS_B200_LIB; fi
  python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;l=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('variant ${v:-head}', round(l['value']/1e6,1), 'M traj/s, step', round(l['ms_per_step'],3), 'decode', round(l['roofline']['kernel_ms'],3))"
done
