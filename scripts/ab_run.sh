# A/B of library variants built by scripts/build_variant.py:  bash scripts/ab_run.sh "" cell2 ...
for v in "$@"; do
  if [ -n "$v" ]; then export SOCIALWAYS_B200_LIB=$PWD/socialways_b200/build/ab/libsw_$v.so; else unset SOCIALWAYS_B200_LIB; fi
  echo "== variant: ${v:-base}"; timeout 200 python scripts/pair_check.py bench 2>&1 | grep -E "pair|Error|error"
done
