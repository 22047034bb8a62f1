# Burst vs sustained (power-capped) time of the pair decode for library variants built by scripts/build_variant.py:
#   VARIANTS="- x2 - x2" bash scripts/sustain_ab.sh      ("-" = the regular library)
# Each variant: 40 back-to-back launches on the bench workload (scripts/pair_sustain.py short); prints the first launches after
# warm-up (burst) and the mean of the last 20 (the board's power cap has engaged by then).
for v in ${VARIANTS:-"-"}; do
  [ "$v" = "-" ] && v=""
  if [ -n "$v" ]; then export SOCIALWAYS_B200_LIB=$PWD/socialways_b200/build/ab/libsw_$v.so; else unset SOCIALWAYS_B200_LIB; fi
  echo "== ${v:-default}"
  python scripts/pair_sustain.py short 2>&1 | tail -1 | awk '{n=0; s=0; for(i=NF-19;i<=NF;i++){s+=$i;n++}; printf "burst %s %s %s  sustained(mean of last 20) %.3f\n",$4,$5,$6,s/n}'
  sleep 2
done
