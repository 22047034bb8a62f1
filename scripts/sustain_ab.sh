for v in ${VARIANTS:-"" n0 n2 "" n0 n2}; do
  if [ -n "$v" ]; then export SOCIALWAYS_B200_LIB=$PWD/socialways_b200/build/ab/libsw_$v.so; else unset SOCIALWAYS_B200_LIB; fi
  [ "$v" = "-" ] && v=""; if [ -n "$v" ]; then export SOCIALWAYS_B200_LIB=$PWD/socialways_b200/build/ab/libsw_$v.so; else unset SOCIALWAYS_B200_LIB; fi; echo "== ${v:-default}"; python scripts/pair_sustain.py short 2>&1 | tail -1 | awk '{n=0; s=0; for(i=NF-19;i<=NF;i++){s+=$i;n++}; printf "burst %s %s %s  sustained(mean of last 20) %.3f\n",$4,$5,$6,s/n}'
  sleep 2
done
