# nstr = 24 / 32 (rolled instantiation, per-solve state in local memory): resident threads per SM vs L2 capacity
for n in 24 32; do
for cfg in 128:0 64:1 32:1 32:2 64:2 128:1; do
  b=${cfg%%:*}; c=${cfg##*:}
  echo -n "nstr $n block $b ctas/SM $c: "
  NSTRS=$n B200DISORT_BLOCK=$b B200DISORT_CTAS_PER_SM=$c python profiles/scripts/sweep_nstr.py 2>&1 | tail -1
done
done
