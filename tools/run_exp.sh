for r in 0 1 2; do
  PG_EXP_XROW_REPEAT=$r timeout 300 python tools/prof_reml.py 10000 50000 10
  PG_EXP_XROW_REPEAT=$r timeout 300 python tools/prof_reml.py 449 100000 6
done
