for v in base pf1 pf2 pf3 sr8 sr2 base pf1 pf2 pf3; do
  if [ $v = base ]; then python profiles/quick_bench.py $v 3; else LD_LIBRARY_PATH=build/variants/$v python profiles/quick_bench.py $v 3; fi
done
