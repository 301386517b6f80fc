#!/usr/bin/env bash
# Same-box A/B of two library builds: diffusionvid_b200/_C/libdvid_b200_prev.so (built by hand from an older
# conv_gemm.cu) against the current library.  usage: tools/ab.sh [bench.py args...]
for l in prev new prev new; do
  if [ $l = prev ]; then export DVID_LIB_PATH=$PWD/diffusionvid_b200/_C/libdvid_b200_prev.so; else unset DVID_LIB_PATH; fi
  echo -n "lib $l: "
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1))"
done
