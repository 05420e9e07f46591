#!/bin/bash
# usage: tools/run_variants2.sh v1 v2 ...   (libraries tools/variants/libnnmd_<v>.so; "prod" = the product library)
for v in "$@"; do
  if [ "$v" = prod ]; then unset NNMD_B200_LIB; else export NNMD_B200_LIB=$PWD/tools/variants/libnnmd_$v.so; fi
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - "$v" <<PY
import json,sys
v=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/var_{v}.json") if l.startswith("{")][-1])
    s=d["roofline"]["stage_ms_per_step"]
    print(f"{v:14s} step {d['ms_per_step']:.2f} ms  angular {s['angular']:.2f} gather {s['gather']:.2f} nlist {s['nlist']:.2f} radial {s['radial']:.2f} mlp {s['mlp']:.2f} contract {s['contract']:.2f} norm {s['normalize']:.2f} | E {d['checksum']['energy']} sumF {d['checksum']['sum_abs_F']:.6f}")
except Exception as e:
    print(v, "ERR", e, open(f"gpurun_out/var_{v}.err").read()[-800:])
PY
done
