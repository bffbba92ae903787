#!/bin/bash
# BASELINE.json configs 3, 4 and 5 on N GPUs of one box (one rank per GPU, NCCL).  Usage: tools/run_configs.sh N [what]
# what: c3 | c4 | c5 | all.  JSON lines land in gpurun_out/r02_*.json
N=${1:-8}; WHAT=${2:-all}; OUT=gpurun_out; mkdir -p $OUT
run() { # name, bench args...
  name=$1; shift
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" > $OUT/$name.json 2> $OUT/$name.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" > $OUT/$name.json 2> $OUT/$name.err; fi
  tail -c 400 $OUT/$name.json | head -c 400 > /dev/null
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "bond-updates/s", [round(x) for x in d["sweep_ms"]], d["roofline"].get("north_star"))
except Exception as e:
    print("$name FAILED", e); print(open("$OUT/$name.err").read()[-800:])
PY
}
if [ "$WHAT" = "c3" ] || [ "$WHAT" = "all" ]; then run r02_config3_${N}gpu --steps 3 --warmup 3 --no-cpu-baseline; fi
if [ "$WHAT" = "c4" ] || [ "$WHAT" = "all" ]; then run r02_config4_${N}gpu --S 784 --D 128 --steps 2 --warmup 4 --no-cpu-baseline; fi
if [ "$WHAT" = "c5" ] || [ "$WHAT" = "all" ]; then
  for D in 16 32 64 128 256; do run r02_config5_D${D}_f64_${N}gpu --D $D --steps 2 --warmup 4 --no-cpu-baseline --no-e2e; done
  for D in 64 128 256; do run r02_config5_D${D}_f32_${N}gpu --D $D --dtype f32 --steps 2 --warmup 4 --no-cpu-baseline --no-e2e; done
fi
