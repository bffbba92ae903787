#!/bin/bash
# BASELINE.json configs 3, 4 and 5 on N GPUs of one box (one rank per GPU, NCCL).
# Usage: tools/run_configs.sh N what...   with what in: c3 | c4 | c5 | c5:<D>:<f64|f32>
# WARM (default 8) warm-up sweeps for the D >= 128 runs: the warm-started split needs the first ~6 sweeps of a training
# run to engage there (its first attempts are refused while the tensors still change a lot between visits).
N=${1:-8}; shift; OUT=gpurun_out; mkdir -p $OUT; WARM=${WARM:-8}
run() { # name, bench args...
  name=$1; shift
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" > $OUT/$name.json 2> $OUT/$name.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" > $OUT/$name.json 2> $OUT/$name.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    js=d.get("jacobi_sweeps") or {}
    print("$name", round(d["value"],1), "bond-updates/s", [round(x) for x in d["sweep_ms"]], "e2e", d.get("e2e") and round(d["e2e"]["value"],1), "warm", js.get("warm_started_fraction"), d["roofline"].get("north_star"))
except Exception as e:
    print("$name FAILED", e); print(open("$OUT/$name.err").read()[-800:])
PY
}
for what in "$@"; do
  case $what in
    c3) run r02_config3_${N}gpu --steps 3 --warmup 3 --no-cpu-baseline;;
    c4) run r02_config4_${N}gpu --S 784 --D 128 --steps 2 --warmup $WARM --no-cpu-baseline;;
    c5) for D in 16 32 64; do run r02_config5_D${D}_f64_${N}gpu --D $D --steps 2 --warmup 4 --no-cpu-baseline --no-e2e; done
        for D in 128 256; do run r02_config5_D${D}_f64_${N}gpu --D $D --steps 2 --warmup $WARM --no-cpu-baseline --no-e2e; done
        run r02_config5_D64_f32_${N}gpu --D 64 --dtype f32 --steps 2 --warmup 4 --no-cpu-baseline --no-e2e
        for D in 128 256; do run r02_config5_D${D}_f32_${N}gpu --D $D --dtype f32 --steps 2 --warmup $WARM --no-cpu-baseline --no-e2e; done;;
    c5:*) D=$(echo $what | cut -d: -f2); T=$(echo $what | cut -d: -f3); W=4; [ $D -ge 128 ] && W=$WARM
        run r02_config5_D${D}_${T}_${N}gpu --D $D --dtype $T --steps 2 --warmup $W --no-cpu-baseline --no-e2e;;
  esac
done
