#!/usr/bin/env bash
# Scaling measurements on one multi-GPU box: bash tools/scale_run.sh OUTDIR "8 4 2 1"
OUT=${1:-gpurun_out/scale}; NS=${2:-"8 4 2 1"}
mkdir -p "$OUT"
run() {  # n, tag, extra args...
  local n=$1 tag=$2; shift 2
  if [ "$n" = 1 ]; then timeout 300 python bench.py --gpus 1 "$@" > "$OUT/${tag}_n$n.json" 2> "$OUT/${tag}_n$n.err"
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus "$n" "$@" > "$OUT/${tag}_n$n.json" 2> "$OUT/${tag}_n$n.err"; fi
  echo "$tag n=$n rc=$? $(head -c 220 "$OUT/${tag}_n$n.json")"
}
first=1
for n in $NS; do
  if [ $first = 1 ]; then run "$n" weak --steps 20 --warmup 3 --no-cpu; first=0; else run "$n" weak --steps 20 --warmup 3 --no-ops --no-cpu; fi
done
for n in $NS; do run "$n" strong_c5 --steps 10 --warmup 3 --scaling strong --workload c5 --quick --repeats 3; done
for n in $NS; do run "$n" strong_c4loss --steps 10 --warmup 3 --scaling strong --workload c4loss --repeats 3; done
