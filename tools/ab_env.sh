#!/bin/bash
# Same-box A/B of environment switches on the default training bench (and optionally the inference workload).
# Usage: tools/ab_env.sh OUT_PREFIX "VAR=val ..." ["VAR=val ..." ...]   ("-" = no switch)
# Each variant runs `bench.py --no-extras --no-cpu-baseline --no-eager-baseline` in its own process and leaves
# gpurun_out/OUT_PREFIX_<i>.json; a summary line per variant goes to stdout.
set -u
prefix=$1; shift
i=0
for v in "$@"; do
  envs=""; [ "$v" != "-" ] && envs="$v"
  out=gpurun_out/${prefix}_${i}.json
  env $envs python bench.py --steps ${AB_STEPS:-10} --warmup ${AB_WARMUP:-4} ${AB_ARGS:-} --no-extras --no-cpu-baseline --no-eager-baseline \
      > "$out" 2> gpurun_out/${prefix}_${i}.err || echo "variant '$v' FAILED rc=$?"
  python - "$out" "$v" <<'P'
import json, os, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    pick = {k: kc[k]["ms"] for k in (os.environ.get("AB_KERNELS") or "rbu_sa_reduce,wgrad_gemm,conv3x3").split(",") if k in kc}
    print(f"{sys.argv[2]:28s} {d['ms_per_step']:8.3f} ms/step  {d['value']:8.1f} {d['unit']}  hbm_kernels {d.get('hbm_kernels', {}).get('ms')} ms frac {d.get('hbm_kernels', {}).get('hbm_frac')}  {pick}")
except Exception as e:
    print(sys.argv[2], "unreadable:", e)
P
  i=$((i+1))
done
