#!/bin/bash
# experiment driver (runs on the GPU box): tools/gpu_exp.sh TAG "variant:mode[:extra bench args]" ...
# variant = main | NAME of variants/libpngb200_NAME.so
tag=$1; shift
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu"
names=""
for spec in "$@"; do
    IFS=: read -r var mode extra <<< "$spec"
    n="${var}_m${mode}"
    names="$names $n"
    if [ "$var" = main ]; then lib=""; else lib="variants/libpngb200_$var.so"; fi
    (PNGB200_LIB=$lib timeout 240 $B --inflate-mode $mode $extra > gpurun_out/${tag}_$n.json 2> gpurun_out/${tag}_$n.err)
done
python - "$tag" $names <<'PY'
import json, sys
tag = sys.argv[1]
for n in sys.argv[2:]:
    try:
        d = json.load(open("gpurun_out/%s_%s.json" % (tag, n))); c = d["inflate_stats_per_step"]; w = max(c["waves"], 1)
        print(n, round(d["value"]), {k: round(v, 1) for k, v in d["roofline"]["stage_ms"].items()},
              {k: round(v / w) for k, v in c["cycles"].items() if v}, "rounds/wave %.2f" % (c["resolve_rounds"] / w), "bit_exact", d.get("bit_exact"))
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/%s_%s.err" % (tag, n)).read()[-400:])
PY
