#!/bin/bash
# round 2, last pass: full GPU suite, lanes-per-bucket at the mid sizes and lanes-per-job of the IPA cross terms re-checked on the
# final kernels, then the default line and the reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests.log; tail -2 gpurun_out/r02_gpu_tests.log
run() { python bench.py --workload msm --log2n $1 --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'e2e=%.4g' % d['e2e']['value'], d['checked']['ok'])"; }
for l in 18 19; do for p in 0 2 4 8; do echo -n "log2n=$l P=$p: "; VKZG_MSM_P=$p run $l; done; done
for j in 8 16; do echo -n "ipa lpj=$j: "; VKZG_FB_LPJ=$j python bench.py --workload ipa --steps 5 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value=%.4g' % d['value'], 'e2e=%.4g' % d['e2e']['value'], d['checked']['ok'])"; done
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["checked"]["ok"])
for k,v in d["also"].items(): print(k, v["value"], v["e2e"]["value"], (v.get("roofline") or {}).get("frac"), v.get("checked",{}).get("ok"))
PY
