python -m pytest tests/test_gpu_commit.py tests/test_gpu_ipa.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ipa value=%.6g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'e2e=%.6g' % d['e2e']['value'], 'frac=%.4f' % d['roofline']['frac'], d['checked']['ok'])"
python bench.py --workload commit --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('commit value=%.6g' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'frac=%.4f' % d['roofline']['frac'], d['checked']['ok'])"
CMD2="python bench.py --batch 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-check"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:k_fixed_base_msm -s 5 -c 2 --csv --log-file gpurun_out/r02k_dram.csv $CMD2 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02k_dram.csv | cut -d, -f5,13- | tail -9
