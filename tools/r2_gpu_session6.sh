python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 12 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_b.json')); r=d['roofline']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'strong', d['strong_scaling']['seconds'], d['strong_scaling']['scene_prep'])
print({k:r[k] for k in ('bound','achieved','frac','dram_frac','l2_frac','algorithmic_bytes_per_step','algorithmic_bytes_per_step_reference_rays','untraced_shadow_rays_per_step')})
print(d['cpu_baseline'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_b_reference.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_b_reference.json')); print(d['impl'], d['value'], d['cpu_baseline']); print(d.get('native_so_loaded'))"
