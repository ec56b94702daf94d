python tools/probe_bvh_build.py 2>&1 | grep -v "^\[" | tail -24
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 1500 gpurun_out/r2_bench_a.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_a.json'));
r=d['roofline']; print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['strong_scaling'])
print({k:r[k] for k in ('bound','achieved','frac','traffic','dram_gbs','dram_frac','l2_to_l1_gbs','l2_frac','busiest_unit','algorithmic_bytes_per_step','algorithmic_bytes_per_step_reference_rays','untraced_shadow_rays_per_step')})
print(d['traversal'])
print(d['cpu_baseline'])
"
python bench.py --workload sponza --steps 8 --warmup 3 --no-cpu-baseline --strong-spp 0 > gpurun_out/r2_bench_a_sponza.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_a_sponza.json')); r=d['roofline']
print('sponza',d['value'],{k:r[k] for k in ('bound','achieved','frac','dram_frac','l2_frac','busiest_unit')}); print(d['traversal'])"
