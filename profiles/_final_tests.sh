timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_self.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench45.json 2> gpurun_out/r2_bench45.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench45.json'));print(d['ms_per_step'],{k:round(v,3) for k,v in d['config']['stage_ms'].items()},d['config']['aln_md5'][:8],d['config']['host_wall_us'])"
