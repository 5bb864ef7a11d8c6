timeout 600 python -m pytest tests/test_gpu_stages.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench31.json 2> gpurun_out/r2_bench31.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench31.json'));print(d['ms_per_step'],{k:round(v,3) for k,v in d['config']['stage_ms'].items()},d['config']['aln_md5'][:8],d['roofline']['frac'])"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:adaptamer_merge_kernel -s 2 -c 1 -o gpurun_out/r2_merge_c python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_merge_c.log; tail -2 gpurun_out/ncu_merge_c.log
