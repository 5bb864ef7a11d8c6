timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench36.json 2> gpurun_out/r2_bench36.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench36.json'));print(d['ms_per_step'],{k:round(v,3) for k,v in d['config']['stage_ms'].items()},d['config']['aln_md5'][:8],d['roofline']['frac'])"
timeout 600 python bench.py --per-gpu-bp 1000000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench36_1g.json 2> gpurun_out/r2_bench36_1g.err
echo "rc=$?"; tail -3 gpurun_out/r2_bench36_1g.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench36_1g.json'));print(d['value'],d['ms_per_step'],d['e2e'],{k:round(v,2) for k,v in d['config']['stage_ms'].items()},d['config']['alignments'],d['config']['seeds'],d['config']['kmers'],d['config']['extend_cycles'])"
nvidia-smi --query-gpu=memory.used --format=csv
