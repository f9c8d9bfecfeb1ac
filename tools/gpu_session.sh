#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/s11_topo.txt 2>&1
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/s11_bench_n8.json 2> gpurun_out/s11_bench_n8.err
echo "bench n8 exit $?"; tail -2 gpurun_out/s11_bench_n8.err | cut -c1-300
python - <<'PY'
import json
for line in open('gpurun_out/s11_bench_n8.json'):
    if line.startswith('{'):
        b=json.loads(line)
        print('N=8', b['value'], b['ms_per_step'], b['detail']['stage_ms'], 'e2e', b['e2e']['ms_per_step'], 'page', b['e2e_pageable']['ms_per_step'], 'gen', b.get('e2e_generated',{}).get('ms_per_step'))
        print(json.dumps(b['parity']))
        for k,v in (b.get('extra') or {}).items():
            print(k, json.dumps({kk:v.get(kk) for kk in ('ms_per_step','patterns_per_s','rank0_stage_ms','rank0_gemm_tflops_algorithmic','checks','osm','error')}))
PY
KDI_EXCHANGE=nccl timeout 600 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-generated > gpurun_out/s11_bench_n8_nccl.json 2> gpurun_out/s11_bench_n8_nccl.err
echo "bench n8 nccl exit $?"; python -c "
import json
for line in open('gpurun_out/s11_bench_n8_nccl.json'):
    if line.startswith('{'):
        b=json.loads(line); print('N=8 nccl', b['value'], b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'])"
for n in 4 2; do
timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --no-extras --no-generated > gpurun_out/s11_bench_n$n.json 2> gpurun_out/s11_bench_n$n.err
echo "bench n$n exit $?"; python -c "
import json
for line in open('gpurun_out/s11_bench_n$n.json'):
    if line.startswith('{'):
        b=json.loads(line); print('N=$n', b['value'], b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'], b['parity'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/s11_bench_n1.json 2> gpurun_out/s11_bench_n1.err
echo "bench n1 exit $?"; python -c "
import json
b=json.load(open('gpurun_out/s11_bench_n1.json')); print('N=1', b['value'], b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'], 'gen', b['e2e_generated']['ms_per_step'], b['e2e_generated']['rank0_stage_ms'])"
KDI_TIMELINE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29541 --log-dir gpurun_out/s11_logs --redirects 3 bench.py --gpus 8 --steps 3 --warmup 3 --no-extras --no-generated --e2e-steps 1 > /dev/null 2>&1
echo "timeline n8 exit $?"; f=$(ls gpurun_out/s11_logs/*/attempt_0/0/stderr.log | head -1); grep -n "kdi timeline" $f | sed -n 4p; awk '/kdi timeline/{c++} c==4' $f | head -40 > gpurun_out/s11_timeline_n8_rank0.txt; cat gpurun_out/s11_timeline_n8_rank0.txt; rm -rf gpurun_out/s11_logs
