#!/bin/bash
# The command set of the current GPU session (rewritten per session; results land in gpurun_out/).
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s8_topo.txt 2>&1
timeout 600 python tools/schedule_sweep.py > gpurun_out/s8_sweep.jsonl 2> gpurun_out/s8_sweep.err
echo "sweep exit $?"; cat gpurun_out/s8_sweep.jsonl; tail -3 gpurun_out/s8_sweep.err | cut -c1-300
KDI_TIMELINE=1 ROUNDS=1 REPS=3 SETTINGS="flags=0" timeout 300 python tools/schedule_sweep.py > gpurun_out/s8_timeline.out 2> gpurun_out/s8_timeline.txt
echo "timeline exit $?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s8_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/s8_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s8_smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/s8_smoke.log
B="bench.py --gpus 2 --steps 10 --warmup 3 --extras 4,5 --no-generated"
KDI_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 $B > gpurun_out/s8_bench_n2_peer.json 2> gpurun_out/s8_bench_n2_peer.err
echo "bench n2 peer exit $?"; head -c 2500 gpurun_out/s8_bench_n2_peer.json; grep "kdi trace" gpurun_out/s8_bench_n2_peer.json gpurun_out/s8_bench_n2_peer.err | tail -2
M=100000 N=37500 KEEP=50 ROUNDS=2 REPS=2 SETTINGS="flags=0" timeout 300 python tools/schedule_sweep.py > gpurun_out/s8_sweep_c4.jsonl 2> gpurun_out/s8_sweep_c4.err
echo "c4 sweep exit $?"; cat gpurun_out/s8_sweep_c4.jsonl
