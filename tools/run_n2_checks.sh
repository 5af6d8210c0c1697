# 2-GPU checks of the multi-GPU paths (gpurun --gpus 2): peer-store gather vs NCCL gather, frame workload, DataParallel
python bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_n2_peer.json 2> gpurun_out/r02_n2_peer.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_n2_peer.err | tail -20
python bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --gather nccl > gpurun_out/r02_n2_nccl.json 2> gpurun_out/r02_n2_nccl.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_n2_nccl.err | tail -20
python bench.py --gpus 2 --steps 20 --warmup 5 --workload frame > gpurun_out/r02_frame_n2_peer.json 2> gpurun_out/r02_frame_n2_peer.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_frame_n2_peer.err | tail -20
python bench.py --gpus 1 --steps 20 --warmup 5 --workload frame > gpurun_out/r02_frame_n1.json 2> gpurun_out/r02_frame_n1.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-train --no-cpu-baseline > gpurun_out/r02_n1.json 2> gpurun_out/r02_n1.err; tail -c 600 gpurun_out/r02_n1.err
python -m pytest tests/test_gpu_dataparallel.py tests/test_gpu_hostio.py -q -m gpu --timeout 600 2>&1 | tail -15
python - <<'PY'
import json
for f in ('r02_n1','r02_n2_peer','r02_n2_nccl','r02_frame_n2_peer','r02_frame_n1'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, '%.3fM'%(d['value']/1e6), 'e2e %.3fM'%(e.get('value')/1e6), 'sync', e.get('value_host_sync_per_step'), 'plain', e.get('plain_plugin_call',{}).get('value'), e.get('plain_plugin_call',{}).get('value_host_sync_per_step'), d.get('sharded_equals_unsharded'), e.get('result_equals_value_path'), d['ms_per_step'])
    except Exception as ex: print(f, 'ERR', ex)
PY
