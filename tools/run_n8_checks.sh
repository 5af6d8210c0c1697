# 8-GPU checks (gpurun --gpus 8): weak-scaling batch bench at N = 8 and 4, strong-scaling frame at N = 8 (fern + dtu)
for n in 8 4; do
  python bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_n$n.json 2> gpurun_out/r02_n$n.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_n$n.err | tail -5
done
python bench.py --gpus 8 --steps 20 --warmup 5 --workload frame > gpurun_out/r02_frame_n8.json 2> gpurun_out/r02_frame_n8.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_frame_n8.err | tail -5
python bench.py --gpus 8 --steps 20 --warmup 5 --workload frame --scene dtu > gpurun_out/r02_frame_dtu_n8.json 2> gpurun_out/r02_frame_dtu_n8.err; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_frame_dtu_n8.err | tail -5
python bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --gather nccl > gpurun_out/r02_n8_nccl.json 2> gpurun_out/r02_n8_nccl.err
python - <<'PY'
import json
for f in ('r02_n8','r02_n4','r02_n8_nccl','r02_frame_n8','r02_frame_dtu_n8'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, '%.3fM'%(d['value']/1e6), 'e2e %.3fM'%(e.get('value')/1e6), 'sync', e.get('value_host_sync_per_step'), d.get('sharded_equals_unsharded'), d['ms_per_step'], (e.get('gather') or d['config'].get('gather') or '')[:40])
    except Exception as ex: print(f, 'ERR', ex)
PY
