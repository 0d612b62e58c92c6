#!/bin/bash
# Two-GPU evidence (gpurun --gpus 2): the driver's N = 2 bench command (NCCL weight broadcast through c3b_bcast_weights, weak-scaled
# networks, strong-scaled cascade) and the launcher on two ranks.
set -u
mkdir -p gpurun_out
R=${1:-r2n2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err
tail -c 600 gpurun_out/${R}_bench_n2.err
python - "$R" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/%s_bench_n2.json'%sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print('n_gpus', d['n_gpus'], 'broadcast', d['weight_broadcast'])
    for k,v in d['workloads'].items():
        print(k, v['scaling'], 'value %.4g'%v['value'], 'e2e %.4g'%v['e2e']['value'], 'region %.2f'%v['timed_region_s'])
except Exception as e: print('parse failed', e)
PY
# launcher on two ranks: synthetic .npy/.info chunk files, reference-format shards back, checked against a single forward
python - <<'PY'
import os, numpy as np, torch, sys
sys.path.insert(0, '.')
from clair3_b200 import synth
os.makedirs('gpurun_out/_launcher', exist_ok=True)
d='gpurun_out/_launcher'
names=[]
r=np.random.default_rng(0)
for i,n in enumerate([1300, 700, 2100, 40, 1024, 999]):
    p=os.path.join(d,'pileup_chr20_%d'%i)
    np.save(p, synth.pileup_inputs(n, seed=400+i, dtype=np.int8))
    with open(p+'.info','w') as f:
        for j in range(n):
            f.write('chr20:%d:%s\t%d-X%s 3\n'%(10000*i+j, ''.join(r.choice(list('ACGT'),33)), int(r.integers(5,90)), 'ACGT'[j%4]))
    names.append(os.path.basename(p))
open(os.path.join(d,'file_list'),'w').write('\n'.join(names)+'\n')
sd=synth.pileup_state_dict(False, seed=9)
torch.save({k: torch.from_numpy(np.asarray(v)) for k,v in sd.items()}, os.path.join(d,'pileup.pt'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m clair3_b200.launcher \
    --file_list gpurun_out/_launcher/file_list --chkpnt_fn gpurun_out/_launcher/pileup --pileup --out_prefix gpurun_out/_launcher/pred > gpurun_out/${R}_launcher.log 2>&1
tail -3 gpurun_out/${R}_launcher.log
python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0,'.')
from clair3_b200 import synth, launcher
from clair3_b200.model import Clair3_P
d='gpurun_out/_launcher'
files=launcher.read_file_list(d+'/file_list')
sd=synth.pileup_state_dict(False, seed=9)
m=Clair3_P(False,True,18); m.to(torch.device('cuda')); m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k,v in sd.items()}); m.set_option('lstm_tile',64)
ok=True; tot=0
for rank,fl in enumerate(launcher.split_file_list(files,2)):
    x=np.concatenate([np.load(f+'.npy') for f in fl]); y=m(torch.from_numpy(x)).numpy()
    pred=np.load(d+'/pred_%d.prediction'%rank); pos=np.load(d+'/pred_%d.position'%rank)
    want_pos=[l.split('\t')[0] for f in fl for l in open(f+'.info').read().strip().split('\n')]
    ok &= pred.shape==y.shape and float(np.abs(pred-y).max())<1e-5 and [p[0].decode() for p in pos]==want_pos
    tot+=len(pred)
print('launcher 2-rank shards match a single-process forward:', ok, 'rows', tot)
PY
rm -rf gpurun_out/_launcher
