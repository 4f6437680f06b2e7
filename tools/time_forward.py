import os, sys, torch
sys.path.insert(0, '/root/repo'); os.chdir('/root/repo')
import bench
d=None
from group_gan_gcn_gat_b200 import models as MD
torch.backends.cudnn.allow_tf32=False
dev=torch.device('cuda:0')
data=bench.synth_batch(1<<16,1236)
gen=MD.TrajectoryGenerator(obs_len=8,pred_len=12,embedding_dim=16,encoder_h_dim=32,decoder_h_dim=32,mlp_dim=64,noise_dim=(8,),noise_mix_type='global',pooling_type='pool_net',pool_every_timestep=False,bottleneck_dim=8,batch_norm=False,n_heads=1,alpha=0.2)
gen.load_state_dict(bench.load_weights(),strict=True); gen=gen.to(dev).train(); gen.pool_net.precision='bf16'
x={k:data[k].to(dev) for k in ('obs_traj','obs_traj_rel','obs_traj_g','seq_start_end')}
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
with torch.no_grad():
    z=torch.randn(1<<16,8,device=dev)
    for _ in range(3): out=gen(x['obs_traj'],x['obs_traj_rel'],x['seq_start_end'],x['obs_traj_g'],user_noise=z)
    ts=[]
    for _ in range(15):
        flush.zero_(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); out=gen(x['obs_traj'],x['obs_traj_rel'],x['seq_start_end'],x['obs_traj_g'],user_noise=z); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort(); print(os.path.basename(os.environ.get('SGX_LIB','default')),'forward median %.1f us min %.1f us checksum %.6f'%(ts[7]*1e3,ts[0]*1e3,float(out.double().sum())))
