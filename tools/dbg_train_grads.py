import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from conftest import load_golden, state_dict_of
from oracle import sgan_oracle as O
import group_gan_gcn_gat_b200.models as MD
from group_gan_gcn_gat_b200 import losses
torch.backends.cudnn.allow_tf32 = False
DEV = 'cuda:0'
g = load_golden('generator_gat_zara1')
gen = MD.TrajectoryGenerator(obs_len=8, pred_len=12, embedding_dim=16, encoder_h_dim=32, decoder_h_dim=32, mlp_dim=64,
                             noise_dim=(8,), noise_mix_type='global', pooling_type='pool_net', pool_every_timestep=False,
                             bottleneck_dim=8, batch_norm=False, n_heads=1, alpha=0.2)
gen.load_state_dict(state_dict_of(g), strict=True)
gen = gen.to(DEV).train()
obs, obs_rel, grp, sse = g['obs_traj'], g['obs_traj_rel'], g['obs_traj_g'], g['seq_start_end']
torch.manual_seed(0)
up = torch.randn(12, obs.shape[1], 2)
# GPU
rel = gen(obs.to(DEV), obs_rel.to(DEV), sse.to(DEV), grp.to(DEV), user_noise=g['noise'][0].to(DEV))
(rel * up.to(DEV)).sum().backward()
# CPU oracle
sd = {k: v.clone().requires_grad_(True) for k, v in state_dict_of(g).items()}
cfg = dict(pred_len=12, wiring='gat', pooling=True, pool_every_timestep=False, alpha=0.2, n_heads=1)
relc = O.generator_forward(obs, obs_rel, sse, grp, sd, cfg, g['noise'][0])
(relc * up).sum().backward()
print('fwd err', (rel.detach().cpu() - relc.detach()).abs().max().item())
for k, p in gen.named_parameters():
    if sd[k].grad is None: continue
    r = sd[k].grad
    print('%-45s max|ref| %.3e  abs err %.3e  rel %.2e' % (k, r.abs().max().item(), (p.grad.cpu() - r).abs().max().item(),
          (p.grad.cpu() - r).abs().max().item() / r.abs().max().item()))
print(sse.tolist())

# ---- isolate the pool op on the tensors it actually saw ----
cap = {}
def hook(mod, inp, out):
    cap['h'] = inp[0].detach().clone(); cap['pos'] = inp[2].detach().clone(); cap['out'] = out
    out.register_hook(lambda gr: cap.__setitem__('gout', gr.detach().clone()))
hnd = gen.pool_net.register_forward_hook(hook)
gen.zero_grad()
rel = gen(obs.to(DEV), obs_rel.to(DEV), sse.to(DEV), grp.to(DEV), user_noise=g['noise'][0].to(DEV))
(rel * up.to(DEV)).sum().backward()
hnd.remove()
h, pos, gout = cap['h'].cpu(), cap['pos'].cpu(), cap['gout'].cpu()
psd = {k[len('pool_net.'):]: v.clone().requires_grad_(True) for k, v in state_dict_of(g).items() if k.startswith('pool_net.')}
hc = h.clone().requires_grad_(True)
oc = O.pool_hidden_net(hc, sse, pos, psd)
print('pool fwd err', (oc.detach() - cap['out'].detach().cpu()).abs().max().item(), 'out max', oc.abs().max().item(),
      'frac zero', (oc == 0).float().mean().item())
(oc * gout).sum().backward()
# GPU isolated
import group_gan_gcn_gat_b200.modules as M
pm = M.PoolHiddenNet(embedding_dim=16, h_dim=32, mlp_dim=64, bottleneck_dim=8, batch_norm=False)
pm.load_state_dict({k: v.detach() for k, v in psd.items()}); pm = pm.to(DEV)
hg = h.to(DEV).requires_grad_(True)
og = pm(hg, sse.to(DEV), pos.to(DEV))
(og * gout.to(DEV)).sum().backward()
print('isolated dh rel err', ((hg.grad.cpu() - hc.grad).abs().max() / hc.grad.abs().max()).item())
for k, p in pm.named_parameters():
    r = psd[k].grad
    print('isolated %-28s rel %.2e' % (k, ((p.grad.cpu() - r).abs().max() / r.abs().max()).item()))
# how close are the top-2 candidates?
v, idx = O.pool_hidden_net_argmax(h, sse, pos, {k: v.detach() for k, v in psd.items()})
sched_arg = None
from group_gan_gcn_gat_b200 import ops
from group_gan_gcn_gat_b200.schedule import get_schedule
sc = get_schedule(sse, DEV)
l1, l2 = pm.mlp_pre_pool[0], pm.mlp_pre_pool[2]
o2, a2 = ops.pool_fwd(h.reshape(-1, 32).to(DEV), pos.to(DEV), sc.ped_start, sc.ped_end, sc.pair_off, sc.tile_first, sc.n_pairs,
                      pm.spatial_embedding.weight, pm.spatial_embedding.bias, l1.weight, l1.bias, l2.weight, l2.bias, 0)
mism = ((a2.cpu().long() != idx) & (v > 0))
print('argmax mismatches on positive outputs:', int(mism.sum()), 'of', int((v > 0).sum()))

# ---- CPU full graph with the pooled tensor exposed ----
import torch.nn.functional as F
sd2 = {k: v.clone().requires_grad_(True) for k, v in state_dict_of(g).items()}
hE = O.encoder(obs_rel, sd2, 'encoder.')
ph = O.pool_hidden_net(hE, sse, obs[-1], sd2, 'pool_net.')
ph.retain_grad()
ctx = torch.cat([hE.view(-1, 32), ph], 1)
ctx.retain_grad()
c2 = O.gat_encoder(ctx, sse, obs[-1], grp[-1], sd2, 'gatencoder.', 0.2, 1)
c2 = O.add_global_noise(c2, sse, g['noise'][0])
relc2, _ = O.decoder(obs[-1], obs_rel[-1], (c2.unsqueeze(0), torch.zeros(1, c2.shape[0], 32)), sse, sd2, 'decoder.', 12, False)
(relc2 * up).sum().backward()
print('cpu gout vs gpu gout: max abs diff', (ph.grad - gout).abs().max().item(), 'max |gout|', gout.abs().max().item())
d = (ph.grad - gout).abs()
print('diff on active (out>0):', d[oc.detach() > 0].max().item(), ' on inactive:', d[oc.detach() == 0].max().item() if (oc == 0).any() else 0)
print('full-graph GPU dW1 vs isolated GPU dW1:', (gen.pool_net.mlp_pre_pool[0].weight.grad.cpu() - pm.mlp_pre_pool[0].weight.grad.cpu()).abs().max().item())
print('full-graph CPU dW1 vs isolated CPU dW1:', (sd2['pool_net.mlp_pre_pool.0.weight'].grad - psd['mlp_pre_pool.0.weight'].grad).abs().max().item())
print('GPU h vs CPU h: max abs diff', (h.reshape(-1, 32) - hE.detach().reshape(-1, 32)).abs().max().item())
print('GPU pos vs CPU pos', (pos - obs[-1]).abs().max().item())
for k in psd:
    a, b = psd[k].grad, sd2['pool_net.' + k].grad
    print('isolatedCPU(GPU inputs) vs fullCPU', k, (a - b).abs().max().item())
hc2 = hE.detach().clone().requires_grad_(True)
psd4 = {k: v.detach().clone().requires_grad_(True) for k, v in psd.items()}
o4 = O.pool_hidden_net(hc2, sse, obs[-1], psd4)
(o4 * ph.grad).sum().backward()
print('isolatedCPU(CPU inputs) vs fullCPU dW1', (psd4['mlp_pre_pool.0.weight'].grad - sd2['pool_net.mlp_pre_pool.0.weight'].grad).abs().max().item())
o5 = O.pool_hidden_net(hc2, sse, obs[-1], {k: v.detach().clone().requires_grad_(True) for k, v in psd.items()})
print('psd weights equal sd2 weights:', all(torch.equal(psd[k].detach(), sd2['pool_net.' + k].detach()) for k in psd))
