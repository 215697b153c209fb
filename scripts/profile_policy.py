"""Diagnostic: where the host time of one DFO policy action goes (cProfile over 200 actions
of the C2 policy: B = 1, N = 16384, 3 iterations)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ibc_b200.ibc import specs

device = torch.device('cuda', 0)
c = bench.CFG
net, agent, policy = bench.build_agent(c, device)
obs, _ = bench.synthetic_batch(c, 1, 5, device=device)
ts = specs.restart(obs)
for _ in range(5):
  policy.action(ts)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
  policy.action(ts)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(35)
