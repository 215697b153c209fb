"""Diagnostic: host-side profile (cProfile) of the end-to-end C2 train step of bench.py
(pinned host batch -> H2D -> agent.train -> loss read back), 300 steps."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

device = torch.device('cuda', 0)
torch.cuda.set_device(0)
c = bench.CFG
net, agent, policy = bench.build_agent(c, device)
hb = [bench.synthetic_batch(c, c['batch'], 100 + i, pinned=True) for i in range(4)]
def step(i):
  info = agent.train((hb[i % 4], ()))
  return float(info.loss)
for i in range(10): step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(300): step(i)
torch.cuda.synchronize()
print('e2e ms/step %.4f' % ((time.perf_counter() - t0) / 300 * 1e3))
# host time to enqueue a step without the loss read
t0 = time.perf_counter()
for i in range(300): agent.train((hb[i % 4], ()))
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('enqueue-only host ms/step %.4f, drained %.4f' % ((t1 - t0) / 300 * 1e3, (t2 - t0) / 300 * 1e3))
pr = cProfile.Profile(); pr.enable()
for i in range(300): step(i)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
