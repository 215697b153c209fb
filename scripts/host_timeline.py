"""Diagnostic: host-side wall time of the pieces of one end-to-end C2 train step (pinned host
batch -> agent.train -> loss read back), measured with perf_counter around the agent's own calls
(no extra synchronisation: the only sync is the loss read at the end of each step)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

device = torch.device('cuda', 0); torch.cuda.set_device(0)
c = bench.CFG
net, agent, policy = bench.build_agent(c, device)
hb = [bench.synthetic_batch(c, c['batch'], 100 + i, pinned=True) for i in range(4)]
marks = {}
def wrap(obj, name, label):
  f = getattr(obj, name)
  def g(*a, **k):
    t = time.perf_counter(); r = f(*a, **k); marks[label] = marks.get(label, 0.0) + time.perf_counter() - t
    return r
  setattr(obj, name, g)
wrap(agent, '_to_device', 'to_device')
wrap(net, '_flat_obs', 'flat_obs')
wrap(agent, '_make_counter_example_actions', 'negatives+concat')
wrap(net.engine, 'train_grads', 'train_grads call')
wrap(agent._optimizer, 'apply_gradients', 'adam+params_changed')
wrap(agent, '_check_finite', 'check_finite')
wrap(agent, '_loss', 'loss total')
for i in range(10): float(agent.train((hb[i % 4], ())).loss)
marks.clear(); n = 300; tsync = 0.0
t0 = time.perf_counter()
for i in range(n):
  info = agent.train((hb[i % 4], ()))
  t = time.perf_counter(); float(info.loss); tsync += time.perf_counter() - t
tot = time.perf_counter() - t0
print('e2e %.1f us/step; loss read (sync) %.1f us' % (tot / n * 1e6, tsync / n * 1e6))
for k, v in marks.items(): print('  %-22s %.1f us' % (k, v / n * 1e6))
