"""Where do the statistics folds cost time in a 20-step timed region?  (experiment, round 2)"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from bench import make_tape
from qcd_b200 import _lib as qlib
from qcd_b200.batch import EnvBatch
from qcd_b200.targets import make_target
dev = torch.device('cuda', 0)
N, eta, delta, R, K = 65536, 5, 20, 7, 20
target = make_target('SP-ghz5', eta, seed=42)
shared = torch.zeros((qlib.STATS_COPIES, qlib.STATS_STRIDE), dtype=torch.float64, device=dev)
bs = [EnvBatch(N, eta, delta, 'SP-ghz5', target, device=dev, stats_buffer=shared) for _ in range(R)]
tape = torch.as_tensor(make_tape(eta, 16, N, seed=1), device=dev)
stream = torch.cuda.Stream(dev)
out = torch.zeros(qlib.STATS_STRIDE, dtype=torch.float64, device=dev)
def launch(i): bs[i % R].step(tape[i % 16])
def cap(n, fold):
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g, stream=stream):
    for i in range(n): launch(i)
    if fold: bs[0].fold_stats(out=out, zero=True)
  return g
with torch.cuda.stream(stream):
  for i in range(8): launch(i)
  bs[0].fold_stats(out=out, zero=True)
  stream.synchronize()
  g20, g20f = cap(K, False), cap(K, True)
  for g in (g20, g20f): g.replay()
  stream.synchronize()
  def timeit(fn, reps=30):
    ts = []
    for _ in range(reps):
      stream.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(stream); fn(); e1.record(stream); stream.synchronize()
      ts.append(e0.elapsed_time(e1) * 1e3)
    return np.median(ts), np.min(ts)
  print('graph(20)            us', timeit(lambda: g20.replay()))
  print('fold + graph(20)     us', timeit(lambda: (bs[0].fold_stats(out=out, zero=True), g20.replay())))
  print('graph(20) + fold     us', timeit(lambda: (g20.replay(), bs[0].fold_stats(out=out, zero=True))))
  print('fold+graph(20)+fold  us', timeit(lambda: (bs[0].fold_stats(out=out, zero=True), g20.replay(), bs[0].fold_stats(out=out, zero=True))))
  print('graph(20 + fold)     us', timeit(lambda: g20f.replay()))
  print('fold alone           us', timeit(lambda: bs[0].fold_stats(out=out, zero=True)))
  ev = torch.cuda.Event()
  side = torch.cuda.Stream(dev)
  def with_side():
    bs[0].fold_stats(out=out, zero=True); ev.record(stream)
    with torch.cuda.stream(side): side.wait_event(ev)
    g20.replay()
  print('fold+event+side+graph us', timeit(with_side))
