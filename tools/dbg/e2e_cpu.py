import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
eng = bench.build_engine("mspa_c2f_gd_tood_yolov8n", 32, torch.bfloat16, torch.device("cuda", 0), 3)
host = [bench.make_u8(32, 100 + i).pin_memory() for i in range(6)]
for i in range(6):
    eng.collect(eng.submit(host[i % 6]))
torch.cuda.synchronize()
ts, tc, pend = 0.0, 0.0, []
t0 = time.perf_counter()
N = 60
for i in range(N):
    a = time.perf_counter()
    pend.append(eng.submit(host[i % 6]))
    b = time.perf_counter()
    ts += b - a
    if len(pend) == 6:
        eng.collect(pend.pop(0))
        tc += time.perf_counter() - b
while pend:
    eng.collect(pend.pop(0))
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"per step: wall {tot / N * 1e3:.3f} ms, submit() CPU {ts / N * 1e3:.3f} ms, collect() incl. wait {tc / N * 1e3:.3f} ms")
# pieces of submit
s = eng.slots[0]
def t(f, n=200):
    a = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - a) / n * 1e3
print("weights_changed", t(eng.weights_changed))
with torch.cuda.stream(s.stream), torch.no_grad():
    print("stem launch (CPU side)", t(lambda: eng._head(s, s.src)))
    print("graph replay (CPU side)", t(lambda: s.graph.replay(), 50))
