"""Per-kernel time INSIDE the step graph (torch.profiler / CUPTI on graph replays: warm caches, real back-to-back execution --
unlike the ncu launch lists, which serialise and flush).  python tools/prof_step.py [wgan_gp|wgan|gan] > profiles/..._step_kernels.txt"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

kind = sys.argv[1] if len(sys.argv) > 1 else 'wgan_gp'
dev = torch.device('cuda:0')
w = bench.make_cnn_workload(kind, 32 if kind == 'gan' else 64, 0, 1, dev)
for _ in range(3):
    w.body()
run, mode = bench.capture(w.body, False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_(); run()
torch.cuda.synchronize()
REPS = 5
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(REPS):
        flush.zero_(); run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and 'Memset' not in e.name and 'Memcpy' not in e.name]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
# split into steps at the flush kernel (the 256 MiB fill)
steps, cur = [], []
for e in evs:
    if 'FillFunctor<unsigned char>' in e.name:
        if cur: steps.append(cur)
        cur = []
    else:
        cur.append(e)
if cur: steps.append(cur)
steps = [s for s in steps if len(s) > 20]
last = steps[-1]
t0, t1 = last[0].time_range.start, last[-1].time_range.end
busy = sum(e.time_range.end - e.time_range.start for e in last)
print('# %s step, %s; last of %d profiled replays: %d kernels, span %.1f us, sum of kernel durations %.1f us, gaps %.1f us' %
      (kind, mode, len(steps), len(last), t1 - t0, busy, (t1 - t0) - busy))
tot = collections.defaultdict(lambda: [0, 0.0])
for s in steps:
    for e in s:
        k = e.name.split('(')[0].replace('void ', '').replace('sg::', '')[:70]
        tot[k][0] += 1; tot[k][1] += e.time_range.end - e.time_range.start
n = len(steps)
print('%-72s %6s %9s %7s' % ('kernel (per step, mean over %d steps)' % n, 'n', 'us', 'share'))
allus = sum(v[1] for v in tot.values()) / n
for k, (c, us) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print('%-72s %6.1f %9.1f %6.1f%%' % (k, c / n, us / n, 100 * us / n / allus))
print('# ordered kernels of the last step (start us, duration us, gap before us):')
prev = t0
for e in last:
    print('%9.1f %8.1f %6.1f  %s' % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.time_range.start - prev,
                                     e.name.split('(')[0].replace('void ', '').replace('sg::', '')[:80]))
    prev = e.time_range.end
