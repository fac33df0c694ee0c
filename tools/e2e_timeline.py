"""timeline of one host-staged step (CUDA events on the compute and the copy stream), config 5 at bench size"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from phare_b200 import solver as S

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg, _ = bench.bench_config(5, 1)
sol, ops = bench.build_gpu_solver(cfg, 1, dev)
st = S.HostStaging(ops, sol.patches)
marks = []


def mark(name, stream=None):
    e = torch.cuda.Event(enable_timing=True)
    (stream or torch.cuda.current_stream()).record_event(e)
    marks.append((name, e))


def wrap(obj, attr, before=None, after=None, stream=None):
    f = getattr(obj, attr)

    def g(*a, **k):
        if before:
            mark(before, stream() if stream else None)
        r = f(*a, **k)
        if after:
            mark(after, stream() if stream else None)
        return r
    setattr(obj, attr, g)


cs = lambda: st.copy_stream
wrap(st, "upload", "upload>", "upload<")
wrap(st, "snapshot_moments", "snapshot>", "snapshot<")
wrap(st, "download_B", None, "B d2h<", cs)
wrap(st, "download_fields", None, "E d2h<", cs)
wrap(st, "download_moments", None, "moments d2h<", cs)
wrap(sol, "_finish_particles", "finish>", "finish<")
wrap(st, "join", None, "join<")
mi = sol._move_ions
def move(dt, mode):
    mark(f"sweep{mode}>")
    mi(dt, mode)
    mark(f"sweep{mode}<")
sol._move_ions = move
for i in range(4):
    if i == 2:
        marks.clear()
        mark("step0")
    sol.advance_level(cfg.dt, staging=st)
    st.results_become_inputs()
mark("end")
st.sync()
t0 = marks[0][1]
for name, e in marks:
    print(f"{t0.elapsed_time(e):9.3f} ms  {name}")
