"""which torch (ATen) operators the stage-2 / stage-3 training steps still run, by operator and input shape (torch.profiler, record_shapes)
python tools/gpu/glue_probe.py [stage2|stage3]"""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
which = sys.argv[1] if len(sys.argv) > 1 else "stage2"
sys.argv = [sys.argv[0], "2"]
if which == "stage2":
    import runpy
    mod = runpy.run_path("tools/gpu/train_probe.py")
    step = lambda: mod["step"]()                                         # noqa: E731
    ctx = torch.no_grad()
else:
    import runpy
    mod = runpy.run_path("tools/gpu/stage3_probe.py")
    step = lambda: mod["step"]()                                         # noqa: E731
    import contextlib
    ctx = contextlib.nullcontext()
with ctx, profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages(group_by_input_shape=True):
    t = getattr(ev, "self_device_time_total", None)
    if t is None:
        t = ev.self_cuda_time_total
    if t > 0 and ev.key.startswith("aten::"):
        rows.append((t, ev.count, ev.key, str(ev.input_shapes)[:110]))
rows.sort(reverse=True)
print("%s: ATen operators by self device time" % which)
for t, n, k, shp in rows[:28]:
    print("  %8.2f ms %5d x %-28s %s" % (t / 1e3, n, k, shp))
