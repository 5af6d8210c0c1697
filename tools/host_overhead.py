"""Host-side cost of one plugin call (no device sync inside the loop): how much of the e2e step is Python."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import vipnerf_oracle as O  # noqa: E402
from vipnerf_b200 import renderpath  # noqa: E402
from vipnerf_b200.ModelFactory import get_model  # noqa: E402

dev = torch.device('cuda', 0)
model = get_model(bench.model_configs('bf16'), None)
model.load_state_dict(O.synth_state_dict(0))
model = model.to(dev).eval()
host = {k: v.pin_memory() for k, v in O.make_rays('fern', 4096, seed=2).items()}
batch = {k: v.to(dev) for k, v in host.items()}
pc, pf = model._packed_weights('coarse', 'bf16', dev), model._packed_weights('fine', 'bf16', dev)
keys = renderpath.pass_keys(True, False, 0)
with torch.no_grad():
    for name, fn in (('model(batch)', lambda: model(dict(batch))),
                     ('renderpath.render_rays', lambda: renderpath.render_rays(batch, pc, pf, ndc=True, precision='bf16', keys=keys)),
                     ('9 x H2D', lambda: {k: v.to(dev, non_blocking=True) for k, v in host.items()})):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 300
        for _ in range(n):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f'{name:28s} host {1e6 * (t1 - t0) / n:7.1f} us/call   (device drained after {1e3 * (t2 - t1):.1f} ms)')
