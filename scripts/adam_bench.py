"""Times ldp_adam_update alone on the planner's parameter count (69.5 M floats): bytes moved = 28 per parameter."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from latent_diffusion_planning_b200 import _native as N
lib = N.load()
n = 69_500_000
p, g, m, v = (torch.randn(n, device="cuda") for _ in range(4))
v.abs_()
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    N.check(lib.ldp_adam_update(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-4, 0.9, 0.999, 1e-8, 5, 1.0, s))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    N.check(lib.ldp_adam_update(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-4, 0.9, 0.999, 1e-8, 5, 1.0, s))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"adam {ms*1e3:.1f} us per call, {n*28/ms/1e9:.2f} TB/s")
