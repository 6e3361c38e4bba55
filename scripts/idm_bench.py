"""IDM reverse-diffusion loop timing: (B*Ha = 4096 rows, 2D = 530, A = 7), 100 DDPM steps, bf16 path."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import _native, handles as H, params as P  # noqa: E402

D, A, N = 265, 7, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = P.init_params(P.idm_spec(D, A), seed=1)
idm = H.Idm(p, D, A)
g = torch.Generator().manual_seed(0)
s = (torch.rand(N, 2 * D, generator=g) * 2 - 1).cuda()
a = torch.randn(N, A, generator=g).cuda()
lib = _native.load()
for _ in range(3):
    idm.sample(s, a, seed=1, n_steps=100, precision="bf16")
torch.cuda.synchronize()
lib.ldp_launch_count_reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 5
for i in range(reps):
    idm.sample(s, a, seed=i, n_steps=100, precision="bf16")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"idm loop N={N}: {ms:.2f} ms / 100 steps = {ms * 10:.1f} us/step, launches/loop {lib.ldp_launch_count() // reps}, "
      f"useful {N * 3.153e6 * 100 / ms / 1e9:.1f} TF/s")
