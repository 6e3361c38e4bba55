"""IDM sampling at row counts around the 128-row CTA blocks (diagnostics; run under compute-sanitizer for the memcheck record)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from latent_diffusion_planning_b200 import handles as H, params as P  # noqa: E402

D, A = 265, 7
idm = H.Idm(P.init_params(P.idm_spec(D, A), seed=0), D, A)      # s rows are [z_t, z_{t+1}]: 2 x obs_dim columns
g = torch.Generator().manual_seed(0)
for N in [int(x) for x in sys.argv[1:]] or [4097]:
    s = torch.randn(N, 2 * D, generator=g).cuda()
    a = torch.randn(N, A, generator=g).cuda()
    out = idm.sample(s, a, seed=1, n_steps=3, precision="bf16")
    torch.cuda.synchronize()
    print(N, tuple(out.shape), bool(torch.isfinite(out).all()))
