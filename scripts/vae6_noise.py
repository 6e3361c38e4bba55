import sys, torch
sys.path.insert(0, "/root/repo")
from latent_diffusion_planning_b200 import handles as H, params as P
from oracle import ldp_oracle as O
blocks = (128, 256, 256, 256, 256, 256)
for seed in (3, 4, 5):
    p = P.init_params(P.vae_encoder_spec(blocks), seed=seed, perturb=0.1)
    vae = H.VaeEncoder(p, blocks)
    g = torch.Generator().manual_seed(7 + seed)
    img = torch.randint(0, 256, (5, 64, 64, 3), generator=g, dtype=torch.int32).to(torch.uint8)
    with torch.no_grad():
        ref = O.vae_encode_mean(p, img.double() / 255 * 2 - 1, blocks, dtype=torch.float32)
    out = vae.encode(img.cuda(), precision="bf16").cpu()
    d = (out.double() - ref.double())
    print("seed", seed, "rel_l2", float(d.norm() / ref.double().norm()), "max", float(d.abs().max()), "ref max", float(ref.abs().max()))
    vae.close()
