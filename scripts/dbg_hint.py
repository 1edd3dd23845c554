import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mmdm_oracle as O
from cap4d_b200 import B200MMDMUnet
dev = torch.device("cuda:0")
sd = O.init_state_dict(O.TINY_CONFIG, seed=0)
unet = B200MMDMUnet(O.TINY_CONFIG, sd, device=dev)
for (B, V, H, W, R) in [(2, 4, 16, 16, 1), (4, 4, 16, 16, 2)]:
    x, t, ctrl = O.make_inputs(O.TINY_CONFIG, B=B, V=V, H=H, W=W, R=R, seed=7 * B + R, timestep=333)
    kw = dict(timesteps=t.to(dev), context=None, control={k: v.to(dev) for k, v in ctrl.items()})
    full = unet(x.to(dev), **kw).cpu()
    hint = unet(x.to(dev), n_ref_views=R, **kw).cpu()
    ref = O.unet_forward(sd, O.TINY_CONFIG, x, t, ctrl)
    print("shape", (B, V, H, W, R), "full-vs-ref", O.max_rel_err(full[:, R:], ref[:, R:]), "hint-vs-ref",
          O.max_rel_err(hint[:, R:], ref[:, R:]), "hint-vs-full", O.max_rel_err(hint[:, R:], full[:, R:]))
    d = (hint - full).abs()
    print("  per (b,v) max diff:", [[round(float(d[b, v].max()), 5) for v in range(V)] for b in range(B)], "ref absmax", float(ref[:, R:].abs().max()))
    print("  rel-l2 hint-vs-full", O.rel_l2(hint[:, R:], full[:, R:]), "full-vs-ref", O.rel_l2(full[:, R:], ref[:, R:]))
