"""Per-block error budget of the production U-Net against the oracle (fp32, TF32 off, same GPU).

Shared by tests/test_gpu_parity_budget.py (a subset, asserted) and `python tests/parity_budget.py` (the full matrix of
VERDICT r1 item 2: 3 weight/input seeds x timesteps {1, 501, 991} x {R=1 single_ref, R=4 multi_ref} at the bench's
batch shape B = 10, V = 8, 64x64), which writes the table quoted in DESIGN.md."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import mmdm_oracle as O  # noqa: E402


def block_errors(unet, sd_dev, cfg, B, V, H, W, R, seed, timestep, dev):
    """-> (final max-rel over generated views, {block: max|a-b| / max|b|}) for one forward."""
    x, t, ctrl = O.make_inputs(cfg, B=B, V=V, H=H, W=W, R=R, seed=seed, timestep=timestep)
    x, t = x.to(dev), t.to(dev)
    ctrl = {k: v.to(dev) for k, v in ctrl.items()}
    unet.enable_taps(True)
    y = unet(x, timesteps=t, context=None, control=ctrl, n_ref_views=R)
    mine = unet.taps(H, W)
    unet.enable_taps(False)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        want = {}
        with torch.no_grad():
            ref = O.unet_forward(sd_dev, cfg, x, t, ctrl, taps=want)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    assert torch.equal(y[:, :R], ref[:, :R]), "reference views must be exactly x - z_input"
    errs = {}
    G = V - R
    for name, a in mine.items():
        b = want[name]
        if a.shape[0] != b.shape[0]:  # after the reference views were dropped: generated views only
            b = b.reshape(B, V, *b.shape[1:])[:, R:].reshape(B * G, *b.shape[1:])
        else:                          # all views are computed here, but only the generated ones matter downstream
            a = a.reshape(B, V, *a.shape[1:])[:, R:]
            b = b.reshape(B, V, *b.shape[1:])[:, R:]
        errs[name] = O.max_rel_err(a, b)
    return O.max_rel_err(y[:, R:], ref[:, R:]), errs


def run_matrix(dev, seeds=(0, 1, 2), timesteps=(1, 501, 991), Rs=(1, 4), B=10, log=print):
    from cap4d_b200 import B200MMDMUnet

    cfg = O.PRODUCTION_CONFIG
    rows = []
    for seed in seeds:
        sd = O.init_state_dict(cfg, seed=seed)
        unet = B200MMDMUnet(cfg, sd, device=dev)
        sd_dev = {k: v.to(dev) for k, v in sd.items()}
        del sd
        for R in Rs:
            for ts in timesteps:
                final, errs = block_errors(unet, sd_dev, cfg, B, 8, 64, 64, R, 100 * seed + ts, ts, dev)
                worst = max(errs, key=errs.get)
                rows.append(dict(seed=seed, R=R, timestep=ts, final=final, worst_block=worst,
                                 worst_block_err=errs[worst], blocks=errs))
                log(f"seed {seed} R={R} t={ts}: eps max-rel {final:.3e}; worst block {worst} {errs[worst]:.3e}")
        del unet, sd_dev
        torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    rows = run_matrix(dev)
    out = os.path.join(ROOT, "gpurun_out", "r02_parity_budget.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as fh:
        json.dump(rows, fh, indent=1)
    print("worst final", max(r["final"] for r in rows))
