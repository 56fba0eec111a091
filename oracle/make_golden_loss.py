"""Generate tests/golden/forward_loss.npz by running the reference's OWN PharmacophoreDiff.forward (pharmacodiff.py:
162-243) on CPU over the pure-torch shims, with the timestep and Gaussian draws injected so that the CUDA path can
consume exactly the same ones.  Test infrastructure only (see oracle/make_golden.py).

    python oracle/make_golden_loss.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    model, cfg, _ = MG.build_model()
    sizes = [4, 6, 8, 5]
    gb, _ = MG.make_batch(cfg, pocket_seed=5, n_atoms=120, sizes=sizes)
    gen = torch.Generator().manual_seed(2024)
    nf = sum(sizes)
    prot_x = gb.nodes["prot"].data["x_0"].clone()
    centre = prot_x.mean(0, keepdim=True)
    x0 = centre + torch.randn(nf, 3, generator=gen) * 2.5
    types = torch.randint(0, 6, (nf,), generator=gen)
    h0 = torch.nn.functional.one_hot(types, 6).float()
    t_int = torch.tensor([37, 99, 0, 63])
    eps_h = torch.randn(nf, 6, generator=gen)
    eps_x = torch.randn(nf, 3, generator=gen)
    gb.nodes["pharm"].data["x_0"] = x0.clone()
    gb.nodes["pharm"].data["h_0"] = h0.clone()

    draws = [eps_h, eps_x]  # the reference draws h first, then x (pharmacodiff.py:189-192)
    real_randn, real_randint = torch.randn, torch.randint

    def fake_randn(*shape, **kw):
        out = draws.pop(0)
        shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        assert tuple(out.shape) == shp, (out.shape, shp)
        return out.clone()

    def fake_randint(low, high, size, **kw):
        assert tuple(size) == (len(sizes),) and high == model.n_timesteps
        return t_int.clone()

    torch.randn, torch.randint = fake_randn, fake_randint
    try:
        # eval mode = dropout off (the only train / eval difference of the reference forward); gradients stay enabled so
        # that the same call also yields the reference's own backward (pharmacodiff.py:265-297: total = pos + feat loss)
        losses, metrics = model.forward(gb, phase="val")
    finally:
        torch.randn, torch.randint = real_randn, real_randint
    total = torch.stack(list(losses.values())).sum()
    model.zero_grad()
    total.backward()
    names, norms, sums, dead = [], [], [], []
    keep = {}
    for k, p_ in sorted(model.named_parameters()):
        if p_.numel() == 0:
            continue
        if p_.grad is None:
            dead.append(k)
            continue
        names.append(k)
        norms.append(float(p_.grad.double().norm()))
        sums.append(float(p_.grad.double().sum()))
        if k in ("dynamics.pharm_encoder.0.weight", "dynamics.noise_predictor.noise_predictor.to_scalar_output.weight",
                 "dynamics.noise_predictor.conv_layers.0.edge_message_fns.prot_pp_prot.0.Wh",
                 "dynamics.noise_predictor.conv_layers.1.node_update_fns.pharm.1.scalar_to_vector_gates.bias",
                 "dynamics.noise_predictor.conv_layers.0.message_layer_norms.prot.feat_norm.weight"):
            keep["grad__" + k] = p_.grad.detach().numpy().copy()
    out = {k.replace(" ", "_"): np.asarray(float(v), dtype=np.float64) for k, v in {**losses, **metrics}.items()}
    print(out, len(names), "params with grad;", len(dead), "without (dead last-layer protein side):", dead[:3], "...")
    np.savez(os.path.join(MG.GOLD, "forward_loss.npz"), pocket_seed=5, n_atoms=120, sizes=np.asarray(sizes),
             x0=x0.numpy(), h0=h0.numpy(), t_int=t_int.numpy(), eps_h=eps_h.numpy(), eps_x=eps_x.numpy(),
             grad_names=np.asarray(names), grad_norms=np.asarray(norms), grad_sums=np.asarray(sums),
             dead_params=np.asarray(dead), **keep, **out)


if __name__ == "__main__":
    main()
