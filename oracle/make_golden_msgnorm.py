"""Generate tests/golden/message_norm.npz: one denoiser call of the reference's OWN PharmRecDynamicsGVP with a NUMERIC
message_norm (sum aggregation divided by a constant, gvp.py:375-389, 512-517; the constructor default of the reference is 1,
configs/dev.yml uses 'mean'), on CPU over the pure-torch shims.  Test infrastructure only.

    python oracle/make_golden_msgnorm.py
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main(tags=(("n4", 4.0), ("n1", 1)), out_name="message_norm.npz"):
    MG.reference_loader.load()
    import yaml
    from pharmacoforge.config_utils.load_from_config import model_from_config
    from pharmacoforge.utils import get_batch_idxs
    from pharmacoforge_b200.synthetic import synth_state_dict
    save = {}
    for tag, nv in tags:
        cfg = copy.deepcopy(yaml.safe_load(open(os.path.join(MG.reference_loader.REFERENCE_ROOT, "configs", "dev.yml"))))
        cfg["dynamics"]["message_norm"] = nv
        torch.manual_seed(0)
        model = model_from_config(cfg)
        layout = {k: list(v.shape) for k, v in model.state_dict().items()}
        sd = synth_state_dict(layout, seed=0)
        sd["gamma.gamma"] = model.state_dict()["gamma.gamma"]
        model.load_state_dict(sd, strict=True)
        model.eval()
        sizes = [3, 5, 8, 6]
        gb, _ = MG.make_batch(cfg, pocket_seed=3, n_atoms=100, sizes=sizes)
        gen = torch.Generator().manual_seed(77)
        nf = sum(sizes)
        x_t = torch.randn(nf, 3, generator=gen) * 3.0
        h_t = torch.randn(nf, 6, generator=gen)
        tt = torch.tensor([0.37, 0.99, 0.01, 0.5])
        prot_shift = torch.randn(len(sizes), 3, generator=gen) * 2.0
        bi = get_batch_idxs(gb)
        gb.nodes["prot"].data["x_0"] = gb.nodes["prot"].data["x_0"] - gb.nodes["prot"].data["x_0"].mean(0, keepdim=True) \
            + prot_shift[bi["prot"]]
        gb.nodes["pharm"].data["x_t"] = x_t
        gb.nodes["pharm"].data["h_t"] = h_t
        cap = {}
        hooks = []
        for li, conv in enumerate(model.dynamics.noise_predictor.conv_layers):
            def grab(mod, args, out, li=li):
                for nt in ("pharm", "prot"):
                    cap[f"conv{li}_{nt}_h"] = out[nt][0].detach().clone()
                    cap[f"conv{li}_{nt}_v"] = out[nt][2].detach().clone()
            hooks.append(conv.register_forward_hook(grab))
        with torch.no_grad():
            prot_x_in = gb.nodes["prot"].data["x_0"].clone()
            eps_h, eps_x = model.dynamics(gb, tt, bi)
        for h in hooks:
            h.remove()
        if "x_t" not in save:
            save.update(sizes=np.array(sizes, np.int32), n_atoms=np.int32(100), pocket_seed=np.int32(3),
                        prot_x=prot_x_in.numpy(), x_t=x_t.numpy(), h_t=h_t.numpy(), t=tt.numpy())
        save[f"{tag}__norm"] = np.float64(nv)
        save[f"{tag}__eps_h"], save[f"{tag}__eps_x"] = eps_h.numpy(), eps_x.numpy()
        for k, v in cap.items():
            save[f"{tag}__{k}"] = v.numpy()
        print(tag, float(eps_h.abs().max()), float(eps_x.abs().max()))
    np.savez_compressed(os.path.join(MG.GOLD, out_name), **save)
    print(os.path.getsize(os.path.join(MG.GOLD, out_name)))


if __name__ == "__main__":
    main()
    # message_norm = 0: SUM / (edges into the node type per graph / nodes of the type per graph + 1), gvp.py:504-507, with the
    # per-graph edge counts exactly as add_pharm_edges records them (dynamics_gvp.py:219-221)
    main(tags=(("n0", 0),), out_name="message_norm0.npz")
