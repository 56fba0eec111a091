"""Golden fixture for `PharmacophoreDiff.sample()` and the trajectory frames, from the reference's OWN code.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference).  The reference's
`PharmacophoreDiff.sample` (pharmacodiff.py:516-578) is run unmodified over oracle/shims on three pockets with an
explicit per-pocket `init_pharm_com`, `max_batch_size=4` (six graphs -> chunks of 4 + 2) and
`visualize_trajectory=True`, so the fixture pins the chunking, the per-pocket COM indexing, the regrouping and the
values of every trajectory frame (`get_pos_feat_for_visual`, pharmacodiff.py:360-378).  Noise is injected per
chunk from one pre-drawn buffer [T+1, Nf_total, 9] (columns = pharmacophore nodes in pocket-major order).

    python oracle/make_golden_sample.py          # writes tests/golden/sample_multi.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from make_golden import GOLD, InjectedRandn, build_model  # noqa: E402

from pharmacoforge_b200.synthetic import make_pocket  # noqa: E402

POCKETS = [(100, 21), (60, 22), (80, 23)]          # (atoms, seed)
N_PHARMS = [[3, 5], [4], [6, 3, 4]]
MAX_BATCH = 4


def main():
    model, cfg, _ = build_model()
    from pharmacoforge.dataset.protein_pharm_dataset import build_initial_complex_graph

    T = model.n_timesteps
    refs, coms = [], []
    gen = torch.Generator().manual_seed(77)
    for n_atoms, seed in POCKETS:
        pos, onehot = make_pocket(n_atoms, seed=seed)
        # dataset-style reference graph: carries (dummy) ground-truth pharmacophore nodes, as the graphs handed to
        # sample() by the training-time evaluation do.  visualize_trajectory needs them: get_pos_feat_for_visual reads
        # pharm h_0 (pharmacodiff.py:85), which copy_graph creates (as zeros) only when the reference graph has it.
        refs.append(build_initial_complex_graph(torch.from_numpy(pos), torch.from_numpy(onehot),
                                                cutoffs=cfg["graph"]["graph_cutoffs"],
                                                pharm_atom_positions=torch.zeros(2, 3),
                                                pharm_atom_features=torch.zeros(2, 6)))
        coms.append(torch.from_numpy(pos).mean(dim=0) + torch.randn(3, generator=gen))   # ligand COM near the pocket
    init_pharm_com = torch.stack(coms)
    flat = [n for szs in N_PHARMS for n in szs]
    nf_total = sum(flat)
    noise = torch.randn(T + 1, nf_total, 9, generator=torch.Generator().manual_seed(4321))
    node_off = np.concatenate([[0], np.cumsum(flat)])

    orig = model.sample_given_receptor
    real_randn = torch.randn
    state = {"graphs_done": 0}

    def chunked(g, init_pharm_com=None, visualize_trajectory=False):
        b = g.batch_size
        lo, hi = int(node_off[state["graphs_done"]]), int(node_off[state["graphs_done"] + b])
        state["graphs_done"] += b
        torch.randn = InjectedRandn(noise[:, lo:hi])
        try:
            return orig(g, init_pharm_com=init_pharm_com, visualize_trajectory=visualize_trajectory)
        finally:
            torch.randn = real_randn

    model.sample_given_receptor = chunked
    try:
        out = model.sample(refs, N_PHARMS, max_batch_size=MAX_BATCH, init_pharm_com=init_pharm_com,
                           visualize_trajectory=True)
    finally:
        model.sample_given_receptor = orig
    assert [len(o) for o in out] == [len(s) for s in N_PHARMS]
    ph = [p for o in out for p in o]
    np.savez_compressed(
        os.path.join(GOLD, "sample_multi.npz"),
        pockets=np.array(POCKETS, np.int32), n_pharms_flat=np.array(flat, np.int32),
        n_pharms_per_pocket=np.array([len(s) for s in N_PHARMS], np.int32), max_batch_size=np.int32(MAX_BATCH),
        init_pharm_com=init_pharm_com.numpy(), noise=noise.numpy(),
        final_x=torch.cat([p.ph_coords for p in ph]).numpy(),
        final_h=torch.cat([p.g.nodes['pharm'].data['h_0'] for p in ph]).numpy(),
        final_type=torch.cat([p.ph_feats_idxs for p in ph]).numpy().astype(np.int32),
        pos_frames=torch.cat([p.pos_frames for p in ph], dim=1).numpy(),      # [T+1, Nf_total, 3]
        feat_frames=torch.cat([p.feat_frames for p in ph], dim=1).numpy(),    # [T+1, Nf_total, 6]
        xyz_first=np.array(ph[0].to_xyz_file()), traj_xyz_first=np.array(ph[0].traj_to_xyz()))
    print("sample_multi.npz", os.path.getsize(os.path.join(GOLD, "sample_multi.npz")))


if __name__ == "__main__":
    main()
