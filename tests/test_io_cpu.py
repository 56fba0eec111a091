"""Host-side front / back end (pharmacoforge_b200/io.py): PDB pocket extraction, the processed-dataset reader and the
validity metric.  CPU only."""
import numpy as np
import pytest
import torch

from pharmacoforge_b200 import io as pio

ELEMENTS = ['C', 'N', 'O', 'S', 'P', 'F', 'Cl', 'Br', 'I', 'B', 'D']   # configs/dev.yml:57


def pdb_line(rec, serial, name, resname, chain, resseq, x, y, z, element):
    return f"{rec:<6}{serial:>5} {name:<4} {resname:>3} {chain}{resseq:>4}    {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00          {element:>2}"


@pytest.fixture
def pdb_file(tmp_path):
    rows, serial = [], 1
    spec = [("ALA", "A", 1, (0, 0, 0)), ("GLY", "A", 2, (6, 0, 0)), ("LEU", "A", 3, (30, 0, 0)), ("MSE", "A", 4, (1, 1, 0)),
            ("CYS", "B", 7, (0, 5, 0))]
    for resname, chain, num, (x, y, z) in spec:
        for name, el, dx in (("N", "N", 0.0), ("CA", "C", 1.0), ("HA", "H", 1.5), ("SG" if resname == "CYS" else "O",
                                                                                     "S" if resname == "CYS" else "O", 2.0)):
            rows.append(pdb_line("ATOM", serial, name, resname, chain, num, x + dx, y, z, el))
            serial += 1
    rows.append(pdb_line("ATOM", serial, "ZN", "ALA", "A", 1, 0.5, 0.5, 0.5, "ZN"))      # an 'other' element inside ALA 1
    rows.append(pdb_line("HETATM", serial + 1, "O", "HOH", "A", 100, 0.2, 0.2, 0.2, "O"))
    p = tmp_path / "rec.pdb"
    p.write_text("\n".join(rows) + "\nEND\n")
    return p


def test_pocket_from_pdb_by_ligand(pdb_file, tmp_path):
    lig = np.array([[0.5, 0.0, 0.0], [1.5, 0.0, 0.0], [2.0, 1.0, 0.0]], dtype=np.float32)
    pocket, com = pio.pocket_from_pdb(pdb_file, ELEMENTS, pocket_cutoff=4.0, lig_coords=lig)
    # ALA 1 and GLY 2 (closest atom 6.0 - 2.0 = 4.0 away: NOT < 4) ... only ALA 1 by distance; MSE is not a standard
    # amino acid; CYS B7 is 5 A away; water is a HETATM; hydrogens and the Zn atom are dropped
    assert pocket.prot_x.shape == (3, 3) and pocket.prot_h.shape == (3, 11)
    assert torch.allclose(com, torch.from_numpy(lig.mean(0, keepdims=True)))
    assert pocket.prot_h.argmax(1).tolist() == [1, 0, 2]                     # N, C, O
    pocket8, _ = pio.pocket_from_pdb(pdb_file, ELEMENTS, pocket_cutoff=8.0, lig_coords=lig)
    assert pocket8.prot_x.shape[0] == 9 and pocket8.prot_h[:, 3].sum() == 1  # + GLY 2 and CYS B7 (one sulphur)
    sdf = tmp_path / "lig.sdf"
    body = "".join(f"{x:10.4f}{y:10.4f}{z:10.4f} {e:<3} 0  0  0  0  0  0  0  0  0  0  0  0\n"
                   for (x, y, z), e in zip(list(lig) + [np.array([9., 9., 9.])], ["C", "N", "O", "H"]))
    sdf.write_text("lig\n  test\n\n" + f"{4:3d}{0:3d}  0  0  0  0  0  0  0  0999 V2000\n" + body + "M  END\n$$$$\n")
    pocket_s, com_s = pio.pocket_from_pdb(pdb_file, ELEMENTS, pocket_cutoff=4.0, lig_file=sdf)
    assert torch.equal(pocket_s.prot_x, pocket.prot_x) and torch.allclose(com_s, com)


def test_pocket_from_pdb_by_residue_list(pdb_file):
    pocket, com = pio.pocket_from_pdb(pdb_file, ELEMENTS, residue_list=["A:2", "B:7"])
    assert pocket.prot_x.shape[0] == 6
    with pytest.raises(ValueError):
        pio.pocket_from_pdb(pdb_file, ELEMENTS)
    with pytest.raises(KeyError):
        pio.pocket_from_pdb(pdb_file, ELEMENTS, residue_list=["A:55"])


def test_dataset_reader_and_subsampling(tmp_path):
    rng = np.random.default_rng(0)
    truth = []
    for split in (0, 1, 2):
        d = tmp_path / f"split_{split}"
        d.mkdir()
        n_items = 3
        prot_n, ph_n, rp_n = rng.integers(20, 40, n_items), rng.integers(3, 12, n_items), rng.integers(2, 9, n_items)
        mk = lambda cnt: np.stack([np.cumsum(cnt) - cnt, np.cumsum(cnt)], axis=1)
        arrs = dict(prot_pos=rng.normal(size=(prot_n.sum(), 3)), prot_feat=rng.integers(0, 11, prot_n.sum()),
                    pharm_pos=rng.normal(size=(ph_n.sum(), 3)), pharm_feat=rng.integers(0, 6, ph_n.sum()),
                    prot_ph_pos=rng.normal(size=(rp_n.sum(), 3)), prot_ph_feat=rng.integers(0, 6, rp_n.sum()),
                    prot_idx=mk(prot_n), pharm_idx=mk(ph_n), prot_ph_idx=mk(rp_n))
        np.savez(d / "prot_pharm_tensors.npz", **arrs)
        if split in (0, 2):
            for i in range(n_items):
                truth.append((arrs["prot_pos"][arrs["prot_idx"][i, 0]:arrs["prot_idx"][i, 1]],
                              arrs["pharm_feat"][arrs["pharm_idx"][i, 0]:arrs["pharm_idx"][i, 1]]))
    ds = pio.ProteinPharmacophoreDataset([0, 2], tmp_path, ELEMENTS)
    assert len(ds) == 6
    for i, (ppos, pfeat) in enumerate(truth):
        it = ds[i]
        assert np.allclose(it["pocket"].prot_x.numpy(), ppos.astype(np.float32))
        assert it["h_0"].argmax(1).tolist() == pfeat.tolist() and it["pocket"].prot_h.shape[1] == 11
    sub = pio.ProteinPharmacophoreDataset([0, 2], tmp_path, ELEMENTS, subsample_pharms=True, subsample_min=4, subsample_max=8)
    for i in range(len(sub)):
        full = len(truth[i][1])
        n = sub[i]["x_0"].shape[0]
        assert (n == full) if full < 4 else (4 <= n <= min(8, full))


def test_complementarity_matches_reference(golden):
    g = golden("complementarity.npz")
    for c in range(int(g["n_cases"])):
        pt = [pio.PH_TYPES[int(i)] for i in g[f"pt{c}"]]
        rt = [pio.PH_TYPES[int(i)] for i in g[f"rt{c}"]]
        cnt = pio.compute_complementarity(pt, torch.from_numpy(g[f"ppos{c}"]), rt, torch.from_numpy(g[f"rpos{c}"]), True)
        assert int(cnt) == int(g[f"count{c}"])
    from pharmacoforge_b200.diffusion import SampledPharmacophore
    ph = SampledPharmacophore(torch.from_numpy(g["ppos0"]), torch.nn.functional.one_hot(torch.from_numpy(g["pt0"]), 6).float(),
                              pio.PH_TYPES)
    rfeat = torch.nn.functional.one_hot(torch.from_numpy(g["rt0"]), 6).float()
    v = pio.SampleAnalyzer().analyze([ph], [torch.from_numpy(g["rpos0"])], [rfeat])["validity"]
    assert abs(v - int(g["count0"]) / ph.n_ph_centers) < 1e-9


# ------------------------------------------------------------------------------------------------ checkpoint / config I/O
def _meta():
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "lightning_ckpt_meta.json")) as f:
        return json.load(f)


def _write_reference_run(tmp_path, sd, hyper_parameters, config, config_name="config.yaml"):
    """A run directory as the reference's train.py leaves it: <run>/config.yaml + <run>/checkpoints/last.ckpt, the
    checkpoint being the dict Lightning's ModelCheckpoint writes for the reference LightningModule."""
    import torch
    import yaml
    run = tmp_path / "fancy-run_abc123"
    (run / "checkpoints").mkdir(parents=True)
    with open(run / config_name, "w") as f:
        yaml.dump(config, f)
    ckpt = {"epoch": 3, "global_step": 1234, "pytorch-lightning_version": "2.0.9",
            "state_dict": {k: v.clone() for k, v in sd.items()}, "loops": {}, "callbacks": {},
            "optimizer_states": [{"state": {}, "param_groups": [{"lr": 1e-4}]}], "lr_schedulers": [{}],
            "hyper_parameters": hyper_parameters}
    torch.save(ckpt, run / "checkpoints" / "last.ckpt")
    return run


def test_load_reference_format_run_directory(tmp_path, sd):
    """generate_pharmacophores.py:231-269: config.yaml + checkpoints/last.ckpt whose `hyper_parameters` are what the
    reference's own constructor saved (golden fixture written by oracle/make_golden_ckpt.py)."""
    import torch
    from pharmacoforge_b200.checkpoint import find_run_files, load_run
    meta = _meta()
    run = _write_reference_run(tmp_path, sd, meta["hyper_parameters"], meta["config"])
    model, config = load_run(model_dir=run)
    assert not model.training and config == meta["config"]
    got = model.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    for k, v in meta["hyper_parameters"].items():
        assert model.hparams[k] == v, k
    assert model.n_timesteps == 100 and model.dynamics.pf_k == 5 and model.ph_type_map == meta["config"]["dataset"]["ph_type_map"]
    # --ckpt form: run_dir = ckpt.parent.parent
    cfg_file, model_file = find_run_files(ckpt=run / "checkpoints" / "last.ckpt")
    assert cfg_file == run / "config.yaml" and model_file == run / "checkpoints" / "last.ckpt"
    model2, _ = load_run(ckpt=run / "checkpoints" / "last.ckpt")
    assert all(torch.equal(model2.state_dict()[k], sd[k]) for k in sd)


def test_legacy_checkpoint_without_ph_type_map_retries(tmp_path, sd):
    """Checkpoints trained before `ph_type_map` became a constructor argument: load_from_checkpoint raises TypeError, the
    caller retries with the map from the config (generate_pharmacophores.py:264-268); config.yml is accepted too."""
    import pytest
    from pharmacoforge_b200.checkpoint import load_run
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    meta = _meta()
    hp = {k: v for k, v in meta["hyper_parameters"].items() if k != "ph_type_map"}
    run = _write_reference_run(tmp_path, sd, hp, meta["config"], config_name="config.yml")
    with pytest.raises(TypeError):
        PharmacophoreDiff.load_from_checkpoint(run / "checkpoints" / "last.ckpt")
    model, _ = load_run(model_dir=run)
    assert model.ph_type_map == meta["config"]["dataset"]["ph_type_map"]
    (run / "config.yml").unlink()
    with pytest.raises(FileNotFoundError):
        load_run(model_dir=run)


def test_from_config_and_run_dir_round_trip(tmp_path, sd):
    """model_from_config (load_from_config.py:6-32) on the reference's dev.yml reproduces the hyper-parameters the
    reference's constructor records; write_run_dir + save_checkpoint produce a directory load_run reads back."""
    import torch
    import yaml
    from pharmacoforge_b200.checkpoint import load_run, write_run_dir
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    meta = _meta()
    model = PharmacophoreDiff.from_config(meta["config"])
    for k, v in meta["hyper_parameters"].items():
        assert model.hparams[k] == v, k
    assert set(model.state_dict()) == set(sd)
    model.load_state_dict(sd)
    run = write_run_dir(tmp_path / "runs", meta["config"], name="brisk-sun-7", run_id="x1y2z3")
    assert run.name == "brisk-sun-7_x1y2z3" and (run / "checkpoints").is_dir()
    written = yaml.safe_load(open(run / "config.yaml"))
    assert written["resume"] == {"run_id": "x1y2z3"} and written["wandb"]["name"] == "brisk-sun-7"
    assert written["dynamics"] == meta["config"]["dynamics"]
    model.save_checkpoint(run / "checkpoints" / "last.ckpt", epoch=2, global_step=77)
    raw = torch.load(run / "checkpoints" / "last.ckpt", weights_only=False)
    assert {"state_dict", "hyper_parameters", "epoch", "global_step", "pytorch-lightning_version"} <= set(raw)
    back, _ = load_run(model_dir=run)
    assert all(torch.equal(back.state_dict()[k], sd[k]) for k in sd)


def test_write_pharmacophore_file_matches_reference_text(tmp_path):
    """utils/unorganized_utils.py:111-128 on the fixture input: same text, byte for byte."""
    import torch
    from pharmacoforge_b200.io import write_pharmacophore_file
    w = _meta()["xyz_writer"]
    coords = [torch.tensor(c) for c in w["coords"]]
    assert write_pharmacophore_file(coords, w["types"], None) == w["text"]
    write_pharmacophore_file(coords, w["types"], None, filename=tmp_path / "p.xyz")
    assert (tmp_path / "p.xyz").read_text() == w["text"]


def test_pocket_dgl_round_trip_over_the_dgl_shim(golden):
    """Pocket.to_dgl / Pocket.from_dgl against the reference's graph layout (protein_pharm_dataset.py:210-266), run over
    the pure-torch dgl stand-in of oracle/shims (dgl itself is not installable here): the pp edges equal the fixture
    written by the reference's own build_initial_complex_graph."""
    import os
    import sys
    import numpy as np
    import torch
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    try:
        from pharmacoforge_b200.batch import Pocket
        from pharmacoforge_b200.synthetic import make_pocket
        pos, onehot = make_pocket(400, seed=0)
        p = Pocket.from_numpy(pos, onehot)
        g = p.to_dgl({"pp": 3.5, "pf": 8, "fp": 8, "ff": 9})
        assert set(g.ntypes) == {"prot", "pharm", "prot_ph"} and g.num_nodes("prot") == 400 and g.num_nodes("pharm") == 0
        u, v = g.edges(form="uv", etype=("prot", "pp", "prot"))
        order = np.lexsort((u.numpy(), v.numpy()))
        gold = golden("pp_graph_n400_seed0.npz")
        assert np.array_equal(u.numpy()[order], gold["src"]) and np.array_equal(v.numpy()[order], gold["dst"])
        back = Pocket.from_dgl(g)
        assert torch.equal(back.prot_x, p.prot_x) and torch.equal(back.prot_h, p.prot_h)
    finally:
        sys.path.remove(os.path.join(ROOT, "oracle", "shims"))
        for m in [m for m in sys.modules if m == "dgl" or m.startswith("dgl.")]:
            del sys.modules[m]


def test_pdb_altloc_occupancy_hetero_and_mmcif(tmp_path):
    """The hand-written PDB / mmCIF readers follow Bio.PDB's conventions (generate_pharmacophores.py:128-135): one atom
    per (residue, name) -- the alternate location with the highest occupancy --, hetero residues keep their own residue
    id (a standard amino acid stored as HETATM still counts, water and ligands do not), .mmcif is accepted."""
    import numpy as np
    from pharmacoforge_b200.io import pocket_from_pdb, read_mmcif_atoms, read_pdb_atoms

    def atom(rec, serial, name, alt, res, chain, num, x, y, z, occ, el):
        return f"{rec:<6}{serial:>5} {name:<4}{alt}{res:>3} {chain}{num:>4}    {x:>8.3f}{y:>8.3f}{z:>8.3f}{occ:>6.2f}{20.0:>6.2f}          {el:>2}\n"
    pdb = (atom("ATOM", 1, "N", " ", "ALA", "A", 1, 0.0, 0.0, 0.0, 1.0, "N")
           + atom("ATOM", 2, "CA", "A", "ALA", "A", 1, 1.0, 0.0, 0.0, 0.3, "C")
           + atom("ATOM", 3, "CA", "B", "ALA", "A", 1, 1.5, 0.0, 0.0, 0.7, "C")
           + atom("HETATM", 4, "CA", " ", "GLY", "A", 2, 3.0, 0.0, 0.0, 1.0, "C")
           + atom("HETATM", 5, "O", " ", "HOH", "A", 3, 4.0, 0.0, 0.0, 1.0, "O")
           + atom("ATOM", 6, "CA", " ", "LEU", "A", 9, 60.0, 0.0, 0.0, 1.0, "C"))
    f = tmp_path / "rec.pdb"
    f.write_text(pdb)
    atoms = read_pdb_atoms(f)
    assert [a["name"] for a in atoms] == ["N", "CA", "CA", "O", "CA"] and atoms[1]["altloc"] == "B"
    els = ["C", "N", "O", "S"]
    pocket, com = pocket_from_pdb(f, els, pocket_cutoff=8.0, lig_coords=np.array([[2.0, 0.0, 0.0]]))
    assert pocket.prot_x.shape[0] == 3                      # ALA (N, CA alt B) + the HETATM glycine; water and far LEU out
    assert np.allclose(pocket.prot_x.numpy()[1], [1.5, 0.0, 0.0])
    # the same structure as mmCIF
    head = ["group_PDB", "id", "type_symbol", "label_atom_id", "label_alt_id", "label_comp_id", "label_asym_id",
            "label_seq_id", "pdbx_PDB_ins_code", "Cartn_x", "Cartn_y", "Cartn_z", "occupancy", "auth_seq_id",
            "auth_asym_id", "pdbx_PDB_model_num"]
    rows = ["ATOM 1 N N . ALA A 1 ? 0.0 0.0 0.0 1.0 1 A 1", "ATOM 2 C CA A ALA A 1 ? 1.0 0.0 0.0 0.3 1 A 1",
            "ATOM 3 C CA B ALA A 1 ? 1.5 0.0 0.0 0.7 1 A 1", "HETATM 4 C CA . GLY A 2 ? 3.0 0.0 0.0 1.0 2 A 1",
            "HETATM 5 O O . HOH A 3 ? 4.0 0.0 0.0 1.0 3 A 1", "ATOM 6 C CA . LEU A 9 ? 60.0 0.0 0.0 1.0 9 A 1"]
    cif = tmp_path / "rec.mmcif"
    cif.write_text("data_x\nloop_\n" + "".join(f"_atom_site.{h}\n" for h in head) + "\n".join(rows) + "\n#\n")
    assert [(a["name"], a["altloc"], a["het"]) for a in read_mmcif_atoms(cif)] == \
        [(a["name"], a["altloc"], a["het"]) for a in atoms]
    pocket2, _ = pocket_from_pdb(cif, els, pocket_cutoff=8.0, lig_coords=np.array([[2.0, 0.0, 0.0]]))
    assert np.array_equal(pocket2.prot_x.numpy(), pocket.prot_x.numpy())
