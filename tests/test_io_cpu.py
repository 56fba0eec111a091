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
