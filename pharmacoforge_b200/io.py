"""Host-side front end and back end around the hot path -- plain Python / numpy, no
BioPython, RDKit or DGL:

  * `pocket_from_pdb`: the pocket extraction of generate_pharmacophores.py:120-220 (`process_ligand_and_pocket`): the
    standard-amino-acid residues with an atom within `pocket_cutoff` of the reference ligand (or an explicit
    `chain:resnum` list), heavy atoms only, element one-hot over `prot_elements` with the 'other' column dropped and
    those atoms removed.  PDB / mmCIF atom records (alternate locations resolved by occupancy, hetero residues kept
    under their own residue ids, as Bio.PDB does) and V2000 SDF coordinates are parsed by hand.
  * `ProteinPharmacophoreDataset`: the processed CrossDocked tensors (`prot_pharm_tensors.npz`,
    protein_pharm_dataset.py:18-207) as (Pocket, pharmacophore x_0 / h_0, receptor pharmacophores) items, with the
    reference's pharmacophore subsampling, plus `collate` to a training `GraphBatch`.
  * `compute_complementarity` / `SampleAnalyzer`: the validity metric of analysis/metrics.py:9-86.
"""
from __future__ import annotations

import random
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .batch import Pocket

STANDARD_AA = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO",
               "SER", "THR", "TRP", "TYR", "VAL"}
PH_TYPES = ["Aromatic", "HydrogenDonor", "HydrogenAcceptor", "PositiveIon", "NegativeIon", "Hydrophobic"]  # constants.py


def element_fixer(element: str) -> str:
    """generate_pharmacophores.py:97-102: 'CL' -> 'Cl'."""
    return element[0] + element[1:].lower() if len(element) > 1 else element


def read_pdb_atoms(path) -> List[dict]:
    """ATOM / HETATM records of every model, as Bio.PDB.PDBParser presents them (generate_pharmacophores.py:128-135):
    chain, residue id (hetero flag, resseq, icode), resname, atom name, element, xyz, model index.  Alternate
    locations collapse to ONE atom per (residue, atom name): the location with the highest occupancy, the first one on
    ties -- BioPython's DisorderedAtom selection -- at the position of its first record."""
    atoms: List[dict] = []
    slot: Dict[tuple, int] = {}
    model = 0
    for line in Path(path).read_text().splitlines():
        rec = line[:6]
        if rec.startswith("ENDMDL"):
            model += 1
            continue
        if rec not in ("ATOM  ", "HETATM"):
            continue
        name = line[12:16].strip()
        resname = line[17:20].strip()
        element = line[76:78].strip() if len(line) >= 78 and line[76:78].strip() else "".join(c for c in name if c.isalpha())[:1]
        het = " " if rec == "ATOM  " else ("W" if resname in ("HOH", "WAT") else "H_" + resname)
        try:
            occ = float(line[54:60])
        except ValueError:
            occ = 1.0
        a = dict(het=het, name=name, altloc=line[16], resname=resname, chain=line[21], resseq=int(line[22:26]),
                 icode=line[26], element=element_fixer(element.upper()), occupancy=occ, model=model,
                 xyz=(float(line[30:38]), float(line[38:46]), float(line[46:54])))
        key = (model, a["chain"], het, a["resseq"], a["icode"], name)
        if key in slot:
            if a["altloc"] != " " and occ > atoms[slot[key]]["occupancy"]:
                atoms[slot[key]] = a
            continue
        slot[key] = len(atoms)
        atoms.append(a)
    return atoms


def read_mmcif_atoms(path) -> List[dict]:
    """The `_atom_site` loop of an mmCIF file (generate_pharmacophores.py:130-131 uses Bio.PDB.MMCIFParser): same records as
    `read_pdb_atoms`, author chain / residue numbering, alternate locations resolved the same way."""
    lines = Path(path).read_text().splitlines()
    cols: List[str] = []
    rows: List[List[str]] = []
    i = 0
    while i < len(lines):
        if lines[i].strip() == "loop_" and i + 1 < len(lines) and lines[i + 1].strip().startswith("_atom_site."):
            i += 1
            while i < len(lines) and lines[i].strip().startswith("_atom_site."):
                cols.append(lines[i].strip().split(".", 1)[1])
                i += 1
            while i < len(lines) and lines[i].strip() and not lines[i].startswith(("#", "loop_", "_")):
                rows.append(lines[i].split())
                i += 1
            break
        i += 1
    if not cols:
        raise ValueError(f"no _atom_site loop in {path}")
    c = {k: j for j, k in enumerate(cols)}
    get = lambda r, k, d="": r[c[k]].strip('"\'') if k in c else d
    atoms: List[dict] = []
    slot: Dict[tuple, int] = {}
    models: Dict[str, int] = {}
    for r in rows:
        model = models.setdefault(get(r, "pdbx_PDB_model_num", "1"), len(models))
        resname = get(r, "auth_comp_id") or get(r, "label_comp_id")
        name = get(r, "auth_atom_id") or get(r, "label_atom_id")
        chain = get(r, "auth_asym_id") or get(r, "label_asym_id")
        resseq = int(get(r, "auth_seq_id") or get(r, "label_seq_id"))
        icode = get(r, "pdbx_PDB_ins_code", "?")
        icode = " " if icode in ("?", ".") else icode
        alt = get(r, "label_alt_id", ".")
        alt = " " if alt in ("?", ".") else alt
        het = " " if get(r, "group_PDB", "ATOM") == "ATOM" else ("W" if resname in ("HOH", "WAT") else "H_" + resname)
        occ = float(get(r, "occupancy", "1.0"))
        a = dict(het=het, name=name, altloc=alt, resname=resname, chain=chain, resseq=resseq, icode=icode,
                 element=element_fixer(get(r, "type_symbol").upper()), occupancy=occ, model=model,
                 xyz=(float(get(r, "Cartn_x")), float(get(r, "Cartn_y")), float(get(r, "Cartn_z"))))
        key = (model, chain, het, resseq, icode, name)
        if key in slot:
            if alt != " " and occ > atoms[slot[key]]["occupancy"]:
                atoms[slot[key]] = a
            continue
        slot[key] = len(atoms)
        atoms.append(a)
    return atoms


def read_sdf_coords(path, remove_hydrogen: bool = True) -> np.ndarray:
    """Atom coordinates of the first V2000 mol block of an SDF file (generate_pharmacophores.py:68-95 uses RDKit)."""
    lines = Path(path).read_text().splitlines()
    n_atoms = int(lines[3][:3])
    out = []
    for line in lines[4:4 + n_atoms]:
        x, y, z, el = float(line[0:10]), float(line[10:20]), float(line[20:30]), line[31:34].strip()
        if remove_hydrogen and el == "H":
            continue
        out.append((x, y, z))
    return np.asarray(out, dtype=np.float32).reshape(-1, 3)


def onehot_encode_elements(elements: Iterable[str], element_map: Dict[str, int]) -> np.ndarray:
    """generate_pharmacophores.py:104-117."""
    idx = np.fromiter((element_map.get(e, element_map["other"]) for e in elements), int)
    out = np.zeros((idx.size, len(element_map)), dtype=np.float32)
    out[np.arange(idx.size), idx] = 1
    return out


def pocket_from_pdb(rec_file, prot_elements: Sequence[str], pocket_cutoff: float = 8.0, lig_file=None,
                    lig_coords: Optional[np.ndarray] = None, residue_list: Sequence[str] = (),
                    remove_hydrogen: bool = True) -> Tuple[Pocket, torch.Tensor]:
    """-> (Pocket, init_com [1, 3]): what `process_ligand_and_pocket` feeds `build_initial_complex_graph`
    (generate_pharmacophores.py:120-205).  Ligand given as an SDF file or raw coordinates, or a residue list
    ['A:101', ...]; init_com is the ligand COM (or the COM of the listed residues' atoms)."""
    if lig_file is None and lig_coords is None and len(residue_list) == 0:
        raise ValueError("Either reference ligand or pocket residue list must be provided.")
    suffix = Path(rec_file).suffix
    if suffix == ".pdb":
        atoms = read_pdb_atoms(rec_file)
    elif suffix == ".mmcif":
        atoms = read_mmcif_atoms(rec_file)
    else:
        raise ValueError(f"unsupported receptor file type: {suffix}, must be .pdb or .mmcif")
    use_ligand = lig_file is not None or lig_coords is not None
    if not use_ligand:
        atoms = [a for a in atoms if a["model"] == 0]   # the residue list indexes rec_struct[0] (:165-166)
    residues: Dict[tuple, List[dict]] = {}
    for a in atoms:                                   # insertion-ordered, like BioPython's get_residues() (all models)
        residues.setdefault((a["model"], a["chain"], a["het"], a["resseq"], a["icode"]), []).append(a)
    if use_ligand:
        lig = read_sdf_coords(lig_file, remove_hydrogen) if lig_coords is None else np.asarray(lig_coords, dtype=np.float32)
        init_com = torch.from_numpy(lig.mean(axis=0).reshape(1, 3).astype(np.float32))
        chosen = []
        for key, res in residues.items():
            if res[0]["resname"] not in STANDARD_AA:
                continue
            xyz = np.asarray([a["xyz"] for a in res], dtype=np.float32)
            d = np.sqrt(((lig[:, None, :] - xyz[None, :, :]) ** 2).sum(-1)).min()
            if d < pocket_cutoff:
                chosen.append(key)
        if not chosen:
            raise ValueError("no valid pocket residues found.")
    else:
        chosen = []
        for spec in residue_list:
            chain, num = spec.split(":")
            key = (0, chain, " ", int(num), " ")
            if key not in residues:
                raise KeyError(f"residue {spec} not found in {rec_file}")
            chosen.append(key)
        xyz = np.asarray([a["xyz"] for k in chosen for a in residues[k]], dtype=np.float32)
        init_com = torch.from_numpy(xyz.mean(axis=0).reshape(1, 3))
    pocket_atoms = [a for k in chosen for a in residues[k] if not (remove_hydrogen and a["element"] == "H")]
    element_map = {e: i for i, e in enumerate(list(prot_elements) + ["other"])}
    onehot = onehot_encode_elements([a["element"] for a in pocket_atoms], element_map)
    keep = onehot[:, -1] != 1                          # drop 'other' atoms (generate_pharmacophores.py:189-196)
    pos = np.asarray([a["xyz"] for a in pocket_atoms], dtype=np.float32)[keep]
    return Pocket.from_numpy(pos, onehot[keep, :-1]), init_com


TYPE_IDX_TO_ELEM = ["P", "S", "F", "N", "O", "C"]


def write_pharmacophore_file(coords_list, atom_types_list, pharm_type_map=None, filename=None):
    """utils/unorganized_utils.py:111-128: several pharmacophores as concatenated xyz blocks (`n` line, then one
    `<element> x y z` line per centre, element = TYPE_IDX_TO_ELEM[type index]); returns the text when no filename."""
    out = ""
    for coords, atom_types in zip(coords_list, atom_types_list):
        assert len(coords) == len(atom_types)
        elems = [TYPE_IDX_TO_ELEM[int(i)] for i in atom_types]
        out += f"{len(coords)}\n"
        for i in range(len(coords)):
            out += f"{elems[i]} {coords[i, 0]:.3f} {coords[i, 1]:.3f} {coords[i, 2]:.3f}\n"
    if filename is None:
        return out
    Path(filename).write_text(out)


def pocket_to_dgl(pocket: Pocket, graph_cutoffs: dict, pharm_x: Optional[torch.Tensor] = None,
                  pharm_h: Optional[torch.Tensor] = None):
    """A reference-style pocket graph (`build_initial_complex_graph`, protein_pharm_dataset.py:210-266): node types
    prot / pharm / prot_ph, the pp radius edges (src = neighbour, dst = centre, ascending), node data x_0 / h_0.
    Needs dgl (and torch_cluster for nothing: the pp edges are computed here).  Inverse of `Pocket.from_dgl`."""
    import dgl
    x = pocket.prot_x.float()
    n = x.shape[0]
    r = float(graph_cutoffs["pp"])
    src = torch.zeros(0, dtype=torch.long)
    dst = torch.zeros(0, dtype=torch.long)
    if r > 0 and n > 0:
        d = x[:, None, :] - x[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        hit = d2 < r * r                                    # [centre, neighbour], self included (torch_cluster rule)
        rank = torch.cumsum(hit.long(), dim=1)
        keep = hit & (rank <= 100 + 1) & ~torch.eye(n, dtype=torch.bool)
        dst, src = keep.nonzero(as_tuple=True)
    n_pharm = 0 if pharm_x is None else pharm_x.shape[0]
    g = dgl.heterograph({("prot", "pp", "prot"): (src, dst), ("prot", "pf", "pharm"): ([], []),
                         ("pharm", "ff", "pharm"): ([], []), ("pharm", "fp", "prot"): ([], [])},
                        num_nodes_dict={"prot": n, "pharm": n_pharm, "prot_ph": 0})
    if pharm_x is not None:
        g.nodes["pharm"].data["x_0"] = pharm_x
        g.nodes["pharm"].data["h_0"] = pharm_h
    g.nodes["prot"].data["x_0"] = pocket.prot_x
    g.nodes["prot"].data["h_0"] = pocket.prot_h
    return g


class ProteinPharmacophoreDataset:
    """protein_pharm_dataset.py:18-207 without DGL: item i -> dict(pocket, x_0, h_0, prot_ph_pos, prot_ph_feat).
    `processed_data_dir` holds one sub-directory per split, each with `prot_pharm_tensors.npz` (keys pharm_pos,
    pharm_feat, prot_pos, prot_feat, prot_ph_pos, prot_ph_feat and the [n, 2] start / end index arrays pharm_idx,
    prot_idx, prot_ph_idx)."""

    def __init__(self, split_idxs: Sequence[int], processed_data_dir, prot_elements: Sequence[str],
                 ph_type_map: Sequence[str] = PH_TYPES, subsample_pharms: bool = False, subsample_min: int = 3,
                 subsample_max: int = 9, **kwargs):
        self.prot_elements, self.ph_type_map = list(prot_elements), list(ph_type_map)
        self.subsample_pharms, self.subsample_min, self.subsample_max = subsample_pharms, subsample_min, subsample_max
        root = Path(processed_data_dir)
        if not root.exists():
            raise FileNotFoundError(f"Could not find processed data directory at {root}")
        keys = ("pharm_pos", "pharm_feat", "prot_pos", "prot_feat", "prot_ph_pos", "prot_ph_feat")
        data = {k: [] for k in keys}
        idx = {"pharm_idx": [], "prot_idx": [], "prot_ph_idx": []}
        for split_dir in sorted(root.iterdir()):
            if not split_dir.is_dir() or int(split_dir.name.split("_")[-1][-1]) not in split_idxs:
                continue
            z = np.load(split_dir / "prot_pharm_tensors.npz")
            for k in keys:
                data[k].append(z[k])
            for k in idx:                              # make the per-split [start, end) indices global (:103-121)
                off = idx[k][-1][-1, 1] if idx[k] else 0
                idx[k].append(z[k] + off)
        if not data["prot_pos"]:
            raise FileNotFoundError(f"no split directories for splits {list(split_idxs)} under {root}")
        for k in keys:
            setattr(self, k, torch.from_numpy(np.concatenate(data[k], axis=0)))
        for k in idx:
            setattr(self, k, np.concatenate(idx[k], axis=0))

    def __len__(self):
        return self.prot_idx.shape[0]

    def __getitem__(self, i):
        f0, f1 = self.pharm_idx[i]
        p0, p1 = self.prot_idx[i]
        r0, r1 = self.prot_ph_idx[i]
        one_hot = torch.nn.functional.one_hot
        pharm_pos = self.pharm_pos[f0:f1].float()
        pharm_feat = one_hot(self.pharm_feat[f0:f1].long(), len(self.ph_type_map)).float()
        if self.subsample_pharms and len(pharm_pos) > self.subsample_min - 1:     # :159-168
            hi = min(self.subsample_max, len(pharm_pos))
            n = self.subsample_min if self.subsample_min == hi else random.randint(self.subsample_min, hi)
            sel = random.sample(range(len(pharm_pos)), n)
            pharm_pos, pharm_feat = pharm_pos[sel], pharm_feat[sel]
        pocket = Pocket(self.prot_pos[p0:p1].float().contiguous(),
                        one_hot(self.prot_feat[p0:p1].long(), len(self.prot_elements)).float())
        return dict(pocket=pocket, x_0=pharm_pos, h_0=pharm_feat, prot_ph_pos=self.prot_ph_pos[r0:r1].float(),
                    prot_ph_feat=one_hot(self.prot_ph_feat[r0:r1].long(), len(self.ph_type_map)).float())

    @staticmethod
    def collate(items: Sequence[dict], model, device=None):
        """One training graph per item -> GraphBatch with the ground truth attached (dgl.batch of the item graphs)."""
        g = model.make_batch([it["pocket"] for it in items], [[int(it["x_0"].shape[0])] for it in items], device=device)
        return g.set_pharmacophores(torch.cat([it["x_0"] for it in items]), torch.cat([it["h_0"] for it in items]))


MATCHING_TYPES = {"Aromatic": ["Aromatic", "PositiveIon"], "HydrogenDonor": ["HydrogenAcceptor"],
                  "HydrogenAcceptor": ["HydrogenDonor"], "PositiveIon": ["NegativeIon", "Aromatic"],
                  "NegativeIon": ["PositiveIon"], "Hydrophobic": ["Hydrophobic"]}
MATCHING_DISTANCE = {"Aromatic": 7, "Hydrophobic": 5, "HydrogenAcceptor": 4, "HydrogenDonor": 4, "NegativeIon": 5,
                     "PositiveIon": 5}


def compute_complementarity(pharm_types: Sequence[str], pharm_pos: torch.Tensor, prot_ph_types: Sequence[str],
                            prot_ph_pos: torch.Tensor, return_count: bool = False):
    """analysis/metrics.py:54-86: pharmacophore centres within the type's matching distance of a complementary
    receptor feature (count, or fraction of the centres)."""
    dist = torch.cdist(pharm_pos.float(), prot_ph_pos.float())
    cut = torch.tensor([MATCHING_DISTANCE[t] for t in pharm_types], dtype=dist.dtype, device=dist.device).reshape(-1, 1)
    match = torch.tensor([[r in MATCHING_TYPES[t] for r in prot_ph_types] for t in pharm_types], dtype=torch.bool,
                         device=dist.device).reshape(len(pharm_types), len(prot_ph_types))
    count = ((dist <= cut) & match).any(dim=1).sum()
    return count if return_count else count / max(len(pharm_types), 1)


class SampleAnalyzer:
    """analysis/metrics.py:7-35: validity = complementary centres / all centres over a list of sampled pharmacophores,
    each paired with the receptor pharmacophore features of its pocket."""

    def analyze(self, samples, prot_ph_pos: Sequence[torch.Tensor], prot_ph_feat: Sequence[torch.Tensor]):
        num = den = 0
        for ph, rpos, rfeat in zip(samples, prot_ph_pos, prot_ph_feat):
            rtypes = [PH_TYPES[int(i)] for i in rfeat.argmax(dim=1)]
            num += int(compute_complementarity(ph.ph_types, ph.ph_coords, rtypes, rpos, return_count=True))
            den += ph.n_ph_centers
        return {"validity": num / max(den, 1)}
