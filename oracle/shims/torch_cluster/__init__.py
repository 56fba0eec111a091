"""Pure-PyTorch restatement of the torch_cluster calls on PharmacoForge's hot path.

TEST INFRASTRUCTURE ONLY.  torch_cluster (conda `pyg` channel, unpinned in the
reference's install_things.sh:5) is not installable here.  Call sites restated:
protein_pharm_dataset.py:235 (radius_graph), dynamics_gvp.py:194,196,202,211
(knn_graph, radius_graph, knn, radius).  Semantics follow the published
torch_cluster CUDA kernels:

  * radius(x, y, r, batch_x, batch_y, max_num_neighbors) -> [2, E], row 0 =
    index into y (query), row 1 = index into x; a pair is kept iff it is in the
    same batch element and the squared distance is STRICTLY below r*r; per query
    the first `max_num_neighbors` hits in ascending x index are kept.
  * knn(x, y, k, batch_x, batch_y) -> row 0 = y (query) index, row 1 = x index;
    the k smallest squared distances in ascending order, ties keep the lower x
    index; fewer than k candidates -> fewer rows.
  * radius_graph / knn_graph: the same with x == y, returned as
    (source = neighbour, target = centre), self pairs removed.

Canonical squared distance (so that CPU oracle and CUDA kernels agree bit for
bit at the threshold and in kNN order): ((dx*dx + dy*dy) + dz*dz) in fp32 with
separately rounded multiplies and adds (no FMA contraction).
"""
import torch


def _sqdist(q, c):
    # q [Q,3], c [C,3] -> [Q,C]; explicit op order, each op rounded to fp32
    dx = q[:, None, 0] - c[None, :, 0]
    dy = q[:, None, 1] - c[None, :, 1]
    dz = q[:, None, 2] - c[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def _segments(batch, n, device):
    if batch is None:
        return [(0, n)]
    if n == 0:
        return []
    nb = int(batch.max()) + 1
    counts = torch.bincount(batch, minlength=nb)
    ends = torch.cumsum(counts, 0)
    starts = ends - counts
    return [(int(s), int(e)) for s, e in zip(starts.tolist(), ends.tolist())]


def _pair_segments(x, y, batch_x, batch_y):
    sx = _segments(batch_x, x.shape[0], x.device)
    sy = _segments(batch_y, y.shape[0], y.device)
    nb = max(len(sx), len(sy))
    sx += [(x.shape[0], x.shape[0])] * (nb - len(sx))
    sy += [(y.shape[0], y.shape[0])] * (nb - len(sy))
    return zip(sx, sy)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1, batch_size=None):
    rows, cols = [], []
    r2 = torch.tensor(float(r), dtype=x.dtype) * torch.tensor(float(r), dtype=x.dtype)
    for (xs, xe), (ys, ye) in _pair_segments(x, y, batch_x, batch_y):
        if xe == xs or ye == ys:
            continue
        d = _sqdist(y[ys:ye], x[xs:xe])
        hit = d < r2
        rank = torch.cumsum(hit.to(torch.int64), dim=1)
        hit = hit & (rank <= max_num_neighbors)
        qi, ci = torch.nonzero(hit, as_tuple=True)  # row-major: sorted by (query, candidate)
        rows.append(qi + ys)
        cols.append(ci + xs)
    if not rows:
        return torch.zeros((2, 0), dtype=torch.int64, device=x.device)
    return torch.stack([torch.cat(rows), torch.cat(cols)], dim=0)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target",
                 num_workers=1, batch_size=None):
    ei = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    if flow == "source_to_target":
        row, col = ei[1], ei[0]
    else:
        row, col = ei[0], ei[1]
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col], dim=0)


def knn(x, y, k, batch_x=None, batch_y=None, cosine=False, num_workers=1, batch_size=None):
    assert not cosine
    rows, cols = [], []
    for (xs, xe), (ys, ye) in _pair_segments(x, y, batch_x, batch_y):
        if xe == xs or ye == ys:
            continue
        d = _sqdist(y[ys:ye], x[xs:xe])
        kk = min(k, xe - xs)
        order = torch.sort(d, dim=1, stable=True).indices[:, :kk]  # stable: ties keep lower index
        q = torch.arange(ys, ye, device=x.device)[:, None].expand(-1, kk)
        rows.append(q.reshape(-1))
        cols.append(order.reshape(-1) + xs)
    if not rows:
        return torch.zeros((2, 0), dtype=torch.int64, device=x.device)
    return torch.stack([torch.cat(rows), torch.cat(cols)], dim=0)


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", cosine=False, num_workers=1,
              batch_size=None):
    ei = knn(x, x, k if loop else k + 1, batch, batch, cosine)
    if flow == "source_to_target":
        row, col = ei[1], ei[0]
    else:
        row, col = ei[0], ei[1]
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col], dim=0)
