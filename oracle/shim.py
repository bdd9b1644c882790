"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (dig_b200/).

Pure-torch stand-ins for the four third-party extension packages that the
reference's 3D-graph path calls but that are neither vendored under
/root/reference nor installable here (no network):

    torch-cluster  1.6.0   radius_graph          (pin: docs/environment.yaml:15-19)
    torch-scatter  2.0.9   scatter, scatter_min
    torch-sparse   0.6.13  SparseTensor, matmul
    torch-geometric 2.1.0  nn.{radius_graph,GraphConv,GraphNorm,MessagePassing,inits}, data.*

`install()` registers them in `sys.modules` so that the reference's own model
code (dig/threedgraph/method/*.py, dig/threedgraph/utils/geometric_computing.py)
can be imported and executed VERBATIM.  Each function restates the published
semantics of the pinned upstream version; call sites in the reference are cited.

Parity note: the reference's tests pin nothing at this boundary;
the only known-answer is examples/threedgraph/xyz_to_dat.ipynb (SparseTensor
row-select), which tests/test_oracle.py checks.
"""
import math
import sys
import types

import torch


# --------------------------------------------------------------------------
# torch_scatter
# --------------------------------------------------------------------------
def _expand_index(index, src, dim):
    if index.dim() == src.dim():
        return index
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    """torch_scatter.scatter.  Call sites: spherenet.py:171,211,224; schnet.py:55,81;
    dimenetpp.py:150,190,203; comenet.py:398; utils/geometric_computing.py:75.
    dim_size=None => index.max()+1; empty segments are zero-filled (also for min)."""
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    idx = _expand_index(index, src, dim)
    if reduce in ("sum", "add"):
        res = torch.zeros(shape, dtype=src.dtype, device=src.device)
        res = res.scatter_add(dim, idx, src)
    elif reduce == "mean":
        res = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(dim, idx, src)
        cnt = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(
            dim, idx, torch.ones_like(src))
        res = res / cnt.clamp(min=1)
    elif reduce == "min":
        res = torch.full(shape, float("inf"), dtype=src.dtype, device=src.device)
        res = res.scatter_reduce(dim, idx, src, reduce="amin", include_self=True)
        res = torch.where(torch.isinf(res), torch.zeros_like(res), res)
    elif reduce == "max":
        res = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
        res = res.scatter_reduce(dim, idx, src, reduce="amax", include_self=True)
        res = torch.where(torch.isinf(res), torch.zeros_like(res), res)
    else:
        raise ValueError(reduce)
    if out is not None:
        out.copy_(res)
        return out
    return res


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    """torch_scatter.scatter_min -> (out, argmin).  Call sites comenet.py:304,311,316,325.
    argmin == src.size(dim) for empty segments (relied on at comenet.py:305);
    first occurrence wins ties (torch_scatter CPU semantics)."""
    assert src.dim() == 1 and (dim in (-1, 0))
    n = src.numel()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if n > 0 else 0
    val = torch.full((dim_size,), float("inf"), dtype=src.dtype, device=src.device)
    val = val.scatter_reduce(0, index, src, reduce="amin", include_self=True)
    pos = torch.arange(n, device=src.device)
    is_min = src == val[index]
    cand = torch.where(is_min, pos, torch.full_like(pos, n))
    arg = torch.full((dim_size,), n, dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(0, index, cand, reduce="amin", include_self=True)
    val = torch.where(arg == n, torch.zeros_like(val), val)
    return val, arg


# --------------------------------------------------------------------------
# torch_sparse
# --------------------------------------------------------------------------
class _Storage:
    def __init__(self, row, col, value):
        self._row, self._col, self._value = row, col, value

    def row(self):
        return self._row

    def col(self):
        return self._col

    def value(self):
        return self._value


class SparseTensor:
    """torch_sparse.SparseTensor, the subset used at utils/geometric_computing.py:27-41,54:
    construction from COO (sorted by (row, col), stable, no dedup), row index_select via
    `adj[idx]`, `.set_value(None).sum(dim=1)` (= row counts), `.storage.{row,col,value}()`."""

    def __init__(self, row=None, col=None, value=None, sparse_sizes=None, is_sorted=False):
        if not is_sorted:
            key = row * sparse_sizes[1] + col
            perm = torch.argsort(key, stable=True)
            row, col = row[perm], col[perm]
            value = value[perm] if value is not None else None
        self.storage = _Storage(row, col, value)
        self._sizes = tuple(sparse_sizes)

    def sparse_sizes(self):
        return self._sizes

    def set_value(self, value, layout=None):
        return SparseTensor(self.storage.row(), self.storage.col(), value, self._sizes, is_sorted=True)

    def _rowptr(self):
        cnt = torch.bincount(self.storage.row(), minlength=self._sizes[0])
        ptr = torch.zeros(self._sizes[0] + 1, dtype=torch.long, device=cnt.device)
        ptr[1:] = torch.cumsum(cnt, 0)
        return ptr, cnt

    def __getitem__(self, idx):
        assert isinstance(idx, torch.Tensor) and idx.dtype == torch.long and idx.dim() == 1
        ptr, cnt = self._rowptr()
        n_per = cnt[idx]
        new_row = torch.repeat_interleave(torch.arange(idx.numel(), device=idx.device), n_per)
        start = ptr[idx]
        out_ptr = torch.zeros(idx.numel() + 1, dtype=torch.long, device=idx.device)
        out_ptr[1:] = torch.cumsum(n_per, 0)
        within = torch.arange(new_row.numel(), device=idx.device) - out_ptr[:-1][new_row]
        src = start[new_row] + within
        col = self.storage.col()[src]
        val = self.storage.value()
        val = val[src] if val is not None else None
        return SparseTensor(new_row, col, val, (idx.numel(), self._sizes[1]), is_sorted=True)

    def to_dense(self):
        """Dense [rows, cols] matrix of the values (duplicates add up, like torch_sparse)."""
        val = self.storage.value()
        if val is None:
            val = torch.ones_like(self.storage.row(), dtype=torch.float)
        out = torch.zeros(self._sizes, dtype=val.dtype, device=val.device)
        return out.index_put_((self.storage.row(), self.storage.col()), val, accumulate=True)

    @classmethod
    def from_dense(cls, mat, has_value=True):
        """torch_sparse.SparseTensor.from_dense: the non-zero entries in row-major order."""
        idx = mat.nonzero()
        row, col = idx[:, 0], idx[:, 1]
        return cls(row, col, mat[row, col] if has_value else None, tuple(mat.shape), is_sorted=True)

    def sum(self, dim=None):
        assert dim == 1
        val = self.storage.value()
        if val is None:
            return torch.bincount(self.storage.row(), minlength=self._sizes[0]).to(torch.float)
        out = torch.zeros(self._sizes[0], dtype=val.dtype, device=val.device)
        return out.index_add_(0, self.storage.row(), val)


def _sparse_matmul(src, other, reduce="sum"):  # only so that pronet.py imports
    raise NotImplementedError("torch_sparse.matmul is outside the oracle's scope")


# --------------------------------------------------------------------------
# torch_cluster
# --------------------------------------------------------------------------
def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32,
                 flow="source_to_target", num_workers=1):
    """torch_cluster.radius_graph with the CUDA kernel's ordering:
    one query per node, candidates scanned in ascending index inside the query's graph,
    strict `d2 < r*r`, the first (max_num_neighbors + 1) hits kept (self included in the
    count), self loops removed afterwards.  Result: [2, E] = (source j, target i), sorted
    by (i, j).  d2 is accumulated as fma(diff, diff, acc) over the coordinates, the form
    nvcc contracts the upstream loop to; emulated here in float64.
    Call sites: schnet.py:156, dimenetpp.py:277, spherenet.py:304, comenet.py:294."""
    assert flow == "source_to_target"
    n = x.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long, device=x.device)
    cap = max_num_neighbors if loop else max_num_neighbors + 1
    xd = x.detach().to(torch.float32)
    rr = torch.tensor(float(r) * float(r), dtype=torch.float32, device=x.device)
    rows, cols = [], []
    chunk = 2048
    for s in range(0, n, chunk):
        q = xd[s:s + chunk]
        acc = torch.zeros(q.size(0), n, dtype=torch.float32, device=x.device)
        for d in range(x.size(1)):
            diff = (xd[None, :, d] - q[:, None, d])                      # fp32, exact op
            acc = (diff.double() * diff.double() + acc.double()).float()  # fma emulation
        hit = (acc < rr) & (batch[None, :] == batch[s:s + chunk, None])
        rank = torch.cumsum(hit.to(torch.int32), dim=1)
        hit = hit & (rank <= cap)
        qi, xi = torch.nonzero(hit, as_tuple=True)
        rows.append(qi + s)
        cols.append(xi)
    query = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.long, device=x.device)
    neigh = torch.cat(cols) if cols else torch.zeros(0, dtype=torch.long, device=x.device)
    if not loop:
        keep = query != neigh
        query, neigh = query[keep], neigh[keep]
    return torch.stack([neigh, query], dim=0)


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", cosine=False, num_workers=1):
    """torch_cluster.knn_graph with the CUDA kernel's semantics: every node queries the nodes of its own graph in
    ascending index, keeps the (k + 1 when loop=False) smallest squared distances by strict-`>` insertion (ties: the
    lower index stays in front), then the self pair is dropped.  Returns [2, n*k] = (source = neighbour, target =
    query), grouped by query, nearest first.  Called at ggraph3D/.../geometric_computing.py:14,16."""
    assert flow == "source_to_target" and not cosine
    n = x.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long, device=x.device)
    kk = k if loop else k + 1
    xd = x.detach().to(torch.float32)
    acc = torch.zeros(n, n, dtype=torch.float32, device=x.device)
    for d in range(x.size(1)):
        diff = xd[None, :, d] - xd[:, None, d]
        acc = (diff.double() * diff.double() + acc.double()).float()          # fma emulation, as in radius_graph
    acc = torch.where(batch[None, :] == batch[:, None], acc, torch.full_like(acc, float("inf")))
    order = torch.sort(acc, dim=1, stable=True).indices[:, :kk]               # stable: the lower index wins ties
    valid = torch.gather(acc, 1, order) < float("inf")
    query = torch.arange(n, device=x.device)[:, None].expand_as(order)
    keep = valid if loop else (valid & (order != query))
    return torch.stack([order[keep], query[keep]], dim=0)


# --------------------------------------------------------------------------
# torch_geometric
# --------------------------------------------------------------------------
def glorot_orthogonal(tensor, scale):
    """torch_geometric.nn.inits.glorot_orthogonal (spherenet.py:44-47,126-148)."""
    if tensor is not None:
        torch.nn.init.orthogonal_(tensor.data)
        scale /= ((tensor.size(-2) + tensor.size(-1)) * tensor.var())
        tensor.data *= scale.sqrt()


def glorot(tensor):
    if tensor is not None:
        stdv = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-stdv, stdv)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)


def ones(tensor):
    if tensor is not None:
        tensor.data.fill_(1)


def uniform(size, tensor):
    if tensor is not None:
        bound = 1.0 / math.sqrt(size)
        tensor.data.uniform_(-bound, bound)


def kaiming_uniform(tensor, fan, a):
    if tensor is not None:
        bound = math.sqrt(6 / ((1 + a ** 2) * fan))
        tensor.data.uniform_(-bound, bound)


class _PygLinear(torch.nn.Module):
    """torch_geometric.nn.dense.linear.Linear as used inside GraphConv (PyG 2.1.0):
    weight_initializer=None -> kaiming_uniform(fan=in, a=sqrt(5)); bias_initializer=None ->
    uniform(in)."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        kaiming_uniform(self.weight, fan=self.in_channels, a=math.sqrt(5))
        uniform(self.in_channels, self.bias)

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight, self.bias)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        """torch_geometric 2.1 MessagePassing.propagate for a dense edge_index, flow source_to_target, as ProNet's
        EdgeGraphConv uses it (pronet.py:136): x_j = x[0][edge_index[0]], message(x_j, edge_weight), `aggr` over the
        target edge_index[1]."""
        x = kwargs["x"]
        x_src = x[0] if isinstance(x, (tuple, list)) else x
        x_dst = x[1] if isinstance(x, (tuple, list)) else x
        msg = self.message(x_src[edge_index[0]], kwargs.get("edge_weight"))
        reduce = {"add": "sum"}.get(self.aggr, self.aggr)
        return scatter(msg, edge_index[1], dim=0, dim_size=x_dst.size(0), reduce=reduce)


class GraphConv(MessagePassing):
    """torch_geometric.nn.GraphConv(in, out, aggr='add', bias=True) with an overridable
    `message(x_j, edge_weight)` (subclassed at comenet.py:130-133):
    out = lin_rel(sum_{j->i} message(x_j, w_e)) + lin_root(x_i)."""

    def __init__(self, in_channels, out_channels, aggr="add", bias=True, **kwargs):
        super().__init__(aggr=aggr)
        self.lin_rel = _PygLinear(in_channels, out_channels, bias=bias)
        self.lin_root = _PygLinear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def message(self, x_j, edge_weight):
        return x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j

    def forward(self, x, edge_index, edge_weight=None):
        j, i = edge_index[0], edge_index[1]
        msg = self.message(x[j], edge_weight)
        agg = torch.zeros(x.size(0), msg.size(1), dtype=msg.dtype, device=msg.device)
        agg = agg.index_add_(0, i, msg)
        return self.lin_rel(agg) + self.lin_root(x)


class GraphNorm(torch.nn.Module):
    """torch_geometric.nn.GraphNorm(c, eps=1e-5) (used at comenet.py:160,213)."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.empty(in_channels))
        self.bias = torch.nn.Parameter(torch.empty(in_channels))
        self.mean_scale = torch.nn.Parameter(torch.empty(in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        ones(self.weight)
        zeros(self.bias)
        ones(self.mean_scale)

    def forward(self, x, batch=None):
        if batch is None:
            batch = x.new_zeros(x.size(0), dtype=torch.long)
        b = int(batch.max()) + 1
        mean = scatter(x, batch, dim=0, dim_size=b, reduce="mean")
        out = x - mean.index_select(0, batch) * self.mean_scale
        var = scatter(out.pow(2), batch, dim=0, dim_size=b, reduce="mean")
        std = (var + self.eps).sqrt().index_select(0, batch)
        return self.weight * out / std + self.bias


class GaussianSmearing(torch.nn.Module):  # import-only dependency of comenet/features.py:13
    def __init__(self, start=0.0, stop=5.0, num_gaussians=50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)

    def forward(self, dist):
        dist = dist.view(-1, 1) - self.offset.view(1, -1)
        return torch.exp(self.coeff * torch.pow(dist, 2))


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


class InMemoryDataset:  # import-only
    pass


def _download_url(*a, **k):
    raise RuntimeError("no network")


class DataLoader(torch.utils.data.DataLoader):
    """Minimal torch_geometric.data.DataLoader (run.py:6,53-55): concatenates z/pos/force,
    stacks y, builds the sorted `batch` vector."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        super().__init__(dataset, batch_size, shuffle, collate_fn=self._collate, **kw)

    @staticmethod
    def _collate(items):
        out = Data()
        keys = [k for k, v in items[0].__dict__.items() if isinstance(v, torch.Tensor)]
        for k in keys:
            vals = [getattr(it, k) for it in items]
            if k == "y":
                out.y = torch.cat([v.view(-1) for v in vals])
            else:
                setattr(out, k, torch.cat(vals, dim=0))
        out.batch = torch.cat([torch.full((it.z.size(0),), g, dtype=torch.long)
                               for g, it in enumerate(items)])
        out.num_graphs = len(items)
        return out


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-in modules.  Idempotent."""
    if "torch_scatter" in sys.modules and getattr(sys.modules["torch_scatter"], "_dig_oracle_shim", False):
        return
    _module("torch_scatter", scatter=scatter, scatter_min=scatter_min, _dig_oracle_shim=True)
    _module("torch_sparse", SparseTensor=SparseTensor, matmul=_sparse_matmul)
    _module("torch_cluster", radius_graph=radius_graph, knn_graph=knn_graph)
    inits = _module("torch_geometric.nn.inits", glorot_orthogonal=glorot_orthogonal, glorot=glorot,
                    zeros=zeros, ones=ones, uniform=uniform, kaiming_uniform=kaiming_uniform)
    schnet = _module("torch_geometric.nn.models.schnet", GaussianSmearing=GaussianSmearing)
    models = _module("torch_geometric.nn.models", schnet=schnet)
    nn = _module("torch_geometric.nn", radius_graph=radius_graph, knn_graph=knn_graph, GraphConv=GraphConv,
                 GraphNorm=GraphNorm, MessagePassing=MessagePassing, inits=inits, models=models)
    data = _module("torch_geometric.data", Data=Data, DataLoader=DataLoader,
                   InMemoryDataset=InMemoryDataset, download_url=_download_url)
    _module("torch_geometric", nn=nn, data=data, __version__="2.1.0-oracle-shim")
