"""QM93D / MD17 with the interface of reference dig/threedgraph/dataset/PygQM93D.py:11-117 and PygMD17.py:10-109,
without torch_geometric (not installable here) and without a network: the raw `.npz` must already be at the path the
reference downloads it to (`<root>/qm9/raw/qm9_eV.npz`, `<root>/<name>/raw/<name>_dft.npz`).

The reference materialises one `Data` object per molecule, collates them into one big `Data` + `slices` and caches that
on disk.  Here the concatenated arrays of the npz ARE the storage: `dataset.data` is an attribute bag over them
(`dataset.data.y = dataset.data['U0']` works as in the reference's example), `dataset[i]` is a view, `dataset[index
tensor]` a sub-dataset sharing the storage, and iteration feeds dig_b200.data.DataLoader / collate directly.
"""
import os.path as osp

import numpy as np
import torch
from sklearn.utils import shuffle

from ...data import Molecule

QM9_TARGETS = ['mu', 'alpha', 'homo', 'lumo', 'gap', 'r2', 'zpve', 'U0', 'U', 'H', 'G', 'Cv']


class _Collated:
    """The big collated `Data` of an InMemoryDataset: attribute access, `data[key]`, `data.keys`."""

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    @property
    def keys(self):
        return [k for k, v in vars(self).items() if isinstance(v, torch.Tensor)]


class _InMemory3D:
    url = None

    def __init__(self, folder, raw_name):
        self.folder = folder
        self.raw_dir = osp.join(folder, 'raw')
        path = osp.join(self.raw_dir, raw_name)
        if not osp.exists(path):
            raise FileNotFoundError(
                f"{path} not found.  This environment has no network: fetch {self.url} by hand into {self.raw_dir} "
                "(the reference's InMemoryDataset would download it there)")
        self.data = _Collated()
        self._ptr = None          # [M + 1] node offsets of the per-atom attributes
        self._node_keys = ()
        self._indices = None      # None = all molecules, else a LongTensor of molecule ids (sub-dataset)
        self._load(np.load(path))

    # -- reference API ------------------------------------------------------------------------------------------
    def get_idx_split(self, data_size, train_size, valid_size, seed):
        """reference PygQM93D.py:103-107 / PygMD17.py:90-94 (sklearn shuffle, so the same seed gives the same split)."""
        ids = shuffle(range(data_size), random_state=seed)
        train_idx, val_idx, test_idx = (torch.tensor(ids[:train_size]),
                                        torch.tensor(ids[train_size:train_size + valid_size]),
                                        torch.tensor(ids[train_size + valid_size:]))
        return {'train': train_idx, 'valid': val_idx, 'test': test_idx}

    def __len__(self):
        return int(self._ptr.numel() - 1) if self._indices is None else int(self._indices.numel())

    def __repr__(self):
        return f'{type(self).__name__}({len(self)})'

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)) or (isinstance(idx, torch.Tensor) and idx.dim() == 0):
            i = int(idx)
            if i < 0:
                i += len(self)
            if not 0 <= i < len(self):
                raise IndexError(idx)
            return self._molecule(i if self._indices is None else int(self._indices[i]))
        if isinstance(idx, slice):
            idx = torch.arange(len(self))[idx]
        idx = torch.as_tensor(idx)
        if idx.dtype == torch.bool:
            idx = idx.nonzero().flatten()
        sub = object.__new__(type(self))
        sub.__dict__.update(self.__dict__)
        base = torch.arange(int(self._ptr.numel() - 1)) if self._indices is None else self._indices
        sub._indices = base[idx.long()]
        return sub

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    # -- storage --------------------------------------------------------------------------------------------------
    def _molecule(self, m):
        lo, hi = int(self._ptr[m]), int(self._ptr[m + 1])
        out = Molecule.__new__(Molecule)
        for k in self.data.keys:
            v = getattr(self.data, k)
            setattr(out, k, v[lo:hi] if k in self._node_keys else v[m:m + 1] if v.dim() == 1 else v[m])
        return out


class QM93D(_InMemory3D):
    r"""reference PygQM93D.py:11-117: ~131 k QM9 molecules, 12 targets; `root/qm9/raw/qm9_eV.npz` (keys R, Z, N + targets)."""
    url = 'https://github.com/klicperajo/dimenet/raw/master/data/qm9_eV.npz'

    def __init__(self, root='dataset/', transform=None, pre_transform=None, pre_filter=None):
        if transform is not None or pre_transform is not None or pre_filter is not None:
            raise NotImplementedError("transform / pre_transform / pre_filter are not supported by this reader")
        super().__init__(osp.join(root, 'qm9'), 'qm9_eV.npz')

    def _load(self, raw):
        n = torch.from_numpy(np.asarray(raw['N'], dtype=np.int64))
        self._ptr = torch.zeros(n.numel() + 1, dtype=torch.long)
        self._ptr[1:] = torch.cumsum(n, 0)
        self.data.pos = torch.tensor(np.asarray(raw['R']), dtype=torch.float32)
        self.data.z = torch.tensor(np.asarray(raw['Z']), dtype=torch.int64)
        self._node_keys = ('pos', 'z')
        for name in QM9_TARGETS:
            setattr(self.data, name, torch.tensor(np.asarray(raw[name]), dtype=torch.float32).reshape(-1))
        self.data.y = self.data.mu               # Data(..., y=y_i[0], mu=y_i[0], ...)   PygQM93D.py:92


class MD17(_InMemory3D):
    r"""reference PygMD17.py:10-109: one molecule's MD trajectory with energies and forces;
    `root/<name>/raw/<name>_dft.npz` (keys E, F, R, z)."""

    def __init__(self, root='dataset/', name='benzene_old', transform=None, pre_transform=None, pre_filter=None):
        if transform is not None or pre_transform is not None or pre_filter is not None:
            raise NotImplementedError("transform / pre_transform / pre_filter are not supported by this reader")
        self.name = name
        self.url = 'http://quantum-machine.org/gdml/data/npz/' + name + '_dft.npz'
        super().__init__(osp.join(root, name), name + '_dft.npz')

    def _load(self, raw):
        r = np.asarray(raw['R'])
        frames, atoms = r.shape[0], r.shape[1]
        self._ptr = torch.arange(frames + 1, dtype=torch.long) * atoms
        self.data.pos = torch.tensor(r.reshape(frames * atoms, 3), dtype=torch.float32)
        self.data.z = torch.tensor(np.asarray(raw['z']), dtype=torch.int64).repeat(frames)
        self.data.force = torch.tensor(np.asarray(raw['F']).reshape(frames * atoms, 3), dtype=torch.float32)
        self.data.y = torch.tensor(np.asarray(raw['E']), dtype=torch.float32).reshape(-1)
        self._node_keys = ('pos', 'z', 'force')
