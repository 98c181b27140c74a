"""`.grid` files of the reference (bz2-pickled dict, /root/reference/ns/model/data.py:208-235) straight onto
the device, and the matrix -> edge-list convention its networks use (:22-46).

Only the container and the format are mirrored; the dataset generators (pyamg.gallery / pygmsh based) and the
plotting helpers are out of scope — `mlamg.problems` generates the named shapes.
"""
import bz2
import os
import pickle

import scipy.sparse as sp
import torch

from mlamg import core


class Grid:
    """A, node coordinates and free-form extras — the reference's dataset item (data.py:65-84)."""

    def __init__(self, A_csr, x=None, extra=None):
        self.A = A_csr.tocsr()
        self.x = x
        self.extra = extra if extra is not None else {}
        self._device = None

    @property
    def networkx(self):
        raise NotImplementedError("networkx view is not needed by the device path")

    def device_csr(self, dtype=None):
        """A resident in HBM (cached)"""
        if self._device is None or (dtype is not None and self._device.dtype != dtype):
            self._device = core.DeviceCSR.from_scipy(self.A, dtype)
        return self._device

    def save(self, fname):
        """same on-disk layout as data.py:207-219: A as the (data, indices, indptr) tuple"""
        if '.grid' not in fname:
            fname = fname + '.grid'
        A = self.A.tocsr()
        with bz2.BZ2File(fname, "wb") as f:
            pickle.dump({'A': (A.data, A.indices, A.indptr), 'x': self.x, 'extra': self.extra}, f)

    @staticmethod
    def load(fname):
        """data.py:221-234"""
        if '.grid' not in fname:
            fname = fname + '.grid'
        with bz2.BZ2File(fname, "rb") as f:
            loaded = pickle.load(f)
        extra = loaded['extra'] if 'extra' in loaded else {}
        extra['filename'] = fname
        A = loaded['A']
        if isinstance(A, tuple):
            A = sp.csr_matrix(A)
        return Grid(A, loaded['x'], extra)


def load_dir(directory):
    """every `.grid` file of a directory (data.py:236-242), sorted by name so that dataset order is reproducible"""
    return [Grid.load(os.path.join(directory, f)) for f in sorted(os.listdir(directory)) if '.grid' in f.lower()]


Grid.load_dir = staticmethod(load_dir)


def edge_list(A, with_values=True):
    """edge_index (2 x nnz, CSR order incl. the diagonal) and edge_attr = |A_ij| float32 — the seam between A and
    the networks' per-edge outputs (graph_from_matrix_basic, data.py:39-46), as device tensors."""
    Ad = core.DeviceCSR.wrap(A)
    n = Ad.shape[0]
    rows = torch.repeat_interleave(torch.arange(n, device=Ad.col.device), (Ad.rowptr[1:] - Ad.rowptr[:-1]).long())
    edge_index = torch.stack([rows, Ad.col.long()])
    if not with_values:
        return edge_index
    return edge_index, Ad.val.abs().to(torch.float32)
