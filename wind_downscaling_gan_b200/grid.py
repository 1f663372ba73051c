"""A small numpy stand-in for the slice of the xarray API that the reference's `api.py` touches.

xarray / netCDF4 / rasterio are not available in this image, so `downscale` / `predict` accept these
containers (and convert real xarray objects to them when xarray is importable).  Only labelled 1-D
coordinates and nearest-neighbour selection are needed (api.py:31-62).
"""
import numpy as np


class GridDataset:
    """Variables on named dimensions with 1-D coordinates.

    data_vars: {name: (dims tuple, ndarray)}; coords: {dim name: 1-D ndarray}.
    """

    def __init__(self, data_vars=None, coords=None):
        self.data_vars = {k: (tuple(d), np.asarray(a)) for k, (d, a) in (data_vars or {}).items()}
        self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
        for name, (dims, a) in self.data_vars.items():
            if a.ndim != len(dims):
                raise ValueError(f"{name}: {a.ndim}-D array for dims {dims}")
            for d, n in zip(dims, a.shape):
                if d in self.coords and len(self.coords[d]) != n:
                    raise ValueError(f"{name}: dim {d} has {n} entries but its coordinate has {len(self.coords[d])}")

    @property
    def dims(self):
        out = {k: len(v) for k, v in self.coords.items()}
        for dims, a in self.data_vars.values():
            out.update(dict(zip(dims, a.shape)))
        return out

    def __getitem__(self, name):
        return self.data_vars[name][1]

    def __contains__(self, name):
        return name in self.data_vars

    def var_dims(self, name):
        return self.data_vars[name][0]

    def transposed(self, name, order):
        dims, a = self.data_vars[name]
        return np.transpose(a, [dims.index(d) for d in order])

    def to_npz(self, path):
        np.savez(path, **{f"var__{k}__{'__'.join(d)}": a for k, (d, a) in self.data_vars.items()},
                 **{f"coord__{k}": v for k, v in self.coords.items()})

    @staticmethod
    def from_npz(path):
        dv, co = {}, {}
        with np.load(path, allow_pickle=False) as z:
            for key in z.files:
                parts = key.split("__")
                if parts[0] == "var":
                    dv[parts[1]] = (tuple(parts[2:]), z[key])
                elif parts[0] == "coord":
                    co[parts[1]] = z[key]
        return GridDataset(dv, co)


def from_xarray(obj):
    """Best-effort conversion of an xarray Dataset/DataArray (when the caller has xarray)."""
    if isinstance(obj, GridDataset):
        return obj
    if hasattr(obj, "data_vars"):
        dv = {str(k): (tuple(v.dims), np.asarray(v.values)) for k, v in obj.data_vars.items()}
        co = {str(k): np.asarray(v.values) for k, v in obj.coords.items() if v.ndim == 1}
        return GridDataset(dv, co)
    if hasattr(obj, "dims") and hasattr(obj, "values"):
        name = obj.name or "elevation"
        co = {str(k): np.asarray(v.values) for k, v in obj.coords.items() if v.ndim == 1}
        return GridDataset({name: (tuple(obj.dims), np.asarray(obj.values))}, co)
    raise TypeError(f"cannot interpret {type(obj)} as a gridded dataset")


def nearest_index(labels, targets):
    """Positions of the labels nearest to each target -- `.sel(..., method='nearest')` on a monotonic
    index (pandas `get_indexer(method='nearest')`: on a tie an increasing index picks the larger label,
    a decreasing one the earlier position)."""
    labels = np.asarray(labels, np.float64)
    targets = np.asarray(targets, np.float64)
    if labels.ndim != 1 or len(labels) == 0:
        raise ValueError("need a non-empty 1-D coordinate")
    if len(labels) == 1:
        return np.zeros(len(targets), np.int64)
    inc = labels[-1] >= labels[0]
    lab = labels if inc else labels[::-1]
    right = np.clip(np.searchsorted(lab, targets, side="left"), 0, len(lab) - 1)
    left = np.clip(right - 1, 0, len(lab) - 1)
    dl, dr = np.abs(targets - lab[left]), np.abs(lab[right] - targets)
    if inc:
        idx = np.where(dl < dr, left, right)
    else:  # positions counted in the original (decreasing) order: `left` there is `right` here
        idx = np.where(dr <= dl, right, left)
        idx = len(lab) - 1 - idx
    return idx.astype(np.int64)
