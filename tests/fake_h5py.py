"""A dictionary-backed stand-in for the small part of the h5py API that fake_spectra_b200.savefile uses (File as a context
manager, create_group / require_group, create_dataset with '/'-separated names creating the intermediate groups, item
access by path, attrs, visititems, Dataset as a class, np.array(dataset)).  h5py itself is not installed in this
environment: this exercises the HDF5 BRANCH of the savefile logic (names, attributes, lazy placeholders), not the HDF5
format.  Files are persisted as a pickle so that a second File(..., "r") sees what the first wrote."""
import os
import pickle

import numpy as np


class Dataset:
    def __init__(self, data):
        self._data = np.array(data)

    def __array__(self, dtype=None, copy=None):
        return self._data if dtype is None else self._data.astype(dtype)

    def __getitem__(self, key):
        return self._data[key]

    def len(self):
        return self._data.shape[0]

    @property
    def shape(self):
        return self._data.shape


class Group:
    def __init__(self):
        self._items = {}
        self.attrs = {}

    def _walk(self, path, create=False):
        node = self
        for part in [p for p in path.split("/") if p]:
            if part not in node._items:
                if not create:
                    raise KeyError(path)
                node._items[part] = Group()
            node = node._items[part]
        return node

    def create_group(self, name):
        if name in self._items:
            raise ValueError("group exists: " + name)
        return self._walk(name, create=True)

    def require_group(self, name):
        return self._walk(name, create=True)

    def create_dataset(self, name, data=None):
        parts = [p for p in name.split("/") if p]
        parent = self._walk("/".join(parts[:-1]), create=True)
        if parts[-1] in parent._items:
            raise ValueError("dataset exists: " + name)
        parent._items[parts[-1]] = Dataset(data)
        return parent._items[parts[-1]]

    def __getitem__(self, path):
        return self._walk(path)

    def __setitem__(self, name, value):
        self.create_dataset(name, data=value)

    def __delitem__(self, name):
        del self._items[name]

    def __contains__(self, name):
        try:
            self._walk(name)
            return True
        except KeyError:
            return False

    def keys(self):
        return self._items.keys()

    def visititems(self, func, prefix=""):
        for name, obj in self._items.items():
            path = prefix + name
            func(path, obj)
            if isinstance(obj, Group):
                obj.visititems(func, path + "/")


class File(Group):
    def __init__(self, name, mode="r"):
        super().__init__()
        self._name, self._mode = name, mode
        if mode == "r":
            if not os.path.exists(name):
                raise IOError("unable to open " + str(name))
            with open(name, "rb") as fh:
                root = pickle.load(fh)
            self._items, self.attrs = root._items, root.attrs

    def close(self):
        if self._mode != "r":
            root = Group()
            root._items, root.attrs = self._items, self.attrs
            with open(self._name, "wb") as fh:
                pickle.dump(root, fh)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def is_hdf5(name):
    return os.path.exists(name)
