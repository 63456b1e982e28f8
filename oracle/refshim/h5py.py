"""`h5py` stand-in for running the UNMODIFIED reference in the build container (test infrastructure).

The reference imports h5py in network.py:1 and uses it only in Network.fromModel (:157-228) and save_weights
(:428-472).  This stand-in implements exactly the calls those two functions make — File(path, "r" | "w") as a
context manager, create_dataset / create_group, keys(), [] lookup, dataset[()] and group.values() — on top of an
`.npz` archive whose member names are the HDF5 paths ("weights/weights_0").  With it the reference's own code
writes and reads the checkpoint format of deep-neural-net_b200/mlpcode/checkpoint.py, which is how
oracle/gen_golden.py pins that format.  Like h5py, groups iterate in alphabetical order of the member names."""
import numpy as np


class _Dataset:
    def __init__(self, value):
        self._v = np.asarray(value)

    def __getitem__(self, idx):
        assert idx == (), "the reference only reads whole datasets: ds[()]"
        return self._v[()]


class _Group:
    def __init__(self, store, prefix):
        self._store, self._prefix = store, prefix

    def _names(self):
        p = self._prefix
        return sorted({k[len(p):].split("/")[0] for k in self._store if k.startswith(p)})

    def keys(self):
        return self._names()

    def __contains__(self, name):
        return name in self._names()

    def __getitem__(self, name):
        full = self._prefix + name
        if full in self._store:
            return _Dataset(self._store[full])
        assert any(k.startswith(full + "/") for k in self._store), f"no such member: {full}"
        return _Group(self._store, full + "/")

    def values(self):
        return [self[n] for n in self._names()]

    def create_group(self, name):
        return _Group(self._store, self._prefix + name + "/")

    def create_dataset(self, name, data=None, dtype=None):
        self._store[self._prefix + name] = np.asarray(data, dtype=dtype)


class File(_Group):
    def __init__(self, path, mode="r"):
        assert mode in ("r", "w")
        self._path, self._mode = path, mode
        store = {}
        if mode == "r":
            with np.load(path) as fp:
                store = {k: fp[k] for k in fp.files}
        super().__init__(store, "")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self._mode == "w" and exc[0] is None:
            with open(self._path, "wb") as fh:
                np.savez(fh, **self._store)
        return False
