"""A plain mapping ``name -> array`` with an ``attrs`` dictionary: what this package uses where the reference uses
``xarray.Dataset`` (absent from this image).  ``RTE``, ``SensorsDict`` and the I/O functions read ``ds[name]`` only."""
from collections import OrderedDict


class Dataset(OrderedDict):
    def __init__(self, *args, attrs=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.attrs = dict(attrs or {})

    def copy(self):
        out = Dataset(self)
        out.attrs = dict(self.attrs)
        return out
