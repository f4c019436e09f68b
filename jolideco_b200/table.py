"""Minimal stand-in for the `astropy.table.Table` the reference uses for the loss trace
(jolideco/loss.py:192-250): `Table(names=, dtype=)`, `add_row(dict)`, `t[-1]["total"]`, `t["total"]`,
`len(t)`, `colnames`.  astropy itself is not a dependency of the hot path."""
import numpy as np


class TraceTable:
    def __init__(self, names=None, dtype=None):
        self.colnames = list(names or [])
        self.dtype = list(dtype or [])
        self.rows = []
        self.meta = {}

    def add_row(self, row):
        self.rows.append(dict(row))

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, item):
        if isinstance(item, str):
            return np.array([r[item] for r in self.rows])
        if isinstance(item, slice):
            t = TraceTable(self.colnames, self.dtype)
            t.rows = self.rows[item]
            return t
        return self.rows[item]

    def __iter__(self):
        return iter(self.rows)

    def to_dict(self):
        return {name: self[name].tolist() for name in self.colnames}

    def __repr__(self):
        return f"<TraceTable rows={len(self)} cols={self.colnames}>"
