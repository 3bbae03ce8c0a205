"""
TEST INFRASTRUCTURE: a numpy stand-in for fatiando_b200.harvester.EffectEngine.

Effects come from the CPU oracle (oracle/prism_oracle.py, one reference-order
call per cell and data set, like harvester.py:690-694), the reductions are the
reference's own numpy expressions (harvester.py:434-456), `predicted` is
float32.  It lets the host-side bookkeeping of fatiando_b200/harvester.py run
on machines without a GPU and serves as the checker of the CUDA engine.
"""
import numpy as np

from oracle import prism_oracle as po


class NumpyEngine(object):
    def __init__(self, data):
        self.data = data
        self.props = [d.prop for d in data]
        self.ndata = len(data)
        self.pred = [np.zeros(d.size, dtype='f') for d in data]
        self.rows = {}
        self._next = 0
        self.calls = 0

    def close(self):
        pass

    def release(self, row):
        self.rows.pop(row, None)

    def launches(self):
        return 0

    def _effect(self, d, cell, props):
        if d.prop not in props:
            return np.zeros(d.size, dtype='f')
        b = [np.array([v], dtype=float) for v in (cell.x1, cell.x2, cell.y1, cell.y2, cell.z1, cell.z2)]
        x, y, z = (np.ascontiguousarray(a, dtype=float) for a in (d.x, d.y, d.z))
        if d.field == 'tf':
            from fatiando_b200 import utils
            f = utils.dircos(d.inc, d.dec)
            pm = props[d.prop]
            m = [pm * f[0], pm * f[1], pm * f[2]] if isinstance(pm, (float, int)) else list(pm)
            pr = np.array([m + list(f)], dtype=float)
        else:
            pr = np.array([[float(props[d.prop])]])
        return po.model(d.field, x, y, z, b, pr)

    def effects(self, cells, props):
        out = []
        for c in cells:
            self.rows[self._next] = [self._effect(d, c, props) for d in self.data]
            out.append(self._next)
            self._next += 1
        return np.array(out, dtype=np.int64)

    def accrete(self, row):
        for p, e in zip(self.pred, self.rows[int(row)]):
            p += e

    def evaluate(self, rows):
        self.calls += 1
        misfit = np.zeros((len(rows), self.ndata))
        shape = np.zeros((len(rows), self.ndata))
        for c, row in enumerate(rows):
            for k, d in enumerate(self.data):
                pred = self.pred[k] + self.rows[int(row)][k] if row >= 0 else self.pred[k]
                res = d.observed - pred
                misfit[c, k] = np.sqrt(np.dot(d.weights * res, res)) / d.norm
                alpha = np.sum(d.observed * pred) / d.norm ** 2
                shape[c, k] = np.linalg.norm(alpha * d.observed - pred)
        return misfit, shape

    def predicted(self):
        return [p.copy() for p in self.pred]

    def effect(self, row):
        return [np.asarray(e, dtype=float) for e in self.rows[int(row)]]
