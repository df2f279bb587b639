"""A CPU stand-in for imputecc_b200.sharded.CudaPhases, used ONLY to exercise the sharded driver
(run_sharded_walk + TorchComm) under gloo on a box without GPUs.  It follows the device kernels'
exchange protocol (csrc/select.cuh, csrc/common.cuh kX* layout) with numpy arithmetic; it is test
infrastructure, not a product path."""
import numpy as np
import scipy.sparse as sp
import torch

K_EXTRAS, BINS, CAND = 8, 4096, 8192
X_SQ0, X_SQ1, X_SQ2, X_IN, X_NEW = 0, 1, 2, 3, 4


class Status:
    def __init__(self, steps_done, stop_reason):
        self.steps_done, self.stop_reason = steps_done, stop_reason
        self.max_row_len, self.max_kept, self.fallback_rows = 0, 0, 0


def keys_of(v):
    if v.dtype == np.float32:
        return v.view(np.uint32).astype(np.uint64) ^ np.uint64(0x80000000), 32     # positive values only
    return v.view(np.uint64) ^ np.uint64(0x8000000000000000), 64


class NumpyPhases:
    def __init__(self, P, rows, rp, perc, tol, dtype):
        self.P, self.r0, self.r1 = P.astype(dtype), rows[0], rows[1]
        self.rp, self.perc, self.tol, self.dtype = rp, perc, tol, np.dtype(dtype)
        n = P.shape[0]
        k = self.r1 - self.r0
        self.I = sp.csr_matrix((np.ones(k, dtype=dtype), (np.arange(k), np.arange(self.r0, self.r1))), shape=(k, n))
        self.Q = self.I.copy()
        self.x = [torch.zeros(K_EXTRAS + BINS, dtype=torch.int64), torch.zeros(BINS, dtype=torch.int64),
                  torch.zeros(2 + CAND, dtype=torch.int64)]
        self.step, self.done, self.reason, self.plateau, self.delta_prev = 0, False, 0, 0, 0.0
        self.trace = []

    def exchange(self, which):
        return self.x[which]

    def set_row_hints(self, hints):
        pass

    def row_hints(self, st, perc, issued):
        return (0, 0)

    def propagate(self):
        if self.done:
            return
        scale = self.dtype.type(1 - self.rp)
        restart = self.dtype.type(np.float32(self.rp))
        new = (self.Q @ self.P).multiply(scale) + self.I.multiply(restart)
        new = sp.csr_matrix(new, dtype=self.dtype)
        d = (self.Q - new)
        sumsq = float((d.data.astype(np.float64) ** 2).sum())
        self.new = new
        hi = int(np.floor(sumsq)); f1 = (sumsq - hi) * 2.0 ** 32; m1 = int(np.floor(f1)); m0 = int(np.floor((f1 - m1) * 2.0 ** 32))
        e = self.x[0]
        e[:] = 0
        e[X_SQ0], e[X_SQ1], e[X_SQ2], e[X_IN], e[X_NEW] = m0, m1, hi, self.Q.nnz, new.nnz
        self.keys, self.kbits = keys_of(np.ascontiguousarray(new.data))

    def hist(self, level):
        if self.done:
            return
        if level == 0:
            d = (self.keys >> np.uint64(self.kbits - 12)).astype(np.int64)
            self.x[0][K_EXTRAS:] += torch.from_numpy(np.bincount(d, minlength=BINS))
        else:
            sel = (self.keys >> np.uint64(self.kbits - 12)) == np.uint64(self.prefix)
            d = ((self.keys[sel] >> np.uint64(self.kbits - 24)) & np.uint64(BINS - 1)).astype(np.int64)
            self.x[1][:] = torch.from_numpy(np.bincount(d, minlength=BINS))

    def scan(self, level):
        if self.done:
            return
        if level == 0:
            h = self.x[0][K_EXTRAS:].numpy()
            n = int(h.sum())
            ft = self.dtype.type
            v = ft(n - 1) * (ft(self.perc) / ft(100))
            kf = np.floor(v)
            self.clamp = bool(v >= ft(n - 1)) or int(kf) >= n - 1
            self.k = n - 1 if self.clamp else int(kf)
            self.gamma = ft(0) if self.clamp else ft(v - kf)
            self.n_total, self.below, self.prefix = n, 0, 0
        else:
            h = self.x[1].numpy()
        cum = np.cumsum(h)
        target = self.k - self.below
        b = int(np.searchsorted(cum, target, side="right"))
        self.below += int(cum[b - 1]) if b > 0 else 0
        self.prefix = (self.prefix << 12) | b

    def gather(self):
        if self.done:
            return
        p24 = self.keys >> np.uint64(self.kbits - 24)
        cand = self.keys[p24 == np.uint64(self.prefix)]
        above = self.keys[p24 > np.uint64(self.prefix)]
        blk = self.x[2]
        blk[:] = 0
        blk[0] = cand.size
        blk[1] = int(above.min()) - (1 << 64) if above.size and int(above.min()) >= (1 << 63) else (int(above.min()) if above.size else -1)
        blk[2:2 + cand.size] = torch.from_numpy(cand.view(np.int64).copy())

    def finish(self, gathered):
        if self.done:
            return
        e = self.x[0].numpy()
        cut, applied = float("nan"), False
        if gathered is not None:
            g = gathered.numpy().reshape(-1, 2 + CAND)
            keys = np.sort(np.concatenate([row[2:2 + int(row[0])] for row in g]).view(np.uint64))
            succ = g[:, 1].view(np.uint64).min()
            r = self.k - self.below

            def dec(kk):
                kk = np.uint64(kk)
                if self.kbits == 32:
                    return np.array([kk ^ np.uint64(0x80000000)], dtype=np.uint64).astype(np.uint32).view(np.float32)[0]
                return np.array([kk ^ np.uint64(0x8000000000000000)], dtype=np.uint64).view(np.float64)[0]
            xk = dec(keys[r])
            xk1 = xk if self.clamp else (dec(keys[r + 1]) if r + 1 < keys.size else dec(succ))
            ft = self.dtype.type
            diff = ft(xk1 - xk)
            out = ft(xk + ft(diff * self.gamma))
            if self.gamma >= 0.5:
                out = ft(xk1 - ft(diff * ft(ft(1) - self.gamma)))
            cut = float(out)
            applied = out > 0
        sumsq = (float(e[X_SQ0]) * 2.0 ** -64 + float(e[X_SQ1]) * 2.0 ** -32) + float(e[X_SQ2])
        delta = float(np.sqrt(sumsq))
        self.trace.append(dict(nnz_in=int(e[X_IN]), nnz_new=int(e[X_NEW]), cut=cut, delta=delta))
        self.Q = self.new
        if applied:
            self.Q = sp.csr_matrix(self.Q.multiply(self.Q > self.dtype.type(cut)), dtype=self.dtype)
        if delta < self.tol:
            self.done, self.reason = True, 1
        else:
            if abs(delta - self.delta_prev) < 0.001:
                self.plateau += 1
            self.delta_prev = delta
            if self.plateau == 5:
                self.done, self.reason = True, 2
        self.step += 1
        self.x[0][:] = 0
        self.x[1][:] = 0

    def poll(self):
        return Status(self.step, self.reason if self.done else 0)
