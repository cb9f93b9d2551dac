"""Second, independent restatement of the reference envs in plain Python (TEST INFRASTRUCTURE ONLY).

Written separately from oracle/qg_oracle.hpp (different data structures: Python sets / lists / ints) and
cross-checked against it in tests/test_oracle.py; disagreements are resolved by re-reading the cited Rust
lines (paths relative to /root/reference/rust/src).  Slow: small cases only.
"""
from __future__ import annotations

import struct

import numpy as np

ONE_Q = ("h", "s", "sdg", "sx", "sxdg")


def f32(x):
    return np.float32(x)


class Metrics:  # envs/metrics.rs:18-146
    def __init__(self, n, weights=None):
        self.n = n
        w = {"n_cnots": 0.01, "n_layers_cnots": 0.0, "n_layers": 0.0, "n_gates": 0.0001}
        w.update({k: v for k, v in (weights or {}).items() if k in w})
        self.w = {k: f32(v) for k, v in w.items()}
        self.reset()

    def reset(self):
        self.n_cnots = self.n_gates = 0
        self.layers, self.cx_layers = set(), set()
        self.last = [-1] * self.n
        self.last_cx = [-1] * self.n

    def snapshot(self):
        return (self.n_cnots, len(self.cx_layers), len(self.layers), self.n_gates)

    def _single(self, t):
        if t >= self.n:
            return
        self.n_gates += 1
        self.last[t] += 1
        self.layers.add(self.last[t])

    def _cx(self, c, t):
        if c == t or c >= self.n or t >= self.n:
            return
        self.n_cnots += 1
        self.n_gates += 1
        L = max(self.last[c], self.last[t]) + 1
        self.last[c] = self.last[t] = L
        self.layers.add(L)
        L = max(self.last_cx[c], self.last_cx[t]) + 1
        self.last_cx[c] = self.last_cx[t] = L
        self.cx_layers.add(L)

    def apply(self, name, q):
        name = name.lower()
        if name == "cx":
            self._cx(q[0], q[1])
        elif name == "swap":
            self._cx(q[0], q[1]); self._cx(q[1], q[0]); self._cx(q[0], q[1])
        elif name == "cz":
            self._single(q[1]); self._cx(q[0], q[1]); self._single(q[1])
        else:
            self._single(q[0])

    def penalty(self, prev, now):
        d = [f32(max(a - b, 0)) for a, b in zip(now, prev)]
        w = self.w
        s = f32(w["n_cnots"] * d[0]) + f32(w["n_layers_cnots"] * d[1])
        s = f32(s) + f32(w["n_layers"] * d[2])
        s = f32(s) + f32(w["n_gates"] * d[3])
        return f32(s)


class _Shell:
    """Episode shell shared by Permutation / LinearFunction / Clifford."""

    def __init__(self, n, gateset, difficulty=1, depth_slope=2, max_depth=128, metrics_weights=None, add_inverts=True, track_solution=True):
        self.n, self.gateset = n, [(g.lower().strip(), tuple(q)) for g, q in gateset]
        self.difficulty, self.depth_slope, self.max_depth = difficulty, depth_slope, max_depth
        self.metrics = Metrics(n, metrics_weights)
        self.add_inverts, self.track = add_inverts, track_solution
        self.depth = 1
        self.identity()
        self._internals()
        self.reward_value = f32(1.0 if self.success else 0.0)

    def _internals(self):
        self.success = self.solved()
        self.metrics.reset()
        self.counts = self.metrics.snapshot()
        self.reward_value = f32(1.0 if self.success else 0.0)
        self.inverted = False
        self.sol, self.sol_inv = [], []

    def set_state(self, st):
        self.load(st)
        self.depth = self.max_depth
        self._internals()

    def step(self, action, coin=None):
        penalty = f32(0.0)
        valid = 0 <= action < len(self.gateset)
        if valid:
            name, q = self.gateset[action]
            prev = self.counts
            self.metrics.apply(name, q)
            self.counts = self.metrics.snapshot()
            penalty = self.metrics.penalty(prev, self.counts)
            self.apply(name, q)
        if self.track and (valid or self.PUSH_INVALID):
            (self.sol_inv if self.inverted else self.sol).append(action)
        self.depth = max(self.depth - 1, 0)
        if self.add_inverts and coin:
            self.invert()
            self.inverted = not self.inverted
        self.success = self.solved()
        self.reward_value = f32(f32(1.0 if self.success else 0.0) - penalty)

    def masks(self):
        return [not self.success] * len(self.gateset)

    def is_final(self):
        return self.depth == 0 or self.success

    def reward(self):
        return float(self.reward_value)

    def solution(self):
        return self.sol + self.sol_inv[::-1]


class PermutationRef(_Shell):  # envs/permutation.rs
    PUSH_INVALID = False

    def identity(self):
        self.state = list(range(self.n))

    def load(self, st):
        self.state = [int(x) for x in st]

    def solved(self):
        return all(v == i for i, v in enumerate(self.state))

    def apply(self, name, q):
        if name == "swap":
            a, b = q
            self.state[a], self.state[b] = self.state[b], self.state[a]

    def invert(self):
        inv = [0] * len(self.state)
        for i, v in enumerate(self.state):
            inv[v] = i
        self.state = inv

    def observe(self):
        return [i * self.n + v for i, v in enumerate(self.state)]

    def obs_shape(self):
        return [self.n, self.n]

    def raw_state(self):
        return list(self.state)


class _MatrixRef(_Shell):
    PUSH_INVALID = True

    def identity(self):
        self.m = np.eye(self.dim, dtype=np.uint8)

    def load(self, st):
        self.m = (np.asarray(st, dtype=np.int64) > 0).astype(np.uint8).reshape(self.dim, self.dim)

    def solved(self):
        return bool(np.array_equal(self.m, np.eye(self.dim, dtype=np.uint8)))

    def invert(self):  # any algorithm gives the unique inverse; here: solve with augmented matrix
        d = self.dim
        aug = np.concatenate([self.m.copy(), np.eye(d, dtype=np.uint8)], axis=1)
        for c in range(d):
            piv = next(r for r in range(c, d) if aug[r, c])
            if piv != c:
                aug[[c, piv]] = aug[[piv, c]]
            for r in range(d):
                if r != c and aug[r, c]:
                    aug[r] ^= aug[c]
        self.m = aug[:, d:].copy()

    def observe(self):
        return [int(i) for i in np.nonzero(self.m.reshape(-1))[0]]

    def obs_shape(self):
        return [self.dim, self.dim]

    def raw_state(self):
        return self.m.reshape(-1).tolist()


class LinearFunctionRef(_MatrixRef):  # envs/linear_function.rs
    @property
    def dim(self):
        return self.n

    def apply(self, name, q):
        if name == "cx" and q[0] != q[1]:
            self.m[q[1]] ^= self.m[q[0]]
        elif name == "swap" and q[0] != q[1]:
            self.m[[q[0], q[1]]] = self.m[[q[1], q[0]]]


class CliffordRef(_MatrixRef):  # envs/clifford.rs
    @property
    def dim(self):
        return 2 * self.n

    def apply(self, name, q):
        n, m = self.n, self.m
        a = q[0]
        b = q[1] if len(q) > 1 else None
        if name == "h":
            m[[a, n + a]] = m[[n + a, a]]
        elif name in ("s", "sdg"):
            m[n + a] ^= m[a]
        elif name in ("sx", "sxdg"):
            m[a] ^= m[n + a]
        elif a == b:
            return
        elif name == "cx":
            m[b] ^= m[a]; m[n + a] ^= m[n + b]
        elif name == "cz":
            m[n + a] ^= m[b]; m[n + b] ^= m[a]
        elif name == "swap":
            m[[a, b]] = m[[b, a]]; m[[n + a, n + b]] = m[[n + b, n + a]]


class PauliRef:
    """envs/pauli.rs + pauli/*.rs (no perms; set_state/step/observe/solution)."""

    def __init__(self, n, gateset, max_rotations=5, depth_slope=2, max_depth=128, metrics_weights=None, pauli_layer_reward=0.01, track_solution=True):
        self.n, self.gateset = n, [(g.lower().strip(), tuple(q)) for g, q in gateset]
        self.max_rot = max(max_rotations, 1)
        self.max_depth, self.plr, self.track = max_depth, f32(pauli_layer_reward), track_solution
        self.metrics = Metrics(n, metrics_weights)
        self.set_state([0] + np.eye(2 * n, dtype=np.int64).reshape(-1).tolist())
        self.depth = 1

    def set_state(self, st):
        n = self.n
        st = list(st)
        R = max(st[0], 0)
        tab = [1 if x > 0 else 0 for x in st[1:1 + 4 * n * n]]
        i = 1 + 4 * n * n
        labels = []
        for r in range(R):
            ln = st[i]; i += 1
            lab = "".join(chr(c) for c in st[i:i + ln]); i += ln
            if r < self.max_rot:
                labels.append(lab)
        self.tab = [[tab[r * 2 * n + c] for c in range(2 * n)] for r in range(2 * n)]
        self.rots = []   # each: dict x[], z[], phase, alive
        for lab in labels:
            coeff = lab.rstrip("IXYZ")
            body = lab[len(coeff):]
            canon = coeff.replace("1", "").replace("+", "").replace("j", "i")
            ph = {"": 0, "-i": 1, "-": 2, "i": 3}[canon]
            x = [body[len(body) - 1 - k] in "XY" for k in range(n)]
            z = [body[len(body) - 1 - k] in "ZY" for k in range(n)]
            ys = sum(1 for k in range(n) if x[k] and z[k])
            self.rots.append({"x": x, "z": z, "phase": (ph + ys) % 4})
        R = len(self.rots)
        self.cols = [[int(r["x"][k]) for k in range(n)] + [int(r["z"][k]) for k in range(n)] for r in self.rots]  # data columns
        self.nodes = list(range(R))   # petgraph node index -> rotation id
        self.edges = {(i, j) for i in range(R) for j in range(i) if self._anti(self.rots[i], self.rots[j])}
        self.depth = self.max_depth
        self.success = self.solved()
        self.metrics.reset()
        self.counts = self.metrics.snapshot()
        self.reward_value = f32(1.0 if self.success else 0.0)
        self.sol = []

    @staticmethod
    def _anti(a, b):
        s = sum((a["x"][k] and b["z"][k]) + (a["z"][k] and b["x"][k]) for k in range(len(a["x"])))
        return s % 2 == 1

    def solved(self):
        n = self.n
        return not self.nodes and all(self.tab[r][c] == (1 if r == c else 0) for r in range(2 * n) for c in range(2 * n))

    # row operations act on the tableau rows and on every rotation column entry
    def _xor_rows(self, a, b):
        self.tab[a] = [u ^ v for u, v in zip(self.tab[a], self.tab[b])] if a != b else [0] * (2 * self.n)
        for col in self.cols:
            col[a] = col[a] ^ col[b] if a != b else 0

    def _swap_rows(self, a, b):
        self.tab[a], self.tab[b] = self.tab[b], self.tab[a]
        for col in self.cols:
            col[a], col[b] = col[b], col[a]

    def _h(self, i):
        self._swap_rows(i, self.n + i)
        for r in self.rots:
            x, z = r["x"][i], r["z"][i]
            r["x"][i], r["z"][i] = z, x
            r["phase"] = (r["phase"] + 2 * int(x and z)) % 4

    def _s(self, i):
        self._xor_rows(self.n + i, i)
        for r in self.rots:
            x = r["x"][i]
            r["z"][i] ^= x
            r["phase"] = (r["phase"] + int(x)) % 4

    def _sx(self, i):
        self._xor_rows(i, self.n + i)
        for r in self.rots:
            for op in ("h", "s", "h"):
                x, z = r["x"][i], r["z"][i]
                if op == "h":
                    r["x"][i], r["z"][i] = z, x
                    r["phase"] = (r["phase"] + 2 * int(x and z)) % 4
                else:
                    r["z"][i] ^= x
                    r["phase"] = (r["phase"] + int(x)) % 4

    def _clean(self):
        n = self.n
        out = []
        removed = True
        while removed:
            removed = False
            front = [p for p, rid in enumerate(self.nodes) if not any((rid, other) in self.edges for other in self.nodes)]
            kill = []
            for p in front:
                rid = self.nodes[p]
                col = self.cols[rid]
                w = sum(col[k] | col[n + k] for k in range(n))
                if w <= 1:
                    q = next(k for k in range(n) if col[k] | col[n + k])
                    axis = (1 if col[n + q] else 0) if col[q] else 2
                    out.append((axis, q, rid))
                    kill.append(p)
                    self.cols[rid] = [0] * (2 * n)
                    removed = True
            for p in range(len(self.nodes) - 1, -1, -1):   # retain_nodes: high -> low, swap_remove
                if p in kill:
                    self.nodes[p] = self.nodes[-1]
                    self.nodes.pop()
        return out

    def _cnot(self, i, j):
        self._xor_rows(i, j)
        self._xor_rows(self.n + j, self.n + i)
        for r in self.rots:
            r["x"][i] ^= r["x"][j]
            r["z"][j] ^= r["z"][i]
        return self._clean()

    def _act(self, name, q):
        if name == "h":
            self._h(q[0])
        elif name == "s":
            self._s(q[0])
        elif name == "sdg":
            self._s(q[0]); self._s(q[0]); self._s(q[0])
        elif name == "sx":
            self._sx(q[0])
        elif name == "sxdg":
            self._sx(q[0]); self._sx(q[0]); self._sx(q[0])
        elif name == "cx":
            return self._cnot(q[0], q[1])
        elif name == "cz":
            self._h(q[1]); out = self._cnot(q[0], q[1]); self._h(q[1]); return out
        elif name == "swap":
            return self._cnot(q[0], q[1]) + self._cnot(q[1], q[0]) + self._cnot(q[0], q[1])
        return []

    def step(self, action, coin=None):
        penalty, k = f32(0.0), 0
        if 0 <= action < len(self.gateset):
            name, q = self.gateset[action]
            prev = self.counts
            self.metrics.apply(name, q)
            self.counts = self.metrics.snapshot()
            penalty = self.metrics.penalty(prev, self.counts)
            hv = self._act(name, q)
            k = len(hv)
            if self.track:
                self.sol.append(action)
                for axis, qubit, rid in hv:
                    r = self.rots[rid]
                    ys = sum(1 for t in range(self.n) if r["x"][t] and r["z"][t])
                    ph = (r["phase"] + 4 * self.n - ys) % 4
                    self.sol.append(0x80000000 | (axis << 21) | (qubit << 11) | (rid << 1) | (0 if ph == 2 else 1))
        self.depth = max(self.depth - 1, 0)
        self.success = self.solved()
        self.reward_value = f32(f32(f32(1.0 if self.success else 0.0) - penalty) + f32(self.plr * f32(k)))

    def observe(self):
        n, mc = self.n, 2 * self.n + self.max_rot
        out = []
        for r in range(2 * n):
            row = list(self.tab[r]) + [self.cols[rid][r] for rid in self.nodes[: self.max_rot]]
            out += [r * mc + c for c, v in enumerate(row) if v]
        return out

    def obs_shape(self):
        return [2 * self.n, 2 * self.n + self.max_rot]

    def masks(self):
        return [not self.success] * len(self.gateset)

    def is_final(self):
        return self.depth == 0 or self.success

    def reward(self):
        return float(self.reward_value)

    def solution(self):
        return list(self.sol)

    def raw_state(self):
        n, R = self.n, len(self.rots)
        return [v for r in range(2 * n) for v in (list(self.tab[r]) + [self.cols[k][r] for k in range(R)])]


def make(kind, n, gateset, **kw):
    if kind == 0:
        return PermutationRef(n, gateset, **kw)
    if kind == 1:
        return LinearFunctionRef(n, gateset, **kw)
    if kind == 2:
        return CliffordRef(n, gateset, **kw)
    kw.pop("add_inverts", None)
    return PauliRef(n, gateset, **kw)


def f32_bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]
