"""gym.spaces objects for the facades: the real gym / gymnasium classes when importable, else minimal equivalents.

The reference builds its spaces with gym==0.17-era classes (deepcomp/env/single_ue/variants.py:15-17,255-269,
deepcomp/env/multi_ue/central.py:28,147-152); neither gym nor gymnasium is installed in this image, so the fallbacks
below implement exactly what the facades and RLlib's preprocessors touch: shape, dtype, contains, sample and, for
Dict, alphabetical key order (gym sorts plain-dict keys).
"""
from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - depends on the host environment
    from gym import spaces as _gs
    Discrete, MultiDiscrete, MultiBinary, Box, Dict = _gs.Discrete, _gs.MultiDiscrete, _gs.MultiBinary, _gs.Box, _gs.Dict
    BACKEND = 'gym'
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gymnasium import spaces as _gs
        Discrete, MultiDiscrete, MultiBinary, Box, Dict = (_gs.Discrete, _gs.MultiDiscrete, _gs.MultiBinary, _gs.Box,
                                                           _gs.Dict)
        BACKEND = 'gymnasium'
    except Exception:  # noqa: BLE001
        BACKEND = 'builtin'

        class _Space:
            shape = None
            dtype = None
            _rng = np.random.default_rng()

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)

            def __contains__(self, x):
                return self.contains(x)

        class Discrete(_Space):
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()
                self.dtype = np.int64

            def contains(self, x):
                if isinstance(x, (bool, np.bool_)):
                    return False
                if isinstance(x, (int, np.integer)):
                    return 0 <= int(x) < self.n
                if isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
                    return 0 <= int(x) < self.n
                return False

            def sample(self):
                return int(self._rng.integers(0, self.n))

            def __repr__(self):
                return f"Discrete({self.n})"

        class MultiDiscrete(_Space):
            def __init__(self, nvec):
                self.nvec = np.asarray(nvec, dtype=np.int64)
                self.shape = self.nvec.shape
                self.dtype = np.int64

            def contains(self, x):
                x = np.asarray(x)
                return (x.shape == self.shape and np.issubdtype(x.dtype, np.integer) and bool(np.all(x >= 0))
                        and bool(np.all(x < self.nvec)))

            def sample(self):
                return (self._rng.random(self.shape) * self.nvec).astype(np.int64)

            def __repr__(self):
                return f"MultiDiscrete({self.nvec.tolist()})"

        class MultiBinary(_Space):
            def __init__(self, n):
                self.n = int(n)
                self.shape = (self.n,)
                self.dtype = np.int8

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all((x == 0) | (x == 1)))

            def sample(self):
                return self._rng.integers(0, 2, self.shape).astype(np.int8)

            def __repr__(self):
                return f"MultiBinary({self.n})"

        class Box(_Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                if shape is None:
                    shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
                self.shape = tuple(shape)
                self.dtype = np.dtype(dtype)
                self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

            def sample(self):
                return self._rng.uniform(self.low, self.high).astype(self.dtype)

            def __repr__(self):
                return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        class Dict(_Space):
            def __init__(self, spaces):
                if not isinstance(spaces, OrderedDict):
                    spaces = OrderedDict(sorted(spaces.items()))
                self.spaces = spaces

            def contains(self, x):
                return (isinstance(x, dict) and set(x.keys()) == set(self.spaces.keys())
                        and all(self.spaces[k].contains(x[k]) for k in self.spaces))

            def sample(self):
                return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

            def __getitem__(self, k):
                return self.spaces[k]

            def __repr__(self):
                return "Dict(" + ", ".join(f"{k}: {s!r}" for k, s in self.spaces.items()) + ")"
