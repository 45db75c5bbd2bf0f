"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (deepcomp_b200/).

Stub modules that let the *unmodified* reference env code under /root/reference
(deepcomp/env/**, deepcomp/util/constants.py, deepcomp/util/logs.py) import in a
container that lacks its third-party dependencies (gym, ray, shapely, structlog,
structlog_round, matplotlib, svgpath2mpl -- see SURVEY.md section 8c).

Only two stubs carry arithmetic, and both restate what the missing dependency
does for the calls the hot path makes:

* ``shapely.geometry.Point`` -- shapely 1.7.0 / GEOS point-point distance is
  ``sqrt(dx*dx + dy*dy)`` in double (call sites: station.py:124, movement.py:142);
  equality is coordinate-tuple equality (movement.py:169); ``within``/``touches``
  against the rectangular map polygon are strict-interior / on-boundary tests
  (movement.py:129,165).
* ``gym.spaces.Dict`` -- sorts plain-dict keys alphabetically like gym<0.22
  (central.py:33-44 iterates ``observation_space.spaces.keys()``).

Everything else (logging, plotting) is a no-op.  This file is used by
``oracle/ref_loader.py`` in THIS container only (to pin the restatement in
``oracle/deepcomp_oracle.py`` and to generate ``tests/golden/*.npz``); it cannot
be used on the GPU box because /root/reference does not exist there.
"""
import math
import sys
import types
from collections import OrderedDict

import numpy as np


# --------------------------------------------------------------------------- shapely
class Point:
    """Pure-Python stand-in for shapely.geometry.Point (2-D, immutable)."""
    __slots__ = ('x', 'y')

    def __init__(self, *args):
        if len(args) == 1:
            # Point(np.array([x, y])) as in movement.py:153
            x, y = args[0][0], args[0][1]
        else:
            x, y = args
        object.__setattr__(self, 'x', float(x))
        object.__setattr__(self, 'y', float(y))

    def __setattr__(self, k, v):
        raise AttributeError("Point is immutable")

    def distance(self, other):
        dx = self.x - other.x
        dy = self.y - other.y
        return math.sqrt(dx * dx + dy * dy)

    def __eq__(self, other):
        return isinstance(other, Point) and self.x == other.x and self.y == other.y

    def __hash__(self):
        return hash((self.x, self.y))

    def within(self, poly):
        return poly._contains_strict(self)

    def touches(self, poly):
        return poly._on_boundary(self)

    def buffer(self, r):
        return _Buffered(self, r)

    def __str__(self):
        return f"POINT ({self.x:g} {self.y:g})"

    __repr__ = __str__


class _Buffered:
    def __init__(self, p, r):
        self.center, self.r = p, r


class Polygon:
    """Axis-aligned rectangle is all the hot path ever builds (map.py:27-28, station.py:38-41)."""

    def __init__(self, pts):
        xs = [p[0] for p in pts]
        ys = [p[1] for p in pts]
        self.min_x, self.max_x = min(xs), max(xs)
        self.min_y, self.max_y = min(ys), max(ys)

    def _contains_strict(self, p):
        return self.min_x < p.x < self.max_x and self.min_y < p.y < self.max_y

    def _on_boundary(self, p):
        inside_closed = self.min_x <= p.x <= self.max_x and self.min_y <= p.y <= self.max_y
        return inside_closed and not self._contains_strict(p)


# --------------------------------------------------------------------------- gym
class _Space:
    shape = None

    def contains(self, x):
        return True


class Discrete(_Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()

    def contains(self, x):
        try:
            xi = int(x)
        except (TypeError, ValueError):
            return False
        return 0 <= xi < self.n and xi == x


class MultiBinary(_Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = (self.n,)


class MultiDiscrete(_Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.nvec.shape and bool(np.all(x >= 0)) and bool(np.all(x < self.nvec))


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Dict(_Space):
    def __init__(self, spaces):
        if not isinstance(spaces, OrderedDict):
            spaces = OrderedDict(sorted(spaces.items()))
        self.spaces = spaces


class _GymEnv:
    metadata = {}

    def __init__(self, *a, **k):
        pass


class _NullLogger:
    def bind(self, **kw):
        return self

    def _noop(self, *a, **k):
        return None

    info = debug = warning = error = msg = _noop


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Insert the stubs into sys.modules (idempotent)."""
    if getattr(install, '_done', False):
        return
    # shapely
    geom = _mod('shapely.geometry', Point=Point, Polygon=Polygon)
    _mod('shapely', geometry=geom)
    # gym
    spaces = _mod('gym.spaces', Discrete=Discrete, MultiBinary=MultiBinary, MultiDiscrete=MultiDiscrete,
                  Box=Box, Dict=Dict)
    glogger = _mod('gym.logger', set_level=lambda lvl: None, ERROR=40)
    _mod('gym', Env=_GymEnv, spaces=spaces, logger=glogger)
    # structlog
    stdlib = _mod('structlog.stdlib', LoggerFactory=lambda *a, **k: None, filter_by_level=lambda *a, **k: None)
    dev = _mod('structlog.dev', ConsoleRenderer=lambda *a, **k: None)
    _mod('structlog', get_logger=lambda *a, **k: _NullLogger(), configure=lambda *a, **k: None,
         stdlib=stdlib, dev=dev)
    _mod('structlog_round', FloatRounder=lambda *a, **k: None)
    # ray
    mae = _mod('ray.rllib.env.multi_agent_env', MultiAgentEnv=type('MultiAgentEnv', (), {}))
    renv = _mod('ray.rllib.env', multi_agent_env=mae)
    rllib = _mod('ray.rllib', env=renv)
    _mod('ray', rllib=rllib)

    # matplotlib / svgpath2mpl: imported at module import time (constants.py:8-9,85-88), never run on the hot path
    class _Anything:
        def __getattr__(self, k):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __isub__(self, o):
            return self

        def mean(self, *a, **k):
            return 0

    _any = _Anything()
    pyplot = _mod('matplotlib.pyplot')
    pyplot.__getattr__ = lambda k: _any
    cm = _mod('matplotlib.cm')
    cm.__getattr__ = lambda k: _any
    pe = _mod('matplotlib.patheffects')
    pe.__getattr__ = lambda k: _any
    tr = _mod('matplotlib.transforms', Affine2D=lambda *a, **k: _any)
    anim = _mod('matplotlib.animation')
    anim.__getattr__ = lambda k: _any
    _mod('matplotlib', pyplot=pyplot, cm=cm, patheffects=pe, transforms=tr, animation=anim)
    _mod('svgpath2mpl', parse_path=lambda p: _any)
    install._done = True
