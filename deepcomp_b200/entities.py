"""Scenario descriptors with the reference's constructor signatures (deepcomp/env/entities/{map,station,user}.py,
deepcomp/env/util/movement.py).

In the reference these objects ARE the simulation state (one Python object per UE / BS).  Here they only describe the
scenario that ``env_config`` carries into the env constructor -- the state itself lives in HBM slabs -- and, after a
step, expose read-only views of it (``ue.pos``, ``ue.curr_dr``, ``ue.utility``, ``bs.num_conn_ues`` ...) for callers
such as the reference's Simulation / callbacks that read them.  The env facades are
duck-typed: they accept the reference's own Map/Basestation/User objects just as well.
"""
import math
from collections import namedtuple

Point = namedtuple('Point', ['x', 'y'])

SUPPORTED_SHARING = {'max-cap', 'resource-fair', 'rate-fair', 'proportional-fair'}   # util/constants.py:23
SUPPORTED_UTILITIES = {'log', 'step', 'linear'}                                      # util/constants.py:25


class Map:
    """entities/map.py:7-30 (width/height cast to int)"""

    def __init__(self, width, height, min_x=0, min_y=0):
        self.width, self.height = int(width), int(height)
        self.min_x, self.min_y = min_x, min_y
        self.max_x, self.max_y = min_x + self.width, min_y + self.height
        self.diagonal = math.sqrt(self.width ** 2 + self.height ** 2)

    def seed(self, seed=None):
        """The map RNG only feeds UE arrivals (map.py:52-65), which are out of scope."""

    def __repr__(self):
        return f'{self.width}x{self.height}map'


class Basestation:
    """entities/station.py:13-45"""

    def __init__(self, id, pos, sharing_model):
        assert sharing_model in SUPPORTED_SHARING, f"{sharing_model=} not supported. {SUPPORTED_SHARING=}"
        self.id, self.pos, self.sharing_model = id, pos, sharing_model
        self.num_conn_ues = 0
        self.conn_ues = []

    def __repr__(self):
        return str(self.id)


class RandomWaypoint:
    """util/movement.py:82-105"""

    def __init__(self, map, velocity, pause_duration=2, border_buffer=10):
        assert border_buffer > 0, "Border Buffer must be >0 to avoid placing waypoints on or outside map borders."
        self.map, self.init_velocity = map, velocity
        self.pause_duration, self.border_buffer = pause_duration, border_buffer
        self.velocity = self.waypoint = None
        self.pausing, self.curr_pause = False, 0

    def __str__(self):
        return f"RandomWaypoint({self.init_velocity})"


class UniformMovement:
    """util/movement.py:26-45: move_x / move_y per step, numbers or 'slow' / 'fast'; bounces off the map border"""

    def __init__(self, map, move_x=0, move_y=0):
        self.map, self.init_move_x, self.init_move_y = map, move_x, move_y
        self.move_x = self.move_y = None

    def __str__(self):
        return f"UniformMovement({self.move_x}, {self.move_y})"


class User:
    """entities/user.py:12-50"""

    def __init__(self, id, map, pos_x, pos_y, movement, util_func='log', dr_req=1):
        assert util_func in SUPPORTED_UTILITIES, \
            f"Utility function {util_func} not supported. Supported: {SUPPORTED_UTILITIES}"
        self.id, self.map, self.movement, self.util_func, self.dr_req = id, map, movement, util_func, dr_req
        self.init_pos_x, self.init_pos_y = pos_x, pos_y
        self.pos = None
        self.bs_dr = {}
        self.ewma_dr = 0
        self.curr_dr = 0
        self.utility = -20

    def __repr__(self):
        return str(self.id)

    def __eq__(self, other):
        return type(other) is type(self) and self.id == other.id

    def __hash__(self):
        return hash(self.id)


def create_ues(map, num_static_ues, num_slow_ues, num_fast_ues, util_func='log'):
    """util/env_setup.py:145-161: ids "1".."N", static then slow then fast, all at 'random' positions"""
    ue_list = []
    id = 1
    for velocity, count in ((0, num_static_ues), ('slow', num_slow_ues), ('fast', num_fast_ues)):
        for _ in range(count):
            ue_list.append(User(str(id), map, pos_x='random', pos_y='random',
                                movement=RandomWaypoint(map, velocity=velocity), util_func=util_func))
            id += 1
    return ue_list
