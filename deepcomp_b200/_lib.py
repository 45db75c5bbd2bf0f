"""ctypes binding of libdeepcomp_b200.so (include/deepcomp_b200.h).  No CPU fallback: a missing library is an error."""
import ctypes
import os

from . import build as _build

_LIB = None


class DcbConfig(ctypes.Structure):
    _fields_ = [
        ('abi_version', ctypes.c_int32), ('device', ctypes.c_int32), ('kind', ctypes.c_int32),
        ('reward', ctypes.c_int32), ('num_envs', ctypes.c_int32), ('n_ue', ctypes.c_int32), ('n_bs', ctypes.c_int32),
        ('map_width', ctypes.c_int32), ('map_height', ctypes.c_int32), ('episode_length', ctypes.c_int32),
        ('rand_episodes', ctypes.c_int32), ('auto_reset', ctypes.c_int32), ('pause_duration', ctypes.c_int32),
        ('border_buffer', ctypes.c_int32),
        ('host_bs_xy', ctypes.c_void_p), ('host_sharing', ctypes.c_void_p), ('host_velocity', ctypes.c_void_p),
        ('host_init_xy', ctypes.c_void_p), ('host_seeds', ctypes.c_void_p),
    ]


class DcbOutputs(ctypes.Structure):
    _fields_ = [
        ('obs', ctypes.c_void_p), ('reward', ctypes.c_void_p), ('lost_conn', ctypes.c_void_p),
        ('curr_dr', ctypes.c_void_p), ('utility', ctypes.c_void_p), ('sum_utility', ctypes.c_void_p),
        ('obs_stride', ctypes.c_int64), ('reward_stride', ctypes.c_int64), ('lost_conn_stride', ctypes.c_int64),
        ('curr_dr_stride', ctypes.c_int64), ('utility_stride', ctypes.c_int64), ('sum_utility_stride', ctypes.c_int64),
        ('dbg_obs', ctypes.c_void_p), ('dbg_reward', ctypes.c_void_p), ('dbg_snr', ctypes.c_void_p),
        ('dbg_link_rate', ctypes.c_void_p), ('dbg_curr_dr', ctypes.c_void_p), ('dbg_utility', ctypes.c_void_p),
        ('dbg_sum_utility', ctypes.c_void_p),
    ]


class DcbPolicy(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('noop_interval', ctypes.c_int32), ('epsilon', ctypes.c_double),
                ('host_cluster_masks', ctypes.c_void_p), ('host_fixed_action', ctypes.c_void_p),
                ('seed', ctypes.c_uint64), ('calls_before', ctypes.c_int64)]


class DcbObsVariant(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('dr_mode', ctypes.c_int32), ('dr_cutoff', ctypes.c_double),
                ('curr_dr_obs', ctypes.c_int32), ('ues_at_bs_obs', ctypes.c_int32), ('dist_obs', ctypes.c_int32),
                ('next_dist_obs', ctypes.c_int32)]


class DcbStateHost(ctypes.Structure):
    _fields_ = [('pos', ctypes.c_void_p), ('mask', ctypes.c_void_p), ('ewma', ctypes.c_void_p),
                ('movement', ctypes.c_void_p), ('time', ctypes.c_void_p)]


# every symbol include/deepcomp_b200.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    'dcb_abi_version', 'dcb_last_error', 'dcb_create', 'dcb_destroy', 'dcb_reset', 'dcb_observe', 'dcb_step',
    'dcb_step_many', 'dcb_rollout', 'dcb_step_host', 'dcb_check_errors', 'dcb_get_state', 'dcb_set_state', 'dcb_obs_size',
    'dcb_reward_size', 'dcb_algorithmic_bytes_per_env_step', 'dcb_launch_count', 'dcb_launch_geometry',
    'dcb_kernel_name', 'dcb_set_active_ues', 'dcb_get_active_ues', 'dcb_population_event', 'dcb_get_ue_ids', 'dcb_num_joint_actions', 'dcb_test_actions', 'dcb_set_utility',
    'dcb_set_obs_norm', 'dcb_step_many_host', 'dcb_step_no_move', 'dcb_set_uniform_movement', 'dcb_set_obs_variant', 'dcb_set_interference', 'dcb_extend_waypoints',
]

DCB_ABI_VERSION = 1


class DcbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"deepcomp_b200 error {code}: {msg}")
        self.code = code


def lib_path():
    """In-tree library; DCB_LIB_PATH overrides it (A/B experiments with an alternative build)."""
    return os.environ.get('DCB_LIB_PATH') or _build.LIB


def load():
    """Load (building first if the sources are newer and nvcc is available) the C-ABI library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if path == _build.LIB and _build.needs_build():
        try:
            _build.build()
        except (OSError, Exception) as exc:  # noqa: BLE001 -- nvcc missing / compile error
            if not os.path.exists(path):
                raise ImportError(
                    f"libdeepcomp_b200.so is missing and could not be built ({exc}). deepcomp_b200 has no CPU "
                    f"fallback: run `python -m deepcomp_b200.build` on a machine with nvcc.") from exc
    L = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.dcb_abi_version.restype = ctypes.c_int
    L.dcb_last_error.restype = ctypes.c_char_p
    L.dcb_create.argtypes = [ctypes.POINTER(DcbConfig), ctypes.POINTER(vp)]
    L.dcb_destroy.argtypes = [vp]
    L.dcb_destroy.restype = None
    L.dcb_reset.argtypes = [vp, vp, i32, vp]
    L.dcb_observe.argtypes = [vp, ctypes.POINTER(DcbOutputs), vp]
    L.dcb_step.argtypes = [vp, vp, ctypes.POINTER(DcbOutputs), vp]
    L.dcb_step_no_move.argtypes = [vp, vp, ctypes.POINTER(DcbOutputs), vp]
    L.dcb_set_uniform_movement.argtypes = [vp, vp, vp]
    L.dcb_step_many.argtypes = [vp, vp, i32, ctypes.POINTER(DcbOutputs), vp]
    L.dcb_rollout.argtypes = [vp, ctypes.POINTER(DcbPolicy), i32, vp, ctypes.POINTER(DcbOutputs), vp]
    L.dcb_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.dcb_step_many_host.argtypes = [vp, vp, i32, vp, vp, vp, i32, vp]
    L.dcb_extend_waypoints.argtypes = [vp, vp]
    L.dcb_check_errors.argtypes = [vp, vp]
    L.dcb_get_state.argtypes = [vp, ctypes.POINTER(DcbStateHost)]
    L.dcb_set_state.argtypes = [vp, ctypes.POINTER(DcbStateHost)]
    L.dcb_obs_size.argtypes = [vp]
    L.dcb_obs_size.restype = i64
    L.dcb_reward_size.argtypes = [vp]
    L.dcb_reward_size.restype = i64
    L.dcb_algorithmic_bytes_per_env_step.argtypes = [vp]
    L.dcb_algorithmic_bytes_per_env_step.restype = i64
    L.dcb_launch_count.argtypes = [vp]
    L.dcb_launch_count.restype = i64
    L.dcb_launch_geometry.argtypes = [vp] + [ctypes.POINTER(i32)] * 4
    L.dcb_set_active_ues.argtypes = [vp, i32]
    L.dcb_get_active_ues.argtypes = [vp]
    L.dcb_population_event.argtypes = [vp, i32, i32, vp, vp]
    L.dcb_get_ue_ids.argtypes = [vp, vp]
    L.dcb_set_utility.argtypes = [vp, i32, ctypes.c_double]
    L.dcb_set_obs_norm.argtypes = [vp, i32]
    L.dcb_set_obs_variant.argtypes = [vp, ctypes.POINTER(DcbObsVariant)]
    L.dcb_set_interference.argtypes = [vp, i32]
    L.dcb_num_joint_actions.argtypes = [vp]
    L.dcb_num_joint_actions.restype = i64
    L.dcb_test_actions.argtypes = [vp, i32, i64, i64, vp, vp]
    L.dcb_get_active_ues.restype = i32
    L.dcb_kernel_name.argtypes = [vp]
    L.dcb_kernel_name.restype = ctypes.c_char_p
    if L.dcb_abi_version() != DCB_ABI_VERSION:
        raise ImportError(f"libdeepcomp_b200.so ABI {L.dcb_abi_version()} != {DCB_ABI_VERSION}; rebuild")
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise DcbError(rc, load().dcb_last_error().decode())
