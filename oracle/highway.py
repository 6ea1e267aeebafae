"""oracle/highway.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``oracle/highway_oracle.c`` (the fp64 CPU restatement of
highway-v0 as configured by the reference's ``config/base_config.py:5-39`` and
driven at ``training/routine.py:127,134``).  PARITY UNPINNED for the simulator:
upstream highway-env 1.10.1 (reference ``uv.lock:163-175``) is not available in
this build, see ``highway_oracle.h``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Any, Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "highway_oracle.c")
_LIB = os.path.join(_HERE, "libhighway_oracle.so")
_LIB_LEAN = os.path.join(_HERE, "libhighway_oracle_lean.so")   # CPU-baseline build: no decision bookkeeping

HW_MAX_VEHICLES = 128
HW_MAX_FEATURES = 8
FEATURE_CODES = {"presence": 0, "x": 1, "y": 2, "vx": 3, "vy": 4, "heading": 5, "cos_h": 6, "sin_h": 7}


class HwCfg(C.Structure):
    _fields_ = [
        ("lanes_count", C.c_int32), ("vehicles_count", C.c_int32),
        ("simulation_frequency", C.c_int32), ("policy_frequency", C.c_int32),
        ("initial_lane_id", C.c_int32), ("ego_mode", C.c_int32),
        ("normalize_reward", C.c_int32), ("offroad_terminal", C.c_int32),
        ("duration", C.c_double), ("ego_spacing", C.c_double), ("vehicles_density", C.c_double),
        ("collision_reward", C.c_double), ("right_lane_reward", C.c_double),
        ("high_speed_reward", C.c_double), ("reward_speed_lo", C.c_double),
        ("reward_speed_hi", C.c_double),
        ("obs_vehicles", C.c_int32), ("obs_nfeat", C.c_int32),
        ("obs_feat", C.c_int32 * HW_MAX_FEATURES), ("obs_has_range", C.c_int32 * HW_MAX_FEATURES),
        ("obs_lo", C.c_double * HW_MAX_FEATURES), ("obs_hi", C.c_double * HW_MAX_FEATURES),
        ("obs_normalize", C.c_int32), ("obs_clip", C.c_int32), ("obs_absolute", C.c_int32),
        ("obs_sorted", C.c_int32), ("obs_see_behind", C.c_int32), ("_pad", C.c_int32),
    ]


class HwState(C.Structure):
    _fields_ = [(n, C.c_double * HW_MAX_VEHICLES) for n in
                ("x", "y", "heading", "speed", "target_speed", "delta", "timer", "impact_x", "impact_y")] + \
               [(n, C.c_int32 * HW_MAX_VEHICLES) for n in ("lane", "target_lane", "crashed", "has_impact")] + \
               [("time", C.c_double), ("steps", C.c_int64)]


STATE_F64 = ("x", "y", "heading", "speed", "target_speed", "delta", "timer", "impact_x", "impact_y")
STATE_I32 = ("lane", "target_lane", "crashed", "has_impact")


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (plain C, OpenMP for the batch driver)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(
            os.path.getmtime(_SRC), os.path.getmtime(os.path.join(_HERE, "highway_oracle.h"))):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", _LIB, _SRC, "-lm"])
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-DHW_NO_DECISION_AIDS", "-o", _LIB_LEAN, _SRC, "-lm"])
    return _LIB


def use_lean_build() -> None:
    """Switch this process to the CPU-baseline build of the oracle (bench.py's CPU arms): the same algorithm compiled
    without the parity tests' decision bookkeeping.  Must be called before the first OracleEnv is created."""
    global _lib
    build()
    assert _lib is None or getattr(_lib, "_lean", False), "the oracle library is already loaded"
    if _lib is None:
        _lib = _bind(C.CDLL(_LIB_LEAN))
        _lib._lean = True


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _bind(C.CDLL(build()))
    return _lib


def _bind(L):
    if True:
        L.hw_create.restype = C.c_void_p
        L.hw_create.argtypes = [C.POINTER(HwCfg)]
        L.hw_destroy.argtypes = [C.c_void_p]
        L.hw_num_vehicles.argtypes = [C.c_void_p]
        L.hw_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]
        L.hw_get_state.argtypes = [C.c_void_p, C.POINTER(HwState)]
        L.hw_set_state.argtypes = [C.c_void_p, C.POINTER(HwState)]
        L.hw_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32),
                              C.POINTER(C.c_int32)]
        L.hw_observe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hw_shuffle_perm.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int32, C.c_void_p]
        L.hw_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hw_batch_step.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_uint64, C.c_int32, C.c_int32]
        L.hw_last_min_margin.restype = C.c_double
        L.hw_last_min_margin.argtypes = [C.c_void_p]
        L.hw_record_margin.argtypes = [C.c_void_p, C.c_double]
        L.hw_marginal_keys.restype = C.c_int32
        L.hw_marginal_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.hw_force_decisions.restype = C.c_int32
        L.hw_force_decisions.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.hw_slow_vehicles.argtypes = [C.c_void_p, C.c_void_p]
        L.hw_set_trace.argtypes = [C.c_void_p, C.c_void_p]
    return L


# highway-env defaults the reference config does not override
_ENV_DEFAULTS: Dict[str, Any] = {
    "lanes_count": 4, "vehicles_count": 50, "controlled_vehicles": 1, "initial_lane_id": None,
    "duration": 40, "ego_spacing": 2, "vehicles_density": 1, "collision_reward": -1,
    "right_lane_reward": 0.1, "high_speed_reward": 0.4, "lane_change_reward": 0,
    "reward_speed_range": [20, 30], "normalize_reward": True, "offroad_terminal": False,
    "simulation_frequency": 15, "policy_frequency": 1,
}


def cfg_from_dict(cfg: Dict[str, Any]) -> HwCfg:
    """highway-env config dict -> hw_cfg, with HighwayEnv.default_config() filled in."""
    full = dict(_ENV_DEFAULTS)
    full.update(cfg)
    obs = full.get("observation", {"type": "Kinematics"})
    act = full.get("action", {"type": "DiscreteMetaAction"})
    c = HwCfg()
    c.lanes_count = int(full["lanes_count"])
    c.vehicles_count = int(full["vehicles_count"])
    c.simulation_frequency = int(full["simulation_frequency"])
    c.policy_frequency = int(full["policy_frequency"])
    c.initial_lane_id = -1 if full["initial_lane_id"] is None else int(full["initial_lane_id"])
    c.ego_mode = 0 if act.get("type") == "ContinuousAction" else 1
    c.normalize_reward = int(bool(full["normalize_reward"]))
    c.offroad_terminal = int(bool(full["offroad_terminal"]))
    c.duration = float(full["duration"])
    c.ego_spacing = float(full["ego_spacing"])
    c.vehicles_density = float(full["vehicles_density"])
    c.collision_reward = float(full["collision_reward"])
    c.right_lane_reward = float(full["right_lane_reward"])
    c.high_speed_reward = float(full["high_speed_reward"])
    c.reward_speed_lo, c.reward_speed_hi = map(float, full["reward_speed_range"])
    feats = list(obs.get("features", ["presence", "x", "y", "vx", "vy"]))
    c.obs_vehicles = int(obs.get("vehicles_count", 5))
    c.obs_nfeat = len(feats)
    # KinematicObservation.normalize_obs default ranges (only when features_range is None)
    w = 4.0 * c.lanes_count
    rng = obs.get("features_range") or {"x": [-200.0, 200.0], "y": [-w, w],
                                        "vx": [-80.0, 80.0], "vy": [-80.0, 80.0]}
    for i, f in enumerate(feats):
        c.obs_feat[i] = FEATURE_CODES[f]
        if f in rng:
            c.obs_has_range[i] = 1
            c.obs_lo[i], c.obs_hi[i] = float(rng[f][0]), float(rng[f][1])
    c.obs_normalize = int(bool(obs.get("normalize", True)))
    c.obs_clip = int(bool(obs.get("clip", True)))
    c.obs_absolute = int(bool(obs.get("absolute", False)))
    c.obs_sorted = int(obs.get("order", "sorted") == "sorted")
    c.obs_see_behind = int(bool(obs.get("see_behind", False)))
    return c


class OracleEnv:
    """One fp64 highway-v0 env. ``get_state``/``set_state`` move dicts of numpy arrays."""

    def __init__(self, cfg: Dict[str, Any]):
        self.cfg = cfg_from_dict(cfg)
        self._h = lib().hw_create(C.byref(self.cfg))
        if not self._h:
            raise ValueError("oracle: config exceeds HW_MAX_VEHICLES / HW_MAX_FEATURES")
        self.V = lib().hw_num_vehicles(self._h)
        self.N = self.cfg.obs_vehicles
        self.F = self.cfg.obs_nfeat

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:   # (module globals are gone at interpreter shutdown)
            try:
                _lib.hw_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def reset(self, seed: int, env_id: int = 0, episode: int = 0) -> None:
        lib().hw_reset(self._h, seed, env_id, episode)

    def get_state(self) -> Dict[str, np.ndarray]:
        s = HwState()
        lib().hw_get_state(self._h, C.byref(s))
        out: Dict[str, Any] = {}
        for n in STATE_F64:
            out[n] = np.ctypeslib.as_array(getattr(s, n))[: self.V].copy()
        for n in STATE_I32:
            out[n] = np.ctypeslib.as_array(getattr(s, n))[: self.V].copy()
        out["time"] = float(s.time)
        out["steps"] = int(s.steps)
        return out

    def set_state(self, st: Dict[str, Any]) -> None:
        s = HwState()
        for n in STATE_F64:
            np.ctypeslib.as_array(getattr(s, n))[: self.V] = np.asarray(st[n], dtype=np.float64)
        for n in STATE_I32:
            np.ctypeslib.as_array(getattr(s, n))[: self.V] = np.asarray(st[n], dtype=np.int32)
        s.time = float(st["time"])
        s.steps = int(st["steps"])
        lib().hw_set_state(self._h, C.byref(s))

    def step(self, action) -> tuple:
        a = np.ascontiguousarray(action, dtype=np.float32)
        r = C.c_double()
        te, tr = C.c_int32(), C.c_int32()
        lib().hw_step(self._h, a.ctypes.data, C.byref(r), C.byref(te), C.byref(tr))
        return float(r.value), bool(te.value), bool(tr.value)

    def observe(self, perm: Optional[np.ndarray] = None, with_rows: bool = False):
        obs = np.zeros((self.N, self.F), dtype=np.float32)
        rows = np.zeros(self.N, dtype=np.int32)
        p = None
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.int32)
            p = perm.ctypes.data
        lib().hw_observe(self._h, obs.ctypes.data, p, rows.ctypes.data)
        return (obs, rows) if with_rows else obs

    def min_margin(self) -> float:
        return float(lib().hw_last_min_margin(self._h))

    # -- either-branch parity aid (highway_oracle.h) ------------------------------------------------
    def record_margin(self, below: float) -> None:
        """List the decisions of the following steps / observations decided by less than ``below``."""
        lib().hw_record_margin(self._h, float(below))

    def marginal(self, with_margins: bool = False):
        """Keys of the decisions listed since the last ``step`` began (``decode_key`` names them)."""
        keys = np.zeros(1024, dtype=np.uint64)
        marg = np.zeros(1024, dtype=np.float64)
        n = min(int(lib().hw_marginal_keys(self._h, keys.ctypes.data, marg.ctypes.data, 1024)), 1024)
        out = [int(k) for k in keys[:n]]
        return (out, [float(m) for m in marg[:n]]) if with_margins else out

    def force(self, keys=()) -> None:
        """Decide the listed keys the other way from now on (empty: back to the fp64 outcome)."""
        k = np.asarray(list(keys), dtype=np.uint64)
        if lib().hw_force_decisions(self._h, k.ctypes.data if len(k) else None, len(k)) != 0:
            raise ValueError("oracle: too many forced decisions")

    def trace(self, on: bool = True) -> Optional[np.ndarray]:
        """Per-frame snapshots of the following steps: the returned [frames, V, 7] float64 array is rewritten by
        every ``step`` (x, y, speed, heading, impact_x, impact_y, flags)."""
        if not on:
            lib().hw_set_trace(self._h, None)
            self._trace = None
            return None
        frames = self.cfg.simulation_frequency // self.cfg.policy_frequency
        self._trace = np.zeros((frames, self.V, 7), dtype=np.float64)
        lib().hw_set_trace(self._h, self._trace.ctypes.data)
        return self._trace

    def slow_vehicles(self) -> np.ndarray:
        out = np.zeros(HW_MAX_VEHICLES, dtype=np.uint8)
        lib().hw_slow_vehicles(self._h, out.ctypes.data)
        return out[: self.V].astype(bool)


DECISION_KINDS = ("?", "closest_lane", "on_band", "on_road", "reachable", "x_order", "not_zero_sign", "mobil_safe",
                  "mobil_gain", "abort_ahead", "abort_gap", "speed_below_1", "speed_index", "speed_limit", "precheck",
                  "sat_now", "sat_will", "sat_axis", "obs_close", "obs_behind", "obs_order")


def key_fields(key: int):
    """(kind, frame, a, b, c) of a decision key."""
    return (key >> 32) & 0xff, (key >> 24) & 0xff, (key >> 16) & 0xff, (key >> 8) & 0xff, key & 0xff


def decode_key(key: int) -> str:
    kind, frame, a, b, c = (key >> 32) & 0xff, (key >> 24) & 0xff, (key >> 16) & 0xff, (key >> 8) & 0xff, key & 0xff
    fr = {250: "pre", 251: "end", 252: "obs"}.get(frame, str(frame))
    name = DECISION_KINDS[kind] if kind < len(DECISION_KINDS) else str(kind)
    return f"{name}[frame {fr}: {a},{b},{c}]"


def shuffle_perm(seed: int, env_id: int, draw: int, n: int) -> np.ndarray:
    perm = np.zeros(n, dtype=np.int32)
    lib().hw_shuffle_perm(seed, env_id, draw, n, perm.ctypes.data)
    return perm


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().hw_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


class OracleVecEnv:
    """n independent oracle envs stepped with OpenMP; the CPU baseline of bench.py."""

    def __init__(self, cfg: Dict[str, Any], num_envs: int, seed: int = 0, nthreads: int = 0):
        self.envs = [OracleEnv(cfg) for _ in range(num_envs)]
        self.n = num_envs
        self.seed = seed
        self.nthreads = nthreads
        self._handles = (C.c_void_p * num_envs)(*[e._h for e in self.envs])
        e0 = self.envs[0]
        self.obs = np.zeros((num_envs, e0.N, e0.F), dtype=np.float32)
        self.reward = np.zeros(num_envs, dtype=np.float32)
        self.terminated = np.zeros(num_envs, dtype=np.uint8)
        self.truncated = np.zeros(num_envs, dtype=np.uint8)
        for i, e in enumerate(self.envs):
            e.reset(seed, i, 0)

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        lib().hw_batch_step(self._handles, self.n, a.ctypes.data, self.obs.ctypes.data,
                            self.reward.ctypes.data, self.terminated.ctypes.data,
                            self.truncated.ctypes.data, self.seed, 1, self.nthreads)
        return self.obs, self.reward, self.terminated, self.truncated
