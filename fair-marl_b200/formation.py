"""Formation-family scenarios (``nav_fairassign_fairrew_formation_graph`` / ``..._nofairrew_...``) on the device:
tensor-native env over ``fm_formation_*`` (include/fairmarl.h; kernels in csrc/fm_formation.cu).

A first, correctness-first path: device tensors in, device tensors out, the same dict keys
as ``B200GraphVecEnv.step_tensor`` with this family's shapes -- ``obs [B,N,11]`` (scenario ``observation``, :840-1015),
``node_obs [B,N,E,13]`` (``_get_entity_feat_relative``, :1222-1340), ``adj_env [B,E,E]``, ``reward [B,N]``,
``done [B,N]`` (per-agent early done, environment.py:240-242), ``info [B,N,14]``.  The numpy ``ShareVecEnv`` tuple
interface is not wrapped around it yet.  No CPU path: raises without the library or a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, Optional

import numpy as np

from fair_marl_b200 import _lib


@dataclass
class FormationSimConfig:
    """The argparse fields the formation ``Scenario.make_world`` reads (:48-216) that this path honours."""
    num_agents: int = 3
    num_obstacles: int = 3
    world_size: float = 2.0
    max_speed: Optional[float] = 2.0
    collision_rew: float = 5.0
    goal_rew: float = 5.0
    min_dist_thresh: float = 0.05
    min_obs_dist: float = 0.5
    episode_length: int = 25
    fair_rew: float = 1.0
    zeroshift: float = 5.0
    collaborative: bool = False
    fairness_reward: bool = True       # True: ..._fairrew_... scenario; False: ..._nofairrew_...
    auto_reset: bool = True

    @property
    def num_entities(self) -> int:
        return 2 * self.num_agents + self.num_obstacles

    @classmethod
    def from_args(cls, args: Any, **overrides) -> "FormationSimConfig":
        kw = {f: getattr(args, f) for f in cls.__dataclass_fields__ if hasattr(args, f)}
        name = getattr(args, "scenario_name", "nav_fairassign_fairrew_formation_graph")
        if name not in ("nav_fairassign_fairrew_formation_graph", "nav_fairassign_nofairrew_formation_graph"):
            raise NotImplementedError(f"scenario {name!r} is not one of the two formation scenarios this path covers")
        kw["fairness_reward"] = name == "nav_fairassign_fairrew_formation_graph"
        for unsupported in ("num_walls", "num_scripted_agents"):
            if getattr(args, unsupported, 0):
                raise NotImplementedError(f"{unsupported} > 0 is not supported by the formation kernels")
        if getattr(args, "graph_feat_type", "relative") != "relative":
            raise NotImplementedError("the formation kernels emit the relative node features only")
        if hasattr(args, "num_landmarks") and args.num_landmarks != kw.get("num_agents", 3):
            raise ValueError("the formation scenarios need num_landmarks == num_agents")
        kw.update(overrides)
        return cls(**kw)


class B200FormationVecEnv:
    """B formation envs on one GPU.  ``reset_tensor()`` / ``step_tensor(actions int32 [B,N])`` return dicts of CUDA
    tensors (views of buffers owned by this object, overwritten by the next call)."""

    closed = False

    def __init__(self, cfg: FormationSimConfig, num_envs: int, device: int = 0, seed: int = 0, env_offset: int = 0):
        torch = _lib.require_cuda()
        self.torch, self.lib, self.cfg = torch, _lib.load(), cfg
        self.num_envs, self.num_agents, self.num_entities = int(num_envs), cfg.num_agents, cfg.num_entities
        self.device_index, self.device = int(device), torch.device("cuda", int(device))
        c = _lib.FmFormationConfig(
            num_envs=self.num_envs, num_agents=cfg.num_agents, num_obstacles=cfg.num_obstacles,
            episode_length=cfg.episode_length, env_offset=int(env_offset), seed=int(seed) & (2 ** 64 - 1),
            world_size=cfg.world_size, max_speed=cfg.max_speed if cfg.max_speed is not None else -1.0,
            collision_rew=cfg.collision_rew, goal_rew=cfg.goal_rew, min_dist_thresh=cfg.min_dist_thresh,
            min_obs_dist=cfg.min_obs_dist, fair_rew=cfg.fair_rew, zeroshift=cfg.zeroshift,
            fairness_reward=int(cfg.fairness_reward), collaborative=int(cfg.collaborative), auto_reset=int(cfg.auto_reset))
        self._h = C.c_void_p()
        _lib.check(self.lib.fm_formation_create(C.byref(c), self.device_index, C.byref(self._h)), "fm_formation_create")
        B, N, E = self.num_envs, self.num_agents, self.num_entities
        f32 = dict(dtype=torch.float32, device=self.device)
        self._buf = {
            "obs": torch.zeros((B, N, _lib.FORMATION_OBS_DIM), **f32),
            "node_obs": torch.zeros((B, N, E, _lib.FORMATION_NODE_FEAT_DIM), **f32),
            "adj_env": torch.zeros((B, E, E), **f32), "reward": torch.zeros((B, N), **f32),
            "done": torch.zeros((B, N), dtype=torch.uint8, device=self.device),
            "info": torch.zeros((B, N, _lib.INFO_DIM), **f32)}
        b = self._buf
        self._out = _lib.FmOutputs(b["obs"].data_ptr(), b["node_obs"].data_ptr(), b["adj_env"].data_ptr(),
                                   b["reward"].data_ptr(), b["done"].data_ptr(), b["info"].data_ptr())

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def reset_tensor(self, mask=None) -> Dict[str, Any]:
        t = self.torch
        m = None
        if mask is not None:
            m = t.as_tensor(np.asarray(mask) if not t.is_tensor(mask) else mask).to(device=self.device, dtype=t.uint8).contiguous()
            if tuple(m.shape) != (self.num_envs,):
                raise ValueError(f"mask must have shape ({self.num_envs},)")
        _lib.check(self.lib.fm_formation_reset(self._h, m.data_ptr() if m is not None else None, C.byref(self._out),
                                               self._stream()), "fm_formation_reset")
        if m is not None:
            t.cuda.current_stream(self.device).synchronize()          # keep `m` alive until consumed
        return {k: self._buf[k] for k in ("obs", "node_obs", "adj_env")}

    def step_tensor(self, actions) -> Dict[str, Any]:
        t = self.torch
        if not (t.is_tensor(actions) and actions.is_cuda and actions.dtype == t.int32 and actions.is_contiguous()
                and tuple(actions.shape) == (self.num_envs, self.num_agents)):
            raise ValueError(f"actions must be a contiguous int32 CUDA tensor [{self.num_envs},{self.num_agents}]")
        _lib.check(self.lib.fm_formation_step(self._h, actions.data_ptr(), C.byref(self._out), self._stream()),
                   "fm_formation_step")
        out = dict(self._buf)
        out["done"] = self._buf["done"].view(t.bool)
        return out

    # ------------------------------------------------------------------ state
    def _shapes(self):
        B, N, O = self.num_envs, self.num_agents, self.cfg.num_obstacles
        per = {"pos": (B, N, 2), "vel": (B, N, 2), "landmark_pos": (B, N, 2), "obstacle_pos": (B, O, 2),
               "dist_traveled_mean": (B,), "dist_traveled_stddev": (B,), "step": (B,), "episode": (B,)}
        return {name: per.get(name, (B, N)) for name in _lib.FORMATION_STATE_FIELDS}

    def _dtype(self, name):
        t = self.torch
        return t.int32 if name in _lib.FORMATION_STATE_INT_FIELDS else (t.uint8 if name == "status" else t.float32)

    def _struct(self, tensors) -> _lib.FmFormationState:
        st = _lib.FmFormationState()
        for name in _lib.FORMATION_STATE_FIELDS:
            v = tensors.get(name)
            setattr(st, name, v.data_ptr() if v is not None and v.numel() else None)
        return st

    def get_state(self) -> Dict[str, Any]:
        tensors = {n: self.torch.empty(sh, dtype=self._dtype(n), device=self.device) for n, sh in self._shapes().items()}
        st = self._struct(tensors)
        _lib.check(self.lib.fm_formation_get_state(self._h, C.byref(st), self._stream()), "fm_formation_get_state")
        return tensors

    def set_state(self, state: Dict[str, Any]) -> None:
        t, shapes, tensors = self.torch, self._shapes(), {}
        for name, v in state.items():
            if name not in shapes:
                raise KeyError(name)
            x = t.as_tensor(np.asarray(v) if not t.is_tensor(v) else v).to(device=self.device, dtype=self._dtype(name)).contiguous()
            if tuple(x.shape) != shapes[name]:
                raise ValueError(f"state[{name!r}] must have shape {shapes[name]}, got {tuple(x.shape)}")
            tensors[name] = x
        st = self._struct(tensors)
        _lib.check(self.lib.fm_formation_set_state(self._h, C.byref(st), self._stream()), "fm_formation_set_state")
        t.cuda.current_stream(self.device).synchronize()              # keep `tensors` alive until consumed

    def close(self) -> None:
        if self.closed:
            return
        self.torch.cuda.synchronize(self.device)
        self.lib.fm_formation_destroy(self._h)
        self._h, self.closed = None, True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
