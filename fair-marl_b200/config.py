"""Simulator configuration: the argparse fields ``Scenario.make_world`` reads
(navigation_graph.py:94-129, :144, :188, :208; defaults config.py:176-253, train_mpe.py:72-101)."""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Optional


@dataclass
class SimConfig:
    num_agents: int = 3
    num_obstacles: int = 3
    num_walls: int = 0          # 0, 1 or 2 wall segments (agent-warp kernels at 3 agents / 3 obstacles, else group-per-env; not with graph_feat_type='global')
    world_size: float = 2.0
    max_speed: Optional[float] = 2.0
    collision_rew: float = 5.0
    goal_rew: float = 5.0
    min_dist_thresh: float = 0.05
    episode_length: int = 25
    fair_rew: float = 1.0
    zeroshift: float = 5.0
    max_edge_dist: float = 1.0
    collaborative: bool = False
    # navigation_graph.py (FA+FR) vs nav_graph_goalassign_noFair.py (FA)
    fairness_reward: bool = True
    # 'relative' (11 ego-relative node features) or 'global' (7 absolute ones), navigation_graph.py:941-1035
    graph_feat_type: str = "relative"
    auto_reset: bool = True
    info_every_step: bool = False
    # kernel mapping: 'auto' | 'group' (group-per-env) | 'aw' (agent-warp, compiled for small (N, O));
    # results are identical across mappings
    mapping: str = "auto"

    @property
    def num_entities(self) -> int:
        return 2 * self.num_agents + self.num_obstacles + self.num_walls

    @property
    def node_feat_dim(self) -> int:
        return 11 if self.graph_feat_type == "relative" else 7

    @classmethod
    def from_args(cls, args: Any, **overrides) -> "SimConfig":
        """Build from the reference's flat ``all_args`` Namespace (or any object with those fields)."""
        kw = {}
        for f in fields(cls):
            if hasattr(args, f.name):
                kw[f.name] = getattr(args, f.name)
        if hasattr(args, "num_landmarks") and args.num_landmarks != kw.get("num_agents", 3):
            raise ValueError("navigation_graph needs num_landmarks == num_agents "
                             f"(got {args.num_landmarks} vs {kw.get('num_agents')})")
        if not 0 <= int(getattr(args, "num_walls", 0) or 0) <= 2:
            raise ValueError("navigation_graph places at most 2 walls (navigation_graph.py:289)")
        if getattr(args, "num_walls", 0) and getattr(args, "graph_feat_type", "relative") == "global":
            raise NotImplementedError("wall entities have no global features (navigation_graph.py:1074-1075)")
        if getattr(args, "graph_feat_type", "relative") not in ("relative", "global"):
            raise ValueError(f"graph_feat_type must be 'relative' or 'global' (got {args.graph_feat_type!r})")
        if getattr(args, "num_scripted_agents", 0):
            raise NotImplementedError("scripted agents are not supported")
        scen = getattr(args, "scenario_name", "navigation_graph")
        if scen == "nav_graph_goalassign_noFair":
            kw["fairness_reward"] = False
        elif scen != "navigation_graph":
            raise NotImplementedError(f"scenario {scen!r} is not supported (navigation_graph family only)")
        kw.update(overrides)
        return cls(**kw)
