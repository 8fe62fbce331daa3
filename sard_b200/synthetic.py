"""Synthetic SARD rollouts: transitions and robot-morphology graphs of each env's shape.

MuJoCo cannot run on the build or bench machines, so rollouts are synthesised with the
exact *layout* the reference's sampler produces:

* state  = ``[obs f64[N,D_ctl], edges i64[2,E], stage i64[1], num_nodes i64[1],
  body_ind i64[N], hfield f64[H], obj f64[O], goal f64[G]]``
  (reference ``design_opt/envs/derl_envs/derl_env.py:377-396``)
* graph  = "ant tree": root + ``n_legs`` legs, each leg a chain of depth 1..3; node order is
  root, the legs, then depth-2 nodes, then depth-3 nodes (append order of
  ``derl_env.py:175-182``); edges ``[i,p],[p,i]`` in node order
  (``khrylib/robot/xml_robot.py:624-632``); ``body_ind = int('1'*(depth-1) + str(leg), base)``
  (``xml_robot.py:335-342``, ``derl_env.py:369-375``).
* episode = 5 skeleton-transform steps (stage 0), 1 attribute-transform step (stage 1), then X
  execution steps (stage 2) (``derl_env.py:213-264``).
* action = ``[control(1) | attr(6) | skel(1)]`` per node, unused columns zero
  (``design_opt/models/transform2act_policy.py:370-380``).

The generator is vectorised per episode and emits a :class:`HostRollout` (flat numpy arrays);
``HostRollout.to_traj_batch()`` gives the reference's list-of-states ``TrajBatchDisc`` view.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np

ATTR_DESIGN_DIM = 6
ACTION_DIM = 8  # control 1 + attr 6 + skel 1


@dataclasses.dataclass(frozen=True)
class EnvShape:
    """Observation shape of one SARD task."""
    name: str
    sim_obs_dim: int = 41
    hfield_dim: int = 1
    obj_dim: int = 1
    goal_dim: int = 1
    n_legs: int = 4
    max_body_depth: int = 4

    @property
    def index_base(self) -> int:           # derl_env.py:81
        return max(5, self.n_legs + 1)

    @property
    def attr_fixed_dim(self) -> int:       # derl_env.py:330-359 (depth one-hot + name digits)
        return self.max_body_depth + self.max_body_depth * self.index_base

    @property
    def state_dim(self) -> int:
        return self.attr_fixed_dim + self.sim_obs_dim + ATTR_DESIGN_DIM


ENV_SHAPES: Dict[str, EnvShape] = {
    'point_nav': EnvShape('point_nav', goal_dim=3),
    'patrol': EnvShape('patrol', goal_dim=3),
    'locomotion_ft': EnvShape('locomotion_ft'),
    'locomotion_vt': EnvShape('locomotion_vt', hfield_dim=1410),
    'escape': EnvShape('escape', hfield_dim=1410),
    'manipulation_box': EnvShape('manipulation_box', sim_obs_dim=54, obj_dim=12),
}

# orbits of the D_4 subgroups used to make symmetric morphologies
SYMMETRIC_ORBITS = {
    'H_4': [[1], [2], [3], [4]],
    'H_2': [[1, 3], [2, 4]],
    'K_1': [[1, 2], [3, 4]],
    'H_1_0': [[1, 2, 3, 4]],
}


def _leg_name(leg: int, depth: int) -> str:
    return '1' * (depth - 1) + np.base_repr(leg, 36)


def make_ant_graph(depths: Sequence[int], shape: EnvShape):
    """Topology of one ant tree.

    Returns ``(edges i64[2,E], body_ind i64[N], attr_fixed f64[N,attr_fixed_dim], node_depth)``.
    """
    base = shape.index_base
    names = ['0']
    parent = [-1]
    node_depth = [0]
    prev_level = {}
    for leg, d in enumerate(depths, start=1):
        if d >= 1:
            names.append(_leg_name(leg, 1)); parent.append(0); node_depth.append(1)
            prev_level[leg] = len(names) - 1
    for level in range(2, max(list(depths) + [1]) + 1):
        for leg, d in enumerate(depths, start=1):
            if d >= level:
                names.append(_leg_name(leg, level)); parent.append(prev_level[leg]); node_depth.append(level)
                prev_level[leg] = len(names) - 1
    n = len(names)
    edges = []
    for i in range(1, n):
        edges.append([i, parent[i]])
        edges.append([parent[i], i])
    edges = np.asarray(edges, dtype=np.int64).T.reshape(2, -1)
    body_ind = np.asarray([int(nm, base) for nm in names], dtype=np.int64)
    fixed = np.zeros((n, shape.attr_fixed_dim))
    for i, nm in enumerate(names):
        fixed[i, node_depth[i]] = 1.0
        for j in range(shape.max_body_depth):
            if len(nm) > j:
                fixed[i, shape.max_body_depth + j * base + int(nm[-(j + 1)], 36)] = 1.0
    return edges, body_ind, fixed, np.asarray(node_depth)


@dataclasses.dataclass
class HostRollout:
    """Flat host-side rollout batch; the packed twin of the reference's ``TrajBatchDisc``."""
    shape: EnvShape
    obs: np.ndarray          # [n_total, D_ctl]
    node_ptr: np.ndarray     # i64 [T+1]
    edges: np.ndarray        # i64 [2, e_total], node ids local to each graph
    edge_ptr: np.ndarray     # i64 [T+1]
    stage: np.ndarray        # i64 [T]
    body_ind: np.ndarray     # i64 [n_total]
    hfield: np.ndarray       # [T, H]
    obj: np.ndarray          # [T, O]
    goal: np.ndarray         # [T, G]
    actions: np.ndarray      # [n_total, 8]
    rewards: np.ndarray      # [T]
    masks: np.ndarray        # [T]
    exps: np.ndarray         # [T]

    @property
    def num_transitions(self) -> int:
        return int(self.stage.shape[0])

    @property
    def num_nodes(self) -> np.ndarray:
        return np.diff(self.node_ptr)

    def state(self, t: int) -> list:
        a, b = int(self.node_ptr[t]), int(self.node_ptr[t + 1])
        ea, eb = int(self.edge_ptr[t]), int(self.edge_ptr[t + 1])
        return [self.obs[a:b], self.edges[:, ea:eb], self.stage[t:t + 1],
                np.array([b - a]), self.body_ind[a:b], self.hfield[t], self.obj[t], self.goal[t]]

    def action(self, t: int) -> np.ndarray:
        a, b = int(self.node_ptr[t]), int(self.node_ptr[t + 1])
        return self.actions[a:b]

    def to_traj_batch(self) -> 'SyntheticTrajBatch':
        T = self.num_transitions
        return SyntheticTrajBatch(
            states=[self.state(t) for t in range(T)],
            actions=[self.action(t) for t in range(T)],
            masks=self.masks.copy(), rewards=self.rewards.copy(), exps=self.exps.copy())

    def select(self, idx: np.ndarray) -> 'HostRollout':
        """Sub-batch of transitions ``idx`` (in that order)."""
        idx = np.asarray(idx, dtype=np.int64)
        nn = self.num_nodes[idx]
        ne = np.diff(self.edge_ptr)[idx]
        node_sel = np.concatenate([np.arange(self.node_ptr[t], self.node_ptr[t + 1]) for t in idx]) if len(idx) else np.zeros(0, np.int64)
        edge_sel = np.concatenate([np.arange(self.edge_ptr[t], self.edge_ptr[t + 1]) for t in idx]) if len(idx) else np.zeros(0, np.int64)
        return HostRollout(
            shape=self.shape, obs=self.obs[node_sel],
            node_ptr=np.concatenate([[0], np.cumsum(nn)]).astype(np.int64),
            edges=self.edges[:, edge_sel],
            edge_ptr=np.concatenate([[0], np.cumsum(ne)]).astype(np.int64),
            stage=self.stage[idx], body_ind=self.body_ind[node_sel],
            hfield=self.hfield[idx], obj=self.obj[idx], goal=self.goal[idx],
            actions=self.actions[node_sel], rewards=self.rewards[idx], masks=self.masks[idx],
            exps=self.exps[idx])


@dataclasses.dataclass
class SyntheticTrajBatch:
    """Same attribute set as the reference's ``TrajBatchDisc`` (``design_opt/utils/tools.py:4-16``)."""
    states: List[list]
    actions: List[np.ndarray]
    masks: np.ndarray
    rewards: np.ndarray
    exps: np.ndarray
    next_states: Optional[list] = None


def generate_rollout(shape: EnvShape, num_transitions: int, seed: int = 1234, *,
                     orbits: Optional[List[List[int]]] = None,
                     min_exec: int = 50, max_exec: int = 1000,
                     skel_steps: int = 5, dtype=np.float64,
                     fixed_depth: Optional[int] = None) -> HostRollout:
    """Seeded synthetic rollouts of exactly ``num_transitions`` transitions.

    ``orbits`` lists the leg orbits of the design subgroup: leg depths are equal inside an
    orbit (symmetric morphologies, as the SARD env enforces); default is the trivial subgroup.
    ``fixed_depth`` forces every leg to that depth (fixed-size graphs).
    """
    rng = np.random.default_rng(seed)
    if orbits is None:
        orbits = [[l] for l in range(1, shape.n_legs + 1)]
    D = shape.state_dim
    F = shape.attr_fixed_dim
    S = shape.sim_obs_dim
    cache = {}

    def graph(depths):
        key = tuple(depths)
        if key not in cache:
            cache[key] = make_ant_graph(key, shape)
        return cache[key]

    obs_l, edge_l, nn_l, ne_l, stage_l, body_l = [], [], [], [], [], []
    hf_l, ob_l, go_l, act_l, rew_l, mask_l = [], [], [], [], [], []
    total = 0
    while total < num_transitions:
        final = np.ones(shape.n_legs, dtype=np.int64)
        for orb in orbits:
            d = fixed_depth if fixed_depth is not None else int(rng.integers(1, shape.max_body_depth))
            for leg in orb:
                final[leg - 1] = d
        n_exec = int(rng.integers(min_exec, max_exec + 1))
        design = rng.uniform(-1, 1, size=(1 + shape.n_legs * (shape.max_body_depth - 1), ATTR_DESIGN_DIM))
        ep_len = skel_steps + 1 + n_exec
        # --- skeleton-transform steps: morphology grows towards `final`
        for k in range(skel_steps):
            cur = np.minimum(final, 1 + (k + 1) // 2)
            edges, bind, fixed, _ = graph(cur)
            n = len(bind)
            o = np.zeros((n, D))
            o[:, :F] = fixed
            o[:, F + S:] = design[:n]
            a = np.zeros((n, ACTION_DIM))
            a[1:, -1] = rng.integers(0, 3, size=n - 1)
            obs_l.append(o); edge_l.append(edges); nn_l.append(n); ne_l.append(edges.shape[1])
            stage_l.append(0); body_l.append(bind); act_l.append(a)
            hf_l.append(np.zeros((1, shape.hfield_dim))); ob_l.append(np.zeros((1, shape.obj_dim)))
            go_l.append(np.zeros((1, shape.goal_dim)))
        # --- attribute-transform step
        edges, bind, fixed, _ = graph(final)
        n = len(bind)
        o = np.zeros((n, D)); o[:, :F] = fixed; o[:, F + S:] = design[:n]
        a = np.zeros((n, ACTION_DIM))
        a[:, 1:1 + ATTR_DESIGN_DIM] = np.clip(rng.normal(0.0, 0.1, size=(n, ATTR_DESIGN_DIM)), -1, 1)
        obs_l.append(o); edge_l.append(edges); nn_l.append(n); ne_l.append(edges.shape[1])
        stage_l.append(1); body_l.append(bind); act_l.append(a)
        hf_l.append(np.zeros((1, shape.hfield_dim))); ob_l.append(np.zeros((1, shape.obj_dim)))
        go_l.append(np.zeros((1, shape.goal_dim)))
        # --- execution steps (vectorised)
        design2 = rng.uniform(-1, 1, size=(n, ATTR_DESIGN_DIM))
        o = np.zeros((n_exec, n, D))
        o[:, :, :F] = fixed
        o[:, :, F:F + S] = rng.normal(size=(n_exec, n, S))
        o[:, :, F + S:] = design2
        a = np.zeros((n_exec, n, ACTION_DIM))
        a[:, :, 0] = rng.normal(size=(n_exec, n))
        obs_l.append(o.reshape(n_exec * n, D))
        act_l.append(a.reshape(n_exec * n, ACTION_DIM))
        edge_l.extend([edges] * n_exec); nn_l.extend([n] * n_exec); ne_l.extend([edges.shape[1]] * n_exec)
        stage_l.extend([2] * n_exec); body_l.extend([bind] * n_exec)
        hf_l.append(rng.normal(size=(n_exec, shape.hfield_dim)) if shape.hfield_dim > 1 else np.zeros((n_exec, 1)))
        ob_l.append(rng.normal(size=(n_exec, shape.obj_dim)) if shape.obj_dim > 1 else np.zeros((n_exec, 1)))
        go_l.append(rng.normal(size=(n_exec, shape.goal_dim)) if shape.goal_dim > 1 else np.zeros((n_exec, 1)))
        rew_l.append(rng.normal(0.01, 0.1, size=ep_len))
        m = np.ones(ep_len); m[-1] = 0.0
        mask_l.append(m)
        total += ep_len

    T = num_transitions
    nn = np.asarray(nn_l, dtype=np.int64)[:T]
    ne = np.asarray(ne_l, dtype=np.int64)[:T]
    node_ptr = np.concatenate([[0], np.cumsum(nn)]).astype(np.int64)
    edge_ptr = np.concatenate([[0], np.cumsum(ne)]).astype(np.int64)
    n_tot, e_tot = int(node_ptr[-1]), int(edge_ptr[-1])
    obs = np.concatenate(obs_l)[:n_tot].astype(dtype)
    actions = np.concatenate(act_l)[:n_tot].astype(dtype)
    edges = np.concatenate(edge_l[:T], axis=1)[:, :e_tot]
    body = np.concatenate(body_l[:T])[:n_tot]
    return HostRollout(
        shape=shape, obs=obs, node_ptr=node_ptr, edges=edges, edge_ptr=edge_ptr,
        stage=np.asarray(stage_l, dtype=np.int64)[:T], body_ind=body,
        hfield=np.concatenate(hf_l)[:T].astype(dtype), obj=np.concatenate(ob_l)[:T].astype(dtype),
        goal=np.concatenate(go_l)[:T].astype(dtype), actions=actions,
        rewards=np.concatenate(rew_l)[:T].astype(dtype), masks=np.concatenate(mask_l)[:T].astype(dtype),
        exps=np.ones(T, dtype=dtype))


def make_random_tree(n_nodes: int, rng, max_index: int = 256, fanout_window: int = 6):
    """A random rooted tree with ``n_nodes`` nodes in the reference's edge layout (``[i,p],[p,i]`` in node order,
    ``xml_robot.py:624-632``); node i hangs off one of the ``fanout_window`` nodes before it, so degrees stay small like
    a robot's.  ``body_ind`` cycles through the IndexLinear table (``i % max_index``)."""
    parent = np.zeros(n_nodes, dtype=np.int64)
    for i in range(1, n_nodes):
        parent[i] = int(rng.integers(max(0, i - fanout_window), i))
    child = np.arange(1, n_nodes, dtype=np.int64)
    edges = np.empty((2, 2 * (n_nodes - 1)), dtype=np.int64)
    edges[0, 0::2], edges[1, 0::2] = child, parent[1:]
    edges[0, 1::2], edges[1, 1::2] = parent[1:], child
    return edges, np.arange(n_nodes, dtype=np.int64) % max_index


def generate_sweep_rollout(shape: EnvShape, num_transitions: int, n_nodes: int, seed: int = 1234, *, n_designs: int = 4,
                           max_index: int = 256, dtype=np.float32) -> HostRollout:
    """Scale-sweep rollouts (BASELINE.json config 5: 64-1024-node graphs): ``num_transitions`` EXECUTION-stage
    transitions (stage 2, the stage that makes up > 95 % of a real rollout; the design stages need leg-structured body
    names for the orbit tie, which a random tree does not have) over ``n_designs`` random trees of ``n_nodes`` nodes.
    Episodes are 500 steps long (mask 0 at their last step)."""
    rng = np.random.default_rng(seed)
    D, F, S = shape.state_dim, shape.attr_fixed_dim, shape.sim_obs_dim
    designs = [make_random_tree(n_nodes, rng, max_index) for _ in range(n_designs)]
    T = num_transitions
    which = (np.arange(T) // 500) % n_designs
    obs = np.zeros((T * n_nodes, D), dtype=dtype)
    obs[:, F:F + S] = rng.standard_normal((T * n_nodes, S), dtype=np.float32).astype(dtype)
    obs[:, 0] = 1.0
    actions = np.zeros((T * n_nodes, ACTION_DIM), dtype=dtype)
    actions[:, 0] = rng.standard_normal(T * n_nodes, dtype=np.float32).astype(dtype)
    ne = 2 * (n_nodes - 1)
    edges = np.concatenate([designs[w][0] for w in which], axis=1)
    body = np.concatenate([designs[w][1] for w in which])
    masks = np.ones(T, dtype=dtype)
    masks[499::500] = 0.0
    masks[-1] = 0.0
    ext = lambda d: (rng.standard_normal((T, d), dtype=np.float32).astype(dtype) if d > 1 else np.zeros((T, 1), dtype=dtype))
    return HostRollout(
        shape=shape, obs=obs, node_ptr=np.arange(T + 1, dtype=np.int64) * n_nodes, edges=edges,
        edge_ptr=np.arange(T + 1, dtype=np.int64) * ne, stage=np.full(T, 2, dtype=np.int64), body_ind=body,
        hfield=ext(shape.hfield_dim), obj=ext(shape.obj_dim), goal=ext(shape.goal_dim), actions=actions,
        rewards=rng.normal(0.01, 0.1, size=T).astype(dtype), masks=masks, exps=np.ones(T, dtype=dtype))


class SyntheticEnv:
    """The attributes ``Transform2ActAgent.setup_env`` reads from a DERL env (agent :111-124), for a shape."""

    def __init__(self, shape: EnvShape):
        self.shape = shape
        self.index_base = shape.index_base
        self.attr_fixed_dim = shape.attr_fixed_dim
        self.attr_design_dim = ATTR_DESIGN_DIM
        self.sim_obs_dim = shape.sim_obs_dim
        self.control_action_dim = 1
        self.skel_num_action = 3
        self.sim_obs_hfield_dim = shape.hfield_dim
        self.sim_obs_obj_dim = shape.obj_dim
        self.sim_obs_goal_dim = shape.goal_dim

    def seed(self, seed):
        pass


class SyntheticRolloutEnv(SyntheticEnv):
    """A stand-in simulator with the DERL env's step protocol (``derl_env.py:213-264``): 5 skeleton-transform steps,
    1 attribute-transform step, then ``exec_steps`` execution steps of a fixed symmetric ant tree.  It exists so that
    ``agent.sample()`` / ``agent.optimize()`` -- the host glue around the CUDA path -- can run where MuJoCo cannot;
    rewards are a smooth function of the actions, nothing is simulated."""

    def __init__(self, shape: EnvShape, exec_steps: int = 20, depth: int = 2, seed: int = 0):
        super().__init__(shape)
        self.exec_steps, self.depth = exec_steps, depth
        self.rng = np.random.default_rng(seed)
        self.num = None
        self.t = 0

    def _graph(self):
        legs = self.shape.n_legs
        parents, body, depth_of = [-1], [0], [0]
        prev = {}
        for d in range(1, self.depth + 1):
            for leg in range(1, legs + 1):
                parents.append(0 if d == 1 else prev[leg])
                body.append(int('1' * (d - 1) + str(leg), self.index_base))
                depth_of.append(d)
                prev[leg] = len(parents) - 1
        n = len(parents)
        e = []
        for i in range(1, n):
            e += [[i, parents[i]], [parents[i], i]]
        return n, np.array(e, dtype=np.int64).T.reshape(2, -1), np.array(body, dtype=np.int64), depth_of

    def _state(self, stage: int):
        n, edges, body, depth_of = self._graph()
        sh = self.shape
        obs = np.zeros((n, sh.state_dim))
        for i in range(n):
            obs[i, min(depth_of[i], sh.max_body_depth - 1)] = 1.0
        if stage == 2:
            obs[:, sh.attr_fixed_dim:sh.attr_fixed_dim + sh.sim_obs_dim] = self.rng.standard_normal((n, sh.sim_obs_dim))
        obs[:, -ATTR_DESIGN_DIM:] = self.attr
        return [obs, edges, np.array([stage]), np.array([n]), body, self.rng.standard_normal(sh.hfield_dim),
                self.rng.standard_normal(sh.obj_dim), self.rng.standard_normal(sh.goal_dim)]

    def reset(self, extra_args=None):
        self.num = (extra_args or {}).get('num')
        self.t = 0
        n = self._graph()[0]
        self.attr = np.zeros((n, ATTR_DESIGN_DIM))
        return self._state(0)

    def step(self, action):
        a = np.asarray(action, dtype=np.float64)
        self.t += 1
        if self.t <= 5:
            stage, name, reward = (0 if self.t < 5 else 1), 'skeleton_transform', 0.0
        elif self.t == 6:
            self.attr = np.clip(a[:, 1:1 + ATTR_DESIGN_DIM], -1, 1)
            stage, name, reward = 2, 'attribute_transform', 0.0
        else:
            stage, name = 2, 'execution'
            reward = float(1.0 - np.mean(a[:, 0] ** 2))
        done = self.t >= 6 + self.exec_steps
        info = {'stage': name, 'done': done}
        if done:
            info['done_max_nsteps'] = 1.0
        return self._state(stage), reward, done, info
