"""Drop-in ``Transform2ActValue`` (reference: design_opt/models/transform2act_critic.py)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .engine import ValueEngine
from .nets import GNNSimple, MLP, ObsEncoder, RunningNorm, init_fc_weights
from .packed import Ordering, PackedRollout


class Transform2ActValue(nn.Module):
    """GNN critic with root-node readout; same constructor, ``state_dict`` and ``forward`` contract."""

    def __init__(self, cfg, agent):
        super().__init__()
        self.cfg = cfg
        self.agent = agent
        self.enable_infer = agent.cfg.enable_infer
        if not (cfg.get('design_flag_in_state', False) and cfg.get('onehot_design_flag', False)):
            raise NotImplementedError('the CUDA critic implements design_flag_in_state + onehot_design_flag (derl.yml)')
        self.sym_embed_dim = sum(agent.cfg.num_legs) * 2
        self.state_dim = agent.state_dim + 3 + (self.sym_embed_dim if self.enable_infer else 0)
        self.sim_obs_hfield_dim = agent.sim_obs_hfield_dim
        self.sim_obs_obj_dim = agent.sim_obs_obj_dim
        self.sim_obs_goal_dim = agent.sim_obs_goal_dim
        self.norm = RunningNorm(self.state_dim)
        self.hfield_enc = ObsEncoder(self.sim_obs_hfield_dim)
        self.obj_enc = ObsEncoder(self.sim_obs_obj_dim)
        self.goal_enc = ObsEncoder(self.sim_obs_goal_dim)
        if 'pre_mlp' in cfg or 'gnn_specs' not in cfg or 'mlp' not in cfg:
            raise NotImplementedError('the CUDA critic implements gnn_specs + mlp without pre_mlp (derl.yml)')
        self.gnn = GNNSimple(self.state_dim, cfg['gnn_specs'])
        self.mlp = MLP(self.gnn.out_dim, cfg['mlp'], cfg['htype'])
        self.value_head = nn.Linear(self.mlp.out_dim + 3 * 64, 1)
        init_fc_weights(self.value_head)
        object.__setattr__(self, '_group_ref', agent.policy_net.group if self.enable_infer else None)
        object.__setattr__(self, '_engine', None)

    @property
    def group(self):
        return self._group_ref

    @property
    def engine(self) -> ValueEngine:
        if self._engine is None:
            object.__setattr__(self, '_engine', ValueEngine(self))
        return self._engine

    def prepare(self, opt=None):
        eng = self.engine
        eng.sync(opt)
        eng.set_embeds(self.group.get_cur_subgroup_embeds() if self.enable_infer else None)
        return eng

    def forward(self, x):
        """``x``: list of 8-slot states -> ``[B,1]`` values (critic.py:78-108).  Train mode updates the RunningNorm."""
        dev = self.value_head.weight.device
        eng = self.prepare(eng_opt(self))
        arena = PackedRollout.from_lists(x, device=dev)
        order = Ordering(arena, np.arange(arena.T))
        run = order.run(0, arena.T)
        with torch.cuda.device(dev):
            eng.run(arena, run, 2 if self.training else 0)
        return arena.values.clone().unsqueeze(1)


def eng_opt(module):
    return module._engine.opt if module._engine is not None else None
