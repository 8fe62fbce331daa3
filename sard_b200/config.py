"""Training configuration with the attribute set of the reference's ``design_opt/utils/config.py``.

``Config(cfg_dict)`` takes the parsed YAML (``design_opt/cfg/derl.yml`` layout); ``derl_config()``
returns the reference's default hyper-parameters (derl.yml:1-69 and config.py:83-99) so the bench
and the tests do not need the reference tree.  No directories are created here.
"""
from __future__ import annotations

import copy

import numpy as np

_GNN = {'layer_type': 'graph_conv', 'hdims': [64, 64, 64], 'aggr': 'add', 'bias': True}
_IMLP = {'hdims': [128, 128], 'rescale_linear': True, 'max_index': 256}

DERL_YML = {
    'env_name': 'derl',
    'agent_specs': {'batch_design': True},
    'gamma': 0.995, 'tau': 0.95,
    'policy_specs': {
        'name': 'v5', 'htype': 'tanh',
        'skel_gnn_specs': _GNN, 'skel_index_mlp': _IMLP,
        'control_gnn_specs': _GNN, 'control_index_mlp': _IMLP,
        'attr_gnn_specs': _GNN, 'attr_index_mlp': _IMLP,
        'control_log_std': 0, 'attr_log_std': -2.3, 'fix_control_std': False, 'fix_attr_std': False,
    },
    'policy_optimizer': 'Adam', 'policy_lr': 5.e-5, 'policy_momentum': 0.0, 'policy_weightdecay': 0.0,
    'value_specs': {'htype': 'tanh', 'design_flag_in_state': True, 'onehot_design_flag': True, 'mlp': [512, 256],
                    'gnn_specs': _GNN},
    'value_optimizer': 'Adam', 'value_lr': 3.e-4, 'value_momentum': 0.0, 'value_weightdecay': 0.0,
    'clip_epsilon': 0.2, 'min_batch_size': 50000, 'mini_batch_size': 2048, 'eval_batch_size': 10000,
    'num_optim_epoch': 10, 'max_epoch_num': 1000, 'seed': 1, 'save_model_interval': 100,
    'robot': {'joint_params': {}},
}


class Config:
    def __init__(self, cfg_dict, cfg_id='derl', enable_infer=True, model_dir=None):
        cfg = copy.deepcopy(cfg_dict)
        self.id = cfg_id
        self.model_dir = model_dir
        self.gamma = cfg.get('gamma', 0.99)
        self.tau = cfg.get('tau', 0.95)
        self.agent_specs = cfg.get('agent_specs', dict())
        self.policy_specs = cfg.get('policy_specs', dict())
        self.policy_optimizer = cfg.get('policy_optimizer', 'Adam')
        self.policy_lr = cfg.get('policy_lr', 5e-5)
        self.policy_momentum = cfg.get('policy_momentum', 0.0)
        self.policy_weightdecay = cfg.get('policy_weightdecay', 0.0)
        self.value_specs = cfg.get('value_specs', dict())
        self.value_optimizer = cfg.get('value_optimizer', 'Adam')
        self.value_lr = cfg.get('value_lr', 3e-4)
        self.value_momentum = cfg.get('value_momentum', 0.0)
        self.value_weightdecay = cfg.get('value_weightdecay', 0.0)
        self.adv_clip = cfg.get('adv_clip', np.inf)
        self.clip_epsilon = cfg.get('clip_epsilon', 0.2)
        self.num_optim_epoch = cfg.get('num_optim_epoch', 10)
        self.min_batch_size = cfg.get('min_batch_size', 50000)
        self.mini_batch_size = cfg.get('mini_batch_size', self.min_batch_size)
        self.eval_batch_size = cfg.get('eval_batch_size', 10000)
        self.max_epoch_num = cfg.get('max_epoch_num', 1000)
        self.seed = cfg.get('seed', 1)
        self.save_model_interval = cfg.get('save_model_interval', 100)
        self.enable_infer = enable_infer
        self.scheduled_params = cfg.get('scheduled_params', dict())
        self.env_name = cfg.get('env_name', 'derl')
        self.robot_cfg = cfg.get('robot', dict())
        self.enable_design = True
        # SARD knobs hard-coded in the reference (config.py:83-99)
        n = cfg.get('n_legs', 4)
        self.num_legs = [n]
        self.init_subgroup_name = cfg.get('init_subgroup_name', f'H_{n}>0>H_{n}>{n}')
        self.updating_subgroup = True
        self.struct_subgroup = True
        self.subgroup_alpha_gap = 3


def derl_config(**overrides) -> Config:
    d = copy.deepcopy(DERL_YML)
    d.update(overrides)
    return Config(d)


def sweep_yml(hidden: int = 256) -> dict:
    """derl.yml with every hidden width set to ``hidden`` (BASELINE.json config 5: "hidden 256")."""
    d = copy.deepcopy(DERL_YML)
    gnn = {'layer_type': 'graph_conv', 'hdims': [hidden] * 3, 'aggr': 'add', 'bias': True}
    imlp = {'hdims': [hidden, hidden], 'rescale_linear': True, 'max_index': 256}
    for st in ('skel', 'control', 'attr'):
        d['policy_specs'][f'{st}_gnn_specs'] = copy.deepcopy(gnn)
        d['policy_specs'][f'{st}_index_mlp'] = copy.deepcopy(imlp)
    d['value_specs']['gnn_specs'] = copy.deepcopy(gnn)
    return d
