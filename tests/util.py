"""Helpers shared by the tests: golden-case loading and oracle plumbing."""
import json
import os

import numpy as np
import torch

from sard_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ('trivial', 'h2_alpha2', 'k1')


class Case:
    """One golden case (tests/golden/case_*.npz) written by make_golden.py from the reference."""

    def __init__(self, tag):
        z = np.load(os.path.join(GOLDEN, f'case_{tag}.npz'))
        self.z = {k: z[k] for k in z.files}
        self.meta = json.loads(bytes(self.z['meta.json']).decode())
        sh = self.meta['shape']
        self.shape = synthetic.EnvShape(**sh)
        f64 = lambda k: self.z['roll.' + k].astype(np.float64)
        self.roll = synthetic.HostRollout(
            shape=self.shape, obs=f64('obs'), node_ptr=self.z['roll.node_ptr'], edges=self.z['roll.edges'],
            edge_ptr=self.z['roll.edge_ptr'], stage=self.z['roll.stage'], body_ind=self.z['roll.body_ind'],
            hfield=f64('hfield'), obj=f64('obj'), goal=f64('goal'), actions=f64('actions'),
            rewards=f64('rewards'), masks=f64('masks'), exps=np.ones(self.meta['T']))
        self.orbits = {}
        for r, m in self.z['meta.orbit_rep']:
            self.orbits.setdefault(int(r), []).append(int(m))
        self.embeds = torch.from_numpy(self.z['meta.embeds'].astype(np.float64))

    def sd(self, prefix, dtype=torch.float64, requires_grad=True):
        """State dict stored under ``prefix`` ('init.policy.', 'upd.value.', ...) as torch tensors."""
        out = {}
        for k, v in self.z.items():
            if not k.startswith(prefix):
                continue
            t = torch.from_numpy(np.array(v))
            name = k[len(prefix):]
            if t.is_floating_point():
                t = t.to(dtype)
                is_buffer = name.split('.')[-1] in ('mean', 'var', 'std') and 'norm' in name
                if requires_grad and not is_buffer:
                    t.requires_grad_(True)
            out[name] = t
        return out

    def get(self, key):
        return self.z[key].astype(np.float64) if self.z[key].dtype == np.float32 else self.z[key]

    def spec(self):
        from oracle import sard_oracle as O
        return O.Spec(attr_fixed_dim=self.shape.attr_fixed_dim, index_base=self.shape.index_base,
                      embeds=self.embeds, orbits=self.orbits)

    def states(self, dtype=torch.float64):
        out = []
        for t in range(self.roll.num_transitions):
            s = self.roll.state(t)
            out.append([torch.tensor(x, dtype=dtype if i != 1 else torch.long) if i in (0, 1, 5, 6, 7) else x
                        for i, x in enumerate(s)])
        return out

    def actions(self, dtype=torch.float64):
        return [torch.tensor(self.roll.action(t), dtype=dtype) for t in range(self.roll.num_transitions)]


def warm_norms(case, psd, vsd):
    """Load the reference's RunningNorm buffers after its warm-up pass."""
    for k, v in case.z.items():
        if k.startswith('warm.policy.'):
            psd[k[len('warm.policy.'):]].copy_(torch.from_numpy(np.array(v)).to(psd[k[len('warm.policy.'):]].dtype))
        if k.startswith('warm.value.'):
            vsd[k[len('warm.value.'):]].copy_(torch.from_numpy(np.array(v)).to(vsd[k[len('warm.value.'):]].dtype))


def assert_close(a, b, rtol, atol, what=''):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f'{what}: max violation at {i}: got {a[i]!r} want {b[i]!r} '
                             f'(|err|={err[i]:.3e}, tol={tol[i]:.3e}); max|err|={err.max():.3e}')


# ---------------------------------------------------------------------------------------------
# product-side helpers (GPU tests)
# ---------------------------------------------------------------------------------------------
def small_config(subgroup, mini_batch=32, n_epochs=2, T=96, max_index=40):
    """The configuration make_golden.py used (small_cfg there), as a sard_b200 Config."""
    from sard_b200.config import Config
    gnn = {'layer_type': 'graph_conv', 'hdims': [16, 16, 16], 'aggr': 'add', 'bias': True}
    imlp = {'hdims': [16, 16], 'rescale_linear': True, 'max_index': max_index}
    d = {
        'agent_specs': {'batch_design': True}, 'gamma': 0.995, 'tau': 0.95,
        'policy_specs': {'htype': 'tanh', 'skel_gnn_specs': gnn, 'skel_index_mlp': imlp, 'control_gnn_specs': gnn,
                         'control_index_mlp': imlp, 'attr_gnn_specs': gnn, 'attr_index_mlp': imlp,
                         'control_log_std': 0, 'attr_log_std': -2.3, 'fix_control_std': False, 'fix_attr_std': False},
        'value_specs': {'htype': 'tanh', 'design_flag_in_state': True, 'onehot_design_flag': True, 'mlp': [32, 16],
                        'gnn_specs': gnn},
        'policy_lr': 5e-5, 'value_lr': 3e-4, 'clip_epsilon': 0.2, 'min_batch_size': T, 'mini_batch_size': mini_batch,
        'num_optim_epoch': n_epochs, 'robot': {'joint_params': {}}, 'init_subgroup_name': subgroup,
    }
    return Config(d)


def build_agent(case, device, prefix='init'):
    """sard_b200 agent with the golden case's weights loaded."""
    import torch
    from sard_b200 import synthetic
    from sard_b200.transform2act_agent import Transform2ActAgent
    cfg = small_config(case.meta['subgroup'], case.meta['mini_batch'], case.meta['n_epochs'], case.meta['T'])
    ag = Transform2ActAgent(cfg, torch.float32, device, 0, 1, env=synthetic.SyntheticEnv(case.shape))
    load_sd(ag.policy_net, case.sd(prefix + '.policy.', torch.float32, requires_grad=False))
    load_sd(ag.value_net, case.sd(prefix + '.value.', torch.float32, requires_grad=False))
    return ag


def load_sd(module, sd):
    own = module.state_dict()
    missing = [k for k in own if k not in sd]
    extra = [k for k in sd if k not in own]
    assert not missing and not extra, (missing, extra)
    module.load_state_dict(sd)


def assert_mostly_close(a, b, rtol, atol, frac=0.999, hard_atol=None, what=''):
    """Like assert_close for at least ``frac`` of the elements; every element within ``hard_atol``.

    Used for weights after Adam steps: an element whose gradient is at fp32 noise level can move by
    +-lr in either direction at the first step, which is inherent to Adam, not to the kernel."""
    a = np.asarray(a, dtype=np.float64).reshape(-1); b = np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape, f'{what}: shape'
    err = np.abs(a - b)
    ok = err <= atol + rtol * np.abs(b)
    assert ok.mean() >= frac, f'{what}: only {ok.mean():.5f} of elements within tolerance (max err {err.max():.3e})'
    if hard_atol is not None:
        assert err.max() <= hard_atol, f'{what}: max err {err.max():.3e} > {hard_atol:.3e}'
