"""CPU: sard_b200.group_action reproduces the reference's dihedral tables (golden KATs)."""
import json
import os

import numpy as np
import pytest

from sard_b200.group_action import LearnedGroup
from tests.util import GOLDEN

with open(os.path.join(GOLDEN, 'group_tables.json')) as f:
    TABLES = json.load(f)


@pytest.mark.parametrize('key', sorted(TABLES))
def test_subgroup_tables(key):
    ns = [int(x) for x in key.split(',')]
    n0 = ns[0]
    G = LearnedGroup(ns, init_subgroup_name=f'H_{n0}>0>H_{n0}>{n0}')
    want = TABLES[key]
    for n, g in G.Gs.items():
        for nm, e in g.group_elements.items():
            w = want['elements'][str(n)][nm]
            assert e.map_to.tolist() == w['map_to']
            np.testing.assert_allclose(e.matrix, w['matrix'], atol=1e-12)
            assert e.permutation_inv.tolist() == w['permutation_inv']
        assert sorted(g.subgroups) == sorted(want['groups'][str(n)])
        for nm, sg in g.subgroups.items():
            w = want['groups'][str(n)][nm]
            assert sorted(sg.elements_set) == w['elements'] and sg.size == w['size']
            assert {str(k): [int(x) for x in v] for k, v in sg.orbits.items()} == w['orbits']
            assert list(map(str, sg.orbits)) == list(w['orbits']) or sorted(map(str, sg.orbits)) == sorted(w['orbits'])
            assert sg.elements_indicator.reshape(-1).tolist() == w['indicator']
            assert sorted(g.subgroup_struct[nm]['sub']) == w['sub']
            assert sorted(g.subgroup_struct[nm]['super']) == w['super']
            for r, (flag, vec) in sg.reprentative_set_defaults.items():
                wf, wv = w['set_defaults'][str(r)]
                assert flag == wf
                if wf:
                    np.testing.assert_allclose(vec, wv, atol=1e-12)
            for r, m in sg.orbits_trans_matrix.items():
                np.testing.assert_allclose(m, w['orbits_trans_matrix'][str(r)], atol=1e-12)
    for name, w in want['names'].items():
        G.set_cur_subgroup_name(name)
        p = G._get_parsed_names()
        for k, v in w['parsed'].items():
            assert p[k] == pytest.approx(v) if isinstance(v, float) else p[k] == v
        assert {str(k): [int(x) for x in v] for k, v in G.get_cur_orbits().items()} == w['orbits']
        np.testing.assert_allclose(G.get_cur_subgroup_embeds().numpy(), w['embeds'], atol=1e-7)
        sym = G.get_cur_subgroup_symmetrizer()
        assert len(sym) == len(w['symmetrizer'])
        for a, b in zip(sym, w['symmetrizer']):
            np.testing.assert_allclose(a['matrix'], b['matrix'], atol=1e-12)
            assert a['permutation_inv'].tolist() == b['permutation_inv']
            assert a['ratio'] == pytest.approx(b['ratio'])
        for cnt, key2 in ((1, 'neighbours_count1'), (25, 'neighbours_count25')):
            G.count = cnt
            try:
                got = sorted(G.get_neibor_subgroup_name())
            except AssertionError:
                got = 'AssertionError'
            G.count = 1
            assert got == w[key2], (name, cnt)


def test_survey_known_answers():
    # D_4 subgroup orbits and the two quoted embeddings
    G = LearnedGroup([4], init_subgroup_name='H_4>0>H_4>4')
    orb = {nm: {int(k): list(map(int, v)) for k, v in sg.orbits.items()} for nm, sg in G.Gs[4].subgroups.items()}
    assert orb['H_1'] == {1: [1, 2, 3, 4]} and orb['H_2'] == {1: [1, 3], 2: [2, 4]}
    assert orb['K_0'] == {1: [1], 2: [2, 4], 3: [3]} and orb['K_1'] == {1: [1, 2], 3: [3, 4]}
    assert orb['K_3'] == {1: [1, 4], 2: [2, 3]} and orb['H_2_1'] == {1: [1, 2, 3, 4]}
    assert len(G.Gs[4].subgroups) == 10 + 1 - 1 or len(G.Gs[4].subgroups) == 10
    assert G.get_cur_subgroup_embeds().tolist() == [1, 0, 0, 0, 0, 0, 0, 0]
    G.set_cur_subgroup_name('H_2>1>H_1>4')
    np.testing.assert_allclose(G.get_cur_subgroup_embeds().numpy(), [1 / 3, 1 / 6, 1 / 3, 1 / 6, 0, 0, 0, 0], atol=1e-6)
    assert G.leg_representatives(5).tolist() == [0, 1, 1, 1, 1]


def test_transform_policy_output_matches_reference_goldens():
    """Orbit tie + subgroup symmetrizer of the rollout path (policy.py:421-498), host utility (CPU tensors)."""
    import torch
    from sard_b200 import synthetic
    from sard_b200.transform2act_policy import Transform2ActPolicy
    from tests.util import small_config
    z = np.load(os.path.join(GOLDEN, 'symmetrizer.npz'))
    shape = synthetic.EnvShape('golden_b', sim_obs_dim=4, hfield_dim=1, obj_dim=1, goal_dim=3)

    class A:
        pass
    for i, name in enumerate(z['names']):
        cfg = small_config(str(name))
        agent = A()
        agent.cfg = cfg
        agent.env = synthetic.SyntheticEnv(shape)
        for k in ('attr_fixed_dim', 'sim_obs_dim', 'attr_design_dim', 'control_action_dim', 'skel_num_action',
                  'sim_obs_hfield_dim', 'sim_obs_obj_dim', 'sim_obs_goal_dim'):
            setattr(agent, k, getattr(agent.env, k))
        agent.state_dim = shape.state_dim
        policy = Transform2ActPolicy(cfg.policy_specs, agent)
        x, body = torch.from_numpy(z[f'{i}.x']), z[f'{i}.body']
        env_out, orbit_out = policy.transform_policy_output(x.clone(), body, attr=True, attr_select=True)
        np.testing.assert_allclose(env_out.numpy(), z[f'{i}.env'], atol=1e-12, err_msg=str(name))
        np.testing.assert_allclose(orbit_out.numpy(), z[f'{i}.orbit'], atol=1e-12, err_msg=str(name))
        tie, _ = policy.transform_policy_output(x.clone(), body, attr=True)
        np.testing.assert_allclose(tie.numpy(), z[f'{i}.tie'], atol=0, err_msg=str(name))
        sk, _ = policy.transform_policy_output(torch.from_numpy(z[f'{i}.skel_in']), body)
        assert np.array_equal(sk.numpy(), z[f'{i}.skel_out']), name


def test_reference_rng_mode_consumes_the_same_numpy_draws():
    """With REFERENCE_RNG the constructors advance the global numpy RNG exactly like the reference's (whose randomised
    orientation probes sit between np.random.seed and the first minibatch shuffle); needs the staged reference."""
    import os
    import numpy as np
    import pytest
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip('oracle/_ref not staged (no /root/reference on this machine)')
    ref_runner.install_shims()
    from design_opt.models.group_action import LearnedGroup as RefGroup
    from sard_b200 import group_action as ga
    for ns, name in (([4], 'H_4>0>H_4>4'), ([4], 'K_1>0>K_1>4'), ([3, 4], 'H_3>0>H_3>3')):
        np.random.seed(7)
        RefGroup(ns, init_subgroup_name=name)
        want = np.random.get_state()[1].copy()
        prev = ga.REFERENCE_RNG
        ga.REFERENCE_RNG = True
        try:
            np.random.seed(7)
            ga.LearnedGroup(ns, init_subgroup_name=name)
            got = np.random.get_state()[1].copy()
        finally:
            ga.REFERENCE_RNG = prev
        assert np.array_equal(got, want), (ns, name)
