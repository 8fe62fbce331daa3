"""CPU: the oracle restatement reproduces the reference's golden vectors (made by make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import sard_oracle as O
from tests.util import CASES, Case, assert_close, warm_norms

TIGHT = dict(rtol=2e-6, atol=2e-6)      # goldens are stored rounded to fp32


@pytest.fixture(scope='module', params=CASES)
def case(request):
    return Case(request.param)


def test_eval_forward_matches_reference(case):
    psd, vsd = case.sd('init.policy.'), case.sd('init.value.')
    st, ac = case.states(), case.actions()
    with torch.no_grad():
        fw = O.policy_forward(psd, case.spec(), st, training=False)
        assert_close(fw['control']['mean'], case.get('eval.control_mean'), what='control mean', **TIGHT)
        assert_close(fw['attr']['mean'], case.get('eval.attr_mean'), what='attr mean', **TIGHT)
        assert_close(torch.log_softmax(fw['skel']['logits'], -1), case.get('eval.skel_logits'), what='skel logits', **TIGHT)
        assert_close(O.policy_log_prob(psd, case.spec(), st, ac, False), case.get('eval.log_prob'), what='log_prob', rtol=2e-6, atol=2e-5)
        assert_close(O.value_forward(vsd, case.spec(), st, False), case.get('eval.value'), what='value', **TIGHT)


def test_running_norm_update_and_warm_forward(case):
    psd, vsd = case.sd('init.policy.'), case.sd('init.value.')
    st, ac = case.states(), case.actions()
    T = len(st)
    with torch.no_grad():
        O.policy_log_prob(psd, case.spec(), st[:T // 2], ac[:T // 2], training=True)
        O.value_forward(vsd, case.spec(), st[:T // 2], training=True)
    for k, v in case.z.items():
        if k.startswith('warm.policy.') and '_norm.' in k:
            assert_close(psd[k[len('warm.policy.'):]], v, what=k, rtol=2e-6, atol=1e-7)
        if k.startswith('warm.value.norm.'):
            assert_close(vsd[k[len('warm.value.'):]], v, what=k, rtol=2e-6, atol=1e-7)
    with torch.no_grad():
        fw = O.policy_forward(psd, case.spec(), st, training=False)
        assert_close(fw['control']['mean'], case.get('warm.control_mean'), what='control mean', rtol=1e-5, atol=1e-5)
        assert_close(fw['attr']['mean'], case.get('warm.attr_mean'), what='attr mean', rtol=1e-5, atol=1e-5)
        assert_close(O.policy_log_prob(psd, case.spec(), st, ac, False), case.get('warm.log_prob'), what='lp', rtol=1e-5, atol=1e-4)
        assert_close(O.value_forward(vsd, case.spec(), st, False), case.get('warm.value'), what='value', rtol=1e-5, atol=1e-5)


def test_gae_matches_reference(case):
    r = torch.from_numpy(case.roll.rewards); m = torch.from_numpy(case.roll.masks)
    v = torch.from_numpy(case.get('warm.value'))
    adv, ret = O.estimate_advantages(r, m, v, 0.995, 0.95)
    assert_close(adv, case.get('gae.advantages'), what='adv', rtol=1e-5, atol=1e-5)
    assert_close(ret, case.get('gae.returns'), what='ret', rtol=1e-5, atol=1e-5)


def test_loss_and_gradients_match_reference(case):
    psd, vsd = case.sd('init.policy.'), case.sd('init.value.')
    warm_norms(case, psd, vsd)
    st, ac = case.states(), case.actions()
    adv = torch.from_numpy(case.get('gae.advantages')); ret = torch.from_numpy(case.get('gae.returns'))
    fixed = torch.from_numpy(case.get('grad.fixed_log_prob'))
    lp, ent = O.policy_log_prob(psd, case.spec(), st, ac, training=True, get_entropy=True)
    loss, kl, valid = O.ppo_loss(lp, fixed, adv, 0.2)
    assert_close(float(loss), case.get('grad.surr_loss'), what='surr', rtol=1e-5, atol=1e-6)
    assert_close(float(ent), case.get('grad.entropy'), what='entropy', rtol=1e-5, atol=1e-6)
    assert_close(kl, case.get('grad.approx_kl'), what='kl', rtol=1e-5, atol=1e-6)
    assert valid == bool(case.z['grad.valid'])
    loss.backward()
    n = 0
    for k, v in case.z.items():
        if k.startswith('grad.policy.'):
            g = psd[k[len('grad.policy.'):]].grad
            assert g is not None, k
            assert_close(g, v, what=k, rtol=1e-4, atol=1e-6 * max(1.0, float(np.abs(v).max())))
            n += 1
    assert n > 30
    vloss = (O.value_forward(vsd, case.spec(), st, training=True) - ret).pow(2).mean()
    assert_close(float(vloss), case.get('grad.value_loss'), what='vloss', rtol=1e-5, atol=1e-6)
    vloss.backward()
    for k, v in case.z.items():
        if k.startswith('grad.value.'):
            assert_close(vsd[k[len('grad.value.'):]].grad, v, what=k, rtol=1e-4, atol=1e-6 * max(1.0, float(np.abs(v).max())))


@pytest.mark.parametrize('tag', ['trivial', 'h2_alpha2'])
def test_update_policy_matches_reference(tag):
    case = Case(tag)
    psd, vsd = case.sd('init.policy.'), case.sd('init.value.')
    warm_norms(case, psd, vsd)
    # the reference's warm pass + grad pass also advanced the norms: replay them
    for k, v in case.z.items():
        if k.startswith('grad.policy_after.'):
            psd[k[len('grad.policy_after.'):]].copy_(torch.from_numpy(np.array(v)).to(psd[k[len('grad.policy_after.'):]].dtype))
        if k.startswith('grad.value_after.'):
            vsd[k[len('grad.value_after.'):]].copy_(torch.from_numpy(np.array(v)).to(vsd[k[len('grad.value_after.'):]].dtype))
    st, ac = case.states(), case.actions()
    adv = torch.from_numpy(case.get('gae.advantages')); ret = torch.from_numpy(case.get('gae.returns'))
    ag = O.OracleAgent(psd, vsd, case.spec(), mini_batch_size=case.meta['mini_batch'], opt_num_epochs=case.meta['n_epochs'])
    logs = ag.update_policy(st, ac, ret, adv, perms=list(case.z['upd.perms']))
    surr = np.array([l['surr_loss'] for l in logs])
    want = float(case.get('upd.surr_loss'))
    assert_close(surr.sum() / (np.count_nonzero(surr) + 1e-6), want, what='mean surr', rtol=1e-3, atol=1e-5)
    for k, v in case.z.items():
        if k.startswith('upd.policy.'):
            assert_close(psd[k[len('upd.policy.'):]].detach(), v, what=k, rtol=1e-4, atol=2e-5)
        if k.startswith('upd.value.'):
            assert_close(vsd[k[len('upd.value.'):]].detach(), v, what=k, rtol=1e-4, atol=2e-5)
