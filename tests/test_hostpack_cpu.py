"""CPU: the native host packer (sard_b200/_hostpack.so) reproduces the pure-Python flattening of the
reference's list-of-states batch bit for bit (integer layouts exact, fp64 -> fp32 casts identical)."""
import numpy as np
import pytest
import torch

from sard_b200 import packed, synthetic


@pytest.mark.parametrize('env,dtype', [('manipulation_box', np.float64), ('locomotion_vt', np.float32), ('point_nav', np.float64)])
def test_native_packer_matches_python_packing(env, dtype):
    assert packed._hostpack is not None, 'native packer not built (python -m sard_b200.build)'
    roll = synthetic.generate_rollout(synthetic.ENV_SHAPES[env], 700, seed=4, min_exec=5, max_exec=60, dtype=dtype)
    b = roll.to_traj_batch()
    got = packed.host_tensors_from_lists(b.states, b.actions, b.rewards, b.masks, pin=False, nthreads=3)
    want = packed.host_tensors_from_rollout(packed.host_rollout_from_lists(b.states, b.actions, b.rewards, b.masks), pin=False)
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), k
    # and both equal the generator's flat arrays
    assert np.array_equal(got['node_ptr'].numpy(), roll.node_ptr)
    assert np.array_equal(got['esrc'].numpy(), roll.edges[0]) and np.array_equal(got['edst'].numpy(), roll.edges[1])
    assert np.array_equal(got['obs'].numpy(), roll.obs.astype(np.float32))


def test_native_packer_rejects_inconsistent_states():
    roll = synthetic.generate_rollout(synthetic.ENV_SHAPES['point_nav'], 20, seed=1, min_exec=3, max_exec=5)
    b = roll.to_traj_batch()
    b.states[3][0] = b.states[3][0][:, :-1]          # wrong observation width
    with pytest.raises(ValueError):
        packed.host_tensors_from_lists(b.states, b.actions, b.rewards, b.masks, pin=False)


def test_native_packer_threaded_walk_matches_python_packing():
    """>= 4096 transitions: the object walk runs on several threads (read-only, under the caller's GIL)."""
    roll = synthetic.generate_rollout(synthetic.ENV_SHAPES['manipulation_box'], 5000, seed=6, min_exec=5, max_exec=60, dtype=np.float32)
    b = roll.to_traj_batch()
    got = packed.host_tensors_from_lists(b.states, b.actions, b.rewards, b.masks, pin=False)
    want = packed.host_tensors_from_rollout(packed.host_rollout_from_lists(b.states, b.actions, b.rewards, b.masks), pin=False)
    for k in want:
        assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), k
    # a tuple state and a non-numpy slot send the whole batch down the general path: same result
    b.states[17] = tuple(b.states[17])
    import array
    b.states[18][2] = array.array('q', [int(x) for x in np.asarray(b.states[18][2]).reshape(-1)])     # buffer protocol, not numpy
    got2 = packed.host_tensors_from_lists(b.states, b.actions, b.rewards, b.masks, pin=False)
    for k in want:
        assert torch.equal(got2[k], want[k]), k
