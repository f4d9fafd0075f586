"""Trainer-side pieces against fixtures produced by the reference's own classes
(tests/golden/make_golden_trainer.py): normalisers (oracle restatement, bit-exact) and the
checkpoint layout of `MAPPOActorCritic` (parameter names, order, forward pass)."""
import os

import numpy as np
import torch

from _util import GOLDEN_DIR
from oracle.normalization import MeanStdNormalizerOracle, RewardStdNormalizerOracle


def test_oracle_obs_normalizer_is_the_reference():
    g = np.load(os.path.join(GOLDEN_DIR, "normalizers.npz"))
    n = MeanStdNormalizerOracle(shape=g["x"].shape[2:], clip=10, epsilon=1e-8)
    for t in range(g["x"].shape[0]):
        y = n(g["x"][t])
        assert np.array_equal(y, g["y"][t]) and np.array_equal(n.rms.mean, g["mean"][t])
        assert np.array_equal(n.rms.var, g["var"][t]) and n.rms.count == g["count"][t]
    n.read_only = True
    assert np.array_equal(n(g["x_eval"]), g["y_eval"])
    assert np.array_equal(n.rms.mean, g["mean"][-1])          # frozen


def test_oracle_reward_normalizer_is_the_reference():
    g = np.load(os.path.join(GOLDEN_DIR, "normalizers.npz"))
    n = RewardStdNormalizerOracle(gamma=0.99, clip=10, epsilon=1e-8)
    for t in range(g["r"].shape[0]):
        assert np.array_equal(n(g["r"][t], g["d"][t]), g["ry"][t])
        assert float(n.rms.var) == g["rvar"][t]


def test_checkpoint_layout_matches_reference_actor_critic():
    """`DeviceMAPPO.save` writes `agent.ac` under the reference's parameter names; loading the
    reference's weights reproduces its forward pass; optimiser parameter order is the reference's."""
    from marl_gym_pybullet_drones_b200.mappo import ActorCritic
    g = np.load(os.path.join(GOLDEN_DIR, "checkpoint_layout.npz"))
    names = [str(n) for n in g["names"]]
    ac = ActorCritic(obs_dim=72, act_dim=4, num_agents=2, hidden_dim=64)
    sd = ac.reference_state_dict()
    assert sorted(sd.keys()) == sorted(names)
    for n in names:
        assert tuple(sd[n].shape) == g["w:" + n].shape, n
    ac.load_reference_state_dict({n: torch.as_tensor(g["w:" + n]) for n in names})
    obs = torch.as_tensor(g["obs"])
    with torch.no_grad():
        mean = ac.actor(obs.reshape(-1, 72)).reshape(5, 2, 4)
        value = ac.value(obs.reshape(5, 144))
    assert np.allclose(mean.numpy(), g["mean"], atol=1e-6) and np.allclose(value.numpy(), g["value"], atol=1e-6)
    # Adam state is indexed by parameter position: same order as MLPActor / CentralizedCritic.parameters()
    own = {id(p): k for k, p in ac.named_parameters()}
    order = [own[id(p)] for p in ac.actor_parameters()]
    ref = [str(n) for n in g["actor_param_order"]]
    assert [o.replace("actor.", "pi_net.") if o != "logstd" else o for o in order] == ref
    assert [k for k, _ in ac.critic.named_parameters()] == [str(n).replace("v_net.", "") for n in g["critic_param_order"]]
    # round trip through the repo's own key layout still loads
    ac2 = ActorCritic(72, 4, 2, 64)
    ac2.load_reference_state_dict(ac.state_dict())
    assert torch.equal(ac2.logstd, ac.logstd)
