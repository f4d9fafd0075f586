"""GatedAdam (host logic, runs on CPU): torch.optim.Adam's arithmetic and checkpoint format, plus the gate."""
import torch

from marl_gym_pybullet_drones_b200.optim import GatedAdam


def _nets():
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 2))
    b = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 2))
    b.load_state_dict(a.state_dict())
    return a, b


def test_gated_adam_matches_torch_adam_and_its_checkpoint_format():
    a, b = _nets()
    ref = torch.optim.Adam(a.parameters(), lr=3e-3)
    opt = GatedAdam(b.parameters(), lr=3e-3)
    x = torch.randn(16, 5)
    for it in range(25):
        for net, o in ((a, ref), (b, opt)):
            o.zero_grad()
            (net(x) - 1.0).pow(2).mean().backward()
            o.step()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=2e-6)      # same formulas; rounding order differs by ulps per step
    sd, rsd = opt.state_dict(), ref.state_dict()
    assert sorted(sd["state"].keys()) == sorted(rsd["state"].keys())
    assert set(sd["param_groups"][0]) == set(rsd["param_groups"][0]), set(sd["param_groups"][0]) ^ set(rsd["param_groups"][0])
    for i in rsd["state"]:
        assert float(sd["state"][i]["step"]) == float(rsd["state"][i]["step"]) == 25.0
        assert torch.allclose(sd["state"][i]["exp_avg"], rsd["state"][i]["exp_avg"], rtol=1e-5, atol=1e-9)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], rsd["state"][i]["exp_avg_sq"], rtol=1e-5, atol=1e-12)
    # both directions load: a torch Adam continues from GatedAdam's state and vice versa
    ref.load_state_dict(sd)
    a2, b2 = _nets()
    opt2 = GatedAdam(b2.parameters(), lr=1.0)
    opt2.load_state_dict(rsd)
    assert opt2.lr == 3e-3 and float(opt2.step_t) == 25.0


def test_gate_off_changes_nothing_not_even_the_step_count():
    _, b = _nets()
    opt = GatedAdam(b.parameters(), lr=1e-2)
    x = torch.randn(8, 5)
    opt.zero_grad()
    b(x).pow(2).mean().backward()
    opt.step()
    before = (opt.flat.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt.step_t.clone())
    opt.zero_grad()
    (b(x) * float("nan")).mean().backward()          # even a poisoned gradient must not leak through a closed gate
    opt.step(torch.tensor(False))
    for t, u in zip(before, (opt.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_t)):
        assert torch.equal(t, u)
    opt.zero_grad()
    b(x).pow(2).mean().backward()
    opt.step(torch.tensor(True))
    assert float(opt.step_t) == 2.0 and not torch.equal(before[0], opt.flat)
    assert opt.state_dict()["state"][0]["exp_avg"].shape == (7, 5)     # per-parameter views of the flat buffers


def test_empty_state_before_the_first_step_like_torch():
    _, b = _nets()
    opt = GatedAdam(b.parameters())
    assert opt.state_dict()["state"] == {}
    opt.load_state_dict({"state": {}, "param_groups": [{"lr": 0.5}]})
    assert opt.lr == 0.5
