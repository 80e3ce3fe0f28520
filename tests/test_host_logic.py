"""Host-side logic that needs no GPU: the flat parameter layout the kernels rely on, and the bookkeeping of the
captured-lockstep mixin (which graph a lockstep replays, what the host mirrors of the device counters read)."""
import pytest
import torch
import torch.nn as nn


def test_flat_params_layout_alignment_spans_and_counts():
    from gymrl_b200.nn import FlatParams

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = nn.Linear(3, 6)        # 18 + 6 elements
            self.h1 = nn.Linear(6, 2)       # 12 elements: a multiple of 4, so the next matrix follows without padding
            self.h2 = nn.Linear(6, 1)

    net = Net()
    order = ["a.weight", "a.bias", "h1.weight", "h2.weight", "h1.bias", "h2.bias"]
    fp = FlatParams(net, order, torch.device("cpu"))
    for name, (off, shape) in fp.views.items():
        if len(shape) >= 2:
            assert off % 4 == 0, name                                   # float4 operand loads
    assert fp.n_params() == sum(p.numel() for p in net.parameters()) == 18 + 6 + 12 + 6 + 2 + 1
    assert fp.numel() % 4 == 0 and fp.numel() >= fp.n_params()
    # sibling heads stacked as one [3, 5] GEMM operand / one bias vector (the layout the trainers' engines use)
    W = fp.span("h1.weight", "h2.weight", 3, 6)
    b = fp.span("h1.bias", "h2.bias", 1, 3).view(3)
    assert torch.equal(W[:2], net.h1.weight.data) and torch.equal(W[2:], net.h2.weight.data)
    assert torch.equal(b[:2], net.h1.bias.data) and torch.equal(b[2:], net.h2.bias.data)
    # parameters are views of the flat buffer: an in-place update of the buffer is what the module sees
    fp.flat.add_(1.0)
    assert torch.equal(net.a.weight.data, fp.p("a.weight"))
    # gradients likewise
    fp.grad.fill_(2.0)
    assert float(net.h2.bias.grad.item()) == 2.0
    with pytest.raises(AssertionError):
        fp.span("a.weight", "h1.weight", 5, 6)                           # a.bias lies in between: not one matrix

    class Odd(nn.Module):                                                # 2 x 5 = 10 elements: padding before the next matrix
        def __init__(self):
            super().__init__()
            self.h1, self.h2 = nn.Linear(5, 2), nn.Linear(5, 1)
    fo = FlatParams(Odd(), ["h1.weight", "h2.weight", "h1.bias", "h2.bias"], torch.device("cpu"))
    assert fo.views["h2.weight"][0] == 12 and fo.numel() > fo.n_params()
    with pytest.raises(AssertionError):
        fo.span("h1.weight", "h2.weight", 3, 5)                          # the engine must not treat padded siblings as one GEMM


def test_lockstep_graph_mixin_phases_and_host_mirrors(monkeypatch):
    from gymrl_b200 import graphs

    class FakeGraph:
        n_kernels = 7

        def __init__(self, log, phase):
            self.log, self.phase = log, phase

        def replay(self):
            self.log.append(("replay", self.phase))

    class Mem:
        capacity, _size_host = 100, 0

        def __len__(self):
            return self._size_host

    class T(graphs.LockstepGraphs):
        def __init__(self, phases, use_graph=True):
            self.cfg = type("C", (), {"use_cuda_graph": use_graph})()
            self.memory, self.B, self.N = Mem(), 8, 4
            self.total_updates = self.act_count = 0
            self.phases, self.log = phases, []

        def _lockstep_phases(self):
            return self.phases

        def _lockstep_body(self):                 # what the real body does to the host mirrors
            self.act_count += 1
            self.memory._size_host = min(self.memory.capacity, self.memory._size_host + self.N)
            if len(self.memory) >= self.B:
                self.total_updates += 1
            self.log.append(("body", (self.total_updates) % self.phases))

    def fake_capture(fn, warmup=True):
        assert warmup is False                    # the mixin ran the eager pass itself
        fn()                                      # the recording pass bumps the host counters once more ...
        return FakeGraph(t.log, t.log[-1][1])

    monkeypatch.setattr(graphs, "capture", fake_capture)
    t = T(phases=2)
    for _ in range(2):                            # replay not yet full (8 transitions needed): eager locksteps, no graph
        t.lockstep()
    assert t.log == [("body", 0), ("body", 1)] and not getattr(t, "_g_lockstep", None)
    assert (t.act_count, t.total_updates, len(t.memory)) == (2, 1, 8)
    for _ in range(6):
        t.lockstep()
    # ... and the mixin puts them back: every lockstep advances each mirror by exactly one step
    assert (t.act_count, t.total_updates, len(t.memory)) == (8, 7, 32)
    assert sorted(t._g_lockstep) == [0, 1]        # one graph per phase, captured on first use, replayed afterwards
    replays = [e for e in t.log if e[0] == "replay"]
    assert [p for _, p in replays] == [0, 1, 0, 1] and t.graph_launches == 4 * FakeGraph.n_kernels
    # cfg.use_cuda_graph = False keeps everything eager
    e = T(phases=2, use_graph=False)
    for _ in range(5):
        e.lockstep()
    assert not getattr(e, "_g_lockstep", None) and e.act_count == 5


def test_fused_adam_state_dict_is_torch_adam_layout():
    """FusedAdam.state_dict / load_state_dict speak torch.optim.Adam's {state, param_groups} layout, so
    utils.model.ModelLoader round-trips `optimizer_state_dict` with reference checkpoints (ref utils/model.py:337-366)."""
    import torch
    import torch.nn as nn
    from gymrl_b200.nn import FlatParams, FusedAdam
    torch.manual_seed(0)
    mk = lambda: nn.Sequential(nn.Linear(4, 8), nn.Tanh(), nn.Linear(8, 2))
    ours, ref = mk(), mk()
    ref.load_state_dict(ours.state_dict())
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3, eps=1e-5)
    for _ in range(3):
        opt.zero_grad(); ref(torch.randn(5, 4)).sum().backward(); opt.step()
    fa = FusedAdam(FlatParams(ours, ["2.weight", "0.bias", "0.weight", "2.bias"], device="cpu"), lr=1e-3, eps=1e-5)  # flat order != parameter order
    assert fa.state_dict()["state"] == {}                      # lazily created, like torch
    fa.load_state_dict(opt.state_dict())
    sd, rd = fa.state_dict(), opt.state_dict()
    assert sd["state"].keys() == rd["state"].keys() and set(sd["param_groups"][0]) == set(rd["param_groups"][0])
    for i in rd["state"]:
        assert torch.equal(sd["state"][i]["exp_avg"], rd["state"][i]["exp_avg"])
        assert torch.equal(sd["state"][i]["exp_avg_sq"], rd["state"][i]["exp_avg_sq"])
        assert float(sd["state"][i]["step"]) == float(rd["state"][i]["step"]) == 3.0
    torch.optim.Adam(ref.parameters(), lr=5e-4, eps=1e-5).load_state_dict(sd)       # and torch accepts ours
    legacy = {"exp_avg": fa.exp_avg.clone(), "exp_avg_sq": fa.exp_avg_sq.clone(), "step": 7, "lr": 1e-4}    # round-1 private layout
    fa.load_state_dict(legacy)
    assert int(fa.step_t.item()) == 7 and fa.param_groups[0]["lr"] == 1e-4
