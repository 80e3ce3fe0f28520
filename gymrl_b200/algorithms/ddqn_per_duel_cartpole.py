"""Double DQN + prioritized replay + dueling head for CartPole-v1 on the B200 engine — same surface as the reference
``algorithms/ddqn_per_duel_cartpole.py`` (Config, DuelingQNetwork, SumTree, PrioritizedReplayBuffer, DDQNPERDuelTrainer).
Everything but the network is `ddqn_per_cartpole` (the two reference files differ only in the network, ref :58-78): one hidden
layer, then the value and advantage streams as ONE [A+1, H] GEMM; Q = V + (A - mean A) is evaluated inside the loss kernel."""
from __future__ import annotations

import torch.nn as nn

from ..mlp import Chain
from ..nn import FlatParams
from .ddqn_per_cartpole import NONE, RELU, Config, DDQNPERTrainer, PrioritizedReplayBuffer, SumTree  # noqa: F401  (re-exported surface)


class DuelingQNetwork(nn.Module):
    """Same module tree / state_dict keys as the reference DuelingQNetwork (ref :58-78)."""

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 256):
        super().__init__()
        self.action_dim, self.hidden_dim = action_dim, hidden_dim
        self.fc1 = nn.Linear(state_dim, hidden_dim)
        self.value_stream = nn.Linear(hidden_dim, 1)
        self.advantage_stream = nn.Linear(hidden_dim, action_dim)

    DUELING = True
    # flat layout: advantage rows then the value row, biases likewise -> one [A+1, H] head (columns :A advantage, column A value)
    PARAM_ORDER = ["fc1.weight", "fc1.bias", "advantage_stream.weight", "value_stream.weight", "advantage_stream.bias", "value_stream.bias"]

    @staticmethod
    def chain(fp: FlatParams, M: int, backward: bool) -> Chain:
        A = fp.views["advantage_stream.weight"][1][0]
        H = fp.views["advantage_stream.weight"][1][1]
        W = fp.span("advantage_stream.weight", "value_stream.weight", A + 1, H)
        b = fp.span("advantage_stream.bias", "value_stream.bias", 1, A + 1).view(A + 1)
        gW = fp.span("advantage_stream.weight", "value_stream.weight", A + 1, H, grad=True)
        gb = fp.span("advantage_stream.bias", "value_stream.bias", 1, A + 1, grad=True).view(A + 1)
        return Chain(fp, [(fp.p("fc1.weight"), fp.p("fc1.bias"), fp.g("fc1.weight"), fp.g("fc1.bias"), RELU), (W, b, gW, gb, NONE)], M, backward)


class DDQNPERDuelTrainer(DDQNPERTrainer):
    NET = DuelingQNetwork


def main():
    config = Config()
    config.num_envs, config.batch_size, config.memory_capacity = 1024, 1024, 1 << 18
    trainer = DDQNPERDuelTrainer(config)
    try:
        trainer.train()
    except KeyboardInterrupt:
        print("\nTraining interrupted.")
    trainer.test()


if __name__ == "__main__":
    main()
