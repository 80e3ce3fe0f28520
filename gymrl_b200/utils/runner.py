"""utils.runner on the engine (reference utils/runner.py:16-226): BasicConfig, make_env, train, evaluate, test,
BenchMark, log_monitors — same names, arguments and agent protocol (agent.memory / net / choose_action / evaluate /
update / learn_step / save_model / load_model, optional state_norm / reward_scaler).

What changes underneath: make_env returns the CUDA env behind a Gymnasium-shaped single-env view (utils/env.py), and
the Normalization / RewardScaling the loop attaches are the device versions.  TensorBoard and loguru are used when they
import, otherwise scalars go to a no-op writer and messages to `logging`.
"""
from __future__ import annotations

import logging
import sys
import time

import numpy as np
import torch

from . import env as _env
from .buffer import *  # noqa: F401,F403  (the reference re-exports the buffer classes through utils.runner)
from .buffer import ReplayBuffer_on_policy, ReplayBuffer_on_policy_v2
from .normalization import Normalization, RewardScaling

try:
    from loguru import logger
    logger.remove()
    logger.add(sys.stdout, level="INFO", format="<green>{time:YYYY-MM-DD HH:mm:ss}</green> | <level>{message}</level>")
except Exception:  # pragma: no cover
    logging.basicConfig(level=logging.INFO, format="%(asctime)s | %(message)s")
    logger = logging.getLogger("gymrl_b200")

    def _catch(*a, **k):
        return (lambda f: f) if not (a and callable(a[0])) else a[0]
    logger.catch = _catch

try:
    from torch.utils.tensorboard import SummaryWriter
except Exception:  # pragma: no cover
    class SummaryWriter:
        def __init__(self, *a, **k): pass
        def add_scalar(self, *a, **k): pass
        def close(self): pass

np.random.seed(int(time.time()))   # the reference seeds NumPy from the wall clock at import (:12)


class BasicConfig:
    def __init__(self):
        self.render_mode = 'rgb_array'
        self.train_eps, self.test_eps, self.eval_freq = 500, 3, 10
        self.max_steps = 20000
        self.lr, self.gamma, self.lamda = 1e-4, 0.99, 0.95
        self.n_states = self.n_actions = self.action_bound = None
        self.use_atari = self.unwrapped = self.load_model = False
        self.save_freq = 50
        self.use_rnn = self.on_policy = None
        self.save_path = './checkpoints/model.pth'
        self.device = torch.device('cuda')   # the engine has no CPU path

    def show(self):
        print('-' * 30 + 'Parameters' + '-' * 30)
        for k, v in vars(self).items():
            print(k, '=', v)
        print('-' * 60)


def log_monitors(writer, monitors, agent, phase, step):
    for key, value in monitors.items():
        if not np.isnan(value):
            writer.add_scalar(f'{phase}/{key}', value, global_step=step)


def make_env(cfg, **kwargs):
    """gym.make on the device + the cfg fields the reference fills in (:69-77).  Must precede the agent constructor."""
    if getattr(cfg, "use_atari", False):
        raise NotImplementedError("Atari / image envs are outside the B200 path (SURVEY §2.1)")
    env = _env.make(cfg.env_name, render_mode=getattr(cfg, "render_mode", None), **kwargs)
    if getattr(cfg, "unwrapped", False):
        env = env.unwrapped
    logger.info(f'Observation Space = Box{env.observation_space.shape}')
    cfg.state_shape = env.observation_space.shape
    cfg.n_states = int(env.observation_space.shape[0])
    if isinstance(env.action_space, _env.Box):
        cfg.action_bound = env.action_space.high[0]
        cfg.n_actions = int(env.action_space.shape[0])
    else:
        cfg.n_actions = int(env.action_space.n)
    cfg.max_steps = int(env.spec.max_episode_steps or cfg.max_steps)
    return env


def _attach_tools(env, agent, cfg):
    if not hasattr(agent, "state_norm"):
        agent.state_norm = Normalization(shape=env.observation_space.shape)
    if not hasattr(agent, "reward_scaler"):
        agent.reward_scaler = RewardScaling(shape=1, gamma=cfg.gamma)
    mem = agent.memory
    cfg.on_policy = (isinstance(mem, (ReplayBuffer_on_policy, ReplayBuffer_on_policy_v2)) or
                     isinstance(mem, list) and isinstance(mem[0], ReplayBuffer_on_policy))
    cfg.use_rnn = hasattr(agent.net, 'reset_hidden')


def _episode_seed():
    return int(np.random.randint(1, 2 ** 31 - 1))


def train_vec(agent, cfg, env=None):
    """train() for cfg.num_envs = N lockstep env copies on the device (the runner loop, vectorised).

    Same loop as the reference's (utils/runner.py:81-166) with every per-step quantity a device tensor of N rows:
      * env        gymrl_b200.ops.VecEnv (auto-reset; the TRUE next observation is what the buffer sees, like the reference's
                   next_state; the observation after an episode end seeds that copy's next episode);
      * state_norm one running statistic over all copies (batch-merged Welford update per lockstep); the first observation of a
                   new episode updates it only at N = 1 (a masked batch update does not exist) — the one documented difference;
      * reward_scaler one discounted-return accumulator R per copy, reset where that copy's episode ended (ref :107);
      * memory.store((s, a, r, done, dw, logp, v, v')) with [N] rows per call; agent.update() whenever memory.size() >= batch_size.
    The agent's choose_action takes a [N, D] device tensor and returns ([N] actions, [N] log-probs, [N] values) (on-policy) or [N]
    actions (off-policy).  At N = 1 the sequence of operations is train()'s.  Stops after cfg.train_eps finished episodes."""
    from .. import ops
    N = int(cfg.num_envs)
    seed = int(getattr(cfg, "seed", 0) or 0)
    env = env or ops.VecEnv(cfg.env_name, N, seed=seed)
    cfg.state_shape = (env.obs_dim,)
    cfg.n_states = env.obs_dim
    if env.n_actions:
        cfg.n_actions = env.n_actions
    cfg.max_steps = env.max_episode_steps
    dev = torch.device("cuda", torch.cuda.current_device())
    if not hasattr(agent, "state_norm"):
        agent.state_norm = Normalization(shape=(env.obs_dim,))
    if not hasattr(agent, "reward_scaler"):
        agent.reward_scaler = RewardScaling(shape=1, gamma=cfg.gamma, num_envs=N)
    mem = agent.memory
    cfg.on_policy = isinstance(mem, ReplayBuffer_on_policy)
    cfg.use_rnn = False
    stamp = time.strftime("%Y%m%d-%H%M%S")
    writer = SummaryWriter(f'./exp/{cfg.algo_name}_{cfg.env_name.replace("/", "-")}_{stamp}')
    logger.info(f'Start training: {N} lockstep env copies')
    obs = env.reset()
    state = agent.state_norm(obs)
    agent.reward_scaler.reset()
    if cfg.on_policy:
        action, log_prob, value = agent.choose_action(state)
    else:
        action = agent.choose_action(state)
    last_total, step = 0, 0
    max_lock = int(getattr(cfg, "max_locksteps", 0) or (cfg.train_eps * cfg.max_steps))
    while step < max_lock:
        step += 1
        a_dev = action if torch.is_tensor(action) else torch.as_tensor(np.asarray(action), device=dev)
        obs, reward, terminated, truncated, next_obs = env.step(a_dev.to(torch.int32) if env.n_actions else a_dev.float().reshape(N, -1))
        done = (terminated | truncated)
        reward = agent.reward_scaler(reward)
        next_state = agent.state_norm(next_obs)                      # the true s' (updates the statistic, ref :126)
        if cfg.on_policy:
            nxt_action, nxt_log_prob, nxt_value = agent.choose_action(next_state)
            mem.store((state, action, reward, done, terminated, log_prob, value, nxt_value))
        else:
            mem.store((state, action, reward, next_state, done))
            nxt_action = agent.choose_action(next_state)
        # copies whose episode ended continue from their reset observation (ref :111-113 at the next episode start)
        fresh = agent.state_norm(obs, update=(N == 1 and bool(done.any())))
        dmask = done.bool()
        if bool(dmask.any()):
            agent.reward_scaler.reset(done)
            if cfg.on_policy:
                f_action, f_log_prob, f_value = agent.choose_action(fresh)
                nxt_action = torch.where(dmask, f_action, nxt_action)
                nxt_log_prob = torch.where(dmask, f_log_prob, nxt_log_prob)
                nxt_value = torch.where(dmask, f_value, nxt_value)
            else:
                f_action = agent.choose_action(fresh)
                nxt_action = torch.where(dmask.reshape((N,) + (1,) * (f_action.dim() - 1)), f_action, nxt_action)
            next_state = torch.where(dmask[:, None], fresh, next_state)
        state, action = next_state, nxt_action
        if cfg.on_policy:
            log_prob, value = nxt_log_prob, nxt_value
        if mem.size() >= cfg.batch_size:
            log_monitors(writer, agent.update(), agent, 'train', agent.learn_step)
        if step % 50 == 0 or step == max_lock:
            avg, _, total = env.episode_stats(100)
            if total != last_total:
                last_total = total
                log_monitors(writer, {'reward': avg}, agent, 'train', total)
                logger.info(f'Episodes:{total}/{cfg.train_eps}  Avg(100) reward:{avg:.1f}  locksteps:{step}')
            if total >= cfg.train_eps:
                break
    logger.info('Finish training!')
    if hasattr(agent, "save_model"):
        agent.save_model()
    writer.close()
    return env


def train(env, agent, cfg):
    if int(getattr(cfg, "num_envs", 1) or 1) > 1:
        return train_vec(agent, cfg)        # N lockstep copies on the device; `env` (the 1-copy view) is not used
    logger.info('Start training!')
    if cfg.load_model:
        agent.load_model()
    _attach_tools(env, agent, cfg)
    stamp = time.strftime("%Y%m%d-%H%M%S")
    writer = SummaryWriter(f'./exp/{cfg.algo_name}_{cfg.env_name.replace("/", "-")}_{stamp}')
    cfg.show()
    for ep in range(cfg.train_eps):
        ep_reward, ep_step = 0.0, 0
        agent.reward_scaler.reset()
        if cfg.use_rnn:
            agent.net.reset_hidden()
        state, _ = env.reset(seed=_episode_seed())
        state = agent.state_norm(state)
        if cfg.on_policy:
            action, log_prob, value = agent.choose_action(state)
        else:
            action = agent.choose_action(state)
        for _ in range(cfg.max_steps):
            next_state, reward, terminated, truncated, _info = env.step(action)
            done = terminated or truncated
            ep_reward += reward
            ep_step += 1
            reward = agent.reward_scaler(reward)[0]
            next_state = agent.state_norm(next_state)
            if cfg.on_policy:
                # the value of s' comes from a sampled forward whose action is then taken (q17)
                nxt_action, nxt_log_prob, nxt_value = agent.choose_action(next_state)
                item = (state, action, reward, done, terminated, log_prob, value, nxt_value)
                (agent.memory[ep % cfg.batch_size] if cfg.use_rnn else agent.memory).store(item)
                action, log_prob, value = nxt_action, nxt_log_prob, nxt_value
            else:
                agent.memory.store((state, action, reward, next_state, done))
                action = agent.choose_action(next_state)
            state = next_state
            if not cfg.use_rnn and agent.memory.size() >= cfg.batch_size:
                log_monitors(writer, agent.update(), agent, 'train', agent.learn_step)
            if done:
                break
        if cfg.use_rnn and ep % cfg.batch_size == 0 and ep > 0:
            log_monitors(writer, agent.update(), agent, 'train', agent.learn_step)
        log_monitors(writer, {'reward': ep_reward, 'step': ep_step}, agent, 'train', ep)
        logger.info(f'Episode:{ep + 1}/{cfg.train_eps}  Reward:{ep_reward:.0f}  Step:{ep_step:.0f}')
        if (ep + 1) % cfg.eval_freq == 0:
            evaluate(env, agent, cfg, {'writer': writer})
        if (ep + 1) % cfg.save_freq == 0:
            agent.save_model()
    logger.info('Finish training!')
    agent.save_model()
    env.close()
    writer.close()


def _greedy_episode(env, agent, cfg):
    ep_reward, ep_step, done = 0.0, 0, False
    state, _ = env.reset(seed=_episode_seed())
    state = agent.state_norm(state, update=False)
    if cfg.use_rnn:
        agent.net.reset_hidden()
    while not done:
        ep_step += 1
        state, reward, terminated, truncated, _ = env.step(agent.evaluate(state))
        state = agent.state_norm(state, update=False)
        ep_reward += reward
        done = terminated or truncated
    return ep_reward, ep_step


def evaluate(env, agent, cfg, tools):
    ep_reward, ep_step = _greedy_episode(env, agent, cfg)
    log_monitors(tools['writer'], {'reward': ep_reward, 'step': ep_step}, agent, 'eval', agent.learn_step)


def test(env, agent, cfg):
    logger.info('Start test!')
    agent.load_model()
    if not hasattr(agent, "state_norm") or cfg.use_rnn is None:
        _attach_tools(env, agent, cfg)
    for i in range(cfg.test_eps):
        ep_reward, ep_step = _greedy_episode(env, agent, cfg)
        logger.info(f'Episode:{i + 1}/{cfg.train_eps}  Reward:{ep_reward:.0f}  Step:{ep_step:.0f}')
    logger.info('Finish test!')
    env.close()


class BenchMark:
    @staticmethod
    def train(algo, config):
        cfg = config()
        env = make_env(cfg)     # before the agent: the agent constructor reads cfg.n_states / n_actions
        agent = algo(cfg)
        train(env, agent, cfg)

    @staticmethod
    def test(algo, config):
        cfg = config()
        cfg.render_mode = 'human'
        env = make_env(cfg)
        agent = algo(cfg)
        test(env, agent, cfg)
