"""The rollout loop of rex_gym/agents/tools/simulate.py:57-123 on the device: per control step

    prevob -> PPOAlgorithm.perform (filter + policy + value + sample)   one kernel   (csrc/rexsim_agent.cu)
           -> BatchEnv.step (RexGymEnv.step for every env)              one kernel   (csrc/rexsim_kernel.cu)
           -> PPOAlgorithm.experience (filter updates) + memory append  one kernel; the "append" is the kernels writing
                                                                        straight into the [T][N] rollout buffers

Environments that finish are restarted inside the step kernel (auto_reset), which is what simulate() does at the top of
the next iteration (`begin_episode`, :117-120): the observation stored after a `done` is the first one of the new episode.
The whole T-step loop is captured once into a CUDA graph and replayed: no host round trip per step.
"""
import torch

from . import utility


class Rollout(object):
    def __init__(self, env, network, horizon, seed=0, training=True, use_graph=True):
        if not getattr(env._cfg, "auto_reset", 0):
            raise ValueError("Rollout needs an env created with auto_reset=True")
        if network.O != env.obs_dim or network.A != env.action_dim:
            raise ValueError("network (%d -> %d) does not fit the env (%d -> %d)" % (network.O, network.A, env.obs_dim, env.action_dim))
        self.env, self.net, self.T, self.seed, self.training = env, network, int(horizon), int(seed), bool(training)
        N, O, A, T, dev = env.num_envs, env.obs_dim, env.action_dim, self.T, env.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.observ = torch.zeros((T, N, O), **f32)
        self.action = torch.zeros((T, N, A), **f32)
        self.mean = torch.zeros((T, N, A), **f32)
        self.logprob = torch.zeros((T, N), **f32)
        self.value = torch.zeros((T + 1, N), **f32)
        self.reward = torch.zeros((T, N), **f32)
        self.done = torch.zeros((T, N), dtype=torch.uint8, device=dev)
        self._cur = torch.zeros((N, O), **f32)             # observation the next action is computed from
        self._graph = None
        import torch.distributed as dist
        if use_graph and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            use_graph = False      # env-sharded: the per-step filter all-reduce runs through c10d; the plain loop is used
        self._use_graph = use_graph
        self.reset()

    def reset(self):
        """simulate(reset=True): restart every environment."""
        idx = torch.arange(self.env.num_envs, device=self.env.device, dtype=torch.int32)
        self._cur.copy_(self.env.reset(idx))

    def _loop(self):
        env, net, T = self.env, self.net, self.T
        for t in range(T):
            net.perform(self._cur, training=self.training, seed=self.seed, step=0, env_offset=env._cfg.env_offset,
                        out=dict(action=self.action[t], mean=self.mean[t], logprob=self.logprob[t], value=self.value[t]),
                        observ_copy=self.observ[t])
            env.step_into(self.action[t], self._cur, self.reward[t], self.done[t])
            if self.training:
                net.experience(self.observ[t], self.reward[t])
        net.perform(self._cur, training=False, out=dict(value=self.value[T]))       # bootstrap value of the last observation

    def collect(self):
        """Run T control steps for every env; returns views of the rollout buffers (valid until the next collect)."""
        if not self._use_graph:
            self._loop()
        else:
            if self._graph is None:
                torch.cuda.synchronize(self.env.device)
                side = torch.cuda.Stream(device=self.env.device)
                side.wait_stream(torch.cuda.current_stream(self.env.device))
                # warm-up outside capture (lazy kernel loading must not happen while capturing); the warm-up advances the
                # environments and the filters, so everything is snapshotted first and put back afterwards
                snap = (self.env.state_dict(), self.net.state_dict(), self._cur.clone())
                with torch.cuda.stream(side):
                    self._loop()
                torch.cuda.current_stream(self.env.device).wait_stream(side)
                torch.cuda.synchronize(self.env.device)
                self.env.load_state_dict(snap[0]); self.net.load_state_dict(snap[1]); self._cur.copy_(snap[2])
                torch.cuda.synchronize(self.env.device)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._loop()
            self._graph.replay()
        return dict(observ=self.observ, action=self.action, mean=self.mean, logprob=self.logprob, value=self.value,
                    reward=self.reward, done=self.done.view(torch.bool))

    def returns_and_advantages(self, discount, lam=1.0):
        """Done-aware discounted return and GAE over the last rollout (utility.gae_segments)."""
        return utility.gae_segments(self.reward, self.value, self.done, discount, lam)
