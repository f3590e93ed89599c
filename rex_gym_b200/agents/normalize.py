"""StreamingNormalize (rex_gym/agents/ppo/normalize.py:22-144) -- a view on the filter state that lives on the device
inside the agent handle; transform is fused into ForwardGaussianPolicy.perform, update into .experience."""


class StreamingNormalize(object):
    def __init__(self, network, which):
        self._net, self._which = network, which

    @property
    def count(self):
        return self._net.get_filters()[self._which + "_count"]

    @property
    def mean(self):
        return self._net.get_filters()[self._which + "_mean"]

    @property
    def var_sum(self):
        return self._net.get_filters()[self._which + "_var_sum"]

    def reset(self):                                        # normalize.py:101-111
        import numpy as np
        f = self._net.get_filters()
        if self._which == "observ":
            self._net.set_filters(0, np.zeros_like(f["observ_mean"]), np.zeros_like(f["observ_var_sum"]), f["reward_count"], f["reward_mean"], f["reward_var_sum"])
        else:
            self._net.set_filters(f["observ_count"], f["observ_mean"], f["observ_var_sum"], 0, 0.0, 0.0)
