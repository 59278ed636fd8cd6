"""mu_k updaters of the safeguarded evaluation (the reference's `mu_updater_dict[name](mu, param).step(norm, accepted)`
surface, mu_updater.py:18-116).

One table-driven class: per column, mu <- rule(norm, mu) where the column's learned iterate was accepted, unchanged
elsewhere.  `METHOD_ID` are the ids `dladmm_sg_select_update` (csrc/safeguard.cu) takes: the safeguarded forwards fold the
update into the selection kernel and only fall back to these host-side objects for "RM".  The reference's `RMUpdater`
returns a (values, indices) tuple and allocates with a float size (unusable as written); here it is the running maximum
of the last `parameter` norms.  "None" is the reference's BlankUpdater: the threshold becomes 1e10 after the first test.
"""
import torch

_RULES = {
    "EMA": lambda s, mu, p: p * s + (1 - p) * mu,      # exponential moving average
    "GS": lambda s, mu, p: (1 - p) * mu,               # geometric series
    "RT": lambda s, mu, p: s,                          # recent term
}
METHOD_ID = {"EMA": 1, "GS": 2, "RT": 3, "None": 4}   # dladmm_sg_select_update `method`


class _RuleUpdater(object):
    rule = None

    def __init__(self, mu, parameter):
        self.mu, self.parameter = mu, parameter

    def step(self, Sx_L2O_norm, bool_term):
        upd = _RULES[self.rule](Sx_L2O_norm, self.mu, self.parameter)
        self.mu = torch.where(bool_term != 0, upd, self.mu).detach()
        return self.mu


def _rule(name):
    return type(name + "Updater", (_RuleUpdater,), {"rule": name})


class RMUpdater(object):
    def __init__(self, mu, parameter):
        self.window = max(1, int(parameter))
        self.recent = mu.new_zeros((mu.shape[0], self.window))
        self.pointer = 0
        self.step(mu)

    def step(self, Sx_L2O_norm, bool_term=None):
        self.recent[:, self.pointer] = Sx_L2O_norm
        self.pointer = (self.pointer + 1) % self.window
        return self.recent.max(dim=1).values


class BlankUpdater(object):
    def __init__(self, mu, parameter):
        self.like = mu

    def step(self, Sx_L2O_norm, bool_term):
        return torch.full_like(self.like, 1e10)


EMAUpdater, GSUpdater, RTUpdater = _rule("EMA"), _rule("GS"), _rule("RT")
mu_updater_dict = {"EMA": EMAUpdater, "GS": GSUpdater, "RT": RTUpdater, "RM": RMUpdater, "None": BlankUpdater}
