"""mu_k updaters of the safeguarded evaluation -- mirror of the reference's mu_updater.py:18-116.

Each updater holds a per-column vector mu (B,) and is stepped with the norm of S(u^k) of the learned iterate and the
0/1 acceptance vector.  `RMUpdater` in the reference returns a (values, indices) tuple and allocates with a float
size (mu_updater.py:75-94, unusable as written); here it returns the running maximum over the last `parameter` norms.
"""
import torch


class EMAUpdater(object):
    def __init__(self, mu, parameter):
        self.mu, self.parameter = mu, parameter

    def step(self, Sx_L2O_norm, bool_term):
        update = self.parameter * Sx_L2O_norm + (1 - self.parameter) * self.mu
        self.mu = ((1.0 - bool_term) * self.mu + bool_term * update).detach()
        return self.mu


class GSUpdater(object):
    def __init__(self, mu, parameter):
        self.mu, self.parameter = mu, parameter

    def step(self, Sx_L2O_norm, bool_term):
        update = (1 - self.parameter) * self.mu
        self.mu = ((1.0 - bool_term) * self.mu + bool_term * update).detach()
        return self.mu


class RTUpdater(object):
    def __init__(self, mu, parameter):
        self.mu, self.parameter = mu, parameter

    def step(self, Sx_L2O_norm, bool_term):
        self.mu = ((1.0 - bool_term) * self.mu + bool_term * Sx_L2O_norm).detach()
        return self.mu


class RMUpdater(object):
    def __init__(self, mu, parameter):
        self.parameter = max(1, int(parameter))
        self.recent = mu.new_zeros((mu.shape[0], self.parameter))
        self.pointer = 0
        self.step(mu)

    def step(self, Sx_L2O_norm, bool_term=None):
        self.recent[:, self.pointer] = Sx_L2O_norm
        self.pointer = (self.pointer + 1) % self.parameter
        return self.recent.max(dim=1).values


class BlankUpdater(object):
    """No safeguard threshold: mu_k = 1e10 (mu_updater.py:97-108)."""

    def __init__(self, mu, parameter):
        self.like = mu

    def step(self, Sx_L2O_norm, bool_term):
        return torch.full_like(self.like, 1e10)


mu_updater_dict = {"EMA": EMAUpdater, "GS": GSUpdater, "RT": RTUpdater, "RM": RMUpdater, "None": BlankUpdater}
