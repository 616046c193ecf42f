"""Abstract agent interface (ssp_bayes_opt/agents/agent.py:15-42)."""


class Agent:
    def __init__(self):
        pass

    def eval(self, xs):
        """(mu, var, acquisition) of the surrogate at xs."""

    def update(self, x_t, y_t, sigma_t, step_num=0):
        """Fold one observation into the surrogate."""

    def untrusted(self, x, badness=None):
        """Penalise a region in the acquisition."""

    def acquisition_func(self):
        """(objective, jacobian) for scipy.optimize.minimize."""

    def length_scale(self):
        """Current length scale(s)."""
