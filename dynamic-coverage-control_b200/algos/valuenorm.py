"""ValueNorm — the reference's value normaliser (utils/valuenorm.py:8-79) as three float32 on the device.

The state {running_mean, running_mean_sq, debiasing_term} lives in one CUDA tensor that the kernels read and update
in place (GAE denormalises with it, `dcc_mappo_epoch_grads` applies `update` once per epoch before normalising,
algos/mappo.py:107-109).  The methods below mirror the reference's API for callers and tests; they are not on the
hot path.
"""
import torch


class ValueNorm:
    def __init__(self, input_shape=1, beta=0.99999, epsilon=1e-5, device="cuda"):
        assert int(input_shape) == 1
        self.beta, self.epsilon = beta, epsilon
        self.state = torch.zeros(4, dtype=torch.float32, device=device)   # [mean, mean_sq, debias, pad]

    @property
    def running_mean(self):
        return self.state[0:1]

    @property
    def running_mean_sq(self):
        return self.state[1:2]

    @property
    def debiasing_term(self):
        return self.state[2]

    def running_mean_var(self):
        c = self.state[2].clamp(min=self.epsilon)
        mean = self.state[0:1] / c
        var = (self.state[1:2] / c - mean ** 2).clamp(min=1e-2)
        return mean, var

    @torch.no_grad()
    def update(self, input_vector):
        x = torch.as_tensor(input_vector, dtype=torch.float32, device=self.state.device)
        w = self.beta
        self.state[0].mul_(w).add_(x.mean() * (1.0 - w))
        self.state[1].mul_(w).add_((x ** 2).mean() * (1.0 - w))
        self.state[2].mul_(w).add_(1.0 - w)

    def normalize(self, input_vector):
        x = torch.as_tensor(input_vector, dtype=torch.float32, device=self.state.device)
        mean, var = self.running_mean_var()
        return (x - mean) / torch.sqrt(var)

    def denormalize(self, input_vector):
        x = torch.as_tensor(input_vector, dtype=torch.float32, device=self.state.device)
        mean, var = self.running_mean_var()
        return x * torch.sqrt(var) + mean

    def state_dict(self):
        return {"running_mean": self.state[0:1].clone(), "running_mean_sq": self.state[1:2].clone(),
                "debiasing_term": self.state[2].clone()}

    def load_state_dict(self, sd):
        self.state[0] = float(sd["running_mean"])
        self.state[1] = float(sd["running_mean_sq"])
        self.state[2] = float(sd["debiasing_term"])
