"""TEST INFRASTRUCTURE — numpy restatement of the reference trainer's normalisers.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import this module.
Restates the arithmetic of `gym_pybullet_drones/safe_control_gym/math_and_models/normalization.py`:
running moments (:13-58), the mean / std observation normaliser (:64-96) and the return-std reward
normaliser (:99-141), with the operation order the reference uses (the golden outputs of the unmodified
reference classes are reproduced bit for bit: `tests/golden/make_golden_trainer.py` ->
`tests/golden/normalizers.npz`, `tests/test_trainer_golden.py`).
"""
import numpy as np


def fold_batch(mean, var, count, b_mean, b_var, b_count):
    """Running (mean, var, count) after absorbing a batch with moments (b_mean, b_var, b_count): the parallel-variance
    update of normalization.py:44-58, evaluated in the reference's order of operations."""
    n = count + b_count
    shift = b_mean - mean
    mean_out = mean + shift * b_count / n
    spread = var * count + b_var * b_count + np.square(shift) * count * b_count / (count + b_count)
    return mean_out, spread / (count + b_count), b_count + count


class RunningMeanStdOracle:
    """State of normalization.py:24-32 (mean 0, var 1, count = epsilon) + `update` (:34-42)."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean, self.var, self.count = np.zeros(shape, np.float64), np.ones(shape, np.float64), epsilon

    def update(self, arr):
        # moments over the leading (env) axis, population variance, in the array's own dtype like np.mean / np.var
        self.update_from_moments(np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        self.mean, self.var, self.count = fold_batch(self.mean, self.var, self.count, batch_mean, batch_var, batch_count)


class MeanStdNormalizerOracle:
    """(x - running mean) / sqrt(running var + eps), clipped (normalization.py:67-88); `read_only` freezes the statistics."""

    def __init__(self, shape=(), read_only=False, clip=10.0, epsilon=1e-8):
        self.rms = RunningMeanStdOracle(shape=shape)
        self.read_only, self.clip, self.epsilon = read_only, clip, epsilon

    def __call__(self, x):
        x = np.asarray(x)
        if not self.read_only:
            self.rms.update(x)
        scaled = (x - self.rms.mean) / np.sqrt(self.rms.var + self.epsilon)
        return np.clip(scaled, -self.clip, self.clip)


class RewardStdNormalizerOracle(MeanStdNormalizerOracle):
    """Rewards divided by the running std of the discounted return (normalization.py:111-141): the return accumulator
    is updated first, its batch moments folded in, finished envs restart from zero, and only then is x scaled."""

    def __init__(self, gamma=0.99, read_only=False, clip=10.0, epsilon=1e-8):
        super().__init__((), read_only, clip, epsilon)
        self.gamma, self.ret = gamma, None

    def __call__(self, x, dones):
        x = np.asarray(x)
        if not self.read_only:
            self.ret = (np.zeros_like(x) if self.ret is None else self.ret) * self.gamma + x
            self.rms.update(self.ret)
            self.ret[dones.astype(bool)] = 0
        return np.clip(x / np.sqrt(self.rms.var + self.epsilon), -self.clip, self.clip)
