"""TEST INFRASTRUCTURE — numpy restatement of the reference trainer's normalisers.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import this module.
Restates `gym_pybullet_drones/safe_control_gym/math_and_models/normalization.py`:
`RunningMeanStd` (:13-58), `MeanStdNormalizer` (:64-96), `RewardStdNormalizer` (:99-141).
Pinned bit-for-bit against the unmodified reference classes on seeded batches
(`tests/golden/make_golden_trainer.py` -> `tests/golden/normalizers.npz`).
"""
import numpy as np


class RunningMeanStdOracle:
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)                               # :30-32
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr):                                                    # :34-42
        batch_mean = np.mean(arr, axis=0)
        batch_var = np.var(arr, axis=0)
        self.update_from_moments(batch_mean, batch_var, arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):        # :44-58
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + np.square(delta) * self.count * batch_count / (self.count + batch_count)
        self.mean = new_mean
        self.var = m_2 / (self.count + batch_count)
        self.count = batch_count + self.count


class MeanStdNormalizerOracle:
    def __init__(self, shape=(), read_only=False, clip=10.0, epsilon=1e-8):   # :67-80
        self.read_only = read_only
        self.rms = RunningMeanStdOracle(shape=shape)
        self.clip = clip
        self.epsilon = epsilon

    def __call__(self, x):                                                    # :82-88
        x = np.asarray(x)
        if not self.read_only:
            self.rms.update(x)
        return np.clip((x - self.rms.mean) / np.sqrt(self.rms.var + self.epsilon), -self.clip, self.clip)


class RewardStdNormalizerOracle(MeanStdNormalizerOracle):
    def __init__(self, gamma=0.99, read_only=False, clip=10.0, epsilon=1e-8):  # :111-122
        super().__init__((), read_only, clip, epsilon)
        self.gamma = gamma
        self.ret = None

    def __call__(self, x, dones):                                             # :124-141
        x = np.asarray(x)
        if not self.read_only:
            if self.ret is None:
                self.ret = np.zeros_like(x)
            self.ret = self.ret * self.gamma + x
            self.rms.update(self.ret)
            self.ret[dones.astype(bool)] = 0
        return np.clip(x / np.sqrt(self.rms.var + self.epsilon), -self.clip, self.clip)
