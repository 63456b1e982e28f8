"""Learning-rate schedule: a host-side scalar, identical in behaviour to mlpcode/optim.py:12-72.

The "optimiser" proper is the in-place SGD fused into the layer kernels
(layers.py:117-118, 264, 268 -> bnn_bn_bwd_finalize and the BNN_EPI_SGD_F32 GEMM epilogue).
"""
from enum import Enum


class LRSchedulerStrat(Enum):
    na = "No Decay"
    time = "time-based"
    exp = "exponential"
    drop = "drop-based"


_DEFAULT_DECAY = {LRSchedulerStrat.na: 0.0, LRSchedulerStrat.time: 0.0, LRSchedulerStrat.exp: 1.0,
                  LRSchedulerStrat.drop: 1.0}


class LRScheduler:
    def __init__(self, alpha=0.1, decay_rate: float = None,
                 strategy: LRSchedulerStrat = LRSchedulerStrat.na, drop_every_n=10):
        if decay_rate is None:
            decay_rate = self.get_default_decay_rate(strategy)
        elif strategy == LRSchedulerStrat.exp and decay_rate > 1.0:
            raise ValueError("The decay rate must be less than one for exponential decay")
        elif strategy == LRSchedulerStrat.drop and decay_rate < 1:
            raise ValueError("The decay rate must be above 1 for drop based decay "
                             "(otherwise LR will increase)")
        assert decay_rate >= 0.0
        self.strategy = strategy
        self._rate = decay_rate
        self._alpha0 = alpha
        self._alpha = alpha
        self._steps = 0
        self._drop_every = drop_every_n
        self.step()  # the reference steps once in the constructor (optim.py:38)

    @property
    def value(self) -> float:
        return self._alpha

    @staticmethod
    def get_default_decay_rate(strategy: LRSchedulerStrat) -> float:
        if strategy not in _DEFAULT_DECAY:
            raise ValueError("Invalid value for LR Decay Strategy")
        return _DEFAULT_DECAY[strategy]

    def step(self) -> None:
        n = self._steps
        if self.strategy == LRSchedulerStrat.na:
            self._alpha = self._alpha0
        elif self.strategy == LRSchedulerStrat.time:
            self._alpha = 1.0 / (1 + self._rate * n) * self._alpha0
        elif self.strategy == LRSchedulerStrat.exp:
            self._alpha = self._alpha0 * (self._rate ** n)
        elif self.strategy == LRSchedulerStrat.drop:
            if (n + 1) % self._drop_every == 0:
                self._alpha = self._alpha0 / self._rate
        else:
            raise ValueError()
        self._steps = n + 1
