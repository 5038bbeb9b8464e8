"""`pyro.contrib.easyguide.EasyGuide` base (S/guides.py:18): a PyroModule holding the model whose call runs
`self.guide(*args)`.  The grouping helpers of the real class are not used by DRAUPNIRGUIDES and are not provided."""
from __future__ import annotations

from ..nn import PyroModule


class EasyGuide(PyroModule):
    def __init__(self, model):
        super().__init__()
        self.__dict__["model"] = model      # not registered as a sub-module (the real class does the same)
        self.prototype_trace = None

    def guide(self, *args, **kwargs):
        raise NotImplementedError

    def forward(self, *args, **kwargs):
        return self.guide(*args, **kwargs)
