"""`DetachableModule`, `BcosSequential` -- mirror of reference bcos/modules/common.py:8-51."""
from torch import nn

from ..explain import BcosUtilMixin

__all__ = ["DetachableModule", "BcosSequential"]


class DetachableModule(nn.Module):
    """Base of modules whose dynamic weights can be detached from the graph (explanation mode); common.py:8-34."""

    def __init__(self):
        super().__init__()
        self.detach = False

    def set_explanation_mode(self, activate: bool = True) -> None:
        self.detach = activate

    @property
    def is_in_explanation_mode(self) -> bool:
        return self.detach


class BcosSequential(BcosUtilMixin, nn.Sequential):
    """nn.Sequential with the explanation helpers; common.py:37-51."""

    def __init__(self, *args):
        super().__init__(*args)

    @classmethod
    def from_standard_module(cls, mod):
        return cls(*mod._modules.values())
