"""gpcounts_b200 - B200-native per-gene GP fitting (drop-in for GPcounts' ``Fit_GPcounts`` hot path).

    from gpcounts_b200 import Fit_GPcounts          # same API as GPcounts.GPcounts_Module.Fit_GPcounts
"""
__version__ = "0.1.0"

from .GPcounts_Module import Fit_GPcounts  # noqa: E402,F401
