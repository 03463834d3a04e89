"""gpcounts_b200 - B200-native per-gene GP fitting (drop-in for GPcounts' ``Fit_GPcounts`` hot path)."""
__version__ = "0.1.0"
