from .corr import CorrBlock, AltCorrBlock  # noqa: F401
