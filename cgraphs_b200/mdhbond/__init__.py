"""B200-native replacement for the H-bond path of ``cgraphs.mdhbond`` (reference
``cgraphs/mdhbond/__init__.py:1`` exports ``HbondAnalysis``)."""
from .hbond import HbondAnalysis, dict2graph  # noqa: F401
from .batch import HbondBatch  # noqa: F401
