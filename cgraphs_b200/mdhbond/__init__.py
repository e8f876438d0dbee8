"""B200-native replacement for the H-bond path of ``cgraphs.mdhbond`` (reference
``cgraphs/mdhbond/__init__.py:1-3`` exports ``HbondAnalysis``, ``WireAnalysis`` and ``HydrationAnalysis``)."""
from .hbond import HbondAnalysis, dict2graph  # noqa: F401
from .batch import HbondBatch  # noqa: F401
from .waterwire import WireAnalysis  # noqa: F401
from .hydration import HydrationAnalysis  # noqa: F401
