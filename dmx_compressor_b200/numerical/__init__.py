from .format import (  # noqa: F401
    Format,
    Same,
    FixedPoint,
    FloatingPoint,
    BlockFloatingPoint,
    ScaledBlockFloatingPoint,
    MXINT,
    MXFP,
    ROUNDING_MODE,
)
from .observer import DMXObserverBase, DummyObserver, HistogramObserver, MinMaxObserver  # noqa: F401
from .cast import CastTo, CastToDict, CastToFormat  # noqa: F401
