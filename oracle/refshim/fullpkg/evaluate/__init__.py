"""Stand-in for PyPI ``evaluate`` (imported by the reference's modeling/hf.py). TEST SCAFFOLDING ONLY."""


def evaluator(*a, **k):  # pragma: no cover
    raise RuntimeError("evaluate is not installed in this image")


def load(*a, **k):  # pragma: no cover
    raise RuntimeError("evaluate is not installed in this image")
