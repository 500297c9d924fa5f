"""Empty stand-in for PyPI ``pptree`` (imported by the reference's utils/visualization)."""


class Node:  # pragma: no cover
    def __init__(self, *a, **k):
        pass


def print_tree(*a, **k):  # pragma: no cover
    pass
