class Space:  # pragma: no cover
    def __init__(self, *a, **k):
        pass


class Real(Space):  # pragma: no cover
    pass


class Integer(Space):  # pragma: no cover
    pass


class Categorical(Space):  # pragma: no cover
    pass
