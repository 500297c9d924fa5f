"""Stand-in for PyPI ``bidict`` (reference uses only ``.inverse``). TEST SCAFFOLDING ONLY."""


class bidict(dict):
    @property
    def inverse(self):
        return {v: k for k, v in self.items()}
