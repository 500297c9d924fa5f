"""Stand-in for PyPI ``graphviz`` (imported by the reference's utils/fx/visualize_graph.py). TEST SCAFFOLDING ONLY."""


class Digraph:  # pragma: no cover
    def __init__(self, *a, **k):
        pass

    def node(self, *a, **k):
        pass

    def edge(self, *a, **k):
        pass

    def render(self, *a, **k):
        pass
