"""Minimal stand-in for the PyPI ``parse`` package (pinned 1.20.2 by the reference's
poetry.lock), which is absent from this image.  TEST SCAFFOLDING ONLY: it lets the
*unmodified* reference python (numerical/format.py, sparse.py) be imported to generate
golden vectors.  Supports the subset those files use: ``{name}``, ``{name:d|w|l|f}``
and ``{{`` / ``}}`` escapes."""
import re

_TYPES = {
    "d": (r"[-+]?\d+", int),
    "w": (r"\w+", str),
    "l": (r"[A-Za-z]+", str),
    "f": (r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", float),
    "": (r".+?", str),
}


class Result(dict):
    @property
    def named(self):
        return dict(self)


def _compile(fmt):
    out, convs, i = [], {}, 0
    while i < len(fmt):
        c = fmt[i]
        if fmt.startswith("{{", i):
            out.append(re.escape("{")); i += 2
        elif fmt.startswith("}}", i):
            out.append(re.escape("}")); i += 2
        elif c == "{":
            j = fmt.index("}", i)
            name, _, typ = fmt[i + 1:j].partition(":")
            rx, conv = _TYPES[typ]
            out.append(f"(?P<{name}>{rx})")
            convs[name] = conv
            i = j + 1
        else:
            out.append(re.escape(c)); i += 1
    return re.compile("^" + "".join(out) + "$"), convs


def parse(fmt, string, *a, **k):
    rx, convs = _compile(fmt)
    m = rx.match(string)
    if m is None:
        return None
    return Result({n: convs[n](v) for n, v in m.groupdict().items()})
