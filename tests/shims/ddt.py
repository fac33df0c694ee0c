"""A small stand-in for the `ddt` package (absent from this image), enough for the reference's own unittest files
(tests/simulator/**): @ddt on the class expands methods decorated with @data(...) (optionally @unpack) into one test method
per datum."""
import functools

_DATA, _UNPACK = "%ddt_values", "%ddt_unpack"


def data(*values):
    def wrap(fn):
        setattr(fn, _DATA, values)
        return fn
    return wrap


def unpack(fn):
    setattr(fn, _UNPACK, True)
    return fn


def _name(base, i, v):
    try:
        s = str(v)
    except Exception:
        s = ""
    s = "".join(c if c.isalnum() else "_" for c in s)[:60]
    return f"{base}_{i + 1}_{s}"


def ddt(cls):
    for name, fn in list(cls.__dict__.items()):
        if not hasattr(fn, _DATA):
            continue
        for i, v in enumerate(getattr(fn, _DATA)):
            def make(fn=fn, v=v):
                @functools.wraps(fn)
                def test(self, *a, **k):
                    if getattr(fn, _UNPACK, False):
                        if isinstance(v, dict):
                            return fn(self, *a, **{**v, **k})
                        if isinstance(v, (tuple, list)):
                            return fn(self, *v, *a, **k)
                    return fn(self, v, *a, **k)
                return test
            t = make()
            t.ddt_value = v
            t.__name__ = _name(name, i, v)
            setattr(cls, t.__name__, t)
        delattr(cls, name)
    return cls
