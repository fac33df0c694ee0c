"""pybindlibs.dictator (src/initializer/dictator.cpp:41-62): typed `add` functions filling the process-wide
PHAREDict.  Types are enforced like the pybind overloads do (a wrong type is an error, not a silent cast)."""
import numpy as np

from phare_b200.simulator import dict_instance


def _typed(name, check, conv):
    def add(path, value):
        if not check(value):
            raise TypeError(f"{name}(): incompatible function arguments: {type(value).__name__}")
        dict_instance().add(path, conv(value))
    add.__name__ = name
    return add


_is_int = lambda v: isinstance(v, (int, np.integer)) and not isinstance(v, bool)
add_size_t = _typed("add_size_t", lambda v: _is_int(v) and v >= 0, int)
add_optional_size_t = _typed("add_optional_size_t", lambda v: v is None or (_is_int(v) and v >= 0),
                             lambda v: None if v is None else int(v))
add_bool = _typed("add_bool", lambda v: isinstance(v, (bool, np.bool_)), bool)
add_int = _typed("add_int", _is_int, int)
add_vector_int = _typed("add_vector_int", lambda v: all(_is_int(x) for x in v), lambda v: [int(x) for x in v])
add_double = _typed("add_double", lambda v: isinstance(v, (float, int, np.floating, np.integer))
                    and not isinstance(v, bool), float)
add_string = _typed("add_string", lambda v: isinstance(v, (str, bytes)), lambda v: v.decode() if isinstance(v, bytes) else v)  # std::string takes bytes too and comes back as str
add_vector_string = _typed("add_vector_string", lambda v: all(isinstance(x, str) for x in v), list)
addInitFunction1D = _typed("addInitFunction1D", callable, lambda f: f)
addInitFunction2D = _typed("addInitFunction2D", callable, lambda f: f)
addInitFunction3D = _typed("addInitFunction3D", callable, lambda f: f)


def add_array_as_vector(path, array):
    a = np.asarray(array, dtype=np.float64)
    if a.ndim != 1:
        raise RuntimeError("Number of dimensions must be one")
    dict_instance().add(path, a.copy())


def stop():
    dict_instance().stop()
