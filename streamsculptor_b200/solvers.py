"""Stand-ins for the diffrax solver objects the reference passes around (main.py:126,288; fields.py:36).

Only Dopri5 and Dopri8 are on the hot path.  Any object whose class is called Dopri5/Dopri8 (e.g. a real
diffrax.Dopri8(scan_kind='bounded')) is accepted too.
"""


class _Solver:
    order = None

    def __init__(self, scan_kind=None, **_):
        self.scan_kind = scan_kind

    def __repr__(self):
        return f"{type(self).__name__}()"

    def __eq__(self, other):
        return type(self).__name__ == type(other).__name__

    def __hash__(self):
        return hash(type(self).__name__)


class Dopri5(_Solver):
    order = 5


class Dopri8(_Solver):
    order = 8


def solver_id(solver):
    name = solver if isinstance(solver, str) else type(solver).__name__
    if name == "Dopri5":
        return 5
    if name == "Dopri8":
        return 8
    raise NotImplementedError(f"solver {name!r} is not on the B200 hot path (only Dopri5 / Dopri8; no CPU fallback)")
