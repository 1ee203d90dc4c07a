"""Oracle-side QPD instantiation tables (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates, per qubit, what ``VirtualGateEndpoint._circuit_on_index`` extracts from
each ``_instantiations()`` circuit of the reference
(``third_party/qvm/qvm/virtual_gates.py:62-103`` move, ``:154-177`` cz,
``:198-206`` cx, ``:210-220`` cy, ``:230-260`` rzz, ``:299-310`` cp,
``:141-150`` the per-qubit projection).  An instantiation is a pair
``(ops on qubit 0, ops on qubit 1)``; an op is ``("m",)`` for the measurement into
the config bit or ``(gate name, params tuple)``.

Pinned against a dump of the reference classes in
``tests/golden/instantiation_tables.json`` (``tests/test_oracle_golden.py``).
"""
from math import cos, pi, sin

RZZ_EPS = 1e-5          # virtual_gates.py:223
M = ("m",)


def _g(name, *params):
    return (name, tuple(float(p) for p in params))


def cz():
    return [
        ([_g("sdg")], [_g("sdg")]),
        ([_g("s")], [_g("s")]),
        ([M], []),
        ([M], [_g("z")]),
        ([], [M]),
        ([_g("z")], [M]),
    ]


def _sandwich(table, q0_pre=(), q0_post=(), q1_pre=(), q1_post=()):
    return [(list(q0_pre) + a + list(q0_post), list(q1_pre) + b + list(q1_post)) for a, b in table]


def cx():
    return _sandwich(cz(), q1_pre=[_g("h")], q1_post=[_g("h")])


def cy():
    return _sandwich(cx(), q1_pre=[_g("rz", -pi / 2)], q1_post=[_g("rz", pi / 2)])


def move():
    return [
        ([], []),
        ([], [_g("x")]),
        ([_g("h"), M], [_g("h")]),
        ([_g("h"), M], [_g("x"), _g("h")]),
        ([_g("sdg"), _g("h"), M], [_g("h"), _g("s")]),
        ([_g("sdg"), _g("h"), M], [_g("x"), _g("h"), _g("s")]),
        ([M], []),
        ([M], [_g("x")]),
    ]


def rzz(theta):
    m = -theta
    i0 = ([], [])
    i1 = ([_g("z")], [_g("z")])
    if abs(cos(m / 2)) < RZZ_EPS:
        return [i1]
    if abs(sin(m / 2)) < RZZ_EPS:
        return [i0]
    return [i0, i1,
            ([_g("rz", -pi / 2)], [M]),
            ([M], [_g("rz", -pi / 2)]),
            ([_g("rz", pi / 2)], [M]),
            ([M], [_g("rz", pi / 2)])]


def cp(theta):
    """Reference behaviour *as written*: after construction params[0] = -theta/2;
    tables = rz(params[0]/2) on q0, RZZ table at params[0], rz(params[0]/2) on q1."""
    lam = -theta / 2
    return _sandwich(rzz(lam), q0_pre=[_g("rz", lam / 2)], q1_post=[_g("rz", lam / 2)])


def table(kind, theta=None):
    if kind == "move":
        return move()
    if kind == "cz":
        return cz()
    if kind == "cx":
        return cx()
    if kind == "cy":
        return cy()
    if kind == "rzz":
        return rzz(theta)
    if kind == "cp":
        return cp(theta)
    raise KeyError(kind)


def knit_param(kind, theta):
    """The angle ``m_theta`` used by the reference's RZZ-family knit
    (``virtual_gates.py:263``: ``-self._params[0]``; for cp params[0] was already
    overwritten with ``-theta/2``, ``:297``)."""
    if kind == "rzz":
        return -theta
    if kind == "cp":
        return theta / 2
    return None
