"""Oracle-side gate matrices (TEST INFRASTRUCTURE - see oracle/__init__.py).

Written independently of the product's ``circuit.gate_matrix`` so that a wrong
convention in one of the two shows up as a parity failure.  Conventions are
Qiskit's (qiskit-terra 0.25.2.1 ``circuit/library/standard_gates``; third-party,
not vendored under /root/reference): ``U(theta,phi,lam)``, ``RZ = exp(-i Z
theta/2)``, little-endian two-qubit matrices with the first argument on bit 0.
"""
import numpy as np

_I2 = np.eye(2, dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)
_P0 = np.array([[1, 0], [0, 0]], dtype=complex)
_P1 = np.array([[0, 0], [0, 1]], dtype=complex)


def _rot(pauli, theta):
    """exp(-i theta/2 * pauli)"""
    return np.cos(theta / 2) * _I2 - 1j * np.sin(theta / 2) * pauli


def _phase(lam):
    return np.array([[1, 0], [0, np.exp(1j * lam)]], dtype=complex)


def _u3(theta, phi, lam):
    # U = P(phi) RY(theta) P(lam)
    return _phase(phi) @ _rot(_Y, theta) @ _phase(lam)


def _controlled(u):
    """control on bit 0 (first argument), target on bit 1: index = c + 2 t."""
    return np.kron(_I2, _P0) + np.kron(u, _P1)


def matrix(name, params=()):
    p = [float(x) for x in params]
    one = {
        "id": lambda: _I2, "i": lambda: _I2,
        "x": lambda: _X, "y": lambda: _Y, "z": lambda: _Z,
        "h": lambda: (_X + _Z) / np.sqrt(2),
        "s": lambda: _phase(np.pi / 2), "sdg": lambda: _phase(-np.pi / 2),
        "t": lambda: _phase(np.pi / 4), "tdg": lambda: _phase(-np.pi / 4),
        "sx": lambda: np.exp(1j * np.pi / 4) * _rot(_X, np.pi / 2),
        "rx": lambda: _rot(_X, p[0]), "ry": lambda: _rot(_Y, p[0]), "rz": lambda: _rot(_Z, p[0]),
        "p": lambda: _phase(p[0]), "u1": lambda: _phase(p[0]),
        "r": lambda: _rot(np.cos(p[1]) * _X + np.sin(p[1]) * _Y, p[0]),
        "u": lambda: _u3(*p), "u3": lambda: _u3(*p),
        "u2": lambda: _u3(np.pi / 2, p[0], p[1]),
    }
    if name in one:
        return np.asarray(one[name](), dtype=complex)
    two = {
        "cx": lambda: _controlled(_X), "cy": lambda: _controlled(_Y), "cz": lambda: _controlled(_Z),
        "cp": lambda: _controlled(_phase(p[0])),
        "rzz": lambda: np.diag(np.exp(-0.5j * p[0] * np.array([1, -1, -1, 1]))),
        "swap": lambda: np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex),
    }
    if name in two:
        return np.asarray(two[name](), dtype=complex)
    raise KeyError(f"oracle: unknown gate {name!r}")
