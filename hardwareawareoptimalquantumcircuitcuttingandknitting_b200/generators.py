"""Qiskit-free restatement of the benchmark circuit generators (inputs only).

The reference builds its benchmark inputs with Qiskit
(``benchmarks/helper_functions.py:66-127,206-233`` dispatching into
``benchmarks/qcg/**``).  The generators themselves are *not* on the hot path, but
without them there is nothing to feed it, so the six circuits BASELINE.json's
north_star names are restated here as gate lists over ``circuit.QuantumCircuit``:

=======  =====================================================================
``bv``   ``qcg/BernsteinVazirani/bernstein_vazirani.py:61-98`` with the all-ones
         secret of ``helper_functions.py:26-31``
``hwe``  ``qcg/QAOA/hw_efficient_ansatz.py:79-104,118-187`` (``parameters="optimal"``)
``syc``  ``qcg/Supremacy/Qgrid_Sycamore.py:84-176``, ``Qbit_Sycamore.py:10-19``,
         ``ABCD_layer_generation.py:5-70``; grid from ``helper_functions.py:16-24``
``qft``  ``helper_functions.py:83-86`` -> qiskit ``QFT(n, approximation_degree=0,
         do_swaps=False)`` (library source not vendored; restated from the
         published construction: for j = n-1..0: H(j), then CP(pi/2^(j-k)) (j,k)
         for k = j-1..0, truncated to the ``approximation_degree`` nearest ones)
``aqft`` ``helper_functions.py:87-93`` (approximation_degree = n - int(log2 n + 2))
``add``  ``qcg/Arithmetic/ripple_carry_adder.py:117-199`` (a = b = 0, decomposed
         Toffoli)
=======  =====================================================================

Every generator ends with ``measure_all()`` as ``generateBv``/... do
(``helper_functions.py:163-197``).  ``syc`` draws its one-qubit gates from
Python's ``random``; the reference seeds it with ``None``
(``helper_functions.py:66-67``), i.e. is not reproducible - here the caller
passes ``seed`` and it is recorded with every result.
"""
from __future__ import annotations

import math
import random

from .circuit import QuantumCircuit, QuantumRegister

__all__ = ["gen_circ", "factor_int", "gen_bv", "gen_hwea", "gen_sycamore", "gen_qft", "gen_adder",
           "sycamore_layers"]


def factor_int(n: int) -> tuple[int, int]:
    val = math.ceil(math.sqrt(n))
    while True:
        co = int(n / val)
        if val * co == n:
            return val, co
        val -= 1


def _new(n: int) -> QuantumCircuit:
    return QuantumCircuit(QuantumRegister(n, "q"))


def gen_bv(num_qubits: int) -> QuantumCircuit:
    nq = num_qubits - 1                       # secret of nq ones + one ancilla
    qc = _new(num_qubits)
    qc.x(nq)
    for i in range(num_qubits):
        qc.h(i)
    for i in range(nq):
        qc.cx(i, nq)
    for i in range(num_qubits):
        qc.h(i)
    return qc


def gen_hwea(num_qubits: int, depth: int) -> QuantumCircuit:
    n = num_qubits
    theta = [0.0] * (2 * n * (1 + depth))
    theta[0] = math.pi / 2
    for i in range(2 * n, 2 * n + n // 2):
        theta[i] = math.pi
    qc = _new(n)
    p = 0
    for i in range(n):
        qc.u(theta[i + p], 0, 0, i)
    p += n
    for i in range(n):
        qc.u(0, 0, theta[i + p], i)
    p += n
    for _ in range(depth):
        for i in range(n - 1):
            qc.cx(i, i + 1)
        for i in range(n):
            qc.u(theta[i + p], 0, 0, i)
        p += n
        for i in range(n):
            qc.u(0, 0, theta[i + p], i)
        p += n
    return qc


def sycamore_layers(rows: int, cols: int) -> list[list[tuple[int, int]]]:
    """A, B (horizontal) and C, D (vertical) coupler patterns as qubit-index pairs."""
    def horiz(first_even: int):
        out = []
        for r in range(rows):
            start = first_even if r % 2 == 0 else 1 - first_even
            for c in range(start, cols, 2):
                if c != cols - 1:
                    out.append((r * cols + c, r * cols + c + 1))
        return out

    def vert(first_even: int):
        out = []
        for c in range(cols):
            start = first_even if c % 2 == 0 else 1 - first_even
            for r in range(start, rows, 2):
                if r != rows - 1:
                    out.append((r * cols + c, (r + 1) * cols + c))
        return out

    return [horiz(0), horiz(1), vert(0), vert(1)]


def gen_sycamore(num_qubits: int, depth: int, seed: int | None = 0) -> QuantumCircuit:
    rows, cols = factor_int(num_qubits)
    rng = random.Random(seed)                 # same Mersenne-Twister stream as random.seed(seed)
    layers = sycamore_layers(rows, cols)
    order = [0, 1, 2, 3, 2, 3, 0, 1]
    choices = {"X": ("Y", "W"), "Y": ("X", "W"), "W": ("X", "Y")}
    prev: list[str | None] = [None] * num_qubits
    qc = _new(num_qubits)
    for d in range(depth):
        for q in range(num_qubits):
            if prev[q] is None:
                g = ("X", "Y", "W")[rng.randint(0, 2)]
            else:
                g = choices[prev[q]][rng.randint(0, 1)]
            prev[q] = g
            if g == "X":
                qc.rx(math.pi / 2, q)
            elif g == "Y":
                qc.ry(math.pi / 2, q)
            else:
                qc.z(q)                       # sic: the reference applies Z for "W"
        for a, b in layers[order[d % len(order)]]:
            qc.cz(a, b)
    return qc


def gen_qft(num_qubits: int, approximation_degree: int = 0) -> QuantumCircuit:
    n = num_qubits
    qc = _new(n)
    for j in reversed(range(n)):
        qc.h(j)
        num_ent = max(0, j - max(0, approximation_degree - (n - j - 1)))
        for k in reversed(range(j - num_ent, j)):
            qc.cp(math.pi * (2.0 ** (k - j)), j, k)
    return qc


def gen_adder(num_qubits: int) -> QuantumCircuit:
    nbits = int((num_qubits - 2) / 2)
    nq = 2 * nbits + 2
    qc = _new(nq)

    def toffoli(x, y, z):
        qc.h(z); qc.cx(y, z); qc.tdg(z); qc.cx(x, z); qc.t(z); qc.cx(y, z); qc.t(y); qc.tdg(z)
        qc.cx(x, z); qc.cx(x, y); qc.t(z); qc.h(z); qc.t(x); qc.tdg(y); qc.cx(x, y)

    a_idx = [2 * i + 2 for i in range(nbits)]
    for a in a_idx:                            # MAJ ladder
        x, y, z = a - 2, a - 1, a
        qc.cx(z, y); qc.cx(z, x); toffoli(x, y, z)
    qc.cx(a_idx[-1], nq - 1)
    for a in reversed(a_idx):                  # UMA ladder
        x, y, z = a - 2, a - 1, a
        qc.x(y); qc.cx(x, y); toffoli(x, y, z); qc.x(y); qc.cx(z, x); qc.cx(z, y)
    return qc


def gen_circ(name: str, num_qubits: int, depth: int = 1, seed: int | None = 0) -> QuantumCircuit:
    """``genCirc`` (``helper_functions.py:206-233``) for the six north_star circuits."""
    name = name.lower()
    if name == "bv":
        qc = gen_bv(num_qubits)
    elif name == "hwe":
        qc = gen_hwea(num_qubits, depth)
    elif name == "syc":
        qc = gen_sycamore(num_qubits, depth, seed)
    elif name == "qft":
        qc = gen_qft(num_qubits, 0)
    elif name == "aqft":
        qc = gen_qft(num_qubits, num_qubits - int(math.log(num_qubits, 2) + 2))
    elif name == "add":
        qc = gen_adder(num_qubits)
    else:
        raise RuntimeError(f"circName {name} is not supported")
    assert qc.num_qubits == num_qubits
    qc.name = f"{name}_{num_qubits}_{depth}"
    qc.measure_all()
    return qc
