"""Orchestration of the fidelity harness mirror (utilities.py; Utilities.py:36-226) with the device calls replaced
by fakes - CPU only.  The device calls it composes (B200Backend.run, run_virtual_circuit, hellinger_fidelity) have
their own GPU tests."""
import threading
from importlib import import_module

import pytest

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
util = import_module(f"{PKG}.utilities")
circuit = import_module(f"{PKG}.circuit")


class _FakeBackend:
    def __init__(self, tag="ideal"):
        self.tag = tag
        self.threads = []

    def run(self, circ, shots=0):
        self.threads.append(threading.get_ident())
        counts = {"01": 0.25 * shots, "10": 0.75 * shots} if self.tag == "ideal" else {"01": shots}
        return type("J", (), {"result": lambda s: type("R", (), {"get_counts": lambda r: counts})()})()


@pytest.fixture
def patched(monkeypatch):
    calls = {"virt": []}
    monkeypatch.setattr(util, "B200Backend", lambda: _FakeBackend("ideal"))
    monkeypatch.setattr(util.QuasiDistr, "from_counts",
                        staticmethod(lambda counts, **kw: {int(k, 2): v / sum(counts.values())
                                                           for k, v in counts.items()}))

    class _Virt:
        def __init__(self, circ):
            self.backend = None

        def set_backend_for_all(self, be):
            self.backend = be

    monkeypatch.setattr(util, "VirtualCircuit", _Virt)

    def fake_run(virt, shots=0):
        calls["virt"].append((virt.backend.tag, threading.get_ident()))
        return ({1: 0.25, 2: 0.75} if virt.backend.tag == "ideal" else {1: 1.0}), None

    monkeypatch.setattr(util, "run_virtual_circuit", fake_run)
    monkeypatch.setattr(util, "hellinger_fidelity", lambda p, q, num_bits=None: (p, q, num_bits))
    return calls


def _circ():
    qc = circuit.QuantumCircuit(circuit.QuantumRegister(2, "q"))
    qc.h(0)
    qc.measure_all()
    return qc


def test_compare_runs_four_workers_and_pairs_results(patched):
    noisy = _FakeBackend("noisy")
    f_in, f_cut, f_cross = util.compareOriginalCircWithCutCirc(_circ(), _circ(), noisy, 1000)
    ideal, wrong = {1: 0.25, 2: 0.75}, {1: 1.0}
    assert f_in == (ideal, wrong, 2)          # uncut: ideal vs backend
    assert f_cut == (ideal, wrong, 2)         # cut: ideal vs backend
    assert f_cross == (ideal, ideal, 2)       # uncut ideal vs cut ideal: the number benchmark.py:99-102 logs
    assert sorted(t for t, _ in patched["virt"]) == ["ideal", "noisy"]
    main = threading.get_ident()
    assert all(tid != main for _, tid in patched["virt"]) and all(tid != main for tid in noisy.threads)


def test_backend_none_uses_a_second_exact_backend(patched):
    f_in, f_cut, f_cross = util.compareOriginalCircWithCutCirc(_circ(), _circ())
    assert f_in[0] == f_in[1] and f_cut[0] == f_cut[1]


def test_worker_exception_reaches_the_caller(patched, monkeypatch):
    def boom(virt, shots=0):
        raise ValueError("Fragment not found.")
    monkeypatch.setattr(util, "run_virtual_circuit", boom)
    with pytest.raises(ValueError):
        util.getVirtualCircResultFromBackend(_circ(), None, 10)
