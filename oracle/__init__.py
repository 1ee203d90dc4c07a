"""CPU oracle for the fragment-simulation + knit hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker (or the timed CPU
baseline) - never as the thing measured as "ours" or shipped.  The product
package (``hardwareawareoptimalquantumcircuitcuttingandknitting_b200``) never
imports this package and fails loudly when its CUDA library is missing.

What it restates (all paths relative to ``/root/reference``):

===========================  ====================================================
``gates.py``                 Qiskit standard-gate matrices (qiskit-terra 0.25.2.1,
                             third-party, not vendored) - independent of the
                             product's table in ``circuit.py``
``qpd_tables.py``            ``third_party/qvm/qvm/virtual_gates.py:62-103,154-177,
                             198-220,230-260,299-310`` (instantiation tables)
``statevector.py``           exact branching statevector semantics that replace
                             ``AerSimulator().run(..., shots)`` (``run.py:42``;
                             qiskit-aer 0.13.0, third-party, not vendored)
``instantiate.py``           ``virtual_circuit.py:21-48,97-131,183-213``
``sparse_knit.py``           ``quasi_distr.py:6-86``, ``virtual_gates.py:105-124,
                             179-194,262-286``, ``virtual_circuit.py:50-68,133-171,
                             193-194,216-228``
``dense.py``                 dense numpy forms of the same algebra (closed form of
                             SURVEY.md A.3), ``nearest_probability_distribution``
                             and qiskit's ``hellinger_fidelity``
``c/qck_oracle.c``           plain-C restatement of the two heavy loops (statevector
                             gate application, outer-product knit) used as the
                             timed CPU baseline
``ref_loader.py``            imports the *real* ``qvm/quasi_distr.py`` and
                             ``qvm/virtual_gates.py`` from ``/root/reference`` under
                             a qiskit stub (build container only) to generate
                             ``tests/golden``
===========================  ====================================================

Pinning status: the reference has **no tests and no golden
vectors** for this path.  The knit half (tables, QuasiDistr algebra, per-gate
knit, nearest_probability_distribution) is pinned against outputs of the
reference's own code executed in the build container (``tests/golden/*.json``,
generator ``tests/golden/make_golden.py``).  The simulate half replaces
qiskit-aer, which cannot be installed here: it is **parity unpinned** against Aer
itself and is anchored instead on (i) cut == uncut identities through the
reference's own knit code and (ii) analytically known distributions.
"""
