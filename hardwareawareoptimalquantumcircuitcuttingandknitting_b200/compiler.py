"""Fragment circuit -> device programs (the host half of fragment execution).

The reference materialises every instance as a Qiskit circuit
(``third_party/qvm/qvm/virtual_circuit.py:183-213``: copy the fragment, add the
``vgate_c`` register, splice in ``VirtualGateEndpoint.instantiate(label[k])``,
``decompose()``) and hands the list to Aer (``run.py:42``).  Here a fragment is
compiled ONCE into a template whose virtual-gate endpoints are *slots*; an
instance is just a mixed-radix label decoded on the device, where each digit
selects the slot's variant matrices from a table.  Instances that share the same
set of measuring slots (a *pattern*) share one program (``qck_sim_plan``):

* a slot variant is ``pre-matrix, measure?, post-matrix`` (the one-qubit share
  of ``_instantiations()[id]``, ``virtual_gates.py:134-150``; at most one
  measurement, ``:45-50``);
* a measurement whose qubit is used afterwards (gate cuts: the wire continues to
  its final measurement) becomes ``CX(qubit -> fresh ancilla bit)``: the two
  halves of the enlarged state are the two un-normalised branches of SURVEY.md
  A.2; a measurement on a wire that ends there (the old end of a wire cut,
  ``Cutter.py:624-643``) is read off ``|amp|^2`` directly;
* the output row of an instance is indexed by the fragment's finally-measured
  clbits in ascending clbit order (so a fragment's row index is
  ``pext(key, out_mask)``); config bits are folded with sign ``(-1)^bit`` (exact
  mode) or kept as extra row bits (``fold=False``, for the
  reference-faithful pruning mode);
* consecutive one-qubit gates on a wire are multiplied together on the host
  (complex128), also into the following slot's pre-matrices;
* programs whose state (qubits + ancillas) exceeds the on-chip limit are cut into
  *sweeps*: runs of ops that only touch <= ``STREAM_TILE`` qubits, always
  including the ``LOW_RUN`` lowest ones so that every tile is a set of
  contiguous >= 512-byte runs in HBM.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .circuit import Barrier, Gate, Measure, QuantumCircuit, QuantumRegister
from .virtual_gates import VirtualBinaryGate, VirtualGateEndpoint

ONCHIP_MAX_QUBITS = 13      # 2^13 complex128 = 128 KiB of the 227 KiB shared memory
import os as _os

# Register-resident regime (csrc/sim_warp_kernel.inc): fragments of <= WARP_MAX_QUBITS qubits whose programs hold
# only one-qubit gates, cx and cz run ONE WARP PER INSTANCE with the state in registers (lane bits = qubits 0-4:
# shuffle butterflies) and a depth-first walk over the outcomes of mid-circuit measurements instead of ancilla
# bits.  Pair fusion is off for them (a fused 4x4 would need all four amplitudes of a quad in one lane).
# QCK_SIM_WARP=0 keeps them on the shared-memory kernel.
WARP_MAX_QUBITS = 10
WARP_MAX_DEPTH = 8          # QCK_WARP_MAX_DEPTH: measurements whose qubit lives on, per program
WARP = _os.environ.get("QCK_SIM_WARP", "1") != "0"
# Tree-walk simulation (csrc/sim_tree_kernel.inc): register-regime fragments whose every virtual gate
# has ONE endpoint in the fragment run all their instances level by level, every shared prefix once, from ONE
# program (no per-pattern programs).  QCK_SIM_TREE=0 keeps them on the per-instance warp kernel.
TREE = _os.environ.get("QCK_SIM_TREE", "1") != "0"
TREE_MAX_WORK_BYTES = 2 << 30

STREAM_TILE = int(_os.environ.get("QCK_STREAM_TILE", "12"))   # 64 KiB tiles; env override = tuning knob
LOW_RUN = 5                 # tiles always hold qubits 0..4: 32 amplitudes = 512 contiguous bytes
# Prefix sharing (first level): the ops of an on-chip program that precede its first label-dependent
# op are the same for every instance; they run once per program and the instances start from that state.
# "auto": only where instances, not latency, are the bound - many instances and a prefix that is most of the
# program (the sending side of wire cuts: basis change + measurement at the END of the fragment);
# QCK_SHARE_PREFIX=1 / 0 forces it on (whenever a prefix exists) / off.
# Instance de-duplication: label digits that select identical variants in every slot of their gate inside this
# fragment give identical instances; one representative per class is simulated and its row copied to the others
# (qck_rows_broadcast).  The copy is one more dependent launch per fragment: measured on one box, syc-16 d5 (671 of
# 1 296 instances saved per fragment) 0.265 -> 0.284 ms and bv-16 0.143 -> 0.165 ms, but hwe-16 d5 (4 651 of 7 776
# saved) 0.75 -> 0.41 ms and aqft-16 with five wire cuts (31 744 and 24 992 of 32 768 saved) 7.7 -> 3.0 ms.  Hence
# "auto": only when at least DEDUPE_MIN_SAVED instances are saved; QCK_DEDUPE=1 / 0 forces it on / off.
DEDUPE = {"1": True, "0": False}.get(_os.environ.get("QCK_DEDUPE", ""), "auto")
DEDUPE_MIN_SAVED = 2048
SHARE_PREFIX = {"1": True, "0": False}.get(_os.environ.get("QCK_SHARE_PREFIX", ""), "auto")
SHARE_PREFIX_MIN_INSTANCES = 1024       # instances of the fragment
SHARE_PREFIX_MIN_PLAN = 64              # instances of the program (measurement pattern)
SHARE_PREFIX_MIN_FRACTION = 0.5

_I2 = np.eye(2, dtype=np.complex128)


def _is_identity(m: np.ndarray) -> bool:
    # exact comparison with the 2x2 identity (np.array_equal costs ~10 us per call; this runs per gate)
    return m.shape == (2, 2) and m[0, 0] == 1 and m[1, 1] == 1 and m[0, 1] == 0 and m[1, 0] == 0


# Host compiler in C++ (csrc/host_program.cu): the fragment circuit is flattened into integer records + a matrix
# pool (which double as the structure key of the process-wide program cache) and lowered / fused natively.
# QCK_HOST_COMPILE=py keeps the Python originals below (the reference the native compiler is tested against).
NATIVE = _os.environ.get("QCK_HOST_COMPILE", "native") != "py"

IN_G1, IN_G2, IN_CX, IN_CZ, IN_MEASURE, IN_ENDPOINT = 1, 2, 3, 4, 5, 6
T_U1, T_CX, T_CZ, T_U2, T_SLOT, T_MMEAS = 0, 1, 2, 3, 4, 5
_TOP_NAMES = ("u1", "cx", "cz", "u2", "slot", "mmeas")

_gate_matrix_cache: dict = {}
_endpoint_table_cache: dict = {}


def _endpoint_table(vg: VirtualBinaryGate, side: int):
    """-> (pre [n, 2, 2], meas flags, post [n, 2, 2]) of one end of a virtual gate: the one-qubit share of every
    instantiation (``virtual_gates.py:134-150``) as matrix before / measure? / matrix after."""
    key = (type(vg), tuple(vg.params), side)
    hit = _endpoint_table_cache.get(key)
    if hit is not None:
        return hit
    pre, meas, post = [], [], []
    for inst in vg._table():
        a, b, seen = _I2, _I2, False
        for name, params, q in inst:
            if q != side:
                continue
            if name == "measure":
                if seen:
                    raise ValueError("instantiation measures a qubit twice")
                seen = True
                continue
            m = Gate(name, 1, params).to_matrix()
            if seen:
                b = m @ b
            else:
                a = m @ a
        pre.append(a)
        meas.append(seen)
        post.append(b)
    if len(pre) > _lib.MAX_VARIANTS:
        raise NotImplementedError("a virtual gate has more instantiations than QCK_MAX_VARIANTS")
    hit = (np.stack(pre).astype(np.complex128), tuple(meas), np.stack(post).astype(np.complex128))
    if len(_endpoint_table_cache) < 4096:
        _endpoint_table_cache[key] = hit
    return hit


class FlatCircuit:
    """A fragment circuit as the C ABI takes it (``qck_host_lower``): instruction records, endpoint records, matrix
    pool.  ``key`` identifies everything a program depends on."""
    __slots__ = ("n_qubits", "n_clbits", "instr", "endpoints", "pool", "tables", "key")


def flatten(circ: QuantumCircuit, fragment: QuantumRegister) -> FlatCircuit:
    n_q = len(fragment)
    creg_off, off = {}, 0
    for r in circ.cregs:
        creg_off[id(r)] = off
        off += len(r)
    rows, eps, tables = [], [], []
    pool, pool_len, mat_off = [], 0, {}
    add = rows.append
    try:
        for ins in circ.data:
            op = ins.operation
            kind = type(op)
            if kind is Gate or isinstance(op, Gate):
                qs = ins.qubits
                nq = len(qs)
                q0 = qs[0]
                if q0.register is not fragment or (nq == 2 and qs[1].register is not fragment):
                    raise KeyError(q0)
                m = op._matrix
                name = op.name
                if m is None:
                    if nq == 2:
                        if name == "cx":
                            add((IN_CX, q0.index, qs[1].index, -1, -1, -1))
                            continue
                        if name == "cz":
                            add((IN_CZ, q0.index, qs[1].index, -1, -1, -1))
                            continue
                    params = op.params
                    mkey = (name, params[0]) if len(params) == 1 else (name, *params)
                else:
                    mkey = m.tobytes()
                if nq > 2 or nq != op.num_qubits:
                    raise NotImplementedError(f"{op.name}: gates on more than two qubits are not supported")
                o = mat_off.get(mkey)
                if o is None:
                    if m is None:
                        m = _gate_matrix_cache.get(mkey)
                        if m is None:
                            m = np.ascontiguousarray(op.to_matrix(), dtype=np.complex128)
                            if len(_gate_matrix_cache) < 65536:
                                _gate_matrix_cache[mkey] = m
                    if m.size != (4 if nq == 1 else 16):
                        raise ValueError(f"{op.name}: matrix of the wrong size")
                    o = mat_off[mkey] = pool_len
                    pool.append(np.ascontiguousarray(m, dtype=np.complex128).reshape(-1).view(np.float64))
                    pool_len += 2 * m.size
                if nq == 1:
                    add((IN_G1, qs[0].index, 0, -1, o, -1))
                else:
                    add((IN_G2, qs[0].index, qs[1].index, -1, o, -1))
            elif isinstance(op, VirtualGateEndpoint):
                q = ins.qubits[0]
                if q.register is not fragment:
                    raise KeyError(q)
                vg = op.virtual_gate
                tkey = (type(vg), tuple(vg.params), op.qubit_idx)
                o = mat_off.get(tkey)
                pre, meas, post = _endpoint_table(vg, op.qubit_idx)
                if o is None:
                    o = mat_off[tkey] = pool_len
                    pool.append(pre.reshape(-1).view(np.float64))
                    pool.append(post.reshape(-1).view(np.float64))
                    pool_len += 16 * len(pre)
                add((IN_ENDPOINT, q.index, 0, -1, -1, len(eps)))
                eps.append((op.vgate_idx, op.qubit_idx, len(pre), sum(1 << v for v, f in enumerate(meas) if f),
                            o, o + 8 * len(pre)))
                tables.append((pre, meas, post))
            elif isinstance(op, Barrier):
                continue
            elif isinstance(op, Measure):
                q, c = ins.qubits[0], ins.clbits[0]
                if q.register is not fragment:
                    raise KeyError(q)
                add((IN_MEASURE, q.index, 0, creg_off[id(c.register)] + c.index, -1, -1))
            else:
                raise TypeError(f"cannot lower operation {op!r}")
    except KeyError:
        raise ValueError("Circuit contains gates that act on multiple fragments.") from None
    flat = FlatCircuit()
    flat.n_qubits, flat.n_clbits = n_q, off
    flat.instr = np.asarray(rows, dtype=np.int32).reshape(-1, 6)
    flat.endpoints = np.asarray(eps, dtype=np.int32).reshape(-1, 6)
    flat.pool = np.concatenate(pool) if pool else np.zeros(0, np.float64)
    flat.tables = tables
    flat.key = (n_q, off, flat.instr.tobytes(), flat.endpoints.tobytes(), flat.pool.tobytes())
    return flat


def _host_get(lib, ptr, what: int, dtype) -> np.ndarray:
    n = lib.qck_host_program_get(ptr, what, None, 0)
    if n < 0:
        raise RuntimeError("qck_host_program_get failed")
    out = np.empty(n // np.dtype(dtype).itemsize, dtype=dtype)
    if n:
        lib.qck_host_program_get(ptr, what, out.ctypes.data, n)
    return out


@dataclass
class Slot:
    digit: int                  # position of the owning virtual gate in the fragment label
    vgate_idx: int
    side: int
    qubit: int                  # local qubit (state bit position)
    terminal: bool              # nothing acts on the qubit afterwards
    pre: list                   # per variant: 2x2 complex matrix
    meas: list                  # per variant: bool
    post: list                  # per variant: 2x2 complex matrix
    pre_off: int = -1           # matrix-pool offsets (doubles); -1: all identity
    post_off: int = -1


@dataclass
class TreeLevel:
    kind: int                   # _lib.TREE_SLOT / TREE_TERMINAL / TREE_MMEAS
    qubit: int
    digit: int                  # label digit of the slot's gate (-1: mid-circuit measurement of the input)
    pre_off: int
    post_off: int
    choices: list               # [(representative variant, outcome or -1)]
    canon: list                 # per variant: its representative
    meas: list                  # per variant: measures?
    col_bit: int                # TREE_MMEAS: row bit of the outcome
    seg: tuple = (0, 0)         # ops after this branching op


@dataclass
class TreeHost:
    n_base: int
    ops: np.ndarray             # [n_ops, 8] int32: label-independent ops on state bit positions
    seg0: tuple
    levels: list
    free: list                  # [(row bit, state position)]
    n_out_bits: int
    base_sum: int
    node_counts: list           # nodes before each level, then the leaves


@dataclass
class PlanHost:
    pattern: int
    labels: np.ndarray                  # int32 fragment-label indices using this program
    n_state: int
    ops: np.ndarray                     # [n_ops, 8] int32
    sweeps: list                        # [(positions, op_begin, op_end)]
    out_pos: list
    sum_mask: int
    sign_mask: int
    op_base: int = 0                    # offset of this plan's ops in the fragment's device op array
    shared_prefix: bool = False         # on-chip: sweeps[0] holds no label-dependent op and runs once per plan
    warp_base: int = 0                  # > 0: register-resident plan of that many fragment qubits (the state bits
    #                                     above are branch outcomes, not amplitudes the kernel holds)


class FragmentProgram:
    """Template + per-pattern programs of one fragment."""

    def __init__(self, frag_circuit: QuantumCircuit, fragment: QuantumRegister, num_clbits: int,
                 onchip_max: int = ONCHIP_MAX_QUBITS, stream_tile: int = STREAM_TILE,
                 cluster: bool = True, fuse: bool = True, early_bits: int = 0, share_prefix=None,
                 warp=None, flat: "FlatCircuit | None" = None, native=None) -> None:
        self.share_prefix = SHARE_PREFIX if share_prefix is None else share_prefix      # True / False / "auto"
        # register-resident regime wanted?  (only default-shaped programs: the knobs below select other kernels)
        self.warp_wanted = (WARP if warp is None else bool(warp)) and onchip_max == ONCHIP_MAX_QUBITS \
            and self.share_prefix is not True and cluster and fuse and not early_bits
        self.fragment = fragment
        self.n_qubits = len(fragment)
        self.num_clbits = num_clbits
        self.onchip_max = onchip_max
        self.stream_tile = stream_tile
        self.cluster = cluster
        self.fuse = fuse
        # state bits the FIRST sweep of a streaming plan should take into its tile (sharded runs: the rank
        # bits - they then go live while the state is one tile, and the big expansion sweeps stay local)
        self.early_bits = int(early_bits)
        self._pool: list[np.ndarray] = []
        self._pool_len = 0
        self._mat_by_off: dict[int, np.ndarray] = {}
        self.native = NATIVE if native is None else bool(native)
        if self.native:
            self._lower_native(flat if flat is not None else flatten(frag_circuit, fragment))
        else:
            self._lower(frag_circuit)
        self._plans: dict[bool, list[PlanHost]] = {}

    # ------------------------------------------------------------------ template
    def _add_matrix(self, m: np.ndarray) -> int:
        flat = np.ascontiguousarray(m, dtype=np.complex128).reshape(-1).view(np.float64)
        off = self._pool_len
        self._pool.append(flat)
        self._pool_len += flat.size
        self._mat_by_off[off] = np.asarray(m, dtype=np.complex128)
        return off

    def _lower_native(self, flat: "FlatCircuit") -> None:
        """``_lower`` + ``_fuse_pairs`` in C++ (``qck_host_lower``): same tops, slots and matrix pool."""
        lib = _lib.load()
        ptr = C.c_void_p()
        flags = (1 if self.warp_wanted else 0) | (2 if self.fuse else 0)
        rc = lib.qck_host_lower(flat.instr.ctypes.data, len(flat.instr), flat.endpoints.ctypes.data,
                                len(flat.endpoints), flat.pool.ctypes.data, len(flat.pool), flat.n_qubits,
                                flat.n_clbits, flags, WARP_MAX_QUBITS, WARP_MAX_DEPTH, C.byref(ptr))
        if rc == -1:
            raise ValueError("two terminal measurements write the same clbit")
        if rc != 0:
            raise _lib._EXC.get(rc, RuntimeError)(f"qck_host_lower failed ({rc})")
        self._hp, self._hp_lib = ptr, lib
        self._flat_tables = flat.tables
        meta = _host_get(lib, ptr, 11, np.int32).tolist()
        n_dig, n_touched, n_outc = meta[6], meta[8], meta[9]
        self.vgate_indices = meta[10:10 + n_touched]
        self.radix = meta[10 + n_touched:10 + n_touched + n_dig]
        self.out_clbits = meta[10 + n_touched + n_dig:10 + n_touched + n_dig + n_outc]
        self._pool_len = -1                    # the pool is fetched when somebody asks for it
        del self._mat_by_off                   # (materialised on demand, with tops / slots / qubit_order / out_bits)
        self.warp = bool(meta[1])
        self.has_mid_measure = meta[2] > 0
        self.out_mask = 0
        for c in self.out_clbits:
            self.out_mask |= 1 << c
        self.measures_anything = bool(meta[3])
        self.num_labels = 1
        for r in self.radix:
            self.num_labels *= r

    _LAZY = ("tops", "slots", "qubit_order", "out_bits", "_mat_by_off")

    def __getattr__(self, name):
        # native programs keep their op list in the C++ object; the Python views of it (tests, bench statistics,
        # the reference-faithful knit's slot table) are built on first use
        if name in FragmentProgram._LAZY and self.__dict__.get("_hp") is not None:
            self._materialise()
            return self.__dict__[name]
        raise AttributeError(name)

    def work_estimate(self) -> float:
        """Sum over the labels of 2^(state bits) x op records of the label's program (``dist.simulation_work``)."""
        return float(_host_get(self._hp_lib, self._hp, 12, np.float64)[0])

    def _native_pool(self) -> np.ndarray:
        n = self._hp_lib.qck_host_program_get(self._hp, 9, None, 0) // 8
        if n != self._pool_len:
            pool = _host_get(self._hp_lib, self._hp, 9, np.float64)
            self._pool, self._pool_len, self._mats_cache = ([pool] if len(pool) else []), len(pool), None
        return self._pool[0] if self._pool else np.zeros(0)

    def _materialise(self) -> None:
        lib, ptr = self._hp_lib, self._hp
        tops = _host_get(lib, ptr, 1, np.int32).reshape(-1, 4).tolist()
        slot_rows = _host_get(lib, ptr, 2, np.int32).reshape(-1, 9).tolist()
        slot_pre = _host_get(lib, ptr, 3, np.float64).view(np.complex128).reshape(-1, _lib.MAX_VARIANTS, 2, 2)
        order = _host_get(lib, ptr, 4, np.int32).tolist()
        out_bits = _host_get(lib, ptr, 5, np.int32).reshape(-1, 2).tolist()
        frag_qubits = list(self.fragment)
        d = self.__dict__
        d["qubit_order"] = [frag_qubits[i] for i in order]
        cview = self._native_pool().view(np.complex128)
        mats = d["_mat_by_off"] = {}
        slots = []
        for i, (digit, vgate_idx, side, qubit, terminal, n_var, _mask, pre_off, post_off) in enumerate(slot_rows):
            _pre, meas, post = self._flat_tables[i]
            slots.append(Slot(digit=digit, vgate_idx=vgate_idx, side=side, qubit=qubit, terminal=bool(terminal),
                              pre=list(slot_pre[i, :n_var]), meas=list(meas), post=list(post),
                              pre_off=pre_off, post_off=post_off))
            if pre_off >= 0:
                mats[pre_off] = cview[pre_off // 2:pre_off // 2 + 4 * n_var].reshape(n_var, 2, 2)
            if post_off >= 0:
                mats[post_off] = cview[post_off // 2:post_off // 2 + 4 * n_var].reshape(n_var, 2, 2)
        d["slots"] = slots
        out = []
        for kind, a, b, c in tops:
            if kind == T_U1:
                mats[b] = cview[b // 2:b // 2 + 4].reshape(2, 2)
                out.append(("u1", a, b))
            elif kind == T_U2:
                mats[c] = cview[c // 2:c // 2 + 16].reshape(4, 4)
                out.append(("u2", a, b, c))
            elif kind == T_SLOT:
                out.append(("slot", a))
            else:
                out.append((_TOP_NAMES[kind], a, b))
        d["tops"] = out
        d["out_bits"] = [(c, p) for c, p in out_bits]

    def _lower(self, circ: QuantumCircuit) -> None:
        frag_qubits = list(self.fragment)
        qset = set(frag_qubits)
        # (plain gates are the common case: `type(op) is Gate` spares them the abstract-base-class checks)
        instrs = [ins for ins in circ.data
                  if type(ins.operation) is Gate or not (isinstance(ins.operation, Barrier)
                                                         and not isinstance(ins.operation, VirtualGateEndpoint))]
        for ins in instrs:
            for q in ins.qubits:
                if q not in qset:
                    raise ValueError("Circuit contains gates that act on multiple fragments.")
        last_use = {}
        for i, ins in enumerate(instrs):
            for q in ins.qubits:
                last_use[q] = i
        # finally-measured clbits (terminal measurements) define the output row
        terminal = []                    # (clbit index, qubit)
        for i, ins in enumerate(instrs):
            if type(ins.operation) is not Gate and isinstance(ins.operation, Measure) and last_use[ins.qubits[0]] == i:
                terminal.append((circ.clbit_index(ins.clbits[0]), ins.qubits[0]))
        terminal.sort(key=lambda t: t[0])
        if len({c for c, _ in terminal}) != len(terminal):
            raise ValueError("two terminal measurements write the same clbit")
        order = [q for _, q in terminal] + [q for q in frag_qubits if q not in {t[1] for t in terminal}]
        pos = {q: i for i, q in enumerate(order)}
        self.qubit_order = order

        # virtual gates touching this fragment, in circuit (= vgate index) order
        touched = sorted({ins.operation.vgate_idx for ins in instrs
                          if isinstance(ins.operation, VirtualGateEndpoint)})
        self.vgate_indices = touched
        digit_of = {k: d for d, k in enumerate(touched)}
        self.radix = [0] * len(touched)

        out_bits = [(c, pos[q]) for c, q in terminal]      # (clbit, state position or ("anc", i))
        tops: list[tuple] = []
        slots: list[Slot] = []
        pending: dict[int, np.ndarray] = {}

        def flush(q: int) -> None:
            m = pending.pop(q, None)
            if m is not None and not _is_identity(m):
                tops.append(("u1", q, self._add_matrix(m)))

        mid_measures = 0
        for i, ins in enumerate(instrs):
            op = ins.operation
            qs = [pos[q] for q in ins.qubits]
            if type(op) is not Gate and isinstance(op, VirtualGateEndpoint):
                vg: VirtualBinaryGate = op.virtual_gate
                table = vg._table()
                d = digit_of[op.vgate_idx]
                self.radix[d] = len(table)
                pre, meas, post = [], [], []
                for inst in table:
                    a, b, seen = _I2, _I2, False
                    for name, params, q in inst:
                        if q != op.qubit_idx:
                            continue
                        if name == "measure":
                            if seen:
                                raise ValueError("instantiation measures a qubit twice")
                            seen = True
                            continue
                        m = Gate(name, 1, params).to_matrix()
                        if seen:
                            b = m @ b
                        else:
                            a = m @ a
                    pre.append(a)
                    meas.append(seen)
                    post.append(b)
                pend = pending.pop(qs[0], None)
                if pend is not None:
                    pre = [a @ pend for a in pre]
                slot = Slot(digit=d, vgate_idx=op.vgate_idx, side=op.qubit_idx, qubit=qs[0],
                            terminal=(last_use[ins.qubits[0]] == i), pre=pre, meas=meas, post=post)
                if not all(_is_identity(a) for a in pre):
                    slot.pre_off = self._add_matrix(np.stack(pre))
                if not slot.terminal and not all(_is_identity(b) for b in post):
                    slot.post_off = self._add_matrix(np.stack(post))
                tops.append(("slot", len(slots)))
                slots.append(slot)
            elif type(op) is not Gate and isinstance(op, Measure):
                flush(qs[0])
                if last_use[ins.qubits[0]] != i:          # mid-circuit measurement of the input circuit
                    tops.append(("mmeas", qs[0], circ.clbit_index(ins.clbits[0])))
                    mid_measures += 1
            elif isinstance(op, Gate):
                if op.num_qubits == 1:
                    m = op.to_matrix()
                    pending[qs[0]] = m @ pending[qs[0]] if qs[0] in pending else m
                elif op.num_qubits == 2:
                    flush(qs[0]); flush(qs[1])
                    if op._matrix is None and op.name == "cx":
                        tops.append(("cx", qs[0], qs[1]))
                    elif op._matrix is None and op.name == "cz":
                        tops.append(("cz", qs[0], qs[1]))
                    else:
                        tops.append(("u2", qs[0], qs[1], self._add_matrix(op.to_matrix())))
                else:
                    raise NotImplementedError(f"{op.name}: gates on more than two qubits are not supported")
            else:
                raise TypeError(f"cannot lower operation {op!r}")
        # one-qubit gates still pending act on wires that are never measured afterwards: they
        # cannot change any probability and are dropped
        pending.clear()
        self.slots = slots
        branch_points = mid_measures + sum(1 for sl in slots if not sl.terminal and any(sl.meas))
        self.warp = (self.warp_wanted and self.n_qubits <= WARP_MAX_QUBITS and branch_points <= WARP_MAX_DEPTH
                     and not any(t[0] == "u2" for t in tops) and len(tops) + 2 * len(slots) < 60000)
        self.tops = tops if self.warp or not self.fuse else self._fuse_pairs(tops)
        self.out_bits = out_bits
        # the row of an instance has one bit per WRITTEN clbit, in ascending clbit order: the terminal
        # measurements and the mid-circuit measurements of the input circuit (their outcome lives on an
        # ancilla, see _build_plan) - masks, key_mask and row_bits all follow this list
        self.out_clbits = sorted([c for c, _ in out_bits] + [t[2] for t in self.tops if t[0] == "mmeas"])
        self.has_mid_measure = mid_measures > 0
        self.out_mask = 0
        for c in self.out_clbits:
            self.out_mask |= 1 << c
        # does any instance of this fragment measure anything at all?  (run.py:49-58 drops
        # fragments whose circuits have no measurement)
        self.measures_anything = bool(out_bits) or mid_measures > 0 or any(any(s.meas) for s in slots)
        self.num_labels = int(np.prod(self.radix, dtype=np.int64)) if self.radix else 1

    def __del__(self):  # pragma: no cover - best effort
        hp = getattr(self, "_hp", None)
        if hp is not None and hp.value:
            self._hp = None
            try:
                self._hp_lib.qck_host_program_free(hp)
            except Exception:
                pass

    def _native_build(self, stage: int, fold: bool = True) -> int:
        """Run one stage of the C++ planner (0 tree, 1 plans, 2 host image, 3 canonical labels) with the CURRENT
        module knobs, then pick up what it appended to the matrix pool."""
        lib = self._hp_lib
        mode = {True: 1, False: 0}.get(self.share_prefix, 2)
        rc = lib.qck_host_program_configure(self._hp, self.onchip_max, min(self.stream_tile, _lib.MAX_TILE_QUBITS),
                                            int(self.cluster), mode, int(TREE), {True: 1, False: 0}.get(DEDUPE, 2),
                                            C.c_uint64(self.early_bits))
        if rc == 0:
            rc = lib.qck_host_program_build(self._hp, stage, int(bool(fold)))
        if rc == -2:
            raise ValueError("a clbit is written by more than one measurement")
        if rc != 0 and not (stage == 0 and rc == 1):
            raise _lib._EXC.get(rc, RuntimeError)(f"host planner failed ({rc}): program out of range "
                                                  "(state > 40 qubits, an op that does not fit a tile, both ends of "
                                                  "a virtual gate measuring in one fragment, or too many instances)")
        return rc

    def _native_plans(self, fold: bool) -> list:
        self._native_build(1, fold)
        lib, hp, base = self._hp_lib, self._hp, 100 + (20 if fold else 0)
        meta = _host_get(lib, hp, base, np.int64).reshape(-1, 12).tolist()
        labels = _host_get(lib, hp, base + 1, np.int32)
        ops = _host_get(lib, hp, base + 2, np.int32).reshape(-1, 8)
        sweeps = _host_get(lib, hp, base + 3, np.int32).reshape(-1, 43).tolist()
        out_pos = _host_get(lib, hp, base + 4, np.int32).tolist()
        plans, l0, o0, s0, p0 = [], 0, 0, 0, 0
        for (pattern, n_labels, n_state, n_ops, n_sweeps, n_out, shared, warp_base, op_base, sum_mask, sign_mask,
             _z) in meta:
            sw = [(rec[3:3 + rec[0]], rec[1], rec[2]) for rec in sweeps[s0:s0 + n_sweeps]]
            plans.append(PlanHost(pattern, labels[l0:l0 + n_labels], n_state, ops[o0:o0 + n_ops], sw,
                                  out_pos[p0:p0 + n_out], sum_mask & (2 ** 64 - 1), sign_mask & (2 ** 64 - 1),
                                  op_base=op_base, shared_prefix=bool(shared), warp_base=warp_base))
            l0, o0, s0, p0 = l0 + n_labels, o0 + n_ops, s0 + n_sweeps, p0 + n_out
        return plans

    def _native_tree(self):
        if self._native_build(0) != 1:
            return None
        lib, hp = self._hp_lib, self._hp
        n_base, n_out_bits, seg0_b, seg0_e, base_sum, _n_levels, _n_ops, off_ops = _host_get(lib, hp, 50, np.int64).tolist()
        ops = _host_get(lib, hp, 51, np.int32).reshape(-1, 8)
        levels = []
        for rec in _host_get(lib, hp, 52, np.int32).reshape(-1, 58).tolist():
            kind, qubit, digit, pre_off, post_off, col_bit, seg_b, seg_e, n_ch, n_var = rec[:10]
            levels.append(TreeLevel(kind, qubit, digit, pre_off, post_off,
                                    [(rec[10 + 2 * c], rec[11 + 2 * c]) for c in range(n_ch)],
                                    rec[42:42 + n_var], [bool(m) for m in rec[50:50 + n_var]], col_bit, (seg_b, seg_e)))
        free = [tuple(x) for x in _host_get(lib, hp, 53, np.int32).reshape(-1, 2).tolist()]
        counts = _host_get(lib, hp, 54, np.int64).tolist()
        tree = TreeHost(n_base, ops, (seg0_b, seg0_e), levels, free, n_out_bits, base_sum & (2 ** 64 - 1), counts)
        st = _lib.QckSimTreePlan.from_buffer_copy(_host_get(lib, hp, 55, np.uint8).tobytes())
        self._tree_image = (_host_get(lib, hp, 56, np.uint8), off_ops, st)
        return tree

    @property
    def mats(self) -> np.ndarray:
        """The matrix pool (float64, interleaved re/im).  Planning appends the tile-resolved variants
        of ops whose diag qubits stay outside a tile, so read it AFTER ``plans()``."""
        if self.native:
            pool = self._native_pool()
            return pool if len(pool) else np.zeros(8)
        cached = getattr(self, "_mats_cache", None)
        if cached is None or cached[0] != len(self._pool):
            arr = np.concatenate(self._pool) if self._pool else np.zeros(8)
            self._mats_cache = cached = (len(self._pool), arr)
        return cached[1]

    # ------------------------------------------------------------------ gate fusion
    _SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
    _CX01 = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)
    _CZ = np.diag([1, 1, 1, -1]).astype(np.complex128)
    _OVERHEAD = 6.0      # per-op dispatch cost in units of FP64 instructions per amplitude.  (A larger value
    #                      for the streaming kernel - fewer, denser passes: bv-32 90 -> 61 - measured slower:
    #                      hwe-32 d2 77 -> 85 ms; a cx pass is a pure swap, a fused 4x4 is 64 DFMA per quad.)

    def _fuse_pairs(self, tops: list[tuple]) -> list[tuple]:
        """Merge runs of gates that act inside one qubit pair into a single 4x4 unitary when a
        cost model (FP64 work per amplitude + per-op overhead) says the dense product is cheaper
        than the separate gates.  Slots and mid-circuit measurements are barriers on their qubit.
        Only reorders gates on disjoint qubits; per-qubit order is preserved."""
        out: list[tuple] = []
        group_of: dict[int, dict] = {}          # qubit -> open pair group
        pending: dict[int, list] = {}           # qubit -> 1q ops waiting for a partner

        def op_cost(t) -> float:
            if t[0] == "u1":
                m = self._mat_by_off[t[2]]
                return 2.0 if (m[0, 1] == 0 and m[1, 0] == 0) else 8.0
            if t[0] in ("cx", "cz"):
                return 1.0
            return 16.0

        def embed(t, a, b) -> np.ndarray:
            if t[0] == "u1":
                u = self._mat_by_off[t[2]]
                m = np.zeros((4, 4), dtype=np.complex128)
                if t[1] == a:                       # kron(I, u): acts on index bit 0
                    m[0:2, 0:2] = u
                    m[2:4, 2:4] = u
                else:                               # kron(u, I): acts on index bit 1
                    m[0::2, 0::2] = u
                    m[1::2, 1::2] = u
                return m
            m = self._CX01 if t[0] == "cx" else self._CZ if t[0] == "cz" else self._mat_by_off[t[3]].reshape(4, 4)
            return m if (t[1], t[2]) == (a, b) else self._SWAP @ m @ self._SWAP

        def close(g) -> None:
            for q in (g["a"], g["b"]):
                if group_of.get(q) is g:
                    del group_of[q]
            ops = g["ops"]
            separate = sum(op_cost(t) + self._OVERHEAD for t in ops)
            if len(ops) > 1 and separate > 16.0 + self._OVERHEAD:
                m = np.eye(4, dtype=np.complex128)
                for t in ops:
                    m = embed(t, g["a"], g["b"]) @ m
                out.append(("u2", g["a"], g["b"], self._add_matrix(m)))
            else:
                out.extend(ops)

        def flush_pending(q) -> None:
            out.extend(pending.pop(q, []))

        for t in tops:
            if t[0] == "u1":
                q = t[1]
                if q in group_of:
                    group_of[q]["ops"].append(t)
                else:
                    pending.setdefault(q, []).append(t)
            elif t[0] in ("cx", "cz", "u2"):
                a, b = t[1], t[2]
                g = group_of.get(a)
                if g is not None and g is group_of.get(b):
                    g["ops"].append(t)
                    continue
                for q in (a, b):
                    if q in group_of:
                        close(group_of[q])
                g = {"a": a, "b": b, "ops": pending.pop(a, []) + pending.pop(b, []) + [t]}
                group_of[a] = group_of[b] = g
            else:                                   # slot / mmeas: barrier on its qubit
                q = self.slots[t[1]].qubit if t[0] == "slot" else t[1]
                if q in group_of:
                    close(group_of[q])
                flush_pending(q)
                out.append(t)
        for g in list({id(g): g for g in group_of.values()}.values()):
            close(g)
        for q in list(pending):
            flush_pending(q)
        return out

    # ------------------------------------------------------------------ tree-walk program
    def tree(self):
        """The fragment as ONE tree program (TreeHost), or None when it is not eligible: register regime only, at
        least one branching op, every virtual gate with exactly one endpoint here, folded rows."""
        cached = getattr(self, "_tree", False)
        if cached is not False:
            return cached
        self._tree = None
        if not (self.warp and TREE):
            return None
        if self.native:
            self._tree = self._native_tree()
            return self._tree
        rows, levels, seen = [], [], set()
        seg_begin, seg0 = 0, None
        mmeas_clbits = [t[2] for t in self.tops if t[0] == "mmeas"]
        bits = sorted([(c, p) for c, p in self.out_bits] + [(c, ("mmeas", c)) for c in mmeas_clbits],
                      key=lambda t: t[0])
        row_bit_of_clbit = {c: j for j, (c, _) in enumerate(bits)}
        free = [(j, p) for j, (c, p) in enumerate(bits) if not isinstance(p, tuple)]

        def close_segment():
            nonlocal seg_begin, seg0
            seg = (seg_begin, len(rows))
            if levels:
                levels[-1].seg = seg
            else:
                seg0 = seg
            seg_begin = len(rows)

        for top in self.tops:
            if top[0] == "u1":
                rows.append([_lib.OP_U1, top[1], 0, top[2], -1, 0, 0, 0])
            elif top[0] == "cx":
                rows.append([_lib.OP_CX, top[1], top[2], 0, -1, 0, 0, 0])
            elif top[0] == "cz":
                rows.append([_lib.OP_CZ, top[1], top[2], 0, -1, 0, 0, 0])
            elif top[0] == "mmeas":
                close_segment()
                levels.append(TreeLevel(_lib.TREE_MMEAS, top[1], -1, -1, -1, [(0, 0), (0, 1)], [0], [True],
                                        row_bit_of_clbit[top[2]]))
            elif top[0] == "slot":
                slot = self.slots[top[1]]
                if slot.digit in seen:
                    return None                      # both ends of a virtual gate in one fragment
                seen.add(slot.digit)
                close_segment()
                r = self.radix[slot.digit]
                first, canon = {}, []
                for v in range(r):
                    key = ((slot.pre[v] + 0.0).tobytes(), bool(slot.meas[v]), (slot.post[v] + 0.0).tobytes())
                    canon.append(first.setdefault(key, v))
                choices = []
                for v in sorted(set(canon)):
                    if slot.meas[v] and not slot.terminal:
                        choices += [(v, 0), (v, 1)]
                    else:
                        choices.append((v, -1))
                levels.append(TreeLevel(_lib.TREE_TERMINAL if slot.terminal else _lib.TREE_SLOT, slot.qubit, slot.digit,
                                        slot.pre_off, slot.post_off if not slot.terminal else -1, choices, canon,
                                        [bool(m) for m in slot.meas], -1))
            else:
                return None
        close_segment()
        if not levels or len(levels) > _lib.TREE_MAX_LEVELS or any(len(l.choices) > _lib.TREE_MAX_CHOICES for l in levels):
            return None
        if len(seen) != len(self.radix):
            return None
        counts = [1]
        for l in levels:
            counts.append(counts[-1] * len(l.choices))
        state_bytes = 16 << max(self.n_qubits, 5)
        inner = max(counts[1:-1], default=0)
        if 2 * inner * state_bytes + counts[-1] * (8 << len(free)) > TREE_MAX_WORK_BYTES or counts[-1] >= 2 ** 31:
            return None
        used = {p for _, p in free}
        base_sum = sum(1 << b for b in range(self.n_qubits) if b not in used)
        ops = np.asarray(rows, dtype=np.int32).reshape(-1, 8) if rows else np.zeros((0, 8), np.int32)
        self._tree = TreeHost(self.n_qubits, ops, seg0, levels, free, len(bits), base_sum, counts)
        return self._tree

    # ------------------------------------------------------------------ identical instances
    def canonical_labels(self) -> np.ndarray:
        """int32 [num_labels]: the smallest label whose instance is identical to this one (itself for a
        representative).  Two variants of a gate are identical in this fragment when every slot of that gate here has
        bit-identical pre / post matrices and the same measure flag for both."""
        cached = getattr(self, "_canon", None)
        if cached is not None:
            return cached
        if self.native:
            self._native_build(3)
            self._canon = _host_get(self._hp_lib, self._hp, 60, np.int32)
            return self._canon
        n_dig = len(self.radix)
        maps = []
        for d in range(n_dig):
            first: dict = {}
            m = np.zeros(self.radix[d], dtype=np.int64)
            for v in range(self.radix[d]):
                key = tuple(((s.pre[v] + 0.0).tobytes(), bool(s.meas[v]), (s.post[v] + 0.0).tobytes())
                            for s in self.slots if s.digit == d)
                m[v] = first.setdefault(key, v)
            maps.append(m)
        labels = np.arange(self.num_labels, dtype=np.int64)
        if n_dig:
            digits = self.label_digits(labels)
            canon = np.stack([maps[d][digits[:, d]] for d in range(n_dig)], axis=0)
            labels = np.ravel_multi_index(tuple(canon), self.radix)
        self._canon = labels.astype(np.int32)
        return self._canon

    # ------------------------------------------------------------------ patterns
    def label_digits(self, labels: np.ndarray) -> np.ndarray:
        if not self.radix:
            return np.zeros((len(labels), 0), dtype=np.int64)
        return np.stack(np.unravel_index(np.asarray(labels, dtype=np.int64), self.radix), axis=1)

    def plans(self, fold: bool = True) -> list[PlanHost]:
        if fold in self._plans:
            return self._plans[fold]
        if self.native:
            self._plans[fold] = self._native_plans(fold)
            return self._plans[fold]
        labels = np.arange(self.num_labels, dtype=np.int64)
        digits = self.label_digits(labels)
        pat = np.zeros(self.num_labels, dtype=np.int64)
        for s, slot in enumerate(self.slots):
            flags = np.asarray(slot.meas, dtype=np.int64)[digits[:, slot.digit]]
            pat |= flags << s
        plans = []
        op_base = 0
        for p in np.unique(pat):
            plan = self._build_plan(int(p), labels[pat == p].astype(np.int32), fold)
            plan.op_base = op_base
            op_base += len(plan.ops)
            plans.append(plan)
        self._plans[fold] = plans
        return plans

    def _plan_template(self):
        """Everything ``_build_plan`` needs that does not depend on the measurement pattern: the op rows as if every
        non-terminal slot measured, with per-row marks - ``opt`` (slot whose pattern bit decides whether the row
        exists, -1: always), ``alloc`` (the row is the CX onto a fresh ancilla), ``clbit`` (mid-circuit measurement
        of the input circuit: the clbit it writes, else -1).  (The per-pattern Python emit loop this replaces was
        a third of the cold compile of hwe-16 d5: 64 programs x 150 ops.)"""
        cached = getattr(self, "_template", None)
        if cached is not None:
            return cached
        rows, opt, alloc, clbit = [], [], [], []

        def emit(kind, q0, q1=0, mat=0, sel=-1, stride=0, o=-1, a=False, c=-1):
            rows.append([kind, q0, q1, mat, sel, stride, 0, 0])
            opt.append(o)
            alloc.append(a)
            clbit.append(c)

        for top in self.tops:
            if top[0] == "u1":
                emit(_lib.OP_U1, top[1], 0, top[2])
            elif top[0] == "cx":
                emit(_lib.OP_CX, top[1], top[2])
            elif top[0] == "cz":
                emit(_lib.OP_CZ, top[1], top[2])
            elif top[0] == "u2":
                emit(_lib.OP_U2, top[1], top[2], top[3])
            elif top[0] == "mmeas":
                emit(_lib.OP_CX, top[1], 0, a=True, c=top[2])
            else:
                s = top[1]
                slot = self.slots[s]
                if slot.pre_off >= 0:
                    emit(_lib.OP_U1, slot.qubit, 0, slot.pre_off, slot.digit, 8)
                if not slot.terminal:
                    emit(_lib.OP_CX, slot.qubit, 0, o=s, a=True)
                if slot.post_off >= 0:
                    emit(_lib.OP_U1, slot.qubit, 0, slot.post_off, slot.digit, 8)
        self._template = (np.asarray(rows, dtype=np.int32).reshape(-1, 8), np.asarray(opt, dtype=np.int64),
                          np.asarray(alloc, dtype=bool), np.asarray(clbit, dtype=np.int64),
                          [t[1] for t in self.tops if t[0] == "slot"])
        return self._template

    def _build_plan(self, pattern: int, labels: np.ndarray, fold: bool) -> PlanHost:
        n = self.n_qubits
        t_rows, t_opt, t_alloc, t_clbit, slot_order = self._plan_template()
        keep = (t_opt < 0) | (((pattern >> np.maximum(t_opt, 0)) & 1) == 1)
        ops_arr = t_rows[keep]                          # fancy indexing: a copy
        alloc = t_alloc[keep]
        cum = np.cumsum(alloc)                          # ancillas allocated up to and including each row
        ops_arr[:, 6] = n + cum                         # n_live of a row = n + ancillas so far (its own included)
        at = np.nonzero(alloc)[0]
        ops_arr[at, 2] = n + cum[at] - 1                # the CX target: the ancilla the row allocates
        n_anc = int(cum[-1]) if len(cum) else 0
        anc_of_slot: dict[int, int] = {}
        extra_out: list[tuple[int, int]] = []           # (clbit, ancilla position) of mid-circuit measurements
        opt_k, clbit_k = t_opt[keep], t_clbit[keep]
        for r in at.tolist():
            if clbit_k[r] >= 0:
                extra_out.append((int(clbit_k[r]), int(ops_arr[r, 2])))
            else:
                anc_of_slot[int(opt_k[r])] = int(ops_arr[r, 2])
        cfg_pos: dict[int, int] = {}                    # digit -> state bit carrying the config bit
        for s in slot_order:
            if not (pattern >> s) & 1:
                continue
            slot = self.slots[s]
            if slot.digit in cfg_pos:
                raise NotImplementedError("both ends of a virtual gate measure inside one fragment")
            cfg_pos[slot.digit] = slot.qubit if slot.terminal else anc_of_slot[s]
        n_state = n + n_anc
        if n_state > 40:
            raise NotImplementedError(f"instance state of {n_state} qubits is out of range")
        bits = sorted(self.out_bits + extra_out)
        if len({c for c, _ in bits}) != len(bits):
            raise ValueError("a clbit is written by more than one measurement")
        out_pos = [p for _, p in bits]
        if not fold:
            out_pos += [cfg_pos.get(d, -1) for d in range(len(self.radix))]
        used = {p for p in out_pos if p >= 0}
        sum_mask = 0
        for b in range(n_state):
            if b not in used:
                sum_mask |= 1 << b
        sign_mask = 0
        if fold:
            for p in cfg_pos.values():
                sign_mask |= 1 << p
        shared = False
        if self.warp:
            # one "sweep" over the whole program; the kernel holds the n fragment qubits and walks the outcomes
            # of the n_anc branch points depth first (positions listed for the plan interpreter: all n_state bits)
            return PlanHost(pattern, labels, n_state, ops_arr, [(list(range(n_state)), 0, len(ops_arr))], out_pos,
                            sum_mask, sign_mask, warp_base=n)
        if n_state <= self.onchip_max:
            sweeps = [(list(range(n_state)), 0, len(ops_arr))]
            dep = np.nonzero(ops_arr[:, 4] >= 0)[0]       # ops that select their matrix by a label digit
            split = int(dep[0]) if len(dep) else 0
            if split > 0 and (self.share_prefix is True or (
                    self.share_prefix == "auto" and self.num_labels >= SHARE_PREFIX_MIN_INSTANCES
                    and len(labels) >= SHARE_PREFIX_MIN_PLAN and split >= SHARE_PREFIX_MIN_FRACTION * len(ops_arr))):
                shared = True
                sweeps = [(list(range(n_state)), 0, split), (list(range(n_state)), split, len(ops_arr))]
        else:
            ops_arr, sweeps = _schedule_sweeps(ops_arr, n_state, self.stream_tile, self)
        # register clusters pay in the latency-bound on-chip kernel; in the streaming kernel the plain
        # passes (matrix in registers, no per-op dispatch) measured faster (hwe-30 d3: 62 -> 45 ms)
        if self.cluster and n_state <= self.onchip_max:
            ops_arr, sweeps = _cluster_sweeps(ops_arr, sweeps)
        return PlanHost(pattern, labels, n_state, ops_arr, sweeps, out_pos, sum_mask, sign_mask, shared_prefix=shared)

    @property
    def row_bits(self) -> int:
        """log2 of the folded output row length."""
        return len(self.out_clbits)       # terminal measurements + mid-circuit measurements of the input circuit

    def row_len(self, fold: bool = True) -> int:
        bits = self.row_bits + (0 if fold else len(self.radix))
        return 1 << bits


def _op_qubits(op) -> tuple[int, ...]:
    return (int(op[1]),) if op[0] == _lib.OP_U1 else (int(op[1]), int(op[2]))


_DIAG_TOL = 1e-13           # |entry| below this counts as a structural zero of a fused matrix

_X2 = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Z2 = np.diag([1, -1]).astype(np.complex128)


def _op_roles(row, program) -> tuple[int, int]:
    """-> (need mask, diag mask) of one op record.  A qubit is *diag* when the op is block diagonal
    with respect to it (it never mixes amplitudes that differ in that bit): the control of a cx, both
    qubits of cz / cp / rzz, the qubit of rz / z / s / t / p.  Such a qubit need not be in the tile of
    the sweep that applies the op: its bit is a constant of the tile and just selects a variant
    (qck.h: QCK_OP_U1X / QCK_OP_PHASE).  Diag uses of one qubit also commute with each other."""
    kind, q0, q1, mat, sel = row[0], row[1], row[2], row[3], row[4]
    if kind == _lib.OP_U1:
        if sel < 0:
            m = program._mat_by_off[mat]
            if abs(m[0, 1]) < _DIAG_TOL and abs(m[1, 0]) < _DIAG_TOL:
                return 0, 1 << q0
        return 1 << q0, 0
    if kind == _lib.OP_CX:
        return 1 << q1, 1 << q0
    if kind == _lib.OP_CZ:
        return 0, (1 << q0) | (1 << q1)
    m = np.abs(program._mat_by_off[mat].reshape(4, 4)) >= _DIAG_TOL
    r, c = np.nonzero(m)
    d0 = bool(np.all((r & 1) == (c & 1)))         # block diagonal w.r.t. q0 (index bit 0)
    d1 = bool(np.all((r >> 1) == (c >> 1)))       # ... w.r.t. q1 (index bit 1)
    need = (0 if d0 else 1 << q0) | (0 if d1 else 1 << q1)
    diag = (1 << q0 if d0 else 0) | (1 << q1 if d1 else 0)
    return need, diag


def _cond_matrices(row, program, ext: int) -> tuple[np.ndarray, np.ndarray]:
    """The two 2x2 matrices a two-qubit op applies to its OTHER qubit when its (diag) qubit `ext`
    is 0 / 1."""
    kind, q0, q1, mat = row[0], row[1], row[2], row[3]
    if kind == _lib.OP_CX:
        return _I2, _X2
    if kind == _lib.OP_CZ:
        return _I2, _Z2
    m = program._mat_by_off[mat].reshape(4, 4)
    if ext == q1:
        return m[0:2, 0:2].copy(), m[2:4, 2:4].copy()
    return m[0::2, 0::2].copy(), m[1::2, 1::2].copy()


def _phase_scalars(row, program) -> np.ndarray:
    """Diagonal of an op all of whose qubits are diag: 2 scalars (one qubit) or 4 (index
    bit(q0) + 2 bit(q1))."""
    kind, mat = row[0], row[3]
    if kind == _lib.OP_U1:
        return np.diag(program._mat_by_off[mat]).copy()
    if kind == _lib.OP_CZ:
        return np.array([1, 1, 1, -1], dtype=np.complex128)
    return np.diag(program._mat_by_off[mat].reshape(4, 4)).copy()


def _schedule_sweeps(ops: np.ndarray, n_state: int, tile: int, program):
    """Greedy list scheduling: each sweep takes, in program order, every op that is not blocked by
    an earlier unscheduled op and whose *need* qubits still fit into the tile (bit-mask sets).  Diag
    qubits (see _op_roles) do not have to fit; when they end up outside the tile the op is emitted
    in its tile-resolved form:

    * one diag qubit outside: a one-qubit op on the other qubit whose matrix is selected by the
      outside bit; consecutive ones on the same tile qubit are chained into ONE ``OP_U1X`` header
      followed by ``OP_TERM`` records (the device multiplies the selected matrices once per tile);
    * every qubit outside: a scalar per tile; all of them in a sweep are collected into one
      ``OP_PHASE`` header + terms at the start of the sweep (they commute with everything in it).
    """
    tile = min(tile, n_state)
    low = max(0, min(LOW_RUN, tile - 2))     # always leave room for a two-qubit gate
    rows = ops.tolist()
    roles = [_op_roles(r, program) for r in rows]
    remaining = list(range(len(rows)))
    new_ops, sweeps = [], []
    while remaining:
        tile_set = (1 << low) - 1
        if not sweeps and bin(tile_set | program.early_bits).count("1") <= tile - 2:
            tile_set |= program.early_bits & ((1 << n_state) - 1)
        blocked_full = blocked_diag = 0
        taken, rest = [], []
        for i in remaining:
            need, diag = roles[i]
            conflict = (need & (blocked_full | blocked_diag)) or (diag & blocked_full)
            if not conflict and bin(tile_set | need).count("1") <= tile:
                tile_set |= need
                taken.append(i)
            else:
                blocked_full |= need
                blocked_diag |= diag
                rest.append(i)
        if not taken:
            raise NotImplementedError("an op does not fit into a tile")
        b = 0
        while bin(tile_set).count("1") < tile:   # pad with the lowest free qubits: longer contiguous runs
            tile_set |= 1 << b
            b += 1
        positions = [q for q in range(n_state) if (tile_set >> q) & 1]
        local = {p: j for j, p in enumerate(positions)}
        begin = len(new_ops)
        items: list = []                     # rows, or open chains {"q": tile qubit, "terms": [...]}
        chain_of: dict[int, dict] = {}       # tile qubit -> its open chain
        phase_terms: list[list[int]] = []
        for i in taken:
            r = list(rows[i])
            qs = _op_qubits(r)
            outside = [q for q in qs if q not in local]
            r[6] = 0                             # n_live is meaningless inside a tile
            if not outside:
                for q in qs:
                    chain_of.pop(q, None)        # anything else on the qubit closes its chain
                r[1] = local[r[1]]
                if r[0] != _lib.OP_U1:
                    r[2] = local[r[2]]
                items.append(r)
            elif len(outside) == len(qs):        # nothing in the tile: a scalar per tile
                sc = _phase_scalars(r, program)
                off = program._add_matrix(sc)
                phase_terms.append([_lib.OP_TERM, outside[0], outside[1] if len(outside) > 1 else -1, off,
                                    -1, len(sc), 0, 1])
            else:                                # one diag qubit outside: conditional op on the other one
                ext = outside[0]
                a = qs[0] if qs[1] == ext else qs[1]
                m0, m1 = _cond_matrices(r, program, ext)
                off = program._add_matrix(np.stack([m0, m1]))
                ch = chain_of.get(a)
                if ch is None or len(ch["terms"]) >= MAX_TERMS:
                    ch = {"q": local[a], "terms": []}
                    chain_of[a] = ch
                    items.append(ch)
                ch["terms"].append([_lib.OP_TERM, ext, -1, off, -1, 8, 0, 1])
        for c0 in range(0, len(phase_terms), MAX_TERMS):
            chunk = phase_terms[c0:c0 + MAX_TERMS]
            new_ops.append([_lib.OP_PHASE, 0, len(chunk), 0, -1, 0, 0, 1])
            new_ops.extend(chunk)
        for it in items:
            if isinstance(it, dict):
                new_ops.append([_lib.OP_U1X, it["q"], len(it["terms"]), 0, -1, 0, 0, 0])
                new_ops.extend(it["terms"])
            else:
                new_ops.append(it)
        sweeps.append((positions, begin, len(new_ops)))
        remaining = rest
    return np.asarray(new_ops, dtype=np.int32).reshape(-1, 8), sweeps


MAX_CLUSTER_OPS = 32        # QCK_MAX_CLUSTER_OPS: a cluster must fit the staged chunk
MAX_TERMS = 30              # terms per OP_U1X / OP_PHASE header (header + terms must fit the staged chunk)


def _cluster_sweeps(ops: np.ndarray, sweeps: list):
    """Group the ops of every sweep into clusters acting on <= 3 tile qubits: list scheduling in
    ``qck_host_cluster_ops`` (csrc/host_compile.cu; the Python original lives on as its reference in
    tests/cluster_reference.py).  A cluster is a QCK_OP_CLUSTER header followed by its member ops, whose
    qubits are re-expressed as ranks among the cluster's three ascending positions."""
    lib = _lib.load()
    ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 8)
    out = np.empty((2 * len(ops) + 1, 8), dtype=np.int32)
    n_out = C.c_int()
    new_sweeps, w = [], 0
    for positions, b, e in sweeps:
        rc = lib.qck_host_cluster_ops(ops[b:e].ctypes.data, e - b, len(positions), MAX_CLUSTER_OPS,
                                      out[w:].ctypes.data, C.byref(n_out))
        if rc != 0:
            raise ValueError(f"qck_host_cluster_ops failed ({rc}): an op addresses a qubit outside its tile")
        new_sweeps.append((positions, w, w + n_out.value))
        w += n_out.value
    return out[:w].copy(), new_sweeps


# ---------------------------------------------------------------------- device side
class _PinnedRing:
    """Per-thread pinned staging ring: program blobs are tiny (KBs), but ``pin_memory()`` per
    array costs a cudaHostAlloc each (hundreds of microseconds).  Slices are handed out
    sequentially; on wrap-around the device is synchronised so no in-flight copy is overwritten."""

    def __init__(self, nbytes: int = 8 << 20) -> None:
        import torch
        self.torch = torch
        self.buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self.off = 0

    def take(self, nbytes: int):
        nbytes = (nbytes + 255) & ~255
        if nbytes > self.buf.numel():
            return self.torch.empty(nbytes, dtype=self.torch.uint8).pin_memory()
        if self.off + nbytes > self.buf.numel():
            self.torch.cuda.synchronize()
            self.off = 0
        out = self.buf[self.off:self.off + nbytes]
        self.off += nbytes
        return out


_ring_tls = __import__("threading").local()


def _pinned_ring() -> _PinnedRing:
    ring = getattr(_ring_tls, "ring", None)
    if ring is None:
        ring = _ring_tls.ring = _PinnedRing()
    return ring


class FragmentExecutor:
    """Uploads a FragmentProgram and runs all (or some) of its instances on the GPU."""

    def __init__(self, program: FragmentProgram, device, fold: bool = True) -> None:
        import torch
        self.torch = torch
        self.program = program
        self.device = torch.device(device)
        self.fold = fold
        self.tree = program.tree() if fold else None
        if self.tree is not None:
            host = program.__dict__.setdefault("_tree_image", None)     # (the native planner leaves it there)
            if host is None:
                host = program._tree_image = self._build_tree_image(program, self.tree)
            self._blob, self._off_ops, self._tree_struct = host
            self.plans, self._structs, self._dedupe = [], [], None
            self.h2d_bytes = int(self._blob.nbytes)
            self.d_blob = None
            self.row_len = program.row_len(fold)
            self.max_state = program.n_qubits
            self.streaming = False
            self._work = None
            self._work_bytes = None
            return
        if not program.native:
            self.plans = program.plans(fold)
        # Host image of the program (blob + plan structs): a pure function of the program, built once and
        # shared by every executor of it (programs are cached process-wide; run() only ever copies the
        # struct templates, so sharing across threads is safe).
        host = program.__dict__.setdefault("_host_images", {}).get(fold)
        if host is None:
            host = (self._native_host_image(program, fold) if program.native
                    else self._build_host_image(program, self.plans))
            program._host_images[fold] = host
        (self._blob, self._off_ops, self._off_labels, self._labels_host, self._sweep_arrays, self._structs,
         self._dedupe) = host
        self.h2d_bytes = int(self._blob.nbytes)
        self.d_blob = None
        self.row_len = program.row_len(fold)
        self.max_state = max(st.n_state_qubits for st, _o, _c in self._structs)
        self.streaming = self.max_state > program.onchip_max and not program.warp
        self._work = None
        self._work_bytes = None

    def __getattr__(self, name):
        if name == "plans" and "program" in self.__dict__:      # native programs: unpacked on first use only
            self.plans = self.program.plans(self.fold)
            return self.plans
        raise AttributeError(name)

    @staticmethod
    def _build_tree_image(program: FragmentProgram, tree: "TreeHost"):
        """-> (blob [mats f64 | ops i32], offset of the ops, qck_sim_tree_plan with the host fields filled)."""
        mats = np.ascontiguousarray(program.mats, dtype=np.float64)
        ops = tree.ops if len(tree.ops) else np.zeros((1, 8), np.int32)
        off_ops = (mats.nbytes + 255) & ~255
        blob = np.zeros(off_ops + ops.nbytes, dtype=np.uint8)
        blob[:mats.nbytes] = mats.view(np.uint8)
        blob[off_ops:] = np.ascontiguousarray(ops).view(np.uint8).reshape(-1)
        st = _lib.QckSimTreePlan()
        st.n_base, st.n_levels, st.n_digits, st.n_out_bits = tree.n_base, len(tree.levels), len(program.radix), tree.n_out_bits
        st.seg0_begin, st.seg0_end = tree.seg0
        st.n_free = len(tree.free)
        for r, (j, p) in enumerate(tree.free):
            st.free_bit[r], st.free_pos[r] = j, p
        st.base_sum = tree.base_sum
        for k, r in enumerate(program.radix):
            st.radix[k] = r
        for i, lv in enumerate(tree.levels):
            L = st.level[i]
            L.seg_begin, L.seg_end = lv.seg
            L.kind, L.qubit, L.digit, L.pre_off, L.post_off = lv.kind, lv.qubit, lv.digit, lv.pre_off, lv.post_off
            L.n_choices, L.col_bit = len(lv.choices), lv.col_bit
            L.meas_mask = sum(1 << v for v, m in enumerate(lv.meas) if m)
            L.canon = sum(c << (4 * v) for v, c in enumerate(lv.canon))
            for v in range(8):
                L.first_choice[v] = -1
            for c, (v, o) in enumerate(lv.choices):
                L.choice_variant[c], L.choice_outcome[c] = v, o
                if L.first_choice[v] < 0:
                    L.first_choice[v] = c
        return blob, off_ops, st

    @staticmethod
    def _native_host_image(program: FragmentProgram, fold: bool):
        """``_build_host_image`` in C++: the blob and the ``qck_sim_plan`` structs come filled (their ``sweeps``
        pointers address memory owned by the program's native object, which the program keeps alive)."""
        program._native_build(2, fold)
        lib, hp, base = program._hp_lib, program._hp, 100 + (20 if fold else 0)
        off_ops, off_labels, off_extra, off_src, dedupe, n_plans, _nb = _host_get(lib, hp, base + 5, np.int64).tolist()
        blob = _host_get(lib, hp, base + 6, np.uint8)
        raw = _host_get(lib, hp, base + 7, np.uint8).tobytes()
        size = C.sizeof(_lib.QckSimPlan)
        lab = _host_get(lib, hp, base + 9, np.int64).reshape(-1, 2).tolist()
        structs = [(_lib.QckSimPlan.from_buffer_copy(raw, i * size), lab[i][0], lab[i][1]) for i in range(n_plans)]
        labels = blob[off_labels:off_labels + 4 * (lab[-1][0] + lab[-1][1] if lab else 0)].view(np.int32)
        dedupe_info = None
        if dedupe:
            reps = _host_get(lib, hp, base + 8, np.int64).reshape(-1, 2).tolist()
            dedupe_info = (off_extra, [tuple(r) for r in reps], off_src)
        return blob, off_ops, off_labels, labels, None, structs, dedupe_info

    @staticmethod
    def _build_host_image(program: FragmentProgram, plans: list):
        ops = np.concatenate([p.ops for p in plans]) if plans else np.zeros((0, 8), np.int32)
        if len(ops) == 0:
            ops = np.zeros((1, 8), np.int32)
        labels = np.concatenate([p.labels for p in plans]).astype(np.int32)
        # representatives only (per plan, same order) + the source row of every label: used when the whole
        # label range is simulated and some instances are identical (qck_rows_broadcast)
        src = program.canonical_labels()
        saved = int((src != np.arange(len(src))).sum())
        dedupe = saved > 0 and (DEDUPE is True or (DEDUPE == "auto" and saved >= DEDUPE_MIN_SAVED))
        reps = [p.labels[src[p.labels] == p.labels] for p in plans] if dedupe else []
        extra = np.concatenate(reps + [src]).astype(np.int32) if dedupe else np.zeros(0, np.int32)
        # one host blob [mats f64 | ops i32 | labels i32 | representatives i32 | sources i32] -> one H2D copy
        mats = np.ascontiguousarray(program.mats, dtype=np.float64)
        off_ops = (mats.nbytes + 255) & ~255
        off_labels = (off_ops + ops.nbytes + 255) & ~255
        off_extra = (off_labels + labels.nbytes + 255) & ~255
        blob = np.zeros(off_extra + extra.nbytes, dtype=np.uint8)
        blob[:mats.nbytes] = mats.view(np.uint8)
        blob[off_ops:off_ops + ops.nbytes] = np.ascontiguousarray(ops).view(np.uint8).reshape(-1)
        blob[off_labels:off_labels + labels.nbytes] = labels.view(np.uint8)
        blob[off_extra:] = extra.view(np.uint8)
        rep_ranges, w = [], 0
        for r in reps:
            rep_ranges.append((w, len(r)))
            w += len(r)
        dedupe_info = (off_extra, rep_ranges, off_extra + 4 * w) if dedupe else None
        sweep_arrays, structs = [], []
        off = 0
        for p in plans:
            arr = (_lib.QckSweep * len(p.sweeps))()
            for i, (positions, b, e) in enumerate(p.sweeps):
                if p.warp_base:
                    positions = positions[:p.warp_base]
                arr[i].n_tile = len(positions)
                arr[i].op_begin = p.op_base + b
                arr[i].op_end = p.op_base + e
                arr[i].flags = (int(np.isin(p.ops[b:e, 0], (_lib.OP_U1X, _lib.OP_PHASE)).any())
                                | 2 * int((p.ops[b:e, 0] == _lib.OP_CLUSTER).any())
                                | (_lib.SWEEP_SHARED if p.shared_prefix and i == 0 else 0)
                                | ((_lib.SWEEP_WARP | (p.warp_base << 8)) if p.warp_base else 0))
                for j, x in enumerate(positions):
                    arr[i].pos[j] = x
            sweep_arrays.append(arr)
            st = _lib.QckSimPlan()
            st.n_state_qubits = p.n_state
            st.n_sweeps = len(p.sweeps)
            st.sweeps = arr
            st.n_digits = len(program.radix)
            for k, r in enumerate(program.radix):
                st.radix[k] = r
            st.n_out_bits = len(p.out_pos)
            for j, x in enumerate(p.out_pos):
                st.out_pos[j] = x
            st.sum_mask = p.sum_mask
            st.sign_mask = p.sign_mask
            structs.append((st, off, len(p.labels)))
            off += len(p.labels)
        return blob, off_ops, off_labels, labels, sweep_arrays, structs, dedupe_info

    def upload(self) -> None:
        """Host -> device copy of matrices, ops and label lists through pinned memory (one copy;
        part of the e2e timed region)."""
        torch = self.torch
        stage = _pinned_ring().take(self._blob.nbytes)
        stage[:self._blob.nbytes].numpy()[:] = self._blob
        self.d_blob = torch.empty(self._blob.nbytes, dtype=torch.uint8, device=self.device)
        self.d_blob.copy_(stage[:self._blob.nbytes], non_blocking=True)

    def alloc_out(self, label_range=None):
        """The table a run fills: [num_labels, row_len] float64 (zero-filled when only a label range is written)."""
        alloc = self.torch.zeros if label_range is not None else self.torch.empty
        return alloc((self.program.num_labels, self.row_len), dtype=self.torch.float64, device=self.device)

    def plan_struct(self, i: int = 0) -> "_lib.QckSimPlan":
        """A private copy of plan i's ``qck_sim_plan`` with the device pointers of THIS executor filled in
        (for direct C-ABI calls such as ``qck_sim_statevector``); uploads the program if necessary."""
        if self.d_blob is None:
            self.upload()
        st = _lib.QckSimPlan.from_buffer_copy(self._structs[i][0])
        st.d_ops = self.d_blob.data_ptr() + self._off_ops
        st.d_mats = self.d_blob.data_ptr()
        return st

    def run(self, handle: "_lib.Handle", out=None, label_range: tuple[int, int] | None = None,
            scratch_tag=0, defer_broadcast: bool = False):
        """-> device tensor [num_labels, row_len] float64 (rows outside label_range untouched / 0).
        ``scratch_tag``: executors whose runs overlap on the device (one ``qck_sim_region``) pass different tags
        and get different work buffers; ``defer_broadcast``: inside a region the rows of identical instances
        can only be copied after the region has been joined - the caller then calls ``finish(handle, out)``."""
        torch = self.torch
        if self.d_blob is None:
            self.upload()
        prog = self.program
        if out is None:
            out = self.alloc_out(label_range)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.tree is not None:
            st = _lib.QckSimTreePlan.from_buffer_copy(self._tree_struct)
            st.d_ops = self.d_blob.data_ptr() + self._off_ops
            st.d_mats = self.d_blob.data_ptr()
            if self._work_bytes is None:
                self._work_bytes = int(handle.lib.qck_sim_tree_work_bytes(C.byref(st)))
            work = handle.scratch(torch, self._work_bytes, self.device, stream, scratch_tag)
            l0, l1 = label_range if label_range is not None else (0, prog.num_labels)
            handle.check(handle.lib.qck_sim_tree(handle.ptr, C.byref(st), l0, l1, out.data_ptr(), self.row_len,
                                                 work.data_ptr(), work.numel(), stream))
            self._pending_broadcast = False
            return out
        # Scratch comes from the handle's cache, keyed by (device, stream): it is fetched on EVERY run for the
        # stream this run is enqueued on (work on one stream is ordered; another stream gets its own buffer).
        # Only the size is remembered: the shared-prefix snapshots, or as many streaming states as fit.
        if self._work_bytes is None:
            if not self.streaming:
                self._work_bytes = sum(16 << st.n_state_qubits for st, _o, _c in self._structs
                                       if st.sweeps[0].flags & _lib.SWEEP_SHARED)
            else:
                per = 16 << self.max_state
                n = 1
                if prog.num_labels > 1:          # several instances in flight: bounded by free memory
                    free, _ = torch.cuda.mem_get_info(self.device)
                    n = max(1, min(prog.num_labels, int(free * 0.5) // per, 64))
                self._work_bytes = n * per
        if not self._work_bytes:
            self._work = None
        elif self._work_bytes > handle.SCRATCH_CACHE_MAX:
            # too large for the handle's cache (a 64 GiB state): this executor owns it, per stream
            if self._work is None or self._work_key != (stream, scratch_tag):
                self._work = None                      # (free the old one first)
                self._work = handle.scratch(torch, self._work_bytes, self.device, stream, scratch_tag)
                self._work_key = (stream, scratch_tag)
        else:
            self._work = handle.scratch(torch, self._work_bytes, self.device, stream, scratch_tag)
        n = len(self._structs)
        plans = (_lib.QckSimPlan * n)()
        label_ptrs = (C.c_void_p * n)()
        counts = (C.c_int64 * n)()
        ops_ptr = self.d_blob.data_ptr() + self._off_ops
        mats_ptr = self.d_blob.data_ptr()
        dedupe = self._dedupe if label_range is None else None      # a clipped range may miss its representatives
        for i, (st, off, count) in enumerate(self._structs):
            labels_ptr = self.d_blob.data_ptr() + self._off_labels + 4 * off
            if dedupe is not None:
                rep_off, count = dedupe[1][i]
                labels_ptr = self.d_blob.data_ptr() + dedupe[0] + 4 * rep_off
            if label_range is not None:
                # label lists are ascending inside a plan: clip to the requested range
                host = self._labels_host[off:off + count]
                lo = int(np.searchsorted(host, label_range[0], side="left"))
                hi = int(np.searchsorted(host, label_range[1], side="left"))
                labels_ptr += 4 * lo
                count = max(0, hi - lo)
            plans[i] = st                    # a copy of the shared template; pointers go into the copy
            plans[i].d_ops = ops_ptr
            plans[i].d_mats = mats_ptr
            label_ptrs[i] = labels_ptr
            counts[i] = count
        work_ptr = self._work.data_ptr() if self._work is not None else None
        work_bytes = self._work.numel() if self._work is not None else 0
        handle.check(handle.lib.qck_sim_fragments_batch(handle.ptr, n, plans, label_ptrs, counts, out.data_ptr(),
                                                        self.row_len, work_ptr, work_bytes, stream))
        self._pending_broadcast = dedupe is not None
        if not defer_broadcast:
            self.finish(handle, out)
        return out

    def finish(self, handle: "_lib.Handle", out) -> None:
        """Rows of instances identical to their representative (qck_rows_broadcast), on the current stream."""
        if getattr(self, "_pending_broadcast", False):
            stream = self.torch.cuda.current_stream(self.device).cuda_stream
            handle.check(handle.lib.qck_rows_broadcast(handle.ptr, out.data_ptr(), self.row_len, self.row_len,
                                                       self.d_blob.data_ptr() + self._dedupe[2],
                                                       self.program.num_labels, stream))
            self._pending_broadcast = False
