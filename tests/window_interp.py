"""CPU interpreter of the device programs the fused window executor launches (TEST INFRASTRUCTURE).

`qi_debug_lower` (host-only entry of the C ABI) serialises what `qi_apply_circuit` would launch for a gate list:
window passes (their window qubits, op lists and phase tables) and per-gate-kernel steps.  This module states what
every device op MEANS, independently of the CUDA code (quant_iron_b200/csrc/window.cu, run_ops), as whole-vector
numpy operations, so that the host half of the executor -- scheduling, CNOT absorption, diagonal merging, phase
tables, slot masks, tile predicates -- is checked against the oracle on machines without a GPU.
It is not a fallback: nothing in the product imports it.
"""
import ctypes as C
import struct

import numpy as np

WK_X, WK_RX, WK_RXS, WK_REAL, WK_U2, WK_REALUP, WK_REALUM, WK_RXU, WK_RXSU, WK_DIAG, WK_RZ, WK_TABLE, WK_SCALE, WK_NEG, WK_REALL, WK_RXL = range(1, 17)
PAIR_KINDS = set(range(WK_X, WK_RXSU + 1)) | {WK_REALL, WK_RXL}
LAST_PAIR = WK_RXSU
IK_H, IK_X, IK_Y, IK_U2, IK_DIAG, IK_RZ, IK_SWAP, IK_MATCH = 1, 2, 3, 4, 5, 6, 7, 8
CLS_LANE, CLS_REG, CLS_TILE = 1, 2, 3

DOP = np.dtype([("kind", "u1"), ("tpos", "u1"), ("hub_cls", "u1"), ("hub_bit", "u1"), ("nchunks", "u1"), ("has_reg", "u1"),
                ("c_lval", "u1"), ("code", "u1"), ("c_lane", "<u4"), ("c_reg", "<u4"), ("c_tile", "<u8"), ("c_tval", "<u8"),
                ("t_lane", "<u4"), ("t_reg", "<u4"), ("t_tile", "<u8"), ("m", "<f8", 8)])
PHYS = np.dtype([("kind", "<i4"), ("t0", "<i4"), ("t1", "<i4"), ("pad", "<i4"), ("cmask", "<u8"), ("p", "<f8", 8)])
assert DOP.itemsize == 112 and PHYS.itemsize == 88


def lower(circuit, n, rank=0, world=1, regs=0, phys=None):
    """Serialised programs for a Circuit of this package (one run of operator gates); `phys` = the logical -> physical
    qubit map the run starts under (64 entries, default identity)."""
    from quant_iron_b200 import _ffi
    runs = circuit._lower()
    assert len(runs) == 1 and runs[0][0] == "ops"
    used = C.c_uint64()
    cap = 1 << 20
    pmap = (C.c_uint8 * 64)(*phys) if phys is not None else None
    while True:
        blob = (C.c_uint8 * cap)()
        st = _ffi.lib.qi_debug_lower(n, rank, world, pmap, runs[0][1], runs[0][2], regs, blob, cap, C.byref(used))
        if st == 0:
            return bytes(blob[:used.value])
        if used.value > cap:
            cap = used.value
            continue
        _ffi.check(st)


def parse(blob):
    off = 0

    def u64():
        nonlocal off
        v = struct.unpack_from("<Q", blob, off)[0]
        off += 8
        return v

    steps = []
    for _ in range(u64()):
        tag = u64()
        if tag == 1:
            steps.append(("simple", np.frombuffer(blob, PHYS, 1, off)[0]))
            off += PHYS.itemsize
        elif tag == 2:                # one k_tile launch: 11 tile qubits, rounds of (4 register qubits, 7 thread-bit qubits, ops)
            tile_qubits = [u64() for _ in range(11)]
            tile_out = [u64() for _ in range(11)]
            rounds = []
            for _ in range(u64()):
                regs = [u64() for _ in range(4)]
                thr = [u64() for _ in range(7)]
                nops = u64()
                ops = np.frombuffer(blob, DOP, nops, off)
                off += nops * DOP.itemsize
                rounds.append((regs, thr, ops))
            steps.append(("tile", tile_qubits, rounds, tile_out))
        else:
            R = u64()
            regs = [u64() for _ in range(8)][:R]
            nops = u64()
            ops = np.frombuffer(blob, DOP, nops, off)
            off += nops * DOP.itemsize
            steps.append(("pass", R, regs, ops))
    na = u64()
    arena = np.frombuffer(blob, "<c16", na, off)
    off += 16 * na
    phys = list(blob[off:off + 64])
    assert off + 64 == len(blob)
    return steps, arena, phys


def _pair(v, idx0, tb, kind, m):
    """2x2 update of the pairs (i, i | tb) for i in idx0 (target bit clear)."""
    a0, a1 = v[idx0], v[idx0 | tb]
    if kind == WK_X:
        n0, n1 = a1, a0
    elif kind == WK_RX:
        n0, n1 = m[0] * a0 - 1j * m[1] * a1, m[0] * a1 - 1j * m[1] * a0
    elif kind == WK_RXS:
        n0, n1 = m[0] * a1 - 1j * m[1] * a0, m[0] * a0 - 1j * m[1] * a1
    elif kind == WK_REAL:
        n0, n1 = m[0] * a0 + m[1] * a1, m[2] * a0 + m[3] * a1
    elif kind == WK_U2:
        m00, m01, m10, m11 = (complex(m[0], m[1]), complex(m[2], m[3]), complex(m[4], m[5]), complex(m[6], m[7]))
        n0, n1 = m00 * a0 + m01 * a1, m10 * a0 + m11 * a1
    elif kind in (WK_REALUP, WK_REALUM):      # lean forms (option "lean"): [[1, p], [q, +-1]]
        n0, n1 = a0 + m[0] * a1, m[1] * a0 + (a1 if kind == WK_REALUP else -a1)
    elif kind == WK_RXU:                      # [[1, -i t], [-i t, 1]]
        n0, n1 = a0 - 1j * m[0] * a1, a1 - 1j * m[0] * a0
    elif kind == WK_RXSU:                     # the same with its inputs swapped
        n0, n1 = a1 - 1j * m[0] * a0, a0 - 1j * m[0] * a1
    elif kind == WK_REALL:                    # k_tile, lifted real 2x2: the second row acts on the NEW a0 (m = k00, k01, k10/k00, det/k00)
        n0 = m[0] * a0 + m[1] * a1
        n1 = m[3] * a1 + m[2] * n0
    elif kind == WK_RXL:                      # k_tile, lifted RX (m = c, s, 1/c, s/c)
        n0 = m[0] * a0 - 1j * m[1] * a1
        n1 = m[2] * a1 - 1j * m[3] * n0
    else:
        raise AssertionError(f"unknown pair kind {kind}")
    v[idx0] = n0
    v[idx0 | tb] = n1


def run_pass(v, nl, R, regs, ops, arena, lane_qubits=(0, 1, 2, 3, 4)):
    """One op list under one layout: `regs` = physical qubits of the slot bits, `lane_qubits` = physical qubits of the
    lane-index bits (k_window: qubits 0..4; a k_tile round: the 7 thread-index bits); every other qubit is a tile bit."""
    S = 1 << R
    NT = 1 << len(lane_qubits)
    idx = np.arange(1 << nl, dtype=np.uint64)
    lane = np.zeros(1 << nl, dtype=np.uint32)
    for j, q in enumerate(lane_qubits):
        lane |= (((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(j)).astype(np.uint32)
    slot = np.zeros(1 << nl, dtype=np.uint32)
    for j, q in enumerate(regs):
        slot |= (((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(j)).astype(np.uint32)
    tile = np.zeros(1 << nl, dtype=np.uint64)
    t = 0
    for q in range(nl):
        if q in regs or q in lane_qubits:
            continue
        tile |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(t)
        t += 1
    for op in ops:
        kind = int(op["kind"])
        m = op["m"]
        tile_ok = (tile & np.uint64(op["c_tile"])) == np.uint64(op["c_tval"])
        lane_ok = (lane & np.uint32(op["c_lane"])) == np.uint32(op["c_lval"])
        c_reg = int(op["c_reg"])
        slot_ok = ((np.uint32(c_reg) >> slot) & np.uint32(1)).astype(bool)
        if kind in PAIR_KINDS:
            tpos = int(op["tpos"])
            tbit = lane_qubits[tpos] if tpos < 5 else regs[tpos - 5]
            assert tpos >= 5 or NT == 32, "a k_tile round has no lane-pair ops"
            tb = np.uint64(1 << tbit)
            # the predicate is evaluated on the member with the target bit clear; controls never include the target
            sel = tile_ok & lane_ok & slot_ok & ((idx & tb) == 0)
            _pair(v, idx[sel], tb, kind, m)
            continue
        active = tile_ok & lane_ok
        if kind == WK_SCALE:                  # product of the deferred scalars of the launch: unit-form gates (real) and P-form diagonals (complex)
            v[active & slot_ok] *= complex(m[0], m[1])
        elif kind == WK_DIAG:
            v[active & slot_ok] *= complex(m[0], m[1])
        elif kind == WK_NEG:                    # a phase of exactly -1 (Z, CZ): sign flips
            v[active & slot_ok] *= -1.0
        elif kind == WK_RZ:
            p0, p1 = complex(m[0], m[1]), complex(m[2], m[3])
            t_thread = ((tile & np.uint64(op["t_tile"])) != 0) | ((lane & np.uint32(op["t_lane"])) != 0)
            t_slot = (slot & np.uint32(op["t_reg"])) != 0
            ph = np.where(t_slot | t_thread, p1, p0)
            sel = active & slot_ok
            v[sel] *= ph[sel]
        elif kind == WK_TABLE:
            base = int(np.frombuffer(np.float64(m[0]).tobytes(), "<i8")[0])
            hub_cls, hub_bit = int(op["hub_cls"]), int(op["hub_bit"])
            if hub_cls == CLS_TILE:
                active = active & (((tile >> np.uint64(hub_bit)) & np.uint64(1)) == 1)
            elif hub_cls == CLS_LANE:
                active = active & (((lane >> np.uint32(hub_bit)) & np.uint32(1)) == 1)
            elif hub_cls == CLS_REG:
                active = active & (((slot >> np.uint32(hub_bit)) & np.uint32(1)) == 1)
            f = arena[base + lane.astype(np.int64)]
            for k in range(int(op["nchunks"])):
                f = f * arena[base + NT + S + 256 * k + ((tile >> np.uint64(8 * k)) & np.uint64(255)).astype(np.int64)]
            if op["has_reg"]:
                f = f * arena[base + NT + slot.astype(np.int64)]
            v[active] *= f[active]
        else:
            raise AssertionError(f"unknown device op kind {kind}")


def run_simple(v, nl, g):
    idx = np.arange(1 << nl, dtype=np.uint64)
    cm = np.uint64(g["cmask"])
    ctrl = (idx & cm) == cm
    kind, t0, t1, p = int(g["kind"]), int(g["t0"]), int(g["t1"]), g["p"]
    if kind in (IK_H, IK_X, IK_Y, IK_U2):
        tb = np.uint64(1 << t0)
        i0 = idx[ctrl & ((idx & tb) == 0)]
        a0, a1 = v[i0], v[i0 | tb]
        if kind == IK_H:
            v[i0], v[i0 | tb] = p[0] * (a0 + a1), p[0] * (a0 - a1)
        elif kind == IK_X:
            v[i0], v[i0 | tb] = a1, a0
        elif kind == IK_Y:
            v[i0], v[i0 | tb] = -1j * a1, 1j * a0
        else:
            _pair(v, i0, tb, WK_U2, p)
    elif kind == IK_DIAG:
        sel = ctrl if t0 < 0 else ctrl & ((idx & np.uint64(1 << t0)) != 0)
        v[sel] *= complex(p[0], p[1])
    elif kind == IK_RZ:
        hi = (idx & np.uint64(1 << t0)) != 0 if t0 >= 0 else np.zeros(1 << nl, bool)
        ph = np.where(hi, complex(p[2], p[3]), complex(p[0], p[1]))
        v[ctrl] *= ph[ctrl]
    elif kind == IK_SWAP:
        ba, bb = np.uint64(1 << t0), np.uint64(1 << t1)
        i = idx[ctrl & ((idx & ba) != 0) & ((idx & bb) == 0)]
        j = i ^ (ba | bb)
        v[i], v[j] = v[j].copy(), v[i].copy()
    elif kind == IK_MATCH:
        b1, b2 = np.uint64(1 << t0), np.uint64(1 << t1)
        l = idx[ctrl & ((idx & (b1 | b2)) == 0)]
        a01, a10, a11 = v[l | b1], v[l | b2], v[l | b1 | b2]
        e1, e2 = complex(p[2], p[3]), complex(p[4], p[5])
        v[l | b1] = p[0] * a01 - e1 * p[1] * a10
        v[l | b2] = p[1] * a01 + e1 * p[0] * a10
        v[l | b1 | b2] = a11 * e2
    else:
        raise AssertionError(f"unknown physical gate kind {kind}")


def execute(blob, v, nl):
    """Run the serialised programs on the local vector `v` (length 2^nl, PHYSICAL order) in place; returns the final
    logical -> physical qubit map."""
    steps, arena, phys = parse(blob)
    for st in steps:
        if st[0] == "simple":
            run_simple(v, nl, st[1])
        elif st[0] == "tile":
            tile_qubits, rounds, tile_out = st[1], st[2], st[3]
            assert tile_qubits[:5] == [0, 1, 2, 3, 4] and sorted(tile_qubits) == tile_qubits and len(set(tile_qubits)) == 11
            assert sorted(tile_out) == tile_qubits, "the store layout permutes the tile's own positions"
            for k, (regs, thr, ops) in enumerate(rounds):
                assert sorted(regs + thr) == tile_qubits, "a round's register + thread qubits are the tile qubits"
                if k == 0:                              # load layout: coalesced 512-byte rows
                    assert thr[:5] == [0, 1, 2, 3, 4], "first round: lanes on positions 0..4"
                if k == len(rounds) - 1:                # store layout: the lanes hold what goes to positions 0..4
                    dest = dict(zip(tile_qubits, tile_out))
                    assert [dest[q] for q in thr[:5]] == [0, 1, 2, 3, 4], "last round: lanes on the bits stored to positions 0..4"
                run_pass(v, nl, 4, regs, ops, arena, lane_qubits=tuple(thr))
            if tile_out != tile_qubits:                 # the pass stores its tile with the local bits in a new order
                idx = np.arange(1 << nl, dtype=np.uint64)
                tmask = np.uint64(sum(1 << p for p in tile_qubits))
                new = idx & ~tmask
                for pin, pout in zip(tile_qubits, tile_out):
                    new |= ((idx >> np.uint64(pin)) & np.uint64(1)) << np.uint64(pout)
                w = np.empty_like(v)
                w[new] = v
                v[:] = w
        else:
            run_pass(v, nl, st[1], st[2], st[3], arena)
    return phys, steps


def to_logical(v, n, phys):
    """Undo a (lazy SWAP) relabelling: logical index L lives at the physical index with bit q of L at position phys[q]."""
    if all(phys[q] == q for q in range(n)):
        return v
    L = np.arange(1 << n, dtype=np.uint64)
    P = np.zeros(1 << n, dtype=np.uint64)
    for q in range(n):
        P |= ((L >> np.uint64(q)) & np.uint64(1)) << np.uint64(phys[q])
    return v[P]


# ---- fused Pauli-exponential sequences (csrc/pauli_window.cu) ---------------------------------------------------
PXOP = np.dtype([("xl", "<u4"), ("xr", "<u4"), ("zl", "<u4"), ("zr", "<u4"), ("zt", "<u8"), ("k0", "<u4"), ("pad", "<u4"),
                 ("c", "<f8"), ("s", "<f8")])
PEXP = np.dtype([("x", "<u8"), ("z", "<u8"), ("k0", "<i4"), ("pad", "u1", 12), ("ch", "<f8", 2), ("sh", "<f8", 2)])
assert PXOP.itemsize == 48 and PEXP.itemsize == 64


def lower_pauli(strings, factors, n, rank=0, world=1, phys=None):
    """Serialised programs of qi_apply_pauli_exp_sequence for PauliStrings of this package (on shard `rank` of `world`
    under the qubit map `phys` when given)."""
    from quant_iron_b200 import _ffi
    arr = (_ffi.QiPauliTerm * len(strings))()
    keep = []
    for i, ps in enumerate(strings):
        rec, k = ps.term()
        arr[i] = rec
        keep.append(k)
    flat = []
    for f in factors:
        flat += [complex(f).real, complex(f).imag]
    used = C.c_uint64()
    cap = 1 << 20
    while True:
        blob = (C.c_uint8 * cap)()
        pmap = (C.c_uint8 * 64)(*phys) if phys is not None else None
        st = _ffi.lib.qi_debug_pauli_lower(n, rank, world, pmap, arr, len(strings), _ffi.dbl_array(flat), blob, cap, C.byref(used))
        if st == 0:
            return bytes(blob[:used.value])
        if used.value > cap:
            cap = used.value
            continue
        _ffi.check(st)


def _parity(x):
    x = x.copy()
    for sh in (32, 16, 8, 4, 2, 1):
        x ^= x >> np.uint64(sh)
    return (x & np.uint64(1)).astype(np.int64)


_IPOW = np.array([1, 1j, -1, -1j], dtype=np.complex128)


def execute_pauli(blob, v, n):
    """psi <- cosh(a) psi + sinh(a) P psi per term, (P psi)[i] = i^(k0 + 2 popc(i & z)) psi[i ^ x], in place on `v`."""
    off = 0

    def u64():
        nonlocal off
        val = struct.unpack_from("<Q", blob, off)[0]
        off += 8
        return val

    idx = np.arange(1 << n, dtype=np.uint64)
    lane = idx & np.uint64(31)
    passes, singles = 0, 0
    for _ in range(u64()):
        if u64():
            e = np.frombuffer(blob, PEXP, 1, off)[0]
            off += PEXP.itemsize
            singles += 1
            k = (int(e["k0"]) + 2 * _parity(idx & np.uint64(e["z"]))) & 3
            ch, sh = complex(*e["ch"]), complex(*e["sh"])
            v[:] = ch * v + sh * (_IPOW[k] * v[idx ^ np.uint64(e["x"])])
            continue
        R = u64()
        regs = [u64() for _ in range(8)][:R]
        nops = u64()
        ops = np.frombuffer(blob, PXOP, nops, off)
        off += nops * PXOP.itemsize
        passes += 1
        slot = np.zeros(1 << n, dtype=np.uint64)
        for j, q in enumerate(regs):
            slot |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(j)
        tile = np.zeros(1 << n, dtype=np.uint64)
        t = 0
        for q in range(5, n):
            if q in regs:
                continue
            tile |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(t)
            t += 1
        for op in ops:
            x = int(op["xl"])
            for j, q in enumerate(regs):
                if (int(op["xr"]) >> j) & 1:
                    x |= 1 << q
            flip = _parity(tile & np.uint64(op["zt"])) + _parity(lane & np.uint64(op["zl"])) + _parity(slot & np.uint64(op["zr"]))
            k = (int(op["k0"]) + 2 * flip) & 3
            v[:] = float(op["c"]) * v + float(op["s"]) * (_IPOW[k] * v[idx ^ np.uint64(x)])
    assert off == len(blob)
    return passes, singles


# ---- batched SumOp expectation (csrc/pauli_window.cu, k_pauli_expect_window) ---------------------------------------
def lower_expect(strings, n):
    from quant_iron_b200 import _ffi
    arr = (_ffi.QiPauliTerm * len(strings))()
    keep = []
    for i, ps in enumerate(strings):
        rec, k = ps.term()
        arr[i] = rec
        keep.append(k)
    used = C.c_uint64()
    cap = 1 << 20
    while True:
        blob = (C.c_uint8 * cap)()
        st = _ffi.lib.qi_debug_expect_lower(n, arr, len(strings), blob, cap, C.byref(used))
        if st == 0:
            return bytes(blob[:used.value])
        if used.value > cap:
            cap = used.value
            continue
        _ffi.check(st)


def expect_groups(blob, v, n):
    """sum over the grouped terms of <psi| c P |psi> = sum_i conj(psi[i]) c i^(k0 + 2 popc(i & z)) psi[i ^ x]; returns
    (value, number of groups, indices of the terms left to the per-term kernel)."""
    off = 0

    def u64():
        nonlocal off
        val = struct.unpack_from("<Q", blob, off)[0]
        off += 8
        return val

    idx = np.arange(1 << n, dtype=np.uint64)
    lane = idx & np.uint64(31)
    total = 0j
    ngroups = u64()
    for _ in range(ngroups):
        R = u64()
        regs = [u64() for _ in range(8)][:R]
        nops = u64()
        ops = np.frombuffer(blob, PXOP, nops, off)
        off += nops * PXOP.itemsize
        slot = np.zeros(1 << n, dtype=np.uint64)
        for j, q in enumerate(regs):
            slot |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(j)
        tile = np.zeros(1 << n, dtype=np.uint64)
        t = 0
        for q in range(5, n):
            if q in regs:
                continue
            tile |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(t)
            t += 1
        for op in ops:
            x = int(op["xl"])
            for j, q in enumerate(regs):
                if (int(op["xr"]) >> j) & 1:
                    x |= 1 << q
            flip = _parity(tile & np.uint64(op["zt"])) + _parity(lane & np.uint64(op["zl"])) + _parity(slot & np.uint64(op["zr"]))
            k = (int(op["k0"]) + 2 * flip) & 3
            total += complex(op["c"], op["s"]) * np.vdot(v, _IPOW[k] * v[idx ^ np.uint64(x)])
    left = [u64() for _ in range(u64())]
    assert off == len(blob)
    return total, ngroups, left
