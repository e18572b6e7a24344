// gates.cu -- Operator::apply for every reference operator: validation (operator.rs:214-273),
// parameter resolution, and the per-gate ("simple") kernels: one HBM pass per gate, visiting only
// the amplitudes the gate can change (controls and the phase-gate target are folded into the index
// expansion, so a CNOT touches 1/2 and a CP 1/4 of the state).
// The fused register-window executor lives in window.cu; these kernels serve small states
// (n < 9), gates the window kernel does not take, and the "path"=1 option.
#include <algorithm>

#include "common.cuh"

namespace qi {

static const int kBlock = 256;

// ---- kernels ------------------------------------------------------------------------------------
struct U2 { amp_t m00, m01, m10, m11; };

// Each thread handles kUnroll independent items (loads first, then arithmetic, then stores) so that
// 2 * kUnroll 16-byte requests are in flight per thread.
static const int kUnroll = 4;

template <int IK>
__global__ void __launch_bounds__(256) k_pair(amp_t* __restrict__ a, uint64_t total, BitInsert ins, uint64_t tbit, U2 u) {
    const uint64_t quarter = (total + kUnroll - 1) / kUnroll;
    const uint64_t k0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k0 >= quarter) return;
    uint64_t i0[kUnroll];
    amp_t a0[kUnroll], a1[kUnroll];
    bool on[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const uint64_t k = k0 + (uint64_t)j * quarter;
        on[j] = k < total;
        i0[j] = expand_index(on[j] ? k : 0, ins);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; j++) if (on[j]) { a0[j] = ld_amp(a + i0[j]); a1[j] = ld_amp(a + (i0[j] | tbit)); }
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        if (!on[j]) continue;
        amp_t r0, r1;
        if (IK == IK_H) {            // operator.rs:401-402: s*(a0+a1), s*(a0-a1)
            const double s = u.m00.x;
            r0 = cscale(s, cadd(a0[j], a1[j]));
            r1 = cscale(s, csub(a0[j], a1[j]));
        } else if (IK == IK_X) {     // operator.rs:579-580
            r0 = a1[j]; r1 = a0[j];
        } else if (IK == IK_Y) {     // operator.rs:588-589: -i*a1, i*a0
            r0 = make_double2(a1[j].y, -a1[j].x);
            r1 = make_double2(-a0[j].y, a0[j].x);
        } else {                     // operator.rs:2246-2247
            r0 = cadd(cmul(u.m00, a0[j]), cmul(u.m01, a1[j]));
            r1 = cadd(cmul(u.m10, a0[j]), cmul(u.m11, a1[j]));
        }
        st_amp(a + i0[j], r0);
        st_amp(a + (i0[j] | tbit), r1);
    }
}

// amp *= phase on every index the expansion produces (target bit and controls forced to 1)
__global__ void __launch_bounds__(256) k_diag(amp_t* __restrict__ a, uint64_t total, BitInsert ins, amp_t phase) {
    const uint64_t quarter = (total + kUnroll - 1) / kUnroll;
    const uint64_t k0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k0 >= quarter) return;
    uint64_t idx[kUnroll];
    amp_t v[kUnroll];
    bool on[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const uint64_t k = k0 + (uint64_t)j * quarter;
        on[j] = k < total;
        idx[j] = expand_index(on[j] ? k : 0, ins);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; j++) if (on[j]) v[j] = ld_amp(a + idx[j]);
#pragma unroll
    for (int j = 0; j < kUnroll; j++) if (on[j]) st_amp(a + idx[j], cmul(v[j], phase));
}

// RZ (operator.rs:2013-2029): every index with the controls set; phase by the target bit
__global__ void __launch_bounds__(256) k_rz(amp_t* __restrict__ a, uint64_t total, BitInsert ins, uint64_t tbit, amp_t p0, amp_t p1) {
    const uint64_t quarter = (total + kUnroll - 1) / kUnroll;
    const uint64_t k0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k0 >= quarter) return;
    uint64_t idx[kUnroll];
    amp_t v[kUnroll];
    bool on[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const uint64_t k = k0 + (uint64_t)j * quarter;
        on[j] = k < total;
        idx[j] = expand_index(on[j] ? k : 0, ins);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; j++) if (on[j]) v[j] = ld_amp(a + idx[j]);
#pragma unroll
    for (int j = 0; j < kUnroll; j++) if (on[j]) st_amp(a + idx[j], cmul(v[j], (idx[j] & tbit) ? p1 : p0));
}

// SWAP (operator.rs:800-813): expansion fixes t_a = 1, t_b = 0; partner flips both
__global__ void __launch_bounds__(256) k_swap(amp_t* __restrict__ a, uint64_t total, BitInsert ins, uint64_t flip) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    uint64_t i = expand_index(k, ins);
    uint64_t j = i ^ flip;
    amp_t x = ld_amp(a + i), y = ld_amp(a + j);
    st_amp(a + i, y);
    st_amp(a + j, x);
}

// Matchgate (operator.rs:987-1007) on (q1, q1+1)
__global__ void __launch_bounds__(256) k_match(amp_t* __restrict__ a, uint64_t total, BitInsert ins, uint64_t b1, uint64_t b2,
                                               double ch, double sh, amp_t e1, amp_t e2) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    uint64_t l = expand_index(k, ins);
    uint64_t i01 = l | b1, i10 = l | b2, i11 = l | b1 | b2;
    amp_t a01 = ld_amp(a + i01), a10 = ld_amp(a + i10), a11 = ld_amp(a + i11);
    amp_t e1s = make_double2(e1.x * sh, e1.y * sh), e1c = make_double2(e1.x * ch, e1.y * ch);
    st_amp(a + i01, csub(cscale(ch, a01), cmul(e1s, a10)));
    st_amp(a + i10, cadd(cscale(sh, a01), cmul(e1c, a10)));
    st_amp(a + i11, cmul(a11, e2));
}

BitInsert make_insert(const std::vector<int>& zero_positions, const std::vector<int>& one_positions) {
    BitInsert b;
    memset(&b, 0, sizeof(b));
    std::vector<int> all(zero_positions);
    all.insert(all.end(), one_positions.begin(), one_positions.end());
    std::sort(all.begin(), all.end());
    b.n = (int)all.size();
    for (int i = 0; i < b.n && i < 62; i++) b.pos[i] = (uint8_t)all[i];
    for (int p : one_positions) b.ones |= 1ull << p;
    return b;
}

static std::vector<int> bits_of(uint64_t m) {
    std::vector<int> v;
    for (int i = 0; i < 64; i++) if ((m >> i) & 1) v.push_back(i);
    return v;
}

// ---- validation (operator.rs:214-273 + per-operator extras) ------------------------------------
int validate_gate(const qi_state* s, const qi_gate* g) {
    if (!g) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "gate is NULL");
    if (g->kind < QI_GATE_H || g->kind > QI_GATE_MATCHGATE) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)g->kind, 0, "unknown gate kind");
    if (g->num_controls && !g->controls) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "controls is NULL");
    const uint32_t expected = (g->kind == QI_GATE_SWAP) ? 2u : 1u;
    const uint32_t n = s->num_qubits;
    if (g->num_targets != expected) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, g->num_targets, 0, "Invalid number of qubits");
    for (uint32_t i = 0; i < expected; i++)
        if (g->targets[i] >= n) return fail(QI_ERR_INVALID_QUBIT_INDEX, g->targets[i], n, "Invalid qubit index");
    for (uint32_t c = 0; c < g->num_controls; c++) {
        uint32_t cq = g->controls[c];
        if (cq >= n) return fail(QI_ERR_INVALID_QUBIT_INDEX, cq, n, "Invalid qubit index");
        for (uint32_t i = 0; i < expected; i++)
            if (cq == g->targets[i]) return fail(QI_ERR_OVERLAPPING_CONTROL_AND_TARGET, cq, g->targets[i], "Control qubit overlaps target");
    }
    if (expected > 1 && g->targets[0] == g->targets[1]) return fail(QI_ERR_INVALID_QUBIT_INDEX, g->targets[0], n, "duplicate target");
    switch (g->kind) {
        case QI_GATE_CNOT:     // operator.rs:677-679
            if (g->num_controls != 1) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, g->num_controls, 0, "CNOT needs one control");
            break;
        case QI_GATE_TOFFOLI:  // operator.rs:1065-1072
            if (g->num_controls != 2 || g->controls[0] == g->controls[1])
                return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, g->num_controls, 0, "Toffoli needs two distinct controls");
            break;
        case QI_GATE_MATCHGATE:  // operator.rs:903-905
            if (g->targets[0] == n - 1) return fail(QI_ERR_INVALID_QUBIT_INDEX, g->targets[0], n, "Matchgate target is the last qubit");
            break;
        default: break;
    }
    if (!s->consistent) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, s->num_qubits, 0, "state vector length is not 2^num_qubits");
    return QI_OK;
}

// A Matchgate whose control list names its own partner qubit t+1 passes the reference's validation (controls are only
// checked against the target, operator.rs:214-273) and means something definite there: check_controls(i01) fails for
// every pair (bit t+1 of i01 is 0), so the 01/10 rotation never runs, and only `amp11 * e^{i phi2}` is applied where the
// remaining controls are set (operator.rs:940-1007).  That is exactly PhaseShift(phi2) on t under the same control list
// (both evaluate (cos phi2, sin phi2) and one complex multiply), so such records are rewritten once at the ABI entry
// and nothing below ever sees a Matchgate whose control aliases its partner.
const qi_gate* normalise_gates(const qi_gate* gates, uint64_t count, std::vector<qi_gate>* own) {
    auto aliased = [](const qi_gate& g) {
        if (g.kind != QI_GATE_MATCHGATE) return false;
        for (uint32_t c = 0; c < g.num_controls; c++) if (g.controls[c] == g.targets[0] + 1) return true;
        return false;
    };
    uint64_t first = 0;
    while (first < count && !aliased(gates[first])) first++;
    if (first == count) return gates;
    own->assign(gates, gates + count);
    for (uint64_t i = first; i < count; i++) {
        if (!aliased(gates[i])) continue;
        qi_gate& g = (*own)[i];
        g.kind = QI_GATE_P;
        g.params[0] = gates[i].params[2];
    }
    return own->data();
}

// numeric parameters, computed on the host with the reference's expressions
static void resolve_params(const qi_gate* g, PhysGate* o) {
    const double is2 = 1.0 / std::sqrt(2.0);   // operator.rs:316, 1288
    memset(o->p, 0, sizeof(o->p));
    switch (g->kind) {
        case QI_GATE_H: o->kind = IK_H; o->p[0] = is2; break;
        case QI_GATE_X: case QI_GATE_CNOT: case QI_GATE_TOFFOLI: o->kind = IK_X; break;
        case QI_GATE_Y: o->kind = IK_Y; break;
        case QI_GATE_I: o->kind = IK_NOP; break;
        case QI_GATE_Z: o->kind = IK_DIAG; o->p[0] = -1.0; o->p[1] = 0.0; break;
        case QI_GATE_S: o->kind = IK_DIAG; o->p[0] = 0.0; o->p[1] = 1.0; break;
        case QI_GATE_SDG: o->kind = IK_DIAG; o->p[0] = 0.0; o->p[1] = -1.0; break;
        case QI_GATE_T: o->kind = IK_DIAG; o->p[0] = is2; o->p[1] = is2; break;
        case QI_GATE_TDG: o->kind = IK_DIAG; o->p[0] = is2; o->p[1] = -is2; break;
        case QI_GATE_P: o->kind = IK_DIAG; o->p[0] = std::cos(g->params[0]); o->p[1] = std::sin(g->params[0]); break;
        case QI_GATE_RX: {   // operator.rs:1729-1730 as a 2x2: [[c, -i s], [-i s, c]]
            double h = g->params[0] / 2.0, c = std::cos(h), s = std::sin(h);
            o->kind = IK_U2;
            o->p[0] = c; o->p[3] = -s; o->p[5] = -s; o->p[6] = c;
            break;
        }
        case QI_GATE_RY: {   // operator.rs:1871-1872: [[c, -s], [s, c]]
            double h = g->params[0] / 2.0, c = std::cos(h), s = std::sin(h);
            o->kind = IK_U2;
            o->p[0] = c; o->p[2] = -s; o->p[4] = s; o->p[6] = c;
            break;
        }
        case QI_GATE_RZ: {   // operator.rs:1997-1998
            double h = g->params[0] / 2.0;
            o->kind = IK_RZ;
            o->p[0] = std::cos(h); o->p[1] = -std::sin(h); o->p[2] = std::cos(h); o->p[3] = std::sin(h);
            break;
        }
        case QI_GATE_U2: o->kind = IK_U2; memcpy(o->p, g->params, 8 * sizeof(double)); break;
        case QI_GATE_SWAP: o->kind = IK_SWAP; break;
        case QI_GATE_MATCHGATE: {   // operator.rs:982-985
            o->kind = IK_MATCH;
            o->p[0] = std::cos(g->params[0] / 2.0);
            o->p[1] = std::sin(g->params[0] / 2.0);
            amp_t e1 = h_cexp(make_double2(0.0, g->params[1])), e2 = h_cexp(make_double2(0.0, g->params[2]));
            o->p[2] = e1.x; o->p[3] = e1.y; o->p[4] = e2.x; o->p[5] = e2.y;
            break;
        }
    }
}

// logical record -> physical record on this rank (handles the logical->physical qubit map and,
// for sharded states, controls/targets that live in the rank bits; see shard.cu)
int shard_prepare_gate(qi_state* s, const qi_gate* g, PhysGate* o, bool* skip);

int prepare_gate(qi_state* s, const qi_gate* g, PhysGate* o, bool* skip) {
    resolve_params(g, o);
    *skip = false;
    if (s->world > 1) return shard_prepare_gate(s, g, o, skip);
    o->t0 = s->phys[g->targets[0]];
    o->t1 = (g->kind == QI_GATE_SWAP) ? s->phys[g->targets[1]] : (g->kind == QI_GATE_MATCHGATE ? s->phys[g->targets[0] + 1] : -1);
    o->cmask = 0;
    for (uint32_t c = 0; c < g->num_controls; c++) o->cmask |= 1ull << s->phys[g->controls[c]];
    return QI_OK;
}

// ---- one pass with the per-gate kernels -----------------------------------------------------
int launch_simple_gate(qi_state* s, const PhysGate& g) {
    Context& c = ctx();
    const int n = (int)s->n_local;
    std::vector<int> ctrl = bits_of(g.cmask);
    const int nc = (int)ctrl.size();
    const double state_bytes = 16.0 * (double)s->len;
    auto blocks = [](uint64_t total) { return (unsigned)((total + kBlock - 1) / kBlock); };
    auto blocks4 = [](uint64_t total) { uint64_t q = (total + kUnroll - 1) / kUnroll; return (unsigned)((q + kBlock - 1) / kBlock); };
    switch (g.kind) {
        case IK_NOP: return QI_OK;
        case IK_H: case IK_X: case IK_Y: case IK_U2: {
            if (n - 1 - nc < 0) return QI_OK;
            uint64_t total = 1ull << (n - 1 - nc);
            BitInsert ins = make_insert({g.t0}, ctrl);
            U2 u;
            u.m00 = make_double2(g.p[0], g.p[1]); u.m01 = make_double2(g.p[2], g.p[3]);
            u.m10 = make_double2(g.p[4], g.p[5]); u.m11 = make_double2(g.p[6], g.p[7]);
            LaunchScope ls(KF_PAIR, 2.0 * state_bytes / (double)(1ull << nc));
            uint64_t tb = 1ull << g.t0;
            if (g.kind == IK_H) k_pair<IK_H><<<blocks4(total), kBlock, 0, c.stream>>>(s->d, total, ins, tb, u);
            else if (g.kind == IK_X) k_pair<IK_X><<<blocks4(total), kBlock, 0, c.stream>>>(s->d, total, ins, tb, u);
            else if (g.kind == IK_Y) k_pair<IK_Y><<<blocks4(total), kBlock, 0, c.stream>>>(s->d, total, ins, tb, u);
            else k_pair<IK_U2><<<blocks4(total), kBlock, 0, c.stream>>>(s->d, total, ins, tb, u);
            return check_launch("k_pair");
        }
        case IK_DIAG: {
            // t0 < 0: the target lives in the rank bits and is set on this rank -> phase on all
            // amplitudes whose (local) controls are set
            std::vector<int> ones(ctrl);
            if (g.t0 >= 0) ones.push_back(g.t0);
            int nf = (int)ones.size();
            if (n - nf < 0) return QI_OK;
            uint64_t total = 1ull << (n - nf);
            BitInsert ins = make_insert({}, ones);
            LaunchScope ls(KF_DIAG, 2.0 * state_bytes / (double)(1ull << nf));
            k_diag<<<blocks4(total), kBlock, 0, c.stream>>>(s->d, total, ins, make_double2(g.p[0], g.p[1]));
            return check_launch("k_diag");
        }
        case IK_RZ: {
            uint64_t total = 1ull << (n - nc);
            BitInsert ins = make_insert({}, ctrl);
            LaunchScope ls(KF_DIAG, 2.0 * state_bytes / (double)(1ull << nc));
            uint64_t tb = g.t0 >= 0 ? (1ull << g.t0) : 0ull;
            k_rz<<<blocks4(total), kBlock, 0, c.stream>>>(s->d, total, ins, tb, make_double2(g.p[0], g.p[1]),
                                                         make_double2(g.p[2], g.p[3]));
            return check_launch("k_rz");
        }
        case IK_SWAP: {
            if (n - 2 - nc < 0) return QI_OK;
            uint64_t total = 1ull << (n - 2 - nc);
            std::vector<int> ones(ctrl);
            ones.push_back(g.t0);
            BitInsert ins = make_insert({g.t1}, ones);
            LaunchScope ls(KF_SWAP, 2.0 * state_bytes / (double)(2ull << nc));
            k_swap<<<blocks(total), kBlock, 0, c.stream>>>(s->d, total, ins, (1ull << g.t0) | (1ull << g.t1));
            return check_launch("k_swap");
        }
        case IK_MATCH: {
            if (n - 2 - nc < 0) return QI_OK;
            uint64_t total = 1ull << (n - 2 - nc);
            BitInsert ins = make_insert({g.t0, g.t1}, ctrl);
            LaunchScope ls(KF_MATCH, 2.0 * state_bytes * 0.75 / (double)(1ull << nc));
            k_match<<<blocks(total), kBlock, 0, c.stream>>>(s->d, total, ins, 1ull << g.t0, 1ull << g.t1, g.p[0], g.p[1],
                                                            make_double2(g.p[2], g.p[3]), make_double2(g.p[4], g.p[5]));
            return check_launch("k_match");
        }
    }
    return fail(QI_ERR_UNKNOWN, 0, 0, "unhandled internal gate kind");
}

// Undo lazy SWAP relabelling: physically swap bit positions until logical qubit q sits at bit q.
int canonicalise(qi_state* s) {
    if (s->identity_layout()) return QI_OK;
    if (s->world > 1) return fail(QI_ERR_PEER, 0, 0, "canonicalise is not available on sharded states");
    std::vector<PhysGate> swaps;
    uint8_t phys[64];
    memcpy(phys, s->phys, sizeof(phys));
    for (uint32_t p = 0; p < s->num_qubits; p++) {
        if (phys[p] == p) continue;
        const int pp = phys[p];                 // logical p currently lives at physical pp
        int q = -1;                             // logical qubit currently at physical p
        for (uint32_t k = 0; k < s->num_qubits; k++) if (phys[k] == p) q = (int)k;
        PhysGate g;
        memset(&g, 0, sizeof(g));
        g.kind = IK_SWAP;
        g.t0 = (int)p;
        g.t1 = pp;
        swaps.push_back(g);
        phys[p] = (uint8_t)p;
        if (q >= 0) phys[q] = (uint8_t)pp;
    }
    if (ctx().opt_path != 1 && window_supported(s)) QI_TRY(run_circuit_windowed(s, swaps));
    else for (const PhysGate& g : swaps) QI_TRY(launch_simple_gate(s, g));
    for (int i = 0; i < 64; i++) s->phys[i] = (uint8_t)i;
    return QI_OK;
}

}  // namespace qi

using namespace qi;

extern "C" {

int qi_unitary2_check(const double m[8]) {
    // operator.rs:2092-2118
    if (!m) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "matrix is NULL");
    const double tol = 2.220446049250313e-16 * 2.0;
    amp_t a = make_double2(m[0], m[1]), b = make_double2(m[2], m[3]), c = make_double2(m[4], m[5]), d = make_double2(m[6], m[7]);
    auto nsq = [](amp_t z) { return z.x * z.x + z.y * z.y; };
    if (std::fabs((nsq(a) + nsq(b)) - 1.0) > tol) return fail(QI_ERR_NON_UNITARY_MATRIX, 0, 0, "Non-unitary matrix");
    if (std::fabs((nsq(c) + nsq(d)) - 1.0) > tol) return fail(QI_ERR_NON_UNITARY_MATRIX, 0, 0, "Non-unitary matrix");
    amp_t dot = cadd(cmul(a, cconj(c)), cmul(b, cconj(d)));
    if (nsq(dot) > tol * tol) return fail(QI_ERR_NON_UNITARY_MATRIX, 0, 0, "Non-unitary matrix");
    return QI_OK;
}

int qi_apply_gate(qi_state* s, const qi_gate* gate) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    QI_TRY(validate_gate(s, gate));
    QI_TRY(ensure_ctx());
    return qi_apply_circuit(s, gate, 1);
}

int qi_apply_circuit(qi_state* s, const qi_gate* gates, uint64_t count) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    if (count && !gates) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "gates is NULL");
    // Gate::apply validates as it goes (circuit.rs:167-169); the first failing gate aborts the run.
    // Here the whole run is validated up front, so a failure leaves the state untouched.
    for (uint64_t i = 0; i < count; i++) QI_TRY(validate_gate(s, &gates[i]));
    std::vector<qi_gate> own;
    gates = normalise_gates(gates, count, &own);
    QI_TRY(ensure_ctx());
    Context& c = ctx();
    const bool use_window = (c.opt_path != 1) && window_supported(s);
    if (s->world > 1) return apply_circuit_sharded(s, gates, count, use_window);     // staged around the exchanges (shard.cu)
    std::vector<PhysGate> run;
    run.reserve(count);
    auto flush = [&]() -> int {
        if (run.empty()) return QI_OK;
        int st = QI_OK;
        if (use_window) st = run_circuit_windowed(s, run, true);
        else for (const PhysGate& g : run) { st = launch_simple_gate(s, g); if (st != QI_OK) break; }
        run.clear();
        return st;
    };
    for (uint64_t i = 0; i < count; i++) {
        PhysGate pg;
        bool skip = false;
        if (gates[i].kind == QI_GATE_SWAP && gates[i].num_controls == 0 && c.opt_lazy_swap) {
            // an uncontrolled SWAP is a relabelling of the logical->physical qubit map: no data moves now
            // (gates queued before it were mapped with the old labels, later ones use the new ones)
            std::swap(s->phys[gates[i].targets[0]], s->phys[gates[i].targets[1]]);
            continue;
        }
        QI_TRY(prepare_gate(s, &gates[i], &pg, &skip));
        if (skip || pg.kind == IK_NOP) continue;
        run.push_back(pg);
    }
    return flush();
}

// Host-only: how the fused executor would split a gate list into passes on a single device (no device access).
// rows[8*i..] = {per-gate-kernel step?, window qubits used, lane-pair ops, register-pair ops, diagonal ops, phase-table ops,
//               absorbed CNOTs, 0}
int qi_debug_schedule(uint32_t num_qubits, const qi_gate* gates, uint64_t count, int window_regs, int32_t* rows,
                      uint64_t max_rows, uint64_t* n_rows) {
    qi_state s;
    s.num_qubits = num_qubits;
    s.n_local = num_qubits;
    s.len = 1ull << num_qubits;
    for (int i = 0; i < 64; i++) s.phys[i] = (uint8_t)i;
    std::vector<PhysGate> run;
    for (uint64_t i = 0; i < count; i++) QI_TRY(validate_gate(&s, &gates[i]));
    std::vector<qi_gate> own;
    gates = normalise_gates(gates, count, &own);
    for (uint64_t i = 0; i < count; i++) {
        if (gates[i].kind == QI_GATE_SWAP && gates[i].num_controls == 0) { std::swap(s.phys[gates[i].targets[0]], s.phys[gates[i].targets[1]]); continue; }
        PhysGate pg;
        bool skip = false;
        QI_TRY(prepare_gate(&s, &gates[i], &pg, &skip));
        if (!skip && pg.kind != IK_NOP) run.push_back(pg);
    }
    std::vector<std::vector<int>> summary;
    QI_TRY(debug_schedule(&s, run, window_regs ? window_regs : 4, &summary));
    uint64_t n = 0;
    for (auto& r : summary) {
        if (n >= max_rows) break;
        for (int k = 0; k < 8; k++) rows[8 * n + k] = r[k];
        n++;
    }
    if (n_rows) *n_rows = summary.size();
    return QI_OK;
}

// Host-only: the device programs the fused executor would launch for a gate list on one device (`rank`/`world` > 1: on
// that shard of a sharded state; `phys` = logical -> physical qubit map it runs under, NULL = identity), serialised for
// a CPU interpreter (layout: window.cu, debug_lower).
// *used = bytes needed; QI_ERR_INVALID_ARGUMENT with payload[0] = needed size when `capacity` is too small.
int qi_debug_lower(uint32_t num_qubits, int rank, int world, const uint8_t* phys, const qi_gate* gates, uint64_t count, int window_regs,
                   uint8_t* blob, uint64_t capacity, uint64_t* used) {
    if (count && !gates) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "gates is NULL");
    if (world < 1 || (world & (world - 1)) || world > 16 || rank < 0 || rank >= world) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)world, 0, "bad rank / world");
    int p = 0;
    while ((1 << p) < world) p++;
    qi_state s;
    s.num_qubits = num_qubits;
    s.n_local = num_qubits - (uint32_t)p;
    s.len = 1ull << s.n_local;
    s.rank = rank;
    s.world = world;
    for (int i = 0; i < 64; i++) s.phys[i] = phys ? phys[i] : (uint8_t)i;
    if (!window_supported(&s)) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, num_qubits, 0, "too few local qubits for the window executor");
    std::vector<PhysGate> run;
    for (uint64_t i = 0; i < count; i++) QI_TRY(validate_gate(&s, &gates[i]));
    std::vector<qi_gate> own;
    gates = normalise_gates(gates, count, &own);
    for (uint64_t i = 0; i < count; i++) {
        if (world == 1 && ctx().opt_lazy_swap && gates[i].kind == QI_GATE_SWAP && gates[i].num_controls == 0) {
            std::swap(s.phys[gates[i].targets[0]], s.phys[gates[i].targets[1]]);
            continue;
        }
        PhysGate pg;
        bool skip = false;
        QI_TRY(prepare_gate(&s, &gates[i], &pg, &skip));
        if (!skip && pg.kind != IK_NOP) run.push_back(pg);
    }
    int R = window_regs ? window_regs : 4;
    while (R > 3 && (int)s.n_local < 5 + R) R--;
    std::vector<uint8_t> out;
    std::vector<int> relabel;
    QI_TRY(debug_lower(&s, run, R, &out, &relabel));
    if (!relabel.empty())          // tile passes left the qubits at other positions
        for (uint32_t q = 0; q < num_qubits; q++) s.phys[q] = (uint8_t)relabel[s.phys[q]];
    // the final logical -> physical map (lazy SWAPs), so the interpreter can undo it
    out.insert(out.end(), s.phys, s.phys + 64);
    if (used) *used = out.size();
    if (out.size() > capacity || !blob) return fail(QI_ERR_INVALID_ARGUMENT, out.size(), capacity, "blob too small");
    memcpy(blob, out.data(), out.size());
    return QI_OK;
}

}  // extern "C"
