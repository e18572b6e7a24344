// window.cu -- the fused register-window executor behind qi_apply_circuit (Circuit::execute's gate
// loop, circuit.rs:160-172, which in the reference is one clone + one full sweep per gate).
//
// One HBM pass = every amplitude is loaded once, any number of gates is applied to it in registers,
// and it is stored once.  Work unit: a WARP TILE of 32 * 2^R amplitudes.
//   lanes      <-> physical qubits 0..4   (32 consecutive amplitudes = one coalesced 512 B access)
//   registers  <-> R "window" qubits chosen per pass (any positions >= 5); slot s of a thread holds
//                  the amplitude whose window bits spell s
// A gate whose target is a window qubit pairs two registers of the same thread (no data movement);
// a target on qubits 0..4 pairs two lanes (warp shuffles); controls and diagonal gates can sit on
// ANY qubit, because every thread knows the full index of each amplitude it holds.  No shared
// memory, no block-level synchronisation.
//
// The host scheduler walks the gate list greedily: a gate joins the current pass if it commutes
// with every gate deferred so far (two gates commute when on each shared qubit both act diagonally)
// and its non-diagonal targets fit the window (5 lane qubits + R free choices).
#include <algorithm>

#include "common.cuh"

namespace qi {

enum { WK_H = 1, WK_X, WK_Y, WK_RX, WK_REAL, WK_U2, WK_DIAG, WK_RZ };

struct WOp {              // 88 bytes
    uint32_t kind;        // WK_*
    uint32_t tpos;        // target: 0..4 = lane bit, 5+j = register bit j (pair ops only)
    uint64_t cmask;       // physical index bits that must all be 1 (for WK_DIAG: includes the target bit)
    uint64_t tmask;       // WK_RZ: the target bit in the physical index
    double m[8];          // WK_U2: m00,m01,m10,m11 (re,im); WK_RX: c,s; WK_REAL: m00,m01,m10,m11; WK_H: 1/sqrt2;
                          // WK_DIAG: phase (re,im); WK_RZ: phase0 (re,im), phase1 (re,im)
};

template <int R>
struct WParams {
    BitInsert ins;              // zero-insert positions of the R window qubits (ascending)
    uint64_t off[1 << R];       // slot -> index offset
    uint32_t nops;
};

__device__ __forceinline__ amp_t shfl_xor_amp(amp_t v, int mask) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}

// ---- pair gate on register bit B --------------------------------------------------------------
template <int R, int B>
__device__ __forceinline__ void reg_pair_op(amp_t (&v)[1 << R], uint64_t base, const uint64_t* off, uint32_t kind,
                                            uint64_t cmask, const double* __restrict__ m) {
#pragma unroll
    for (int p = 0; p < (1 << (R - 1)); p++) {
        const int s0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        const int s1 = s0 | (1 << B);
        const bool ok = ((base | off[s0]) & cmask) == cmask;
        const amp_t a0 = v[s0], a1 = v[s1];
        amp_t r0, r1;
        switch (kind) {
            case WK_H: {
                const double s = m[0];
                r0 = cscale(s, cadd(a0, a1));
                r1 = cscale(s, csub(a0, a1));
                break;
            }
            case WK_X: r0 = a1; r1 = a0; break;
            case WK_Y: r0 = make_double2(a1.y, -a1.x); r1 = make_double2(-a0.y, a0.x); break;
            case WK_RX: {   // [[c, -i s], [-i s, c]]
                const double c = m[0], s = m[1];
                r0 = make_double2(c * a0.x + s * a1.y, c * a0.y - s * a1.x);
                r1 = make_double2(c * a1.x + s * a0.y, c * a1.y - s * a0.x);
                break;
            }
            case WK_REAL: {  // real 2x2
                r0 = make_double2(m[0] * a0.x + m[1] * a1.x, m[0] * a0.y + m[1] * a1.y);
                r1 = make_double2(m[2] * a0.x + m[3] * a1.x, m[2] * a0.y + m[3] * a1.y);
                break;
            }
            default: {       // WK_U2
                const amp_t m00 = make_double2(m[0], m[1]), m01 = make_double2(m[2], m[3]);
                const amp_t m10 = make_double2(m[4], m[5]), m11 = make_double2(m[6], m[7]);
                r0 = cadd(cmul(m00, a0), cmul(m01, a1));
                r1 = cadd(cmul(m10, a0), cmul(m11, a1));
                break;
            }
        }
        v[s0] = ok ? r0 : a0;
        v[s1] = ok ? r1 : a1;
    }
}

// ---- pair gate on lane bit tpos (warp shuffles) --------------------------------------------------
template <int R>
__device__ __forceinline__ void lane_pair_op(amp_t (&v)[1 << R], uint64_t base, const uint64_t* off, uint32_t kind,
                                             uint32_t tpos, uint64_t cmask, const double* __restrict__ m, int lane) {
    const int xm = 1 << tpos;
    const bool hi = (lane >> tpos) & 1;     // this lane holds the |1> member of the pair
    // own/partner coefficients: new = cA * mine + cB * partner
    amp_t cA, cB;
    switch (kind) {
        case WK_H: cA = make_double2(hi ? -m[0] : m[0], 0.0); cB = make_double2(m[0], 0.0); break;
        case WK_X: cA = make_double2(0.0, 0.0); cB = make_double2(1.0, 0.0); break;
        case WK_Y: cA = make_double2(0.0, 0.0); cB = make_double2(0.0, hi ? 1.0 : -1.0); break;
        case WK_RX: cA = make_double2(m[0], 0.0); cB = make_double2(0.0, -m[1]); break;
        case WK_REAL: cA = make_double2(hi ? m[3] : m[0], 0.0); cB = make_double2(hi ? m[2] : m[1], 0.0); break;
        default:
            cA = hi ? make_double2(m[6], m[7]) : make_double2(m[0], m[1]);
            cB = hi ? make_double2(m[4], m[5]) : make_double2(m[2], m[3]);
            break;
    }
    const bool real_only = (kind == WK_H || kind == WK_REAL);
#pragma unroll
    for (int s = 0; s < (1 << R); s++) {
        const amp_t mine = v[s];
        const amp_t other = shfl_xor_amp(mine, xm);
        const bool ok = ((base | off[s]) & cmask) == cmask;   // cmask never contains the target bit
        amp_t r;
        if (kind == WK_X) r = other;
        else if (real_only) r = make_double2(cA.x * mine.x + cB.x * other.x, cA.x * mine.y + cB.x * other.y);
        else r = cadd(cmul(cA, mine), cmul(cB, other));
        v[s] = ok ? r : mine;
    }
}

template <int R>
__device__ __forceinline__ void diag_op(amp_t (&v)[1 << R], uint64_t base, const uint64_t* off, uint64_t mask, amp_t ph) {
#pragma unroll
    for (int s = 0; s < (1 << R); s++) {
        const bool ok = ((base | off[s]) & mask) == mask;
        const amp_t r = cmul(v[s], ph);
        v[s] = ok ? r : v[s];
    }
}

template <int R>
__device__ __forceinline__ void rz_op(amp_t (&v)[1 << R], uint64_t base, const uint64_t* off, uint64_t cmask, uint64_t tmask,
                                      amp_t p0, amp_t p1) {
#pragma unroll
    for (int s = 0; s < (1 << R); s++) {
        const uint64_t idx = base | off[s];
        const bool ok = (idx & cmask) == cmask;
        const amp_t ph = (idx & tmask) ? p1 : p0;
        const amp_t r = cmul(v[s], ph);
        v[s] = ok ? r : v[s];
    }
}

template <int R>
__global__ void __launch_bounds__(128) k_window(amp_t* __restrict__ a, uint64_t ntiles, const __grid_constant__ WParams<R> P,
                                                const WOp* __restrict__ ops) {
    constexpr int S = 1 << R;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t tile = warp; tile < ntiles; tile += nwarps) {
        const uint64_t base = expand_index((tile << 5) | (uint64_t)lane, P.ins);
        amp_t v[S];
#pragma unroll
        for (int s = 0; s < S; s++) v[s] = a[base + P.off[s]];
#pragma unroll 1
        for (uint32_t o = 0; o < P.nops; o++) {
            const WOp* op = ops + o;
            const uint32_t kind = op->kind, tpos = op->tpos;
            const uint64_t cmask = op->cmask;
            if (kind == WK_DIAG) {
                diag_op<R>(v, base, P.off, cmask, make_double2(op->m[0], op->m[1]));
            } else if (kind == WK_RZ) {
                rz_op<R>(v, base, P.off, cmask, op->tmask, make_double2(op->m[0], op->m[1]), make_double2(op->m[2], op->m[3]));
            } else if (tpos < 5) {
                lane_pair_op<R>(v, base, P.off, kind, tpos, cmask, op->m, lane);
            } else {
                switch (tpos - 5) {
                    case 0: reg_pair_op<R, 0>(v, base, P.off, kind, cmask, op->m); break;
                    case 1: if (R > 1) reg_pair_op<R, (R > 1 ? 1 : 0)>(v, base, P.off, kind, cmask, op->m); break;
                    case 2: if (R > 2) reg_pair_op<R, (R > 2 ? 2 : 0)>(v, base, P.off, kind, cmask, op->m); break;
                    case 3: if (R > 3) reg_pair_op<R, (R > 3 ? 3 : 0)>(v, base, P.off, kind, cmask, op->m); break;
                    case 4: if (R > 4) reg_pair_op<R, (R > 4 ? 4 : 0)>(v, base, P.off, kind, cmask, op->m); break;
                    default: break;
                }
            }
        }
#pragma unroll
        for (int s = 0; s < S; s++) a[base + P.off[s]] = v[s];
    }
}

// ---- host: scheduling ------------------------------------------------------------------------------
static const int kR = 4;            // register qubits per pass
static const int kLaneQubits = 5;

struct GateUse {
    uint64_t n_use;   // qubits used non-diagonally (targets of H/X/Y/U2/SWAP)
    uint64_t d_use;   // qubits used diagonally (controls, targets of phase gates)
};

static GateUse uses_of(const PhysGate& g) {
    GateUse u{0, g.cmask};
    switch (g.kind) {
        case IK_H: case IK_X: case IK_Y: case IK_U2: u.n_use = 1ull << g.t0; break;
        case IK_SWAP: u.n_use = (1ull << g.t0) | (1ull << g.t1); break;
        case IK_DIAG: case IK_RZ: if (g.t0 >= 0) u.d_use |= 1ull << g.t0; break;
        default: break;
    }
    return u;
}

static void classify_u2(const double* p, WOp* op) {
    const bool rx_form = p[1] == 0.0 && p[2] == 0.0 && p[4] == 0.0 && p[7] == 0.0 && p[0] == p[6] && p[3] == p[5];
    const bool real_form = p[1] == 0.0 && p[3] == 0.0 && p[5] == 0.0 && p[7] == 0.0;
    memset(op->m, 0, sizeof(op->m));
    if (real_form) { op->kind = WK_REAL; op->m[0] = p[0]; op->m[1] = p[2]; op->m[2] = p[4]; op->m[3] = p[6]; }
    else if (rx_form) { op->kind = WK_RX; op->m[0] = p[0]; op->m[1] = -p[3]; }
    else { op->kind = WK_U2; memcpy(op->m, p, 8 * sizeof(double)); }
}

struct Pass {
    std::vector<int> regs;      // window qubits (physical positions >= 5), position j <-> register bit j after sort
    std::vector<WOp> ops;       // tpos filled after the window is final
    std::vector<int> op_target; // physical target of each pair op (-1 for diagonal ops)
    double unfused_bytes = 0.0;
};

bool window_supported(const qi_state* s) {
    return s->consistent && (int)s->n_local >= kLaneQubits + kR;
}

static bool window_takes(const PhysGate& g) {
    return g.kind == IK_H || g.kind == IK_X || g.kind == IK_Y || g.kind == IK_U2 || g.kind == IK_DIAG || g.kind == IK_RZ ||
           g.kind == IK_SWAP;
}

static void push_pair(Pass& ps, uint32_t kind, int target, uint64_t cmask, const double* m8) {
    WOp op;
    memset(&op, 0, sizeof(op));
    op.kind = kind;
    op.cmask = cmask;
    if (m8) memcpy(op.m, m8, 8 * sizeof(double));
    ps.ops.push_back(op);
    ps.op_target.push_back(target);
}

static void lower_gate(Pass& ps, const PhysGate& g) {
    WOp op;
    memset(&op, 0, sizeof(op));
    switch (g.kind) {
        case IK_H: { double m[8] = {g.p[0]}; push_pair(ps, WK_H, g.t0, g.cmask, m); break; }
        case IK_X: push_pair(ps, WK_X, g.t0, g.cmask, nullptr); break;
        case IK_Y: push_pair(ps, WK_Y, g.t0, g.cmask, nullptr); break;
        case IK_U2: classify_u2(g.p, &op); push_pair(ps, op.kind, g.t0, g.cmask, op.m); break;
        case IK_SWAP: {   // SWAP(a,b) = CX(b->a) CX(a->b) CX(b->a), each under the original controls
            push_pair(ps, WK_X, g.t0, g.cmask | (1ull << g.t1), nullptr);
            push_pair(ps, WK_X, g.t1, g.cmask | (1ull << g.t0), nullptr);
            push_pair(ps, WK_X, g.t0, g.cmask | (1ull << g.t1), nullptr);
            break;
        }
        case IK_DIAG: {
            op.kind = WK_DIAG;
            op.cmask = g.cmask | (g.t0 >= 0 ? (1ull << g.t0) : 0ull);
            op.m[0] = g.p[0]; op.m[1] = g.p[1];
            ps.ops.push_back(op);
            ps.op_target.push_back(-1);
            break;
        }
        case IK_RZ: {
            op.kind = WK_RZ;
            op.cmask = g.cmask;
            op.tmask = g.t0 >= 0 ? (1ull << g.t0) : 0ull;
            memcpy(op.m, g.p, 4 * sizeof(double));
            if (g.t0 < 0) { /* target in the rank bits: caller resolved the phase into p[0..1] */ op.m[2] = g.p[0]; op.m[3] = g.p[1]; }
            ps.ops.push_back(op);
            ps.op_target.push_back(-1);
            break;
        }
        default: break;
    }
}

static double gate_unfused_bytes(const qi_state* s, const PhysGate& g) {
    const double full = 32.0 * (double)s->len;
    int nc = __builtin_popcountll(g.cmask);
    double f = 1.0;
    if (g.kind == IK_DIAG) f = g.t0 >= 0 ? 0.5 : 1.0;
    if (g.kind == IK_SWAP) f = 0.5;
    return full * f / (double)(1ull << nc);
}

// finalise a pass on the host: pad + sort the window, slot offsets, op target positions
template <int R>
static void finalise_pass(const qi_state* s, Pass& ps, WParams<R>* P) {
    const int n = (int)s->n_local;
    std::sort(ps.regs.begin(), ps.regs.end());
    // pad the window with unused qubits (lowest free positions first: better locality)
    for (int q = kLaneQubits; q < n && (int)ps.regs.size() < R; q++)
        if (std::find(ps.regs.begin(), ps.regs.end(), q) == ps.regs.end()) ps.regs.push_back(q);
    std::sort(ps.regs.begin(), ps.regs.end());
    memset(P, 0, sizeof(*P));
    P->ins = make_insert(ps.regs, {});
    for (int sidx = 0; sidx < (1 << R); sidx++) {
        uint64_t o = 0;
        for (int j = 0; j < R; j++) if ((sidx >> j) & 1) o |= 1ull << ps.regs[j];
        P->off[sidx] = o;
    }
    for (size_t i = 0; i < ps.ops.size(); i++) {
        int t = ps.op_target[i];
        if (t < 0) continue;
        if (t < kLaneQubits) ps.ops[i].tpos = (uint32_t)t;
        else ps.ops[i].tpos = (uint32_t)(kLaneQubits + (std::find(ps.regs.begin(), ps.regs.end(), t) - ps.regs.begin()));
    }
    P->nops = (uint32_t)ps.ops.size();
}

template <int R>
static int launch_pass(qi_state* s, const WParams<R>& P, const WOp* d_ops_slot) {
    Context& c = ctx();
    const uint64_t ntiles = s->len >> (kLaneQubits + R);
    const int warps_per_block = 4;
    uint64_t blocks = (ntiles + warps_per_block - 1) / warps_per_block;
    const uint64_t cap = (uint64_t)c.sm_count * 4 * 8;      // 8 waves of 4 resident blocks per SM
    if (blocks > cap) blocks = cap;
    LaunchScope ls(KF_WINDOW, 32.0 * (double)s->len);
    k_window<R><<<(unsigned)blocks, warps_per_block * 32, 0, c.stream>>>(s->d, ntiles, P, d_ops_slot);
    return check_launch("k_window");
}

// staging for op programs: pinned host buffer + device buffer, reused across calls
static int ensure_ops(size_t count) {
    Context& c = ctx();
    if (!c.ops_event) QI_CUDA(cudaEventCreateWithFlags(&c.ops_event, cudaEventDisableTiming));
    if (c.ops_cap >= count) return QI_OK;
    QI_CUDA(cudaStreamSynchronize(c.stream));
    if (c.h_ops) cudaFreeHost(c.h_ops);
    if (c.d_ops) cudaFree(c.d_ops);
    c.h_ops = c.d_ops = nullptr;
    size_t cap = count < 4096 ? 4096 : count * 2;
    QI_CUDA(cudaMallocHost(&c.h_ops, cap * sizeof(WOp)));
    QI_CUDA(cudaMalloc(&c.d_ops, cap * sizeof(WOp)));
    c.ops_cap = cap;
    return QI_OK;
}

struct Step { bool simple; size_t gate; Pass pass; size_t op_offset; };

int run_circuit_windowed(qi_state* s, const std::vector<PhysGate>& gates) {
    Context& c = ctx();
    const bool fuse = c.opt_fuse != 0;
    const size_t G = gates.size();
    std::vector<char> done(G, 0);
    std::vector<Step> steps;
    size_t first = 0;           // first gate not yet scheduled
    const size_t kLookahead = 4096;
    while (first < G) {
        if (done[first]) { first++; continue; }
        if (!window_takes(gates[first])) {
            Step st{true, first, Pass(), 0};
            steps.push_back(std::move(st));
            done[first++] = 1;
            continue;
        }
        Pass ps;
        uint64_t blocked_any = 0, blocked_n = 0, window_mask = 0;
        size_t scanned = 0;
        for (size_t i = first; i < G && scanned < kLookahead; i++) {
            if (done[i]) continue;
            scanned++;
            const PhysGate& g = gates[i];
            if (!window_takes(g)) break;      // barrier: nothing may move across an unsupported gate
            GateUse u = uses_of(g);
            bool take = ((u.n_use & blocked_any) == 0) && ((u.d_use & blocked_n) == 0);
            if (take) {
                // non-diagonal targets above the lane qubits must be (or become) window qubits
                uint64_t need = u.n_use & ~((1ull << kLaneQubits) - 1) & ~window_mask;
                if ((int)ps.regs.size() + __builtin_popcountll(need) > kR) take = false;
                else
                    for (int q = kLaneQubits; q < 64 && need; q++)
                        if ((need >> q) & 1) { ps.regs.push_back(q); window_mask |= 1ull << q; need &= ~(1ull << q); }
            }
            if (take) {
                lower_gate(ps, g);
                ps.unfused_bytes += gate_unfused_bytes(s, g);
                done[i] = 1;
                if (!fuse) break;
            } else {
                blocked_any |= u.n_use | u.d_use;
                blocked_n |= u.n_use;
            }
        }
        if (ps.ops.empty()) return fail(QI_ERR_UNKNOWN, 0, 0, "scheduler made no progress");
        Step st{false, 0, std::move(ps), 0};
        steps.push_back(std::move(st));
    }
    // one staging copy for the whole run, then back-to-back launches
    size_t total_ops = 0;
    std::vector<WParams<kR>> params(steps.size());
    for (size_t i = 0; i < steps.size(); i++) {
        if (steps[i].simple) continue;
        finalise_pass<kR>(s, steps[i].pass, &params[i]);
        steps[i].op_offset = total_ops;
        total_ops += steps[i].pass.ops.size();
    }
    if (total_ops) {
        QI_TRY(ensure_ops(total_ops));
        QI_CUDA(cudaEventSynchronize(c.ops_event));      // previous run's copy has left the pinned buffer
        WOp* h = (WOp*)c.h_ops;
        for (const Step& st : steps)
            if (!st.simple) memcpy(h + st.op_offset, st.pass.ops.data(), st.pass.ops.size() * sizeof(WOp));
        QI_CUDA(cudaMemcpyAsync(c.d_ops, c.h_ops, total_ops * sizeof(WOp), cudaMemcpyHostToDevice, c.stream));
        QI_CUDA(cudaEventRecord(c.ops_event, c.stream));
    }
    for (size_t i = 0; i < steps.size(); i++) {
        if (steps[i].simple) QI_TRY(launch_simple_gate(s, gates[steps[i].gate]));
        else QI_TRY(launch_pass<kR>(s, params[i], (const WOp*)c.d_ops + steps[i].op_offset));
    }
    return QI_OK;
}

}  // namespace qi
