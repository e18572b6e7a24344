// window.cu -- the fused register-window executor behind qi_apply_circuit (Circuit::execute's gate
// loop, circuit.rs:160-172, which in the reference is one clone + one full sweep per gate).
//
// One HBM pass = every amplitude is loaded once, any number of gates is applied to it in registers,
// and it is stored once.  Work unit: a WARP TILE of 32 * 2^R amplitudes.
//   lanes      <-> physical qubits 0..4   (32 consecutive amplitudes = one coalesced 512 B access)
//   registers  <-> R "window" qubits chosen per pass (any positions >= 5); slot s of a thread holds
//                  the amplitude whose window bits spell s
//   tile index <-> all remaining qubits (uniform across the warp)
// A gate whose target is a window qubit pairs two registers of the same thread (no data movement);
// a target on qubits 0..4 pairs two lanes (warp shuffles); controls and diagonal gates can sit on
// ANY qubit: a control on a tile bit is a warp-uniform skip, on a lane bit a per-thread predicate,
// on a register bit a per-slot uniform predicate (a host-expanded slot mask, which also carries
// negative controls).  The default kernel uses no shared memory and no block synchronisation.
// The op program of a pass travels in the kernel's parameter space (constant bank, uniform loads).
//
// Diagonal gates that meet in a pass are merged into PHASE-TABLE ops: a group "if hub bit set,
// multiply by prod_j f_j(bit_j)" (the QFT's H + controlled-phase ladder is exactly one such group
// per Hadamard) is applied with one lookup per lane/slot/tile-chunk table instead of one sweep per
// gate (subroutine.rs:93-106 emits n-1-i controlled phases after H(i)).
//
// The host scheduler walks the gate list greedily: a gate joins the current pass if it commutes
// with every gate deferred so far (two gates commute when on each shared qubit both act diagonally)
// and its non-diagonal targets fit the window (5 lane qubits + R free choices).  A CNOT next to a
// single-qubit gate on its target is absorbed into it (absorb_cnot_*).
//
// Performance notes that shaped the code (DESIGN.md 3.1 has the measurements):
//   * gates without register-bit controls take straight-line variants: per-slot predicate regions
//     serialise the slots and expose one shuffle / FP64 latency per slot;
//   * the kernel is instantiated by pass content (k_window<R, LANES, U2K>): passes without lane
//     gates / complex 2x2 gates run code that does not carry those paths;
//   * k_window_tma is a TMA-prefetched persistent variant kept behind the "tma" option (slower).
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdlib>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "common.cuh"
#include "window_layout.cuh"

namespace qi {

// pair kinds first (<= kLastPairKind); RXS = RX o X.  The *U kinds are the LEAN forms (option "lean", off by default, not yet
// measured): the gate divided by its top-left entry g, so that two of the four coefficients are +-1 and a pair costs 4 DFMA
// instead of 4 DMUL + 4 DFMA; the g of all lean gates of a pass are multiplied on the host and applied once (folded into a
// phase table when the pass has an unconditional one, else one WK_SCALE op).
//   REALUP = [[1, p], [q, 1]]   REALUM = [[1, p], [q, -1]]   (m[0] = p, m[1] = q; H / g = REALUM with p = q = 1)
//   RXU = [[1, -i t], [-i t, 1]]   RXSU = RXU with its inputs swapped   (m[0] = t = tan(theta / 2))
enum { WK_X = 1, WK_RX, WK_RXS, WK_REAL, WK_U2, WK_REALUP, WK_REALUM, WK_RXU, WK_RXSU, WK_DIAG, WK_RZ, WK_TABLE, WK_SCALE,
       WK_NEG /* WK_DIAG with phase -1 (Z, CZ): sign flips, no FP64 work */,
       // k_tile only: LIFTED forms of REAL / RX -- the same 2x2 as two in-place accumulations per amplitude (see fast routines)
       WK_REALL /* m = (k00, k01, k10 / k00, det / k00) */, WK_RXL /* m = (c, s, 1 / c, s / c) */ };
static const int kLastPairKind = WK_RXSU;

static const int kMaxOps = 120;      // per launch (parameter space: 120 * 112 B + header < 16 KB)
// CTA-tile kernel (k_tile, below)
static const int kTileBits = 11;            // local bits of a CTA tile
static const int kTileThreads = 128;        // 2^(kTileBits - 4)
static const int kTileThrBits = 7;
static const int kTileWindow = kTileBits - kLaneQubits;   // window qubits per pass (6)
static const int kMaxRounds = 24;
static const int kMaxTileOps = 232;         // parameter space: 232 * 112 B + rounds + header < 32 KiB

struct alignas(16) DOp {     // device op, 112 bytes; the first 32 bytes are everything the dispatch needs (two 16-byte loads)
    uint8_t kind;            // WK_*
    uint8_t tpos;            // pair ops: 0..4 lane bit, 5+j register bit j
    uint8_t hub_cls;         // WK_TABLE: class of the hub bit (CLS_*)
    uint8_t hub_bit;         // WK_TABLE: lane bit / slot bit / compact tile bit index
    uint8_t nchunks;         // WK_TABLE: number of 8-bit tile chunks with a table
    uint8_t has_reg;         // WK_TABLE: register table is not all ones
    uint8_t c_lval;          // value the c_lane bits of the lane / thread index must have: (idx & c_lane) == c_lval
    uint8_t code;            // k_tile: index of the op's specialised straight-line routine (fast_code), 0 = generic interpreter
    uint32_t c_lane, c_reg;  // c_lane: control bits in lane space (k_window: lane bits 0..4; k_tile: the 7 thread-index bits).
                             // c_reg: SLOT MASK -- bit s is set iff slot s passes the register-bit controls (positive and
                             // negative), expanded on the host
    uint64_t c_tile;         // control bits in compact tile-index space (positive and negative controls)
    uint64_t c_tval;         // value those bits must have: (tile & c_tile) == c_tval
    uint32_t t_lane, t_reg;  // WK_RZ: target bit in lane space / slot space (0 if elsewhere)
    uint64_t t_tile;         // WK_RZ: target bit in tile space
    double m[8];             // matrix / phases; WK_TABLE: m[0] holds the table offset (as integer bits)
};

static_assert(sizeof(DOp) == 112, "DOp layout (the CPU interpreter in tests/ parses it)");
static_assert(sizeof(PhysGate) == 88, "PhysGate layout (the CPU interpreter in tests/ parses it)");

template <int R>
struct WProgram {
    BitInsert ins;              // zero-insert positions of the R window qubits (ascending)
    uint64_t off[1 << R];       // slot -> index offset
    uint32_t nops;
    uint32_t flags;             // bit 0: prefetch the warp's next tile into L2 (option "prefetch"; lean instantiations only)
    const amp_t* tables;        // phase-table arena
    DOp ops[kMaxOps];
};

// the state is streamed once per pass and is far larger than L2: optional streaming cache hints
#ifndef QI_NO_STREAM_HINTS
#define QI_LD(p) __ldcs(p)
#define QI_ST(p, v) __stcs(p, v)
#else
#define QI_LD(p) (*(p))
#define QI_ST(p, v) (*(p) = (v))
#endif

__device__ __forceinline__ amp_t shfl_xor_amp(amp_t v, int mask) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}

// ---- in-place 2x2 updates -------------------------------------------------------------------------------
// Written as PTX with tied ("+d") operands: every amplitude keeps its virtual register across an op, so the op loop has no
// phi nodes at its merge points.  The plain C++ forms made ptxas finish every variant in its own register assignment and
// restore the canonical one with ~64 MOVs per op at the loop back-edge (17 % of all executed instructions, ncu r02b).
// Same operation order as the C++ forms after FMA contraction: one product rounded, then one fused multiply-add.
// REAL: a0' = k0 a0 + k1 a1, a1' = k2 a0 + k3 a1 (real coefficients, applied to .x and .y alike)
__device__ __forceinline__ void upd_real(amp_t& a0, amp_t& a1, double k0, double k1, double k2, double k3) {
    asm("{\n\t.reg .f64 t0, t1, t2, t3;\n\t"
        "mul.f64 t0, %4, %0;\n\tmul.f64 t1, %6, %0;\n\tmul.f64 t2, %4, %1;\n\tmul.f64 t3, %6, %1;\n\t"
        "fma.rn.f64 %0, %5, %2, t0;\n\tfma.rn.f64 %1, %5, %3, t2;\n\tfma.rn.f64 %2, %7, %2, t1;\n\tfma.rn.f64 %3, %7, %3, t3;\n\t}"
        : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(k0), "d"(k1), "d"(k2), "d"(k3));
}
// RX: a0' = c a0 - i s a1, a1' = c a1 - i s a0      (x' = c x + s y_other, y' = c y - s x_other)
__device__ __forceinline__ void upd_rx(amp_t& a0, amp_t& a1, double c, double sn) {
    asm("{\n\t.reg .f64 t0, t1, t2, t3, ns;\n\tneg.f64 ns, %5;\n\t"
        "mul.f64 t0, %5, %3;\n\tmul.f64 t1, ns, %2;\n\tmul.f64 t2, %5, %1;\n\tmul.f64 t3, ns, %0;\n\t"      // s a1y, -s a1x, s a0y, -s a0x
        "fma.rn.f64 %0, %4, %0, t0;\n\tfma.rn.f64 %1, %4, %1, t1;\n\tfma.rn.f64 %2, %4, %2, t2;\n\tfma.rn.f64 %3, %4, %3, t3;\n\t}"
        : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(c), "d"(sn));
}
// RXS = RX with its inputs swapped: a0' = c a1 - i s a0, a1' = c a0 - i s a1
__device__ __forceinline__ void upd_rxs(amp_t& a0, amp_t& a1, double c, double sn) {
    asm("{\n\t.reg .f64 t0, t1, t2, t3, ns;\n\tneg.f64 ns, %5;\n\t"
        "mul.f64 t0, %5, %1;\n\tmul.f64 t1, ns, %0;\n\tmul.f64 t2, %5, %3;\n\tmul.f64 t3, ns, %2;\n\t"      // s a0y, -s a0x, s a1y, -s a1x
        "fma.rn.f64 t0, %4, %2, t0;\n\tfma.rn.f64 t1, %4, %3, t1;\n\tfma.rn.f64 %2, %4, %0, t2;\n\tfma.rn.f64 %3, %4, %1, t3;\n\t"
        "mov.f64 %0, t0;\n\tmov.f64 %1, t1;\n\t}"
        : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(c), "d"(sn));
}
// lean unit forms: [[1, p], [q, +-1]]
template <bool MINUS>
__device__ __forceinline__ void upd_realu(amp_t& a0, amp_t& a1, double p, double q) {
    if (MINUS)
        asm("{\n\t.reg .f64 t0, t1, n2, n3;\n\tneg.f64 n2, %2;\n\tneg.f64 n3, %3;\n\t"
            "fma.rn.f64 t0, %4, %2, %0;\n\tfma.rn.f64 t1, %4, %3, %1;\n\tfma.rn.f64 %2, %5, %0, n2;\n\tfma.rn.f64 %3, %5, %1, n3;\n\t"
            "mov.f64 %0, t0;\n\tmov.f64 %1, t1;\n\t}"
            : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(p), "d"(q));
    else
        asm("{\n\t.reg .f64 t0, t1;\n\t"
            "fma.rn.f64 t0, %4, %2, %0;\n\tfma.rn.f64 t1, %4, %3, %1;\n\tfma.rn.f64 %2, %5, %0, %2;\n\tfma.rn.f64 %3, %5, %1, %3;\n\t"
            "mov.f64 %0, t0;\n\tmov.f64 %1, t1;\n\t}"
            : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(p), "d"(q));
}
// RXU = [[1, -i t], [-i t, 1]]: a0' = a0 - i t a1, a1' = a1 - i t a0;  SWAPPED (RXSU): a0' = a1 - i t a0, a1' = a0 - i t a1
template <bool SWAPPED>
__device__ __forceinline__ void upd_rxu(amp_t& a0, amp_t& a1, double t) {
    if (!SWAPPED)
        asm("{\n\t.reg .f64 t0, t1, nt;\n\tneg.f64 nt, %4;\n\t"
            "fma.rn.f64 t0, %4, %3, %0;\n\tfma.rn.f64 t1, nt, %2, %1;\n\tfma.rn.f64 %2, %4, %1, %2;\n\tfma.rn.f64 %3, nt, %0, %3;\n\t"
            "mov.f64 %0, t0;\n\tmov.f64 %1, t1;\n\t}"
            : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(t));
    else
        asm("{\n\t.reg .f64 t0, t1, t2, t3, nt;\n\tneg.f64 nt, %4;\n\t"
            "fma.rn.f64 t0, %4, %1, %2;\n\tfma.rn.f64 t1, nt, %0, %3;\n\tfma.rn.f64 t2, %4, %3, %0;\n\tfma.rn.f64 t3, nt, %2, %1;\n\t"
            "mov.f64 %0, t0;\n\tmov.f64 %1, t1;\n\tmov.f64 %2, t2;\n\tmov.f64 %3, t3;\n\t}"
            : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(t));
}

// a *= f (complex): x' = x fx - y fy, y' = x fy + y fx
__device__ __forceinline__ void upd_cmul(amp_t& a, const amp_t f) {
    asm("{\n\t.reg .f64 t0, t1, t2;\n\t"
        "mul.f64 t0, %1, %3;\n\tneg.f64 t0, t0;\n\tmul.f64 t1, %1, %2;\n\tfma.rn.f64 t2, %0, %3, t1;\n\tfma.rn.f64 %0, %0, %2, t0;\n\tmov.f64 %1, t2;\n\t}"
        : "+d"(a.x), "+d"(a.y) : "d"(f.x), "d"(f.y));
}
__device__ __forceinline__ void upd_scale(amp_t& a, const double g) {
    asm("mul.f64 %0, %0, %2;\n\tmul.f64 %1, %1, %2;" : "+d"(a.x), "+d"(a.y) : "d"(g));
}

// ---- pair gate on register bit B --------------------------------------------------------------
// COND = false: no register-bit controls, straight-line code.  COND = true: the slot predicate
// bit s0 of the slot mask c_reg is warp-uniform (c_reg comes from the constant bank, s0 is a literal), so a
// controlled gate costs a uniform branch per pair, not a select per register.
// Kinds: X (swap), RX ([[c,-is],[-is,c]]), RXS (RX with its inputs swapped = RX o X = X o RX), REAL (real 2x2;
// H is REAL), U2 (complex 2x2; Y is U2).
template <int R, int B, int KIND, bool COND>
__device__ __forceinline__ void reg_pair_kind(amp_t (&v)[1 << R], const uint32_t c_reg, const double* __restrict__ m) {
    double k0 = 0, k1 = 0, k2 = 0, k3 = 0, k4 = 0, k5 = 0, k6 = 0, k7 = 0;
    if (KIND == WK_RX || KIND == WK_RXS) { k0 = m[0]; k1 = m[1]; }
    if (KIND == WK_REAL) { k0 = m[0]; k1 = m[1]; k2 = m[2]; k3 = m[3]; }
    if (KIND == WK_U2) { k0 = m[0]; k1 = m[1]; k2 = m[2]; k3 = m[3]; k4 = m[4]; k5 = m[5]; k6 = m[6]; k7 = m[7]; }
    if (KIND == WK_REALUP || KIND == WK_REALUM || KIND == WK_RXU || KIND == WK_RXSU) { k0 = m[0]; k1 = m[1]; }
#pragma unroll
    for (int p = 0; p < (1 << (R - 1)); p++) {
        const int s0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        const int s1 = s0 | (1 << B);
        if (!COND || ((c_reg >> s0) & 1u)) {
            if (KIND == WK_X) {
                const amp_t a0 = v[s0], a1 = v[s1];
                v[s0] = a1; v[s1] = a0;
            } else if (KIND == WK_RX) upd_rx(v[s0], v[s1], k0, k1);
            else if (KIND == WK_RXS) upd_rxs(v[s0], v[s1], k0, k1);
            else if (KIND == WK_REAL) upd_real(v[s0], v[s1], k0, k1, k2, k3);
            else if (KIND == WK_REALUP) upd_realu<false>(v[s0], v[s1], k0, k1);
            else if (KIND == WK_REALUM) upd_realu<true>(v[s0], v[s1], k0, k1);
            else if (KIND == WK_RXU) upd_rxu<false>(v[s0], v[s1], k0);
            else if (KIND == WK_RXSU) upd_rxu<true>(v[s0], v[s1], k0);
            else {
                const amp_t a0 = v[s0], a1 = v[s1];
                const amp_t m00 = make_double2(k0, k1), m01 = make_double2(k2, k3);
                const amp_t m10 = make_double2(k4, k5), m11 = make_double2(k6, k7);
                v[s0] = cadd(cmul(m00, a0), cmul(m01, a1));
                v[s1] = cadd(cmul(m10, a0), cmul(m11, a1));
            }
        }
    }
}

template <int R>
constexpr uint32_t kAllSlots = (R >= 5) ? 0xffffffffu : ((1u << (1 << R)) - 1u);

template <int R, int B, bool U2K, bool LEAN>
__device__ __forceinline__ void reg_pair_op(amp_t (&v)[1 << R], uint32_t kind, uint32_t c_reg, const double* __restrict__ m) {
    if (LEAN && kind >= WK_REALUP) {       // lean forms: uncontrolled ones straight-line, halves of an absorbed CNOT predicated
        if (c_reg == kAllSlots<R>) {
            switch (kind) {
                case WK_REALUP: reg_pair_kind<R, B, WK_REALUP, false>(v, 0, m); return;
                case WK_REALUM: reg_pair_kind<R, B, WK_REALUM, false>(v, 0, m); return;
                case WK_RXU: reg_pair_kind<R, B, WK_RXU, false>(v, 0, m); return;
                default: reg_pair_kind<R, B, WK_RXSU, false>(v, 0, m); return;
            }
        }
        switch (kind) {
            case WK_REALUP: reg_pair_kind<R, B, WK_REALUP, true>(v, c_reg, m); return;
            case WK_REALUM: reg_pair_kind<R, B, WK_REALUM, true>(v, c_reg, m); return;
            case WK_RXU: reg_pair_kind<R, B, WK_RXU, true>(v, c_reg, m); return;
            default: reg_pair_kind<R, B, WK_RXSU, true>(v, c_reg, m); return;
        }
    }
    // no register-bit controls (the bulk of every circuit): straight-line variants without slot predicates -- the
    // per-pair predicate regions keep the compiler from interleaving the pairs, which exposes the FP64 latency.
    // (X and U2 stay predicated: their straight-line forms spill hundreds of bytes.)
    if (c_reg == kAllSlots<R>) {
        switch (kind) {
            case WK_REAL: reg_pair_kind<R, B, WK_REAL, false>(v, 0, m); return;
            case WK_RX: reg_pair_kind<R, B, WK_RX, false>(v, 0, m); return;
            case WK_RXS: reg_pair_kind<R, B, WK_RXS, false>(v, 0, m); return;
            default: break;
        }
    }
    switch (kind) {
        case WK_X: reg_pair_kind<R, B, WK_X, true>(v, c_reg, m); break;
        case WK_RX: reg_pair_kind<R, B, WK_RX, true>(v, c_reg, m); break;
        case WK_RXS: reg_pair_kind<R, B, WK_RXS, true>(v, c_reg, m); break;
        case WK_REAL: reg_pair_kind<R, B, WK_REAL, true>(v, c_reg, m); break;
        default: if (U2K) reg_pair_kind<R, B, WK_U2, true>(v, c_reg, m); break;
    }
}

// ---- pair gate on lane bit tpos (warp shuffles; every lane takes part in the exchange) -----------
// A lane whose lane-bit controls are off keeps its value: its coefficients become (1, 0); its partner
// differs only in the target bit, so it sees the same controls.
// ALL = every slot takes part (no register-bit controls): straight-line code, so the shuffles of a gate are issued
// back to back (one exposed latency per gate) instead of one latency-exposed shuffle group per predicated slot.
template <int R, bool ALL, bool U2K, bool LEAN>
__device__ __forceinline__ void lane_pair_impl(amp_t (&v)[1 << R], uint32_t kind, uint32_t tpos, uint32_t c_reg, bool thread_ok,
                                               const double* __restrict__ m, int lane) {
    constexpr int S = 1 << R;
    const int xm = 1 << tpos;
    const bool hi = (lane >> tpos) & 1;     // this lane holds the |1> member of the pair
    if (LEAN && kind >= WK_REALUP) {
        if (kind == WK_REALUP || kind == WK_REALUM) {
            // [[1, p], [q, r]], r = +-1: |0> lane: mine + p other; |1> lane: r mine + q other
            double cB = hi ? m[1] : m[0];
            const bool neg = hi && kind == WK_REALUM && thread_ok;
            if (!thread_ok) cB = 0.0;
#pragma unroll
            for (int s = 0; s < S; s++) {
                if (ALL || ((c_reg >> s) & 1u)) {
                    const amp_t mine = v[s];
                    const amp_t other = shfl_xor_amp(mine, xm);
                    const amp_t base = neg ? cneg(mine) : mine;
                    v[s] = make_double2(fma(cB, other.x, base.x), fma(cB, other.y, base.y));
                }
            }
            return;
        }
        // RXU: mine' = mine - i t other.  RXSU (inputs swapped): mine' = other - i t mine
        const bool swapped = (kind == WK_RXSU) && thread_ok;
        const double t = thread_ok ? m[0] : 0.0;
#pragma unroll
        for (int s = 0; s < S; s++) {
            if (ALL || ((c_reg >> s) & 1u)) {
                const amp_t mine = v[s];
                const amp_t other = shfl_xor_amp(mine, xm);
                const amp_t P = swapped ? other : mine, Q = swapped ? mine : other;
                v[s] = make_double2(fma(t, Q.y, P.x), fma(-t, Q.x, P.y));
            }
        }
        return;
    }
    if (kind == WK_X) {
        const int src = thread_ok ? (lane ^ xm) : lane;
#pragma unroll
        for (int s = 0; s < S; s++)
            if (ALL || ((c_reg >> s) & 1u))
                v[s] = make_double2(__shfl_sync(0xffffffffu, v[s].x, src), __shfl_sync(0xffffffffu, v[s].y, src));
        return;
    }
    if (kind == WK_REAL) {
        double cA = hi ? m[3] : m[0], cB = hi ? m[2] : m[1];
        if (!thread_ok) { cA = 1.0; cB = 0.0; }
#pragma unroll
        for (int s = 0; s < S; s++) {
            if (ALL || ((c_reg >> s) & 1u)) {
                const amp_t mine = v[s];
                const amp_t other = shfl_xor_amp(mine, xm);
                v[s] = make_double2(cA * mine.x + cB * other.x, cA * mine.y + cB * other.y);
            }
        }
        return;
    }
    if (kind == WK_RX || kind == WK_RXS) {
        // RX: mine' = c*mine - i*s*other.  RXS (inputs swapped): mine' = c*other - i*s*mine
        const bool swapped = (kind == WK_RXS) && thread_ok;
        double c = m[0], sn = m[1];
        if (!thread_ok) { c = 1.0; sn = 0.0; }
#pragma unroll
        for (int s = 0; s < S; s++) {
            if (ALL || ((c_reg >> s) & 1u)) {
                const amp_t mine = v[s];
                const amp_t other = shfl_xor_amp(mine, xm);
                const amp_t P = swapped ? other : mine, Q = swapped ? mine : other;
                v[s] = make_double2(c * P.x + sn * Q.y, c * P.y - sn * Q.x);
            }
        }
        return;
    }
    if (!U2K) return;
    amp_t cA = hi ? make_double2(m[6], m[7]) : make_double2(m[0], m[1]);
    amp_t cB = hi ? make_double2(m[4], m[5]) : make_double2(m[2], m[3]);
    if (!thread_ok) { cA = make_double2(1.0, 0.0); cB = make_double2(0.0, 0.0); }
#pragma unroll
    for (int s = 0; s < S; s++) {
        if (ALL || ((c_reg >> s) & 1u)) {
            const amp_t mine = v[s];
            const amp_t other = shfl_xor_amp(mine, xm);
            v[s] = cadd(cmul(cA, mine), cmul(cB, other));
        }
    }
}

template <int R, bool U2K, bool LEAN>
__device__ __forceinline__ void lane_pair_op(amp_t (&v)[1 << R], uint32_t kind, uint32_t tpos, uint32_t c_reg, bool thread_ok,
                                             const double* __restrict__ m, int lane) {
    if (c_reg == kAllSlots<R>) lane_pair_impl<R, true, U2K, LEAN>(v, kind, tpos, c_reg, thread_ok, m, lane);
    else lane_pair_impl<R, false, U2K, LEAN>(v, kind, tpos, c_reg, thread_ok, m, lane);
}

// ---- the op program on one register tile -------------------------------------------------------------
// LANES = the program contains pair gates on lane qubits; programs without them run an instantiation that does not
// carry the shuffle code at all (smaller, fewer live registers).
// one device op on a register tile (the tile predicate has been checked by the caller)
// NT = entries of a phase table's lane part: 32 (k_window: lanes) or 128 (k_tile: the thread index of the round's layout)
template <int R, bool LANES, bool U2K, bool LEAN, int NT>
__device__ __forceinline__ void exec_op(amp_t (&v)[1 << R], const uint64_t tile, const int lane, const DOp& op, const amp_t* __restrict__ tables) {
    constexpr int S = 1 << R;
    const bool thread_ok = ((uint32_t)lane & op.c_lane) == op.c_lval;
    const uint32_t kind = op.kind, c_reg = op.c_reg, tpos = op.tpos;

    if (LANES && kind <= kLastPairKind && tpos < 5) {             // pair gate across lanes
        lane_pair_op<R, U2K, LEAN>(v, kind, tpos, c_reg, thread_ok, op.m, lane);
        return;
    }
    if (!thread_ok) return;                                     // lane-bit controls: skip at op granularity
    if (kind == WK_DIAG) {
        const amp_t ph = make_double2(op.m[0], op.m[1]);
#pragma unroll
        for (int s = 0; s < S; s++)
            if ((c_reg >> s) & 1u) upd_cmul(v[s], ph);
    } else if (kind == WK_RZ) {
        const amp_t p0 = make_double2(op.m[0], op.m[1]), p1 = make_double2(op.m[2], op.m[3]);
        const uint32_t t_reg = op.t_reg;
        if (t_reg == 0 && c_reg == kAllSlots<R>) {                 // target outside the registers: one phase per thread
            const bool t_thread = ((tile & op.t_tile) != 0) || (((uint32_t)lane & op.t_lane) != 0);
            const amp_t pt = t_thread ? p1 : p0;
#pragma unroll
            for (int s = 0; s < S; s++) upd_cmul(v[s], pt);
        } else {
            const bool t_thread = ((tile & op.t_tile) != 0) || (((uint32_t)lane & op.t_lane) != 0);
            const amp_t pt = t_thread ? p1 : p0;
#pragma unroll
            for (int s = 0; s < S; s++)
                if ((c_reg >> s) & 1u) {
                    if (s & t_reg) upd_cmul(v[s], p1);
                    else upd_cmul(v[s], pt);
                }
        }
    } else if (kind == WK_TABLE) {
        const amp_t* __restrict__ tab = tables + (uint64_t)__double_as_longlong(op.m[0]);
        const uint32_t hub_cls = op.hub_cls, hub_bit = op.hub_bit;
        if (hub_cls == CLS_TILE && !((tile >> hub_bit) & 1)) return;
        if (hub_cls == CLS_LANE && !((lane >> hub_bit) & 1)) return;
        amp_t f = tab[lane];                                        // lane table (NT entries)
        const uint32_t nch = op.nchunks;
        for (uint32_t k = 0; k < nch; k++)                          // tile chunk tables (256 entries each)
            f = cmul(f, __ldg(tab + NT + S + 256 * k + ((tile >> (8 * k)) & 255)));
        const uint32_t hub_slot = hub_cls == CLS_REG ? (1u << hub_bit) : 0u;
        if (op.has_reg) {
#pragma unroll
            for (int s = 0; s < S; s++)
                if ((s & hub_slot) == hub_slot) upd_cmul(v[s], cmul(f, __ldg(tab + NT + s)));
        } else {
#pragma unroll
            for (int s = 0; s < S; s++)
                if ((s & hub_slot) == hub_slot) upd_cmul(v[s], f);
        }
    } else if (kind == WK_NEG) {
#pragma unroll
        for (int s = 0; s < S; s++)
            if ((c_reg >> s) & 1u) { v[s].x = -v[s].x; v[s].y = -v[s].y; }
    } else if (LEAN && kind == WK_SCALE) {                          // the product of the pass's deferred gate scales
        const double g = op.m[0];
#pragma unroll
        for (int s = 0; s < S; s++) upd_scale(v[s], g);
    } else {
        switch (tpos - 5) {
            case 0: reg_pair_op<R, 0, U2K, LEAN>(v, kind, c_reg, op.m); break;
            case 1: reg_pair_op<R, (R > 1 ? 1 : 0), U2K, LEAN>(v, kind, c_reg, op.m); break;
            case 2: reg_pair_op<R, (R > 2 ? 2 : 0), U2K, LEAN>(v, kind, c_reg, op.m); break;
            case 3: reg_pair_op<R, (R > 3 ? 3 : 0), U2K, LEAN>(v, kind, c_reg, op.m); break;
            default: reg_pair_op<R, (R > 4 ? 4 : 0), U2K, LEAN>(v, kind, c_reg, op.m); break;
        }
    }
}

template <int R, bool LANES, bool U2K, bool LEAN = false, int NT = 32>
__device__ __forceinline__ void run_ops(amp_t (&v)[1 << R], const uint64_t tile, const int lane, const DOp* __restrict__ ops, const uint32_t nops,
                                        const amp_t* __restrict__ tables) {
#pragma unroll 1
    for (uint32_t o = 0; o < nops; o++) {
        const DOp& op = ops[o];
        if ((tile & op.c_tile) != op.c_tval) continue;                // warp-uniform control (positive and negative bits)
        exec_op<R, LANES, U2K, LEAN, NT>(v, tile, lane, op, tables);
    }
}

// ---- k_tile: the op set of the CTA-tile kernel -------------------------------------------------------------------
// On the CTA-tile kernel a pass is compute-bound, so the op loop is built around what ptxas does with 64 live amplitude
// registers in a loop (measured on SASS and with ncu, profiles/r02_*):
//  * a 2x2 written as  n0 = k0 a0 + k1 a1, n1 = k2 a0 + k3 a1  leaves its results in fresh registers, and every path through
//    the loop body then ends with ~64 moves back to the canonical assignment -- and ONE such path (a register swap, a
//    complex multiply) makes ptxas shuffle the whole file at every merge point of the loop: 4 x 64 moves per op.
//  * so EVERY op of this kernel is in place instruction by instruction -- each instruction overwrites the one operand
//    that dies there:
//      real 2x2 (LIFTING):  a0 <- k0 a0;  a0 <- a0 + k1 a1;  a1 <- k3' a1;  a1 <- a1 + k2' a0     k2' = k2 / k0, k3' = det / k0
//                           (the second row acts on the NEW a0; same 8 FP64 instructions per pair, no moves)
//      RX:                  the same with the off-diagonal products crossing x and y
//      phase e^{i phi}:     three shears  x <- x - t y;  y <- y + s x;  x <- x - t y   (t = tan(phi/2), s = sin phi; 3 FMA, not 4)
//      X / CNOT:            xor swaps (PTX, so that they stay three xors and never become a register renaming)
//  * the divisions are done on the host; a 2x2 whose pivot |k0| is below 0.05 is lowered as X followed by the 2x2 with its
//    columns swapped (a unitary's other pivot is then ~1), RX likewise through RX(theta) = -i X RX(theta - pi); a complex
//    2x2 is lowered to phase . real rotation . phase.  Rounding stays within ~20 ulp of the amplitudes an op touches.
//  * pair gates never see a register-bit control: the round builder keeps the controls of a controlled gate out of the
//    round's register qubits (they become thread or tile predicates); only X has variants under one register-bit control.
//   codes: 1..4 REALL on register bit B; 5..8 RXL; 9..12 X; 13..36 X under one register-bit control (C, CV);
//          40 TABLE, 41 NEG, 42 DIAG, 43 RZ, 44 SCALE.
static const double kLiftMinPivot = 0.05;
enum { FC_TABLE = 40, FC_NEG, FC_DIAG, FC_RZ, FC_SCALE, FC_REALUP = 45 /* + B */, FC_REALUM = 49 /* + B */, FC_RXU = 53 /* + B */ };
static inline int fast_code(int kind, int B, uint32_t pos, uint32_t neg) {
    if ((kind == WK_REALUP || kind == WK_REALUM || kind == WK_RXU) && B >= 0 && B <= 3 && (pos | neg) == 0)
        return (kind == WK_REALUP ? FC_REALUP : kind == WK_REALUM ? FC_REALUM : FC_RXU) + B;
    if (kind == WK_TABLE) return FC_TABLE;
    if (kind == WK_NEG) return FC_NEG;
    if (kind == WK_DIAG) return FC_DIAG;
    if (kind == WK_RZ) return FC_RZ;
    if (kind == WK_SCALE) return FC_SCALE;
    if (B < 0 || B > 3 || ((pos | neg) >> B) & 1u) return 0;
    if ((pos | neg) == 0) return kind == WK_REALL ? 1 + B : kind == WK_RXL ? 5 + B : kind == WK_X ? 9 + B : 0;
    if (kind != WK_X || __builtin_popcount(pos | neg) != 1) return 0;
    const int C = __builtin_ctz(pos | neg), CI = C < B ? C : C - 1, CV = pos ? 1 : 0;
    return 13 + (B * 3 + CI) * 2 + CV;
}

template <int B>
__device__ __forceinline__ void fast_reall(amp_t (&v)[16], const double k0, const double k1, const double k2, const double k3) {
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int s0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1)), s1 = s0 | (1 << B);
        v[s0].x *= k0; v[s0].x = fma(k1, v[s1].x, v[s0].x);
        v[s0].y *= k0; v[s0].y = fma(k1, v[s1].y, v[s0].y);
        v[s1].x *= k3; v[s1].x = fma(k2, v[s0].x, v[s1].x);
        v[s1].y *= k3; v[s1].y = fma(k2, v[s0].y, v[s1].y);
    }
}
// RX = [[c, -i s], [-i s, c]]:  a0' = c a0 - i s a1;  a1' = (1 / c) a1 - i (s / c) a0'      (k = c, s, 1 / c, s / c)
template <int B>
__device__ __forceinline__ void fast_rxl(amp_t (&v)[16], const double k0, const double k1, const double k2, const double k3) {
    const double n1 = -k1, n3 = -k3;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int s0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1)), s1 = s0 | (1 << B);
        v[s0].x *= k0; v[s0].x = fma(k1, v[s1].y, v[s0].x);
        v[s0].y *= k0; v[s0].y = fma(n1, v[s1].x, v[s0].y);
        v[s1].x *= k2; v[s1].x = fma(k3, v[s0].y, v[s1].x);
        v[s1].y *= k2; v[s1].y = fma(n3, v[s0].x, v[s1].y);
    }
}
// UNIT forms (uncontrolled gates only; the omitted factor g of every such gate of a launch is multiplied on the host and applied
// by one WK_SCALE op at the end of the launch): half the FP64 instructions of the lifted forms.
//   REALUP / REALUM = [[1, p], [q, +-1]]:  a0' = a0 + p a1,  a1' = q a0 +- a1        (H / g: p = q = 1, minus)
template <int B, bool MINUS>
__device__ __forceinline__ void fast_realu(amp_t (&v)[16], const double p, const double q) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int s0 = ((i >> B) << (B + 1)) | (i & ((1 << B) - 1)), s1 = s0 | (1 << B);
        const double tx = v[s0].x, ty = v[s0].y;
        v[s0].x = fma(p, v[s1].x, tx);
        v[s0].y = fma(p, v[s1].y, ty);
        v[s1].x = fma(q, tx, MINUS ? -v[s1].x : v[s1].x);
        v[s1].y = fma(q, ty, MINUS ? -v[s1].y : v[s1].y);
    }
}
//   RXU = [[1, -i t], [-i t, 1]]:  a0' = a0 - i t a1,  a1' = a1 - i t a0
template <int B>
__device__ __forceinline__ void fast_rxu(amp_t (&v)[16], const double t) {
    const double nt = -t;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int s0 = ((i >> B) << (B + 1)) | (i & ((1 << B) - 1)), s1 = s0 | (1 << B);
        const double tx = v[s0].x, ty = v[s0].y;
        v[s0].x = fma(t, v[s1].y, tx);
        v[s0].y = fma(nt, v[s1].x, ty);
        v[s1].x = fma(t, ty, v[s1].x);
        v[s1].y = fma(nt, tx, v[s1].y);
    }
}
__device__ __forceinline__ void xor_swap(amp_t& a, amp_t& b) {
    asm volatile("xor.b64 %0, %0, %1;\n\txor.b64 %1, %1, %0;\n\txor.b64 %0, %0, %1;" : "+d"(a.x), "+d"(b.x));
    asm volatile("xor.b64 %0, %0, %1;\n\txor.b64 %1, %1, %0;\n\txor.b64 %0, %0, %1;" : "+d"(a.y), "+d"(b.y));
}
template <int B>
__device__ __forceinline__ void fast_x(amp_t (&v)[16]) {
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int s0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        xor_swap(v[s0], v[s0 | (1 << B)]);
    }
}
template <int B, int CI, int CV>
__device__ __forceinline__ void fast_cx(amp_t (&v)[16]) {
    constexpr int C = CI < B ? CI : CI + 1;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int s0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1));
        if (((s0 >> C) & 1) == CV) xor_swap(v[s0], v[s0 | (1 << B)]);
    }
}

// ---- in-place phase multiplication ----------------------------------------------------------------------------
// a <- e^{i phi} a is a rotation of (x, y); as three shears it is in place instruction by instruction and costs 3 FMA instead
// of 4 mul/FMA plus the register moves of the plain complex product:
//     x <- x - t y;  y <- y + s x;  x <- x - t y        t = tan(phi / 2) = fy / (1 + fx),  s = sin phi = fy
// For fx < 0 the amplitude is negated first (sign flips, in place) and the rotation uses -f, so 1 + fx >= 1 always.
// Every diagonal factor of a unitary circuit has modulus one (phases of Z/S/T/P/RZ and their products).
struct Rot { double nt, s; bool neg; };
__host__ __device__ __forceinline__ Rot make_rot(double fx, double fy) {
    Rot r;
    r.neg = fx < 0.0;
    if (r.neg) { fx = -fx; fy = -fy; }
    r.nt = -fy / (1.0 + fx);
    r.s = fy;
    return r;
}
// x <- -x where `mask` is 0x80000000, x where it is 0: one xor on the high word (ALU pipe, in place; no FP64 issue slot)
__device__ __forceinline__ void flip_sign(double& x, uint32_t mask) {
    x = __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}
__device__ __forceinline__ void rot_inplace(amp_t& a, const Rot& r) {
    const uint32_t m = r.neg ? 0x80000000u : 0u;
    flip_sign(a.x, m);
    flip_sign(a.y, m);
    a.x = fma(r.nt, a.y, a.x);
    a.y = fma(r.s, a.x, a.y);
    a.x = fma(r.nt, a.y, a.x);
}
// host-side packing of a rotation into one amp_t: (nt with the negate flag in its lowest mantissa bit, s)
static inline amp_t pack_rot(amp_t f) {
    Rot r = make_rot(f.x, f.y);
    uint64_t bits;
    memcpy(&bits, &r.nt, 8);
    bits = (bits & ~1ull) | (r.neg ? 1ull : 0ull);
    memcpy(&r.nt, &bits, 8);
    return make_double2(r.nt, r.s);
}
__device__ __forceinline__ Rot unpack_rot(double nt, double s) {
    Rot r;
    r.nt = nt; r.s = s;
    r.neg = (__double2loint(nt) & 1) != 0;
    return r;
}

#define QI_CX6(B) \
    case 13 + (B * 3 + 0) * 2 + 0: fast_cx<B, 0, 0>(v); break; case 13 + (B * 3 + 0) * 2 + 1: fast_cx<B, 0, 1>(v); break; \
    case 13 + (B * 3 + 1) * 2 + 0: fast_cx<B, 1, 0>(v); break; case 13 + (B * 3 + 1) * 2 + 1: fast_cx<B, 1, 1>(v); break; \
    case 13 + (B * 3 + 2) * 2 + 0: fast_cx<B, 2, 0>(v); break; case 13 + (B * 3 + 2) * 2 + 1: fast_cx<B, 2, 1>(v); break;

// the op program of one round on the CTA-tile kernel
__device__ __forceinline__ void run_ops_tile(amp_t (&v)[16], const uint64_t tile, const int t, const DOp* __restrict__ ops, const uint32_t nops,
                                             const amp_t* __restrict__ tables) {
    constexpr int S = 16, NT = kTileThreads;
    if (nops == 0) return;
    // everything the dispatch and the common routines need is loaded one op AHEAD (warp-uniform loads: uniform registers), so
    // the constant-bank latency of op o+1 hides behind the arithmetic of op o instead of heading a dependent chain per op
    uint4 nh = *reinterpret_cast<const uint4*>(ops);                                   // kind..code | c_lane | c_reg
    ulonglong2 nt2 = *reinterpret_cast<const ulonglong2*>(reinterpret_cast<const char*>(ops) + 16);   // c_tile | c_tval
    double nk0 = ops[0].m[0], nk1 = ops[0].m[1], nk2 = ops[0].m[2], nk3 = ops[0].m[3];
#pragma unroll 1
    for (uint32_t o = 0; o < nops; o++) {
        const DOp& op = ops[o];
        const uint4 h = nh;
        const ulonglong2 tl = nt2;
        const double k0 = nk0, k1 = nk1, k2 = nk2, k3 = nk3;
        if (o + 1 < nops) {
            const DOp& nx = ops[o + 1];
            nh = *reinterpret_cast<const uint4*>(&nx);
            nt2 = *reinterpret_cast<const ulonglong2*>(reinterpret_cast<const char*>(&nx) + 16);
            nk0 = nx.m[0]; nk1 = nx.m[1]; nk2 = nx.m[2]; nk3 = nx.m[3];
        }
        // tile-uniform controls and thread-bit controls: an op that does not apply becomes the empty case
        const bool on = (tile & tl.x) == tl.y && ((uint32_t)t & h.z) == ((h.y >> 16) & 0xffu);
        const uint32_t code = on ? (h.y >> 24) : 0u;
        const uint32_t c_reg = h.w;
        switch (code) {
            case 1: fast_reall<0>(v, k0, k1, k2, k3); break;
            case 2: fast_reall<1>(v, k0, k1, k2, k3); break;
            case 3: fast_reall<2>(v, k0, k1, k2, k3); break;
            case 4: fast_reall<3>(v, k0, k1, k2, k3); break;
            case 5: fast_rxl<0>(v, k0, k1, k2, k3); break;
            case 6: fast_rxl<1>(v, k0, k1, k2, k3); break;
            case 7: fast_rxl<2>(v, k0, k1, k2, k3); break;
            case 8: fast_rxl<3>(v, k0, k1, k2, k3); break;
            case 9: fast_x<0>(v); break;
            case 10: fast_x<1>(v); break;
            case 11: fast_x<2>(v); break;
            case 12: fast_x<3>(v); break;
            QI_CX6(0) QI_CX6(1) QI_CX6(2) QI_CX6(3)
            case FC_REALUP + 0: fast_realu<0, false>(v, k0, k1); break;
            case FC_REALUP + 1: fast_realu<1, false>(v, k0, k1); break;
            case FC_REALUP + 2: fast_realu<2, false>(v, k0, k1); break;
            case FC_REALUP + 3: fast_realu<3, false>(v, k0, k1); break;
            case FC_REALUM + 0: fast_realu<0, true>(v, k0, k1); break;
            case FC_REALUM + 1: fast_realu<1, true>(v, k0, k1); break;
            case FC_REALUM + 2: fast_realu<2, true>(v, k0, k1); break;
            case FC_REALUM + 3: fast_realu<3, true>(v, k0, k1); break;
            case FC_RXU + 0: fast_rxu<0>(v, k0); break;
            case FC_RXU + 1: fast_rxu<1>(v, k0); break;
            case FC_RXU + 2: fast_rxu<2>(v, k0); break;
            case FC_RXU + 3: fast_rxu<3>(v, k0); break;
            case FC_TABLE: {
                const amp_t* __restrict__ tab = tables + (uint64_t)__double_as_longlong(op.m[0]);
                const uint32_t hub_cls = op.hub_cls, hub_bit = op.hub_bit;
                if (hub_cls == CLS_TILE && !((tile >> hub_bit) & 1)) break;
                if (hub_cls == CLS_LANE && !((t >> hub_bit) & 1)) break;
                amp_t f = tab[t];                                             // thread table (128 entries)
                const uint32_t nch = op.nchunks;
                for (uint32_t k = 0; k < nch; k++)                            // tile chunk tables (256 entries each)
                    f = cmul(f, __ldg(tab + NT + S + 256 * k + ((tile >> (8 * k)) & 255)));
                // the thread's factor as a plain complex product (2 DMUL + 2 DFMA per amplitude, written with explicit roundings
                // so that the JIT modules, which emit the same four instructions, stay bit-identical): the three-shear form
                // needs one division per thread and op (a subroutine call in straight-line code) and two sign flips per amplitude
                const uint32_t hub_slot = hub_cls == CLS_REG ? (1u << hub_bit) : 0u;
#pragma unroll
                for (int s = 0; s < S; s++)
                    if ((s & hub_slot) == hub_slot) {
                        const double t0 = __dmul_rn(f.y, v[s].y), t1 = __dmul_rn(f.y, v[s].x);
                        v[s].x = __fma_rn(f.x, v[s].x, -t0);
                        v[s].y = __fma_rn(f.x, v[s].y, t1);
                    }
                if (op.has_reg) {                                             // slot factors: packed rotations behind the chunk tables
                    const amp_t* __restrict__ sl = tab + NT + S + 256 * nch;
#pragma unroll
                    for (int s = 0; s < S; s++)
                        if ((s & hub_slot) == hub_slot) { const amp_t e = __ldg(sl + s); rot_inplace(v[s], unpack_rot(e.x, e.y)); }
                }
                break;
            }
            case FC_NEG:
#pragma unroll
                for (int s = 0; s < S; s++) {
                    const uint32_t m = (c_reg << (31 - s)) & 0x80000000u;      // slot s selected: flip, else keep
                    flip_sign(v[s].x, m);
                    flip_sign(v[s].y, m);
                }
                break;
            case FC_DIAG: {
                const Rot r = unpack_rot(k2, k3);
#pragma unroll
                for (int s = 0; s < S; s++)
                    if ((c_reg >> s) & 1u) rot_inplace(v[s], r);
                break;
            }
            case FC_RZ: {
                const Rot r0 = unpack_rot(op.m[4], op.m[5]), r1 = unpack_rot(op.m[6], op.m[7]);
                const uint32_t t_reg = op.t_reg;
                const bool t_thread = ((tile & op.t_tile) != 0) || (((uint32_t)t & op.t_lane) != 0);
                Rot rt;
                rt.nt = t_thread ? r1.nt : r0.nt; rt.s = t_thread ? r1.s : r0.s; rt.neg = t_thread ? r1.neg : r0.neg;
#pragma unroll
                for (int s = 0; s < S; s++)
                    if ((c_reg >> s) & 1u) {
                        if (s & t_reg) rot_inplace(v[s], r1);
                        else rot_inplace(v[s], rt);
                    }
                break;
            }
            case FC_SCALE:
                if (k1 == 0.0) {
#pragma unroll
                    for (int s = 0; s < S; s++) { v[s].x *= k0; v[s].y *= k0; }
                } else {               // complex scalar (k0, k1): the global phases the P-form diagonal ops left out
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const double t0 = __dmul_rn(k1, v[s].y), t1 = __dmul_rn(k1, v[s].x);
                        v[s].x = __fma_rn(k0, v[s].x, -t0);
                        v[s].y = __fma_rn(k0, v[s].y, t1);
                    }
                }
                break;
            default: break;
        }
    }
}

// resident blocks per SM the register budget is tuned for: 8 amplitudes/thread -> 8 blocks,
// 16 -> 4 blocks (128 registers), 32 -> 2 blocks
#ifndef QI_WINDOW_BLOCKS4
#define QI_WINDOW_BLOCKS4 4
#endif
#define QI_WINDOW_BLOCKS(R) ((R) <= 3 ? 8 : ((R) == 4 ? QI_WINDOW_BLOCKS4 : 2))

// direct variant: every thread loads its 2^R amplitudes itself (coalesced 512 B per warp access)
template <int R, bool LANES, bool U2K, bool LEAN = false>
__global__ void __launch_bounds__(128, QI_WINDOW_BLOCKS(R)) k_window(amp_t* __restrict__ a, uint64_t ntiles, const __grid_constant__ WProgram<R> P) {
    constexpr int S = 1 << R;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t tile = warp; tile < ntiles; tile += nwarps) {
        const uint64_t base = expand_index((tile << 5) | (uint64_t)lane, P.ins);
        amp_t v[S];
#pragma unroll
        for (int s = 0; s < S; s++) v[s] = QI_LD(a + base + P.off[s]);
        if (LEAN && (P.flags & 1u) && (lane & 7) == 0 && tile + nwarps < ntiles) {
            // experimental (unmeasured): pull the warp's NEXT tile into L2 while this one is computed -- costs no registers,
            // unlike a second tile in flight; one prefetch per 128-byte line (lanes 0, 8, 16, 24)
            const uint64_t nbase = expand_index(((tile + nwarps) << 5) | (uint64_t)lane, P.ins);
#pragma unroll
            for (int s = 0; s < S; s++) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + nbase + P.off[s]));
        }
        run_ops<R, LANES, U2K, LEAN>(v, tile, lane, P.ops, P.nops, P.tables);
#pragma unroll
        for (int s = 0; s < S; s++) QI_ST(a + base + P.off[s], v[s]);
    }
}

// ---- TMA-prefetched variant ------------------------------------------------------------------------------
// Persistent grid (one block slot per SM and resident block).  Every warp owns an 8 KiB (2^R x 512 B) staging
// buffer in shared memory and one mbarrier: while it works on tile t in registers, the 2^R bulk copies
// (cp.async.bulk, 512 B each, L2 evict-first) of its NEXT tile are already in flight, so HBM stays busy during
// the compute phase of a fused pass instead of only during the load phase of whichever warps happen to be there.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "QI_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra QI_DONE;\n\t"
        "bra QI_WAIT;\n\t"
        "QI_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load_evict_first(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(0x12F0000000000000ull)
                 : "memory");
}

template <int R, bool LANES, bool U2K>
__global__ void __launch_bounds__(128, QI_WINDOW_BLOCKS(R)) k_window_tma(amp_t* __restrict__ a, uint64_t ntiles, const __grid_constant__ WProgram<R> P) {
    constexpr int S = 1 << R;
    extern __shared__ __align__(128) unsigned char smem_raw[];     // [4 warps][S * 32 amplitudes] then 4 mbarriers
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    amp_t* buf = reinterpret_cast<amp_t*>(smem_raw) + w * (S * 32);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 4 * S * 32 * sizeof(amp_t)) + w;
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint64_t warp = (uint64_t)blockIdx.x * 4 + w;
    const uint64_t nwarps = (uint64_t)gridDim.x * 4;
    // lane s (< 2^R) fetches slot s of the tile: 32 consecutive amplitudes = 512 B
    auto issue = [&](uint64_t tile) {
        if (lane == 0) mbar_expect_tx(bar, S * 512);
        __syncwarp();
        if (lane < S) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer was last read through the generic proxy
            bulk_load_evict_first(buf + lane * 32, a + expand_index(tile << 5, P.ins) + P.off[lane], 512, bar);
        }
    };
    uint64_t tile = warp;
    uint32_t parity = 0;
    if (tile < ntiles) issue(tile);
    for (; tile < ntiles; tile += nwarps) {
        mbar_wait(bar, parity);
        parity ^= 1;
        amp_t v[S];
#pragma unroll
        for (int s = 0; s < S; s++) v[s] = buf[s * 32 + lane];
        __syncwarp();
        if (tile + nwarps < ntiles) issue(tile + nwarps);
        run_ops<R, LANES, U2K>(v, tile, lane, P.ops, P.nops, P.tables);
        const uint64_t base = expand_index((tile << 5) | (uint64_t)lane, P.ins);
#pragma unroll
        for (int s = 0; s < S; s++) QI_ST(a + base + P.off[s], v[s]);
    }
}

// ---- CTA-tile variant (k_tile): two-level pass --------------------------------------------------------------
// The warp-tile kernel above holds 9 qubits per pass (5 lane qubits whose gates cost shuffles + 4 register qubits), so a
// deep circuit needs one HBM pass per ~10 gates, and the passes that carry lane-qubit gates run at 2-5x the HBM floor.
// k_tile makes the unit of a pass a CTA TILE of 2^11 amplitudes: physical qubits 0..4 (so that every global access is a
// coalesced 512-byte row) plus SIX window qubits chosen per pass.  128 threads hold 16 amplitudes each in registers and
// work in ROUNDS: in a round, 4 of the 11 tile qubits are register qubits (every gate on them pairs two registers of one
// thread, no data movement, no shuffles, low qubits included); between rounds the registers are regrouped through a
// 32 KiB shared-memory image of the tile (one 16-byte store + one 16-byte load per amplitude, XOR-swizzled, ONE
// __syncthreads per regroup: a thread writes the locations it owns under the current layout -- the ones it read last --
// and reads the ones it owns under the next layout, so consecutive regroups cannot race).  The first and the last round
// use the IO layout (register qubits = 4 of the window qubits), the only one whose global accesses coalesce.
// Measured on the prototype (tools/micro/tile_proto.cu, profiles/r02_tile_proto_microbench.txt): a regroup costs ~0.3 ms
// at 30 qubits against 5.5 ms for the HBM pass it replaces; a register gate costs 0.15-0.2 ms (the FP64 pipe).

struct TRound {                 // 64 bytes
    uint16_t sswz[16];          // swizzled tile-local offset of slot s (register bits of this round)
    uint8_t thr_pos[kTileThrBits];   // tile-local bit of thread-index bit k
    uint8_t pad;
    uint16_t first_op, nops;
    uint32_t pad2[5];
};
static_assert(sizeof(TRound) == 64, "TRound layout");

struct TProgram {
    uint8_t tpos[kTileBits];    // physical position of tile-local bit j (ascending; tpos[0..4] = 0..4)
    uint8_t nrounds;
    uint8_t tpos_out[kTileBits];   // position the content of local bit j is stored to (a permutation of tpos: the pass may leave
    uint8_t pad0;                  // its 11 qubits in any order -- the five the next pass wants at positions 0..4)
    uint32_t flags;
    uint64_t goff_in[16];       // slot -> global index offset under the first round's layout (loads)
    uint64_t goff_out[16];      // ... under the last round's layout (stores)
    const amp_t* tables;        // phase-table arena
    TRound rounds[kMaxRounds];
    DOp ops[kMaxTileOps];
};
static_assert(sizeof(TProgram) <= 32000, "kernel parameter space");

// XOR-fold of the 16-byte-unit index: a quarter warp (8 lanes, thread bits 0..2) hits 8 distinct 16-byte bank groups whenever
// the three tile-local bits behind thread bits 0..2 fall into three different classes mod 3 (the host orders them so)
__host__ __device__ __forceinline__ uint32_t tile_swz(uint32_t j) { return j ^ ((j >> 3) & 7u) ^ ((j >> 6) & 7u) ^ ((j >> 9) & 3u); }

__global__ void __launch_bounds__(kTileThreads, 4) k_tile(amp_t* __restrict__ a, uint64_t ntiles, const __grid_constant__ TProgram P) {
    __shared__ __align__(16) amp_t sm[1 << kTileBits];
    __shared__ uint16_t lbs[kMaxRounds][kTileThreads];       // swizzled tile-local base of thread t in round r
    const int t = threadIdx.x;
    const int nr = P.nrounds;
    for (int r = 0; r < nr; r++) {
        uint32_t b = 0;
#pragma unroll
        for (int k = 0; k < kTileThrBits; k++) b |= ((t >> k) & 1u) << P.rounds[r].thr_pos[k];
        lbs[r][t] = (uint16_t)tile_swz(b);
    }
    // IO layouts (first round: loads, last round: stores): thread bit k <-> tile-local bit thr_pos[k] <-> physical tpos[...]
    uint64_t g_in = 0, g_out = 0;
#pragma unroll
    for (int k = 0; k < kTileThrBits; k++) {
        g_in |= (uint64_t)((t >> k) & 1u) << P.tpos[P.rounds[0].thr_pos[k]];
        g_out |= (uint64_t)((t >> k) & 1u) << P.tpos_out[P.rounds[nr - 1].thr_pos[k]];
    }
    __syncthreads();
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint64_t tb = tile;
#pragma unroll 1
        for (int j = 0; j < kTileBits; j++) tb = insert_zero(tb, P.tpos[j]);
        amp_t v[16];
        {
            const amp_t* __restrict__ g = a + tb + g_in;
#pragma unroll
            for (int s = 0; s < 16; s++) v[s] = QI_LD(g + P.goff_in[s]);
        }
        run_ops_tile(v, tile, t, P.ops + P.rounds[0].first_op, P.rounds[0].nops, P.tables);
#pragma unroll 1
        for (int r = 1; r < nr; r++) {
            const TRound& prev = P.rounds[r - 1];
            const TRound& cur = P.rounds[r];
            const uint32_t wb = lbs[r - 1][t], rb = lbs[r][t];
#pragma unroll
            for (int s = 0; s < 16; s++) sm[wb ^ prev.sswz[s]] = v[s];
            __syncthreads();
#pragma unroll
            for (int s = 0; s < 16; s++) v[s] = sm[rb ^ cur.sswz[s]];
            run_ops_tile(v, tile, t, P.ops + cur.first_op, cur.nops, P.tables);
        }
        {
            amp_t* __restrict__ g = a + tb + g_out;
#pragma unroll
            for (int s = 0; s < 16; s++) QI_ST(g + P.goff_out[s], v[s]);
        }
    }
}

// ---- host: scheduling ------------------------------------------------------------------------------
static const int kDefaultR = 4;     // register qubits per pass (option "window_regs": 3, 4 or 5)

struct GateUse {
    uint64_t n_use;   // qubits used non-diagonally (targets of H/X/Y/U2/SWAP)
    uint64_t d_use;   // qubits used diagonally (controls, targets of phase gates)
};

static GateUse uses_of(const PhysGate& g) {
    GateUse u{0, g.cmask};
    switch (g.kind) {
        case IK_H: case IK_X: case IK_Y: case IK_U2: u.n_use = 1ull << g.t0; break;
        case IK_SWAP: case IK_MATCH: u.n_use = (1ull << g.t0) | (1ull << g.t1); break;
        case IK_DIAG: case IK_RZ: if (g.t0 >= 0) u.d_use |= 1ull << g.t0; break;
        default: break;
    }
    return u;
}

// host-side op with PHYSICAL masks; lowered to a DOp once the pass's window is final
struct HOp {
    uint32_t kind = 0;
    int target = -1;            // pair ops: physical target
    uint64_t cmask = 0;         // physical bits that must be 1 (WK_DIAG: includes the target)
    uint64_t nmask = 0;         // physical bits that must be 0 (the complement half of an absorbed CNOT)
    uint64_t tmask = 0;         // WK_RZ: physical target bit
    double m[8] = {0};
    int group = -1;             // WK_TABLE: index into Pass::groups
    int pair = -1;              // the two halves of an absorbed CNOT share an id (lean lowering treats them as one gate)
};

// a mergeable diagonal group: [hub set] * prod_j (bit_j ? f1_j : f0_j)
struct DiagGroup {
    int hub = -1;                        // physical hub qubit; -1 = unconditional
    int hub_alt = -1;                    // second hub candidate while only one 2-qubit member is present
    std::vector<int> bits;               // member qubits
    std::vector<amp_t> f0, f1;           // factor when the member bit is 0 / 1
    uint64_t blocked_since = 0;          // qubits used non-diagonally by ops appended after this group
    size_t op_index = 0;                 // position of the group's op in Pass::ops
    int members = 0;                     // number of merged gates
};

struct Pass {
    std::vector<int> regs;               // window qubits (physical positions >= 5)
    std::vector<HOp> ops;
    std::vector<DiagGroup> groups;
    int next_pair = 0;
    bool absorb = true;                  // fold CNOTs into neighbouring gates (k_window); tile passes keep them as register swaps
};

static void classify_u2(const double* p, HOp* op) {
    const bool rx_form = p[1] == 0.0 && p[2] == 0.0 && p[4] == 0.0 && p[7] == 0.0 && p[0] == p[6] && p[3] == p[5];
    const bool real_form = p[1] == 0.0 && p[3] == 0.0 && p[5] == 0.0 && p[7] == 0.0;
    memset(op->m, 0, sizeof(op->m));
    if (real_form) { op->kind = WK_REAL; op->m[0] = p[0]; op->m[1] = p[2]; op->m[2] = p[4]; op->m[3] = p[6]; }
    else if (rx_form) { op->kind = WK_RX; op->m[0] = p[0]; op->m[1] = -p[3]; }
    else { op->kind = WK_U2; memcpy(op->m, p, 8 * sizeof(double)); }
}

static int window_regs(const qi_state* s) {
    int r = ctx().opt_window_regs ? ctx().opt_window_regs : kDefaultR;
    if (r < 3) r = 3;
    if (r > 5) r = 5;
    while (r > 3 && (int)s->n_local < kLaneQubits + r) r--;
    return r;
}

bool window_supported(const qi_state* s) {
    return s->consistent && (int)s->n_local >= kLaneQubits + 3;
}

static bool window_takes(const PhysGate& g) {
    return g.kind == IK_H || g.kind == IK_X || g.kind == IK_Y || g.kind == IK_U2 || g.kind == IK_DIAG || g.kind == IK_RZ ||
           g.kind == IK_SWAP;
}

static void append_op(Pass& ps, const HOp& op, uint64_t n_use) {
    ps.ops.push_back(op);
    for (DiagGroup& g : ps.groups) g.blocked_since |= n_use;
}

// ---- CNOT absorption --------------------------------------------------------------------------------
// A CNOT (X with one control c) next to an uncontrolled single-qubit gate G on the same target t costs a full
// register-swap op of its own.  Since (G X)(a0, a1) = G(a1, a0) and X only permutes, the pair is replaced by
//   [c = 1] G' = G X  (or X G)   and   [c = 0] G
// two half-populated ops that together do the work of ONE G.  Exact: X permutes, G' has the columns (rows) of
// G swapped.  Only controls on tile or register bits qualify (a lane-bit control would idle half the lanes in
// both ops), and nothing between the two gates may touch t or change c.
static bool absorbable_kind(uint32_t k) { return k == WK_REAL || k == WK_RX || k == WK_RXS || k == WK_U2; }

// G' = G o X (swap_inputs) or X o G (swap outputs); for the kinds above both are again one of the kinds
static void compose_with_x(HOp* g, bool x_first) {
    double* m = g->m;
    if (g->kind == WK_RX) { g->kind = WK_RXS; return; }          // RX is symmetric and persymmetric: RX X = X RX
    if (g->kind == WK_RXS) { g->kind = WK_RX; return; }
    if (g->kind == WK_REAL) {                                     // m = (k00, k01, k10, k11)
        if (x_first) { std::swap(m[0], m[1]); std::swap(m[2], m[3]); }      // columns
        else { std::swap(m[0], m[2]); std::swap(m[1], m[3]); }              // rows
        return;
    }
    // WK_U2: m = (m00, m01, m10, m11) as (re, im)
    if (x_first) { std::swap(m[0], m[2]); std::swap(m[1], m[3]); std::swap(m[4], m[6]); std::swap(m[5], m[7]); }
    else { std::swap(m[0], m[4]); std::swap(m[1], m[5]); std::swap(m[2], m[6]); std::swap(m[3], m[7]); }
}

// index of the last op of the pass that touches qubit t, or -1; `clean` = no op after it n-uses `protect`
static int last_touch(const Pass& ps, int t, uint64_t protect, bool* clean) {
    *clean = true;
    for (int i = (int)ps.ops.size() - 1; i >= 0; i--) {
        const HOp& h = ps.ops[i];
        if (h.kind == 0) continue;
        uint64_t n = 0, d = 0;
        if (h.kind == WK_TABLE) {
            const DiagGroup& g = ps.groups[h.group];
            if (g.hub >= 0) d |= 1ull << g.hub;
            if (g.hub_alt >= 0) d |= 1ull << g.hub_alt;
            for (int q : g.bits) d |= 1ull << q;
            d |= h.cmask | h.tmask;
        } else if (h.kind == WK_DIAG || h.kind == WK_RZ) d = h.cmask | h.tmask;
        else { n = 1ull << h.target; d = h.cmask | h.nmask; }
        if ((n | d) & (1ull << t)) return i;
        if (n & protect) *clean = false;
    }
    return -1;
}

// `g` (uncontrolled pair op on t) is about to be appended: if the last op on t is a single-control X, fold it in
static bool absorb_cnot_before(Pass& ps, HOp* g) {
    if (!absorbable_kind(g->kind) || g->cmask || g->nmask) return false;
    bool clean = false;
    const int i = last_touch(ps, g->target, 0, &clean);
    if (i < 0) return false;
    HOp& x = ps.ops[i];
    if (x.kind != WK_X || x.target != g->target || x.nmask || __builtin_popcountll(x.cmask) != 1) return false;
    if (x.cmask & ((1ull << kLaneQubits) - 1)) return false;
    // nothing after the X may have changed its control qubit
    bool ctrl_clean = false;
    (void)last_touch(ps, g->target, x.cmask, &ctrl_clean);
    if (!ctrl_clean) return false;
    const uint64_t c = x.cmask;
    HOp on = *g;                 // control = 1: G X, takes the X's place
    compose_with_x(&on, true);
    on.cmask = c;
    on.pair = g->pair = ps.next_pair++;
    x = on;
    g->nmask = c;                // control = 0: plain G, appended by the caller
    return true;
}

// `x` (single-control X on t) is about to be appended: if the last op on t is an uncontrolled G, fold the X in
static bool absorb_cnot_after(Pass& ps, const HOp& x, HOp* off_half) {
    if (x.kind != WK_X || x.nmask || __builtin_popcountll(x.cmask) != 1) return false;
    if (x.cmask & ((1ull << kLaneQubits) - 1)) return false;
    bool clean = false;
    const int i = last_touch(ps, x.target, x.cmask, &clean);
    if (i < 0 || !clean) return false;
    HOp& g = ps.ops[i];
    if (!absorbable_kind(g.kind) || g.target != x.target || g.cmask || g.nmask) return false;
    *off_half = g;               // control = 0: plain G, appended by the caller
    off_half->nmask = x.cmask;
    compose_with_x(&g, false);   // control = 1: X G, stays in G's place
    g.cmask = x.cmask;
    g.pair = off_half->pair = ps.next_pair++;
    return true;
}

static void push_pair(Pass& ps, uint32_t kind, int target, uint64_t cmask, const double* m8) {
    HOp op;
    op.kind = kind;
    op.target = target;
    op.cmask = cmask;
    if (m8) memcpy(op.m, m8, 8 * sizeof(double));
    if (ctx().opt_absorb && ps.absorb) {
        HOp off;
        if (absorb_cnot_before(ps, &op)) { append_op(ps, op, 1ull << target); return; }     // op now carries nmask
        if (absorb_cnot_after(ps, op, &off)) { append_op(ps, off, 1ull << target); return; }
    }
    append_op(ps, op, 1ull << target);
}

// try to merge a diagonal gate into an open group of this pass; returns false if it must be a plain op
static bool merge_diag(Pass& ps, const PhysGate& g) {
    // normal form: optional hub h, member bit j with factors (f0, f1)
    int h = -1, h2 = -1, j = -1;
    amp_t f0 = make_double2(1.0, 0.0), f1 = make_double2(1.0, 0.0);
    const int nc = __builtin_popcountll(g.cmask);
    if (g.t0 < 0 || nc > 1) return false;
    if (g.kind == IK_DIAG) {
        f1 = make_double2(g.p[0], g.p[1]);
        j = g.t0;
        if (nc == 1) { h = __builtin_ctzll(g.cmask); h2 = g.t0; }     // symmetric: either qubit can be the hub
    } else {   // IK_RZ
        f0 = make_double2(g.p[0], g.p[1]);
        f1 = make_double2(g.p[2], g.p[3]);
        j = g.t0;
        if (nc == 1) h = __builtin_ctzll(g.cmask);
    }
    const uint64_t qmask = (1ull << j) | (h >= 0 ? (1ull << h) : 0ull);
    for (int gi = (int)ps.groups.size() - 1; gi >= 0; gi--) {
        DiagGroup& grp = ps.groups[gi];
        if (grp.blocked_since & qmask) continue;
        int member = j;
        if (h < 0) { if (grp.hub >= 0 || grp.hub_alt >= 0) continue; }            // unconditional gate: unconditional group only
        else {
            if (grp.hub < 0 && grp.hub_alt < 0) continue;
            if (grp.hub_alt >= 0) {
                // the group holds one symmetric 2-qubit member {hub, hub_alt}: fix the hub now
                if (h == grp.hub || (h2 >= 0 && h2 == grp.hub)) { /* keep */ }
                else if (h == grp.hub_alt || (h2 >= 0 && h2 == grp.hub_alt)) {
                    std::swap(grp.hub, grp.hub_alt);
                    grp.bits[0] = grp.hub_alt;
                } else continue;
                grp.hub_alt = -1;
            }
            if (h == grp.hub) member = j;
            else if (h2 >= 0 && h2 == grp.hub) member = h;                      // symmetric gate seen from the other side
            else continue;
        }
        // merge factors if the member bit is already present
        bool found = false;
        for (size_t k = 0; k < grp.bits.size(); k++)
            if (grp.bits[k] == member) { grp.f0[k] = cmul(grp.f0[k], f0); grp.f1[k] = cmul(grp.f1[k], f1); found = true; break; }
        if (!found) { grp.bits.push_back(member); grp.f0.push_back(f0); grp.f1.push_back(f1); }
        grp.members++;
        return true;
    }
    // open a new group
    DiagGroup grp;
    grp.hub = h;
    grp.hub_alt = (g.kind == IK_DIAG && nc == 1) ? h2 : -1;
    grp.bits.push_back(j);
    grp.f0.push_back(f0);
    grp.f1.push_back(f1);
    grp.members = 1;
    grp.op_index = ps.ops.size();
    HOp op;
    op.kind = WK_TABLE;
    op.group = (int)ps.groups.size();
    // remember the original gate so a single-member group can fall back to a plain op
    op.cmask = g.cmask | (g.kind == IK_DIAG ? (1ull << g.t0) : 0ull);
    op.tmask = g.kind == IK_RZ ? (1ull << g.t0) : 0ull;
    memcpy(op.m, g.p, 4 * sizeof(double));
    op.target = g.kind == IK_RZ ? -2 : -1;      // -2 marks "plain form is WK_RZ"
    ps.ops.push_back(op);                        // diagonal: does not block other groups
    ps.groups.push_back(grp);
    return true;
}

static void lower_gate(Pass& ps, const PhysGate& g, bool merge) {
    HOp op;
    switch (g.kind) {
        case IK_H: { double m[8] = {g.p[0], g.p[0], g.p[0], -g.p[0]}; push_pair(ps, WK_REAL, g.t0, g.cmask, m); break; }   // H = s*[[1,1],[1,-1]]
        case IK_X: push_pair(ps, WK_X, g.t0, g.cmask, nullptr); break;
        case IK_Y: { double m[8] = {0, 0, 0, -1, 0, 1, 0, 0}; push_pair(ps, WK_U2, g.t0, g.cmask, m); break; }   // [[0,-i],[i,0]]
        case IK_U2: classify_u2(g.p, &op); push_pair(ps, op.kind, g.t0, g.cmask, op.m); break;
        case IK_SWAP: {   // SWAP(a,b) = CX(b->a) CX(a->b) CX(b->a), each under the original controls
            push_pair(ps, WK_X, g.t0, g.cmask | (1ull << g.t1), nullptr);
            push_pair(ps, WK_X, g.t1, g.cmask | (1ull << g.t0), nullptr);
            push_pair(ps, WK_X, g.t0, g.cmask | (1ull << g.t1), nullptr);
            break;
        }
        case IK_DIAG:
            if (merge && merge_diag(ps, g)) break;
            op.kind = WK_DIAG;
            op.cmask = g.cmask | (g.t0 >= 0 ? (1ull << g.t0) : 0ull);
            op.m[0] = g.p[0]; op.m[1] = g.p[1];
            append_op(ps, op, 0);
            break;
        case IK_RZ:
            if (merge && merge_diag(ps, g)) break;
            op.kind = WK_RZ;
            op.cmask = g.cmask;
            op.tmask = g.t0 >= 0 ? (1ull << g.t0) : 0ull;
            memcpy(op.m, g.p, 4 * sizeof(double));
            if (g.t0 < 0) { op.m[2] = g.p[0]; op.m[3] = g.p[1]; }   // target in the rank bits: phase pre-resolved
            append_op(ps, op, 0);
            break;
        default: break;
    }
}

// ---- late placement of unconditional phase tables ----------------------------------------------------------------
// merge_diag puts a diagonal gate into the latest open group that no later op blocks, else opens a new group at the
// current end of the pass; a pass that spans several layers ends up with several unconditional tables although a table
// op costs as much as two Hadamards.  Every uncontrolled single-qubit diagonal gate commutes with every op that does
// not use its qubit non-diagonally, so inside the pass it may sit anywhere between the previous and the next such op:
// an interval.  The fewest tables that serve all gates is the minimum piercing set of those intervals -- greedy by
// the earliest right end, i.e. every table is placed as LATE as its most urgent member allows.
static void repack_unconditional_tables(Pass& ps) {
    const size_t n = ps.ops.size();
    struct Item { int q; amp_t f0, f1; size_t lo, hi; };
    std::vector<Item> items;
    std::vector<char> drop(n, 0);
    auto n_uses = [&](const HOp& h, int q) { return h.kind >= WK_X && h.kind <= kLastPairKind && h.target == q; };
    size_t tables = 0;
    for (size_t i = 0; i < n; i++) {
        const HOp& h = ps.ops[i];
        if (h.kind != WK_TABLE) continue;
        const DiagGroup& g = ps.groups[h.group];
        if (g.hub >= 0 || g.hub_alt >= 0) continue;
        drop[i] = 1;
        tables++;
        for (size_t k = 0; k < g.bits.size(); k++) {
            Item it{g.bits[k], g.f0[k], g.f1[k], 0, n};
            for (size_t j = i; j-- > 0;) if (n_uses(ps.ops[j], it.q)) { it.lo = j + 1; break; }
            for (size_t j = i + 1; j < n; j++) if (n_uses(ps.ops[j], it.q)) { it.hi = j; break; }
            items.push_back(it);
        }
    }
    if (tables < 2) return;
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.hi < b.hi; });
    std::vector<std::vector<HOp>> insert_at(n + 1);
    std::vector<char> used(items.size(), 0);
    size_t new_tables = 0;
    for (size_t a = 0; a < items.size(); a++) {
        if (used[a]) continue;
        const size_t point = items[a].hi;           // the table sits just before op `point`
        DiagGroup grp;
        for (size_t b = a; b < items.size(); b++) {
            if (used[b] || items[b].lo > point) continue;
            used[b] = 1;
            grp.members++;
            bool found = false;
            for (size_t k = 0; k < grp.bits.size(); k++)
                if (grp.bits[k] == items[b].q) { grp.f0[k] = cmul(grp.f0[k], items[b].f0); grp.f1[k] = cmul(grp.f1[k], items[b].f1); found = true; break; }
            if (!found) { grp.bits.push_back(items[b].q); grp.f0.push_back(items[b].f0); grp.f1.push_back(items[b].f1); }
        }
        HOp op;
        op.kind = WK_TABLE;
        op.group = (int)ps.groups.size();
        op.target = -2;                              // a single-member group falls back to the plain RZ form diag(f0, f1)
        op.tmask = 1ull << grp.bits[0];
        op.m[0] = grp.f0[0].x; op.m[1] = grp.f0[0].y; op.m[2] = grp.f1[0].x; op.m[3] = grp.f1[0].y;
        if (grp.bits.size() == 1) grp.members = 1;   // several gates on ONE qubit are one plain op
        ps.groups.push_back(grp);
        insert_at[point].push_back(op);
        new_tables++;
    }
    if (new_tables >= tables) return;                // nothing gained: keep the pass as it was built
    std::vector<HOp> out;
    out.reserve(n + new_tables);
    for (size_t i = 0; i <= n; i++) {
        for (const HOp& t : insert_at[i]) out.push_back(t);
        if (i < n && !drop[i]) out.push_back(ps.ops[i]);
    }
    for (size_t i = 0; i < out.size(); i++)
        if (out[i].kind == WK_TABLE) ps.groups[out[i].group].op_index = i;
    ps.ops.swap(out);
}

// ---- host: lowering a pass to device form -----------------------------------------------------------
// (Layout / make_layout / split_mask: window_layout.cuh)

// build the tables of one group: [lane(32) | slot(2^R) | chunk0(256) | chunk1(256) ...]
static void build_tables(const Layout& L, const DiagGroup& g, std::vector<amp_t>& arena, DOp* d, double scale = 1.0) {
    const int S = 1 << L.R;
    const int nch_total = (L.ntile_bits + 7) / 8;
    int used_chunks = 0;
    bool has_reg = false;
    for (size_t k = 0; k < g.bits.size(); k++) {
        int q = g.bits[k];
        if (L.cls[q] == CLS_TILE) used_chunks = std::max(used_chunks, L.idx[q] / 8 + 1);
        if (L.cls[q] == CLS_REG) has_reg = true;
    }
    (void)nch_total;
    const size_t off = arena.size();
    const int NT = L.nt;
    arena.resize(off + NT + S + 256 * (size_t)used_chunks, make_double2(1.0, 0.0));
    amp_t* lane_t = arena.data() + off;
    amp_t* slot_t = lane_t + NT;
    amp_t* chunk_t = slot_t + S;
    for (size_t k = 0; k < g.bits.size(); k++) {
        const int q = g.bits[k];
        const amp_t f0 = g.f0[k], f1 = g.f1[k];
        if (L.cls[q] == CLS_LANE) {
            for (int i = 0; i < NT; i++) lane_t[i] = cmul(lane_t[i], ((i >> L.idx[q]) & 1) ? f1 : f0);
        } else if (L.cls[q] == CLS_REG) {
            for (int i = 0; i < S; i++) slot_t[i] = cmul(slot_t[i], ((i >> L.idx[q]) & 1) ? f1 : f0);
        } else {
            const int ch = L.idx[q] / 8, b = L.idx[q] % 8;
            for (int i = 0; i < 256; i++) chunk_t[256 * ch + i] = cmul(chunk_t[256 * ch + i], ((i >> b) & 1) ? f1 : f0);
        }
    }
    if (scale != 1.0)                      // lean lowering: the pass's deferred gate scale rides on an unconditional table
        for (int i = 0; i < NT; i++) lane_t[i] = make_double2(lane_t[i].x * scale, lane_t[i].y * scale);
    if (NT == kTileThreads && has_reg) {   // k_tile applies the slot factors as packed in-place rotations (exec_op_tile)
        std::vector<amp_t> packed(S);
        for (int i = 0; i < S; i++) packed[i] = pack_rot(arena[off + NT + i]);
        arena.insert(arena.end(), packed.begin(), packed.end());
    }
    d->kind = WK_TABLE;
    d->nchunks = (uint8_t)used_chunks;
    d->has_reg = has_reg ? 1 : 0;
    d->hub_cls = CLS_NONE;
    d->hub_bit = 0;
    if (g.hub >= 0) { d->hub_cls = (uint8_t)L.cls[g.hub]; d->hub_bit = (uint8_t)L.idx[g.hub]; }
    long long o = (long long)off;
    memcpy(&d->m[0], &o, sizeof(o));
}

// bit sl set iff slot sl has every positive register-bit control on and every negative one off
static uint32_t slot_mask(int R, uint32_t pos, uint32_t neg) {
    uint32_t m = 0;
    for (int sl = 0; sl < (1 << R); sl++)
        if (((uint32_t)sl & pos) == pos && ((uint32_t)sl & neg) == 0) m |= 1u << sl;
    return m;
}

// ---- lean lowering (option "lean") --------------------------------------------------------------------
// G = g * G' with G' in one of the unit forms of the WK_*U kinds.  Only gates that act on the WHOLE state qualify, so that
// g is a global scalar: uncontrolled REAL / RX / RXS ops, and the two halves of an absorbed CNOT when both give the same g
// (they act on complementary halves of the state).  Returns the product of the g (1.0 = nothing converted).
static bool lean_form(const HOp& h, double* g, HOp* lean) {
    const double kMin = 0.3;                 // keep p, q, t = O(1): below this the scaled form is used
    *lean = h;
    memset(lean->m, 0, sizeof(lean->m));
    if (h.kind == WK_RX || h.kind == WK_RXS) {
        if (std::fabs(h.m[0]) < kMin) return false;
        *g = h.m[0];
        lean->kind = h.kind == WK_RX ? WK_RXU : WK_RXSU;
        lean->m[0] = h.m[1] / h.m[0];
        return true;
    }
    if (h.kind == WK_REAL) {                 // (k00, k01, k10, k11)
        if (std::fabs(h.m[0]) < kMin) return false;
        if (h.m[3] == h.m[0]) lean->kind = WK_REALUP;
        else if (h.m[3] == -h.m[0]) lean->kind = WK_REALUM;
        else return false;
        *g = h.m[0];
        lean->m[0] = h.m[1] / h.m[0];
        lean->m[1] = h.m[2] / h.m[0];
        return true;
    }
    return false;
}

static double lean_convert(std::vector<HOp>& ops) {
    const size_t n = ops.size();
    std::vector<char> ok(n, 0);
    std::vector<double> g(n, 1.0);
    std::vector<HOp> lean(n);
    for (size_t i = 0; i < n; i++) ok[i] = ops[i].kind != 0 && lean_form(ops[i], &g[i], &lean[i]);
    size_t converted = 0;
    std::vector<char> take(n, 0);
    for (size_t i = 0; i < n; i++) {
        if (!ok[i]) continue;
        const HOp& h = ops[i];
        if (h.pair < 0) { if (h.cmask == 0 && h.nmask == 0) { take[i] = 1; converted++; } continue; }
        for (size_t j = i + 1; j < n; j++)
            if (ops[j].kind != 0 && ops[j].pair == h.pair) {
                const bool halves = (h.cmask && !h.nmask && ops[j].nmask == h.cmask && !ops[j].cmask) ||
                                    (h.nmask && !h.cmask && ops[j].cmask == h.nmask && !ops[j].nmask);
                if (ok[j] && halves && g[j] == g[i]) { take[i] = 1; take[j] = 2; converted++; }     // 2 = its g is already counted
                break;
            }
    }
    if (converted < 3) return 1.0;           // the scale op would cost more than the lean forms save
    double scale = 1.0;
    for (size_t i = 0; i < n; i++) {
        if (!take[i]) continue;
        if (take[i] == 1) scale *= g[i];
        ops[i] = lean[i];
    }
    return scale;
}

// one host op -> one device op under layout L (`scale`: the pass's deferred lean scale, consumed by the first
// unconditional table)
static void lower_op(const Layout& L, const Pass& ps, const HOp& h, double* scale, std::vector<DOp>& dops, std::vector<amp_t>& arena) {
    DOp d;
    memset(&d, 0, sizeof(d));
    HOp op = h;
    if (op.kind == WK_TABLE) {
        const DiagGroup& g = ps.groups[op.group];
        if (g.members >= 2) {
            const bool carries = *scale != 1.0 && g.hub < 0 && g.hub_alt < 0;      // unconditional: touches every amplitude
            build_tables(L, g, arena, &d, carries ? *scale : 1.0);
            if (carries) *scale = 1.0;
            dops.push_back(d);
            return;
        }
        op.kind = (op.target == -2) ? WK_RZ : WK_DIAG;      // single member: plain op is cheaper
        if (op.kind == WK_DIAG) { op.m[0] = h.m[0]; op.m[1] = h.m[1]; }
    }
    if (op.kind == WK_DIAG && op.m[0] == -1.0 && op.m[1] == 0.0) op.kind = WK_NEG;
    d.kind = (uint8_t)op.kind;
    {
        uint32_t pos_lane = 0, pos_reg = 0, neg_lane = 0, neg_reg = 0;
        uint64_t pos_tile = 0, neg_tile = 0;
        split_mask(L, op.cmask, &pos_lane, &pos_reg, &pos_tile);
        split_mask(L, op.nmask, &neg_lane, &neg_reg, &neg_tile);      // k_window: neg_lane is always 0 (absorb_cnot refuses lane controls)
        d.c_lane = pos_lane | neg_lane;
        d.c_lval = (uint8_t)pos_lane;
        d.c_reg = slot_mask(L.R, pos_reg, neg_reg);
        d.c_tile = pos_tile | neg_tile;
        d.c_tval = pos_tile;
    }
    if (op.kind == WK_RZ) split_mask(L, op.tmask, &d.t_lane, &d.t_reg, &d.t_tile);
    memcpy(d.m, op.m, sizeof(d.m));
    if (op.kind != WK_DIAG && op.kind != WK_RZ && op.kind != WK_NEG) {
        const int t = op.target;
        d.tpos = (uint8_t)(L.cls[t] == CLS_LANE ? L.idx[t] : kLaneQubits + L.idx[t]);
    }
    dops.push_back(d);
}

static void push_scale_op(const Layout& L, double scale, std::vector<DOp>& dops) {      // no table to ride on: one real scale op
    DOp d;
    memset(&d, 0, sizeof(d));
    d.kind = WK_SCALE;
    d.c_reg = slot_mask(L.R, 0, 0);
    d.m[0] = scale;
    dops.push_back(d);
}

// ---- k_tile lowering: one host op -> one or more in-place device ops under the layout of its round ------------------
// (the op set and why every op is in place: "k_tile: the op set" above)
// `scale`: product of the factors the launch's unit-form gates leave out (applied by one WK_SCALE op at the end of the launch)
static int lower_op_tile(const Layout& L, const Pass& ps, const HOp& h, std::vector<DOp>& dops, std::vector<amp_t>& arena, amp_t* scale) {
    DOp base;
    memset(&base, 0, sizeof(base));
    uint32_t pos_lane = 0, pos_reg = 0, neg_lane = 0, neg_reg = 0;
    uint64_t pos_tile = 0, neg_tile = 0;
    {
        uint64_t cm = h.cmask, nm = h.nmask;
        if (h.kind == WK_TABLE) { cm = 0; nm = 0; }            // a table's cmask only remembers its first gate (single-member fallback)
        split_mask(L, cm, &pos_lane, &pos_reg, &pos_tile);
        split_mask(L, nm, &neg_lane, &neg_reg, &neg_tile);
    }
    auto controls = [&](DOp* d, uint32_t preg, uint32_t nreg) {
        d->c_lane = pos_lane | neg_lane;
        d->c_lval = (uint8_t)pos_lane;
        d->c_reg = slot_mask(L.R, preg, nreg);
        d->c_tile = pos_tile | neg_tile;
        d->c_tval = pos_tile;
    };
    auto emit = [&](DOp d, uint32_t preg, uint32_t nreg, int B) -> int {
        d.code = (uint8_t)fast_code((int)d.kind, B, preg, nreg);
        if (d.code == 0) return fail(QI_ERR_UNKNOWN, d.kind, 0, "internal: tile op without an in-place routine");
        dops.push_back(d);
        return QI_OK;
    };
    // diagonal op multiplying by `f0` where the target bit is 0 and by `f1` where it is 1 (same controls as the host op);
    // tmask = 0: one phase f1 on everything the controls select (the target already sits in the control masks)
    auto emit_phase = [&](amp_t f0, amp_t f1, uint64_t tmask) -> int {
        DOp d = base;
        if (tmask == 0) {
            const bool neg = f1.x == -1.0 && f1.y == 0.0;
            d.kind = neg ? WK_NEG : WK_DIAG;
            controls(&d, pos_reg, neg_reg);
            d.m[0] = f1.x; d.m[1] = f1.y;
            const amp_t r = pack_rot(f1);
            d.m[2] = r.x; d.m[3] = r.y;
            return emit(d, pos_reg, neg_reg, -1);
        }
        d.kind = WK_RZ;
        controls(&d, pos_reg, neg_reg);
        split_mask(L, tmask, &d.t_lane, &d.t_reg, &d.t_tile);
        d.m[0] = f0.x; d.m[1] = f0.y; d.m[2] = f1.x; d.m[3] = f1.y;
        const amp_t r0 = pack_rot(f0), r1 = pack_rot(f1);
        d.m[4] = r0.x; d.m[5] = r0.y; d.m[6] = r1.x; d.m[7] = r1.y;
        return emit(d, pos_reg, neg_reg, -1);
    };
    auto emit_x = [&](int target) -> int {
        DOp d = base;
        d.kind = WK_X;
        controls(&d, pos_reg, neg_reg);
        d.tpos = (uint8_t)(kLaneQubits + L.idx[target]);
        return emit(d, pos_reg, neg_reg, L.idx[target]);
    };
    // real 2x2 (k00, k01, k10, k11) in lifted form; a small pivot is moved away by an X in front (M = (M X) X)
    auto emit_real = [&](int target, double k0, double k1, double k2, double k3) -> int {
        if (pos_reg | neg_reg) return fail(QI_ERR_UNKNOWN, 0, 0, "internal: pair gate under a register-bit control in a tile round");
        if (std::fabs(k0) < kLiftMinPivot) {
            QI_TRY(emit_x(target));
            std::swap(k0, k1);
            std::swap(k2, k3);
        }
        DOp d = base;
        controls(&d, 0, 0);
        d.tpos = (uint8_t)(kLaneQubits + L.idx[target]);
        const bool uncontrolled = !(pos_lane | neg_lane) && !(pos_tile | neg_tile);
        if (uncontrolled && ctx().opt_tile_lean && (k3 == k0 || k3 == -k0)) {       // unit form: k0 . [[1, k1 / k0], [k2 / k0, +-1]]
            d.kind = k3 == k0 ? WK_REALUP : WK_REALUM;
            d.m[0] = k1 / k0; d.m[1] = k2 / k0;
            scale->x *= k0; scale->y *= k0;
            return emit(d, 0, 0, L.idx[target]);
        }
        d.kind = WK_REALL;
        d.m[0] = k0; d.m[1] = k1; d.m[2] = k2 / k0; d.m[3] = (k0 * k3 - k1 * k2) / k0;
        return emit(d, 0, 0, L.idx[target]);
    };
    // RX-form (c, s): [[c, -i s], [-i s, c]]; small |c|: RX = (-i X) . RX' with c' = s, s' = -c
    auto emit_rx = [&](int target, double c, double sn) -> int {
        if (pos_reg | neg_reg) return fail(QI_ERR_UNKNOWN, 0, 0, "internal: pair gate under a register-bit control in a tile round");
        const bool pivot = std::fabs(c) < kLiftMinPivot;
        if (pivot) { const double c2 = sn, s2 = -c; c = c2; sn = s2; }
        DOp d = base;
        controls(&d, 0, 0);
        d.tpos = (uint8_t)(kLaneQubits + L.idx[target]);
        const bool uncontrolled = !(pos_lane | neg_lane) && !(pos_tile | neg_tile);
        if (uncontrolled && ctx().opt_tile_lean) {                                  // unit form: c . [[1, -i t], [-i t, 1]]
            d.kind = WK_RXU;
            d.m[0] = sn / c;
            scale->x *= c; scale->y *= c;
        } else {
            d.kind = WK_RXL;
            d.m[0] = c; d.m[1] = sn; d.m[2] = 1.0 / c; d.m[3] = sn / c;
        }
        QI_TRY(emit(d, 0, 0, L.idx[target]));
        if (pivot) {
            QI_TRY(emit_x(target));
            QI_TRY(emit_phase(make_double2(0.0, -1.0), make_double2(0.0, -1.0), 0));
        }
        return QI_OK;
    };

    if (h.kind == WK_TABLE) {
        const DiagGroup& g = ps.groups[h.group];
        // A small UNCONDITIONAL group is a product of single-qubit diagonals diag(f0, f1) (the RZ gates of a layer that met in one
        // table).  Each is applied in P FORM: f0 goes into the launch's scalar, diag(1, f1 / f0) is a phase on the HALF of the
        // amplitudes whose bit is set -- 8 of a thread's 16 slots when the qubit is a register qubit (24 FMA), a predicate that
        // skips whole threads / tiles otherwise -- instead of a table op (thread-table and chunk-table loads, 64 FP64
        // instructions on all 16 slots, 48 more when register qubits take part).  Large groups (QFT ladders) stay tables.
        if (ctx().opt_tile_pform && g.members >= 2 && g.hub < 0 && g.hub_alt < 0 && g.bits.size() <= (size_t)std::min(4, ctx().opt_tile_pform)) {
            for (size_t k = 0; k < g.bits.size(); k++) {
                const amp_t f0 = g.f0[k], f1 = g.f1[k];
                const double n0 = f0.x * f0.x + f0.y * f0.y;
                if (n0 == 0.0) return fail(QI_ERR_UNKNOWN, 0, 0, "internal: zero diagonal factor");
                const amp_t ratio = make_double2((f1.x * f0.x + f1.y * f0.y) / n0, (f1.y * f0.x - f1.x * f0.y) / n0);      // f1 / f0
                *scale = cmul(*scale, f0);
                pos_lane = pos_reg = neg_lane = neg_reg = 0; pos_tile = neg_tile = 0;
                split_mask(L, 1ull << g.bits[k], &pos_lane, &pos_reg, &pos_tile);
                if (!(ratio.x == 1.0 && ratio.y == 0.0)) QI_TRY(emit_phase(make_double2(1.0, 0.0), ratio, 0));
            }
            return QI_OK;
        }
        // The same for a small group UNDER A HUB (controlled phases that share a qubit, e.g. the CZs a brick-work layer leaves
        // on one qubit): with the hub set, member k multiplies by diag(f0, f1).  prod f0 is ONE phase on the hub half; each
        // f1 / f0 is a phase on the quarter (hub, bit k) -- and a plain sign flip (free in the modules) when the member is a CZ.
        if (ctx().opt_tile_pform && g.members >= 2 && g.hub >= 0 && g.hub_alt < 0 && g.bits.size() <= (size_t)std::min(4, ctx().opt_tile_pform)) {
            amp_t F0 = make_double2(1.0, 0.0);
            std::vector<amp_t> ratio(g.bits.size());
            for (size_t k = 0; k < g.bits.size(); k++) {
                const amp_t f0 = g.f0[k], f1 = g.f1[k];
                const double n0 = f0.x * f0.x + f0.y * f0.y;
                if (n0 == 0.0) return fail(QI_ERR_UNKNOWN, 0, 0, "internal: zero diagonal factor");
                ratio[k] = (f0.x == 1.0 && f0.y == 0.0) ? f1 : make_double2((f1.x * f0.x + f1.y * f0.y) / n0, (f1.y * f0.x - f1.x * f0.y) / n0);
                F0 = cmul(F0, f0);
            }
            auto under = [&](uint64_t mask) {
                pos_lane = pos_reg = neg_lane = neg_reg = 0; pos_tile = neg_tile = 0;
                split_mask(L, mask, &pos_lane, &pos_reg, &pos_tile);
            };
            if (!(F0.x == 1.0 && F0.y == 0.0)) { under(1ull << g.hub); QI_TRY(emit_phase(make_double2(1.0, 0.0), F0, 0)); }
            for (size_t k = 0; k < g.bits.size(); k++) {
                if (ratio[k].x == 1.0 && ratio[k].y == 0.0) continue;
                under((1ull << g.hub) | (1ull << g.bits[k]));
                QI_TRY(emit_phase(make_double2(1.0, 0.0), ratio[k], 0));
            }
            return QI_OK;
        }
        if (g.members >= 2) {
            DOp d = base;
            build_tables(L, g, arena, &d, 1.0);
            d.c_tval = d.c_tile = 0;
            d.c_reg = slot_mask(L.R, 0, 0);
            return emit(d, 0, 0, -1);
        }
        // single member: the plain op (cmask / tmask / m remember the gate)
        split_mask(L, h.cmask, &pos_lane, &pos_reg, &pos_tile);
        neg_lane = neg_reg = 0; neg_tile = 0;
        if (h.target == -2) return emit_phase(make_double2(h.m[0], h.m[1]), make_double2(h.m[2], h.m[3]), h.tmask);
        return emit_phase(make_double2(1.0, 0.0), make_double2(h.m[0], h.m[1]), 0);
    }
    switch (h.kind) {
        case WK_DIAG: return emit_phase(make_double2(1.0, 0.0), make_double2(h.m[0], h.m[1]), 0);
        case WK_RZ: return emit_phase(make_double2(h.m[0], h.m[1]), make_double2(h.m[2], h.m[3]), h.tmask);
        case WK_X: return emit_x(h.target);
        case WK_REAL: return emit_real(h.target, h.m[0], h.m[1], h.m[2], h.m[3]);
        case WK_RX: return emit_rx(h.target, h.m[0], h.m[1]);
        case WK_RXS: QI_TRY(emit_x(h.target)); return emit_rx(h.target, h.m[0], h.m[1]);      // RX o X
        case WK_U2: {
            // U = diag(p0, p1) . [[c, -s], [s, c]] . diag(1, q),  c = |m00|, s = |m10|  (exact for a unitary U)
            const amp_t m00 = make_double2(h.m[0], h.m[1]), m01 = make_double2(h.m[2], h.m[3]);
            const amp_t m10 = make_double2(h.m[4], h.m[5]), m11 = make_double2(h.m[6], h.m[7]);
            const double c = std::hypot(m00.x, m00.y), sn = std::hypot(m10.x, m10.y);
            const amp_t p0 = c > 0.0 ? make_double2(m00.x / c, m00.y / c) : make_double2(1.0, 0.0);
            const amp_t p1 = sn > 0.0 ? make_double2(m10.x / sn, m10.y / sn) : make_double2(1.0, 0.0);
            amp_t q;
            if (c >= sn) { const amp_t z = cmul(m11, cconj(p1)); q = make_double2(z.x / c, z.y / c); }
            else { const amp_t z = cmul(m01, cconj(p0)); q = make_double2(-z.x / sn, -z.y / sn); }
            const uint64_t tb = 1ull << h.target;
            QI_TRY(emit_phase(make_double2(1.0, 0.0), q, tb));
            QI_TRY(emit_real(h.target, c, -sn, sn, c));
            return emit_phase(p0, p1, tb);
        }
        default: return fail(QI_ERR_UNKNOWN, h.kind, 0, "internal: unknown host op kind in a tile round");
    }
}

static void lower_pass(const qi_state* s, const Pass& ps, int R, std::vector<DOp>& dops, std::vector<amp_t>& arena, Layout* Lout) {
    Layout L = make_layout(s, ps.regs, R);
    *Lout = L;
    std::vector<HOp> ops(ps.ops);
    double scale = ctx().opt_lean ? lean_convert(ops) : 1.0;
    for (const HOp& h : ops) {
        if (h.kind == 0) continue;              // absorbed into a neighbour (absorb_cnot)
        lower_op(L, ps, h, &scale, dops, arena);
    }
    if (scale != 1.0) push_scale_op(L, scale, dops);
}

// ---- host: a tile pass = rounds ---------------------------------------------------------------------------
struct TilePlan {                // a CTA-tile pass in the physical coordinates it runs under
    int pin[kTileBits];          // ascending physical positions of the tile's local bits (pin[0..4] = 0..4)
    int pout[kTileBits];         // position the content of local bit j is stored to: a permutation of pin (identity = no relabelling)
};
struct TileRoundHost {
    int regs[4];                 // tile-local bits held in registers (ascending)
    int thr[kTileThrBits];       // tile-local bit of thread-index bit k
    size_t first_op = 0, nops = 0;
};
struct TileLaunch {
    int tile_qubits[kTileBits];  // physical position of tile-local bit j (ascending)
    int tile_out[kTileBits];     // physical position the content of local bit j is stored to
    std::vector<TileRoundHost> rounds;
    std::vector<DOp> dops;
};

static GateUse uses_of_hop(const Pass& ps, const HOp& h) {
    GateUse u{0, 0};
    if (h.kind == WK_TABLE) {
        const DiagGroup& g = ps.groups[h.group];
        if (g.hub >= 0) u.d_use |= 1ull << g.hub;
        if (g.hub_alt >= 0) u.d_use |= 1ull << g.hub_alt;
        for (int q : g.bits) u.d_use |= 1ull << q;
        u.d_use |= h.cmask | h.tmask;
    } else if (h.kind == WK_DIAG || h.kind == WK_RZ) u.d_use = h.cmask | h.tmask;
    else if (h.kind >= WK_X && h.kind <= kLastPairKind) { u.n_use = 1ull << h.target; u.d_use = h.cmask | h.nmask; }
    return u;
}

// Split the ops of a pass into rounds of <= 4 register qubits.  Same commutation rule as the pass scheduler: an op may
// move ahead of the ops skipped before it when on every shared qubit both act diagonally; a pair op needs its target
// among the round's register qubits.  Returns op indices per round and the register qubits (physical) of each round.
static void form_rounds(const Pass& ps, const std::vector<HOp>& ops, uint64_t tile_mask, std::vector<std::vector<size_t>>* round_ops,
                        std::vector<std::vector<int>>* round_regs, std::vector<uint64_t>* round_keep_out) {
    std::vector<size_t> rest;
    for (size_t i = 0; i < ops.size(); i++) if (ops[i].kind != 0) rest.push_back(i);
    while (!rest.empty()) {
        std::vector<size_t> take, keep;
        std::vector<int> Q;
        uint64_t qmask = 0, blocked_any = 0, blocked_n = 0;
        uint64_t keep_out = 0;             // controls of the pair gates taken so far: they must not become register qubits
        for (size_t i : rest) {
            const HOp& h = ops[i];
            const GateUse u = uses_of_hop(ps, h);
            bool ok = ((u.n_use & blocked_any) == 0) && ((u.d_use & blocked_n) == 0);
            const uint64_t ctrl = u.n_use ? (h.cmask | h.nmask) : 0ull;
            // the in-place pair routines take no register-bit controls; X has variants under exactly one
            const int reg_ctrl_ok = (h.kind == WK_X && __builtin_popcountll(ctrl) == 1) ? 1 : 0;
            if (ok && u.n_use) {
                const uint64_t ko = reg_ctrl_ok ? keep_out : (keep_out | ctrl);
                if (__builtin_popcountll(ctrl & qmask) > reg_ctrl_ok) ok = false;
                else if (__builtin_popcountll(tile_mask & ~ko) < 4) ok = false;       // four register qubits must remain possible
                else if (!(u.n_use & qmask)) {
                    if ((int)Q.size() < 4 && !(u.n_use & ko)) { Q.push_back(h.target); qmask |= u.n_use; }
                    else ok = false;
                }
                if (ok) keep_out = ko;
            }
            if (ok) take.push_back(i);
            else { keep.push_back(i); blocked_any |= u.n_use | u.d_use; blocked_n |= u.n_use; }
        }
        round_ops->push_back(take);
        round_regs->push_back(Q);
        round_keep_out->push_back(keep_out);
        rest.swap(keep);
    }
}

// Layout of a round.  `low` (5 tile-local bits, in order) makes it an IO layout: those bits become thread bits 0..4 -- the
// lanes -- so that a warp's global access is one contiguous 512-byte row (loads: the local bits at positions 0..4; stores:
// the local bits whose content goes to positions 0..4); the register bits then avoid them.  Without `low` the thread
// bits are ordered for bank-conflict-free regroups: tile_swz folds local bits 3-5, 6-8, 9-10 onto 0-2, so the three
// local bits behind thread bits 0..2 should fall into three different classes.
static bool make_tile_round(const std::vector<int>& regs_local, const int* low, uint32_t avoid_local, TileRoundHost* r) {
    std::vector<int> loc(regs_local);
    auto is_low = [&](int j) { if (!low) return false; for (int k = 0; k < kLaneQubits; k++) if (low[k] == j) return true; return false; };
    for (int j : loc) if (is_low(j)) return false;
    // pad to four register bits; never with a bit that controls one of the round's pair gates (`avoid_local`)
    for (int j = kTileBits - 1; j >= 0 && (int)loc.size() < 4; j--)
        if (!is_low(j) && !((avoid_local >> j) & 1u) && std::find(loc.begin(), loc.end(), j) == loc.end()) loc.push_back(j);
    if ((int)loc.size() < 4) return false;
    std::sort(loc.begin(), loc.end());
    for (int k = 0; k < 4; k++) r->regs[k] = loc[k];
    std::vector<int> free_bits;
    for (int j = 0; j < kTileBits; j++)
        if (std::find(loc.begin(), loc.end(), j) == loc.end() && !is_low(j)) free_bits.push_back(j);
    std::vector<int> order;
    if (low) {
        for (int k = 0; k < kLaneQubits; k++) order.push_back(low[k]);
        for (int j : free_bits) order.push_back(j);
    } else {
        auto cls = [](int j) { return j < 9 ? j % 3 : j - 9; };
        bool have[3] = {false, false, false};
        for (int j : free_bits) if (!have[cls(j)]) { have[cls(j)] = true; order.push_back(j); }
        for (int j : free_bits) if (std::find(order.begin(), order.end(), j) == order.end()) order.push_back(j);
    }
    for (int k = 0; k < kTileThrBits; k++) r->thr[k] = order[k];
    return true;
}

static Layout round_layout(const qi_state* s, const int tile_qubits[kTileBits], const TileRoundHost& r) {
    Layout L;
    L.R = 4;
    L.nt = kTileThreads;
    const int n = (int)s->n_local;
    int local_of[64];
    for (int q = 0; q < 64; q++) local_of[q] = -1;
    for (int j = 0; j < kTileBits; j++) local_of[tile_qubits[j]] = j;
    int t = 0;
    for (int q = 0; q < 64; q++) {
        L.cls[q] = CLS_NONE; L.idx[q] = 0;
        if (q >= n) continue;
        const int j = local_of[q];
        if (j < 0) { L.cls[q] = CLS_TILE; L.idx[q] = t++; continue; }
        int slot = -1;
        for (int k = 0; k < 4; k++) if (r.regs[k] == j) slot = k;
        if (slot >= 0) { L.cls[q] = CLS_REG; L.idx[q] = slot; L.regs.push_back(q); continue; }
        for (int k = 0; k < kTileThrBits; k++) if (r.thr[k] == j) { L.cls[q] = CLS_LANE; L.idx[q] = k; }
    }
    L.ntile_bits = t;
    return L;
}

// device ops lower_op_tile emits for a host op (launch packing)
static size_t tile_op_count(const Pass& ps, const HOp& h) {
    switch (h.kind) {
        case WK_REAL: return std::fabs(h.m[0]) < kLiftMinPivot ? 2 : 1;
        case WK_RX: return std::fabs(h.m[0]) < kLiftMinPivot ? 3 : 1;
        case WK_RXS: return std::fabs(h.m[0]) < kLiftMinPivot ? 4 : 2;
        case WK_U2: return 4;
        case WK_TABLE: {                  // a small group lowers to one P-form op per member qubit (+ one for the hub half)
            const DiagGroup& g = ps.groups[h.group];
            const size_t pf = (size_t)std::min(4, std::max(0, ctx().opt_tile_pform));
            if (g.members >= 2 && g.hub_alt < 0 && g.bits.size() <= pf) return g.bits.size() + (g.hub >= 0 ? 1 : 0);
            return 1;
        }
        default: return 1;
    }
}

// lower one pass to one or more k_tile launches; `plan` = the tile's positions and where its content is stored to
// `carry`: the scalar earlier launches of this run left out (unit-form gates, P-form diagonals); a scalar commutes with everything,
// so it travels from launch to launch and is applied once, by the LAST launch of the run (`last_of_run`) -- or earlier when its
// magnitude leaves [2^-200, 2^200] (the stored amplitudes grow by its inverse)
static int lower_tile_pass(const qi_state* s, const Pass& ps, const TilePlan& plan, std::vector<TileLaunch>& launches, std::vector<amp_t>& arena,
                           amp_t* carry, bool last_of_run) {
    int local_of[64];
    for (int q = 0; q < 64; q++) local_of[q] = -1;
    for (int j = 0; j < kTileBits; j++) local_of[plan.pin[j]] = j;

    const std::vector<HOp>& ops = ps.ops;      // (the lean unit forms of k_window are not used here)
    std::vector<std::vector<size_t>> round_ops;
    std::vector<std::vector<int>> round_regs;
    std::vector<uint64_t> round_keep;          // per round: qubits that control one of its pair gates (never register qubits)
    uint64_t tile_mask = 0;
    for (int j = 0; j < kTileBits; j++) tile_mask |= 1ull << plan.pin[j];
    form_rounds(ps, ops, tile_mask, &round_ops, &round_regs, &round_keep);
    // a round never carries more ops than one launch holds
    {
        std::vector<std::vector<size_t>> ro;
        std::vector<std::vector<int>> rq;
        std::vector<uint64_t> rk;
        const size_t kChunk = 40;                 // host ops per round chunk (an op lowers to at most 5 device ops)
        for (size_t r = 0; r < round_ops.size(); r++)
            for (size_t first = 0; first < std::max<size_t>(1, round_ops[r].size()); first += kChunk) {
                ro.emplace_back(round_ops[r].begin() + first, round_ops[r].begin() + std::min(round_ops[r].size(), first + kChunk));
                rq.push_back(round_regs[r]);
                rk.push_back(round_keep[r]);
            }
        if (ro.empty()) { ro.emplace_back(); rq.emplace_back(); rk.push_back(0); }
        round_ops.swap(ro);
        round_regs.swap(rq);
        round_keep.swap(rk);
    }
    const size_t nrounds = round_ops.size();
    const int low_id[kLaneQubits] = {0, 1, 2, 3, 4};
    int low_out_last[kLaneQubits];
    for (int k = 0; k < kLaneQubits; k++)
        for (int j = 0; j < kTileBits; j++) if (plan.pout[j] == k) low_out_last[k] = j;
    size_t ri = 0;
    while (ri < nrounds) {
        // rounds of this launch
        size_t end = ri, nops = 0;
        auto cost = [&](size_t r) { size_t c = 0; for (size_t i : round_ops[r]) c += tile_op_count(ps, ops[i]); return c; };
        while (end < nrounds && (end == ri || (nops + cost(end) <= (size_t)kMaxTileOps - 1 && (end - ri) + 3 <= (size_t)kMaxRounds))) {
            nops += cost(end);
            end++;
        }
        const bool last_launch = end == nrounds;
        const int* low_out = last_launch ? low_out_last : low_id;
        const bool same_io = !memcmp(low_out, low_id, sizeof(low_id));
        TileLaunch tl;
        amp_t scale = *carry;
        for (int j = 0; j < kTileBits; j++) { tl.tile_qubits[j] = plan.pin[j]; tl.tile_out[j] = last_launch ? plan.pout[j] : plan.pin[j]; }
        for (size_t r = ri; r < end; r++) {
            std::vector<int> loc;
            for (int q : round_regs[r]) loc.push_back(local_of[q]);
            TileRoundHost rd;
            uint32_t avoid = 0;
            for (int j = 0; j < kTileBits; j++) if ((round_keep[r] >> plan.pin[j]) & 1) avoid |= 1u << j;
            bool is_in = false, is_out = false;
            if (r == ri && r + 1 == end) {                    // only round: loads and stores
                if (same_io && make_tile_round(loc, low_id, avoid, &rd)) is_in = is_out = true;
                else if (make_tile_round(loc, low_id, avoid, &rd)) is_in = true;
            } else if (r == ri) is_in = make_tile_round(loc, low_id, avoid, &rd);
            else if (r + 1 == end) is_out = make_tile_round(loc, low_out, avoid, &rd);
            if (!is_in && !is_out && !make_tile_round(loc, nullptr, avoid, &rd))
                return fail(QI_ERR_UNKNOWN, 0, 0, "internal: no register layout for a tile round");
            if (r == ri && !is_in) {                          // a launch starts in the load layout: empty round first
                TileRoundHost io;
                make_tile_round(std::vector<int>(), low_id, 0, &io);
                io.first_op = tl.dops.size();
                tl.rounds.push_back(io);
            }
            rd.first_op = tl.dops.size();
            const Layout L = round_layout(s, tl.tile_qubits, rd);
            for (size_t i : round_ops[r]) QI_TRY(lower_op_tile(L, ps, ops[i], tl.dops, arena, &scale));
            if (tl.dops.size() > (size_t)kMaxTileOps) return fail(QI_ERR_UNKNOWN, tl.dops.size(), 0, "internal: tile launch holds too many ops");
            rd.nops = tl.dops.size() - rd.first_op;
            tl.rounds.push_back(rd);
            if (r + 1 == end && !is_out) {                    // ... and ends in the store layout
                TileRoundHost io;
                make_tile_round(std::vector<int>(), low_out, 0, &io);
                io.first_op = tl.dops.size();
                tl.rounds.push_back(io);
            }
        }
        const auto unit = [](amp_t z) { return z.x == 1.0 && z.y == 0.0; };
        {
            const double mag2 = scale.x * scale.x + scale.y * scale.y;
            const bool in_range = mag2 > 3.0e-121 && mag2 < 3.0e120;          // |scale| within [2^-200, 2^200]
            if (!(last_launch && last_of_run) && in_range) { *carry = scale; scale = make_double2(1.0, 0.0); }
            else *carry = make_double2(1.0, 0.0);
        }
        if (!unit(scale)) {                 // an unconditional phase table of the launch (it multiplies EVERY amplitude) carries the factor for free
            for (DOp& t : tl.dops)
                if (t.kind == WK_TABLE && t.hub_cls == CLS_NONE && !t.c_tile && !t.c_lane) {
                    long long off;
                    memcpy(&off, &t.m[0], 8);
                    for (int i = 0; i < kTileThreads; i++) arena[(size_t)off + i] = cmul(arena[(size_t)off + i], scale);
                    scale = make_double2(1.0, 0.0);
                    break;
                }
        }
        if (!unit(scale)) {                 // else: one op on every amplitude in the last round (real: 2 DMUL per amplitude; complex: a product)
            DOp d;
            memset(&d, 0, sizeof(d));
            d.kind = WK_SCALE;
            d.code = (uint8_t)FC_SCALE;
            d.c_reg = 0xffffu;
            d.m[0] = scale.x;
            d.m[1] = scale.y;
            tl.dops.push_back(d);
            tl.rounds.back().nops++;
        }
        launches.push_back(std::move(tl));
        ri = end;
    }
    return QI_OK;
}

static int launch_tile(qi_state* s, const TileLaunch& tl, const amp_t* d_tables) {
    Context& c = ctx();
    static TProgram P;            // 30 KB: not on the stack
    memset(&P, 0, sizeof(P));
    for (int j = 0; j < kTileBits; j++) { P.tpos[j] = (uint8_t)tl.tile_qubits[j]; P.tpos_out[j] = (uint8_t)tl.tile_out[j]; }
    P.nrounds = (uint8_t)tl.rounds.size();
    P.tables = d_tables;
    for (size_t r = 0; r < tl.rounds.size(); r++) {
        const TileRoundHost& h = tl.rounds[r];
        TRound& d = P.rounds[r];
        for (int sl = 0; sl < 16; sl++) {
            uint32_t o = 0;
            for (int k = 0; k < 4; k++) if ((sl >> k) & 1) o |= 1u << h.regs[k];
            d.sswz[sl] = (uint16_t)tile_swz(o);
        }
        for (int k = 0; k < kTileThrBits; k++) d.thr_pos[k] = (uint8_t)h.thr[k];
        d.first_op = (uint16_t)h.first_op;
        d.nops = (uint16_t)h.nops;
    }
    const TileRoundHost& in = tl.rounds.front();
    const TileRoundHost& out = tl.rounds.back();
    for (int sl = 0; sl < 16; sl++) {
        uint64_t oi = 0, oo = 0;
        for (int k = 0; k < 4; k++)
            if ((sl >> k) & 1) { oi |= 1ull << tl.tile_qubits[in.regs[k]]; oo |= 1ull << tl.tile_out[out.regs[k]]; }
        P.goff_in[sl] = oi;
        P.goff_out[sl] = oo;
    }
    memcpy(P.ops, tl.dops.data(), tl.dops.size() * sizeof(DOp));
    const uint64_t ntiles = s->len >> kTileBits;
    uint64_t blocks = std::min<uint64_t>(ntiles, (uint64_t)c.sm_count * 32);
    LaunchScope ls(KF_TILE, 32.0 * (double)s->len);
    k_tile<<<(unsigned)blocks, kTileThreads, 0, c.stream>>>(s->d, ntiles, P);
    return check_launch("k_tile");
}

#include "tile_jit.cuh"

// a tile pass prepared for the JIT path: its module (queued, ready or failed) and the coefficient block of THIS execution
struct TileJit {
    jit::Entry* e = nullptr;
    std::vector<double> coef;
    double fp64 = 0.0;           // FP64 instructions per thread and tile (weighted by control regions)
    int ctas = 4, groups = 1, prefetch = 0, stage = 0;
    bool ready = false;          // the module was assembled when this execution collected its coefficients: launch it
};
// upper bound of the coefficients a launch's module reads (4 per pair op, 4 per phase op, 2 + 32 per table)
static size_t tile_coef_bound(const TileLaunch& tl) {
    size_t n = 0;
    for (const DOp& d : tl.dops) n += d.kind == WK_TABLE ? 34 : 4;
    return n;
}
static bool jit_wanted(const qi_state* s) {
    const Context& c = ctx();
    return c.opt_jit >= 2 || (c.opt_jit == 1 && (int)s->n_local >= c.opt_jit_min_qubits);
}
// `arena_copy`: filled (once per circuit execution) the first time a new structure needs the tables on a worker thread
static int prepare_tile_jit(const qi_state* s, const TileLaunch& tl, const std::vector<amp_t>& arena, std::shared_ptr<const std::vector<amp_t>>* arena_copy, TileJit* out) {
    Context& c = ctx();
    if (!jit_wanted(s) || !jit::driver().ok) return QI_OK;
    int ctas = (c.opt_jit_ctas >= 3 && c.opt_jit_ctas <= 6) ? c.opt_jit_ctas : 4;
    const uint64_t ntiles = s->len >> kTileBits;
    int groups = c.opt_jit_groups == 4 ? 4 : (c.opt_jit_groups == 2 ? 2 : 1);
    while ((uint64_t)groups > ntiles) groups >>= 1;
    if (ctas != 4) groups = 1;
    const int pf = c.opt_jit_prefetch;
    const int stage = c.opt_jit_stage ? 1 : 0;
    if (stage) { groups = 1; ctas = 3; }
    const uint64_t key = jit::structure_key(tl, arena.data(), ctas, groups, pf, stage) ^ (0x9e3779b97f4a7c15ull * (uint64_t)std::max(0, c.opt_jit_smem_kb));
    jit::Entry* e = jit::find(key);
    if (!e) {
        if (tile_coef_bound(tl) * 8 + 64 > 32000) return QI_OK;      // parameter space: leave this launch to k_tile
        if (!*arena_copy) *arena_copy = std::make_shared<const std::vector<amp_t>>(arena);
        jit::Job job;
        job.tl = std::make_shared<const TileLaunch>(tl);
        job.arena = *arena_copy;
        job.ctas = ctas; job.groups = groups; job.prefetch = pf; job.stage = stage;
        const unsigned smem = std::max<unsigned>(jit::smem_bytes(groups, (int)tl.rounds.size(), stage), (unsigned)std::max(0, c.opt_jit_smem_kb) * 1024u);
        e = jit::enqueue(key, std::move(job), c.device, smem);
    }
    out->e = e;
    out->ctas = ctas; out->groups = groups; out->prefetch = pf; out->stage = stage;
    return QI_OK;
}
// this execution's coefficient block, in the order the module's text reads it (dry run of the generator)
static int collect_tile_coef(const TileLaunch& tl, const amp_t* arena, TileJit* tj) {
    if (!tj->e || tj->ready || tj->e->state.load(std::memory_order_acquire) != 1) return QI_OK;
    QI_TRY(jit::generate(tl, arena, tj->ctas, tj->groups, tj->prefetch, tj->stage, nullptr, &tj->coef, &tj->fp64));
    tj->ready = true;
    return QI_OK;
}
// launches the pass's module when it is ready; *launched = false leaves the pass to k_tile
static int launch_tile_jit(qi_state* s, const TileJit& tj, const amp_t* d_tables, bool* launched) {
    *launched = false;
    if (!tj.e || !tj.ready) return QI_OK;       // (a module that finished assembling after the coefficients were collected waits for the next execution)
    Context& c = ctx();
    uint64_t ntiles = s->len >> kTileBits;
    const uint64_t G = (uint64_t)tj.e->groups;
    const uint64_t blocks = std::min<uint64_t>(ntiles / G, (uint64_t)c.sm_count * 32 / G);
    static const double zero = 0.0;
    void* params[] = {(void*)&s->d, (void*)&ntiles, (void*)&d_tables, tj.coef.empty() ? (void*)&zero : (void*)tj.coef.data()};
    LaunchScope ls(KF_TILE_JIT, 32.0 * (double)s->len);
    const CUresult r = jit::driver().LaunchKernel(tj.e->fn, (unsigned)blocks, 1, 1, (unsigned)(kTileThreads * G), 1, 1, tj.e->smem, (CUstream)c.stream, params, nullptr);
    if (r != CUDA_SUCCESS) return fail(QI_ERR_CUDA, (uint64_t)r, 0, "cuLaunchKernel failed for a JIT tile module");
    *launched = true;
    {
        jit::Cache& jc = jit::cache();
        std::lock_guard<std::mutex> lk(jc.mu);
        jc.fp64_warp_instr += tj.fp64 * (kTileThreads / 32) * (double)ntiles;
    }
    return check_launch("qi_tile_jit");
}

template <int R>
static int launch_program(qi_state* s, const Layout& L, const DOp* ops, size_t nops, const amp_t* d_tables) {
    Context& c = ctx();
    WProgram<R> P;
    memset(&P, 0, sizeof(P));
    fill_offsets<R>(L, &P.ins, P.off);
    P.tables = d_tables;
    const uint64_t ntiles = s->len >> (kLaneQubits + R);
    const int warps_per_block = 4;
    uint64_t blocks = (ntiles + warps_per_block - 1) / warps_per_block;
    const uint64_t cap = (uint64_t)c.sm_count * 5 * 8;      // several waves of resident blocks per SM
    if (blocks > cap) blocks = cap;
    for (size_t first = 0; first < nops; first += kMaxOps) {
        size_t cnt = std::min<size_t>(kMaxOps, nops - first);
        P.nops = (uint32_t)cnt;
        P.flags = c.opt_prefetch ? 1u : 0u;
        memcpy(P.ops, ops + first, cnt * sizeof(DOp));
        // instantiation by content: programs without lane gates / without complex 2x2 gates run kernels that do not
        // carry that code (smaller, fewer live registers)
        bool lanes = false, u2k = false, lean = false;
        for (size_t k = 0; k < cnt; k++) {
            const DOp& o = P.ops[k];
            if (o.kind >= WK_X && o.kind <= kLastPairKind) { lanes |= o.tpos < kLaneQubits; u2k |= o.kind == WK_U2; lean |= o.kind >= WK_REALUP; }
            lean |= o.kind == WK_SCALE;
        }
        LaunchScope ls(KF_WINDOW, 32.0 * (double)s->len);
        const unsigned threads = warps_per_block * 32;
        if (lean) {      // programs with lean forms (option "lean"): their own instantiations, the default ones stay as they are
            if (lanes && u2k) k_window<R, true, true, true><<<(unsigned)blocks, threads, 0, c.stream>>>(s->d, ntiles, P);
            else if (lanes) k_window<R, true, false, true><<<(unsigned)blocks, threads, 0, c.stream>>>(s->d, ntiles, P);
            else if (u2k) k_window<R, false, true, true><<<(unsigned)blocks, threads, 0, c.stream>>>(s->d, ntiles, P);
            else k_window<R, false, false, true><<<(unsigned)blocks, threads, 0, c.stream>>>(s->d, ntiles, P);
        } else if (c.opt_tma) {
            uint64_t pblocks = (uint64_t)c.sm_count * QI_WINDOW_BLOCKS(R);       // persistent: every block resident
            if (pblocks > blocks) pblocks = blocks;
            const size_t smem = (size_t)warps_per_block * (32u << R) * sizeof(amp_t) + warps_per_block * sizeof(uint64_t);
            static bool configured = false;      // per instantiation of launch_program
            if (!configured) {
                QI_CUDA(cudaFuncSetAttribute(k_window_tma<R, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                QI_CUDA(cudaFuncSetAttribute(k_window_tma<R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                QI_CUDA(cudaFuncSetAttribute(k_window_tma<R, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                QI_CUDA(cudaFuncSetAttribute(k_window_tma<R, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured = true;
            }
            if (lanes && u2k) k_window_tma<R, true, true><<<(unsigned)pblocks, threads, smem, c.stream>>>(s->d, ntiles, P);
            else if (lanes) k_window_tma<R, true, false><<<(unsigned)pblocks, threads, smem, c.stream>>>(s->d, ntiles, P);
            else if (u2k) k_window_tma<R, false, true><<<(unsigned)pblocks, threads, smem, c.stream>>>(s->d, ntiles, P);
            else k_window_tma<R, false, false><<<(unsigned)pblocks, threads, smem, c.stream>>>(s->d, ntiles, P);
        } else {
            if (lanes && u2k) k_window<R, true, true><<<(unsigned)blocks, threads, 0, c.stream>>>(s->d, ntiles, P);
            else if (lanes) k_window<R, true, false><<<(unsigned)blocks, threads, 0, c.stream>>>(s->d, ntiles, P);
            else if (u2k) k_window<R, false, true><<<(unsigned)blocks, threads, 0, c.stream>>>(s->d, ntiles, P);
            else k_window<R, false, false><<<(unsigned)blocks, threads, 0, c.stream>>>(s->d, ntiles, P);
        }
        QI_TRY(check_launch("k_window"));
    }
    return QI_OK;
}

// staging for phase tables: pinned host buffer + device buffer, reused across calls
static int ensure_tables(size_t count) {
    Context& c = ctx();
    if (!c.ops_event) QI_CUDA(cudaEventCreateWithFlags(&c.ops_event, cudaEventDisableTiming));
    if (c.ops_cap >= count) return QI_OK;
    QI_CUDA(cudaStreamSynchronize(c.stream));
    if (c.h_ops) cudaFreeHost(c.h_ops);
    if (c.d_ops) cudaFree(c.d_ops);
    c.h_ops = c.d_ops = nullptr;
    size_t cap = count < (1u << 16) ? (1u << 16) : count * 2;
    QI_CUDA(cudaMallocHost(&c.h_ops, cap * sizeof(amp_t)));
    QI_CUDA(cudaMalloc(&c.d_ops, cap * sizeof(amp_t)));
    c.ops_cap = cap;
    return QI_OK;
}

struct Step { bool simple; PhysGate sgate; Pass pass; int R; TilePlan plan; };

// a fixed bit at position 0 makes the per-gate kernel touch every other amplitude: half of every 32-byte
// sector is wasted and a full window pass is faster (measured: 6.6 ms vs 5.5 ms at 30 qubits)
static bool fixes_bit0(const PhysGate& g) {
    return (g.cmask & 1ull) || (g.kind == IK_DIAG && g.t0 == 0);
}

// fraction of the amplitudes a gate can change (f)
static double touched_fraction(const PhysGate& g) {
    double f = 1.0 / (double)(1ull << __builtin_popcountll(g.cmask));
    if (g.kind == IK_DIAG && g.t0 >= 0) f *= 0.5;
    if (g.kind == IK_SWAP) f *= 0.5;
    return f;
}

// greedy pass construction (host only, no device access)
static int schedule_passes(const qi_state* s, const std::vector<PhysGate>& gates, bool fuse, int R, std::vector<Step>& steps) {
    const size_t G = gates.size();
    std::vector<char> done(G, 0);
    size_t first = 0;           // first gate not yet scheduled
    const size_t kLookahead = 4096;
    while (first < G) {
        if (done[first]) { first++; continue; }
        if (!window_takes(gates[first])) {
            steps.push_back(Step{true, gates[first], Pass(), R, TilePlan()});
            done[first++] = 1;
            continue;
        }
        Pass ps;
        uint64_t blocked_any = 0, blocked_n = 0, window_mask = 0;
        size_t scanned = 0, taken = 0, last_taken = 0;
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
                if ((int)ps.regs.size() + __builtin_popcountll(need) > R) take = false;
                else
                    for (int q = kLaneQubits; q < 64 && need; q++)
                        if ((need >> q) & 1) { ps.regs.push_back(q); window_mask |= 1ull << q; need &= ~(1ull << q); }
            }
            if (take) {
                lower_gate(ps, g, fuse);
                done[i] = 1;
                taken++;
                last_taken = i;
                if (!fuse) break;
            } else {
                blocked_any |= u.n_use | u.d_use;
                blocked_n |= u.n_use;
            }
        }
        if (ps.ops.empty()) return fail(QI_ERR_UNKNOWN, 0, 0, "scheduler made no progress");
        if (taken == 1 && touched_fraction(gates[last_taken]) <= 0.5 && !fixes_bit0(gates[last_taken])) {
            // a lone gate that can change at most half of the amplitudes: the per-gate kernel visits only
            // those (controls and the phase target are folded into its index expansion) and beats a full pass
            steps.push_back(Step{true, gates[last_taken], Pass(), R, TilePlan()});
            continue;
        }
        if (fuse && ctx().opt_late_tables) repack_unconditional_tables(ps);
        steps.push_back(Step{false, PhysGate(), std::move(ps), R, TilePlan()});   // (measured: the 8-amplitude kernel streams ~6% slower than R = 4)
    }
    return QI_OK;
}

// ---- pass construction for the CTA-tile kernel ------------------------------------------------------------------
// A tile pass works on 11 qubits: whatever sits at physical positions 0..4 (coalescing) plus six more.  Two things make it
// hold far more of a circuit than "first come" window allocation:
//  * the tile is CHOSEN: for every seed qubit a candidate tile is grown breadth-first over the interaction graph of the
//    gates still to run (qubits that share a gate are neighbours), the greedy gate selection is simulated for each
//    candidate and the one that takes the most gates wins -- on a nearest-neighbour circuit the winners are windows of
//    contiguous qubits, which a pass digs into as a trapezoid of layers;
//  * the five LOW positions are a cache, not a fixed set of qubits (`permute`): a pass stores its tile with the 11 local
//    bits in any order at no cost (the last regroup writes whatever layout it likes), so the five of its qubits that the
//    NEXT tile wants are left at positions 0..4 and the tile slides over the register: tile k+1 only has to share five
//    qubits with tile k.  The relabelling is folded into the state's logical -> physical map afterwards (like a lazy SWAP).
// Gates are given in the physical coordinates at entry ("q"); pos[q] tracks where q lives now, and every pass / simple
// step is translated to the positions current when it runs.
struct TileSched {
    int n = 0;
    std::vector<int> pos, at;            // q -> physical position now; position -> q
};

static uint64_t remap_mask(const std::vector<int>& pos, uint64_t m) {
    uint64_t o = 0;
    for (int q = 0; m; q++, m >>= 1) if (m & 1) o |= 1ull << pos[q];
    return o;
}

static void remap_gate(const std::vector<int>& pos, PhysGate* g) {
    if (g->t0 >= 0) g->t0 = pos[g->t0];
    if (g->t1 >= 0) g->t1 = pos[g->t1];
    g->cmask = remap_mask(pos, g->cmask);
}

static void remap_pass(const std::vector<int>& pos, Pass* ps) {
    for (HOp& h : ps->ops) {
        if (h.kind == 0) continue;
        if (h.target >= 0) h.target = pos[h.target];
        h.cmask = remap_mask(pos, h.cmask);
        h.nmask = remap_mask(pos, h.nmask);
        h.tmask = remap_mask(pos, h.tmask);
    }
    for (DiagGroup& g : ps->groups) {
        if (g.hub >= 0) g.hub = pos[g.hub];
        if (g.hub_alt >= 0) g.hub_alt = pos[g.hub_alt];
        for (int& q : g.bits) q = pos[q];
        g.blocked_since = remap_mask(pos, g.blocked_since);
    }
    for (int& q : ps->regs) q = pos[q];
}

static int schedule_tile_passes(const qi_state* s, const std::vector<PhysGate>& gates, bool fuse, bool permute, std::vector<Step>& steps,
                                std::vector<int>* final_pos, bool restore = false) {
    const size_t G = gates.size();
    const int n = (int)s->n_local;
    TileSched ts;
    ts.n = n;
    ts.pos.resize(64); ts.at.resize(64);
    for (int q = 0; q < 64; q++) ts.pos[q] = ts.at[q] = q;
    std::vector<char> done(G, 0);
    std::vector<GateUse> use(G);
    for (size_t i = 0; i < G; i++) use[i] = uses_of(gates[i]);
    size_t first = 0;
    const size_t kLookahead = 4096, kGraph = 512;
    int prev_step = -1;                       // last tile step whose output layout is still open
    std::vector<int> prev_tile;               // its qubits (q coordinates)

    // light simulation of the greedy selection for a candidate tile: gates it would take (non-diagonal ones count 1, diagonal 1/4)
    auto score_tile = [&](uint64_t tmask) {
        uint64_t blocked_any = 0, blocked_n = 0;
        size_t scanned = 0;
        int score4 = 0;
        for (size_t i = first; i < G && scanned < kLookahead; i++) {
            if (done[i]) continue;
            scanned++;
            if (!window_takes(gates[i])) break;
            const GateUse& u = use[i];
            if ((u.n_use & blocked_any) == 0 && (u.d_use & blocked_n) == 0 && (u.n_use & ~tmask) == 0) score4 += u.n_use ? 4 : 1;
            else { blocked_any |= u.n_use | u.d_use; blocked_n |= u.n_use; if ((blocked_n & tmask) == tmask) break; }
        }
        return score4;
    };

    while (first < G) {
        if (done[first]) { first++; continue; }
        if (!window_takes(gates[first])) {
            prev_step = -1;                                   // the open layout stays as it is
            PhysGate g = gates[first];
            remap_gate(ts.pos, &g);
            steps.push_back(Step{true, g, Pass(), kTileWindow, TilePlan()});
            done[first++] = 1;
            continue;
        }
        uint64_t resident = 0;                                // q's at positions 0..4
        for (int p = 0; p < kLaneQubits; p++) resident |= 1ull << ts.at[p];
        uint64_t prev_mask = 0;
        for (int q : prev_tile) prev_mask |= 1ull << q;
        const bool slide = permute && prev_step >= 0;         // this tile may pick its five low qubits among the previous tile's
        // interaction graph of the gates ahead
        std::vector<uint64_t> adj(n, 0);
        {
            size_t scanned = 0;
            for (size_t i = first; i < G && scanned < kGraph; i++) {
                if (done[i]) continue;
                scanned++;
                const uint64_t m = (use[i].n_use | use[i].d_use);
                for (int q = 0; q < n; q++) if ((m >> q) & 1) adj[q] |= m;
            }
        }
        // neighbours are visited in the order in which the gates ahead first act NON-diagonally on them (program order), not by
        // qubit index: for a QFT in the bit-reversed layout its own swaps leave behind, index order fills the tile with the
        // qubits whose Hadamards come LAST (12 passes instead of 5)
        std::vector<int> by_use(n);
        {
            std::vector<size_t> first_use(n, (size_t)-1);
            size_t scanned = 0;
            for (size_t i = first; i < G && scanned < kGraph; i++) {
                if (done[i]) continue;
                scanned++;
                for (int q = 0; q < n; q++) if (((use[i].n_use >> q) & 1) && first_use[q] == (size_t)-1) first_use[q] = i;
            }
            for (int q = 0; q < n; q++) by_use[q] = q;
            if (ctx().opt_tile_bfs_by_use)
                std::stable_sort(by_use.begin(), by_use.end(), [&](int a, int b) { return first_use[a] < first_use[b]; });
        }
        const uint64_t need_first = use[first].n_use;         // progress: the oldest gate must fit
        uint64_t best_tile = 0;
        int best_score = -1;
        for (int seed = 0; seed < n; seed++) {
            // breadth-first order from the seed (the oldest gate's targets first)
            std::vector<int> order;
            uint64_t seen = 0;
            for (int q = 0; q < n; q++) if ((need_first >> q) & 1) { order.push_back(q); seen |= 1ull << q; }
            if (!((seen >> seed) & 1)) { order.push_back(seed); seen |= 1ull << seed; }
            for (size_t k = 0; k < order.size() && (int)order.size() < n; k++) {
                uint64_t nb = adj[order[k]] & ~seen;
                for (int qi = 0; qi < n && nb; qi++) { const int q = by_use[qi]; if ((nb >> q) & 1) { order.push_back(q); seen |= 1ull << q; nb &= ~(1ull << q); } }
            }
            for (int d = 1; (int)order.size() < n && d < n; d++) {        // isolated qubits: nearest positions first
                for (int sgn = -1; sgn <= 1; sgn += 2) {
                    const int q = seed + sgn * d;
                    if (q >= 0 && q < n && !((seen >> q) & 1)) { order.push_back(q); seen |= 1ull << q; }
                }
            }
            uint64_t tile = 0;
            int cnt = 0;
            if (!slide) {
                tile = resident; cnt = kLaneQubits;
                for (int q : order) { if (cnt >= kTileBits) break; if (!((tile >> q) & 1)) { tile |= 1ull << q; cnt++; } }
            } else {
                // eleven qubits in breadth-first order, at least five of them from the previous tile
                int from_prev = 0;
                std::vector<int> chosen;
                for (int q : order) {
                    if ((int)chosen.size() >= kTileBits) break;
                    const bool in_prev = (prev_mask >> q) & 1;
                    const int left = kTileBits - (int)chosen.size();
                    if (!in_prev && left <= kLaneQubits - from_prev) continue;      // keep room for the previous tile's share
                    chosen.push_back(q);
                    from_prev += in_prev;
                }
                for (int q : chosen) tile |= 1ull << q;
                cnt = (int)chosen.size();
            }
            if (cnt < kTileBits || (need_first & ~tile)) continue;
            if (tile == best_tile) continue;
            const int sc = fuse ? score_tile(tile) : 1;
            if (sc > best_score) { best_score = sc; best_tile = tile; }
            if (!fuse) break;
        }
        if (best_score < 0) return fail(QI_ERR_UNKNOWN, 0, 0, "tile scheduler found no tile");
        // the five low qubits of this pass; a sliding tile takes them from the previous tile, whose output layout is fixed now
        if (slide) {
            std::vector<int> low;
            for (int p = 0; p < kLaneQubits; p++) if ((best_tile >> ts.at[p]) & 1) low.push_back(ts.at[p]);      // already low: stay
            for (int q : prev_tile) if ((int)low.size() < kLaneQubits && ((best_tile >> q) & 1) && ts.pos[q] >= kLaneQubits) low.push_back(q);
            if ((int)low.size() < kLaneQubits) return fail(QI_ERR_UNKNOWN, 0, 0, "tile scheduler: fewer than five shared qubits");
            uint64_t low_mask = 0;
            for (int q : low) low_mask |= 1ull << q;
            // evicted low qubits trade places with the newly low ones
            std::vector<int> evicted, incoming;
            for (int p = 0; p < kLaneQubits; p++) if (!((low_mask >> ts.at[p]) & 1)) evicted.push_back(ts.at[p]);
            for (int q : low) if (ts.pos[q] >= kLaneQubits) incoming.push_back(q);
            TilePlan& pl = steps[prev_step].plan;
            for (size_t k = 0; k < evicted.size(); k++) {
                const int qa = evicted[k], qb = incoming[k];
                const int pa = ts.pos[qa], pb = ts.pos[qb];
                for (int j = 0; j < kTileBits; j++) {           // both are in the previous tile: swap their destinations
                    if (pl.pout[j] == pa) pl.pout[j] = pb;
                    else if (pl.pout[j] == pb) pl.pout[j] = pa;
                }
                ts.pos[qa] = pb; ts.pos[qb] = pa;
                ts.at[pa] = qb; ts.at[pb] = qa;
            }
        }
        // the real selection
        Pass ps;
        ps.absorb = ctx().opt_tile_absorb != 0;
        uint64_t blocked_any = 0, blocked_n = 0;
        size_t scanned = 0, taken = 0, last_taken = 0;
        for (size_t i = first; i < G && scanned < kLookahead; i++) {
            if (done[i]) continue;
            scanned++;
            const PhysGate& g = gates[i];
            if (!window_takes(g)) break;
            const GateUse& u = use[i];
            const bool take = ((u.n_use & blocked_any) == 0) && ((u.d_use & blocked_n) == 0) && (u.n_use & ~best_tile) == 0;
            if (take) {
                lower_gate(ps, g, fuse);
                done[i] = 1;
                taken++;
                last_taken = i;
                if (!fuse) break;
            } else {
                blocked_any |= u.n_use | u.d_use;
                blocked_n |= u.n_use;
            }
        }
        if (taken == 0) return fail(QI_ERR_UNKNOWN, 0, 0, "tile scheduler made no progress");
        if (taken == 1 && touched_fraction(gates[last_taken]) <= 0.5) {
            PhysGate g = gates[last_taken];
            remap_gate(ts.pos, &g);
            if (!fixes_bit0(g)) {       // lone gate on at most half of the amplitudes: the per-gate kernel (see schedule_passes)
                prev_step = -1;
                steps.push_back(Step{true, g, Pass(), kTileWindow, TilePlan()});
                continue;
            }
        }
        if (fuse && ctx().opt_late_tables) repack_unconditional_tables(ps);
        remap_pass(ts.pos, &ps);
        Step st{false, PhysGate(), std::move(ps), kTileWindow, TilePlan()};
        std::vector<int> ppos;
        prev_tile.clear();
        for (int q = 0; q < n; q++) if ((best_tile >> q) & 1) { ppos.push_back(ts.pos[q]); prev_tile.push_back(q); }
        std::sort(ppos.begin(), ppos.end());
        for (int j = 0; j < kTileBits; j++) st.plan.pin[j] = st.plan.pout[j] = ppos[j];
        steps.push_back(std::move(st));
        prev_step = (int)steps.size() - 1;
    }
    // ---- restore the layout the run started with (option tile_restore; states that run on modules) -------------------------
    // Sliding tiles leave the qubits permuted.  A later execution of the same circuit would then be scheduled from another
    // layout and share no pass structure -- no module -- with this one.  The permutation is undone here: first inside the
    // last tile, whose output order is still open, then by relabel-only passes (no ops: one HBM pass each) over positions
    // 0..4 plus six chosen positions; every such pass sends home each qubit whose home position is in its tile.
    if (permute && restore) {
        auto displaced = [&]() { int d = 0; for (int q = 0; q < n; q++) d += ts.pos[q] != q; return d; };
        // place the qubits of a tile (given by its 11 positions): home if the home is in the tile, the free positions otherwise
        auto settle = [&](TilePlan& pl) {
            // pl.pin = positions of the tile before the pass; the pass may send the content of pin[j] to any pout[j] (a permutation of pin).
            // Work on the CURRENT placement: position -> qubit (ts.at), restricted to the tile's positions.
            std::vector<int> tile_pos(pl.pout, pl.pout + kTileBits);          // current positions of the tile's content (after earlier swaps)
            uint64_t tmask = 0;
            for (int p : tile_pos) tmask |= 1ull << p;
            std::vector<int> qubits;
            for (int p : tile_pos) qubits.push_back(ts.at[p]);
            std::vector<int> new_pos(qubits.size(), -1);
            uint64_t used = 0;
            for (size_t k = 0; k < qubits.size(); k++)
                if ((tmask >> qubits[k]) & 1) { new_pos[k] = qubits[k]; used |= 1ull << qubits[k]; }      // home is in the tile
            for (size_t k = 0; k < qubits.size(); k++) {
                if (new_pos[k] >= 0) continue;
                if (!((used >> tile_pos[k]) & 1)) { new_pos[k] = tile_pos[k]; used |= 1ull << tile_pos[k]; }   // stay if the place is free
            }
            for (size_t k = 0; k < qubits.size(); k++) {
                if (new_pos[k] >= 0) continue;
                for (int p : tile_pos) if (!((used >> p) & 1)) { new_pos[k] = p; used |= 1ull << p; break; }
            }
            for (int j = 0; j < kTileBits; j++)
                for (size_t k = 0; k < tile_pos.size(); k++)
                    if (pl.pout[j] == tile_pos[k]) { pl.pout[j] = new_pos[k]; break; }
            for (size_t k = 0; k < qubits.size(); k++) { ts.pos[qubits[k]] = new_pos[k]; ts.at[new_pos[k]] = qubits[k]; }
        };
        if (prev_step >= 0 && displaced()) settle(steps[prev_step].plan);
        int guard = 0;
        while (displaced() && guard++ < 64) {
            // six high positions: follow the homes of displaced qubits, starting with those that sit at positions 0..4
            std::vector<int> chosen;
            uint64_t cm = 0;
            auto add = [&](int p) { if (p >= kLaneQubits && p < n && !((cm >> p) & 1) && (int)chosen.size() < kTileBits - kLaneQubits) { chosen.push_back(p); cm |= 1ull << p; return true; } return false; };
            for (int p = 0; p < kLaneQubits; p++) if (ts.at[p] != p) add(ts.at[p]);                  // homes of the low sitters
            for (size_t k = 0; k < chosen.size(); k++) { const int q = ts.at[chosen[k]]; if (q != chosen[k]) add(q); }   // ... and of whoever sits there
            for (int p = kLaneQubits; p < n && (int)chosen.size() < kTileBits - kLaneQubits; p++)
                if (ts.at[p] != p && add(p))
                    for (size_t k = chosen.size() - 1; k < chosen.size(); k++) { const int q = ts.at[chosen[k]]; if (q != chosen[k]) add(q); }
            for (int p = kLaneQubits; p < n && (int)chosen.size() < kTileBits - kLaneQubits; p++) add(p);     // fill up
            if ((int)chosen.size() < kTileBits - kLaneQubits) return fail(QI_ERR_UNKNOWN, 0, 0, "tile scheduler: cannot form a relabel pass");
            Step st{false, PhysGate(), Pass(), kTileWindow, TilePlan()};
            std::vector<int> ppos;
            for (int p = 0; p < kLaneQubits; p++) ppos.push_back(p);
            for (int p : chosen) ppos.push_back(p);
            std::sort(ppos.begin(), ppos.end());
            for (int j = 0; j < kTileBits; j++) st.plan.pin[j] = st.plan.pout[j] = ppos[j];
            const int before = displaced();
            settle(st.plan);
            if (displaced() >= before) return fail(QI_ERR_UNKNOWN, 0, 0, "tile scheduler: relabel pass made no progress");
            steps.push_back(std::move(st));
        }
        if (displaced()) return fail(QI_ERR_UNKNOWN, 0, 0, "tile scheduler: layout not restored");
    }
    if (final_pos) { final_pos->assign(ts.pos.begin(), ts.pos.end()); }
    return QI_OK;
}

// ---- peephole: a controlled X next to a Hadamard on its target becomes a controlled Z --------------------------------
// H X = Z H and X H = H Z hold bit for bit in the reference's arithmetic (X only permutes, Z only negates, and
// s*(a1 + a0) == s*(a0 + a1)), so  [C..X(t), H(t)] == [H(t), C..Z(t)]  and  [H(t), C..X(t)] == [C..Z(t), H(t)]  exactly.
// The controlled Z is diagonal: it merges into the phase table the pass carries anyway, whereas the X costs a register
// swap op of its own or (absorbed) turns the neighbouring gate into two predicated half-populated ops.  An X also slides
// past RX-form 2x2 gates on its target (RX X = X RX bit for bit: both orders evaluate the same products).  The neighbour
// need not be adjacent in the list: the next (previous) gate that touches the target commutes with everything between
// it and the X, so it is pulled next to the X first.  Random H/RX/RZ + CNOT layers: 5 of 9 CNOTs are rewritten.
static bool is_rx_form(const PhysGate& g) {
    const double* p = g.p;
    return g.kind == IK_U2 && g.cmask == 0 && p[1] == 0.0 && p[2] == 0.0 && p[4] == 0.0 && p[7] == 0.0 && p[0] == p[6] && p[3] == p[5];
}

static void rewrite_cx_next_to_h(std::vector<PhysGate>& gates) {
    const int G = (int)gates.size();
    if (G < 2) return;
    const int kWindow = 256;                       // gates scanned on either side of an X
    std::vector<PhysGate> pool(gates);             // nodes 0..G-1: original gates; node G: list sentinel; then the inserted controlled-Z gates
    pool.reserve(G + 1 + G / 2);
    pool.push_back(PhysGate());
    std::vector<int> next(G + 1), prev(G + 1);     // doubly linked list through the sentinel
    for (int i = 0; i <= G; i++) { next[i] = i == G ? 0 : i + 1; prev[i] = i == 0 ? G : i - 1; }
    auto unlink = [&](int i) { next[prev[i]] = next[i]; prev[next[i]] = prev[i]; };
    auto link_after = [&](int at, int id) { next[id] = next[at]; prev[id] = at; prev[next[at]] = id; next[at] = id; };
    auto insert_after = [&](int at, const PhysGate& g) {
        const int id = (int)pool.size();
        pool.push_back(g);
        next.push_back(0); prev.push_back(0);
        link_after(at, id);
    };
    bool any = false;
    for (int i = 0; i < G; i++) {
        const PhysGate x = pool[i];
        if (x.kind != IK_X) continue;
        const uint64_t tb = 1ull << x.t0, cm = x.cmask;
        PhysGate cz;
        memset(&cz, 0, sizeof(cz));
        cz.kind = IK_DIAG; cz.t0 = x.t0; cz.t1 = -1; cz.cmask = cm; cz.p[0] = -1.0; cz.p[1] = 0.0;
        bool done = false;
        // forward: the next gate that touches the target.  Nothing between the X and that gate touches the target, so the gate
        // (a single-qubit gate on the target alone) commutes with everything in between and may be pulled back to the X:
        //   H:        [X, .., H]  = [X, H, ..]  = [H, CZ, ..]
        //   RX-form:  [X, .., RX] = [X, RX, ..] = [RX, X, ..]   and the scan goes on from the X's new place
        int j = next[i];
        for (int k = 0; k < kWindow && j != G; k++) {
            const PhysGate& g = pool[j];
            const GateUse u = uses_of(g);
            const int nj = next[j];
            if ((u.n_use | u.d_use) & tb) {
                if (g.kind == IK_H && g.cmask == 0 && g.t0 == x.t0) {
                    unlink(j); link_after(prev[i], j);          // H right before the X ...
                    insert_after(j, cz); unlink(i);             // ... and the X becomes the CZ behind it
                    done = any = true;
                } else if (is_rx_form(g) && g.t0 == x.t0) {
                    unlink(j); link_after(prev[i], j);          // RX hops over the X
                    any = true;
                    j = nj;
                    continue;
                }
                break;
            }
            j = nj;
        }
        if (done) continue;
        // backward, mirrored: [H, .., X] = [.., H, X] = [.., CZ, H];  [RX, .., X] = [.., X, RX]
        j = prev[i];
        for (int k = 0; k < kWindow && j != G; k++) {
            const PhysGate& g = pool[j];
            const GateUse u = uses_of(g);
            const int pj = prev[j];
            if ((u.n_use | u.d_use) & tb) {
                if (g.kind == IK_H && g.cmask == 0 && g.t0 == x.t0) {
                    unlink(j); link_after(i, j);                // H right behind the X ...
                    insert_after(prev[i], cz); unlink(i);       // ... and the X becomes the CZ in front of it
                    any = true;
                } else if (is_rx_form(g) && g.t0 == x.t0) {
                    unlink(j); link_after(i, j);
                    any = true;
                    j = pj;
                    continue;
                }
                break;
            }
            j = pj;
        }
    }
    if (!any) return;
    std::vector<PhysGate> out;
    out.reserve(pool.size());
    for (int i = next[G]; i != G; i = next[i]) out.push_back(pool[i]);
    gates.swap(out);
}

// passes run on the CTA-tile kernel (k_tile) when the state has enough local qubits; option "tile" = 0 keeps the warp-tile kernel
// Runs of one or two gates stay on the warp-tile kernel: a lone gate is a pure HBM pass, and k_window streams it at 0.97-0.99 of
// the measured copy bandwidth at every target qubit, the tile kernels at 0.77-1.0 depending on how many CTAs fit an SM
// (profiles/r02_single_gate_executors.txt).
static bool tile_mode(const qi_state* s, size_t ngates) {
    const Context& c = ctx();
    return c.opt_tile && !c.opt_tma && (int)s->n_local >= std::max(kTileBits, c.opt_tile_min_qubits) && ngates >= (size_t)std::max(1, c.opt_tile_min_gates);
}

// `allow_relabel`: tile passes may leave the qubits of the state at other physical positions (folded into s->phys, like a
// lazy SWAP).  Callers that need the layout they came with (chunk views, shards, canonicalise) pass false.
static int run_circuit_windowed_impl(qi_state* s, const std::vector<PhysGate>& gates_in, bool allow_relabel, bool force_window);
int run_circuit_windowed(qi_state* s, const std::vector<PhysGate>& gates_in, bool allow_relabel) {
    return run_circuit_windowed_impl(s, gates_in, allow_relabel, false);
}
static int run_circuit_windowed_impl(qi_state* s, const std::vector<PhysGate>& gates_in, bool allow_relabel, bool force_window) {
    Context& c = ctx();
    std::vector<PhysGate> rewritten;
    if (c.opt_cz_rewrite && c.opt_fuse) { rewritten = gates_in; rewrite_cx_next_to_h(rewritten); }
    const std::vector<PhysGate>& gates = (c.opt_cz_rewrite && c.opt_fuse) ? rewritten : gates_in;
    const bool tile = !force_window && tile_mode(s, gates.size());
    const int R = tile ? kTileWindow : window_regs(s);
    std::vector<Step> steps;
    std::vector<int> final_pos;
    // Sliding tiles leave the state in another qubit order after every execution, so the next execution of the same circuit is
    // scheduled from another layout and shares no pass structure with this one: modules would never be reused.  States
    // that run on JIT modules therefore keep their layout (more, cheaper passes: the modules are FP64-bound, not HBM-bound).
    const bool restore = jit_wanted(s) && c.opt_tile_restore;       // sliding tiles whose permutation is undone at the end of the run
    const bool relabel = tile && allow_relabel && s->world == 1 && c.opt_tile_slide && (!jit_wanted(s) || restore);
    if (tile) QI_TRY(schedule_tile_passes(s, gates, c.opt_fuse != 0, relabel, steps, &final_pos, restore));
    else QI_TRY(schedule_passes(s, gates, c.opt_fuse != 0, R, steps));
    // lower every pass, upload all phase tables in one copy, then launch back to back
    std::vector<std::vector<DOp>> dops(steps.size());
    std::vector<Layout> layouts(steps.size());
    std::vector<std::vector<TileLaunch>> tiles(steps.size());
    std::vector<amp_t> arena;
    amp_t carry = make_double2(1.0, 0.0);
    for (size_t i = 0; i < steps.size(); i++) {
        if (steps[i].simple) continue;
        if (tile) {
            bool last = true;
            for (size_t k = i + 1; k < steps.size(); k++) last &= steps[k].simple;
            QI_TRY(lower_tile_pass(s, steps[i].pass, steps[i].plan, tiles[i], arena, &carry, last || !c.opt_tile_carry));
        }
        else lower_pass(s, steps[i].pass, steps[i].R, dops[i], arena, &layouts[i]);
    }
    if (!arena.empty()) {
        QI_TRY(ensure_tables(arena.size()));
        QI_CUDA(cudaEventSynchronize(c.ops_event));      // the previous run's copy has left the pinned buffer
        memcpy(c.h_ops, arena.data(), arena.size() * sizeof(amp_t));
        QI_CUDA(cudaMemcpyAsync(c.d_ops, c.h_ops, arena.size() * sizeof(amp_t), cudaMemcpyHostToDevice, c.stream));
        QI_CUDA(cudaEventRecord(c.ops_event, c.stream));
    }
    std::vector<std::vector<TileJit>> jits(steps.size());
    std::shared_ptr<const std::vector<amp_t>> arena_copy;
    if (tile && jit_wanted(s)) {
        for (size_t i = 0; i < steps.size(); i++) {
            jits[i].resize(tiles[i].size());
            for (size_t k = 0; k < tiles[i].size(); k++) QI_TRY(prepare_tile_jit(s, tiles[i][k], arena, &arena_copy, &jits[i][k]));
        }
        if (c.opt_jit >= 2) jit::drain();            // every module of this circuit is assembled (in parallel) before the first launch
        bool any_ready = false;
        for (size_t i = 0; i < steps.size(); i++)
            for (size_t k = 0; k < tiles[i].size(); k++) {
                QI_TRY(collect_tile_coef(tiles[i][k], arena.data(), &jits[i][k]));
                any_ready |= jits[i][k].ready;
            }
        // First sight of a circuit made mostly of diagonal gates (QFT: 92 % controlled phases): none of its modules exists yet
        // (they are being assembled now), and for such circuits the warp-tile kernel with its merged phase tables beats the
        // INTERPRETING tile kernel (33-qubit QFT: 450-517 ms vs 565-680 ms; the modules take 288 ms) -- this execution runs on it.
        if (!any_ready && c.opt_jit == 1 && !gates.empty()) {
            size_t diag = 0;
            for (const PhysGate& g : gates) diag += (g.kind == IK_DIAG || g.kind == IK_RZ) ? 1 : 0;
            if (10 * diag > 6 * gates.size()) return run_circuit_windowed_impl(s, gates_in, allow_relabel, true);
        }
    }
    for (size_t i = 0; i < steps.size(); i++) {
        if (steps[i].simple) QI_TRY(launch_simple_gate(s, steps[i].sgate));
        else if (tile) {
            for (size_t k = 0; k < tiles[i].size(); k++) {
                bool launched = false;
                if (!jits[i].empty()) QI_TRY(launch_tile_jit(s, jits[i][k], (const amp_t*)c.d_ops, &launched));
                if (!launched) QI_TRY(launch_tile(s, tiles[i][k], (const amp_t*)c.d_ops));
            }
        }
        else if (steps[i].R == 3) QI_TRY(launch_program<3>(s, layouts[i], dops[i].data(), dops[i].size(), (const amp_t*)c.d_ops));
        else if (steps[i].R == 4) QI_TRY(launch_program<4>(s, layouts[i], dops[i].data(), dops[i].size(), (const amp_t*)c.d_ops));
        else QI_TRY(launch_program<5>(s, layouts[i], dops[i].data(), dops[i].size(), (const amp_t*)c.d_ops));
    }
    if (relabel)                  // where the tile passes left the qubits
        for (uint32_t q = 0; q < s->num_qubits; q++) s->phys[q] = (uint8_t)final_pos[s->phys[q]];
    return QI_OK;
}

// host-only: schedule a run of physical gates and report the passes (tuning / tests; no device access)
int debug_schedule(const qi_state* s, const std::vector<PhysGate>& gates, int R, std::vector<std::vector<int>>* summary) {
    std::vector<Step> steps;
    QI_TRY(schedule_passes(s, gates, true, R, steps));
    for (const Step& st : steps) {
        int lane_ops = 0, reg_ops = 0, diag_ops = 0, table_ops = 0, absorbed = 0;
        for (const HOp& h : st.pass.ops) {
            if (h.kind == 0) continue;
            if (h.nmask) absorbed++;
            if (h.kind == WK_TABLE) table_ops++;
            else if (h.kind == WK_DIAG || h.kind == WK_RZ) diag_ops++;
            else if (h.target < kLaneQubits) lane_ops++;
            else reg_ops++;
        }
        summary->push_back({st.simple ? 1 : 0, (int)st.pass.regs.size(), lane_ops, reg_ops, diag_ops, table_ops, absorbed, 0});
    }
    return QI_OK;
}

// host-only: schedule AND lower a run of physical gates; the device programs are serialised into `blob` so that a test
// can interpret them on the CPU (tests/test_window_lowering.py).  Layout (all fields 8-byte aligned):
//   u64 nsteps, then per step: u64 simple;
//     simple = 1: PhysGate (raw);   simple = 0: u64 R, u64 regs[8] (sorted window qubits), u64 nops, DOp[nops] (raw, 112 B each)
//   then u64 arena_count and arena_count double2 phase-table entries (DOp::m[0] of a table op is an offset into it)
//     simple = 2 (CTA-tile pass, one record per k_tile launch): u64 tile_pos[11], u64 tile_out[11] (position the content of local bit j
//                 is stored to), u64 nrounds, then per round
//                 u64 regs[4] (physical qubits of slot bits 0..3), u64 thr[7] (physical qubit of thread-index bit k), u64 nops, DOp[nops]
int debug_lower(const qi_state* s, const std::vector<PhysGate>& gates_in, int R, std::vector<uint8_t>* blob, std::vector<int>* relabel_out) {
    std::vector<PhysGate> gates(gates_in);
    if (ctx().opt_cz_rewrite && ctx().opt_fuse) rewrite_cx_next_to_h(gates);
    std::vector<Step> steps;
    const bool tile = tile_mode(s, gates.size());
    std::vector<int> final_pos;
    const bool restore = jit_wanted(s) && ctx().opt_tile_restore;
    const bool relabel = tile && s->world == 1 && ctx().opt_tile_slide && (!jit_wanted(s) || restore);
    if (tile) QI_TRY(schedule_tile_passes(s, gates, ctx().opt_fuse != 0, relabel, steps, &final_pos, restore));
    else QI_TRY(schedule_passes(s, gates, ctx().opt_fuse != 0, R, steps));
    std::vector<amp_t> arena;
    auto put = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; blob->insert(blob->end(), b, b + n); };
    auto put64 = [&](uint64_t v) { put(&v, 8); };
    put64(steps.size());
    if (tile) {
        uint64_t nrec = 0;
        amp_t carry = make_double2(1.0, 0.0);
        std::vector<std::vector<TileLaunch>> tiles(steps.size());
        for (size_t i = 0; i < steps.size(); i++) {
            if (steps[i].simple) { nrec++; continue; }
            bool last = true;
            for (size_t k = i + 1; k < steps.size(); k++) last &= steps[k].simple;
            QI_TRY(lower_tile_pass(s, steps[i].pass, steps[i].plan, tiles[i], arena, &carry, last || !ctx().opt_tile_carry));
            nrec += tiles[i].size();
        }
        blob->clear();
        if (ctx().opt_debug_ptx) {          // the PTX of every tile launch, separated by a marker line (assembled by the CPU tests with ptxas)
            for (size_t i = 0; i < steps.size(); i++)
                for (const TileLaunch& tl : tiles[i]) {
                    std::string text;
                    std::vector<double> coef;
                    double fp64 = 0.0;
                    QI_TRY(jit::generate(tl, arena.data(), ctx().opt_jit_stage ? 3 : ((ctx().opt_jit_ctas >= 3 && ctx().opt_jit_ctas <= 6) ? ctx().opt_jit_ctas : 4), (ctx().opt_jit_ctas != 4 || ctx().opt_jit_stage) ? 1 : std::max(1, ctx().opt_jit_groups), ctx().opt_jit_prefetch, ctx().opt_jit_stage ? 1 : 0, &text, &coef, &fp64));
                    char mark[64];
                    snprintf(mark, sizeof(mark), "//---PASS fp64=%.1f coef=%zu---\n", fp64, coef.size());
                    put(mark, strlen(mark));
                    put(text.data(), text.size());
                }
            return QI_OK;
        }
        put64(nrec);
        for (size_t i = 0; i < steps.size(); i++) {
            if (steps[i].simple) { put64(1); put(&steps[i].sgate, sizeof(PhysGate)); continue; }
            for (const TileLaunch& tl : tiles[i]) {
                put64(2);
                for (int j = 0; j < kTileBits; j++) put64((uint64_t)tl.tile_qubits[j]);
                for (int j = 0; j < kTileBits; j++) put64((uint64_t)tl.tile_out[j]);
                put64(tl.rounds.size());
                for (const TileRoundHost& r : tl.rounds) {
                    for (int k = 0; k < 4; k++) put64((uint64_t)tl.tile_qubits[r.regs[k]]);
                    for (int k = 0; k < kTileThrBits; k++) put64((uint64_t)tl.tile_qubits[r.thr[k]]);
                    put64(r.nops);
                    put(tl.dops.data() + r.first_op, r.nops * sizeof(DOp));
                }
            }
        }
        put64(arena.size());
        put(arena.data(), arena.size() * sizeof(amp_t));
        if (relabel_out && relabel) *relabel_out = final_pos;
        return QI_OK;
    }
    for (const Step& st : steps) {
        put64(st.simple ? 1 : 0);
        if (st.simple) { put(&st.sgate, sizeof(PhysGate)); continue; }
        std::vector<DOp> dops;
        Layout L;
        lower_pass(s, st.pass, st.R, dops, arena, &L);
        put64((uint64_t)st.R);
        for (int j = 0; j < 8; j++) put64(j < (int)L.regs.size() ? (uint64_t)L.regs[j] : 0ull);
        put64(dops.size());
        put(dops.data(), dops.size() * sizeof(DOp));
    }
    put64(arena.size());
    put(arena.data(), arena.size() * sizeof(amp_t));
    return QI_OK;
}

void jit_drain() { jit::drain(); }
void jit_stats(uint64_t* modules, uint64_t* failed, uint64_t* pending, double* assemble_ms, double* fp64_warp_instr, int reset) {
    jit::Cache& c = jit::cache();
    std::lock_guard<std::mutex> lk(c.mu);
    *modules = c.assembled; *failed = c.failed; *pending = (uint64_t)c.pending; *assemble_ms = c.assemble_ms;
    *fp64_warp_instr = c.fp64_warp_instr;
    if (reset) c.fp64_warp_instr = 0.0;
}

}  // namespace qi
