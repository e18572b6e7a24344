// pauli_window.cu -- a SEQUENCE of PauliString::apply_exp_factor calls fused into register-window
// passes (Trotter batching).
//
// The reference's Trotter drivers (time_evolution.rs:45-66, 89-115) call apply_exp_factor once per
// Hamiltonian term and step, each ~8 sweeps over the state (pauli_string.rs:237-262).  pauli.cu
// already makes each call ONE pass at HBM roofline; here consecutive calls share a pass: the warp
// tile of window.cu (lanes <-> physical qubits 0..4, register slots <-> R window qubits, tile index
// <-> the rest) holds 512 amplitudes in registers and every term whose X/Y factors sit on lane or
// window qubits is applied to them in place,
//     psi[i] <- cosh(a) psi[i] + sinh(a) i^(3 nY + 2 popc(i & z)) psi[i ^ x],
// with the same operation order as k_pauli_exp_pair (results track the per-term path to the last
// bit up to FMA contraction).  Z factors may sit on ANY qubit (they only enter the power of i).
//
// Host scheduler: greedy over the sequence; a term joins the current pass if its X/Y qubits fit the
// window and it commutes with every term deferred so far (two Pauli strings commute iff they
// anticommute on an even number of qubits), so later Trotter steps start while the far end of the
// chain is still waiting for its window (a wavefront over the chain).
#include "common.cuh"
#include "window_layout.cuh"

namespace qi {

static const int kMaxPX = 192;       // ops per launch (parameter space: 192 * 48 B + header < 16 KB)
static const int kPauliR = 4;

struct PXOp {                // 48 bytes
    uint32_t xl, xr;         // flipped bits in lane space / slot space
    uint32_t zl, zr;         // sign bits in lane space / slot space
    uint64_t zt;             // sign bits in compact tile-index space
    uint32_t k0;             // power of i common to the whole state (see PauliExp::k0), + 1 when sinh is imaginary
    uint32_t pad;
    double c, s;             // psi <- c psi + s i^k P-permuted psi, c and s real
};

template <int R>
struct PXProgram {
    BitInsert ins;
    uint64_t off[1 << R];
    uint32_t nops;
    uint32_t unit;           // the pass runs in unit form (every |c| >= 0.05): ops hold (1, s / c), `scale` = product of the c
    double scale;
    PXOp ops[kMaxPX];
};

__device__ __forceinline__ amp_t px_shfl(amp_t v, int mask) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}

// mine <- c*mine + s * i^k * other, written for real c and s: i^k is a component swap (SW = k odd, uniform per
// op) and two signs, which are folded into the coefficients: s_re multiplies the source of the new real part,
// s_im the source of the new imaginary part; `par` (slot parity of the Z mask) flips both.
// x with its sign bit xor-ed by `m` (0 or 0x80000000): one LOP3 on the high word, bit-identical to a conditional negation
__device__ __forceinline__ double px_sign(double x, uint32_t m) { return __hiloint2double(__double2hiint(x) ^ (int)m, __double2loint(x)); }
// parity of (s & zr) for every slot s at once: bit s of the result (xor of the bit patterns of the set bits of zr); computed once
// per op -- the per-slot popc / compare / select chains were 21 % FSEL + 5 % POPC of the executed instructions (ncu, round 2)
__device__ __forceinline__ uint32_t px_parity_mask(uint32_t zr) {
    return ((zr & 1u) ? 0xAAAAAAAAu : 0u) ^ ((zr & 2u) ? 0xCCCCCCCCu : 0u) ^ ((zr & 4u) ? 0xF0F0F0F0u : 0u) ^ ((zr & 8u) ? 0xFF00FF00u : 0u) ^
           ((zr & 16u) ? 0xFFFF0000u : 0u);
}
// UNIT: the pass runs in unit form -- every op is (1 / c) of itself (mine + (s / c) i^k other: one FMA per component instead of
// a product and an FMA) and the product of the c is applied once, at the end of the pass (PXProgram::scale)
template <bool SW, bool UNIT>
__device__ __forceinline__ amp_t px_mix(amp_t mine, amp_t other, double c, double s_re, double s_im, uint32_t m) {
    const double kr = px_sign(s_re, m), ki = px_sign(s_im, m);
    if (UNIT) {
        const double src_re = SW ? other.y : other.x, src_im = SW ? other.x : other.y;
        return make_double2(fma(kr, src_re, mine.x), fma(ki, src_im, mine.y));
    }
    const double src_re = SW ? other.y : other.x, src_im = SW ? other.x : other.y;
    return make_double2(c * mine.x + kr * src_re, c * mine.y + ki * src_im);
}

// one term on the register tile; M = slot xor mask (compile time so every v[] index is a literal)
template <int R, int M, bool SW, bool UNIT>
__device__ __forceinline__ void px_apply(amp_t (&v)[1 << R], const uint32_t xl, const uint32_t zr, const double c, const double s_re,
                                         const double s_im) {
    constexpr int S = 1 << R;
    const uint32_t pm = px_parity_mask(zr);
#define QI_PX_SIGN(slot) ((pm << (31 - (slot))) & 0x80000000u)
    if (M == 0) {
        if (xl) {
#pragma unroll
            for (int s = 0; s < S; s++) v[s] = px_mix<SW, UNIT>(v[s], px_shfl(v[s], xl), c, s_re, s_im, QI_PX_SIGN(s));
        } else {
#pragma unroll
            for (int s = 0; s < S; s++) v[s] = px_mix<SW, UNIT>(v[s], v[s], c, s_re, s_im, QI_PX_SIGN(s));
        }
        return;
    }
    constexpr int HB = (M & 16) ? 16 : (M & 8) ? 8 : (M & 4) ? 4 : (M & 2) ? 2 : 1;
    // the lane-flip test is hoisted out of the slot loop: inside it, it cost one branch per slot pair plus the register moves
    // that set up the "no shuffle" defaults (SASS, round 2)
    if (xl) {
#pragma unroll
        for (int s0 = 0; s0 < S; s0++) {
            if (s0 & HB) continue;
            const int s1 = s0 ^ M;
            const amp_t a = v[s0], b = v[s1];
            const amp_t oa = px_shfl(a, xl), ob = px_shfl(b, xl);
            v[s0] = px_mix<SW, UNIT>(a, ob, c, s_re, s_im, QI_PX_SIGN(s0));     // (P psi)[s0] comes from slot s1
            v[s1] = px_mix<SW, UNIT>(b, oa, c, s_re, s_im, QI_PX_SIGN(s1));
        }
    } else {
#pragma unroll
        for (int s0 = 0; s0 < S; s0++) {
            if (s0 & HB) continue;
            const int s1 = s0 ^ M;
            const amp_t a = v[s0], b = v[s1];
            v[s0] = px_mix<SW, UNIT>(a, b, c, s_re, s_im, QI_PX_SIGN(s0));
            v[s1] = px_mix<SW, UNIT>(b, a, c, s_re, s_im, QI_PX_SIGN(s1));
        }
    }
#undef QI_PX_SIGN
}

template <int R, bool SW, bool UNIT>
__device__ __forceinline__ void px_dispatch(amp_t (&v)[1 << R], const uint32_t xr, const uint32_t xl, const uint32_t zr, const double c,
                                            const double s_re, const double s_im) {
    constexpr int S = 1 << R;
#define QI_PX_CASE(m) case m: px_apply<R, ((m) < S ? (m) : 0), SW, UNIT>(v, xl, zr, c, s_re, s_im); break;
    switch (xr) {
        QI_PX_CASE(0) QI_PX_CASE(1) QI_PX_CASE(2) QI_PX_CASE(3) QI_PX_CASE(4) QI_PX_CASE(5) QI_PX_CASE(6) QI_PX_CASE(7)
        QI_PX_CASE(8) QI_PX_CASE(9) QI_PX_CASE(10) QI_PX_CASE(11) QI_PX_CASE(12) QI_PX_CASE(13) QI_PX_CASE(14) QI_PX_CASE(15)
        default: break;
    }
#undef QI_PX_CASE
}

template <int R, bool UNIT>
__global__ void __launch_bounds__(128, (R <= 3 ? 8 : 4)) k_pauli_window(amp_t* __restrict__ a, uint64_t ntiles, const __grid_constant__ PXProgram<R> P) {
    constexpr int S = 1 << R;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t tile = warp; tile < ntiles; tile += nwarps) {
        const uint64_t base = expand_index((tile << 5) | (uint64_t)lane, P.ins);
        amp_t v[S];
#pragma unroll
        for (int s = 0; s < S; s++) v[s] = ld_amp(a + base + P.off[s]);
#pragma unroll 1
        for (uint32_t o = 0; o < P.nops; o++) {
            const PXOp& op = P.ops[o];
            // power of i for this thread's amplitudes: k0 + 2*(Z-mask parity of the tile and lane bits); the slot
            // bits add their parity inside px_mix
            const uint32_t k0 = op.k0;
            const bool flip = (__popcll(tile & op.zt) + __popc((uint32_t)lane & op.zl)) & 1;
            const bool nre = (((k0 & 3) == 1) || ((k0 & 3) == 2)) != flip;      // i^1 = (-y, x), i^2 = (-x, -y), i^3 = (y, -x)
            const bool nim = ((k0 & 3) >= 2) != flip;
            const double s_re = nre ? -op.s : op.s, s_im = nim ? -op.s : op.s;
            if (k0 & 1) px_dispatch<R, true, UNIT>(v, op.xr, op.xl, op.zr, op.c, s_re, s_im);
            else px_dispatch<R, false, UNIT>(v, op.xr, op.xl, op.zr, op.c, s_re, s_im);
        }
        if (UNIT) {
            const double g = P.scale;
#pragma unroll
            for (int s = 0; s < S; s++) { v[s].x *= g; v[s].y *= g; }
        }
#pragma unroll
        for (int s = 0; s < S; s++) st_amp(a + base + P.off[s], v[s]);
    }
}

bool pauli_window_supported(const qi_state* s) {
    return s->consistent && (int)s->n_local >= kLaneQubits + kPauliR;
}

static inline bool anticommute(const PauliExp& a, const PauliExp& b) {
    return (__builtin_popcountll(a.x & b.z) + __builtin_popcountll(a.z & b.x)) & 1;
}

// the window kernel takes terms whose cosh is real and whose sinh is real or imaginary (every Trotter / time-
// evolution term and every apply_exp with a real coefficient) and whose X/Y factors above the lane qubits fit
// the window; anything else runs alone on the per-term kernels
static inline bool fusable(const PauliExp& t) {
    const uint64_t lane_mask = (1ull << kLaneQubits) - 1;
    return t.ch.y == 0.0 && (t.sh.x == 0.0 || t.sh.y == 0.0) && __builtin_popcountll(t.x & ~lane_mask) <= kPauliR;
}

struct PxPass { std::vector<int> regs; std::vector<size_t> terms; };

// greedy pass construction (host only).  A step with an empty pass means "term `single` runs alone on the
// per-term kernel" (more X/Y factors above the lane qubits than the window holds).
struct PxStep { PxPass pass; size_t single; };

// `emit` is called as soon as a step is final, so the device starts on pass k while the host builds k+1.
template <typename Emit>
static int schedule_pauli(const std::vector<PauliExp>& seq, Emit emit) {
    const size_t n = seq.size();
    const uint64_t lane_mask = (1ull << kLaneQubits) - 1;
    const size_t kLookahead = 256, kMaxDeferred = 96;
    std::vector<char> done(n, 0);
    size_t first = 0;
    while (first < n) {
        if (done[first]) { first++; continue; }
        if (!fusable(seq[first])) {
            QI_TRY(emit(PxStep{PxPass(), first}));
            done[first++] = 1;
            continue;
        }
        PxPass ps;
        uint64_t window_mask = 0, deferred_support = 0;
        std::vector<size_t> deferred;
        size_t scanned = 0;
        for (size_t i = first; i < n && scanned < kLookahead && ps.terms.size() < (size_t)kMaxPX; i++) {
            if (done[i]) continue;
            scanned++;
            const PauliExp& t = seq[i];
            bool ok = fusable(t);
            if (ok && ((t.x | t.z) & deferred_support))
                for (size_t d : deferred)
                    if (anticommute(seq[d], t)) { ok = false; break; }
            if (ok) {
                uint64_t need = t.x & ~lane_mask & ~window_mask;
                if ((int)ps.regs.size() + __builtin_popcountll(need) > kPauliR) ok = false;
                else
                    for (int q = kLaneQubits; q < 64 && need; q++)
                        if ((need >> q) & 1) { ps.regs.push_back(q); window_mask |= 1ull << q; need &= ~(1ull << q); }
            }
            if (ok) {
                ps.terms.push_back(i);
                done[i] = 1;
            } else {
                deferred.push_back(i);
                deferred_support |= t.x | t.z;
                if (deferred.size() >= kMaxDeferred) break;
            }
        }
        QI_TRY(emit(PxStep{std::move(ps), 0}));
    }
    return QI_OK;
}

// device program of one pass (host only)
static int build_px_program(const qi_state* s, const std::vector<PauliExp>& seq, const PxPass& ps, PXProgram<kPauliR>* Pout, Layout* Lout, bool allow_unit = false) {
    constexpr int R = kPauliR;
    Layout L = make_layout(s, ps.regs, R);
    PXProgram<R>& P = *Pout;
    memset(&P, 0, sizeof(P));
    fill_offsets<R>(L, &P.ins, P.off);
    P.nops = (uint32_t)ps.terms.size();
    for (size_t k = 0; k < ps.terms.size(); k++) {
        const PauliExp& t = seq[ps.terms[k]];
        PXOp& d = P.ops[k];
        uint64_t xt = 0;
        split_mask(L, t.x, &d.xl, &d.xr, &xt);
        if (xt) return fail(QI_ERR_UNKNOWN, 0, 0, "pauli window: X/Y factor outside the window");
        split_mask(L, t.z, &d.zl, &d.zr, &d.zt);
        // sinh = i*s: one more power of i (exact); sinh = s: as is
        const bool imag = t.sh.x == 0.0 && t.sh.y != 0.0;
        d.k0 = (uint32_t)((t.k0 + (imag ? 1 : 0)) & 3);
        d.c = t.ch.x;
        d.s = imag ? t.sh.y : t.sh.x;
    }
    P.scale = 1.0;
    if (allow_unit && P.nops >= 2) {
        bool ok = true;
        for (uint32_t k = 0; k < P.nops; k++) ok &= std::fabs(P.ops[k].c) >= 0.05;
        if (ok) {
            P.unit = 1;
            for (uint32_t k = 0; k < P.nops; k++) { P.scale *= P.ops[k].c; P.ops[k].s /= P.ops[k].c; P.ops[k].c = 1.0; }
        }
    }
    if (Lout) *Lout = L;
    return QI_OK;
}

static int launch_pauli_pass(qi_state* s, const std::vector<PauliExp>& seq, const PxPass& ps) {
    Context& c = ctx();
    constexpr int R = kPauliR;
    PXProgram<R> P;
    QI_TRY(build_px_program(s, seq, ps, &P, nullptr, c.opt_pauli_unit != 0));
    const uint64_t ntiles = s->len >> (kLaneQubits + R);
    const int warps_per_block = 4;
    uint64_t blocks = (ntiles + warps_per_block - 1) / warps_per_block;
    const uint64_t cap = (uint64_t)c.sm_count * 5 * 8;
    if (blocks > cap) blocks = cap;
    LaunchScope ls(KF_PAULI_WINDOW, 32.0 * (double)s->len);
    if (P.unit) k_pauli_window<R, true><<<(unsigned)blocks, warps_per_block * 32, 0, c.stream>>>(s->d, ntiles, P);
    else k_pauli_window<R, false><<<(unsigned)blocks, warps_per_block * 32, 0, c.stream>>>(s->d, ntiles, P);
    return check_launch("k_pauli_window");
}

// apply seq[0], seq[1], ... in order (all X/Y factors must already be on local qubits)
int run_pauli_exp_batch(qi_state* s, const std::vector<PauliExp>& seq) {
    if (seq.empty()) return QI_OK;
    return schedule_pauli(seq, [&](const PxStep& st) -> int {
        if (st.pass.terms.empty()) return pauli_exp_single(s, seq[st.single]);
        return launch_pauli_pass(s, seq, st.pass);
    });
}

// host-only view of the schedule: terms per pass (0 = a term that ran alone)
int debug_pauli_schedule(const std::vector<PauliExp>& seq, std::vector<int>* terms_per_pass) {
    return schedule_pauli(seq, [&](const PxStep& st) -> int {
        terms_per_pass->push_back((int)st.pass.terms.size());
        return QI_OK;
    });
}

// host-only: schedule AND lower a sequence; the device programs are serialised for the CPU interpreter in tests/
// (tests/window_interp.py).  Layout: u64 nsteps, then per step: u64 single;
//   single = 1: PauliExp (raw, 64 B);   single = 0: u64 R, u64 regs[8] (sorted window qubits), u64 nops, PXOp[nops] (raw, 48 B each)
static_assert(sizeof(PXOp) == 48 && sizeof(PauliExp) == 64, "layouts parsed by the CPU interpreter in tests/");
int debug_pauli_lower(const qi_state* s, const std::vector<PauliExp>& seq, std::vector<uint8_t>* blob) {
    auto put = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; blob->insert(blob->end(), b, b + n); };
    auto put64 = [&](uint64_t v) { put(&v, 8); };
    const size_t head = blob->size();
    put64(0);
    uint64_t nsteps = 0;
    QI_TRY(schedule_pauli(seq, [&](const PxStep& st) -> int {
        nsteps++;
        if (st.pass.terms.empty()) { put64(1); put(&seq[st.single], sizeof(PauliExp)); return QI_OK; }
        PXProgram<kPauliR> P;
        Layout L;
        QI_TRY(build_px_program(s, seq, st.pass, &P, &L));
        put64(0);
        put64((uint64_t)kPauliR);
        for (int j = 0; j < 8; j++) put64(j < (int)L.regs.size() ? (uint64_t)L.regs[j] : 0ull);
        put64(P.nops);
        put(P.ops, P.nops * sizeof(PXOp));
        return QI_OK;
    }));
    memcpy(blob->data() + head, &nsteps, 8);
    return QI_OK;
}

// ---- SumOp::expectation_value: many terms per read-only pass ------------------------------------------------
// <psi| c P |psi> = sum_i conj(psi[i]) c i^k(i) psi[i ^ x] (pauli_string.rs:491-502 for one term).  Terms whose X/Y
// factors fit a window share one read of the state; no ordering constraint (nothing is written), so terms are
// grouped first-fit.  Every thread adds its terms' contributions into one complex accumulator; the block sums
// go to `partials` and are added by k_expect_final in a fixed order (deterministic; the association differs from
// the reference's term-by-term sum, well inside the 1e-10 relative bar).
template <int R, int M, bool SW>
__device__ __forceinline__ void px_expect(const amp_t (&v)[1 << R], const uint32_t xl, const uint32_t zr, const bool nre, const bool nim,
                                          double& tr, double& ti) {
    constexpr int S = 1 << R;
    const uint32_t pm = px_parity_mask(zr);
    const uint32_t mre = nre ? 0x80000000u : 0u, mim = nim ? 0x80000000u : 0u;
    auto term = [&](int s, amp_t o) {
        const uint32_t par = (pm << (31 - s)) & 0x80000000u;
        double pre = SW ? o.y : o.x, pim = SW ? o.x : o.y;       // i^k o: component swap for odd k, then signs
        pre = px_sign(pre, mre ^ par);
        pim = px_sign(pim, mim ^ par);
        tr += v[s].x * pre + v[s].y * pim;                       // conj(v) * (pre + i pim)
        ti += v[s].x * pim - v[s].y * pre;
    };
    if (xl) {
#pragma unroll
        for (int s = 0; s < S; s++) term(s, px_shfl(v[s ^ M], xl));
    } else {
#pragma unroll
        for (int s = 0; s < S; s++) term(s, v[s ^ M]);
    }
}

template <int R, bool SW>
__device__ __forceinline__ void px_expect_dispatch(const amp_t (&v)[1 << R], const uint32_t xr, const uint32_t xl, const uint32_t zr,
                                                   const bool nre, const bool nim, double& tr, double& ti) {
    constexpr int S = 1 << R;
#define QI_PX_CASE(m) case m: px_expect<R, ((m) < S ? (m) : 0), SW>(v, xl, zr, nre, nim, tr, ti); break;
    switch (xr) {
        QI_PX_CASE(0) QI_PX_CASE(1) QI_PX_CASE(2) QI_PX_CASE(3) QI_PX_CASE(4) QI_PX_CASE(5) QI_PX_CASE(6) QI_PX_CASE(7)
        QI_PX_CASE(8) QI_PX_CASE(9) QI_PX_CASE(10) QI_PX_CASE(11) QI_PX_CASE(12) QI_PX_CASE(13) QI_PX_CASE(14) QI_PX_CASE(15)
        default: break;
    }
#undef QI_PX_CASE
}

template <int R>
__global__ void __launch_bounds__(128, 4) k_pauli_expect_window(const amp_t* __restrict__ a, uint64_t ntiles, const __grid_constant__ PXProgram<R> P,
                                                                 double2* __restrict__ partials) {
    constexpr int S = 1 << R;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    double ax = 0.0, ay = 0.0;
    for (uint64_t tile = warp; tile < ntiles; tile += nwarps) {
        const uint64_t base = expand_index((tile << 5) | (uint64_t)lane, P.ins);
        amp_t v[S];
#pragma unroll
        for (int s = 0; s < S; s++) v[s] = ld_amp(a + base + P.off[s]);
#pragma unroll 1
        for (uint32_t o = 0; o < P.nops; o++) {
            const PXOp& op = P.ops[o];
            const uint32_t k0 = op.k0;
            const bool flip = (__popcll(tile & op.zt) + __popc((uint32_t)lane & op.zl)) & 1;
            const bool nre = (((k0 & 3) == 1) || ((k0 & 3) == 2)) != flip;
            const bool nim = ((k0 & 3) >= 2) != flip;
            double tr = 0.0, ti = 0.0;
            if (k0 & 1) px_expect_dispatch<R, true>(v, op.xr, op.xl, op.zr, nre, nim, tr, ti);
            else px_expect_dispatch<R, false>(v, op.xr, op.xl, op.zr, nre, nim, tr, ti);
            ax += op.c * tr - op.s * ti;                          // coefficient (c, s) = (re, im)
            ay += op.c * ti + op.s * tr;
        }
    }
    // block sum: shuffle tree, then one value per warp through shared memory (fixed order)
    __shared__ double2 wsum[4];
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_down_sync(0xffffffffu, ax, o);
        ay += __shfl_down_sync(0xffffffffu, ay, o);
    }
    if (lane == 0) wsum[threadIdx.x >> 5] = make_double2(ax, ay);
    __syncthreads();
    if (threadIdx.x == 0) {
        double2 t = wsum[0];
        for (int w = 1; w < 4; w++) { t.x += wsum[w].x; t.y += wsum[w].y; }
        partials[blockIdx.x] = t;
    }
}

// terms: x/z/k0 as in PauliExp, ch = the term's coefficient.  Launches one kernel per window group with `grid`
// blocks, group j writing partials[j * grid ..]; returns the number of groups.  Terms that do not fit a window
// are returned in `leftover` (indices into `terms`) for the per-term kernel.
struct ExpectGroup { uint64_t window = 0; int nregs = 0; std::vector<size_t> terms; };

// first-fit grouping of the terms by the window their X/Y factors need (host only)
static void plan_expect_groups(const std::vector<PauliExp>& terms, int max_groups, std::vector<ExpectGroup>* groups_out,
                               std::vector<size_t>* leftover) {
    constexpr int R = kPauliR;
    const uint64_t lane_mask = (1ull << kLaneQubits) - 1;
    std::vector<ExpectGroup>& groups = *groups_out;
    for (size_t i = 0; i < terms.size(); i++) {
        const uint64_t need = terms[i].x & ~lane_mask;
        if (__builtin_popcountll(need) > R) { leftover->push_back(i); continue; }
        ExpectGroup* best = nullptr;
        for (ExpectGroup& g : groups) {
            if (g.terms.size() >= (size_t)kMaxPX) continue;
            if (g.nregs + __builtin_popcountll(need & ~g.window) <= R) { best = &g; break; }
        }
        if (!best) {
            if ((int)groups.size() >= max_groups) { leftover->push_back(i); continue; }
            groups.emplace_back();
            best = &groups.back();
        }
        best->nregs += __builtin_popcountll(need & ~best->window);
        best->window |= need;
        best->terms.push_back(i);
    }
}

// device program of one group (host only): (c, s) = the term's coefficient (re, im)
static int build_expect_program(const qi_state* s, const std::vector<PauliExp>& terms, const ExpectGroup& g, PXProgram<kPauliR>* Pout,
                                Layout* Lout) {
    constexpr int R = kPauliR;
    std::vector<int> regs;
    for (int q = kLaneQubits; q < 64; q++) if ((g.window >> q) & 1) regs.push_back(q);
    Layout L = make_layout(s, regs, R);
    PXProgram<R>& P = *Pout;
    memset(&P, 0, sizeof(P));
    fill_offsets<R>(L, &P.ins, P.off);
    P.nops = (uint32_t)g.terms.size();
    for (size_t k = 0; k < g.terms.size(); k++) {
        const PauliExp& t = terms[g.terms[k]];
        PXOp& d = P.ops[k];
        uint64_t xt = 0;
        split_mask(L, t.x, &d.xl, &d.xr, &xt);
        if (xt) return fail(QI_ERR_UNKNOWN, 0, 0, "pauli window: X/Y factor outside the window");
        split_mask(L, t.z, &d.zl, &d.zr, &d.zt);
        d.k0 = (uint32_t)(t.k0 & 3);
        d.c = t.ch.x;
        d.s = t.ch.y;
    }
    if (Lout) *Lout = L;
    return QI_OK;
}

int run_pauli_expect_batch(const qi_state* s, const std::vector<PauliExp>& terms, int grid, double2* partials, int max_groups,
                           int* groups_used, std::vector<size_t>* leftover) {
    Context& c = ctx();
    constexpr int R = kPauliR;
    std::vector<ExpectGroup> groups;
    plan_expect_groups(terms, max_groups, &groups, leftover);
    const uint64_t ntiles = s->len >> (kLaneQubits + R);
    for (size_t gi = 0; gi < groups.size(); gi++) {
        PXProgram<R> P;
        QI_TRY(build_expect_program(s, terms, groups[gi], &P, nullptr));
        LaunchScope ls(KF_EXPECT, 16.0 * (double)s->len);
        k_pauli_expect_window<R><<<grid, 128, 0, c.stream>>>(s->d, ntiles, P, partials + (size_t)gi * grid);
        QI_TRY(check_launch("k_pauli_expect_window"));
    }
    *groups_used = (int)groups.size();
    return QI_OK;
}

// host-only: the read-only window programs of a batched expectation value, serialised for the CPU interpreter in tests/:
// u64 ngroups, per group: u64 R, u64 regs[8], u64 nops, PXOp[nops]; then u64 nleft and the indices of the terms that
// found no group (they run on the per-term kernel)
int debug_expect_lower(const qi_state* s, const std::vector<PauliExp>& terms, std::vector<uint8_t>* blob) {
    auto put = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; blob->insert(blob->end(), b, b + n); };
    auto put64 = [&](uint64_t v) { put(&v, 8); };
    std::vector<ExpectGroup> groups;
    std::vector<size_t> left;
    plan_expect_groups(terms, 512, &groups, &left);
    put64(groups.size());
    for (const ExpectGroup& g : groups) {
        PXProgram<kPauliR> P;
        Layout L;
        QI_TRY(build_expect_program(s, terms, g, &P, &L));
        put64((uint64_t)kPauliR);
        for (int j = 0; j < 8; j++) put64(j < (int)L.regs.size() ? (uint64_t)L.regs[j] : 0ull);
        put64(P.nops);
        put(P.ops, P.nops * sizeof(PXOp));
    }
    put64(left.size());
    for (size_t i : left) put64(i);
    return QI_OK;
}

}  // namespace qi
