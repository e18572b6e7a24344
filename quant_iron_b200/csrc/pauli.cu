// pauli.cu -- PauliString apply / exp and SumOp apply / expectation (pauli_string.rs), Trotter
// loops (time_evolution.rs).
//
// A Pauli string is reduced on the host to three masks over PHYSICAL bit positions:
//   xmask  bits flipped (X and Y factors),  zmask  bits that contribute a sign (Y and Z factors),
//   ny     number of Y factors.
// Applying the single Paulis one after the other (pauli_string.rs:172-184) gives exactly
//   (P psi)[i] = i^(3*ny + 2*popc(i & zmask)) * psi[i ^ xmask]
// (each factor is a permutation times a power of i, all exact in floating point), so one fused
// pass replaces the reference's clone + one sweep per factor; exp(alpha P) psi =
// cosh(alpha) psi + sinh(alpha) P psi (pauli_string.rs:251-261) is one pass too.
#include <algorithm>

#include "common.cuh"

namespace qi {

static const int kBlock = 256;

struct Masks { uint64_t x, z; int ny; uint64_t high; /* rank bits of this shard, pre-shifted */ };

__device__ __forceinline__ int pauli_pow(uint64_t full_index, const Masks& m) {
    return (3 * m.ny + 2 * __popcll(full_index & m.z)) & 3;
}

// ---- apply ------------------------------------------------------------------------------------
// diagonal string (xmask == 0): psi[i] <- coeff * i^k(i) * psi[i]
__global__ void __launch_bounds__(256) k_pauli_diag(amp_t* __restrict__ a, uint64_t len, Masks m, amp_t coeff, int with_coeff) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        amp_t v = mul_i_pow(a[i], pauli_pow(i | m.high, m));
        a[i] = with_coeff ? cmul(v, coeff) : v;
    }
}

// general string: pairs (i, j = i ^ xmask), i with the pivot bit (highest x bit) clear
__global__ void __launch_bounds__(256) k_pauli_pair(amp_t* __restrict__ a, uint64_t pairs, int pivot, Masks m, amp_t coeff, int with_coeff) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pairs) return;
    uint64_t i = insert_zero(k, pivot), j = i ^ m.x;
    amp_t ai = a[i], aj = a[j];
    amp_t ni = mul_i_pow(aj, pauli_pow(i | m.high, m));
    amp_t nj = mul_i_pow(ai, pauli_pow(j | m.high, m));
    if (with_coeff) { ni = cmul(ni, coeff); nj = cmul(nj, coeff); }
    a[i] = ni;
    a[j] = nj;
}

// ---- exp --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pauli_exp_diag(amp_t* __restrict__ a, uint64_t len, Masks m, amp_t ch, amp_t sh) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        amp_t v = a[i];
        a[i] = cadd(cmul(v, ch), cmul(mul_i_pow(v, pauli_pow(i | m.high, m)), sh));
    }
}

__global__ void __launch_bounds__(256) k_pauli_exp_pair(amp_t* __restrict__ a, uint64_t pairs, int pivot, Masks m, amp_t ch, amp_t sh) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pairs) return;
    uint64_t i = insert_zero(k, pivot), j = i ^ m.x;
    amp_t ai = a[i], aj = a[j];
    amp_t pi = mul_i_pow(aj, pauli_pow(i | m.high, m));   // (P psi)[i]
    amp_t pj = mul_i_pow(ai, pauli_pow(j | m.high, m));   // (P psi)[j]
    a[i] = cadd(cmul(ai, ch), cmul(pi, sh));               // state*cosh + P state*sinh, pauli_string.rs:255-261
    a[j] = cadd(cmul(aj, ch), cmul(pj, sh));
}

// ---- expectation -------------------------------------------------------------------------------
__device__ __forceinline__ double2 block_sum2(double2 v) {
    __shared__ double2 sh[32];
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, o);
        v.y += __shfl_down_sync(0xffffffffu, v.y, o);
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : make_double2(0.0, 0.0);
    if (w == 0)
        for (int o = 16; o > 0; o >>= 1) {
            v.x += __shfl_down_sync(0xffffffffu, v.x, o);
            v.y += __shfl_down_sync(0xffffffffu, v.y, o);
        }
    return v;
}

// sum_i conj(psi[i]) * (coeff * i^k(i) * psi[i ^ x])   (pauli_string.rs:491-502 for one term)
__global__ void __launch_bounds__(256) k_pauli_expect(const amp_t* __restrict__ a, uint64_t len, Masks m, amp_t coeff, double2* partials) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double2 acc = make_double2(0.0, 0.0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        amp_t phi = cmul(mul_i_pow(a[i ^ m.x], pauli_pow(i | m.high, m)), coeff);
        amp_t t = cmul(cconj(a[i]), phi);
        acc.x += t.x;
        acc.y += t.y;
    }
    double2 r = block_sum2(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// per-term totals (fixed tree), then the terms are added in order by thread 0 (pauli_string.rs:506)
__global__ void k_expect_final(const double2* partials, int per_term, int terms, double2* out) {
    extern __shared__ double2 term_sums[];
    for (int t = 0; t < terms; t++) {
        const double2* p = partials + (size_t)t * per_term;
        double2 acc = make_double2(0.0, 0.0);
        for (int i = threadIdx.x; i < per_term; i += blockDim.x) { acc.x += p[i].x; acc.y += p[i].y; }
        double2 r = block_sum2(acc);
        if (threadIdx.x == 0) term_sums[t] = r;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double2 tot = make_double2(0.0, 0.0);
        for (int t = 0; t < terms; t++) { tot.x += term_sums[t].x; tot.y += term_sums[t].y; }
        out[0] = tot;
    }
}

// out += psi-derived term, used by SumOp::apply: out[i] (+)= coeff * i^k(i) * psi[i ^ x]
__global__ void __launch_bounds__(256) k_pauli_accumulate(amp_t* __restrict__ out, const amp_t* __restrict__ a, uint64_t len, Masks m, amp_t coeff, int first) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        amp_t phi = cmul(mul_i_pow(a[i ^ m.x], pauli_pow(i | m.high, m)), coeff);
        out[i] = first ? phi : cadd(out[i], phi);
    }
}

// ---- host side -----------------------------------------------------------------------------------
static int term_masks(const qi_state* s, const qi_pauli_term* t, Masks* m) {
    if (!t) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "term is NULL");
    if (t->num_ops && (!t->qubits || !t->paulis)) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "term arrays are NULL");
    m->x = m->z = 0; m->ny = 0; m->high = 0;
    uint64_t seen = 0;
    for (uint32_t i = 0; i < t->num_ops; i++) {
        uint32_t q = t->qubits[i];
        // each single-Pauli apply validates its target (operator.rs:481 -> 228-231)
        if (q >= s->num_qubits) return fail(QI_ERR_INVALID_QUBIT_INDEX, q, s->num_qubits, "Invalid qubit index");
        if ((seen >> q) & 1) return fail(QI_ERR_INVALID_ARGUMENT, q, 0, "Duplicate Pauli string operator for qubit");
        seen |= 1ull << q;
        uint64_t bit = 1ull << s->phys[q];
        switch (t->paulis[i]) {
            case 1: m->x |= bit; break;
            case 2: m->x |= bit; m->z |= bit; m->ny++; break;
            case 3: m->z |= bit; break;
            default: return fail(QI_ERR_INVALID_ARGUMENT, t->paulis[i], 0, "pauli must be 1 (X), 2 (Y) or 3 (Z)");
        }
    }
    if (!s->consistent) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, s->num_qubits, 0, "state vector length is not 2^num_qubits");
    m->high = (uint64_t)s->rank << s->n_local;
    return QI_OK;
}

static int highest_bit(uint64_t m) { int p = -1; while (m) { p++; m >>= 1; } return p; }

int pauli_global_exchange(qi_state* s, Masks* m);   // shard.cu hook (no-op for world == 1)

static int apply_masks(qi_state* s, Masks m, amp_t coeff, int with_coeff) {
    Context& c = ctx();
    const uint64_t local_mask = s->len - 1;
    if ((m.x & ~local_mask) != 0) return fail(QI_ERR_PEER, 0, 0, "X/Y factor on a global qubit needs an exchange first");
    LaunchScope ls(KF_PAULI, 32.0 * (double)s->len);
    if ((m.x & local_mask) == 0) {
        k_pauli_diag<<<grid_for(s->len, kBlock), kBlock, 0, c.stream>>>(s->d, s->len, m, coeff, with_coeff);
    } else {
        uint64_t pairs = s->len >> 1;
        k_pauli_pair<<<(unsigned)((pairs + kBlock - 1) / kBlock), kBlock, 0, c.stream>>>(s->d, pairs, highest_bit(m.x), m, coeff, with_coeff);
    }
    return check_launch("pauli_apply");
}

static int exp_masks(qi_state* s, Masks m, amp_t alpha) {
    Context& c = ctx();
    const uint64_t local_mask = s->len - 1;
    if ((m.x & ~local_mask) != 0) return fail(QI_ERR_PEER, 0, 0, "X/Y factor on a global qubit needs an exchange first");
    amp_t ch = h_ccosh(alpha), sh = h_csinh(alpha);
    LaunchScope ls(KF_PAULI_EXP, 32.0 * (double)s->len);
    if ((m.x & local_mask) == 0) {
        k_pauli_exp_diag<<<grid_for(s->len, kBlock), kBlock, 0, c.stream>>>(s->d, s->len, m, ch, sh);
    } else {
        uint64_t pairs = s->len >> 1;
        k_pauli_exp_pair<<<(unsigned)((pairs + kBlock - 1) / kBlock), kBlock, 0, c.stream>>>(s->d, pairs, highest_bit(m.x), m, ch, sh);
    }
    return check_launch("pauli_exp");
}

int shard_localise_mask(qi_state* s, const qi_pauli_term* t);   // shard.cu: make every X/Y qubit local

// one term on the per-term kernels (k0 already holds the rank-bit signs; 3*(3*k0) = k0 mod 4)
int pauli_exp_single(qi_state* s, const PauliExp& t) {
    Context& c = ctx();
    Masks m{t.x, t.z, (3 * t.k0) & 3, 0};
    LaunchScope ls(KF_PAULI_EXP, 32.0 * (double)s->len);
    if (t.x == 0) {
        k_pauli_exp_diag<<<grid_for(s->len, kBlock), kBlock, 0, c.stream>>>(s->d, s->len, m, t.ch, t.sh);
    } else {
        uint64_t pairs = s->len >> 1;
        k_pauli_exp_pair<<<(unsigned)((pairs + kBlock - 1) / kBlock), kBlock, 0, c.stream>>>(s->d, pairs, highest_bit(t.x), m, t.ch, t.sh);
    }
    return check_launch("pauli_exp");
}

// reduce exp(factor * coefficient * P) to masks under the state's CURRENT layout
static int make_exp(const qi_state* s, const qi_pauli_term& t, amp_t factor, PauliExp* e, bool* needs_exchange) {
    Masks m;
    QI_TRY(term_masks(s, &t, &m));
    const amp_t alpha = cmul(make_double2(t.coefficient[0], t.coefficient[1]), factor);      // pauli_string.rs:239
    *needs_exchange = false;
    if (t.num_ops == 0) {            // pauli_string.rs:241-244: state * exp(alpha)
        e->x = e->z = 0; e->k0 = 0; e->ch = h_cexp(alpha); e->sh = make_double2(0.0, 0.0);
        return QI_OK;
    }
    const uint64_t local_mask = s->len - 1;
    *needs_exchange = (m.x & ~local_mask) != 0;
    e->x = m.x;
    e->z = m.z & local_mask;
    e->k0 = (3 * m.ny + 2 * __builtin_popcountll(m.high & m.z)) & 3;
    e->ch = h_ccosh(alpha);
    e->sh = h_csinh(alpha);
    return QI_OK;
}

// state <- exp(f_k c_k P_k) state for the terms order[0], order[1], ... (apply_exp_factor in sequence).
// Consecutive terms are fused into register-window passes (pauli_window.cu) unless the state is too small
// or fusion is switched off; a term with an X/Y factor on a rank bit first flushes the pending batch and
// exchanges that qubit into the local bits.
static int exp_sequence(qi_state* s, const qi_pauli_term* terms, const std::vector<uint32_t>& order, const std::vector<amp_t>& factors) {
    Context& c = ctx();
    const bool batch = c.opt_fuse && c.opt_path != 1 && pauli_window_supported(s);
    // apply the terms order[take[0]], order[take[1]], ... under the current layout (all their X/Y factors local)
    auto run = [&](const std::vector<size_t>& take) -> int {
        std::vector<PauliExp> pending;
        for (size_t k : take) {
            const qi_pauli_term& t = terms[order[k]];
            PauliExp e;
            bool ex = false;
            QI_TRY(make_exp(s, t, factors[k], &e, &ex));
            if (ex) return fail(QI_ERR_PEER, 0, 0, "X/Y factor on a global qubit inside a stage");
            if (batch) { pending.push_back(e); continue; }
            if (t.num_ops == 0) { double z[2] = {e.ch.x, e.ch.y}; QI_TRY(qi_scale(s, z)); }
            else QI_TRY(pauli_exp_single(s, e));
        }
        return run_pauli_exp_batch(s, pending);
    };
    if (s->world == 1) {
        std::vector<size_t> all(order.size());
        for (size_t k = 0; k < all.size(); k++) all[k] = k;
        return run(all);
    }
    // sharded: stages around the exchanges (shard_pauli_walk), on logical-qubit masks
    std::vector<uint64_t> lx(order.size(), 0), lz(order.size(), 0);
    for (size_t k = 0; k < order.size(); k++) {
        const qi_pauli_term& t = terms[order[k]];
        for (uint32_t i = 0; i < t.num_ops; i++) {
            if (t.paulis[i] != 3) lx[k] |= 1ull << t.qubits[i];
            if (t.paulis[i] != 1) lz[k] |= 1ull << t.qubits[i];
        }
    }
    return shard_pauli_walk(s, lx, lz, run, false, nullptr);
}

}  // namespace qi

using namespace qi;

extern "C" {

int qi_apply_pauli_string(qi_state* s, const qi_pauli_term* term, int with_coefficient) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    Masks m;
    QI_TRY(term_masks(s, term, &m));
    QI_TRY(ensure_ctx());
    amp_t coeff = make_double2(term->coefficient[0], term->coefficient[1]);
    if (term->num_ops == 0) {
        // pauli_string.rs:141-143 / 173-177: empty string = coefficient (or identity)
        if (!with_coefficient) return QI_OK;
        return qi_scale(s, term->coefficient);
    }
    if (s->world > 1) { QI_TRY(shard_localise_mask(s, term)); QI_TRY(term_masks(s, term, &m)); }
    return apply_masks(s, m, coeff, with_coefficient);
}

int qi_apply_pauli_exp(qi_state* s, const qi_pauli_term* term, const double factor[2]) {
    if (!s || !factor) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    Masks m;
    QI_TRY(term_masks(s, term, &m));
    QI_TRY(ensure_ctx());
    amp_t alpha = cmul(make_double2(term->coefficient[0], term->coefficient[1]), make_double2(factor[0], factor[1]));  // pauli_string.rs:239
    if (term->num_ops == 0) {
        amp_t e = h_cexp(alpha);          // pauli_string.rs:241-244
        double z[2] = {e.x, e.y};
        return qi_scale(s, z);
    }
    if (s->world > 1) { QI_TRY(shard_localise_mask(s, term)); QI_TRY(term_masks(s, term, &m)); }
    return exp_masks(s, m, alpha);
}

int qi_expect_pauli_sum(const qi_state* s, const qi_pauli_term* terms, uint64_t count, double out[2]) {
    if (!s || !out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    out[0] = out[1] = 0.0;
    if (count == 0) return QI_OK;     // pauli_string.rs:486-489
    if (!terms) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "terms is NULL");
    Masks m;
    for (uint64_t k = 0; k < count; k++) QI_TRY(term_masks(s, &terms[k], &m));
    QI_TRY(ensure_ctx());
    Context& c = ctx();
    int g = c.sm_count * 4;
    uint64_t need = (s->len + kBlock - 1) / kBlock;
    if (need < (uint64_t)g) g = (int)need;
    // every launch (one window group of many terms, or one term on the per-term kernel) fills one SLOT of g block
    // partials; k_expect_final adds the slots in order
    const uint64_t kSlots = 512;
    QI_TRY(ensure_partials((size_t)g * 2 * kSlots));
    double2* partials = (double2*)c.d_partials;
    double tot[2] = {0.0, 0.0};
    uint64_t used = 0;
    auto flush_slots = [&]() -> int {
        if (!used) return QI_OK;
        k_expect_final<<<1, kBlock, used * sizeof(double2), c.stream>>>(partials, g, (int)used, (double2*)c.d_result);
        QI_TRY(check_launch("pauli_expect"));
        QI_CUDA(cudaMemcpyAsync(c.h_result, c.d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        QI_CUDA(cudaStreamSynchronize(c.stream));
        tot[0] += c.h_result[0];
        tot[1] += c.h_result[1];
        used = 0;
        return QI_OK;
    };
    auto single = [&](const PauliExp& e) -> int {
        if (used == kSlots) QI_TRY(flush_slots());
        Masks mk{e.x, e.z, (3 * e.k0) & 3, 0};
        LaunchScope ls(KF_EXPECT, (e.x ? 32.0 : 16.0) * (double)s->len);
        k_pauli_expect<<<g, kBlock, 0, c.stream>>>(s->d, s->len, mk, e.ch, partials + used * g);
        used++;
        return QI_OK;
    };
    // terms whose X/Y factors fit a register window share one read of the state (pauli_window.cu)
    const bool batch = c.opt_fuse && c.opt_path != 1 && pauli_window_supported(s);
    std::vector<PauliExp> pend;
    auto flush_batch = [&]() -> int {
        while (!pend.empty()) {
            if (used == kSlots) QI_TRY(flush_slots());
            std::vector<size_t> left;
            int groups = 0;
            QI_TRY(run_pauli_expect_batch(s, pend, g, partials + used * g, (int)(kSlots - used), &groups, &left));
            used += (uint64_t)groups;
            // terms that found no group: wider than a window -> per-term kernel; out of slots -> next round
            std::vector<PauliExp> retry;
            for (size_t idx : left) {
                if (__builtin_popcountll(pend[idx].x & ~31ull) > 4) QI_TRY(single(pend[idx]));
                else retry.push_back(pend[idx]);
            }
            pend.swap(retry);
            if (!pend.empty()) QI_TRY(flush_slots());
        }
        return QI_OK;
    };
    const uint64_t local_mask = s->len - 1;
    for (uint64_t k = 0; k < count; k++) {
        const qi_pauli_term& t = terms[k];
        QI_TRY(term_masks(s, &t, &m));
        if (m.x & ~local_mask) {
            // X/Y on a rank bit pairs amplitudes across ranks: bring that qubit into the local bits first.  The
            // exchange changes the layout, so what is pending (masks under the old layout) is launched before it.
            QI_TRY(flush_batch());
            QI_TRY(shard_localise_mask(const_cast<qi_state*>(s), &t));
            QI_TRY(term_masks(s, &t, &m));
        }
        PauliExp e;
        e.x = m.x;
        e.z = m.z & local_mask;
        e.k0 = (3 * m.ny + 2 * __builtin_popcountll(m.high & m.z)) & 3;
        e.ch = make_double2(t.coefficient[0], t.coefficient[1]);
        e.sh = make_double2(0.0, 0.0);
        if (batch) pend.push_back(e);
        else QI_TRY(single(e));
    }
    QI_TRY(flush_batch());
    QI_TRY(flush_slots());
    if (s->world > 1) QI_TRY(shard_allreduce_sum(const_cast<qi_state*>(s), tot, 2));
    out[0] = tot[0];
    out[1] = tot[1];
    return QI_OK;
}

int qi_apply_pauli_sum(const qi_state* s, const qi_pauli_term* terms, uint64_t count, qi_state** out) {
    if (!s || !out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    if (s->world > 1) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "SumOp::apply needs a second full state; not available on sharded states");
    if (count && !terms) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "terms is NULL");
    std::vector<Masks> ms(count);
    for (uint64_t k = 0; k < count; k++) QI_TRY(term_masks(s, &terms[k], &ms[k]));
    QI_TRY(qi_state_clone(s, out));
    if (count == 0) {                 // pauli_string.rs:454-457: state * 0.0
        double z[2] = {0.0, 0.0};
        return qi_scale(*out, z);
    }
    Context& c = ctx();
    for (uint64_t k = 0; k < count; k++) {
        LaunchScope ls(KF_PAULI, 48.0 * (double)s->len);
        k_pauli_accumulate<<<grid_for(s->len, kBlock), kBlock, 0, c.stream>>>((*out)->d, s->d, s->len, ms[k],
            make_double2(terms[k].coefficient[0], terms[k].coefficient[1]), k == 0);
    }
    return check_launch("pauli_sum_apply");
}

int qi_apply_pauli_exp_sequence(qi_state* s, const qi_pauli_term* terms, uint64_t count, const double* factors) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    if (count == 0) return QI_OK;
    if (!terms || !factors) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    Masks m;
    for (uint64_t k = 0; k < count; k++) QI_TRY(term_masks(s, &terms[k], &m));      // validate before touching the state
    QI_TRY(ensure_ctx());
    std::vector<uint32_t> order(count);
    std::vector<amp_t> f(count);
    for (uint64_t k = 0; k < count; k++) { order[k] = (uint32_t)k; f[k] = make_double2(factors[2 * k], factors[2 * k + 1]); }
    return exp_sequence(s, terms, order, f);
}

int qi_trotter_evolve(qi_state* s, const qi_pauli_term* terms, uint64_t count, double dt, uint64_t steps, int order) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    if (count == 0) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, 0, 0, "empty Hamiltonian");   // time_evolution.rs:147-149
    if (!terms) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "terms is NULL");
    if (order != 1 && order != 2) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)order, 0, "order must be 1 or 2");
    if (count > 0xffffffffull) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "too many terms");
    // validate every term once, before touching the state
    Masks m;
    for (uint64_t k = 0; k < count; k++) QI_TRY(term_masks(s, &terms[k], &m));
    QI_TRY(ensure_ctx());
    // the whole evolution is one sequence of apply_exp_factor calls; it is handed to the batching executor in
    // chunks of whole steps so that passes can span step boundaries
    const amp_t f1 = make_double2(0.0, -dt), f2 = make_double2(0.0, -dt / 2.0);
    const uint64_t per_step = order == 1 ? count : 2 * count;
    const uint64_t steps_per_chunk = std::max<uint64_t>(1, 16384 / per_step);
    std::vector<uint32_t> seq;
    std::vector<amp_t> f;
    for (uint64_t step0 = 0; step0 < steps; step0 += steps_per_chunk) {
        const uint64_t ns = std::min(steps_per_chunk, steps - step0);
        seq.clear();
        f.clear();
        for (uint64_t st = 0; st < ns; st++) {
            if (order == 1) {
                for (uint64_t k = 0; k < count; k++) { seq.push_back((uint32_t)k); f.push_back(f1); }           // 57-63
            } else {
                for (uint64_t k = 0; k < count; k++) { seq.push_back((uint32_t)k); f.push_back(f2); }           // 102-105
                for (uint64_t k = count; k-- > 0;) { seq.push_back((uint32_t)k); f.push_back(f2); }             // 108-111
            }
        }
        QI_TRY(exp_sequence(s, terms, seq, f));
    }
    return QI_OK;
}

int qi_debug_pauli_schedule(uint32_t num_qubits, const qi_pauli_term* terms, uint64_t count, uint64_t repeats,
                            int32_t* terms_per_pass, uint64_t max_passes, uint64_t* n_passes) {
    if (!terms || !n_passes || num_qubits == 0 || num_qubits > 62) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "bad argument");
    qi_state host;                   // layout only: no device memory is touched
    host.num_qubits = host.n_local = num_qubits;
    host.len = 1ull << num_qubits;
    for (int q = 0; q < 64; q++) host.phys[q] = (uint8_t)q;
    std::vector<PauliExp> one(count), seq;
    for (uint64_t k = 0; k < count; k++) {
        bool ex = false;
        QI_TRY(make_exp(&host, terms[k], make_double2(0.0, -0.01), &one[k], &ex));
    }
    for (uint64_t r = 0; r < repeats; r++) seq.insert(seq.end(), one.begin(), one.end());
    std::vector<int> per;
    QI_TRY(debug_pauli_schedule(seq, &per));
    *n_passes = per.size();
    for (size_t i = 0; i < per.size() && i < max_passes; i++) terms_per_pass[i] = per[i];
    return QI_OK;
}

// Host-only: the device programs qi_apply_pauli_exp_sequence would launch on one device (fused register-window passes and
// terms that run alone), serialised for the CPU interpreter in tests/ (layout: pauli_window.cu, debug_pauli_lower).
int qi_debug_pauli_lower(uint32_t num_qubits, int rank, int world, const uint8_t* phys, const qi_pauli_term* terms, uint64_t count,
                         const double* factors, uint8_t* blob, uint64_t capacity, uint64_t* used) {
    if (!terms || !factors || num_qubits == 0 || num_qubits > 62) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "bad argument");
    if (world < 1 || (world & (world - 1)) || world > 8 || rank < 0 || rank >= world) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)world, 0, "bad rank / world");
    int p = 0;
    while ((1 << p) < world) p++;
    qi_state host;                   // layout only: no device memory is touched
    host.num_qubits = num_qubits;
    host.n_local = num_qubits - (uint32_t)p;
    host.len = 1ull << host.n_local;
    host.rank = rank;
    host.world = world;
    for (int q = 0; q < 64; q++) host.phys[q] = phys ? phys[q] : (uint8_t)q;
    if (!pauli_window_supported(&host)) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, num_qubits, 0, "too few qubits for the window executor");
    std::vector<PauliExp> seq(count);
    for (uint64_t k = 0; k < count; k++) {
        bool ex = false;
        memset(&seq[k], 0, sizeof(PauliExp));
        QI_TRY(make_exp(&host, terms[k], make_double2(factors[2 * k], factors[2 * k + 1]), &seq[k], &ex));
        if (ex) return fail(QI_ERR_PEER, k, 0, "X/Y factor on a global qubit inside a stage");
    }
    std::vector<uint8_t> out;
    QI_TRY(debug_pauli_lower(&host, seq, &out));
    if (used) *used = out.size();
    if (out.size() > capacity || !blob) return fail(QI_ERR_INVALID_ARGUMENT, out.size(), capacity, "blob too small");
    memcpy(blob, out.data(), out.size());
    return QI_OK;
}

// Host-only: the stages qi_apply_pauli_exp_sequence runs in on `world` ranks (terms per stage, the qubit map each stage runs
// under, the exchange that follows) -- record layout: shard.cu, debug_shard_pauli_stages.
int qi_debug_shard_pauli_stages(uint32_t total_qubits, int world, const qi_pauli_term* terms, uint64_t count, uint64_t* out,
                                uint64_t capacity, uint64_t* used) {
    if (world != 1 && world != 2 && world != 4 && world != 8) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)world, 0, "world must be 1, 2, 4 or 8");
    if (count && !terms) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "terms is NULL");
    std::vector<uint64_t> lx(count, 0), lz(count, 0);
    for (uint64_t k = 0; k < count; k++)
        for (uint32_t i = 0; i < terms[k].num_ops; i++) {
            const uint32_t q = terms[k].qubits[i];
            if (q >= total_qubits) return fail(QI_ERR_INVALID_QUBIT_INDEX, q, total_qubits, "Invalid qubit index");
            if (terms[k].paulis[i] != 3) lx[k] |= 1ull << q;
            if (terms[k].paulis[i] != 1) lz[k] |= 1ull << q;
        }
    std::vector<uint64_t> rec;
    QI_TRY(debug_shard_pauli_stages(total_qubits, world, lx, lz, &rec));
    if (used) *used = rec.size();
    if (rec.size() > capacity || !out) return fail(QI_ERR_INVALID_ARGUMENT, rec.size(), capacity, "buffer too small");
    memcpy(out, rec.data(), rec.size() * sizeof(uint64_t));
    return QI_OK;
}

// Host-only: the read-only window programs qi_expect_pauli_sum would launch on one device (first-fit groups of terms that
// share a register window) and the terms left to the per-term kernel (layout: pauli_window.cu, debug_expect_lower).
int qi_debug_expect_lower(uint32_t num_qubits, const qi_pauli_term* terms, uint64_t count, uint8_t* blob, uint64_t capacity,
                          uint64_t* used) {
    if (!terms || num_qubits == 0 || num_qubits > 62) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "bad argument");
    qi_state host;                   // layout only: no device memory is touched
    host.num_qubits = host.n_local = num_qubits;
    host.len = 1ull << num_qubits;
    for (int q = 0; q < 64; q++) host.phys[q] = (uint8_t)q;
    if (!pauli_window_supported(&host)) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, num_qubits, 0, "too few qubits for the window executor");
    std::vector<PauliExp> seq(count);
    Masks m;
    for (uint64_t k = 0; k < count; k++) {
        QI_TRY(term_masks(&host, &terms[k], &m));
        memset(&seq[k], 0, sizeof(PauliExp));
        seq[k].x = m.x;
        seq[k].z = m.z;
        seq[k].k0 = (3 * m.ny) & 3;
        seq[k].ch = make_double2(terms[k].coefficient[0], terms[k].coefficient[1]);
    }
    std::vector<uint8_t> out;
    QI_TRY(debug_expect_lower(&host, seq, &out));
    if (used) *used = out.size();
    if (out.size() > capacity || !blob) return fail(QI_ERR_INVALID_ARGUMENT, out.size(), capacity, "blob too small");
    memcpy(blob, out.data(), out.size());
    return QI_OK;
}

}  // extern "C"
