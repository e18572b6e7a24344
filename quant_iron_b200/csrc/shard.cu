// shard.cu -- state sharded over 2/4/8 GPUs of one NVSwitch domain (new work,
// the reference is single-process and its oracle is the single-device result).
//
// Layout: amplitude index = [rank bits (top log2 P) | local bits].  One process per GPU; every rank
// maps every peer's amplitude buffer and flag block through CUDA IPC (handles travel over
// torch.distributed on the host side), so kernels address peer HBM directly over NVLink.
//
//   * gates on local targets run the ordinary kernels; a control on a rank bit enables/disables the
//     whole rank; a DIAGONAL gate on a rank-bit target (Z,S,T,P,RZ, every QFT controlled phase) is a
//     rank-dependent phase -- no communication;
//   * a non-diagonal gate on a global qubit first swaps that qubit with a local one:
//     k_exchange swaps half of this shard with half of the partner's (rank ^ 2^g) IN PLACE through
//     peer loads/stores -- each rank moves one quarter shard out and one quarter in per direction,
//     no staging buffer (a 64 GiB receive buffer does not fit beside a 128 GiB shard);
//     the logical->physical qubit map absorbs the swap, nothing is swapped back;
//   * an uncontrolled SWAP gate is only a relabelling of the map (also on a single GPU);
//   * cross-rank synchronisation is a device-side flag barrier on the engine stream (no host sync);
//   * reductions: each rank publishes its partial in its flag block, all ranks read all partials
//     and add them in rank order (deterministic, identical on every rank).
#include <algorithm>
#include <functional>

#include "common.cuh"

namespace qi {

static const size_t kFlagBytes = 64 * 1024;      // per-rank flag block
static const int kScratchOff = 64;               // in 8-byte words: [0..63] barrier flags, then 2 x 512 doubles
static const int kScratchDoubles = 512;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct PeerPtrs { unsigned long long* f[8]; };

// thread j: tell peer j "rank `me` reached epoch e", then wait until peer j has told me the same
// A peer that has not arrived within `timeout_cycles` (option "peer_timeout_s", default 120 s) is dead or wedged: the
// kernel records the epoch and TRAPS -- the exchange behind the barrier must not run against a shard that is still in
// flight, so the context is poisoned and every later call on this rank fails loudly (QI_ERR_CUDA) instead of computing on
// half-exchanged data.
__global__ void k_barrier(PeerPtrs peers, unsigned long long* mine, int me, int world, unsigned long long epoch,
                          unsigned long long* timeout_flag, long long timeout_cycles) {
    int j = threadIdx.x;
    if (j >= world) return;
    __threadfence_system();
    st_release_sys(peers.f[j] + me, epoch);
    long long t0 = clock64();
    while (ld_acquire_sys(mine + j) < epoch) {
        if (clock64() - t0 > timeout_cycles) {
            *timeout_flag = epoch;
            __threadfence_system();
            __trap();
        }
    }
    __threadfence_system();
}

// In-place swap of k rank bits G with k local bits L (k = 1: a pairwise half-shard swap; k > 1: an
// all-to-all that moves (1 - 2^-k) of every shard).  Element (rank bits rho, local bits lambda) trades
// places with element (rank bits lambda, local bits rho); elements with lambda == rho stay.  For the
// block this rank trades with the rank whose G bits spell lambda, one launch of this kernel swaps
//   mine[expand(k) | or_mine]  <->  peer[expand(k) | or_peer]
// over the half of the block selected by a further local bit h (folded into or_mine / or_peer); the peer
// handles the other half, so every element pair is touched by exactly one rank and the work is balanced.
__global__ void __launch_bounds__(256) k_exchange(amp_t* __restrict__ mine, amp_t* __restrict__ peer, uint64_t total,
                                                  BitInsert ins, uint64_t or_mine, uint64_t or_peer) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += 4 * stride) {
        uint64_t j[4];
        amp_t a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) j[u] = (k + u * stride < total) ? expand_index(k + u * stride, ins) : ~0ull;
#pragma unroll
        for (int u = 0; u < 4; u++) if (j[u] != ~0ull) { a[u] = mine[j[u] | or_mine]; b[u] = peer[j[u] | or_peer]; }
#pragma unroll
        for (int u = 0; u < 4; u++) if (j[u] != ~0ull) { mine[j[u] | or_mine] = b[u]; peer[j[u] | or_peer] = a[u]; }
    }
}

static int log2i(int w) { int p = 0; while ((1 << p) < w) p++; return p; }

static int barrier(qi_state* s) {
    Context& c = ctx();
    if (!s->attached) return fail(QI_ERR_PEER, 0, 0, "shard is not attached to its peers (qi_shard_attach)");
    PeerPtrs pp;
    for (int r = 0; r < 8; r++) pp.f[r] = s->peer_flags[r];
    s->epoch++;
    LaunchScope ls(KF_BARRIER, 0.0);
    const long long cycles = (long long)std::max(1, c.opt_peer_timeout_s) * 2000000000ll;      // clock64 ticks at <= 2 GHz
    k_barrier<<<1, 32, 0, c.stream>>>(pp, s->flags, s->rank, s->world, s->epoch, s->flags + 32, cycles);
    return check_launch("k_barrier");
}

// position of logical qubit q / logical qubit at physical position p
static int logical_at(const qi_state* s, int p) {
    for (uint32_t q = 0; q < s->num_qubits; q++) if (s->phys[q] == p) return (int)q;
    return -1;
}

// swap global positions G[i] with local positions L[i] (all i at once)
int exchange_multi(qi_state* s, const std::vector<int>& G, const std::vector<int>& L) {
    Context& c = ctx();
    const int nl = (int)s->n_local, k = (int)G.size();
    if (k == 0) return QI_OK;
    for (int i = 0; i < k; i++)
        if (G[i] < nl || L[i] >= nl || L[i] < 0) return fail(QI_ERR_PEER, 0, 0, "bad exchange positions");
    int h = nl - 1;
    while (h >= 0 && std::find(L.begin(), L.end(), h) != L.end()) h--;
    if (h < 0) return fail(QI_ERR_PEER, 0, 0, "shard too small to exchange");
    uint64_t rho = 0;                                   // this rank's value on the G bits
    for (int i = 0; i < k; i++) rho |= (uint64_t)((s->rank >> (G[i] - nl)) & 1) << i;
    auto deposit = [&](uint64_t v) { uint64_t m = 0; for (int i = 0; i < k; i++) m |= ((v >> i) & 1ull) << L[i]; return m; };
    std::vector<int> zeros(L);
    zeros.push_back(h);
    BitInsert ins = make_insert(zeros, {});
    const uint64_t total = s->len >> (k + 1);
    QI_TRY(barrier(s));                                 // everyone's earlier kernels are complete
    // XOR schedule: in step d every rank trades with rank value rho ^ d, a perfect matching, so no
    // GPU's NVLink port is ever the target of more than one peer at a time
    for (uint64_t d = 1; d < (1ull << k); d++) {
        const uint64_t lambda = rho ^ d;
        int partner = s->rank;
        for (int i = 0; i < k; i++) partner = (partner & ~(1 << (G[i] - nl))) | ((int)((lambda >> i) & 1) << (G[i] - nl));
        const uint64_t hsel = (rho < lambda) ? 0ull : (1ull << h);     // the lower rank value takes the h = 0 half
        LaunchScope ls(KF_EXCHANGE, 16.0 * (double)s->len / (double)(1ull << k));
        k_exchange<<<c.sm_count * 8, 256, 0, c.stream>>>(s->d, s->peer_amp[partner], total, ins, deposit(lambda) | hsel, deposit(rho) | hsel);
        QI_TRY(check_launch("k_exchange"));
    }
    QI_TRY(barrier(s));                                 // every partner's half has landed before anything reads it
    // per direction and GPU: (1 - 2^-k) of a shard (half of it written/read by this rank, half by its partners)
    const uint64_t moved = (uint64_t)((double)(16ull * s->len) * (1.0 - 1.0 / (double)(1ull << k)));
    s->bytes_sent += moved;
    s->bytes_recv += moved;
    s->exchanges++;
    s->exchanged_qubits += (uint64_t)k;
    for (int i = 0; i < k; i++) {
        int qg = logical_at(s, G[i]), ql = logical_at(s, L[i]);
        if (qg >= 0) s->phys[qg] = (uint8_t)L[i];
        if (ql >= 0) s->phys[ql] = (uint8_t)G[i];
    }
    return QI_OK;
}

int exchange_global_local(qi_state* s, int global_phys, int local_phys) {
    return exchange_multi(s, std::vector<int>{global_phys}, std::vector<int>{local_phys});
}

// choose the local position to evict: the one whose logical qubit is needed (non-diagonally) latest
// index (relative to `upcoming`) of the next gate that uses logical qubit q non-diagonally; ~0 if none
static uint64_t next_nuse(int q, const qi_gate* upcoming, uint64_t n_upcoming) {
    for (uint64_t i = 0; i < n_upcoming && i < 4096; i++) {
        const qi_gate& g = upcoming[i];
        bool diag = g.kind == QI_GATE_Z || g.kind == QI_GATE_S || g.kind == QI_GATE_SDG || g.kind == QI_GATE_T ||
                    g.kind == QI_GATE_TDG || g.kind == QI_GATE_P || g.kind == QI_GATE_RZ || g.kind == QI_GATE_I;
        bool lazy_swap = g.kind == QI_GATE_SWAP && g.num_controls == 0;
        if (diag || lazy_swap) continue;
        bool hit = (int)g.targets[0] == q || (g.kind == QI_GATE_SWAP && (int)g.targets[1] == q) ||
                   (g.kind == QI_GATE_MATCHGATE && (int)g.targets[0] + 1 == q);
        if (hit) return i;
    }
    return ~0ull;
}

static int pick_local_slot(const qi_state* s, uint64_t avoid_phys, const qi_gate* upcoming, uint64_t n_upcoming,
                           uint64_t* evicted_next = nullptr) {
    const int nl = (int)s->n_local;
    int best = -1;
    uint64_t best_next = 0;
    for (int p = nl - 1; p >= 0; p--) {
        if ((avoid_phys >> p) & 1) continue;
        uint64_t next = next_nuse(logical_at(s, p), upcoming, n_upcoming);
        // prefer high positions on ties (p counts down), and keep the lane qubits 0..4 unless nothing else is free
        if (best < 0 || next > best_next || (next == best_next && p >= 5 && best < 5)) { best = p; best_next = next; }
        if (next == ~0ull && p >= 5) break;
    }
    if (evicted_next) *evicted_next = best_next;
    return best;
}

static bool is_diag_kind(int k) {
    return k == QI_GATE_Z || k == QI_GATE_S || k == QI_GATE_SDG || k == QI_GATE_T || k == QI_GATE_TDG || k == QI_GATE_P ||
           k == QI_GATE_RZ || k == QI_GATE_I;
}

// physical positions (of this gate's non-diagonal targets) that sit in the rank bits
static uint64_t global_targets(const qi_state* s, const qi_gate* g) {
    if (is_diag_kind(g->kind)) return 0;
    if (g->kind == QI_GATE_SWAP && g->num_controls == 0) return 0;     // relabelled
    uint64_t m = 0;
    const int nl = (int)s->n_local;
    auto add = [&](uint32_t q) { int p = s->phys[q]; if (p >= nl) m |= 1ull << p; };
    add(g->targets[0]);
    if (g->kind == QI_GATE_SWAP) add(g->targets[1]);
    if (g->kind == QI_GATE_MATCHGATE) add(g->targets[0] + 1);
    return m;
}


static uint64_t target_positions(const qi_state* s, const qi_gate* g) {
    uint64_t m = 1ull << s->phys[g->targets[0]];
    if (g->kind == QI_GATE_SWAP) m |= 1ull << s->phys[g->targets[1]];
    if (g->kind == QI_GATE_MATCHGATE) m |= 1ull << s->phys[g->targets[0] + 1];
    return m;
}

// Decide which (global, local) position pairs to swap before gate `g` can run: the gate's own global
// targets, plus (the fused queue is being flushed anyway) the other global qubits that upcoming gates use
// non-diagonally, each only if the qubit it evicts is needed later than the qubit it brings in.
// Works on a scratch copy of the qubit map so that the engine and the host-only planner share it.
static int plan_exchange(const qi_state* s, const qi_gate* g, uint64_t rem, std::vector<int>* G, std::vector<int>* L) {
    qi_state t;
    t.num_qubits = s->num_qubits; t.n_local = s->n_local; t.len = s->len; t.world = s->world; t.rank = s->rank;
    memcpy(t.phys, s->phys, sizeof(t.phys));
    auto relabel = [&](int gp, int lp) {
        int qg = logical_at(&t, gp), ql = logical_at(&t, lp);
        if (qg >= 0) t.phys[qg] = (uint8_t)lp;
        if (ql >= 0) t.phys[ql] = (uint8_t)gp;
        G->push_back(gp);
        L->push_back(lp);
    };
    uint64_t used_local = 0, used_global = 0, gm;
    while ((gm = global_targets(&t, g) & ~used_global) != 0) {
        int gp = 63 - __builtin_clzll(gm);
        int lp = pick_local_slot(&t, target_positions(&t, g) | used_local, g + 1, rem);
        if (lp < 0) return fail(QI_ERR_PEER, 0, 0, "no local qubit available for the exchange");
        relabel(gp, lp);
        used_local |= 1ull << lp;
        used_global |= 1ull << gp;
    }
    uint64_t protect = target_positions(&t, g) | used_local;
    for (uint64_t j = 0; j < rem && j < 512; j++) {
        const qi_gate* gj = g + 1 + j;
        if (gj->kind < QI_GATE_H || gj->kind > QI_GATE_MATCHGATE) break;
        uint64_t gmj = global_targets(&t, gj) & ~used_global;
        while (gmj) {
            int gp = 63 - __builtin_clzll(gmj);
            gmj &= ~(1ull << gp);
            uint64_t ev_next = 0;
            int lp = pick_local_slot(&t, protect | target_positions(&t, gj), g + 1, rem, &ev_next);
            if (lp < 0 || ev_next <= j) continue;
            relabel(gp, lp);
            protect |= 1ull << lp;
            used_global |= 1ull << gp;
        }
    }
    return QI_OK;
}


// ---- staged execution of a gate list on a sharded state ------------------------------------------------
// A global<->local exchange costs about as much as four fused passes over the shard, so the gate list is cut
// into STAGES: under the current qubit layout a stage takes every gate that (a) has no non-diagonal target
// on a rank bit and (b) commutes with every gate deferred so far (on each shared qubit both act diagonally --
// the rule the window scheduler uses).  Gates behind a deferred gate keep flowing as long as they stay
// outside its light cone, so one exchange serves a deep trapezoid of the circuit instead of one layer.
// Then ONE multi-qubit exchange brings in what the first deferred gate needs (plus what the lookahead over
// the deferred list says will be needed soon) and the next stage starts.  Host-only decisions that depend on
// the gate list alone: every rank takes the same ones, and the planner (qi_shard_plan) replays them.
void logical_uses(const qi_gate& g, uint64_t* n_use, uint64_t* d_use) {
    *n_use = 0;
    *d_use = 0;
    for (uint32_t c = 0; c < g.num_controls; c++) *d_use |= 1ull << g.controls[c];
    if (is_diag_kind(g.kind)) { *d_use |= 1ull << g.targets[0]; return; }
    *n_use |= 1ull << g.targets[0];
    if (g.kind == QI_GATE_SWAP) *n_use |= 1ull << g.targets[1];
    if (g.kind == QI_GATE_MATCHGATE) *n_use |= 1ull << (g.targets[0] + 1);
}

// run(take): execute gates[take[0]], gates[take[1]], ... under the current layout.  exchange(G, L): swap the
// global positions G with the local positions L and update s->phys.
template <typename RunFn, typename ExchangeFn>
static int staged_walk(qi_state* s, const qi_gate* gates, uint64_t count, RunFn run, ExchangeFn exchange) {
    std::vector<uint64_t> pending(count), rest, take;
    for (uint64_t i = 0; i < count; i++) pending[i] = i;
    std::vector<qi_gate> ahead;
    while (!pending.empty()) {
        take.clear();
        rest.clear();
        uint64_t def_any = 0, def_n = 0;
        for (uint64_t i : pending) {
            uint64_t n_use, d_use;
            logical_uses(gates[i], &n_use, &d_use);
            const bool ok = global_targets(s, &gates[i]) == 0 && (n_use & def_any) == 0 && (d_use & def_n) == 0;
            if (ok) take.push_back(i);
            else { rest.push_back(i); def_any |= n_use | d_use; def_n |= n_use; }
        }
        QI_TRY(run(take));
        if (rest.empty()) break;
        // the first deferred gate is deferred only because of a global target: bring it in (lookahead = the
        // deferred list, in order)
        ahead.clear();
        for (uint64_t i : rest) ahead.push_back(gates[i]);
        std::vector<int> G, L;
        QI_TRY(plan_exchange(s, ahead.data(), ahead.size() - 1, &G, &L));
        if (G.empty()) return fail(QI_ERR_PEER, 0, 0, "staged execution made no progress");
        QI_TRY(exchange(G, L));
        pending.swap(rest);
    }
    return QI_OK;
}

// engine entry (qi_apply_circuit on a sharded state): segments between lazily relabelled SWAPs are staged
int apply_circuit_sharded(qi_state* s, const qi_gate* gates, uint64_t count, bool use_window) {
    Context& c = ctx();
    std::vector<PhysGate> phys_run;
    auto run = [&](const std::vector<uint64_t>& take, const qi_gate* seg) -> int {
        phys_run.clear();
        for (uint64_t i : take) {
            PhysGate pg;
            bool skip = false;
            QI_TRY(prepare_gate(s, &seg[i], &pg, &skip));
            if (!skip && pg.kind != IK_NOP) phys_run.push_back(pg);
        }
        if (phys_run.empty()) return QI_OK;
        if (use_window) return run_circuit_windowed(s, phys_run);
        for (const PhysGate& g : phys_run) QI_TRY(launch_simple_gate(s, g));
        return QI_OK;
    };
    uint64_t first = 0;
    while (first < count) {
        uint64_t end = first;
        while (end < count && !(c.opt_lazy_swap && gates[end].kind == QI_GATE_SWAP && gates[end].num_controls == 0)) end++;
        const qi_gate* seg = gates + first;
        QI_TRY(staged_walk(s, seg, end - first,
                           [&](const std::vector<uint64_t>& take) { return run(take, seg); },
                           [&](const std::vector<int>& G, const std::vector<int>& L) { return exchange_multi(s, G, L); }));
        if (end < count) std::swap(s->phys[gates[end].targets[0]], s->phys[gates[end].targets[1]]);    // the lazy SWAP itself
        first = end + 1;
    }
    return QI_OK;
}

// logical record -> physical record on this rank
int shard_prepare_gate(qi_state* s, const qi_gate* g, PhysGate* o, bool* skip) {
    const int nl = (int)s->n_local;
    *skip = false;
    o->cmask = 0;
    for (uint32_t c = 0; c < g->num_controls; c++) {
        int p = s->phys[g->controls[c]];
        if (p >= nl) {
            if (!((s->rank >> (p - nl)) & 1)) { *skip = true; return QI_OK; }   // control on a rank bit that is 0 here
        } else o->cmask |= 1ull << p;
    }
    int t0 = s->phys[g->targets[0]];
    o->t1 = -1;
    if (o->kind == IK_DIAG || o->kind == IK_RZ) {
        if (t0 >= nl) {
            const int bit = (s->rank >> (t0 - nl)) & 1;
            if (o->kind == IK_DIAG) { if (!bit) { *skip = true; return QI_OK; } }
            else { if (bit) { o->p[0] = o->p[2]; o->p[1] = o->p[3]; } o->kind = IK_DIAG; }   // RZ: this rank's phase
            t0 = -1;
        }
        o->t0 = t0;
        return QI_OK;
    }
    if (t0 >= nl) return fail(QI_ERR_PEER, 0, 0, "internal: non-diagonal target still global");
    o->t0 = t0;
    if (g->kind == QI_GATE_SWAP) o->t1 = s->phys[g->targets[1]];
    if (g->kind == QI_GATE_MATCHGATE) o->t1 = s->phys[g->targets[0] + 1];
    if (o->t1 >= nl) return fail(QI_ERR_PEER, 0, 0, "internal: second target still global");
    return QI_OK;
}

// make every X/Y qubit of a Pauli term local (exchanges as needed)
int shard_localise_mask(qi_state* s, const qi_pauli_term* t) {
    const int nl = (int)s->n_local;
    for (;;) {
        uint64_t avoid = 0;
        int gp = -1;
        for (uint32_t i = 0; i < t->num_ops; i++) {
            if (t->paulis[i] == 3) continue;
            int p = s->phys[t->qubits[i]];
            if (p >= nl) gp = p; else avoid |= 1ull << p;
        }
        if (gp < 0) return QI_OK;
        int lp = pick_local_slot(s, avoid, nullptr, 0);
        if (lp < 0) return fail(QI_ERR_PEER, 0, 0, "no local qubit available for the exchange");
        QI_TRY(exchange_global_local(s, gp, lp));
    }
}

// ---- staged execution of a sequence of Pauli exponentials on a sharded state ----------------------------
// Same idea as staged_walk, with the exact Pauli commutation test: under the current layout a stage takes every
// term whose X/Y factors are all local and that commutes with every term deferred so far (two strings commute
// iff they anticommute on an even number of qubits); then ONE multi-qubit exchange brings in the global X/Y
// qubits the deferred terms need next, evicting the local qubits whose next X/Y use is furthest away.
// lx / lz: per term, LOGICAL qubit masks of its X-or-Y and Y-or-Z factors.  run(take) executes the terms take[..]
// (indices into lx) under the current layout; dry = planner mode (relabel only, no device access).
int shard_pauli_walk(qi_state* s, const std::vector<uint64_t>& lx, const std::vector<uint64_t>& lz,
                     const std::function<int(const std::vector<size_t>&)>& run, bool dry, uint64_t* exchanges,
                     const std::function<void(const std::vector<int>&, const std::vector<int>&)>& on_exchange) {
    const int nl = (int)s->n_local;
    const size_t kMaxDeferred = 512;
    std::vector<size_t> pending(lx.size()), rest, take, deferred;
    for (size_t i = 0; i < lx.size(); i++) pending[i] = i;
    auto x_is_local = [&](uint64_t xm) {
        for (uint32_t q = 0; q < s->num_qubits; q++)
            if (((xm >> q) & 1) && s->phys[q] >= nl) return false;
        return true;
    };
    while (!pending.empty()) {
        take.clear(); rest.clear(); deferred.clear();
        uint64_t def_support = 0;
        bool saturated = false;
        for (size_t k : pending) {
            bool ok = !saturated && x_is_local(lx[k]);
            if (ok && ((lx[k] | lz[k]) & def_support))
                for (size_t d : deferred)
                    if ((__builtin_popcountll(lx[d] & lz[k]) + __builtin_popcountll(lz[d] & lx[k])) & 1) { ok = false; break; }
            if (ok) take.push_back(k);
            else {
                rest.push_back(k);
                if (!saturated) {
                    deferred.push_back(k);
                    def_support |= lx[k] | lz[k];
                    if (deferred.size() >= kMaxDeferred) saturated = true;       // keep the test cheap: defer the tail wholesale
                }
            }
        }
        QI_TRY(run(take));
        if (rest.empty()) break;
        if (x_is_local(lx[rest[0]])) { pending.swap(rest); continue; }          // deferred by the cap only: next stage takes it
        // global positions to bring in: those of the first deferred term, plus the other global qubits the next
        // deferred terms flip; local positions to evict: furthest next X/Y use (never a qubit the first term flips)
        auto next_use = [&](int logical) -> size_t {
            for (size_t j = 0; j < rest.size() && j < 4096; j++) if ((lx[rest[j]] >> logical) & 1) return j;
            return ~(size_t)0;
        };
        std::vector<int> G, L;
        uint64_t want = lx[rest[0]];
        for (size_t j = 1; j < rest.size() && j < 64; j++) want |= lx[rest[j]];
        uint64_t used_local = 0;
        for (int gp = (int)s->num_qubits - 1; gp >= nl; gp--) {
            const int qg = logical_at(s, gp);
            if (qg < 0 || !((want >> qg) & 1)) continue;
            const bool mandatory = (lx[rest[0]] >> qg) & 1;
            int best = -1;
            size_t best_next = 0;
            for (int p = nl - 1; p >= 0; p--) {
                if ((used_local >> p) & 1) continue;
                const int ql = logical_at(s, p);
                if (ql >= 0 && ((lx[rest[0]] >> ql) & 1)) continue;
                const size_t nu = ql < 0 ? ~(size_t)0 : next_use(ql);
                if (best < 0 || nu > best_next || (nu == best_next && p >= 5 && best < 5)) { best = p; best_next = nu; }
            }
            if (best < 0) { if (mandatory) return fail(QI_ERR_PEER, 0, 0, "no local qubit available for the exchange"); continue; }
            if (!mandatory && best_next <= next_use(qg)) continue;              // the evicted qubit would be needed sooner
            G.push_back(gp);
            L.push_back(best);
            used_local |= 1ull << best;
        }
        if (G.empty()) return fail(QI_ERR_PEER, 0, 0, "staged Pauli execution made no progress");
        if (on_exchange) on_exchange(G, L);
        if (dry) {
            for (size_t k = 0; k < G.size(); k++) {
                int qg = logical_at(s, G[k]), ql = logical_at(s, L[k]);
                if (qg >= 0) s->phys[qg] = (uint8_t)L[k];
                if (ql >= 0) s->phys[ql] = (uint8_t)G[k];
            }
        } else {
            QI_TRY(exchange_multi(s, G, L));
        }
        if (exchanges) (*exchanges)++;
        pending.swap(rest);
    }
    return QI_OK;
}

// Host-only: the stages a sequence of Pauli exponentials runs in on `world` ranks, as a u64 stream a test replays on the
// CPU (tests/test_sharded_emulation.py): nstages, per stage phys[64], ntake, take[] (term indices), nex, G[], L[]; then the
// final phys[64].  Same decisions as the engine (shard_pauli_walk).
int debug_shard_pauli_stages(uint32_t total_qubits, int world, const std::vector<uint64_t>& lx, const std::vector<uint64_t>& lz,
                             std::vector<uint64_t>* rec) {
    qi_state s;
    s.num_qubits = total_qubits;
    s.n_local = total_qubits - log2i(world);
    s.len = 1ull << s.n_local;
    s.world = world;
    for (int i = 0; i < 64; i++) s.phys[i] = (uint8_t)i;
    rec->push_back(0);
    uint64_t nstages = 0;
    size_t ex_slot = 0;
    auto put_phys = [&]() { for (int i = 0; i < 64; i++) rec->push_back(s.phys[i]); };
    QI_TRY(shard_pauli_walk(&s, lx, lz,
        [&](const std::vector<size_t>& take) -> int {
            nstages++;
            put_phys();
            rec->push_back(take.size());
            for (size_t k : take) rec->push_back(k);
            ex_slot = rec->size();
            rec->push_back(0);
            return QI_OK;
        },
        true, nullptr,
        [&](const std::vector<int>& G, const std::vector<int>& L) {
            (*rec)[ex_slot] = G.size();
            for (int g : G) rec->push_back((uint64_t)g);
            for (int l : L) rec->push_back((uint64_t)l);
        }));
    put_phys();
    (*rec)[0] = nstages;
    return QI_OK;
}

// sum host values across ranks in rank order; every rank gets the identical result
int shard_allreduce_sum(qi_state* s, double* vals, int count) {
    if (s->world == 1) return QI_OK;
    Context& c = ctx();
    if (count > kScratchDoubles) return fail(QI_ERR_PEER, (uint64_t)count, 0, "allreduce payload too large");
    const int parity = (int)(s->reduce_count++ & 1);
    double* my_scratch = (double*)(s->flags + kScratchOff) + parity * kScratchDoubles;
    QI_CUDA(cudaMemcpyAsync(my_scratch, vals, count * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    QI_TRY(barrier(s));
    std::vector<double> all((size_t)s->world * count);
    for (int r = 0; r < s->world; r++) {
        const double* src = (const double*)(s->peer_flags[r] + kScratchOff) + parity * kScratchDoubles;
        QI_CUDA(cudaMemcpyAsync(all.data() + (size_t)r * count, src, count * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    }
    QI_CUDA(cudaStreamSynchronize(c.stream));
    unsigned long long timed_out = 0;
    QI_CUDA(cudaMemcpy(&timed_out, s->flags + 32, sizeof(timed_out), cudaMemcpyDeviceToHost));
    if (timed_out) return fail(QI_ERR_PEER, timed_out, 0, "device barrier timed out: a peer rank is gone");
    for (int i = 0; i < count; i++) {
        double acc = 0.0;
        for (int r = 0; r < s->world; r++) acc += all[(size_t)r * count + i];
        vals[i] = acc;
    }
    return QI_OK;
}

static int alloc_shard(uint32_t total_qubits, int rank, int world, qi_state** out) {
    if (!out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "out is NULL");
    if (world != 1 && world != 2 && world != 4 && world != 8) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)world, 0, "world must be 1, 2, 4 or 8");
    if (rank < 0 || rank >= world) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)rank, 0, "rank out of range");
    const int p = log2i(world);
    if (total_qubits == 0) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, 0, 0, "Invalid number of qubits: 0");
    if ((int)total_qubits < p + 7) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, total_qubits, 0, "a shard needs at least 7 local qubits");
    QI_TRY(ensure_ctx());
    qi_state* s = new qi_state();
    s->num_qubits = total_qubits;
    s->n_local = total_qubits - p;
    s->len = 1ull << s->n_local;
    s->consistent = true;
    s->rank = rank;
    s->world = world;
    for (int i = 0; i < 64; i++) s->phys[i] = (uint8_t)i;
    cudaError_t e = cudaMalloc(&s->d, s->len * sizeof(amp_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->flags, kFlagBytes);
    if (e == cudaSuccess) e = cudaMemset(s->flags, 0, kFlagBytes);
    if (e != cudaSuccess) { if (s->d) cudaFree(s->d); if (s->flags) cudaFree(s->flags); delete s; return cuda_fail(e, "cudaMalloc(shard)"); }
    s->peer_amp[rank] = s->d;
    s->peer_flags[rank] = s->flags;
    if (world == 1) s->attached = true;
    *out = s;
    return QI_OK;
}

}  // namespace qi

using namespace qi;

extern "C" {

void qi_shard_release(qi_state* s) {
    if (!s || s->world <= 1) { if (s && s->flags) { cudaFree(s->flags); s->flags = nullptr; } return; }
    for (int r = 0; r < s->world; r++) {
        if (r == s->rank) continue;
        if (s->peer_amp[r]) cudaIpcCloseMemHandle(s->peer_amp[r]);
        if (s->peer_flags[r]) cudaIpcCloseMemHandle(s->peer_flags[r]);
    }
    if (s->flags) cudaFree(s->flags);
    s->flags = nullptr;
}

int qi_shard_new_zero(uint32_t n, int rank, int world, qi_state** out) {
    QI_TRY(alloc_shard(n, rank, world, out));
    QI_TRY(fill_state(*out, make_double2(0.0, 0.0)));
    if (rank == 0) QI_TRY(set_amplitude(*out, 0, make_double2(1.0, 0.0)));
    return QI_OK;
}

int qi_shard_new_plus(uint32_t n, int rank, int world, qi_state** out) {
    QI_TRY(alloc_shard(n, rank, world, out));
    // 1/sqrt(2^n) with the reference's expression (state.rs:230); 2^n is exact in f64 up to n = 1023
    return fill_state(*out, make_double2(1.0 / std::sqrt(std::ldexp(1.0, (int)n)), 0.0));
}

int qi_shard_new_basis_n(uint32_t n, uint64_t k, int rank, int world, qi_state** out) {
    if (n >= 64 || k >= (1ull << n)) return fail(QI_ERR_INVALID_QUBIT_INDEX, k, n, "basis index out of range");
    QI_TRY(alloc_shard(n, rank, world, out));
    qi_state* s = *out;
    QI_TRY(fill_state(s, make_double2(0.0, 0.0)));
    if ((k >> s->n_local) == (uint64_t)rank) QI_TRY(set_amplitude(s, k & (s->len - 1), make_double2(1.0, 0.0)));
    return QI_OK;
}

int qi_shard_export(qi_state* s, uint8_t* handles) {
    if (!s || !handles) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == QI_IPC_HANDLE_BYTES, "IPC handle size");
    QI_TRY(ensure_ctx());
    cudaIpcMemHandle_t h;
    QI_CUDA(cudaIpcGetMemHandle(&h, s->d));
    memcpy(handles, &h, sizeof(h));
    QI_CUDA(cudaIpcGetMemHandle(&h, s->flags));
    memcpy(handles + QI_IPC_HANDLE_BYTES, &h, sizeof(h));
    return QI_OK;
}

int qi_shard_attach(qi_state* s, const uint8_t* all_handles) {
    if (!s || !all_handles) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    QI_TRY(ensure_ctx());
    for (int r = 0; r < s->world; r++) {
        if (r == s->rank) continue;
        cudaIpcMemHandle_t h;
        void* p = nullptr;
        memcpy(&h, all_handles + (size_t)r * 2 * QI_IPC_HANDLE_BYTES, sizeof(h));
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { set_error((uint64_t)e, (uint64_t)r, "cudaIpcOpenMemHandle(amplitudes of rank %d): %s", r, cudaGetErrorString(e)); cudaGetLastError(); return QI_ERR_PEER; }
        s->peer_amp[r] = (amp_t*)p;
        memcpy(&h, all_handles + (size_t)r * 2 * QI_IPC_HANDLE_BYTES + QI_IPC_HANDLE_BYTES, sizeof(h));
        e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { set_error((uint64_t)e, (uint64_t)r, "cudaIpcOpenMemHandle(flags of rank %d): %s", r, cudaGetErrorString(e)); cudaGetLastError(); return QI_ERR_PEER; }
        s->peer_flags[r] = (unsigned long long*)p;
    }
    s->attached = true;
    return QI_OK;
}

int qi_shard_rank(const qi_state* s) { return s ? s->rank : 0; }
int qi_shard_world(const qi_state* s) { return s ? s->world : 1; }

int qi_shard_comm_stats(const qi_state* s, uint64_t* sent, uint64_t* recv, uint64_t* exchanges) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    if (sent) *sent = s->bytes_sent;
    if (recv) *recv = s->bytes_recv;
    if (exchanges) *exchanges = s->exchanges;
    return QI_OK;
}

int qi_state_layout(const qi_state* s, uint8_t* phys, uint32_t* n_local) {
    if (!s || !phys) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    memcpy(phys, s->phys, 64);
    if (n_local) *n_local = s->n_local;
    return QI_OK;
}

// Host-only planner: how many global<->local exchanges a gate list needs on `world` ranks, following
// exactly the decisions the engine takes (no device access; used by tests and DESIGN.md tables).
int qi_shard_plan(uint32_t total_qubits, int world, const qi_gate* gates, uint64_t count, uint64_t* exchanges,
                  uint64_t* comm_free_global_gates, uint8_t* final_phys) {
    if (world != 1 && world != 2 && world != 4 && world != 8) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)world, 0, "world must be 1, 2, 4 or 8");
    qi_state s;
    s.num_qubits = total_qubits;
    s.n_local = total_qubits - log2i(world);
    s.len = 1ull << s.n_local;
    s.world = world;
    for (int i = 0; i < 64; i++) s.phys[i] = (uint8_t)i;
    uint64_t ex = 0, exq = 0, freeg = 0;
    for (uint64_t i = 0; i < count; i++) QI_TRY(validate_gate(&s, &gates[i]));
    std::vector<qi_gate> own;
    gates = normalise_gates(gates, count, &own);
    uint64_t first = 0;
    while (first < count) {
        uint64_t end = first;
        while (end < count && !(gates[end].kind == QI_GATE_SWAP && gates[end].num_controls == 0)) end++;
        const qi_gate* seg = gates + first;
        QI_TRY(staged_walk(&s, seg, end - first,
            [&](const std::vector<uint64_t>& take) -> int {
                for (uint64_t i : take) {
                    const qi_gate* g = &seg[i];
                    bool touches_global = s.phys[g->targets[0]] >= s.n_local;
                    for (uint32_t c = 0; c < g->num_controls; c++) touches_global |= s.phys[g->controls[c]] >= s.n_local;
                    if (touches_global) freeg++;
                }
                return QI_OK;
            },
            [&](const std::vector<int>& G, const std::vector<int>& L) -> int {
                for (size_t k = 0; k < G.size(); k++) {
                    int qg = logical_at(&s, G[k]), ql = logical_at(&s, L[k]);
                    if (qg >= 0) s.phys[qg] = (uint8_t)L[k];
                    if (ql >= 0) s.phys[ql] = (uint8_t)G[k];
                }
                ex++;
                exq += G.size();
                return QI_OK;
            }));
        if (end < count) std::swap(s.phys[gates[end].targets[0]], s.phys[gates[end].targets[1]]);
        first = end + 1;
    }
    if (exchanges) *exchanges = ex;
    (void)exq;
    if (comm_free_global_gates) *comm_free_global_gates = freeg;
    if (final_phys) memcpy(final_phys, s.phys, 64);
    return QI_OK;
}

// Host-only: the stages apply_circuit_sharded runs for a gate list on `world` ranks, as a u64 stream a test can replay
// on the CPU (tests/test_sharded_emulation.py): nstages, then per stage
//   phys[64] (logical -> physical map the stage runs under), ntake, take[ntake] (gate indices, circuit order),
//   nex, G[nex], L[nex] (the exchange after the stage: global positions G swapped with local positions L; 0 = none)
// and finally phys[64] after the last stage.  Same decisions as the engine (staged_walk, plan_exchange, lazy SWAPs).
int qi_debug_shard_stages(uint32_t total_qubits, int world, const qi_gate* gates, uint64_t count, uint64_t* out, uint64_t capacity,
                          uint64_t* used) {
    if (world != 1 && world != 2 && world != 4 && world != 8) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)world, 0, "world must be 1, 2, 4 or 8");
    if (count && !gates) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "gates is NULL");
    qi_state s;
    s.num_qubits = total_qubits;
    s.n_local = total_qubits - log2i(world);
    s.len = 1ull << s.n_local;
    s.world = world;
    for (int i = 0; i < 64; i++) s.phys[i] = (uint8_t)i;
    for (uint64_t i = 0; i < count; i++) QI_TRY(validate_gate(&s, &gates[i]));
    std::vector<qi_gate> own;
    gates = normalise_gates(gates, count, &own);
    std::vector<uint64_t> rec;
    rec.push_back(0);
    uint64_t nstages = 0;
    auto put_phys = [&]() { for (int i = 0; i < 64; i++) rec.push_back(s.phys[i]); };
    const bool lazy = ctx().opt_lazy_swap != 0;
    uint64_t first = 0;
    while (first < count) {
        uint64_t end = first;
        while (end < count && !(lazy && gates[end].kind == QI_GATE_SWAP && gates[end].num_controls == 0)) end++;
        const qi_gate* seg = gates + first;
        size_t ex_slot = 0;            // where the pending stage's exchange record starts
        QI_TRY(staged_walk(&s, seg, end - first,
            [&](const std::vector<uint64_t>& take) -> int {
                nstages++;
                put_phys();
                rec.push_back(take.size());
                for (uint64_t i : take) rec.push_back(first + i);
                ex_slot = rec.size();
                rec.push_back(0);                      // no exchange unless one follows
                return QI_OK;
            },
            [&](const std::vector<int>& G, const std::vector<int>& L) -> int {
                rec[ex_slot] = G.size();
                for (int g : G) rec.push_back((uint64_t)g);
                for (int l : L) rec.push_back((uint64_t)l);
                for (size_t k = 0; k < G.size(); k++) {
                    int qg = logical_at(&s, G[k]), ql = logical_at(&s, L[k]);
                    if (qg >= 0) s.phys[qg] = (uint8_t)L[k];
                    if (ql >= 0) s.phys[ql] = (uint8_t)G[k];
                }
                return QI_OK;
            }));
        if (end < count) std::swap(s.phys[gates[end].targets[0]], s.phys[gates[end].targets[1]]);
        first = end + 1;
    }
    put_phys();
    rec[0] = nstages;
    if (used) *used = rec.size();
    if (rec.size() > capacity || !out) return fail(QI_ERR_INVALID_ARGUMENT, rec.size(), capacity, "buffer too small");
    memcpy(out, rec.data(), rec.size() * sizeof(uint64_t));
    return QI_OK;
}

// Host-only planner for Pauli-exp sequences (`repeats` repetitions of the term list, e.g. Trotter steps): number of
// exchanges the engine performs on `world` ranks (no device access; same decisions on every rank).
int qi_shard_plan_pauli(uint32_t total_qubits, int world, const qi_pauli_term* terms, uint64_t count, uint64_t repeats,
                        uint64_t* exchanges, uint64_t* stages) {
    if (world != 1 && world != 2 && world != 4 && world != 8) return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)world, 0, "world must be 1, 2, 4 or 8");
    if ((count && !terms) || !exchanges) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    qi_state s;
    s.num_qubits = total_qubits;
    s.n_local = total_qubits - log2i(world);
    s.len = 1ull << s.n_local;
    s.world = world;
    for (int i = 0; i < 64; i++) s.phys[i] = (uint8_t)i;
    std::vector<uint64_t> lx, lz;
    for (uint64_t r = 0; r < repeats; r++)
        for (uint64_t k = 0; k < count; k++) {
            uint64_t x = 0, z = 0;
            for (uint32_t i = 0; i < terms[k].num_ops; i++) {
                const uint32_t q = terms[k].qubits[i];
                if (q >= total_qubits) return fail(QI_ERR_INVALID_QUBIT_INDEX, q, total_qubits, "Invalid qubit index");
                if (terms[k].paulis[i] != 3) x |= 1ull << q;
                if (terms[k].paulis[i] != 1) z |= 1ull << q;
            }
            lx.push_back(x);
            lz.push_back(z);
        }
    *exchanges = 0;
    uint64_t nst = 0;
    QI_TRY(shard_pauli_walk(&s, lx, lz, [&](const std::vector<size_t>&) { nst++; return QI_OK; }, true, exchanges));
    if (stages) *stages = nst;
    return QI_OK;
}

}  // extern "C"
