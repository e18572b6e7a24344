// common.cuh -- shared internals of libqiron_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "qiron_b200.h"

namespace qi {

typedef double2 amp_t;  // Complex<f64>: .x = re, .y = im

// ---- error plumbing ------------------------------------------------------------------------
void set_error(uint64_t p0, uint64_t p1, const char* fmt, ...);
int fail(int status, uint64_t p0 = 0, uint64_t p1 = 0, const char* msg = "");
int cuda_fail(cudaError_t e, const char* what);

#define QI_CUDA(expr)                                          \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) return ::qi::cuda_fail(_e, #expr); \
    } while (0)

#define QI_TRY(expr)                  \
    do {                              \
        int _s = (expr);              \
        if (_s != QI_OK) return _s;   \
    } while (0)

// ---- engine context --------------------------------------------------------------------------
enum KernelFamily {
    KF_INIT = 0, KF_PAIR, KF_DIAG, KF_SWAP, KF_MATCH, KF_WINDOW, KF_PAULI, KF_PAULI_EXP, KF_EXPECT,
    KF_REDUCE, KF_ELEMENTWISE, KF_PROB, KF_SCAN, KF_SAMPLE, KF_COLLAPSE, KF_EXCHANGE, KF_BARRIER,
    KF_PAULI_WINDOW, KF_TILE, KF_TILE_JIT, KF_COUNT
};
extern const char* const kFamilyNames[KF_COUNT];

struct ProfEvent { cudaEvent_t start, stop; int family; };

struct Context {
    bool ready = false;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    // scratch for reductions: partial sums (double2 per block per slot) + result
    double* d_partials = nullptr;     // kPartialSlots * 2 doubles
    double* d_result = nullptr;       // small result area (device)
    double* h_result = nullptr;       // pinned host mirror
    size_t partial_capacity = 0;      // in doubles
    // options
    int opt_path = 0;                 // 0 auto, 1 simple kernels, 2 window kernels
    int opt_fuse = 1;
    int opt_profile = 0;
    int opt_window_regs = 0;          // 0 = default
    int opt_lazy_swap = 1;            // uncontrolled SWAP = relabelling of the qubit map
    int opt_tma = 0;                  // window passes: 1 = TMA-prefetched persistent kernel (measured 2.4% slower), 0 = direct loads
    int opt_absorb = 1;               // fold a CNOT into the neighbouring single-qubit gate on its target (window.cu)
    int opt_late_tables = 1;          // window passes: unconditional phase tables placed as late as possible (fewest tables per pass)
    int opt_prefetch = 0;             // lean instantiations only: L2 prefetch of the warp's next tile (unmeasured: off)
    int opt_cz_rewrite = 1;           // fused executor: a controlled X next to a Hadamard on its target becomes a controlled Z (bit-exact)
    int opt_tile = 1;                 // window passes on the CTA-tile kernel (k_tile: 11 qubits per pass, rounds regrouped through shared memory)
    int opt_peer_timeout_s = 120;     // sharded states: a peer missing at a device barrier for this long traps the kernel (shard.cu)
    int opt_tile_absorb = 0;          // tile passes: CNOT absorption (two predicated half-ops per pair) -- off: register swaps are cheaper there
    int opt_tile_slide = 1;           // tile passes leave the qubits the next tile wants at positions 0..4 (relabelling the qubit map)
    int opt_tile_min_gates = 3;       // ... and runs of at least this many gates (a lone gate is a pure HBM pass: k_window streams it best)
    int opt_tile_min_qubits = 18;     // ... for states with at least this many local qubits (>= 11)
    int opt_jit = 1;                  // tile passes as circuit-specialised straight-line PTX (tile_jit.cuh): 0 never, 1 assembled in the background
                                      // after a pass structure is first seen (k_tile runs it meanwhile), 2 assembled before the first launch
    int opt_jit_min_qubits = 24;      // jit = 1 only for states with at least this many local qubits (a module costs ~1 s of one host core)
    int opt_jit_ctas = 4;             // resident CTAs per SM the modules are assembled for (4 = 128 registers, 3 = 168)
    int opt_jit_smem_kb = 0;          // modules request at least this much dynamic shared memory (56 = at most 4 CTAs per SM however few registers a small module needs)
    int opt_jit_stage = 0;            // modules bring the CTA's next tile into shared memory with bulk async copies (cp.async.bulk + mbarrier) while the current one is computed
    int opt_jit_prefetch = 0;         // modules prefetch the CTA's next tile into L2 while the current one is computed
    int opt_tile_bfs_by_use = 1;      // tile scheduler: candidate tiles grow through neighbours in program order of their first non-diagonal use
    int opt_pauli_unit = 1;           // fused Pauli-exp passes in unit form (one FMA per component and op, one scale per pass)
    int opt_tile_restore = 0;         // states that run on modules: sliding tiles + relabel-only passes that restore the layout at the end of a run
    int opt_tile_carry = 1;           // the scalar the unit-form / P-form ops leave out travels across the launches of a run and is applied once
    int opt_tile_pform = 4;           // tile passes: unconditional diagonal groups of at most this many qubits are applied as P-form phase ops instead of a table (0 = always a table)
    int opt_tile_lean = 1;            // tile passes: uncontrolled H / RY / RX in unit form (2 FP64 instructions per amplitude instead of 4), one scale op per launch
    int opt_jit_groups = 1;           // tiles a module's CTA works on side by side (1, 2 or 4 x 128 threads; measured: 1 = 2 > 4)
    int opt_debug_ptx = 0;            // qi_debug_lower returns the PTX of the tile passes instead of the op blob (CPU-side syntax checks)
    int opt_lean = 0;                 // window passes: unit-form H / RX / real 2x2 with one deferred scale per pass (unmeasured: off)
    int64_t opt_pool_mb = 4096;       // device-buffer cache: at most this many MiB are kept for reuse (0 = off)
    // stats
    uint64_t launches[KF_COUNT] = {0};
    double alg_bytes[KF_COUNT] = {0};
    double total_ms[KF_COUNT] = {0};
    std::vector<ProfEvent> pending;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t timer_start = nullptr, timer_stop = nullptr;
    // op-program staging for the window executor
    void* h_ops = nullptr;
    void* d_ops = nullptr;
    size_t ops_cap = 0;
    cudaEvent_t ops_event = nullptr;
    // host-resident execution (host_pipeline.cu): copy stream + per-chunk events
    int opt_host_chunk_qubits = 3;    // the top k qubits index 2^k chunks that are uploaded / downloaded one by one (0 = no pipelining)
    int opt_host_min_qubits = 26;     // smaller states take the plain upload / execute / download sequence
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_sync = nullptr;
    cudaEvent_t chunk_in[16] = {nullptr}, chunk_out[16] = {nullptr};
};
Context& ctx();
int ensure_ctx();
int ensure_partials(size_t doubles);
// Device buffers through a small size-keyed cache: cudaMalloc/cudaFree cost up to milliseconds each (and cudaFree
// synchronises), which would dominate the reference's functional API (`&self -> State` clones once per gate) on small
// states.  Everything runs on the one engine stream, so a cached buffer can be handed out again without a sync.
int dev_alloc(void** p, size_t bytes);
void dev_free(void* p, size_t bytes);
void dev_pool_trim();                 // release every cached buffer

// Kernel launch bookkeeping: counts launches, accumulates algorithmic bytes and, when the
// "profile" option is on, brackets the launch with CUDA events on the engine stream.
struct LaunchScope {
    int family;
    cudaEvent_t start = nullptr, stop = nullptr;
    LaunchScope(int family, double bytes);
    ~LaunchScope();
};
int check_launch(const char* what);

// ---- state -------------------------------------------------------------------------------------
}  // namespace qi

struct qi_state {
    qi::amp_t* d = nullptr;     // local amplitudes
    uint64_t len = 0;           // local amplitude count
    uint32_t num_qubits = 0;    // logical qubits of the whole (possibly sharded) state
    uint32_t n_local = 0;       // index bits held locally (== num_qubits when world == 1)
    bool consistent = true;     // len == 2^n_local (false only for State{..} literals, state_tests.rs:145)
    // sharding
    int rank = 0, world = 1;
    uint8_t phys[64];           // logical qubit -> physical bit position (identity unless exchanged)
    qi::amp_t* peer_amp[8] = {nullptr};
    unsigned long long* flags = nullptr;          // this rank's flag/scratch block (device)
    unsigned long long* peer_flags[8] = {nullptr};
    unsigned long long epoch = 0;
    uint64_t bytes_sent = 0, bytes_recv = 0, exchanges = 0, exchanged_qubits = 0;
    uint64_t reduce_count = 0;          // parity selects the allreduce scratch buffer
    bool attached = false;
    bool identity_layout() const { for (uint32_t q = 0; q < num_qubits; q++) if (phys[q] != q) return false; return true; }
};

namespace qi {

// ---- bit tricks --------------------------------------------------------------------------------
// Expand a compact index over the free bits into a full index: insert a zero bit at each of the
// `n` ascending positions `pos`, then OR `ones` (the positions that are fixed to 1).
struct BitInsert {
    int n;
    uint8_t pos[62];
    uint64_t ones;
};

__host__ __device__ __forceinline__ uint64_t insert_zero(uint64_t k, int p) {
    return ((k >> p) << (p + 1)) | (k & ((1ull << p) - 1ull));
}

__host__ __device__ __forceinline__ uint64_t expand_index(uint64_t k, const BitInsert& b) {
#pragma unroll 1
    for (int i = 0; i < b.n; i++) k = insert_zero(k, b.pos[i]);
    return k | b.ones;
}

BitInsert make_insert(const std::vector<int>& zero_positions, const std::vector<int>& one_positions);

// ---- complex helpers (same operation order as num-complex; FMA contraction is left to nvcc) ------
__host__ __device__ __forceinline__ amp_t cmul(amp_t a, amp_t b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ amp_t cadd(amp_t a, amp_t b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ amp_t csub(amp_t a, amp_t b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ amp_t cscale(double s, amp_t a) { return make_double2(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ amp_t cneg(amp_t a) { return make_double2(-a.x, -a.y); }
__host__ __device__ __forceinline__ amp_t cconj(amp_t a) { return make_double2(a.x, -a.y); }
// multiply by i^k, exact
__host__ __device__ __forceinline__ amp_t mul_i_pow(amp_t a, int k) {
    switch (k & 3) {
        case 0: return a;
        case 1: return make_double2(-a.y, a.x);
        case 2: return make_double2(-a.x, -a.y);
        default: return make_double2(a.y, -a.x);
    }
}

// host complex functions with num-complex's formulas
static inline amp_t h_cexp(amp_t a) { double e = std::exp(a.x); return make_double2(e * std::cos(a.y), e * std::sin(a.y)); }
static inline amp_t h_ccosh(amp_t a) { return make_double2(std::cosh(a.x) * std::cos(a.y), std::sinh(a.x) * std::sin(a.y)); }
static inline amp_t h_csinh(amp_t a) { return make_double2(std::sinh(a.x) * std::cos(a.y), std::cosh(a.x) * std::sin(a.y)); }

// splitmix64 shared-seed stream: u = (x >> 11) * 2^-53
__host__ __device__ __forceinline__ uint64_t splitmix64_at(uint64_t seed, uint64_t k) {
    uint64_t z = seed + (k + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double uniform_at(uint64_t seed, uint64_t k) {
    return (double)(splitmix64_at(seed, k) >> 11) * (1.0 / 9007199254740992.0);
}

// 128-bit global accesses
// (streaming cache hints: a state is read once and written once per pass and is far larger than L2;
//  measured +2.5% on 30-qubit passes)
__device__ __forceinline__ amp_t ld_amp(const amp_t* p) { return __ldcs(p); }
__device__ __forceinline__ void st_amp(amp_t* p, amp_t v) { __stcs(p, v); }

// internal gate kinds after parameter resolution
enum { IK_NOP = 0, IK_H, IK_X, IK_Y, IK_U2, IK_DIAG, IK_RZ, IK_SWAP, IK_MATCH };

// ---- physical (already validated, already mapped) single-operator record ----------------------
struct PhysGate {
    int kind;            // IK_*
    int t0, t1;          // physical target bit positions (t1 only for SWAP; MATCHGATE: t1 = partner)
    uint64_t cmask;      // physical control bits that must be 1 (local bits only after rank filtering)
    double p[8];         // resolved numeric parameters (see resolve_gate)
};

// one exp(alpha P) factor, reduced to masks over PHYSICAL local bit positions (pauli.cu)
struct PauliExp {
    uint64_t x, z;       // flipped bits (X, Y factors) / sign bits (Y, Z factors), local bits only
    int k0;              // (3 nY + 2 popc(rank bits & z)) & 3
    amp_t ch, sh;        // cosh(alpha), sinh(alpha); an empty string is (exp(alpha), 0) with x = z = 0
};
// pauli.cu
int pauli_exp_single(qi_state* s, const PauliExp& t);          // one term, one pass (per-term kernels)
// pauli_window.cu
bool pauli_window_supported(const qi_state* s);
int run_pauli_exp_batch(qi_state* s, const std::vector<PauliExp>& seq);
int run_pauli_expect_batch(const qi_state* s, const std::vector<PauliExp>& terms, int grid, double2* partials, int max_groups,
                           int* groups_used, std::vector<size_t>* leftover);      // ch = coefficient of each term
int debug_pauli_schedule(const std::vector<PauliExp>& seq, std::vector<int>* terms_per_pass);
int debug_pauli_lower(const qi_state* s, const std::vector<PauliExp>& seq, std::vector<uint8_t>* blob);
int debug_expect_lower(const qi_state* s, const std::vector<PauliExp>& terms, std::vector<uint8_t>* blob);

// gates.cu
int validate_gate(const qi_state* s, const qi_gate* g);
const qi_gate* normalise_gates(const qi_gate* gates, uint64_t count, std::vector<qi_gate>* own);   // gates.cu: aliased-control Matchgates -> P
int launch_simple_gate(qi_state* s, const PhysGate& g);      // one pass with the per-gate kernels
// window.cu
int run_circuit_windowed(qi_state* s, const std::vector<PhysGate>& gates, bool allow_relabel = false);
bool window_supported(const qi_state* s);
int debug_schedule(const qi_state* s, const std::vector<PhysGate>& gates, int R, std::vector<std::vector<int>>* summary);
int debug_lower(const qi_state* s, const std::vector<PhysGate>& gates, int R, std::vector<uint8_t>* blob, std::vector<int>* relabel_out = nullptr);
// shard.cu
int shard_prepare_gate(qi_state* s, const qi_gate* g, PhysGate* out, bool* skip);
int apply_circuit_sharded(qi_state* s, const qi_gate* gates, uint64_t count, bool use_window);   // staged around exchanges
int prepare_gate(qi_state* s, const qi_gate* g, PhysGate* o, bool* skip);                        // gates.cu
int shard_allreduce_sum(qi_state* s, double* host_vals, int count);
int exchange_global_local(qi_state* s, int global_phys, int local_phys);
void logical_uses(const qi_gate& g, uint64_t* n_use, uint64_t* d_use);      // LOGICAL qubits a gate uses non-diagonally / diagonally
// reduce.cu
int reduce_norm_sqr(const qi_state* s, double* out_local);
int reduce_inner(const qi_state* a, const qi_state* b, double out_local[2]);

int grid_for(uint64_t work_items, int block, int max_waves = 8);
// state.cu
int fill_state(qi_state* s, amp_t v);
int set_amplitude(qi_state* s, uint64_t local_index, amp_t v);
// gates.cu: undo lazy SWAP relabelling (physical swaps until logical qubit q sits at bit q)
int canonicalise(qi_state* s);
int shard_localise_mask(qi_state* s, const qi_pauli_term* t);
// staged execution of Pauli-exp sequences around exchanges (shard.cu); lx / lz = logical X-or-Y / Y-or-Z masks per term
int shard_pauli_walk(qi_state* s, const std::vector<uint64_t>& lx, const std::vector<uint64_t>& lz,
                     const std::function<int(const std::vector<size_t>&)>& run, bool dry, uint64_t* exchanges,
                     const std::function<void(const std::vector<int>&, const std::vector<int>&)>& on_exchange = nullptr);
int debug_shard_pauli_stages(uint32_t total_qubits, int world, const std::vector<uint64_t>& lx, const std::vector<uint64_t>& lz,
                             std::vector<uint64_t>* rec);

}  // namespace qi
