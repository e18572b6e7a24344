/*
 * qiron_b200.h -- C ABI of the B200-native state-vector engine for quant-iron.
 *
 * This is the drop-in boundary: a flat extern "C" library that a host
 * facade (Rust -sys crate, C++ header, Python ctypes) binds in place of quant-iron's rayon/OpenCL
 * backends.  Every entry point cites the reference item it replaces as file:line under the
 * reference root (LordSaumya/quant-iron v2.0.0).
 *
 * Conventions
 *   - Amplitudes are Complex<f64>, interleaved (re, im), 16 bytes each, resident in device memory.
 *     Qubit q is bit q of the amplitude index (LSB = qubit 0), as in operator.rs:347-349.
 *   - States are opaque handles owned by the caller.  Gate entry points mutate IN PLACE; the
 *     reference's `&State -> State` methods are clone + in-place on the facade side.
 *   - Every function returns a qi_status.  Codes 1..15 map 1:1 onto quant-iron's `enum Error`
 *     (errors.rs:3-97); the variant's usize payloads are returned by qi_last_error().
 *   - Validation order equals validate_qubits (operator.rs:214-273) so the same variant fires.
 *   - Calls are asynchronous on the engine's CUDA stream; they synchronise only where a host-visible
 *     value is produced (to_host, scalars, samples).
 *   - There is no CPU fallback: without a CUDA device every compute entry returns QI_ERR_CUDA.
 */
#ifndef QIRON_B200_H
#define QIRON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qi_state qi_state; /* replaces `pub struct State { state_vector: Vec<Complex<f64>>, num_qubits }` state.rs:74-81 */

typedef enum qi_status {
    QI_OK = 0,
    QI_ERR_INVALID_NUMBER_OF_MEASUREMENTS = 1,   /* errors.rs:11  payload[0] = n */
    QI_ERR_OVERLAPPING_CONTROL_AND_TARGET = 2,   /* errors.rs:22  payload = (control, target) */
    QI_ERR_INVALID_NUMBER_OF_QUBITS = 3,         /* errors.rs:31  payload[0] = n */
    QI_ERR_INVALID_QUBIT_INDEX = 4,              /* errors.rs:40  payload = (index, num_qubits) */
    QI_ERR_STATE_VECTOR_NOT_NORMALISED = 5,      /* errors.rs:44 */
    QI_ERR_NON_UNITARY_MATRIX = 6,               /* errors.rs:48 */
    QI_ERR_INVALID_NUMBER_OF_INPUTS = 7,         /* errors.rs:57  payload = (actual, expected) */
    QI_ERR_MISMATCHED_NUMBER_OF_PARAMETERS = 8,  /* errors.rs:66  payload = (expected, actual) */
    QI_ERR_UNKNOWN = 9,                          /* errors.rs:70 */
    QI_ERR_CUDA = 10,                            /* takes the slot of OpenCLError (errors.rs:78); payload[0] = cudaError_t */
    QI_ERR_CONTEXT_LOCK = 11,                    /* GpuContextLockError (errors.rs:82); kept for numbering, never returned */
    QI_ERR_CIRCUIT_MACRO = 12,                   /* errors.rs:86 (host-side only) */
    QI_ERR_INVALID_INPUT_VALUE = 13,             /* errors.rs:90  payload[0] = value */
    QI_ERR_ZERO_NORM = 14,                       /* errors.rs:94 */
    QI_ERR_INVALID_PAULI_STRING_COEFFICIENT = 15,/* errors.rs:98  payload = bit patterns of (re, im) */
    QI_ERR_INVALID_ARGUMENT = 16,                /* NULL handle / malformed record: no reference counterpart */
    QI_ERR_PEER = 17                             /* sharded state: peer mapping / exchange failure */
} qi_status;

/* payload and message of the last non-OK status on the calling thread */
void qi_last_error(uint64_t payload[2], char* msg, size_t msg_len);
const char* qi_version(void);

/* ---- engine ------------------------------------------------------------------------------- */
int qi_init(int device);                       /* bind the engine to a CUDA device (default: current device) */
int qi_synchronize(void);                      /* wait for all queued work */
int qi_device_info(char* name, size_t name_len, int* sm_count, uint64_t* total_mem, uint64_t* free_mem);
/* options: "path" = 0 auto | 1 force the simple per-gate kernels | 2 force the window kernels;
 *          "fuse" = 1/0 fusion in qi_apply_circuit, Pauli-exp sequences and expectation values;
 *          "profile" = 1/0 per-kernel event timing; "window_regs" = 3|4|5 register qubits per window pass
 *          (default 4); "lazy_swap" = 1/0 uncontrolled SWAP as a relabelling; "absorb" = 1/0 fold CNOTs into the
 *          neighbouring single-qubit gate; "tma" = 0/1 TMA-prefetched variant of the window kernel;
 *          "late_tables" = 1/0 unconditional phase tables placed as late as their members allow (fewest per pass);
 *          "lean" = 0/1 window passes apply H / RX / real 2x2 gates in unit form with one deferred scale per pass
 *          (half the FP64 instructions per gate; results differ from the default by rounding only; off until measured);
 *          "prefetch" = 0/1 (with "lean" = 1 only) L2 prefetch of every warp's next tile (off until measured);
 *          "host_chunk_qubits", "host_min_qubits": see qi_execute_host;
 *          "tile" = 1/0 fused passes on the CTA-tile executor (11 qubits per HBM pass) for states of >= "tile_min_qubits";
 *          "jit" = tile passes as circuit-specialised straight-line sm_100a modules, assembled from generated PTX by the
 *          driver (csrc/tile_jit.cuh): 0 never | 1 (default) in the background once a pass structure has been seen, for
 *          states of >= "jit_min_qubits" local qubits (the interpreting kernel runs the pass meanwhile; results are
 *          bit-identical) | 2 before the first launch; module skeleton variants kept for A/B (all measured slower or equal,
 *          DESIGN 3.0): "jit_ctas" = 4|3|5|6 CTAs per SM, "jit_groups" = 1|2|4 tiles per CTA, "jit_prefetch" = 0|1|2 L2
 *          prefetch of the next tile, "jit_stage" = 0|1 next tile staged in shared memory by bulk async copies,
 *          "jit_smem_kb" = minimum dynamic shared memory of a module (caps occupancy);
 *          tile lowering: "tile_lean" = 1/0 uncontrolled H / RY / RX in unit form, "tile_pform" = 0..4 diagonal groups of at
 *          most that many qubits as P-form phase ops instead of tables, "tile_carry" = 1/0 one scale op per run,
 *          "tile_slide" = 1/0 sliding tiles (interpreting kernel only), "tile_restore" = 0/1 sliding tiles + layout-restoring
 *          relabel passes for states that run on modules, "tile_min_gates" = 3 shorter runs stay on the window kernel;
 *          "peer_timeout_s" = seconds a sharded state waits for a peer at a device barrier before the kernel traps */
int qi_set_option(const char* name, int64_t value);
/* wait until every queued tile module is assembled (benchmarks: call after the first execution of a circuit) */
int qi_jit_drain(void);
/* modules assembled / rejected by the driver / still queued, the host milliseconds spent assembling, and the FP64 warp
 * instructions launched on modules since the last qi_stats_reset (static count per pass, instructions under a control
 * weighted by the fraction of threads the control selects) -- the numerator of the FP64-pipe roofline in bench.py */
int qi_jit_stats(uint64_t* modules, uint64_t* failed, uint64_t* pending, double* assemble_ms, double* fp64_warp_instr);

/* kernel accounting for bench.py (gpu_launches, roofline.achieved) */
typedef struct qi_kernel_stat {
    char name[32];
    uint64_t launches;
    double total_ms;          /* only when the "profile" option is on */
    double algorithmic_bytes; /* sum over launches of the bytes the pass must move */
} qi_kernel_stat;
int qi_stats_reset(void);
int qi_stats_get(qi_kernel_stat* out, int capacity, int* count);
/* CUDA-event stopwatch on the engine stream */
int qi_timer_start(void);
int qi_timer_stop(float* elapsed_ms);

/* ---- state container (state.rs:99-373, 397-453, 801-945, 2687-2862) --------------------------- */
int qi_state_new_zero(uint32_t num_qubits, qi_state** out);                   /* state.rs:165-178 */
int qi_state_new_basis_n(uint32_t num_qubits, uint64_t n, qi_state** out);    /* state.rs:194-210 */
int qi_state_new_plus(uint32_t num_qubits, qi_state** out);                   /* state.rs:225-237 */
int qi_state_new_minus(uint32_t num_qubits, qi_state** out);                  /* state.rs:252-290 */
int qi_state_new_ghz(uint32_t num_qubits, qi_state** out);                    /* state.rs:305-325 */
/* State::new (checked, state.rs:99-127) when check != 0; the struct literal State{..} when check == 0
 * (len need not be 2^num_qubits then, as in state_tests.rs:145-150). amps = len interleaved (re,im). */
int qi_state_from_host(const double* amps, uint64_t len, uint32_t num_qubits, int check, qi_state** out);
int qi_state_to_host(const qi_state* s, double* amps, uint64_t len);          /* read `state_vector` */
int qi_state_upload(qi_state* s, const double* amps, uint64_t len);           /* overwrite `state_vector` (this rank's shard) from host */
int qi_state_clone(const qi_state* s, qi_state** out);                        /* #[derive(Clone)] state.rs:69 */
void qi_state_free(qi_state* s);
uint32_t qi_state_num_qubits(const qi_state* s);                              /* state.rs:431 */
uint64_t qi_state_len(const qi_state* s);
int qi_state_amplitude(const qi_state* s, uint64_t n, double out[2]);         /* state.rs:448-453 */
int qi_state_init_random(qi_state* s, uint64_t seed);  /* synthetic normalised state */
void* qi_state_device_ptr(qi_state* s);                /* raw device pointer (interop: torch / NCCL plumbing) */

int qi_inner_product(const qi_state* a, const qi_state* b, double out[2]);    /* state.rs:890-917 */
int qi_norm_sqr(const qi_state* s, double* out);                              /* state.rs:117, 924-929 */
int qi_normalise(qi_state* s);                                                /* state.rs:924-945 */
int qi_scale(qi_state* s, const double z[2]);                                 /* state.rs:2687-2776 */
int qi_add(qi_state* a, const qi_state* b);                                   /* state.rs:2779-2800 */
int qi_sub(qi_state* a, const qi_state* b);                                   /* state.rs:2841-2862 */
int qi_conj(qi_state* s);                                                     /* state.rs:397-403 */
int qi_tensor_product(const qi_state* a, const qi_state* b, qi_state** out);  /* state.rs:801-836 */

/* ---- operators (operator.rs) -------------------------------------------------------------- */
typedef enum qi_gate_kind {
    QI_GATE_H = 1,         /* Hadamard::apply        operator.rs:303-424 */
    QI_GATE_X = 2,         /* Pauli::X               operator.rs:474-606 */
    QI_GATE_Y = 3,         /* Pauli::Y */
    QI_GATE_Z = 4,         /* Pauli::Z */
    QI_GATE_I = 5,         /* Identity::apply        operator.rs:1112-1123 */
    QI_GATE_S = 6,         /* PhaseS                 operator.rs:1160-1216 */
    QI_GATE_SDG = 7,       /* PhaseSdag              operator.rs:1352-1408 */
    QI_GATE_T = 8,         /* PhaseT                 operator.rs:1253-1315 */
    QI_GATE_TDG = 9,       /* PhaseTdag              operator.rs:1445-1507 */
    QI_GATE_P = 10,        /* PhaseShift             operator.rs:1565-1624   params[0] = angle */
    QI_GATE_RX = 11,       /* RotateX                operator.rs:1674-1767   params[0] = angle */
    QI_GATE_RY = 12,       /* RotateY                operator.rs:1817-1908 */
    QI_GATE_RZ = 13,       /* RotateZ                operator.rs:1958-2035 */
    QI_GATE_U2 = 14,       /* Unitary2::apply        operator.rs:2209-2266   params = m00,m01,m10,m11 as (re,im) */
    QI_GATE_CNOT = 15,     /* CNOT::apply            operator.rs:667-685     exactly one control */
    QI_GATE_SWAP = 16,     /* SWAP::apply            operator.rs:731-820     two targets */
    QI_GATE_TOFFOLI = 17,  /* Toffoli::apply         operator.rs:1055-1075   two distinct controls */
    QI_GATE_MATCHGATE = 18 /* Matchgate::apply       operator.rs:893-1014    params = theta, phi1, phi2 */
} qi_gate_kind;

/* One `Gate::Operator(Box<dyn Operator>, targets, controls)` record (gate.rs:23). */
typedef struct qi_gate {
    int32_t kind;              /* qi_gate_kind */
    uint32_t num_targets;
    uint32_t targets[2];
    uint32_t num_controls;
    const uint32_t* controls;  /* may be NULL when num_controls == 0 */
    double params[8];
} qi_gate;

/* Operator::apply (operator.rs:165-170), in place. */
int qi_apply_gate(qi_state* s, const qi_gate* gate);
/* Circuit::execute's gate loop (circuit.rs:160-172) over a run of operator gates.  All records are
 * validated first; then the run is scheduled into fused register-window passes. */
int qi_apply_circuit(qi_state* s, const qi_gate* gates, uint64_t count);
/* Circuit::execute (circuit.rs:160-172) on a HOST-resident state vector, the form the reference's caller has
 * (`State.state_vector` is a host Vec, state.rs:74-81): same result as qi_state_upload + qi_apply_circuit +
 * qi_state_to_host on `s` (the device working buffer, len = 2^num_qubits amplitudes; it holds the final state
 * afterwards), but the two PCIe copies overlap the circuit: the device buffer is cut into 2^k chunks by its top k
 * qubits, the gates that commute ahead of everything touching those qubits non-diagonally run chunk by chunk behind
 * the uploads, the gates that commute behind everything else run chunk by chunk ahead of the downloads
 * (csrc/host_pipeline.cu).  amps_in and amps_out may be the same buffer; pinned memory is needed for the overlap,
 * not for correctness.  Options: "host_chunk_qubits" = k (default 3, 0 = plain sequence), "host_min_qubits" (default
 * 26: smaller states take the plain sequence).  Sharded states and lists with relabelled SWAPs take the plain sequence. */
int qi_execute_host(qi_state* s, const qi_gate* gates, uint64_t count, const double* amps_in, double* amps_out, uint64_t len);
/* host-only: the execution order qi_execute_host uses -- order[0 .. n_front) chunk by chunk behind the uploads,
 * the next n_middle on the whole state, the last n_back chunk by chunk ahead of the downloads (no device access) */
int qi_host_pipeline_plan(uint32_t num_qubits, const qi_gate* gates, uint64_t count, int chunk_qubits, uint64_t* order,
                          uint64_t* n_front, uint64_t* n_middle, uint64_t* n_back);
/* Unitary2::new's unitarity check (operator.rs:2092-2118): QI_OK or QI_ERR_NON_UNITARY_MATRIX */
int qi_unitary2_check(const double m[8]);

/* ---- PauliString / SumOp (pauli_string.rs) -------------------------------------------------- */
typedef struct qi_pauli_term {
    uint32_t num_ops;
    const uint32_t* qubits;    /* distinct qubits */
    const uint8_t* paulis;     /* 1 = X, 2 = Y, 3 = Z */
    double coefficient[2];
} qi_pauli_term;

/* PauliString::apply (pauli_string.rs:139-151) when with_coefficient != 0, apply_operators (172-184) otherwise */
int qi_apply_pauli_string(qi_state* s, const qi_pauli_term* term, int with_coefficient);
/* PauliString::apply_exp_factor (pauli_string.rs:237-262): psi <- cosh(a) psi + sinh(a) P psi, a = coefficient*factor.
 * apply_exp (198-223) is factor = 1; apply_exp_neg_i_dt (281-287) is factor = (0,-dt) after the Im(coeff)==0 check. */
int qi_apply_pauli_exp(qi_state* s, const qi_pauli_term* term, const double factor[2]);
/* SumOp::expectation_value (pauli_string.rs:485-507) */
int qi_expect_pauli_sum(const qi_state* s, const qi_pauli_term* terms, uint64_t count, double out[2]);
/* SumOp::apply (pauli_string.rs:453-466): out = sum_k P_k psi (new state) */
int qi_apply_pauli_sum(const qi_state* s, const qi_pauli_term* terms, uint64_t count, qi_state** out);
/* a run of apply_exp_factor calls (pauli_string.rs:237-262) applied in order, term k with factors[2k], factors[2k+1]:
 * what the Trotter loops (time_evolution.rs:57-63, 102-111) and a run of Gate::PauliTimeEvolution gates
 * (gate.rs:116-118) do.  Consecutive terms share register-window passes. */
int qi_apply_pauli_exp_sequence(qi_state* s, const qi_pauli_term* terms, uint64_t count, const double* factors);
/* trotter_evolve_state (time_evolution.rs:140-167); order 1 = First (45-66), 2 = Second (89-115); the whole
 * evolution is one such sequence */
int qi_trotter_evolve(qi_state* s, const qi_pauli_term* terms, uint64_t count, double dt, uint64_t steps, int order);
/* host-only: passes the batching scheduler builds for `repeats` repetitions of the term list on a
 * `num_qubits` state (terms per pass; 0 = a term that runs alone); no device access */
int qi_debug_pauli_schedule(uint32_t num_qubits, const qi_pauli_term* terms, uint64_t count, uint64_t repeats,
                            int32_t* terms_per_pass, uint64_t max_passes, uint64_t* n_passes);

/* ---- measurement (state.rs:525-784) --------------------------------------------------------- */
typedef enum qi_basis { QI_BASIS_COMPUTATIONAL = 0, QI_BASIS_X = 1, QI_BASIS_Y = 2, QI_BASIS_CUSTOM = 3 } qi_basis;

/* un-normalised marginal table over `m` qubits, bin bit j <-> qubits[j] (state.rs:559-588); out = 2^m host doubles */
int qi_probabilities(const qi_state* s, const uint32_t* qubits, uint32_t m, double* out);
/* `shots` draws from one table by a device prefix scan; draw k uses u_k of the shared-seed stream
 * (splitmix64, u = (x >> 11) * 2^-53); rule: first bin with u < cumsum, else last (state.rs:601-619) */
int qi_sample(const qi_state* s, const uint32_t* qubits, uint32_t m, uint64_t shots, uint64_t seed, uint64_t* bins);
/* collapse onto `bin` and renormalise (state.rs:622-654), in place */
int qi_collapse(qi_state* s, const uint32_t* qubits, uint32_t m, uint64_t bin);
/* State::measure (state.rs:525-730), in place: basis change, draw u_{draw_index} of `seed`, collapse, basis
 * change back.  m = 0 measures every qubit (531-541); outcomes[j] = (bin >> j) & 1 (657-660).
 * custom_u = 8 doubles (row-major 2x2, (re,im)) for QI_BASIS_CUSTOM, else NULL. */
int qi_measure(qi_state* s, int basis, const double* custom_u, const uint32_t* qubits, uint32_t m,
               uint64_t seed, uint64_t draw_index, uint8_t* outcomes, uint64_t* bin);
/* the u of the shared-seed stream, for host-side checks */
double qi_uniform(uint64_t seed, uint64_t k);

/* ---- sharded state over 2/4/8 GPUs (new work) --------------------------------- */
#define QI_IPC_HANDLE_BYTES 64
/* one shard per process: `total_qubits` logical qubits, the top log2(world) are global */
int qi_shard_new_zero(uint32_t total_qubits, int rank, int world, qi_state** out);
int qi_shard_new_plus(uint32_t total_qubits, int rank, int world, qi_state** out);
int qi_shard_new_basis_n(uint32_t total_qubits, uint64_t n, int rank, int world, qi_state** out);
/* export this shard's amplitude buffer + its flag block as CUDA IPC handles (2 * QI_IPC_HANDLE_BYTES) */
int qi_shard_export(qi_state* s, uint8_t* handles);
/* map every peer's buffers; `all_handles` = world * 2 * QI_IPC_HANDLE_BYTES, gathered by the host (torch.distributed) */
int qi_shard_attach(qi_state* s, const uint8_t* all_handles);
/* this rank's shard as stored: PHYSICAL bit order (qi_state_layout gives the logical -> physical map), len = 2^n_local
 * amplitudes.  (qi_state_to_host refuses sharded states: a shard is not the reference's `state_vector`.) */
int qi_shard_to_host(const qi_state* s, double* amps, uint64_t len);
int qi_shard_rank(const qi_state* s);
int qi_shard_world(const qi_state* s);
/* bytes this rank moved over NVLink and the number of exchanges since creation */
int qi_shard_comm_stats(const qi_state* s, uint64_t* bytes_sent, uint64_t* bytes_received, uint64_t* exchanges);
/* logical qubit -> physical bit position (64 entries) and the number of local index bits: uncontrolled SWAP
 * gates and global<->local exchanges only relabel this map.  qi_state_to_host / qi_state_amplitude undo it. */
int qi_state_layout(const qi_state* s, uint8_t* phys, uint32_t* n_local);
/* host-only planner (no device): number of global<->local exchanges the engine performs for a gate list on
 * `world` ranks, the number of gates touching a global qubit that need NO communication, and the final map */
int qi_shard_plan(uint32_t total_qubits, int world, const qi_gate* gates, uint64_t count, uint64_t* exchanges,
                  uint64_t* comm_free_global_gates, uint8_t* final_phys);

/* host-only planner for Pauli-exp sequences (`repeats` repetitions of the term list, e.g. Trotter steps) on `world`
 * ranks: exchanges the engine performs and the number of stages it runs between them */
int qi_shard_plan_pauli(uint32_t total_qubits, int world, const qi_pauli_term* terms, uint64_t count, uint64_t repeats,
                        uint64_t* exchanges, uint64_t* stages);

/* host-only: how the fused executor splits a gate list into passes on one device; rows[8*i..] =
 * {per-gate-kernel step?, window qubits used, lane-pair ops, register-pair ops, diagonal ops, phase-table ops,
 *  CNOTs absorbed into a neighbouring gate, 0} */
int qi_debug_schedule(uint32_t num_qubits, const qi_gate* gates, uint64_t count, int window_regs, int32_t* rows,
                      uint64_t max_rows, uint64_t* n_rows);

/* host-only: the device programs (window passes with their op lists and phase tables, per-gate-kernel steps) the fused
 * executor would launch for a gate list, serialised so that tests can interpret them on the CPU and compare with the
 * oracle without a GPU (tests/test_window_lowering.py; blob layout: csrc/window.cu, debug_lower).  rank / world > 1:
 * the programs of that shard (or host-pipeline chunk); phys = the logical -> physical qubit map (64 entries, NULL =
 * identity).  *used = bytes written / needed. */
int qi_debug_lower(uint32_t num_qubits, int rank, int world, const uint8_t* phys, const qi_gate* gates, uint64_t count,
                   int window_regs, uint8_t* blob, uint64_t capacity, uint64_t* used);
/* host-only: the stages the sharded executor runs for a gate list on `world` ranks (the gates of each stage, the qubit map
 * it runs under, the global<->local exchange that follows), as a u64 stream a test replays on the CPU together with
 * qi_debug_lower (tests/test_sharded_emulation.py; record layout: csrc/shard.cu) */
int qi_debug_shard_stages(uint32_t total_qubits, int world, const qi_gate* gates, uint64_t count, uint64_t* out,
                          uint64_t capacity, uint64_t* used);

/* host-only: the same for a sequence of apply_exp_factor calls (term k with factors[2k], factors[2k+1]): the fused
 * Pauli-exp window passes and the terms that run alone (blob layout: csrc/pauli_window.cu, debug_pauli_lower); rank /
 * world / phys as in qi_debug_lower */
int qi_debug_pauli_lower(uint32_t num_qubits, int rank, int world, const uint8_t* phys, const qi_pauli_term* terms,
                         uint64_t count, const double* factors, uint8_t* blob, uint64_t capacity, uint64_t* used);
/* host-only: the stages a Pauli-exp sequence runs in on `world` ranks (same record layout as qi_debug_shard_stages, term
 * indices instead of gate indices) */
int qi_debug_shard_pauli_stages(uint32_t total_qubits, int world, const qi_pauli_term* terms, uint64_t count, uint64_t* out,
                                uint64_t capacity, uint64_t* used);

/* host-only: the read-only window programs of a batched SumOp expectation value (groups of terms sharing one read of the
 * state) and the indices of the terms left to the per-term kernel (blob layout: csrc/pauli_window.cu, debug_expect_lower) */
int qi_debug_expect_lower(uint32_t num_qubits, const qi_pauli_term* terms, uint64_t count, uint8_t* blob, uint64_t capacity,
                          uint64_t* used);

#ifdef __cplusplus
}
#endif
#endif /* QIRON_B200_H */
