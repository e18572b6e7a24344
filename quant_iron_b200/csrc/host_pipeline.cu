// host_pipeline.cu -- Circuit::execute on a HOST-resident state vector (circuit.rs:160-172).
//
// The reference's State is a host Vec<Complex<f64>> (state.rs:74-81): a drop-in caller hands over host memory and
// wants host memory back.  Done naively that is upload, circuit, download, strictly one after the other, and at
// 30 qubits the two 16 GiB PCIe copies cost 60 % of the circuit's own time.  This entry overlaps them with the
// circuit instead.  No new kernel is involved; the gate list is only cut differently:
//
//   * the top k qubits (default 3) index 2^k CHUNKS of the device buffer.  A chunk is exactly what a shard is
//     to the multi-GPU path (shard.cu): a gate without a non-diagonal target on a chunk-index qubit acts inside
//     every chunk on its own (a control on a chunk-index qubit enables/disables the chunk, a diagonal target
//     there is a chunk-dependent phase -- shard_prepare_gate does both);
//   * FRONT = every gate that can be moved ahead of all gates that need a chunk-index qubit non-diagonally
//     (same commutation rule as the window scheduler and staged_walk: two gates commute when on each shared
//     qubit both act diagonally).  The front gates run chunk by chunk, chunk c as soon as its upload has
//     landed and while chunk c+1 is still on the wire;
//   * BACK = every remaining gate that can be moved behind all others, found by the same scan in reverse;
//     it runs chunk by chunk with the download of chunk c behind it, overlapping the work on chunk c+1;
//   * MIDDLE = the rest, run on the whole state by the ordinary executor.
//
// For a brick-work circuit the front and the back are the two light-cone-free trapezoids below the chunk-index
// qubits (about 30 % of the gates each at 30 qubits, depth 40).  The result is the same state the plain sequence
// produces up to the rounding of a different pass grouping (gates are still applied one by one).
// Streams: copies on a dedicated copy stream, kernels on the engine stream, one event per chunk and direction.
#include <algorithm>

#include "common.cuh"

namespace qi {

struct HostPlan {
    std::vector<uint64_t> front, middle, back;      // gate indices, each in circuit order
};

// host only: split the gate list (logical qubits = physical positions, identity layout) for 2^k chunks
static void plan_host_pipeline(uint32_t n, int k, const qi_gate* gates, uint64_t count, HostPlan* plan) {
    const uint64_t top = (k > 0) ? (((1ull << k) - 1ull) << (n - (uint32_t)k)) : 0ull;
    std::vector<uint64_t> rest;
    uint64_t def_any = 0, def_n = 0;
    for (uint64_t i = 0; i < count; i++) {
        uint64_t n_use, d_use;
        logical_uses(gates[i], &n_use, &d_use);
        const bool ok = (n_use & top) == 0 && (n_use & def_any) == 0 && (d_use & def_n) == 0;
        if (ok) plan->front.push_back(i);
        else { rest.push_back(i); def_any |= n_use | d_use; def_n |= n_use; }
    }
    def_any = def_n = 0;
    for (size_t j = rest.size(); j-- > 0;) {
        const uint64_t i = rest[j];
        uint64_t n_use, d_use;
        logical_uses(gates[i], &n_use, &d_use);
        const bool ok = (n_use & top) == 0 && (n_use & def_any) == 0 && (d_use & def_n) == 0;
        if (ok) plan->back.push_back(i);
        else { plan->middle.push_back(i); def_any |= n_use | d_use; def_n |= n_use; }
    }
    std::reverse(plan->back.begin(), plan->back.end());
    std::reverse(plan->middle.begin(), plan->middle.end());
}

// run gates[idx[..]] on `st` (the whole state, or a chunk view that looks like a shard)
static int run_subset(qi_state* st, const qi_gate* gates, const std::vector<uint64_t>& idx) {
    std::vector<PhysGate> run;
    run.reserve(idx.size());
    for (uint64_t i : idx) {
        PhysGate pg;
        bool skip = false;
        QI_TRY(prepare_gate(st, &gates[i], &pg, &skip));
        if (!skip && pg.kind != IK_NOP) run.push_back(pg);
    }
    if (run.empty()) return QI_OK;
    if (ctx().opt_path != 1 && window_supported(st)) return run_circuit_windowed(st, run);
    for (const PhysGate& g : run) QI_TRY(launch_simple_gate(st, g));
    return QI_OK;
}

static int ensure_copy_stream() {
    Context& c = ctx();
    if (c.copy_stream) return QI_OK;
    QI_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    QI_CUDA(cudaEventCreateWithFlags(&c.copy_sync, cudaEventDisableTiming));
    for (int i = 0; i < 16; i++) {
        QI_CUDA(cudaEventCreateWithFlags(&c.chunk_in[i], cudaEventDisableTiming));
        QI_CUDA(cudaEventCreateWithFlags(&c.chunk_out[i], cudaEventDisableTiming));
    }
    return QI_OK;
}

static bool has_lazy_swap(const qi_gate* gates, uint64_t count) {
    if (!ctx().opt_lazy_swap) return false;
    for (uint64_t i = 0; i < count; i++)
        if (gates[i].kind == QI_GATE_SWAP && gates[i].num_controls == 0) return true;
    return false;
}

static int execute_host_pipelined(qi_state* s, const qi_gate* gates, uint64_t count, const double* in, double* out, int k) {
    Context& c = ctx();
    QI_TRY(ensure_copy_stream());
    const uint32_t n = s->num_qubits;
    const int chunks = 1 << k;
    const uint64_t clen = s->len >> k;                       // amplitudes per chunk
    const size_t cbytes = (size_t)clen * sizeof(amp_t);
    HostPlan plan;
    plan_host_pipeline(n, k, gates, count, &plan);

    qi_state view;                                           // chunk c seen as shard c of 2^k (never attached, never exchanged)
    view.num_qubits = n;
    view.n_local = n - (uint32_t)k;
    view.len = clen;
    view.consistent = true;
    view.world = chunks;
    for (int i = 0; i < 64; i++) view.phys[i] = (uint8_t)i;

    // everything queued so far on the engine stream may still read or write the buffer
    QI_CUDA(cudaEventRecord(c.copy_sync, c.stream));
    QI_CUDA(cudaStreamWaitEvent(c.copy_stream, c.copy_sync, 0));
    for (int ch = 0; ch < chunks; ch++) {
        QI_CUDA(cudaMemcpyAsync(s->d + (uint64_t)ch * clen, in + 2ull * (uint64_t)ch * clen, cbytes, cudaMemcpyHostToDevice, c.copy_stream));
        QI_CUDA(cudaEventRecord(c.chunk_in[ch], c.copy_stream));
    }
    for (int i = 0; i < 64; i++) s->phys[i] = (uint8_t)i;   // host data is in logical (identity) order
    for (int ch = 0; ch < chunks; ch++) {
        QI_CUDA(cudaStreamWaitEvent(c.stream, c.chunk_in[ch], 0));
        view.d = s->d + (uint64_t)ch * clen;
        view.rank = ch;
        QI_TRY(run_subset(&view, gates, plan.front));
    }
    QI_TRY(run_subset(s, gates, plan.middle));
    for (int ch = 0; ch < chunks; ch++) {
        view.d = s->d + (uint64_t)ch * clen;
        view.rank = ch;
        QI_TRY(run_subset(&view, gates, plan.back));
        QI_CUDA(cudaEventRecord(c.chunk_out[ch], c.stream));
        QI_CUDA(cudaStreamWaitEvent(c.copy_stream, c.chunk_out[ch], 0));
        QI_CUDA(cudaMemcpyAsync(out + 2ull * (uint64_t)ch * clen, s->d + (uint64_t)ch * clen, cbytes, cudaMemcpyDeviceToHost, c.copy_stream));
    }
    return QI_OK;
}

}  // namespace qi

using namespace qi;

extern "C" {

int qi_execute_host(qi_state* s, const qi_gate* gates, uint64_t count, const double* amps_in, double* amps_out, uint64_t len) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    if (count && !gates) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "gates is NULL");
    if (len != s->len) return fail(QI_ERR_INVALID_ARGUMENT, len, s->len, "length mismatch");
    if (len && (!amps_in || !amps_out)) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "host buffer is NULL");
    for (uint64_t i = 0; i < count; i++) QI_TRY(validate_gate(s, &gates[i]));
    std::vector<qi_gate> own;
    gates = normalise_gates(gates, count, &own);
    QI_TRY(ensure_ctx());
    Context& c = ctx();
    const int k = c.opt_host_chunk_qubits;
    const bool pipelined = k > 0 && s->world == 1 && s->consistent && count > 0 && (int)s->num_qubits >= c.opt_host_min_qubits &&
                           (int)s->num_qubits >= k + 8 && !has_lazy_swap(gates, count);
    if (!pipelined) {
        // same device path, one step after the other (sharded states, small states, lists with relabelled SWAPs)
        QI_TRY(qi_state_upload(s, amps_in, len));
        QI_TRY(qi_apply_circuit(s, gates, count));
        return qi_state_to_host(s, amps_out, len);
    }
    const int st = execute_host_pipelined(s, gates, count, amps_in, amps_out, k);
    // host-visible result: wait for the downloads (and, on failure, for whatever is still using the host buffers)
    cudaError_t e1 = cudaStreamSynchronize(c.copy_stream);
    cudaError_t e2 = cudaStreamSynchronize(c.stream);
    QI_TRY(st);
    if (e1 != cudaSuccess) return cuda_fail(e1, "cudaStreamSynchronize(copy stream)");
    if (e2 != cudaSuccess) return cuda_fail(e2, "cudaStreamSynchronize(engine stream)");
    return QI_OK;
}

// Host-only: the execution order qi_execute_host uses for 2^chunk_qubits chunks -- order[0 .. n_front) run chunk by
// chunk behind the uploads, the next n_middle on the whole state, the last n_back chunk by chunk ahead of the downloads.
int qi_host_pipeline_plan(uint32_t num_qubits, const qi_gate* gates, uint64_t count, int chunk_qubits, uint64_t* order,
                          uint64_t* n_front, uint64_t* n_middle, uint64_t* n_back) {
    if (count && (!gates || !order)) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    if (chunk_qubits < 0 || chunk_qubits > 4 || (uint32_t)chunk_qubits >= num_qubits || num_qubits > 62)
        return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)chunk_qubits, num_qubits, "chunk_qubits out of range");
    qi_state s;
    s.num_qubits = num_qubits;
    s.n_local = num_qubits;
    s.len = 1ull << num_qubits;
    for (int i = 0; i < 64; i++) s.phys[i] = (uint8_t)i;
    for (uint64_t i = 0; i < count; i++) QI_TRY(validate_gate(&s, &gates[i]));
    std::vector<qi_gate> own;
    gates = normalise_gates(gates, count, &own);
    HostPlan plan;
    plan_host_pipeline(num_qubits, chunk_qubits, gates, count, &plan);
    uint64_t o = 0;
    for (uint64_t i : plan.front) order[o++] = i;
    for (uint64_t i : plan.middle) order[o++] = i;
    for (uint64_t i : plan.back) order[o++] = i;
    if (n_front) *n_front = plan.front.size();
    if (n_middle) *n_middle = plan.middle.size();
    if (n_back) *n_back = plan.back.size();
    return QI_OK;
}

}  // extern "C"
