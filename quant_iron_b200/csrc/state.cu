// state.cu -- device-resident State container: constructors, host transfer, element-wise
// arithmetic and the deterministic reductions (norm, inner product).
// Reference: src/components/state.rs (cited per entry point in include/qiron_b200.h).
#include "common.cuh"

namespace qi {

static const int kBlock = 256;
static const int kReduceBlocksPerSM = 4;

// ---- fills ------------------------------------------------------------------------------------
__global__ void k_fill(amp_t* a, uint64_t len, amp_t v) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) a[i] = v;
}

// new_minus (state.rs:252-290): sign = parity of popcount of the FULL index (rank bits included)
__global__ void k_fill_minus(amp_t* a, uint64_t len, double v, uint64_t high_bits) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    int hp = __popcll(high_bits) & 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        int par = (__popcll(i) & 1) ^ hp;
        a[i] = make_double2(par ? -v : v, 0.0);
    }
}

__global__ void k_set_one(amp_t* a, uint64_t idx, amp_t v) { a[idx] = v; }

__global__ void k_random_state(amp_t* a, uint64_t len, uint64_t seed, uint64_t index_base) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        uint64_t g = index_base + i;
        double u1 = uniform_at(seed, 2 * g), u2 = uniform_at(seed, 2 * g + 1);
        double r = sqrt(-2.0 * log(u1 + 1.1102230246251565e-16));
        double s, c;
        sincos(6.283185307179586 * u2, &s, &c);
        a[i] = make_double2(r * c, r * s);
    }
}

// ---- element-wise -----------------------------------------------------------------------------
__global__ void k_scale(amp_t* a, uint64_t len, amp_t z) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) a[i] = cmul(a[i], z);
}
__global__ void k_div_real(amp_t* a, uint64_t len, double d) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        amp_t v = a[i];
        a[i] = make_double2(v.x / d, v.y / d);
    }
}
template <int SIGN>
__global__ void k_addsub(amp_t* a, const amp_t* b, uint64_t len) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        amp_t x = a[i], y = b[i];
        a[i] = SIGN > 0 ? cadd(x, y) : csub(x, y);
    }
}
__global__ void k_conj(amp_t* a, uint64_t len) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) a[i].y = -a[i].y;
}
__global__ void k_tensor(amp_t* out, const amp_t* a, const amp_t* b, uint64_t len, int nb) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t mb = (1ull << nb) - 1ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
        out[i] = cmul(a[i >> nb], b[i & mb]);
}

// ---- deterministic reductions -----------------------------------------------------------------
// Stage 1: fixed grid, each thread a fixed strided subsequence, warp-shuffle tree, shared-memory
// tree across warps -> one partial per block.  Stage 2: one block sums the partials in a fixed
// tree.  Same launch configuration => bit-identical result run to run.
__device__ __forceinline__ double2 block_sum(double2 v) {
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
    if (w == 0) {
        for (int o = 16; o > 0; o >>= 1) {
            v.x += __shfl_down_sync(0xffffffffu, v.x, o);
            v.y += __shfl_down_sync(0xffffffffu, v.y, o);
        }
    }
    return v;  // valid in thread 0
}

__global__ void k_norm_partial(const amp_t* a, uint64_t len, double2* partials) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        amp_t v = a[i];
        acc += v.x * v.x + v.y * v.y;
    }
    double2 r = block_sum(make_double2(acc, 0.0));
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

__global__ void k_inner_partial(const amp_t* a, const amp_t* b, uint64_t len, double2* partials) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double2 acc = make_double2(0.0, 0.0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        amp_t t = cmul(cconj(a[i]), b[i]);   // a.conj() * b, state.rs:906
        acc.x += t.x;
        acc.y += t.y;
    }
    double2 r = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// sums `count` partials per slot; slot s reads partials[s*count .. (s+1)*count)
__global__ void k_final_sum(const double2* partials, int count, double2* out) {
    const double2* p = partials + (size_t)blockIdx.x * count;
    double2 acc = make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < count; i += blockDim.x) { acc.x += p[i].x; acc.y += p[i].y; }
    double2 r = block_sum(acc);
    if (threadIdx.x == 0) out[blockIdx.x] = r;
}

static int reduce_grid(uint64_t len) {
    int g = ctx().sm_count * kReduceBlocksPerSM;
    uint64_t need = (len + kBlock - 1) / kBlock;
    if (need < (uint64_t)g) g = (int)(need ? need : 1);
    return g;
}

static int fetch_result(int slots, double* out) {
    Context& c = ctx();
    QI_CUDA(cudaMemcpyAsync(c.h_result, c.d_result, sizeof(double) * 2 * slots, cudaMemcpyDeviceToHost, c.stream));
    QI_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < 2 * slots; i++) out[i] = c.h_result[i];
    return QI_OK;
}

int reduce_norm_sqr(const qi_state* s, double* out_local) {
    Context& c = ctx();
    if (s->len == 0) { *out_local = 0.0; return QI_OK; }
    int g = reduce_grid(s->len);
    QI_TRY(ensure_partials((size_t)g * 2));
    {
        LaunchScope ls(KF_REDUCE, 16.0 * (double)s->len);
        k_norm_partial<<<g, kBlock, 0, c.stream>>>(s->d, s->len, (double2*)c.d_partials);
    }
    k_final_sum<<<1, kBlock, 0, c.stream>>>((double2*)c.d_partials, g, (double2*)c.d_result);
    QI_TRY(check_launch("norm_sqr"));
    double r[2];
    QI_TRY(fetch_result(1, r));
    *out_local = r[0];
    return QI_OK;
}

int reduce_inner(const qi_state* a, const qi_state* b, double out_local[2]) {
    Context& c = ctx();
    if (a->len == 0) { out_local[0] = out_local[1] = 0.0; return QI_OK; }
    int g = reduce_grid(a->len);
    QI_TRY(ensure_partials((size_t)g * 2));
    {
        LaunchScope ls(KF_REDUCE, 32.0 * (double)a->len);
        k_inner_partial<<<g, kBlock, 0, c.stream>>>(a->d, b->d, a->len, (double2*)c.d_partials);
    }
    k_final_sum<<<1, kBlock, 0, c.stream>>>((double2*)c.d_partials, g, (double2*)c.d_result);
    QI_TRY(check_launch("inner_product"));
    return fetch_result(1, out_local);
}

// ---- allocation ---------------------------------------------------------------------------------
static int alloc_state(uint32_t num_qubits, uint64_t len, qi_state** out) {
    QI_TRY(ensure_ctx());
    qi_state* s = new qi_state();
    s->len = len;
    s->num_qubits = num_qubits;
    s->n_local = num_qubits;
    s->consistent = (num_qubits < 64) && (len == (1ull << num_qubits));
    for (int i = 0; i < 64; i++) s->phys[i] = (uint8_t)i;
    if (len) {
        int st = dev_alloc((void**)&s->d, len * sizeof(amp_t));
        if (st != QI_OK) { delete s; return st; }
    }
    *out = s;
    return QI_OK;
}

int set_amplitude(qi_state* s, uint64_t local_index, amp_t v) {
    k_set_one<<<1, 1, 0, ctx().stream>>>(s->d, local_index, v);
    return check_launch("set_amplitude");
}

int fill_state(qi_state* s, amp_t v) {
    if (!s->len) return QI_OK;
    LaunchScope ls(KF_INIT, 16.0 * (double)s->len);
    k_fill<<<grid_for(s->len, kBlock), kBlock, 0, ctx().stream>>>(s->d, s->len, v);
    return check_launch("fill");
}

}  // namespace qi

using namespace qi;

extern "C" {

static int check_nq(uint32_t n) {
    if (n == 0) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, 0, 0, "Invalid number of qubits: 0");
    if (n > 40) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, n, 0, "state does not fit device memory");
    return QI_OK;
}

int qi_state_new_zero(uint32_t n, qi_state** out) {
    if (!out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "out is NULL");
    QI_TRY(check_nq(n));
    QI_TRY(alloc_state(n, 1ull << n, out));
    QI_TRY(fill_state(*out, make_double2(0.0, 0.0)));
    k_set_one<<<1, 1, 0, ctx().stream>>>((*out)->d, 0, make_double2(1.0, 0.0));
    return check_launch("new_zero");
}

int qi_state_new_basis_n(uint32_t n, uint64_t k, qi_state** out) {
    if (!out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "out is NULL");
    // state.rs:196-202: the index check comes before the qubit-count check
    if (n >= 64 || k >= (1ull << n)) return fail(QI_ERR_INVALID_QUBIT_INDEX, k, n, "basis index out of range");
    QI_TRY(check_nq(n));
    QI_TRY(alloc_state(n, 1ull << n, out));
    QI_TRY(fill_state(*out, make_double2(0.0, 0.0)));
    k_set_one<<<1, 1, 0, ctx().stream>>>((*out)->d, k, make_double2(1.0, 0.0));
    return check_launch("new_basis_n");
}

int qi_state_new_plus(uint32_t n, qi_state** out) {
    if (!out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "out is NULL");
    QI_TRY(check_nq(n));
    QI_TRY(alloc_state(n, 1ull << n, out));
    double dim = (double)(1ull << n);
    return fill_state(*out, make_double2(1.0 / std::sqrt(dim), 0.0));   // state.rs:230
}

int qi_state_new_minus(uint32_t n, qi_state** out) {
    if (!out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "out is NULL");
    QI_TRY(check_nq(n));
    QI_TRY(alloc_state(n, 1ull << n, out));
    double dim = (double)(1ull << n);
    qi_state* s = *out;
    LaunchScope ls(KF_INIT, 16.0 * (double)s->len);
    k_fill_minus<<<grid_for(s->len, 256), 256, 0, ctx().stream>>>(s->d, s->len, 1.0 / std::sqrt(dim), 0);
    return check_launch("new_minus");
}

int qi_state_new_ghz(uint32_t n, qi_state** out) {
    if (!out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "out is NULL");
    QI_TRY(check_nq(n));
    QI_TRY(alloc_state(n, 1ull << n, out));
    QI_TRY(fill_state(*out, make_double2(0.0, 0.0)));
    const double a = 0.70710678118654752440;  // FRAC_1_SQRT_2, state.rs:310
    k_set_one<<<1, 1, 0, ctx().stream>>>((*out)->d, 0, make_double2(a, 0.0));
    k_set_one<<<1, 1, 0, ctx().stream>>>((*out)->d, (1ull << n) - 1, make_double2(a, 0.0));
    return check_launch("new_ghz");
}

int qi_state_from_host(const double* amps, uint64_t len, uint32_t num_qubits, int check, qi_state** out) {
    if (!out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "out is NULL");
    if (check) {
        // State::new, state.rs:99-127
        if (len == 0) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, 0, 0, "empty state vector");
        if (len & (len - 1)) {
            uint64_t fl = 0;
            while ((2ull << fl) <= len) fl++;
            return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, fl, 0, "length is not a power of two");
        }
        num_qubits = 0;
        while ((1ull << num_qubits) < len) num_qubits++;
    }
    if (len && !amps) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "amps is NULL");
    QI_TRY(alloc_state(num_qubits, len, out));
    qi_state* s = *out;
    if (len) {
        cudaError_t e = cudaMemcpyAsync(s->d, amps, len * sizeof(amp_t), cudaMemcpyHostToDevice, ctx().stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);
        if (e != cudaSuccess) { qi_state_free(s); *out = nullptr; return cuda_fail(e, "H2D copy"); }
    }
    if (check) {
        double nsq = 0.0;
        int st = reduce_norm_sqr(s, &nsq);
        if (st != QI_OK) { qi_state_free(s); *out = nullptr; return st; }
        double tol = 2.220446049250313e-16 * (double)len;     // state.rs:118
        if (std::fabs(nsq - 1.0) > tol) {
            qi_state_free(s);
            *out = nullptr;
            return fail(QI_ERR_STATE_VECTOR_NOT_NORMALISED, 0, 0, "State vector is not normalised");
        }
    }
    return QI_OK;
}

int qi_shard_to_host(const qi_state* s, double* amps, uint64_t len) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    if (len != s->len) return fail(QI_ERR_INVALID_ARGUMENT, len, s->len, "length mismatch");
    QI_TRY(ensure_ctx());
    if (len) QI_CUDA(cudaMemcpyAsync(amps, s->d, len * sizeof(amp_t), cudaMemcpyDeviceToHost, ctx().stream));
    QI_CUDA(cudaStreamSynchronize(ctx().stream));
    return QI_OK;
}

int qi_state_to_host(const qi_state* s, double* amps, uint64_t len) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    if (len != s->len) return fail(QI_ERR_INVALID_ARGUMENT, len, s->len, "length mismatch");
    if (s->world > 1)      // a shard is stored in physical bit order and holds 1 / world of the vector: never hand it out as `state_vector`
        return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)s->world, 0, "sharded state: read shards with qi_shard_to_host (physical order, see qi_state_layout)");
    QI_TRY(ensure_ctx());
    if (!s->identity_layout()) QI_TRY(canonicalise(const_cast<qi_state*>(s)));   // lazy SWAPs / tile relabelling become real
    if (len) {
        QI_CUDA(cudaMemcpyAsync(amps, s->d, len * sizeof(amp_t), cudaMemcpyDeviceToHost, ctx().stream));
    }
    QI_CUDA(cudaStreamSynchronize(ctx().stream));
    return QI_OK;
}

int qi_state_upload(qi_state* s, const double* amps, uint64_t len) {
    if (!s || (len && !amps)) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    if (len != s->len) return fail(QI_ERR_INVALID_ARGUMENT, len, s->len, "length mismatch");
    QI_TRY(ensure_ctx());
    if (len) QI_CUDA(cudaMemcpyAsync(s->d, amps, len * sizeof(amp_t), cudaMemcpyHostToDevice, ctx().stream));
    QI_CUDA(cudaStreamSynchronize(ctx().stream));
    for (int i = 0; i < 64; i++) s->phys[i] = (uint8_t)i;     // host data is in logical (identity) order
    return QI_OK;
}

int qi_state_clone(const qi_state* s, qi_state** out) {
    if (!s || !out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    if (s->world > 1) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "sharded states cannot be cloned (no room for a second shard)");
    QI_TRY(alloc_state(s->num_qubits, s->len, out));
    (*out)->consistent = s->consistent;
    memcpy((*out)->phys, s->phys, sizeof(s->phys));
    if (s->len) {
        LaunchScope ls(KF_ELEMENTWISE, 32.0 * (double)s->len);
        QI_CUDA(cudaMemcpyAsync((*out)->d, s->d, s->len * sizeof(amp_t), cudaMemcpyDeviceToDevice, ctx().stream));
    }
    return QI_OK;
}

void qi_shard_release(qi_state* s);

void qi_state_free(qi_state* s) {
    if (!s) return;
    if (s->world > 1) {                  // shards are IPC-exported: never cached
        if (ctx().ready) cudaStreamSynchronize(ctx().stream);
        qi_shard_release(s);
        if (s->d) cudaFree(s->d);
    } else {
        dev_free(s->d, s->len * sizeof(amp_t));
    }
    delete s;
}

uint32_t qi_state_num_qubits(const qi_state* s) { return s ? s->num_qubits : 0; }
uint64_t qi_state_len(const qi_state* s) { return s ? s->len : 0; }
void* qi_state_device_ptr(qi_state* s) { return s ? (void*)s->d : nullptr; }

int qi_state_amplitude(const qi_state* s, uint64_t n, double out[2]) {
    if (!s || !out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    if (n >= ((uint64_t)s->len << (s->num_qubits - s->n_local)) || (s->world == 1 && n >= s->len))
        return fail(QI_ERR_INVALID_QUBIT_INDEX, n, s->num_qubits, "amplitude index out of range");
    QI_TRY(ensure_ctx());
    uint64_t pidx = n;
    if (s->consistent) {     // logical index -> physical index (lazy SWAPs, global<->local exchanges)
        pidx = 0;
        for (uint32_t q = 0; q < s->num_qubits; q++) pidx |= ((n >> q) & 1ull) << s->phys[q];
    }
    ctx().h_result[0] = ctx().h_result[1] = 0.0;
    if ((pidx >> s->n_local) == (uint64_t)s->rank || s->world == 1)
        QI_CUDA(cudaMemcpyAsync(ctx().h_result, s->d + (pidx & (s->len - 1)), sizeof(amp_t), cudaMemcpyDeviceToHost, ctx().stream));
    QI_CUDA(cudaStreamSynchronize(ctx().stream));
    out[0] = ctx().h_result[0];
    out[1] = ctx().h_result[1];
    if (s->world > 1) QI_TRY(shard_allreduce_sum(const_cast<qi_state*>(s), out, 2));   // only the owner rank holds it
    return QI_OK;
}

int qi_state_init_random(qi_state* s, uint64_t seed) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    QI_TRY(ensure_ctx());
    uint64_t base = (uint64_t)s->rank << s->n_local;
    {
        LaunchScope ls(KF_INIT, 16.0 * (double)s->len);
        k_random_state<<<grid_for(s->len, kBlock), kBlock, 0, ctx().stream>>>(s->d, s->len, seed, base);
    }
    QI_TRY(check_launch("random_state"));
    return qi_normalise(s);
}

int qi_norm_sqr(const qi_state* s, double* out) {
    if (!s || !out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    QI_TRY(ensure_ctx());
    double v = 0.0;
    QI_TRY(reduce_norm_sqr(s, &v));
    if (s->world > 1) QI_TRY(shard_allreduce_sum(const_cast<qi_state*>(s), &v, 1));
    *out = v;
    return QI_OK;
}

int qi_inner_product(const qi_state* a, const qi_state* b, double out[2]) {
    if (!a || !b || !out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    // state.rs:891-897
    if (a->num_qubits == 0 || b->num_qubits == 0) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, 0, 0, "zero qubits");
    if (a->len != b->len) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, a->num_qubits, 0, "length mismatch");
    QI_TRY(ensure_ctx());
    if (memcmp(a->phys, b->phys, sizeof(a->phys)) != 0) {
        if (a->world > 1) return fail(QI_ERR_PEER, 0, 0, "inner product of sharded states with different qubit layouts");
        QI_TRY(canonicalise(const_cast<qi_state*>(a)));
        QI_TRY(canonicalise(const_cast<qi_state*>(b)));
    }
    QI_TRY(reduce_inner(a, b, out));
    if (a->world > 1) QI_TRY(shard_allreduce_sum(const_cast<qi_state*>(a), out, 2));
    return QI_OK;
}

int qi_normalise(qi_state* s) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    double nsq = 0.0;
    QI_TRY(qi_norm_sqr(s, &nsq));
    double norm = std::sqrt(nsq);
    if (norm == 0.0) return fail(QI_ERR_ZERO_NORM, 0, 0, "The state cannot be normalised because it has zero norm.");
    if (norm == 1.0) return QI_OK;   // state.rs:934-936
    LaunchScope ls(KF_ELEMENTWISE, 32.0 * (double)s->len);
    k_div_real<<<grid_for(s->len, kBlock), kBlock, 0, ctx().stream>>>(s->d, s->len, norm);
    return check_launch("normalise");
}

int qi_scale(qi_state* s, const double z[2]) {
    if (!s || !z) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    QI_TRY(ensure_ctx());
    if (!s->len) return QI_OK;
    LaunchScope ls(KF_ELEMENTWISE, 32.0 * (double)s->len);
    k_scale<<<grid_for(s->len, kBlock), kBlock, 0, ctx().stream>>>(s->d, s->len, make_double2(z[0], z[1]));
    return check_launch("scale");
}

static int addsub(qi_state* a, const qi_state* b, int sign) {
    if (!a || !b) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    // the reference panics on mismatched states (state.rs:2780-2785); the ABI reports it instead
    if (a->num_qubits != b->num_qubits || a->len != b->len)
        return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, b->num_qubits, 0, "Cannot add/subtract states with different numbers of qubits");
    QI_TRY(ensure_ctx());
    if (!a->len) return QI_OK;
    if (memcmp(a->phys, b->phys, sizeof(a->phys)) != 0) {
        if (a->world > 1) return fail(QI_ERR_PEER, 0, 0, "sharded states with different qubit layouts");
        QI_TRY(canonicalise(a));
        QI_TRY(canonicalise(const_cast<qi_state*>(b)));
    }
    LaunchScope ls(KF_ELEMENTWISE, 48.0 * (double)a->len);
    if (sign > 0) k_addsub<1><<<grid_for(a->len, kBlock), kBlock, 0, ctx().stream>>>(a->d, b->d, a->len);
    else k_addsub<-1><<<grid_for(a->len, kBlock), kBlock, 0, ctx().stream>>>(a->d, b->d, a->len);
    return check_launch("add/sub");
}
int qi_add(qi_state* a, const qi_state* b) { return addsub(a, b, 1); }
int qi_sub(qi_state* a, const qi_state* b) { return addsub(a, b, -1); }

int qi_conj(qi_state* s) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    QI_TRY(ensure_ctx());
    if (!s->len) return QI_OK;
    LaunchScope ls(KF_ELEMENTWISE, 32.0 * (double)s->len);
    k_conj<<<grid_for(s->len, kBlock), kBlock, 0, ctx().stream>>>(s->d, s->len);
    return check_launch("conj");
}

int qi_tensor_product(const qi_state* a, const qi_state* b, qi_state** out) {
    if (!a || !b || !out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    if (a->num_qubits == 0 || b->num_qubits == 0) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, 0, 0, "zero qubits");
    if (!a->consistent || !b->consistent || a->world > 1 || b->world > 1)
        return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "tensor_product needs whole, consistent states");
    uint32_t n = a->num_qubits + b->num_qubits;
    QI_TRY(check_nq(n));
    QI_TRY(canonicalise(const_cast<qi_state*>(a)));
    QI_TRY(canonicalise(const_cast<qi_state*>(b)));
    QI_TRY(alloc_state(n, 1ull << n, out));
    qi_state* s = *out;
    {
        LaunchScope ls(KF_ELEMENTWISE, 16.0 * (double)s->len);
        k_tensor<<<grid_for(s->len, kBlock), kBlock, 0, ctx().stream>>>(s->d, a->d, b->d, s->len, (int)b->num_qubits);
    }
    QI_TRY(check_launch("tensor_product"));
    // Self::new(new_state_vector): normalisation check, state.rs:835
    double nsq = 0.0;
    int st = reduce_norm_sqr(s, &nsq);
    if (st == QI_OK && std::fabs(nsq - 1.0) > 2.220446049250313e-16 * (double)s->len)
        st = fail(QI_ERR_STATE_VECTOR_NOT_NORMALISED, 0, 0, "State vector is not normalised");
    if (st != QI_OK) { qi_state_free(s); *out = nullptr; }
    return st;
}

}  // extern "C"
