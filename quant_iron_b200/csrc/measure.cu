// measure.cu -- marginal probabilities, seeded sampling by a device prefix scan, collapse, and the
// composed State::measure (state.rs:525-784).
//
// Determinism: every reduction here has a fixed shape (fixed chunking, shuffle/shared-memory trees,
// fixed-order final sums), so a probability table is bit-identical run to run.
// Shared-seed contract (the reference's RNG is unseedable): draw k of a call uses
// u_k = (splitmix64(seed, k) >> 11) * 2^-53 and selects the first bin with u < cumsum of the
// normalised table, falling back to the last bin (state.rs:601-619).
#include <algorithm>

#include "common.cuh"

namespace qi {

static const int kBlock = 256;

struct BinMap {
    int m;                 // measured qubits
    uint8_t pos[64];       // physical position of bin bit j
};

__device__ __forceinline__ uint64_t bin_to_mask(uint64_t bin, const BinMap& bm) {
    uint64_t v = 0;
    for (int j = 0; j < bm.m; j++) v |= ((bin >> j) & 1ull) << bm.pos[j];
    return v;
}

__device__ __forceinline__ double block_sum1(double v) {
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0) for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// (A) few bins, large rest space: block (bin, chunk) reduces its chunk of the rest space.
//     `ins` inserts zeros at the measured positions (ascending); the bin's bits are OR-ed in.
__global__ void __launch_bounds__(256) k_prob_chunks(const amp_t* __restrict__ a, uint64_t rest_total, uint64_t chunk_len,
                                                     BitInsert ins, BinMap bm, double* partials, int chunks) {
    uint64_t bin = blockIdx.y;
    int chunk = blockIdx.x;
    uint64_t fixed = bin_to_mask(bin, bm);
    uint64_t lo = (uint64_t)chunk * chunk_len, hi = lo + chunk_len;
    if (hi > rest_total) hi = rest_total;
    // four independent loads per thread and iteration (one load in flight per thread left the pass latency-bound at 0.66 of the
    // copy bandwidth); the four partial sums are added in a fixed order
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    const uint64_t step = blockDim.x;
    uint64_t r = lo + threadIdx.x;
    for (; r + 3 * step < hi; r += 4 * step) {
        const amp_t v0 = a[expand_index(r, ins) | fixed], v1 = a[expand_index(r + step, ins) | fixed];
        const amp_t v2 = a[expand_index(r + 2 * step, ins) | fixed], v3 = a[expand_index(r + 3 * step, ins) | fixed];
        acc0 += v0.x * v0.x + v0.y * v0.y;
        acc1 += v1.x * v1.x + v1.y * v1.y;
        acc2 += v2.x * v2.x + v2.y * v2.y;
        acc3 += v3.x * v3.x + v3.y * v3.y;
    }
    for (; r < hi; r += step) {
        const amp_t v = a[expand_index(r, ins) | fixed];
        acc0 += v.x * v.x + v.y * v.y;
    }
    double acc = (acc0 + acc1) + (acc2 + acc3);
    acc = block_sum1(acc);
    if (threadIdx.x == 0) partials[bin * chunks + chunk] = acc;
}
// (A') the same when some measured qubits sit at positions 0..2: a bin then owns one half / quarter / eighth of every 32-, 64-
//      or 128-byte piece of the state, and a block that reads ONE bin moves the whole sector for half of it (the marginal over
//      qubits {0, 7, 13, 22, 29} of a 30-qubit state ran at 0.53 of the copy bandwidth).  Here a thread reads all 2^L
//      amplitudes that differ in the L low measured bits -- neighbours in memory, the sector is used completely -- and keeps
//      2^L accumulators; a block covers one combination of the HIGH measured bits.  Same fixed-shape reductions.
struct LowBins {
    int L;                    // measured positions below 3
    uint8_t jlow[3];          // bin bit index of the k-th low measured position
    uint8_t plow[3];          // its physical position
    int nhigh;
    uint8_t jhigh[64];        // bin bit index of the k-th high measured position
};
template <int L>
__global__ void __launch_bounds__(256) k_prob_chunks_low(const amp_t* __restrict__ a, uint64_t rest_total, uint64_t chunk_len,
                                                         BitInsert ins, BinMap bm, LowBins lb, double* partials, int chunks) {
    const uint64_t hb = blockIdx.y;
    const int chunk = blockIdx.x;
    uint64_t bin_h = 0;
    for (int k = 0; k < lb.nhigh; k++) bin_h |= ((hb >> k) & 1ull) << lb.jhigh[k];
    const uint64_t fixed = bin_to_mask(bin_h, bm);
    uint64_t lowmask[1 << L];
    uint64_t bin_of[1 << L];
#pragma unroll
    for (int c = 0; c < (1 << L); c++) {
        uint64_t m = 0, b = bin_h;
#pragma unroll
        for (int k = 0; k < L; k++) if ((c >> k) & 1) { m |= 1ull << lb.plow[k]; b |= 1ull << lb.jlow[k]; }
        lowmask[c] = m;
        bin_of[c] = b;
    }
    uint64_t lo = (uint64_t)chunk * chunk_len, hi = lo + chunk_len;
    if (hi > rest_total) hi = rest_total;
    double acc[1 << L];
#pragma unroll
    for (int c = 0; c < (1 << L); c++) acc[c] = 0.0;
    constexpr int U = L == 1 ? 2 : 1;                 // L = 1: two rest indices per iteration (four loads in flight)
    const uint64_t step = blockDim.x;
    uint64_t r = lo + threadIdx.x;
    for (; r + (U - 1) * step < hi; r += U * step) {
        amp_t v[U << L];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t base = expand_index(r + u * step, ins) | fixed;
#pragma unroll
            for (int c = 0; c < (1 << L); c++) v[(u << L) | c] = a[base | lowmask[c]];
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int c = 0; c < (1 << L); c++) acc[c] += v[(u << L) | c].x * v[(u << L) | c].x + v[(u << L) | c].y * v[(u << L) | c].y;
    }
    for (; r < hi; r += step) {
        const uint64_t base = expand_index(r, ins) | fixed;
#pragma unroll
        for (int c = 0; c < (1 << L); c++) {
            const amp_t v = a[base | lowmask[c]];
            acc[c] += v.x * v.x + v.y * v.y;
        }
    }
#pragma unroll
    for (int c = 0; c < (1 << L); c++) {
        const double t = block_sum1(acc[c]);
        if (threadIdx.x == 0) partials[bin_of[c] * chunks + chunk] = t;
    }
}
__global__ void k_prob_finish(const double* partials, int chunks, uint64_t nbins, double* probs) {
    uint64_t bin = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= nbins) return;
    double acc = 0.0;
    for (int c = 0; c < chunks; c++) acc += partials[bin * chunks + c];
    probs[bin] = acc;
}
// (B) many bins, small rest space: one thread per bin walks the rest space serially.
__global__ void __launch_bounds__(256) k_prob_serial(const amp_t* __restrict__ a, uint64_t nbins, uint64_t rest_total,
                                                     BitInsert ins, BinMap bm, double* probs) {
    uint64_t bin = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= nbins) return;
    uint64_t fixed = bin_to_mask(bin, bm);
    double acc = 0.0;
    for (uint64_t r = 0; r < rest_total; r++) {
        amp_t v = a[expand_index(r, ins) | fixed];
        acc += v.x * v.x + v.y * v.y;
    }
    probs[bin] = acc;
}

// ---- prefix scan (3 phases, fixed shape) ---------------------------------------------------------
static const int kScanItems = 8;                       // per thread
static const int kScanTile = kBlock * kScanItems;      // 2048 per block

__device__ __forceinline__ double block_exclusive_scan(double v, double* total) {
    // v = this thread's sum; returns the exclusive prefix over the block in thread order
    __shared__ double wsum[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        double s = lane < nw ? wsum[lane] : 0.0;
        double si = s;
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        if (lane < nw) wsum[lane] = si - s;     // exclusive warp offsets
        if (lane == nw - 1 && total) *total = si;
    }
    __syncthreads();
    double r = wsum[w] + inc - v;
    __syncthreads();
    return r;
}

// phase 1: per-tile totals of p[i] * scale
__global__ void __launch_bounds__(256) k_scan_tile_sums(const double* __restrict__ p, uint64_t n, double total, double* tile_sums) {
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) if (base + j < n) s += p[base + j] / total;
    __shared__ double tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
// phase 2: exclusive scan of the tile sums, single block, serial over tiles of the tile-sum array
__global__ void __launch_bounds__(256) k_scan_offsets(double* tile_sums, uint64_t ntiles) {
    __shared__ double carry_sh;
    if (threadIdx.x == 0) carry_sh = 0.0;
    __syncthreads();
    for (uint64_t base = 0; base < ntiles; base += kBlock) {
        uint64_t i = base + threadIdx.x;
        double v = i < ntiles ? tile_sums[i] : 0.0;
        __shared__ double tot;
        double ex = block_exclusive_scan(v, &tot);
        double carry = carry_sh;
        if (i < ntiles) tile_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = carry + tot;
        __syncthreads();
    }
}
// phase 3: inclusive cdf[i] = offset(tile) + prefix within the tile
__global__ void __launch_bounds__(256) k_scan_write(const double* __restrict__ p, uint64_t n, double total, const double* tile_offsets, double* cdf) {
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    double v[kScanItems];
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) { v[j] = (base + j < n) ? p[base + j] / total : 0.0; s += v[j]; }
    double ex = block_exclusive_scan(s, nullptr) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int j = 0; j < kScanItems; j++) { ex += v[j]; if (base + j < n) cdf[base + j] = ex; }
}

// first i with u < cdf[i], else n-1
__global__ void k_sample(const double* __restrict__ cdf, uint64_t n, uint64_t seed, uint64_t first_draw, uint64_t shots, uint64_t* bins) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= shots) return;
    double u = uniform_at(seed, first_draw + k);
    uint64_t lo = 0, hi = n;            // invariant: answer in [lo, hi]
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (u < cdf[mid]) hi = mid; else lo = mid + 1;
    }
    bins[k] = lo < n ? lo : n - 1;
}

// ---- collapse -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_collapse_zero(amp_t* __restrict__ a, uint64_t len, uint64_t sel_mask, uint64_t sel_val, uint64_t high, double* partials) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    // a thread owns both amplitudes of a 32-byte sector, so that the zeros of a sector are written by ONE thread (two adjacent
    // 16-byte stores) instead of by two threads of different iterations when qubit 0 is among the measured ones
    const amp_t zero = make_double2(0.0, 0.0);
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; 2 * p < len; p += stride) {
        const uint64_t i0 = 2 * p, i1 = i0 + 1;
        const bool k0 = ((i0 | high) & sel_mask) == sel_val, k1 = i1 < len && ((i1 | high) & sel_mask) == sel_val;
        if (k0) { const amp_t v = a[i0]; acc += v.x * v.x + v.y * v.y; } else a[i0] = zero;
        if (i1 < len) { if (k1) { const amp_t v = a[i1]; acc += v.x * v.x + v.y * v.y; } else a[i1] = zero; }
    }
    acc = block_sum1(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
__global__ void k_sum_doubles(const double* partials, int count, double* out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) acc += partials[i];
    acc = block_sum1(acc);
    if (threadIdx.x == 0) out[0] = acc;
}
__global__ void __launch_bounds__(256) k_div_sel(amp_t* __restrict__ a, uint64_t total, BitInsert ins, double d) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    uint64_t i = expand_index(k, ins);
    amp_t v = a[i];
    a[i] = make_double2(v.x / d, v.y / d);
}

// ---- host side -----------------------------------------------------------------------------------
static int check_qubits(const qi_state* s, const uint32_t* qubits, uint32_t m, std::vector<uint32_t>* actual) {
    // state.rs:531-553
    actual->clear();
    if (m == 0) for (uint32_t q = 0; q < s->num_qubits; q++) actual->push_back(q);
    else {
        if (!qubits) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "qubits is NULL");
        actual->assign(qubits, qubits + m);
    }
    if (actual->size() > s->num_qubits) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, s->num_qubits, 0, "more measured qubits than qubits");
    for (uint32_t q : *actual)
        if (q >= s->num_qubits) return fail(QI_ERR_INVALID_QUBIT_INDEX, q, s->num_qubits, "Invalid qubit index");
    if (!s->consistent) return fail(QI_ERR_INVALID_NUMBER_OF_QUBITS, s->num_qubits, 0, "state vector length is not 2^num_qubits");
    return QI_OK;
}

// device table of 2^k un-normalised marginals over DISTINCT LOCAL bit positions (bin bit j <-> pos[j])
static int local_table(const qi_state* s, const std::vector<int>& pos, double** d_probs_out) {
    Context& c = ctx();
    const int m = (int)pos.size();
    if (m > 34) return fail(QI_ERR_INVALID_INPUT_VALUE, (uint64_t)m, 0, "probability table too large");
    const uint64_t nbins = 1ull << m;
    double* d_probs = nullptr;
    QI_TRY(dev_alloc((void**)&d_probs, nbins * sizeof(double)));
    BinMap bm;
    memset(&bm, 0, sizeof(bm));
    bm.m = m;
    for (int j = 0; j < m; j++) bm.pos[j] = (uint8_t)pos[j];
    std::vector<int> sorted(pos);
    std::sort(sorted.begin(), sorted.end());
    const uint64_t rest_total = s->len >> m;
    BitInsert ins = make_insert(sorted, {});
    int st = QI_OK;
    LowBins lb;
    memset(&lb, 0, sizeof(lb));
    for (int j = 0; j < m; j++) {
        if (pos[j] < 3) { lb.jlow[lb.L] = (uint8_t)j; lb.plow[lb.L] = (uint8_t)pos[j]; lb.L++; }
        else lb.jhigh[lb.nhigh++] = (uint8_t)j;
    }
    if (rest_total >= 1024 && nbins <= 32768) {
        const uint64_t nblocks_y = nbins >> lb.L;             // one block row per combination of the high measured bits
        int chunks = (int)std::min<uint64_t>(std::max<uint64_t>(1, rest_total / 4096), std::max<uint64_t>(1, (uint64_t)(c.sm_count * 8) / nblocks_y));
        if (chunks < 1) chunks = 1;
        uint64_t chunk_len = (rest_total + chunks - 1) / chunks;
        st = ensure_partials((size_t)nbins * chunks);
        if (st == QI_OK) {
            LaunchScope ls(KF_PROB, 16.0 * (double)s->len);
            dim3 grid(chunks, (unsigned)nblocks_y);
            if (lb.L == 0) k_prob_chunks<<<grid, kBlock, 0, c.stream>>>(s->d, rest_total, chunk_len, ins, bm, c.d_partials, chunks);
            else if (lb.L == 1) k_prob_chunks_low<1><<<grid, kBlock, 0, c.stream>>>(s->d, rest_total, chunk_len, ins, bm, lb, c.d_partials, chunks);
            else if (lb.L == 2) k_prob_chunks_low<2><<<grid, kBlock, 0, c.stream>>>(s->d, rest_total, chunk_len, ins, bm, lb, c.d_partials, chunks);
            else k_prob_chunks_low<3><<<grid, kBlock, 0, c.stream>>>(s->d, rest_total, chunk_len, ins, bm, lb, c.d_partials, chunks);
            k_prob_finish<<<(unsigned)((nbins + kBlock - 1) / kBlock), kBlock, 0, c.stream>>>(c.d_partials, chunks, nbins, d_probs);
        }
    } else {
        LaunchScope ls(KF_PROB, 16.0 * (double)s->len);
        k_prob_serial<<<(unsigned)((nbins + kBlock - 1) / kBlock), kBlock, 0, c.stream>>>(s->d, nbins, rest_total, ins, bm, d_probs);
    }
    if (st == QI_OK) st = check_launch("probabilities");
    if (st != QI_OK) { dev_free(d_probs, nbins * sizeof(double)); return st; }
    *d_probs_out = d_probs;
    return QI_OK;
}

// device table of 2^m un-normalised probabilities over `qubits` (bin bit j <-> qubits[j], state.rs:567-571);
// on a sharded state the table is summed over ranks, so every rank holds the same full table
static int device_probabilities(const qi_state* s, const std::vector<uint32_t>& qubits, double** d_probs_out) {
    Context& c = ctx();
    const int m = (int)qubits.size();
    const int nl = (int)s->n_local;
    std::vector<int> pos(m), local_pos;
    bool plain = (s->world == 1);
    for (int j = 0; j < m; j++) {
        pos[j] = s->phys[qubits[j]];
        if (pos[j] < nl) {
            if (std::find(local_pos.begin(), local_pos.end(), pos[j]) == local_pos.end()) local_pos.push_back(pos[j]);
            else plain = false;                 // repeated qubit (the reference does not reject it)
        } else plain = false;
    }
    if (plain) return local_table(s, pos, d_probs_out);
    // general case: local table over the distinct local positions, scattered into the full table on the
    // host (bins whose rank bits / repeated bits disagree stay 0), then summed across ranks
    if (m > 20) return fail(QI_ERR_INVALID_INPUT_VALUE, (uint64_t)m, 0, "too many measured qubits for a sharded state");
    double* d_local = nullptr;
    QI_TRY(local_table(s, local_pos, &d_local));
    const uint64_t nloc = 1ull << local_pos.size(), nbins = 1ull << m;
    std::vector<double> hl(nloc), full(nbins, 0.0);
    cudaError_t e = cudaMemcpyAsync(hl.data(), d_local, nloc * sizeof(double), cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    dev_free(d_local, nloc * sizeof(double));
    if (e != cudaSuccess) return cuda_fail(e, "D2H local table");
    for (uint64_t b = 0; b < nbins; b++) {
        uint64_t lb = 0, seen = 0, val = 0;
        bool ok = true;
        for (int j = 0; j < m && ok; j++) {
            const uint64_t bit = (b >> j) & 1;
            if (pos[j] >= nl) { ok = (((uint64_t)s->rank >> (pos[j] - nl)) & 1) == bit; continue; }
            const int k = (int)(std::find(local_pos.begin(), local_pos.end(), pos[j]) - local_pos.begin());
            if ((seen >> k) & 1) ok = ((val >> k) & 1) == bit;
            else { seen |= 1ull << k; val |= bit << k; lb |= bit << k; }
        }
        if (ok) full[b] = hl[lb];
    }
    if (s->world > 1)
        for (uint64_t off = 0; off < nbins; off += 512)
            QI_TRY(shard_allreduce_sum(const_cast<qi_state*>(s), full.data() + off, (int)std::min<uint64_t>(512, nbins - off)));
    double* d_probs = nullptr;
    QI_TRY(dev_alloc((void**)&d_probs, nbins * sizeof(double)));
    e = cudaMemcpyAsync(d_probs, full.data(), nbins * sizeof(double), cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) { dev_free(d_probs, nbins * sizeof(double)); return cuda_fail(e, "H2D full table"); }
    *d_probs_out = d_probs;
    return QI_OK;
}

static int device_total(const double* d_probs, uint64_t nbins, double* total) {
    // total of the table: fixed-shape reduction (stage 1 over tiles via the scan tile sums)
    Context& c = ctx();
    uint64_t ntiles = (nbins + kScanTile - 1) / kScanTile;
    QI_TRY(ensure_partials(ntiles + 8));
    k_scan_tile_sums<<<(unsigned)ntiles, kBlock, 0, c.stream>>>(d_probs, nbins, 1.0, c.d_partials);
    k_sum_doubles<<<1, kBlock, 0, c.stream>>>(c.d_partials, (int)ntiles, c.d_result);
    QI_TRY(check_launch("prob_total"));
    QI_CUDA(cudaMemcpyAsync(c.h_result, c.d_result, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QI_CUDA(cudaStreamSynchronize(c.stream));
    *total = c.h_result[0];
    return QI_OK;
}

static int sample_from_table(const double* d_probs, uint64_t nbins, uint64_t seed, uint64_t first_draw, uint64_t shots, uint64_t* host_bins) {
    Context& c = ctx();
    double total = 0.0;
    QI_TRY(device_total(d_probs, nbins, &total));
    if (total < 2.220446049250313e-16) return fail(QI_ERR_UNKNOWN, 0, 0, "total probability is zero");   // state.rs:592-594
    const double scale = total;          // kernels divide: prob / total (state.rs:595-598)
    uint64_t ntiles = (nbins + kScanTile - 1) / kScanTile;
    double* d_cdf = nullptr;
    uint64_t* d_bins = nullptr;
    QI_TRY(dev_alloc((void**)&d_cdf, nbins * sizeof(double)));
    cudaError_t e = cudaSuccess;
    int st = dev_alloc((void**)&d_bins, shots * sizeof(uint64_t));
    if (st != QI_OK) { dev_free(d_cdf, nbins * sizeof(double)); return st; }
    st = ensure_partials(ntiles + 8);
    if (st == QI_OK) {
        {
            LaunchScope ls(KF_SCAN, 24.0 * (double)nbins);
            k_scan_tile_sums<<<(unsigned)ntiles, kBlock, 0, c.stream>>>(d_probs, nbins, scale, c.d_partials);
            k_scan_offsets<<<1, kBlock, 0, c.stream>>>(c.d_partials, ntiles);
            k_scan_write<<<(unsigned)ntiles, kBlock, 0, c.stream>>>(d_probs, nbins, scale, c.d_partials, d_cdf);
        }
        {
            LaunchScope ls(KF_SAMPLE, 8.0 * (double)shots);
            k_sample<<<(unsigned)((shots + kBlock - 1) / kBlock), kBlock, 0, c.stream>>>(d_cdf, nbins, seed, first_draw, shots, d_bins);
        }
        st = check_launch("sample");
    }
    if (st == QI_OK) {
        e = cudaMemcpyAsync(host_bins, d_bins, shots * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
        if (e != cudaSuccess) st = cuda_fail(e, "D2H bins");
    }
    dev_free(d_cdf, nbins * sizeof(double));
    dev_free(d_bins, shots * sizeof(uint64_t));
    return st;
}

static int collapse_impl(qi_state* s, const std::vector<uint32_t>& qubits, uint64_t bin) {
    Context& c = ctx();
    uint64_t sel_mask = 0, sel_val = 0;
    for (size_t j = 0; j < qubits.size(); j++) {
        uint64_t b = 1ull << s->phys[qubits[j]];
        sel_mask |= b;
        if ((bin >> j) & 1) sel_val |= b;
    }
    const uint64_t high = (uint64_t)s->rank << s->n_local;
    int g = c.sm_count * 4;
    uint64_t need = (s->len + kBlock - 1) / kBlock;
    if (need < (uint64_t)g) g = (int)need;
    QI_TRY(ensure_partials((size_t)g));
    {
        LaunchScope ls(KF_COLLAPSE, 32.0 * (double)s->len);
        k_collapse_zero<<<g, kBlock, 0, c.stream>>>(s->d, s->len, sel_mask, sel_val, high, c.d_partials);
        k_sum_doubles<<<1, kBlock, 0, c.stream>>>(c.d_partials, g, c.d_result);
    }
    QI_TRY(check_launch("collapse"));
    QI_CUDA(cudaMemcpyAsync(c.h_result, c.d_result, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QI_CUDA(cudaStreamSynchronize(c.stream));
    double nsq = c.h_result[0];
    if (s->world > 1) QI_TRY(shard_allreduce_sum(s, &nsq, 1));
    if (nsq > 2.220446049250313e-16) {       // state.rs:649-654
        const double f = std::sqrt(nsq);
        // only the surviving amplitudes are non-zero: divide those (local selected bits fixed)
        std::vector<int> zeros, ones;
        const uint64_t local_mask = s->len - 1;
        bool rank_selected = ((high & sel_mask) == (sel_val & ~local_mask));
        if (rank_selected) {
            for (int p = 0; p < (int)s->n_local; p++)
                if ((sel_mask >> p) & 1) { if ((sel_val >> p) & 1) ones.push_back(p); else zeros.push_back(p); }
            BitInsert ins = make_insert(zeros, ones);
            uint64_t total = s->len >> (zeros.size() + ones.size());
            LaunchScope ls(KF_COLLAPSE, 32.0 * (double)total);
            k_div_sel<<<(unsigned)((total + kBlock - 1) / kBlock), kBlock, 0, c.stream>>>(s->d, total, ins, f);
            QI_TRY(check_launch("collapse_normalise"));
        }
    }
    else       // State::new(collapsed) re-checks the norm (state.rs:667, 117-121): a bin of probability zero fails there
        return fail(QI_ERR_STATE_VECTOR_NOT_NORMALISED, 0, 0, "collapsed state has zero norm (the selected outcome has probability 0)");
    return QI_OK;
}

static int apply_single_all(qi_state* s, const std::vector<uint32_t>& qubits, int kind, const double* params) {
    for (uint32_t q : qubits) {
        qi_gate g;
        memset(&g, 0, sizeof(g));
        g.kind = kind;
        g.num_targets = 1;
        g.targets[0] = q;
        if (params) memcpy(g.params, params, 8 * sizeof(double));
        QI_TRY(qi_apply_gate(s, &g));
    }
    return QI_OK;
}

}  // namespace qi

using namespace qi;

extern "C" {

int qi_probabilities(const qi_state* s, const uint32_t* qubits, uint32_t m, double* out) {
    if (!s || !out) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    std::vector<uint32_t> q;
    QI_TRY(check_qubits(s, qubits, m, &q));
    QI_TRY(ensure_ctx());
    double* d_probs = nullptr;
    QI_TRY(device_probabilities(s, q, &d_probs));
    uint64_t nbins = 1ull << q.size();
    cudaError_t e = cudaMemcpyAsync(out, d_probs, nbins * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);
    dev_free(d_probs, nbins * sizeof(double));
    if (e != cudaSuccess) return cuda_fail(e, "D2H probabilities");
    return QI_OK;
}

int qi_sample(const qi_state* s, const uint32_t* qubits, uint32_t m, uint64_t shots, uint64_t seed, uint64_t* bins) {
    if (!s || !bins) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "NULL argument");
    if (shots == 0) return fail(QI_ERR_INVALID_NUMBER_OF_MEASUREMENTS, 0, 0, "Invalid number of measurements: 0");   // state.rs:756-758
    std::vector<uint32_t> q;
    QI_TRY(check_qubits(s, qubits, m, &q));
    QI_TRY(ensure_ctx());
    double* d_probs = nullptr;
    QI_TRY(device_probabilities(s, q, &d_probs));
    int st = sample_from_table(d_probs, 1ull << q.size(), seed, 0, shots, bins);
    dev_free(d_probs, (1ull << q.size()) * sizeof(double));
    return st;
}

int qi_collapse(qi_state* s, const uint32_t* qubits, uint32_t m, uint64_t bin) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    std::vector<uint32_t> q;
    QI_TRY(check_qubits(s, qubits, m, &q));
    if (q.size() < 64 && bin >= (1ull << q.size())) return fail(QI_ERR_INVALID_INPUT_VALUE, bin, 0, "bin out of range");
    QI_TRY(ensure_ctx());
    return collapse_impl(s, q, bin);
}

int qi_measure(qi_state* s, int basis, const double* custom_u, const uint32_t* qubits, uint32_t m,
               uint64_t seed, uint64_t draw_index, uint8_t* outcomes, uint64_t* bin_out) {
    if (!s) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    std::vector<uint32_t> q;
    QI_TRY(check_qubits(s, qubits, m, &q));
    QI_TRY(ensure_ctx());
    double udag[8];
    switch (basis) {
        case QI_BASIS_COMPUTATIONAL: break;
        case QI_BASIS_X: QI_TRY(apply_single_all(s, q, QI_GATE_H, nullptr)); break;                 // state.rs:672
        case QI_BASIS_Y:                                                                            // state.rs:689-690
            QI_TRY(apply_single_all(s, q, QI_GATE_SDG, nullptr));
            QI_TRY(apply_single_all(s, q, QI_GATE_H, nullptr));
            break;
        case QI_BASIS_CUSTOM:                                                                       // state.rs:708
            if (!custom_u) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "custom basis needs a matrix");
            QI_TRY(qi_unitary2_check(custom_u));                  // unitary_multi -> Unitary2::new (state.rs:1988)
            QI_TRY(apply_single_all(s, q, QI_GATE_U2, custom_u));
            // calculate_adjoint, state.rs:17-30
            udag[0] = custom_u[0]; udag[1] = -custom_u[1]; udag[2] = custom_u[4]; udag[3] = -custom_u[5];
            udag[4] = custom_u[2]; udag[5] = -custom_u[3]; udag[6] = custom_u[6]; udag[7] = -custom_u[7];
            break;
        default: return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)basis, 0, "unknown basis");
    }
    double* d_probs = nullptr;
    QI_TRY(device_probabilities(s, q, &d_probs));
    uint64_t bin = 0;
    int st = sample_from_table(d_probs, 1ull << q.size(), seed, draw_index, 1, &bin);
    dev_free(d_probs, (1ull << q.size()) * sizeof(double));
    QI_TRY(st);
    QI_TRY(collapse_impl(s, q, bin));
    switch (basis) {
        case QI_BASIS_X: QI_TRY(apply_single_all(s, q, QI_GATE_H, nullptr)); break;                 // state.rs:677-679
        case QI_BASIS_Y:                                                                            // state.rs:695-698
            QI_TRY(apply_single_all(s, q, QI_GATE_H, nullptr));
            QI_TRY(apply_single_all(s, q, QI_GATE_S, nullptr));
            break;
        case QI_BASIS_CUSTOM: QI_TRY(apply_single_all(s, q, QI_GATE_U2, udag)); break;             // state.rs:716-720
        default: break;
    }
    if (outcomes) for (size_t j = 0; j < q.size(); j++) outcomes[j] = (uint8_t)((bin >> j) & 1);   // state.rs:657-660
    if (bin_out) *bin_out = bin;
    return QI_OK;
}

}  // extern "C"
