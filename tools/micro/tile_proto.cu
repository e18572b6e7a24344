// Microbenchmark (experiment, not product code): the access pattern and the round structure of a two-level
// "CTA tile" pass -- a block stages 2^11 amplitudes (low L index bits + 11-L chosen qubits) and works on them in
// ROUNDS, each round holding 4 of the 11 tile qubits in registers (16 amplitudes per thread, 128 threads);
// between rounds the registers are regrouped through 32 KiB of shared memory (one __syncthreads per regroup,
// XOR-folded 16-byte swizzle).  Question 1: does a tile made of 128-byte pieces stream as fast as 512-byte
// pieces?  Question 2: what do the regroups and G register gates per round cost against the 5.3 ms HBM floor?
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

typedef double2 amp_t;
static const int kT = 11;          // tile qubits
static const int kMaxRounds = 6;

struct Round {
    uint16_t lbase_ins[4];         // the 4 register local bits (ascending): zero-insert positions for the thread index
    uint16_t soff[16];             // slot -> local offset
    uint16_t sswz[16];             // swizzled slot offset
    int ngates;                    // register gates applied in this round (cycling over the 4 bits)
};
struct Prog {
    uint8_t tpos[kT];              // physical qubit of local bit j (ascending)
    int nrounds;
    int lean;
    uint64_t goff_first[16];       // global offsets of the slots of round 0 / the last round
    uint64_t goff_last[16];
    Round r[kMaxRounds];
};

__host__ __device__ inline uint32_t swz(uint32_t j) { return j ^ ((j >> 3) & 7) ^ ((j >> 6) & 7) ^ ((j >> 9) & 3); }
__host__ __device__ inline uint64_t ins0(uint64_t k, int p) { return ((k >> p) << (p + 1)) | (k & ((1ull << p) - 1ull)); }

template <int B, bool LEAN>
__device__ __forceinline__ void gate(amp_t (&v)[16], double k0, double k1) {
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int s0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1)), s1 = s0 | (1 << B);
        const amp_t a0 = v[s0], a1 = v[s1];
        if (LEAN) {
            v[s0] = make_double2(fma(k0, a1.x, a0.x), fma(k0, a1.y, a0.y));
            v[s1] = make_double2(fma(k1, a0.x, -a1.x), fma(k1, a0.y, -a1.y));
        } else {
            v[s0] = make_double2(k0 * a0.x + k0 * a1.x, k0 * a0.y + k0 * a1.y);
            v[s1] = make_double2(k0 * a0.x - k0 * a1.x, k0 * a0.y - k0 * a1.y);
        }
    }
}

template <bool LEAN>
__device__ __forceinline__ void run_round(amp_t (&v)[16], int ngates, double k0, double k1) {
#pragma unroll 1
    for (int g = 0; g < ngates; g++) {
        switch (g & 3) {
            case 0: gate<0, LEAN>(v, k0, k1); break;
            case 1: gate<1, LEAN>(v, k0, k1); break;
            case 2: gate<2, LEAN>(v, k0, k1); break;
            default: gate<3, LEAN>(v, k0, k1); break;
        }
    }
}

__global__ void __launch_bounds__(128, 4) k_tile(amp_t* __restrict__ a, uint64_t ntiles, const __grid_constant__ Prog P, double k0, double k1) {
    __shared__ __align__(16) amp_t sm[1 << kT];
    const int t = threadIdx.x;
    // per-thread, per-round local base index (thread bits spread over the non-register local bits)
    uint32_t lb_first, lb_last;
    {
        uint32_t b = t;
        for (int i = 0; i < 4; i++) b = (uint32_t)ins0(b, P.r[0].lbase_ins[i]);
        lb_first = b;
        b = t;
        for (int i = 0; i < 4; i++) b = (uint32_t)ins0(b, P.r[P.nrounds - 1].lbase_ins[i]);
        lb_last = b;
    }
    uint64_t g_first = 0, g_last = 0;
    for (int j = 0; j < kT; j++) {
        if ((lb_first >> j) & 1) g_first |= 1ull << P.tpos[j];
        if ((lb_last >> j) & 1) g_last |= 1ull << P.tpos[j];
    }
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint64_t tb = tile;
        for (int j = 0; j < kT; j++) tb = ins0(tb, P.tpos[j]);
        amp_t v[16];
#pragma unroll
        for (int s = 0; s < 16; s++) v[s] = __ldcs(a + tb + g_first + P.goff_first[s]);
        for (int r = 0; r < P.nrounds; r++) {
            if (P.lean) run_round<true>(v, P.r[r].ngates, k0, k1);
            else run_round<false>(v, P.r[r].ngates, k0, k1);
            if (r + 1 < P.nrounds) {
                uint32_t b = t;
                for (int i = 0; i < 4; i++) b = (uint32_t)ins0(b, P.r[r].lbase_ins[i]);
                uint32_t sb = swz(b);
#pragma unroll
                for (int s = 0; s < 16; s++) sm[sb ^ P.r[r].sswz[s]] = v[s];
                __syncthreads();
                b = t;
                for (int i = 0; i < 4; i++) b = (uint32_t)ins0(b, P.r[r + 1].lbase_ins[i]);
                sb = swz(b);
#pragma unroll
                for (int s = 0; s < 16; s++) v[s] = sm[sb ^ P.r[r + 1].sswz[s]];
            }
        }
#pragma unroll
        for (int s = 0; s < 16; s++) __stcs(a + tb + g_last + P.goff_last[s], v[s]);
    }
}

// reference point: the shipped warp-tile pattern (each warp: 16 x 512 B, slots on 4 window qubits), G gates
__global__ void __launch_bounds__(128, 4) k_warp(amp_t* __restrict__ a, uint64_t ntiles, int q0, int ngates, int lean, double k0, double k1) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t tile = warp; tile < ntiles; tile += nwarps) {
        uint64_t k = (tile << 5) | lane;
        for (int i = 0; i < 4; i++) k = ins0(k, q0 + i);
        amp_t v[16];
#pragma unroll
        for (int s = 0; s < 16; s++) v[s] = __ldcs(a + k + ((uint64_t)s << q0));
        if (lean) run_round<true>(v, ngates, k0, k1); else run_round<false>(v, ngates, k0, k1);
#pragma unroll
        for (int s = 0; s < 16; s++) __stcs(a + k + ((uint64_t)s << q0), v[s]);
    }
}

static Prog make_prog(const std::vector<int>& tile_qubits, const std::vector<std::vector<int>>& round_regs, int gates_per_round, int lean) {
    Prog P{};
    std::vector<int> tq = tile_qubits;
    std::sort(tq.begin(), tq.end());
    for (int j = 0; j < kT; j++) P.tpos[j] = (uint8_t)tq[j];
    P.nrounds = (int)round_regs.size();
    P.lean = lean;
    for (int r = 0; r < P.nrounds; r++) {
        std::vector<int> rb = round_regs[r];
        std::sort(rb.begin(), rb.end());
        for (int i = 0; i < 4; i++) P.r[r].lbase_ins[i] = (uint16_t)rb[i];
        for (int s = 0; s < 16; s++) {
            uint32_t o = 0;
            for (int i = 0; i < 4; i++) if ((s >> i) & 1) o |= 1u << rb[i];
            P.r[r].soff[s] = (uint16_t)o;
            P.r[r].sswz[s] = (uint16_t)swz(o);
        }
        P.r[r].ngates = gates_per_round;
    }
    for (int s = 0; s < 16; s++) {
        uint64_t g0 = 0, g1 = 0;
        for (int j = 0; j < kT; j++) {
            if ((P.r[0].soff[s] >> j) & 1) g0 |= 1ull << tq[j];
            if ((P.r[P.nrounds - 1].soff[s] >> j) & 1) g1 |= 1ull << tq[j];
        }
        P.goff_first[s] = g0;
        P.goff_last[s] = g1;
    }
    return P;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 30;
    const uint64_t len = 1ull << n;
    amp_t* d;
    if (cudaMalloc(&d, len * sizeof(amp_t)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(d, 0, len * sizeof(amp_t));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_tile, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double k0 = 0.70710678118654752, k1 = 1.0;
    const double gb = 32.0 * (double)len / 1e9;
    auto time_tile = [&](const char* name, const Prog& P, int grid_mult) {
        const uint64_t ntiles = len >> kT;
        const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)sms * grid_mult);
        k_tile<<<grid, 128>>>(d, ntiles, P, k0, k1);
        cudaEventRecord(e0);
        const int reps = 5;
        for (int i = 0; i < reps; i++) k_tile<<<grid, 128>>>(d, ntiles, P, k0, k1);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        cudaError_t e = cudaGetLastError();
        printf("%-64s rounds=%d gates/round=%2d lean=%d grid=%ux: %7.3f ms  %6.0f GB/s %s\n", name, P.nrounds, P.r[0].ngates, P.lean, grid_mult, ms, gb / (ms * 1e-3),
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    auto window = [&](int L, int q0) { std::vector<int> t; for (int j = 0; j < L; j++) t.push_back(j); for (int j = 0; j < kT - L; j++) t.push_back(q0 + j); return t; };
    // --- question 1: streaming efficiency of the tile shapes (1 round, 0 gates = copy through registers) ---
    const std::vector<std::vector<int>> top4 = {{7, 8, 9, 10}};
    for (int L : {5, 3, 2, 1}) {
        for (int q0 : {L, 12, 19, n - (kT - L)}) {
            char nm[128]; snprintf(nm, sizeof nm, "copy L=%d window=%d..%d", L, q0, q0 + kT - L - 1);
            time_tile(nm, make_prog(window(L, q0), top4, 0, 0), 16);
        }
    }
    { std::vector<int> t = {0, 1, 2, 5, 9, 13, 17, 21, 25, 27, 29}; if (n >= 30) time_tile("copy L=3 scattered {5,9,13,17,21,25,27,29}", make_prog(t, top4, 0, 0), 16); }
    for (int gm : {4, 8, 32, 64}) time_tile("copy L=3 window=12..19 (grid sweep)", make_prog(window(3, 12), top4, 0, 0), gm);
    // regs on low local bits for the first / last round (uncoalesced direct access) -- to see what it costs
    time_tile("copy L=3 window=12..19 regs={0,1,2,3}", make_prog(window(3, 12), {{0, 1, 2, 3}}, 0, 0), 16);
    time_tile("copy L=3 window=12..19 regs={3,4,5,6}", make_prog(window(3, 12), {{3, 4, 5, 6}}, 0, 0), 16);
    // --- question 2: regroup cost and gate cost ---
    const std::vector<std::vector<int>> R2 = {{7, 8, 9, 10}, {3, 4, 5, 6}}, R2b = {{7, 8, 9, 10}, {3, 4, 5, 6}, {7, 8, 9, 10}},
                                        R3 = {{7, 8, 9, 10}, {3, 4, 5, 6}, {5, 6, 7, 8}, {7, 8, 9, 10}},
                                        R4 = {{7, 8, 9, 10}, {3, 4, 5, 6}, {5, 6, 7, 8}, {4, 6, 8, 10}, {7, 8, 9, 10}},
                                        R5 = {{7, 8, 9, 10}, {3, 4, 5, 6}, {5, 6, 7, 8}, {4, 6, 8, 10}, {3, 5, 7, 9}, {7, 8, 9, 10}},
                                        Rlow = {{7, 8, 9, 10}, {0, 1, 2, 3}, {7, 8, 9, 10}}, Rlow2 = {{7, 8, 9, 10}, {0, 3, 6, 9}, {7, 8, 9, 10}};
    for (int lean : {0, 1}) {
        for (int g : {0, 4, 8, 12, 16}) {
            if (lean && g == 0) continue;
            time_tile("L=3 window=12..19", make_prog(window(3, 12), top4, g, lean), 16);
            time_tile("L=3 window=12..19", make_prog(window(3, 12), R2, g, lean), 16);
            time_tile("L=3 window=12..19", make_prog(window(3, 12), R2b, g, lean), 16);
            time_tile("L=3 window=12..19", make_prog(window(3, 12), R3, g, lean), 16);
            time_tile("L=3 window=12..19", make_prog(window(3, 12), R4, g, lean), 16);
            time_tile("L=3 window=12..19", make_prog(window(3, 12), R5, g, lean), 16);
        }
    }
    time_tile("L=3 window=12..19 regs low {0,1,2,3} (bank conflicts?)", make_prog(window(3, 12), Rlow, 8, 1), 16);
    time_tile("L=3 window=12..19 regs {0,3,6,9} (bank conflicts?)", make_prog(window(3, 12), Rlow2, 8, 1), 16);
    // --- reference: the warp-tile pattern ---
    for (int lean : {0, 1})
        for (int g : {0, 4, 8, 12, 16, 24, 32}) {
            const uint64_t ntiles = len >> 9;
            const unsigned grid = sms * 40;
            k_warp<<<grid, 128>>>(d, ntiles, 12, g, lean, k0, k1);
            cudaEventRecord(e0);
            for (int i = 0; i < 5; i++) k_warp<<<grid, 128>>>(d, ntiles, 12, g, lean, k0, k1);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
            printf("warp-tile window=12..15 gates=%2d lean=%d: %7.3f ms  %6.0f GB/s\n", g, lean, ms, gb / (ms * 1e-3));
        }
    cudaFree(d);
    return 0;
}
