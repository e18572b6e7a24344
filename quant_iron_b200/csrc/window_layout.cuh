// window_layout.cuh -- how one pass of a warp-tile kernel maps physical qubits onto lanes, register
// slots and the tile index (shared by window.cu and pauli_window.cu; host only).
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace qi {

enum { CLS_NONE = 0, CLS_LANE = 1, CLS_REG = 2, CLS_TILE = 3 };

static const int kLaneQubits = 5;    // lanes <-> physical qubits 0..4

struct Layout {
    int R;
    int nt = 32;                        // entries of the lane part of a phase table (32 lanes; k_tile rounds: 128 threads)
    std::vector<int> regs;              // sorted window qubits
    int cls[64];                        // CLS_* per physical bit
    int idx[64];                        // lane bit / slot bit / compact tile bit per physical bit
    int ntile_bits;
};

static inline Layout make_layout(const qi_state* s, std::vector<int> regs, int R) {
    Layout L;
    L.R = R;
    const int n = (int)s->n_local;
    std::sort(regs.begin(), regs.end());
    // pad the window with unused qubits (lowest free positions first: better locality)
    for (int q = kLaneQubits; q < n && (int)regs.size() < R; q++)
        if (std::find(regs.begin(), regs.end(), q) == regs.end()) regs.push_back(q);
    std::sort(regs.begin(), regs.end());
    L.regs = regs;
    int t = 0;
    for (int q = 0; q < 64; q++) {
        L.cls[q] = CLS_NONE; L.idx[q] = 0;
        if (q >= n) continue;
        auto it = std::find(regs.begin(), regs.end(), q);
        if (q < kLaneQubits) { L.cls[q] = CLS_LANE; L.idx[q] = q; }
        else if (it != regs.end()) { L.cls[q] = CLS_REG; L.idx[q] = (int)(it - regs.begin()); }
        else { L.cls[q] = CLS_TILE; L.idx[q] = t++; }
    }
    L.ntile_bits = t;
    return L;
}

static inline void split_mask(const Layout& L, uint64_t phys, uint32_t* lane, uint32_t* reg, uint64_t* tile) {
    *lane = 0; *reg = 0; *tile = 0;
    for (int q = 0; q < 64; q++) {
        if (!((phys >> q) & 1)) continue;
        if (L.cls[q] == CLS_LANE) *lane |= 1u << L.idx[q];
        else if (L.cls[q] == CLS_REG) *reg |= 1u << L.idx[q];
        else if (L.cls[q] == CLS_TILE) *tile |= 1ull << L.idx[q];
    }
}

// slot -> index offset table and zero-insert positions of a layout
template <int R>
static inline void fill_offsets(const Layout& L, BitInsert* ins, uint64_t (&off)[1 << R]) {
    *ins = make_insert(L.regs, {});
    for (int sidx = 0; sidx < (1 << R); sidx++) {
        uint64_t o = 0;
        for (int j = 0; j < R; j++) if ((sidx >> j) & 1) o |= 1ull << L.regs[j];
        off[sidx] = o;
    }
}

}  // namespace qi
