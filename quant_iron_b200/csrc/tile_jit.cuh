// tile_jit.cuh -- circuit-specialised CTA-tile kernels (included by window.cu, inside namespace qi).
//
// k_tile (window.cu) INTERPRETS the op list of a tile pass: measured with ncu (profiles/r02_ncu_tile_*.txt) a pass executes ~190
// warp instructions per op of which ~54 are FP64 -- 36 % is loop / decode / dispatch, the rest sign flips, register swaps and
// predicate arithmetic -- and the FP64 pipe, the unit that bounds a pass with ~100 gates in it, sits at 32 %.
// This file removes the interpreter: the SAME lowered program (TileLaunch: rounds + DOps, produced by lower_tile_pass and
// checked on the CPU by tests/window_interp.py) is written out as straight-line PTX for sm_100a and handed to the driver's
// PTX assembler (cuModuleLoadDataEx; no nvcc / NVRTC / files at run time):
//   * every amplitude component is one virtual f64 register; an op is exactly its FP64 instructions, in place;
//   * gate coefficients are kernel PARAMETERS (ld.param with an immediate offset -> uniform-register / constant-bank operands),
//     so the program text depends on the STRUCTURE of the pass only: the same circuit with other angles reuses the module;
//   * X / CNOT / Toffoli on register qubits are a renaming of virtual registers (no instruction) unless a thread- or
//     tile-bit control makes them conditional (two selp per component);
//   * sign flips (Z, CZ, the negated half of a rotation) are neg.f64, which ptxas folds into the operand modifiers of the
//     next DFMA / DMUL (measured on SASS: no instruction);
//   * control predicates are compare-and-branch around the op (bra.uni for tile-uniform ones).
// The arithmetic of every op follows run_ops_tile instruction for instruction, so results are bit-identical to k_tile
// (tests/test_gpu_parity.py compares them exactly).  Modules are cached by a hash of the text.  Policy (option "jit"):
//   0 = never (k_tile only);  1 = a pass structure seen for the first time runs on k_tile while a worker thread assembles
//   its module, later executions use it (no latency for one-shot circuits; default for states of >= jit_min_qubits local
//   qubits);  2 = assemble synchronously before the first launch (benchmarks, tests).
#pragma once
// (no includes here: window.cu includes <cuda.h> and the standard headers this file needs before it opens namespace qi)

namespace jit {

// ---- driver entry points (resolved through the runtime: the library does not link libcuda) ----------------------------
struct Driver {
    bool ok = false;
    CUresult (*ModuleLoadDataEx)(CUmodule*, const void*, unsigned, CUjit_option*, void**) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
};
static Driver& driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        d.ok = get("cuModuleLoadDataEx", (void**)&d.ModuleLoadDataEx) && get("cuModuleGetFunction", (void**)&d.ModuleGetFunction) &&
               get("cuLaunchKernel", (void**)&d.LaunchKernel) && get("cuModuleUnload", (void**)&d.ModuleUnload) &&
               get("cuFuncSetAttribute", (void**)&d.FuncSetAttribute);
        if (!d.ok) cudaGetLastError();
    });
    return d;
}

// ---- PTX writer -----------------------------------------------------------------------------------------------------
struct Gen {
    std::string s;
    std::vector<double> coef;      // coefficient i lives at p_c + 8 i and in register %c<i>
    int ax[16], ay[16];            // slot -> virtual amplitude register (%a<k>) of its real / imaginary part
    int na = 32, ng = 0, nlab = 0;
    const amp_t* arena = nullptr;
    bool dry = false;              // collect the coefficients only (the module of this structure exists): no text
    bool in_region = false;        // between begin_region and end_region: the code runs under a branch
    double weight = 1.0;           // fraction of the threads inside the current control region
    double fp64 = 0.0;             // FP64 instructions per thread and tile, weighted by the control regions they sit in

    void emit(const char* fmt, ...) {
        if ((fmt[0] == 'f' && fmt[1] == 'm') || (fmt[0] == 'm' && fmt[1] == 'u' && fmt[3] == '.' && fmt[4] == 'f') || (fmt[0] == 'd' && fmt[1] == 'i')) fp64 += weight;
        if (dry) return;
        char buf[320];
        va_list ap;
        va_start(ap, fmt);
        int n = vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        s.append("  ");
        s.append(buf, (size_t)std::min<int>(n, (int)sizeof(buf) - 1));
        s.push_back('\n');
    }
    void canonical() { for (int i = 0; i < 16; i++) { ax[i] = 2 * i; ay[i] = 2 * i + 1; } }
    int c(double v) {              // a coefficient register
        const int i = (int)coef.size();
        coef.push_back(v);
        emit("ld.param.f64 %%c%d, [p_c+%d];", i, 8 * i);
        return i;
    }
    int g() { return ng++; }       // a scratch f64 register %g<k>
    int fresh() { return na++; }

    // ---- control predicates: branch around the op when they fail -------------------------------------------------
    int begin_region(const DOp& d) {
        if (!d.c_tile && !d.c_lane) return -1;
        const int L = nlab++;
        in_region = true;
        weight = std::ldexp(1.0, -(__builtin_popcountll(d.c_tile) + __builtin_popcount(d.c_lane)));
        if (d.c_tile) {
            emit("and.b64 %%rdx, %%tile, %llu;", (unsigned long long)d.c_tile);
            emit("setp.ne.u64 %%pq, %%rdx, %llu;", (unsigned long long)d.c_tval);
            emit("@%%pq bra.uni LS%d;", L);
        }
        if (d.c_lane) {
            emit("and.b32 %%rx, %%t, %u;", d.c_lane);
            emit("setp.ne.u32 %%pq, %%rx, %u;", (unsigned)d.c_lval);
            emit("@%%pq bra LS%d;", L);
        }
        return L;
    }
    void end_region(int L) { weight = 1.0; in_region = false; if (L >= 0 && !dry) { s.append("LS"); s.append(std::to_string(L)); s.append(":\n"); } }
    // %pq = the op applies (no branch)
    void on_pred(const DOp& d) {
        if (d.c_tile) {
            emit("and.b64 %%rdx, %%tile, %llu;", (unsigned long long)d.c_tile);
            emit("setp.eq.u64 %%pq, %%rdx, %llu;", (unsigned long long)d.c_tval);
        }
        if (d.c_lane) {
            emit("and.b32 %%rx, %%t, %u;", d.c_lane);
            if (d.c_tile) emit("setp.eq.and.u32 %%pq, %%rx, %u, %%pq;", (unsigned)d.c_lval);
            else emit("setp.eq.u32 %%pq, %%rx, %u;", (unsigned)d.c_lval);
        }
    }

    // ---- ops (same arithmetic, in the same order per register, as run_ops_tile) -----------------------------------
    void op_reall(const DOp& d) {
        const int B = (int)d.tpos - kLaneQubits;
        const int L = begin_region(d);
        const int k0 = c(d.m[0]), k1 = c(d.m[1]), k2 = c(d.m[2]), k3 = c(d.m[3]);
        int s0[8], s1[8];
        for (int p = 0; p < 8; p++) { s0[p] = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1)); s1[p] = s0[p] | (1 << B); }
        for (int p = 0; p < 8; p++) { emit("mul.f64 %%a%d, %%a%d, %%c%d;", ax[s0[p]], ax[s0[p]], k0); emit("mul.f64 %%a%d, %%a%d, %%c%d;", ay[s0[p]], ay[s0[p]], k0); }
        for (int p = 0; p < 8; p++) {
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", ax[s0[p]], k1, ax[s1[p]], ax[s0[p]]);
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", ay[s0[p]], k1, ay[s1[p]], ay[s0[p]]);
        }
        for (int p = 0; p < 8; p++) { emit("mul.f64 %%a%d, %%a%d, %%c%d;", ax[s1[p]], ax[s1[p]], k3); emit("mul.f64 %%a%d, %%a%d, %%c%d;", ay[s1[p]], ay[s1[p]], k3); }
        for (int p = 0; p < 8; p++) {
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", ax[s1[p]], k2, ax[s0[p]], ax[s1[p]]);
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", ay[s1[p]], k2, ay[s0[p]], ay[s1[p]]);
        }
        end_region(L);
    }
    void op_rxl(const DOp& d) {
        const int B = (int)d.tpos - kLaneQubits;
        const int L = begin_region(d);
        const int k0 = c(d.m[0]), k1 = c(d.m[1]), k2 = c(d.m[2]), k3 = c(d.m[3]);
        const int n1 = g(), n3 = g();
        emit("neg.f64 %%g%d, %%c%d;", n1, k1);
        emit("neg.f64 %%g%d, %%c%d;", n3, k3);
        int s0[8], s1[8];
        for (int p = 0; p < 8; p++) { s0[p] = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1)); s1[p] = s0[p] | (1 << B); }
        for (int p = 0; p < 8; p++) { emit("mul.f64 %%a%d, %%a%d, %%c%d;", ax[s0[p]], ax[s0[p]], k0); emit("mul.f64 %%a%d, %%a%d, %%c%d;", ay[s0[p]], ay[s0[p]], k0); }
        for (int p = 0; p < 8; p++) {
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", ax[s0[p]], k1, ay[s1[p]], ax[s0[p]]);
            emit("fma.rn.f64 %%a%d, %%g%d, %%a%d, %%a%d;", ay[s0[p]], n1, ax[s1[p]], ay[s0[p]]);
        }
        for (int p = 0; p < 8; p++) { emit("mul.f64 %%a%d, %%a%d, %%c%d;", ax[s1[p]], ax[s1[p]], k2); emit("mul.f64 %%a%d, %%a%d, %%c%d;", ay[s1[p]], ay[s1[p]], k2); }
        for (int p = 0; p < 8; p++) {
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", ax[s1[p]], k3, ay[s0[p]], ax[s1[p]]);
            emit("fma.rn.f64 %%a%d, %%g%d, %%a%d, %%a%d;", ay[s1[p]], n3, ax[s0[p]], ay[s1[p]]);
        }
        end_region(L);
    }
    // unit forms (never under a control): results go to fresh registers, the slots are renamed
    void op_realu(const DOp& d, bool minus) {
        const int B = (int)d.tpos - kLaneQubits;
        const int p = c(d.m[0]), q = c(d.m[1]);
        for (int i = 0; i < 8; i++) {
            const int s0 = ((i >> B) << (B + 1)) | (i & ((1 << B) - 1)), s1 = s0 | (1 << B);
            const int nx0 = fresh(), ny0 = fresh(), nx1 = fresh(), ny1 = fresh();
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", nx0, p, ax[s1], ax[s0]);
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", ny0, p, ay[s1], ay[s0]);
            if (minus) {
                emit("neg.f64 %%a%d, %%a%d;", ax[s1], ax[s1]);
                emit("neg.f64 %%a%d, %%a%d;", ay[s1], ay[s1]);
            }
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", nx1, q, ax[s0], ax[s1]);
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", ny1, q, ay[s0], ay[s1]);
            ax[s0] = nx0; ay[s0] = ny0; ax[s1] = nx1; ay[s1] = ny1;
        }
    }
    void op_rxu(const DOp& d) {
        const int B = (int)d.tpos - kLaneQubits;
        const int t = c(d.m[0]);
        const int nt = g();
        emit("neg.f64 %%g%d, %%c%d;", nt, t);
        for (int i = 0; i < 8; i++) {
            const int s0 = ((i >> B) << (B + 1)) | (i & ((1 << B) - 1)), s1 = s0 | (1 << B);
            const int nx0 = fresh(), ny0 = fresh(), nx1 = fresh(), ny1 = fresh();
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", nx0, t, ay[s1], ax[s0]);
            emit("fma.rn.f64 %%a%d, %%g%d, %%a%d, %%a%d;", ny0, nt, ax[s1], ay[s0]);
            emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%a%d;", nx1, t, ay[s0], ax[s1]);
            emit("fma.rn.f64 %%a%d, %%g%d, %%a%d, %%a%d;", ny1, nt, ax[s0], ay[s1]);
            ax[s0] = nx0; ay[s0] = ny0; ax[s1] = nx1; ay[s1] = ny1;
        }
    }
    void op_x(const DOp& d) {
        const int B = (int)d.tpos - kLaneQubits;
        const bool cond = d.c_tile || d.c_lane;
        if (cond) on_pred(d);
        for (int p = 0; p < 8; p++) {
            const int s0 = ((p >> B) << (B + 1)) | (p & ((1 << B) - 1)), s1 = s0 | (1 << B);
            if (!((d.c_reg >> s0) & 1u)) continue;       // register-bit controls (never on bit B): slot mask
            if (!cond) { std::swap(ax[s0], ax[s1]); std::swap(ay[s0], ay[s1]); continue; }
            const int nx0 = fresh(), nx1 = fresh(), ny0 = fresh(), ny1 = fresh();
            emit("selp.f64 %%a%d, %%a%d, %%a%d, %%pq;", nx0, ax[s1], ax[s0]);
            emit("selp.f64 %%a%d, %%a%d, %%a%d, %%pq;", nx1, ax[s0], ax[s1]);
            emit("selp.f64 %%a%d, %%a%d, %%a%d, %%pq;", ny0, ay[s1], ay[s0]);
            emit("selp.f64 %%a%d, %%a%d, %%a%d, %%pq;", ny1, ay[s0], ay[s1]);
            ax[s0] = nx0; ax[s1] = nx1; ay[s0] = ny0; ay[s1] = ny1;
        }
    }
    // x <- -x by an xor on the high word (ALU pipe).  Used wherever ptxas cannot fold a neg.f64 into an operand modifier -- under
    // a branch or a predicate it becomes a DADD on the FP64 pipe, the pipe these modules are bound by (12 % of the FP64
    // instructions of the heaviest benchmark pass were such DADDs)
    void flip_const(int reg) {
        emit("mov.b64 {%%rlo, %%rhi}, %%a%d;", reg);
        emit("xor.b32 %%rhi, %%rhi, 0x80000000;");
        emit("mov.b64 %%a%d, {%%rlo, %%rhi};", reg);
    }
    void neg_slot(int sl) {
        if (in_region) { flip_const(ax[sl]); flip_const(ay[sl]); return; }
        emit("neg.f64 %%a%d, %%a%d;", ax[sl], ax[sl]);
        emit("neg.f64 %%a%d, %%a%d;", ay[sl], ay[sl]);
    }
    // a <- e^{i phi} a as three shears with (nt, s) in registers named by `nt`, `sn` (printf patterns "%%c12" / "%%g3")
    void shear(int sl, const char* nt, const char* sn) {
        emit("fma.rn.f64 %%a%d, %s, %%a%d, %%a%d;", ax[sl], nt, ay[sl], ax[sl]);
        emit("fma.rn.f64 %%a%d, %s, %%a%d, %%a%d;", ay[sl], sn, ax[sl], ay[sl]);
        emit("fma.rn.f64 %%a%d, %s, %%a%d, %%a%d;", ax[sl], nt, ay[sl], ax[sl]);
    }
    static bool rot_neg(double nt) { uint64_t b; memcpy(&b, &nt, 8); return (b & 1ull) != 0; }
    void rot_static(int sl, double nt, double sn, int knt, int ksn) {
        char a[24], b[24];
        snprintf(a, sizeof(a), "%%c%d", knt);
        snprintf(b, sizeof(b), "%%c%d", ksn);
        (void)sn;
        if (rot_neg(nt)) neg_slot(sl);
        shear(sl, a, b);
    }
    void op_neg(const DOp& d) {
        if (d.c_tile || d.c_lane) {        // controlled sign flip (CZ with a thread- or tile-bit control): no branch, a runtime sign mask
            on_pred(d);
            emit("selp.b32 %%rm, 0x80000000, 0, %%pq;");
            for (int sl = 0; sl < 16; sl++) if ((d.c_reg >> sl) & 1u) { flip_runtime(ax[sl]); flip_runtime(ay[sl]); }
            return;
        }
        for (int sl = 0; sl < 16; sl++) if ((d.c_reg >> sl) & 1u) neg_slot(sl);
    }
    void op_diag(const DOp& d) {
        const int L = begin_region(d);
        const int knt = c(d.m[2]), ksn = c(d.m[3]);
        for (int sl = 0; sl < 16; sl++) if ((d.c_reg >> sl) & 1u) rot_static(sl, d.m[2], d.m[3], knt, ksn);
        end_region(L);
    }
    void op_rz(const DOp& d) {
        const int L = begin_region(d);
        const int k0n = c(d.m[4]), k0s = c(d.m[5]), k1n = c(d.m[6]), k1s = c(d.m[7]);
        if (!d.t_tile && !d.t_lane) {
            for (int sl = 0; sl < 16; sl++) {
                if (!((d.c_reg >> sl) & 1u)) continue;
                if (sl & d.t_reg) rot_static(sl, d.m[6], d.m[7], k1n, k1s);
                else rot_static(sl, d.m[4], d.m[5], k0n, k0s);
            }
        } else {
            // target on a thread / tile bit: %pt = the target bit is set -> (nt, s) of r1, else r0
            if (d.t_tile) { emit("and.b64 %%rdx, %%tile, %llu;", (unsigned long long)d.t_tile); emit("setp.ne.u64 %%pt, %%rdx, 0;"); }
            if (d.t_lane) {
                emit("and.b32 %%rx, %%t, %u;", d.t_lane);
                if (d.t_tile) emit("setp.ne.or.u32 %%pt, %%rx, 0, %%pt;");
                else emit("setp.ne.u32 %%pt, %%rx, 0;");
            }
            const int gn = g(), gs = g();
            emit("selp.f64 %%g%d, %%c%d, %%c%d, %%pt;", gn, k1n, k0n);
            emit("selp.f64 %%g%d, %%c%d, %%c%d, %%pt;", gs, k1s, k0s);
            const bool n0 = rot_neg(d.m[4]), n1 = rot_neg(d.m[6]);
            char a[24], b[24];
            snprintf(a, sizeof(a), "%%g%d", gn);
            snprintf(b, sizeof(b), "%%g%d", gs);
            for (int sl = 0; sl < 16; sl++) {
                if (!((d.c_reg >> sl) & 1u)) continue;
                if (sl & d.t_reg) { rot_static(sl, d.m[6], d.m[7], k1n, k1s); continue; }
                if (n0 && n1) neg_slot(sl);
                else if (n1 || n0) {
                    emit(n1 ? "selp.b32 %%rm, 0x80000000, 0, %%pt;" : "selp.b32 %%rm, 0, 0x80000000, %%pt;");
                    flip_runtime(ax[sl]);
                    flip_runtime(ay[sl]);
                }
                shear(sl, a, b);
            }
        }
        end_region(L);
    }
    void op_scale(const DOp& d) {
        const int L = begin_region(d);
        const int k0 = c(d.m[0]);
        if (d.m[1] == 0.0) {
            for (int sl = 0; sl < 16; sl++) { emit("mul.f64 %%a%d, %%a%d, %%c%d;", ax[sl], ax[sl], k0); emit("mul.f64 %%a%d, %%a%d, %%c%d;", ay[sl], ay[sl], k0); }
        } else {                   // complex scalar: the same four instructions per amplitude as run_ops_tile
            const int k1 = c(d.m[1]);
            for (int sl = 0; sl < 16; sl++) {
                const int t0 = g(), t1 = g();
                emit("mul.f64 %%g%d, %%c%d, %%a%d;", t0, k1, ay[sl]);
                emit("mul.f64 %%g%d, %%c%d, %%a%d;", t1, k1, ax[sl]);
                emit("neg.f64 %%g%d, %%g%d;", t0, t0);
                emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%g%d;", ax[sl], k0, ax[sl], t0);
                emit("fma.rn.f64 %%a%d, %%c%d, %%a%d, %%g%d;", ay[sl], k0, ay[sl], t1);
            }
        }
        end_region(L);
    }
    // x <- -x where the runtime mask %rm is 0x80000000 (flip_sign of run_ops_tile)
    void flip_runtime(int reg) {
        emit("mov.b64 {%%rlo, %%rhi}, %%a%d;", reg);
        emit("xor.b32 %%rhi, %%rhi, %%rm;");
        emit("mov.b64 %%a%d, {%%rlo, %%rhi};", reg);
    }
    void op_table(const DOp& d) {
        constexpr int NT = kTileThreads, S = 16;
        int L = begin_region(d);
        if (d.hub_cls == CLS_TILE || d.hub_cls == CLS_LANE) {
            if (L < 0) L = nlab++;
            weight *= 0.5;
            if (d.hub_cls == CLS_TILE) {
                emit("and.b64 %%rdx, %%tile, %llu;", 1ull << d.hub_bit);
                emit("setp.eq.u64 %%pq, %%rdx, 0;");
                emit("@%%pq bra.uni LS%d;", L);
            } else {
                emit("and.b32 %%rx, %%t, %u;", 1u << d.hub_bit);
                emit("setp.eq.u32 %%pq, %%rx, 0;");
                emit("@%%pq bra LS%d;", L);
            }
        }
        long long off;
        memcpy(&off, &d.m[0], 8);
        emit("mad.wide.u32 %%rdy, %%t, 16, %%ptab;");
        emit("ld.global.nc.v2.f64 {%%fx, %%fy}, [%%rdy+%lld];", 16ll * off);
        for (int k = 0; k < (int)d.nchunks; k++) {
            if (k) emit("shr.u64 %%rdx, %%tile, %d;", 8 * k);
            emit("and.b64 %%rdx, %s, 255;", k ? "%rdx" : "%tile");
            emit("shl.b64 %%rdx, %%rdx, 4;");
            emit("add.u64 %%rdx, %%rdx, %%ptab;");
            emit("ld.global.nc.v2.f64 {%%gx, %%gy}, [%%rdx+%lld];", 16ll * (off + NT + S + 256 * k));
            // cmul as nvcc contracts it in run_ops_tile: (fx gx - fy gy, fx gy + fy gx)
            emit("mul.f64 %%h0, %%fy, %%gy;");
            emit("mul.f64 %%h1, %%fy, %%gx;");
            emit("neg.f64 %%h0, %%h0;");
            emit("fma.rn.f64 %%h0, %%fx, %%gx, %%h0;");
            emit("fma.rn.f64 %%h1, %%fx, %%gy, %%h1;");
            emit("mov.f64 %%fx, %%h0;");
            emit("mov.f64 %%fy, %%h1;");
        }
        // the thread's factor as a plain complex product, instruction for instruction what run_ops_tile does
        const uint32_t hub_slot = d.hub_cls == CLS_REG ? (1u << d.hub_bit) : 0u;
        for (int sl = 0; sl < S; sl++) {
            if ((sl & hub_slot) != hub_slot) continue;
            const int t0 = g(), t1 = g();
            emit("mul.f64 %%g%d, %%fy, %%a%d;", t0, ay[sl]);
            emit("mul.f64 %%g%d, %%fy, %%a%d;", t1, ax[sl]);
            emit("neg.f64 %%g%d, %%g%d;", t0, t0);
            emit("fma.rn.f64 %%a%d, %%fx, %%a%d, %%g%d;", ax[sl], ax[sl], t0);
            emit("fma.rn.f64 %%a%d, %%fx, %%a%d, %%g%d;", ay[sl], ay[sl], t1);
        }
        if (d.has_reg) {            // slot factors: packed rotations behind the chunk tables (host arena -> coefficients)
            const amp_t* sl_t = arena + off + NT + S + 256 * (long long)d.nchunks;
            for (int sl = 0; sl < S; sl++) {
                if ((sl & hub_slot) != hub_slot) continue;
                const int knt = c(sl_t[sl].x), ksn = c(sl_t[sl].y);
                rot_static(sl, sl_t[sl].x, sl_t[sl].y, knt, ksn);
            }
        }
        end_region(L);
    }
    int op(const DOp& d) {
        switch (d.kind) {
            case WK_REALL: op_reall(d); return QI_OK;
            case WK_RXL: op_rxl(d); return QI_OK;
            case WK_REALUP: case WK_REALUM:
            case WK_RXU:
                if (d.c_tile || d.c_lane || (d.c_reg & 0xffffu) != 0xffffu) return fail(QI_ERR_UNKNOWN, d.kind, 0, "internal: unit-form tile op under a control");
                if (d.kind == WK_RXU) op_rxu(d); else op_realu(d, d.kind == WK_REALUM);
                return QI_OK;
            case WK_X: op_x(d); return QI_OK;
            case WK_NEG: op_neg(d); return QI_OK;
            case WK_DIAG: op_diag(d); return QI_OK;
            case WK_RZ: op_rz(d); return QI_OK;
            case WK_SCALE: op_scale(d); return QI_OK;
            case WK_TABLE: op_table(d); return QI_OK;
            default: return fail(QI_ERR_UNKNOWN, d.kind, 0, "internal: tile op kind without a PTX form");
        }
    }
};

static uint32_t swz_of_regs(const TileRoundHost& h, int sl) {
    uint32_t o = 0;
    for (int k = 0; k < 4; k++) if ((sl >> k) & 1) o |= 1u << h.regs[k];
    return tile_swz(o);
}

// the whole pass as PTX; `coef` receives the parameter block that goes with this text
// (text == nullptr: only the coefficients, in the order the text of this structure reads them)
// `groups`: tiles a CTA works on side by side (128 threads each).  The module is straight-line code far larger than the 32 KB
// L1.5 instruction cache, so every group of warps that drifts apart is one more instruction stream the SM pulls from L2
// (ncu: `no_instruction` was the top stall with four independent 128-thread CTAs per SM); groups of one CTA meet at every
// regroup barrier and share one stream.
// `stage` (groups == 1): the CTA's NEXT tile is brought into a second 32 KiB shared-memory buffer by the async proxy while the
// current one is computed -- 64 bulk copies of one 512-byte row each (cp.async.bulk ... mbarrier::complete_tx::bytes), issued
// by threads 0..63 and counted by one mbarrier -- and the first round reads its registers from that buffer instead of from
// global memory.  No registers are held across the copy (a register prefetch would need 64 more per thread), the loads
// leave the critical path of the tile, and the row addresses need not coalesce per warp.  3 CTAs per SM (2 x 32 KiB each).
static int generate(const TileLaunch& tl, const amp_t* arena, int ctas_per_sm, int groups, int prefetch, int stage, std::string* text, std::vector<double>* coef, double* fp64_per_thread) {
    Gen gn;
    gn.arena = arena;
    gn.dry = text == nullptr;
    const int nr = (int)tl.rounds.size();
    // ---- prologue: thread constants ----
    gn.emit("mov.u32 %%rx, %%tid.x;");
    gn.emit("and.b32 %%t, %%rx, %d;", kTileThreads - 1);
    gn.emit("shr.u32 %%grp, %%rx, %d;", kTileThrBits);
    gn.emit("ld.param.u64 %%pa, [p_a];");
    gn.emit("cvta.to.global.u64 %%pa, %%pa;");
    gn.emit("ld.param.u64 %%ptab, [p_tab];");
    gn.emit("cvta.to.global.u64 %%ptab, %%ptab;");
    gn.emit("ld.param.u64 %%ntiles, [p_ntiles];");
    gn.emit("mov.u32 %%smb, dsm;");
    gn.emit("add.u32 %%lbsa, %%smb, %d;", (groups + (stage ? 1 : 0)) * (int)(sizeof(amp_t) << kTileBits));     // W tables (one per group: W holds the group's own image address) behind the tile images
    gn.emit("mad.lo.u32 %%lbsa, %%grp, %d, %%lbsa;", 512 * std::max(nr, 1));
    gn.emit("mad.lo.u32 %%smb, %%grp, %d, %%smb;", (int)(sizeof(amp_t) << kTileBits));
    gn.emit("shl.b32 %%rx, %%t, 2;");
    gn.emit("add.u32 %%lbsa, %%lbsa, %%rx;");
    for (int r = 0; r < nr; r++) {           // W_r(t) = &sm[swz(b_r(t))]: the XOR swizzle only touches the three low index bits
        gn.emit("mov.u32 %%ry, 0;");
        for (int k = 0; k < kTileThrBits; k++) {
            gn.emit("bfe.u32 %%rx, %%t, %d, 1;", k);
            gn.emit("shl.b32 %%rx, %%rx, %d;", tl.rounds[r].thr[k]);
            gn.emit("or.b32 %%ry, %%ry, %%rx;");
        }
        // tile_swz(j) = j ^ ((j >> 3) & 7) ^ ((j >> 6) & 7) ^ ((j >> 9) & 3)
        gn.emit("bfe.u32 %%rx, %%ry, 3, 3;");
        gn.emit("bfe.u32 %%rz, %%ry, 6, 3;");
        gn.emit("xor.b32 %%rx, %%rx, %%rz;");
        gn.emit("bfe.u32 %%rz, %%ry, 9, 2;");
        gn.emit("xor.b32 %%rx, %%rx, %%rz;");
        gn.emit("xor.b32 %%ry, %%ry, %%rx;");
        gn.emit("shl.b32 %%ry, %%ry, 4;");
        gn.emit("add.u32 %%ry, %%ry, %%smb;");
        gn.emit("st.shared.u32 [%%lbsa+%d], %%ry;", 512 * r);
    }
    auto thread_offset = [&](const char* dst, const TileRoundHost& h, const int* pos) {      // sum_k bit_k(t) << pos[thr[k]]
        gn.emit("mov.u64 %s, 0;", dst);
        for (int k = 0; k < kTileThrBits; k++) {
            gn.emit("bfe.u32 %%rx, %%t, %d, 1;", k);
            gn.emit("cvt.u64.u32 %%rdx, %%rx;");
            gn.emit("shl.b64 %%rdx, %%rdx, %d;", pos[h.thr[k]]);
            gn.emit("or.b64 %s, %s, %%rdx;", dst, dst);
        }
    };
    thread_offset("%gin", tl.rounds.front(), tl.tile_qubits);
    thread_offset("%gout", tl.rounds.back(), tl.tile_out);
    const int kImage = (int)(sizeof(amp_t) << kTileBits);
    auto insert_zeros = [&](const char* reg, const char* tmp) {          // reg <- reg with zero bits inserted at the tile's positions
        for (int j = 0; j < kTileBits; j++) {
            const int p = tl.tile_qubits[j];
            gn.emit("shr.u64 %s, %s, %d;", tmp, reg, p);
            gn.emit("shl.b64 %s, %s, %d;", tmp, tmp, p + 1);
            gn.emit("and.b64 %s, %s, %llu;", reg, reg, (1ull << p) - 1ull);
            gn.emit("or.b64 %s, %s, %s;", reg, reg, tmp);
        }
    };
    // stage: bulk copies of the tile whose compact index is in %rdx (clobbers %rdx, %rdy)
    auto issue_stage = [&]() {
        insert_zeros("%rdx", "%rdy");
        gn.emit("@%%pt0 mbarrier.arrive.expect_tx.shared::cta.b64 _, [%%mbar], %d;", kImage);
        gn.emit("add.u64 %%rdx, %%rdx, %%rowoff;");
        gn.emit("shl.b64 %%rdx, %%rdx, 4;");
        gn.emit("add.u64 %%rdx, %%rdx, %%pa;");
        gn.emit("@%%pld cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%%rowdst], [%%rdx], 512, [%%mbar];");
    };
    if (stage) {
        const TileRoundHost& io = tl.rounds.front();
        gn.emit("mov.u32 %%rawt, dsm;");
        gn.emit("add.u32 %%mbar, %%rawt, %d;", 2 * kImage + 512 * std::max(nr, 1));
        gn.emit("add.u32 %%rawt, %%rawt, %d;", kImage);
        gn.emit("shl.b32 %%rx, %%t, 9;");
        gn.emit("add.u32 %%rowdst, %%rawt, %%rx;");                     // row t of the raw tile (threads 0..63)
        gn.emit("mov.u32 %%ry, 0;");                                    // this thread's tile-local base under the load layout
        for (int k = 0; k < kTileThrBits; k++) {
            gn.emit("bfe.u32 %%rx, %%t, %d, 1;", k);
            gn.emit("shl.b32 %%rx, %%rx, %d;", io.thr[k] + 4);
            gn.emit("or.b32 %%ry, %%ry, %%rx;");
        }
        gn.emit("add.u32 %%rawt, %%rawt, %%ry;");
        gn.emit("mov.u64 %%rowoff, 0;");                                // global index offset of row t: window bit k of t -> position tile_qubits[5 + k]
        for (int k = 0; k < kTileBits - kLaneQubits; k++) {
            gn.emit("bfe.u32 %%rx, %%t, %d, 1;", k);
            gn.emit("cvt.u64.u32 %%rdx, %%rx;");
            gn.emit("shl.b64 %%rdx, %%rdx, %d;", tl.tile_qubits[kLaneQubits + k]);
            gn.emit("or.b64 %%rowoff, %%rowoff, %%rdx;");
        }
        gn.emit("setp.lt.u32 %%pld, %%t, %d;", 1 << (kTileBits - kLaneQubits));
        gn.emit("setp.eq.u32 %%pt0, %%t, 0;");
        gn.emit("@%%pt0 mbarrier.init.shared::cta.b64 [%%mbar], 1;");
        gn.emit("@%%pt0 fence.mbarrier_init.release.cluster;");
        gn.emit("bar.sync 0;");
        gn.emit("mov.u32 %%phase, 0;");
    }
    gn.emit("mov.u32 %%rx, %%ctaid.x;");
    gn.emit("mad.lo.u32 %%rx, %%rx, %d, %%grp;", groups);
    gn.emit("cvt.u64.u32 %%tile, %%rx;");
    gn.emit("mov.u32 %%rx, %%nctaid.x;");
    gn.emit("mul.lo.u32 %%rx, %%rx, %d;", groups);
    gn.emit("cvt.u64.u32 %%tstep, %%rx;");
    gn.emit("setp.ge.u64 %%pq, %%tile, %%ntiles;");
    gn.emit("@%%pq bra.uni LEND;");
    if (stage) {
        gn.emit("mov.u64 %%rdx, %%tile;");
        issue_stage();
    }
    if (!gn.dry) gn.s.append("LTILE:\n");
    // ---- tile base: zero bits inserted at the tile's positions (ascending) ----
    gn.emit("mov.u64 %%tb, %%tile;");
    for (int j = 0; j < kTileBits; j++) {
        const int p = tl.tile_qubits[j];
        gn.emit("shr.u64 %%rdx, %%tb, %d;", p);
        gn.emit("shl.b64 %%rdx, %%rdx, %d;", p + 1);
        gn.emit("and.b64 %%tb, %%tb, %llu;", (1ull << p) - 1ull);
        gn.emit("or.b64 %%tb, %%tb, %%rdx;");
    }
    gn.canonical();
    gn.emit("add.u64 %%rdy, %%tb, %%gin;");
    gn.emit("shl.b64 %%rdy, %%rdy, 4;");
    gn.emit("add.u64 %%rdy, %%rdy, %%pa;");
    auto slot_offset = [&](const TileRoundHost& h, const int* pos, int sl) {
        uint64_t o = 0;
        for (int k = 0; k < 4; k++) if ((sl >> k) & 1) o |= 1ull << pos[h.regs[k]];
        return o * 16ull;
    };
    if (stage) {
        // wait for this tile's 32 KiB, read the registers under the load layout, then hand the buffer to the next tile's copies
        const int LW = gn.nlab++, LN = gn.nlab++;
        if (!gn.dry) { gn.s.append("LS"); gn.s.append(std::to_string(LW)); gn.s.append(":\n"); }
        gn.emit("mbarrier.try_wait.parity.shared::cta.b64 %%pq, [%%mbar], %%phase;");
        gn.emit("@!%%pq bra LS%d;", LW);
        gn.emit("xor.b32 %%phase, %%phase, 1;");
        const TileRoundHost& io = tl.rounds.front();
        for (int sl = 0; sl < 16; sl++) {
            uint32_t o = 0;
            for (int k = 0; k < 4; k++) if ((sl >> k) & 1) o |= 1u << io.regs[k];
            gn.emit("ld.shared.v2.f64 {%%a%d, %%a%d}, [%%rawt+%u];", gn.ax[sl], gn.ay[sl], o * 16u);
        }
        gn.emit("bar.sync 0;");
        gn.emit("add.u64 %%rdx, %%tile, %%tstep;");
        gn.emit("setp.ge.u64 %%pq, %%rdx, %%ntiles;");
        gn.emit("@%%pq bra.uni LS%d;", LN);
        gn.emit("fence.proxy.async.shared::cta;");
        issue_stage();
        if (!gn.dry) { gn.s.append("LS"); gn.s.append(std::to_string(LN)); gn.s.append(":\n"); }
    } else
    for (int sl = 0; sl < 16; sl++) {
        const uint64_t o = slot_offset(tl.rounds.front(), tl.tile_qubits, sl);
        if (o < (1ull << 31)) gn.emit("ld.global.cs.v2.f64 {%%a%d, %%a%d}, [%%rdy+%llu];", gn.ax[sl], gn.ay[sl], (unsigned long long)o);
        else {
            gn.emit("add.u64 %%rdx, %%rdy, %llu;", (unsigned long long)o);
            gn.emit("ld.global.cs.v2.f64 {%%a%d, %%a%d}, [%%rdx];", gn.ax[sl], gn.ay[sl]);
        }
    }
    auto emit_prefetch = [&]() {
        // the CTA's next tile into L2 while this one is computed (no registers held): one 128-byte line per 8 lanes and slot.
        // A pass of ~40 gates is no longer far above its HBM time, and with 4 CTAs per SM the DRAM latency of the loads at
        // the head of every tile is exposed (ncu: long_scoreboard is the top stall of the unit-form modules).
        // prefetch = 1: issued before the LAST round (the whole GPU streams ~150 MB through the 126 MB L2 per tile period, so a
        // line prefetched at the head of the tile is evicted before it is used); 2: at the head of the tile.
        const int L = gn.nlab++;
        gn.emit("add.u64 %%rdx, %%tile, %%tstep;");
        gn.emit("setp.ge.u64 %%pq, %%rdx, %%ntiles;");
        gn.emit("@%%pq bra.uni LS%d;", L);
        gn.emit("and.b32 %%rx, %%t, 7;");
        gn.emit("setp.ne.u32 %%pq, %%rx, 0;");
        gn.emit("@%%pq bra LS%d;", L);
        for (int j = 0; j < kTileBits; j++) {
            const int p = tl.tile_qubits[j];
            gn.emit("shr.u64 %%rdy, %%rdx, %d;", p);
            gn.emit("shl.b64 %%rdy, %%rdy, %d;", p + 1);
            gn.emit("and.b64 %%rdx, %%rdx, %llu;", (1ull << p) - 1ull);
            gn.emit("or.b64 %%rdx, %%rdx, %%rdy;");
        }
        gn.emit("add.u64 %%rdx, %%rdx, %%gin;");
        gn.emit("shl.b64 %%rdx, %%rdx, 4;");
        gn.emit("add.u64 %%rdx, %%rdx, %%pa;");
        for (int sl = 0; sl < 16; sl++) {
            const uint64_t o = slot_offset(tl.rounds.front(), tl.tile_qubits, sl);
            if (o < (1ull << 31)) gn.emit("prefetch.global.L2 [%%rdx+%llu];", (unsigned long long)o);
            else {
                gn.emit("add.u64 %%rdy, %%rdx, %llu;", (unsigned long long)o);
                gn.emit("prefetch.global.L2 [%%rdy];");
            }
        }
        gn.end_region(L);
    };
    if (stage) prefetch = 0;
    if (prefetch == 2 || (prefetch == 1 && nr == 1)) emit_prefetch();
    for (int r = 0; r < nr; r++) {
        const TileRoundHost& cur = tl.rounds[r];
        if (r > 0) {              // regroup through the shared-memory image of the tile (one barrier: see k_tile)
            const TileRoundHost& prev = tl.rounds[r - 1];
            gn.emit("ld.shared.u32 %%wb, [%%lbsa+%d];", 512 * (r - 1));
            gn.emit("ld.shared.u32 %%rb, [%%lbsa+%d];", 512 * r);
            for (int sl = 0; sl < 16; sl++) {
                const uint32_t j = swz_of_regs(prev, sl), lo = (j & 7u) << 4, hi = (j & ~7u) << 4;
                if (lo) { gn.emit("xor.b32 %%rx, %%wb, %u;", lo); gn.emit("st.shared.v2.f64 [%%rx+%u], {%%a%d, %%a%d};", hi, gn.ax[sl], gn.ay[sl]); }
                else gn.emit("st.shared.v2.f64 [%%wb+%u], {%%a%d, %%a%d};", hi, gn.ax[sl], gn.ay[sl]);
            }
            gn.emit("bar.sync 0;");
            gn.canonical();
            for (int sl = 0; sl < 16; sl++) {
                const uint32_t j = swz_of_regs(cur, sl), lo = (j & 7u) << 4, hi = (j & ~7u) << 4;
                if (lo) { gn.emit("xor.b32 %%rx, %%rb, %u;", lo); gn.emit("ld.shared.v2.f64 {%%a%d, %%a%d}, [%%rx+%u];", gn.ax[sl], gn.ay[sl], hi); }
                else gn.emit("ld.shared.v2.f64 {%%a%d, %%a%d}, [%%rb+%u];", gn.ax[sl], gn.ay[sl], hi);
            }
        }
        if (prefetch == 1 && nr > 1 && r == nr - 1) emit_prefetch();
        for (size_t o = 0; o < cur.nops; o++) QI_TRY(gn.op(tl.dops[cur.first_op + o]));
    }
    gn.emit("add.u64 %%rdy, %%tb, %%gout;");
    gn.emit("shl.b64 %%rdy, %%rdy, 4;");
    gn.emit("add.u64 %%rdy, %%rdy, %%pa;");
    for (int sl = 0; sl < 16; sl++) {
        const uint64_t o = slot_offset(tl.rounds.back(), tl.tile_out, sl);
        if (o < (1ull << 31)) gn.emit("st.global.cs.v2.f64 [%%rdy+%llu], {%%a%d, %%a%d};", (unsigned long long)o, gn.ax[sl], gn.ay[sl]);
        else {
            gn.emit("add.u64 %%rdx, %%rdy, %llu;", (unsigned long long)o);
            gn.emit("st.global.cs.v2.f64 [%%rdx], {%%a%d, %%a%d};", gn.ax[sl], gn.ay[sl]);
        }
    }
    gn.emit("add.u64 %%tile, %%tile, %%tstep;");
    gn.emit("setp.lt.u64 %%pq, %%tile, %%ntiles;");
    if (nr > 1) gn.emit("bar.sync 0;");      // the next tile's first regroup must not overwrite shared memory another warp still reads
    gn.emit("@%%pq bra.uni LTILE;");
    *coef = std::move(gn.coef);
    *fp64_per_thread = gn.fp64;
    if (gn.dry) return QI_OK;
    gn.s.append("LEND:\n  ret;\n}\n");

    char head[1024];
    const size_t nb = std::max<size_t>(coef->size(), 1) * 8;
    snprintf(head, sizeof(head),
             ".version 8.6\n.target sm_100a\n.address_size 64\n\n.extern .shared .align 128 .b8 dsm[];\n\n"
             ".visible .entry qi_tile_jit(.param .u64 p_a, .param .u64 p_ntiles, .param .u64 p_tab, .param .align 16 .b8 p_c[%zu])\n"
             ".maxntid %d, 1, 1\n.minnctapersm %d\n{\n"
             "  .reg .pred %%pq, %%pt, %%pt0, %%pld;\n"
             "  .reg .b32 %%t, %%grp, %%smb, %%lbsa, %%wb, %%rb, %%rx, %%ry, %%rz, %%rm, %%rlo, %%rhi, %%rawt, %%rowdst, %%mbar, %%phase;\n"
             "  .reg .b64 %%pa, %%ptab, %%ntiles, %%tile, %%tstep, %%tb, %%gin, %%gout, %%rdx, %%rdy, %%rowoff;\n"
             "  .reg .f64 %%fx, %%fy, %%gx, %%gy, %%h0, %%h1;\n"
             "  .reg .f64 %%a<%d>;\n  .reg .f64 %%c<%zu>;\n  .reg .f64 %%g<%d>;\n"
             "",
             nb, kTileThreads * groups, std::max(1, ctas_per_sm / groups), gn.na, std::max<size_t>(coef->size(), 1), std::max(gn.ng, 1));
    text->assign(head);
    text->append(gn.s);
    return QI_OK;
}

// ---- module cache and assembler workers --------------------------------------------------------------------------------
struct Entry {
    std::atomic<int> state{0};     // 0 = queued / assembling, 1 = ready, -1 = failed (k_tile runs the pass)
    CUfunction fn = nullptr;
    size_t text_len = 0;
    int groups = 1;                // tiles per CTA (128 threads each)
    unsigned smem = 0;             // dynamic shared memory: the groups' tile images + the W table
};
static unsigned smem_bytes(int groups, int nrounds, int stage) {
    if (stage) return 2u * (unsigned)(sizeof(amp_t) << kTileBits) + 512u * (unsigned)std::max(nrounds, 1) + 16u;      // image, raw tile, W table, mbarrier
    return (unsigned)groups * ((unsigned)(sizeof(amp_t) << kTileBits) + 512u * (unsigned)std::max(nrounds, 1));
}
// a structure seen for the first time: the worker writes the text (generate) AND assembles it, so the thread that executes the
// circuit only pays for the key and for copies of the launch record and (once per circuit) the table arena
struct Job {
    Entry* e = nullptr;
    std::shared_ptr<const TileLaunch> tl;
    std::shared_ptr<const std::vector<amp_t>> arena;
    int ctas = 4, groups = 1, prefetch = 0, stage = 0;
};
struct Cache {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::unordered_map<uint64_t, Entry*> map;
    std::deque<Job> queue;
    int workers = 0, pending = 0, device = 0;
    bool stopping = false;             // process exit: queued jobs are dropped, running ones are waited for (shutdown())
    uint64_t assembled = 0, failed = 0;
    double assemble_ms = 0.0;
    double fp64_warp_instr = 0.0;      // launched on modules since the last stats reset (weighted static count x warps x tiles)
};
static const size_t kMaxModules = 4096;
static Cache& cache() { static Cache* c = new Cache; return *c; }        // leaked on purpose: worker threads may outlive exit()

struct Fnv {
    uint64_t h = 1469598103934665603ull;
    void bytes(const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } }
    void u64(uint64_t v) { bytes(&v, 8); }
};
// everything the text of a pass depends on: positions, rounds, the non-coefficient half of every op, table offsets and the
// negate flags packed into the rotations (generate() reads nothing else but coefficients)
static uint64_t structure_key(const TileLaunch& tl, const amp_t* arena, int ctas_per_sm, int groups, int prefetch, int stage) {
    Fnv f;
    f.u64((uint64_t)ctas_per_sm | ((uint64_t)groups << 8) | ((uint64_t)prefetch << 16) | ((uint64_t)(stage != 0) << 24));
    f.bytes(tl.tile_qubits, sizeof(tl.tile_qubits));
    f.bytes(tl.tile_out, sizeof(tl.tile_out));
    f.u64(tl.rounds.size());
    for (const TileRoundHost& r : tl.rounds) { f.bytes(r.regs, sizeof(r.regs)); f.bytes(r.thr, sizeof(r.thr)); f.u64(r.first_op); f.u64(r.nops); }
    for (const DOp& d : tl.dops) {
        f.bytes(&d, offsetof(DOp, m));
        if (d.kind == WK_DIAG) f.u64(Gen::rot_neg(d.m[2]));
        else if (d.kind == WK_SCALE) f.u64(d.m[1] != 0.0);
        else if (d.kind == WK_RZ) f.u64((uint64_t)Gen::rot_neg(d.m[4]) | ((uint64_t)Gen::rot_neg(d.m[6]) << 1));
        else if (d.kind == WK_TABLE) {
            long long off;
            memcpy(&off, &d.m[0], 8);
            f.u64((uint64_t)off);
            if (d.has_reg) {
                const amp_t* sl_t = arena + off + kTileThreads + 16 + 256 * (long long)d.nchunks;
                uint64_t flags = 0;
                for (int sl = 0; sl < 16; sl++) flags |= (uint64_t)Gen::rot_neg(sl_t[sl].x) << sl;
                f.u64(flags);
            }
        }
    }
    return f.h;
}

static void assemble(Entry* e, const std::string& text) {
    Driver& d = driver();
    char log[4096];
    log[0] = 0;
    CUjit_option opts[] = {CU_JIT_ERROR_LOG_BUFFER, CU_JIT_ERROR_LOG_BUFFER_SIZE_BYTES};
    void* vals[] = {(void*)log, (void*)(uintptr_t)sizeof(log)};
    CUmodule mod = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    CUresult r = d.ModuleLoadDataEx(&mod, text.c_str(), 2, opts, vals);
    if (r == CUDA_SUCCESS) r = d.ModuleGetFunction(&e->fn, mod, "qi_tile_jit");
    if (r == CUDA_SUCCESS && e->smem > 48 * 1024) r = d.FuncSetAttribute(e->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)e->smem);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    Cache& c = cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        c.assemble_ms += ms;
        if (r == CUDA_SUCCESS) c.assembled++; else c.failed++;
    }
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[qiron_b200] tile JIT: the driver rejected a generated module (CUresult %d): %s\n  -> this pass runs on the k_tile interpreter\n", (int)r, log);
        e->state.store(-1, std::memory_order_release);
    } else e->state.store(1, std::memory_order_release);
}

static void worker_main(int device) {
    cudaSetDevice(device);
    cudaFree(0);                                   // binds the primary context to this thread
    Cache& c = cache();
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> lk(c.mu);
            c.cv_work.wait(lk, [&] { return !c.queue.empty() && !c.stopping; });
            job = std::move(c.queue.front());
            c.queue.pop_front();
        }
        std::string text;
        std::vector<double> coef;
        double fp64 = 0.0;
        if (generate(*job.tl, job.arena->data(), job.ctas, job.groups, job.prefetch, job.stage, &text, &coef, &fp64) == QI_OK) {
            job.e->text_len = text.size();
            assemble(job.e, text);
        } else {
            job.e->state.store(-1, std::memory_order_release);
            std::lock_guard<std::mutex> lk(c.mu);
            c.failed++;
        }
        {
            std::lock_guard<std::mutex> lk(c.mu);
            c.pending--;
        }
        c.cv_done.notify_all();
    }
}

static void shutdown();
static Entry* find(uint64_t key) {
    Cache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.map.find(key);
    return it == c.map.end() ? nullptr : it->second;
}
// a structure seen for the first time: its text is queued for assembly.  Returns the entry (state says whether it can be launched).
static Entry* enqueue(uint64_t h, Job&& job, int device, unsigned smem) {
    Cache& c = cache();
    std::unique_lock<std::mutex> lk(c.mu);
    auto it = c.map.find(h);
    if (it != c.map.end()) return it->second;
    Entry* e = new Entry;
    e->groups = job.groups;
    e->smem = smem;
    job.e = e;
    c.map.emplace(h, e);
    // modules are never unloaded: a process that keeps producing new pass structures stops assembling at kMaxModules
    // (the interpreting kernel runs what has no module)
    if (c.stopping || c.map.size() > kMaxModules) { e->state.store(-1, std::memory_order_release); return e; }
    c.device = device;
    const int want = std::max(1, std::min(12, (int)std::thread::hardware_concurrency() - 2));
    c.queue.push_back(std::move(job));
    c.pending++;
    if (c.workers < want && c.workers < (int)c.queue.size()) {
        if (c.workers == 0) std::atexit(shutdown);
        c.workers++;
        std::thread(worker_main, device).detach();
    }
    lk.unlock();
    c.cv_work.notify_one();
    return e;
}
// atexit: the CUDA runtime tears the primary context down in its own exit handler, which runs AFTER this one (ours is registered
// later); a worker must not be inside cuModuleLoadDataEx at that point.  Queued jobs are dropped, running ones finish.
static void shutdown() {
    Cache& c = cache();
    std::unique_lock<std::mutex> lk(c.mu);
    c.stopping = true;
    c.pending -= (int)c.queue.size();
    for (Job& j : c.queue) if (j.e) j.e->state.store(-1, std::memory_order_release);
    c.queue.clear();
    c.cv_done.wait_for(lk, std::chrono::seconds(30), [&] { return c.pending <= 0; });
}
static void drain() {
    Cache& c = cache();
    std::unique_lock<std::mutex> lk(c.mu);
    c.cv_done.wait(lk, [&] { return c.pending == 0; });
}

}  // namespace jit
