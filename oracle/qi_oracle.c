/*
 * qi_oracle.c -- CPU ORACLE for the quant-iron state-vector hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or
 * as the timed CPU baseline.  The product (quant_iron_b200/, csrc/) never
 * links, imports or falls back to it.
 *
 * What it is: a plain-C restatement of the arithmetic of the reference's CPU
 * path (LordSaumya/quant-iron v2.0.0, Rust).  The Rust crate cannot be built in
 * this image (no cargo/rustc), so each function below follows one reference
 * function and cites it as `file:line` relative to the reference root.  The
 * complex arithmetic is num-complex 0.4.6 (Cargo.lock:326), which is not
 * vendored in the reference tree; its published semantics are restated in the
 * c_* helpers: (a+bi)(c+di) = (ac-bd) + (ad+bc)i with plain f64 mul/add, no
 * FMA (build with -ffp-contract=off), exp/cosh/sinh by the textbook formulas.
 *
 * Parity pinning: tests/test_ref_ported_*.py run the reference's own
 * known-answer tests (src/tests/*.rs, cited per test) against this oracle.
 * Items the reference never tests (Subroutine::qft, measurement statistics,
 * Heisenberg term order) are pinned by closed forms instead; see DESIGN.md.
 *
 * Two flavours:
 *   orc_*            in-place, same arithmetic and operation order as the
 *                    reference's sequential branch (the parity oracle);
 *   orc_gate_faithful  same passes as the reference's rayon branch
 *                    (clone + parallel (index,value) updates with one small heap
 *                    allocation per pair + serial scatter): the CPU baseline
 *                    that is timed as "the reference's CPU path".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

/* gate kinds (oracle-private numbering; mirrors include/qiron_b200.h by value) */
enum {
    G_H = 1, G_X = 2, G_Y = 3, G_Z = 4, G_I = 5, G_S = 6, G_SDG = 7, G_T = 8,
    G_TDG = 9, G_P = 10, G_RX = 11, G_RY = 12, G_RZ = 13, G_U2 = 14,
    G_CNOT = 15, G_SWAP = 16, G_TOFFOLI = 17, G_MATCHGATE = 18
};

/* ---- num-complex 0.4.6 semantics ------------------------------------ */
static inline cplx c_mk(double re, double im) { cplx z = { re, im }; return z; }
static inline cplx c_add(cplx a, cplx b) { return c_mk(a.re + b.re, a.im + b.im); }
static inline cplx c_sub(cplx a, cplx b) { return c_mk(a.re - b.re, a.im - b.im); }
static inline cplx c_neg(cplx a) { return c_mk(-a.re, -a.im); }
static inline cplx c_mul(cplx a, cplx b) {
    return c_mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
static inline cplx c_scale(double s, cplx a) { return c_mk(s * a.re, s * a.im); } /* f64 * Complex */
static inline cplx c_scale_r(cplx a, double s) { return c_mk(a.re * s, a.im * s); } /* Complex * f64 */
static inline cplx c_divr(cplx a, double s) { return c_mk(a.re / s, a.im / s); }
static inline cplx c_conj(cplx a) { return c_mk(a.re, -a.im); }
static inline double c_norm_sqr(cplx a) { return a.re * a.re + a.im * a.im; }
static inline cplx c_exp(cplx a) { double e = exp(a.re); return c_mk(e * cos(a.im), e * sin(a.im)); }
static inline cplx c_cosh(cplx a) { return c_mk(cosh(a.re) * cos(a.im), sinh(a.re) * sin(a.im)); }
static inline cplx c_sinh(cplx a) { return c_mk(sinh(a.re) * cos(a.im), cosh(a.re) * sin(a.im)); }

void orc_c_exp(const double* a, double* out) { cplx r = c_exp(c_mk(a[0], a[1])); out[0] = r.re; out[1] = r.im; }
void orc_c_cosh(const double* a, double* out) { cplx r = c_cosh(c_mk(a[0], a[1])); out[0] = r.re; out[1] = r.im; }
void orc_c_sinh(const double* a, double* out) { cplx r = c_sinh(c_mk(a[0], a[1])); out[0] = r.re; out[1] = r.im; }

/* operator.rs:195-199 check_controls */
static inline int controls_set(uint64_t idx, uint64_t cmask) { return (idx & cmask) == cmask; }

static uint64_t mask_of(const uint32_t* q, int n) {
    uint64_t m = 0;
    for (int i = 0; i < n; i++) m |= 1ull << q[i];
    return m;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- generic 2x2 pair sweep ------------------------------------------- */
/* operator.rs:2225-2250 (Unitary2), and the pair index math of 347-349 */
static void sweep_u2(cplx* a, int n, int t, uint64_t cmask, const cplx m[4]) {
    const int64_t half = 1ll << (n - 1);
    const uint64_t tb = 1ull << t;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < half; k++) {
        uint64_t i0 = (((uint64_t)k >> t) << (t + 1)) | ((uint64_t)k & (tb - 1));
        uint64_t i1 = i0 | tb;
        if (!controls_set(i0, cmask)) continue;
        cplx ai = a[i0], aj = a[i1];
        a[i0] = c_add(c_mul(m[0], ai), c_mul(m[1], aj));   /* operator.rs:2246 */
        a[i1] = c_add(c_mul(m[2], ai), c_mul(m[3], aj));   /* operator.rs:2247 */
    }
}

/* multiply by `phase` where target bit = 1 and controls set (operator.rs:1193-1211 etc.) */
static void sweep_phase(cplx* a, int n, int t, uint64_t cmask, cplx phase) {
    const int64_t dim = 1ll << n;
    const uint64_t need = cmask | (1ull << t);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < dim; i++) {
        if (((uint64_t)i & need) == need) a[i] = c_mul(a[i], phase);
    }
}

/*
 * Apply one reference operator in place.
 * targets/controls are qubit indices (bit q of the amplitude index = qubit q).
 * params: angle in params[0] (P, RX, RY, RZ); U2: row-major m00,m01,m10,m11 as
 * (re,im) pairs; MATCHGATE: theta, phi1, phi2.
 * No validation here: the reference's validate_qubits (operator.rs:214-273) is
 * restated on the Python side of the oracle (oracle/refapi.py).
 */
void orc_gate(cplx* a, int n, int kind, const uint32_t* targets, int nt,
              const uint32_t* controls, int nc, const double* params) {
    const uint64_t cmask = mask_of(controls, nc);
    const int t = nt > 0 ? (int)targets[0] : 0;
    const int64_t dim = 1ll << n;
    const int64_t half = 1ll << (n - 1);
    const uint64_t tb = 1ull << t;
    (void)dim;
    switch (kind) {
    case G_I: /* operator.rs:1112-1123: validated clone */
        return;
    case G_H: { /* operator.rs:316, 395-416 */
        const double s = 1.0 / sqrt(2.0);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < half; k++) {
            uint64_t i0 = (((uint64_t)k >> t) << (t + 1)) | ((uint64_t)k & (tb - 1));
            uint64_t i1 = i0 | tb;
            if (!controls_set(i0, cmask)) continue;
            cplx a0 = a[i0], a1 = a[i1];
            a[i0] = c_scale(s, c_add(a0, a1));
            a[i1] = c_scale(s, c_sub(a0, a1));
        }
        return;
    }
    case G_X: case G_CNOT: case G_TOFFOLI: { /* operator.rs:574-582; CNOT 684; Toffoli 1074 */
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < half; k++) {
            uint64_t i0 = (((uint64_t)k >> t) << (t + 1)) | ((uint64_t)k & (tb - 1));
            uint64_t i1 = i0 | tb;
            if (!controls_set(i0, cmask)) continue;
            cplx a0 = a[i0]; a[i0] = a[i1]; a[i1] = a0;
        }
        return;
    }
    case G_Y: { /* operator.rs:583-591: new[i] = -i*amp_j ; new[j] = i*amp_i */
        const cplx ic = c_mk(0.0, 1.0);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < half; k++) {
            uint64_t i0 = (((uint64_t)k >> t) << (t + 1)) | ((uint64_t)k & (tb - 1));
            uint64_t i1 = i0 | tb;
            if (!controls_set(i0, cmask)) continue;
            cplx ai = a[i0], aj = a[i1];
            a[i0] = c_mul(c_neg(ic), aj);
            a[i1] = c_mul(ic, ai);
        }
        return;
    }
    case G_Z: { /* operator.rs:592-596: negate where target bit = 1 */
        const uint64_t need = cmask | tb;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < dim; i++)
            if (((uint64_t)i & need) == need) a[i] = c_neg(a[i]);
        return;
    }
    case G_S:   sweep_phase(a, n, t, cmask, c_mk(0.0, 1.0)); return;   /* operator.rs:1202 */
    case G_SDG: sweep_phase(a, n, t, cmask, c_mk(0.0, -1.0)); return;  /* operator.rs:1394 */
    case G_T: { /* operator.rs:1298-1299 */
        const double s = 1.0 / sqrt(2.0);
        sweep_phase(a, n, t, cmask, c_mk(s, s)); return;
    }
    case G_TDG: { /* operator.rs:1490-1491 */
        const double s = 1.0 / sqrt(2.0);
        sweep_phase(a, n, t, cmask, c_mk(s, -s)); return;
    }
    case G_P: /* operator.rs:1610 */
        sweep_phase(a, n, t, cmask, c_mk(cos(params[0]), sin(params[0]))); return;
    case G_RX: { /* operator.rs:1744-1762 */
        const double h = params[0] / 2.0, c = cos(h), s = sin(h);
        const cplx ic = c_mk(0.0, 1.0);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < half; k++) {
            uint64_t i0 = (((uint64_t)k >> t) << (t + 1)) | ((uint64_t)k & (tb - 1));
            uint64_t i1 = i0 | tb;
            if (!controls_set(i0, cmask)) continue;
            cplx ai = a[i0], aj = a[i1];
            /* cos_half*amp_i - i_complex*sin_half*amp_j   (Complex*f64, then Complex*Complex) */
            a[i0] = c_sub(c_scale(c, ai), c_mul(c_scale_r(ic, s), aj));
            /* -i_complex*sin_half*amp_i + cos_half*amp_j */
            a[i1] = c_add(c_mul(c_scale_r(c_neg(ic), s), ai), c_scale(c, aj));
        }
        return;
    }
    case G_RY: { /* operator.rs:1884-1901 */
        const double h = params[0] / 2.0, c = cos(h), s = sin(h);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < half; k++) {
            uint64_t i0 = (((uint64_t)k >> t) << (t + 1)) | ((uint64_t)k & (tb - 1));
            uint64_t i1 = i0 | tb;
            if (!controls_set(i0, cmask)) continue;
            cplx ai = a[i0], aj = a[i1];
            a[i0] = c_sub(c_scale(c, ai), c_scale(s, aj));
            a[i1] = c_add(c_scale(s, ai), c_scale(c, aj));
        }
        return;
    }
    case G_RZ: { /* operator.rs:2013-2029: touches every amplitude whose controls are set */
        const double h = params[0] / 2.0;
        const cplx p0 = c_mk(cos(h), -sin(h)), p1 = c_mk(cos(h), sin(h));
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < dim; i++) {
            if (!controls_set((uint64_t)i, cmask)) continue;
            a[i] = c_mul(a[i], (((uint64_t)i >> t) & 1) ? p1 : p0);
        }
        return;
    }
    case G_U2: { /* operator.rs:2209-2266 */
        cplx m[4];
        for (int j = 0; j < 4; j++) m[j] = c_mk(params[2 * j], params[2 * j + 1]);
        sweep_u2(a, n, t, cmask, m);
        return;
    }
    case G_SWAP: { /* operator.rs:800-813 */
        const int t2 = (int)targets[1];
        const uint64_t b1 = 1ull << t, b2 = 1ull << t2;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < dim; i++) {
            uint64_t ui = (uint64_t)i;
            if (((ui >> t) & 1) != ((ui >> t2) & 1)) {
                uint64_t j = ui ^ b1 ^ b2;
                if (ui < j && controls_set(ui, cmask)) {
                    cplx ai = a[ui]; a[ui] = a[j]; a[j] = ai;
                }
            }
        }
        return;
    }
    case G_MATCHGATE: { /* operator.rs:982-1007 */
        const int q1 = t, q2 = t + 1;
        const double ch = cos(params[0] / 2.0), sh = sin(params[0] / 2.0);
        const cplx e1 = c_exp(c_mk(0.0, params[1])), e2 = c_exp(c_mk(0.0, params[2]));
        const int64_t quarter = 1ll << (n - 2);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < quarter; i++) {
            uint64_t ui = (uint64_t)i;
            uint64_t k = ((ui >> q1) << (q1 + 1)) | (ui & ((1ull << q1) - 1));
            uint64_t l = ((k >> (q2 - 1)) << q2) | (k & ((1ull << (q2 - 1)) - 1));
            uint64_t i01 = l | (1ull << q1), i10 = l | (1ull << q2), i11 = i01 | i10;
            if (controls_set(i01, cmask)) {
                cplx a01 = a[i01], a10 = a[i10];
                /* cos*amp01 - exp_i_phi1*sin*amp10 ; sin*amp01 + exp_i_phi1*cos*amp10 */
                a[i01] = c_sub(c_scale(ch, a01), c_mul(c_scale_r(e1, sh), a10));
                a[i10] = c_add(c_scale(sh, a01), c_mul(c_scale_r(e1, ch), a10));
            }
            if (controls_set(i11, cmask)) a[i11] = c_mul(a[i11], e2);
        }
        return;
    }
    default:
        return;
    }
}

/* ---- the reference's rayon branch, pass for pass (CPU baseline timing) --- */
typedef struct { uint64_t idx; cplx val; } upd_t;          /* (usize, Complex<f64>) = 24 B */

/*
 * operator.rs:339-360 (and the same shape in every pair-type operator):
 *   1. new_state_vec = state.state_vector.clone()           (serial memcpy)
 *   2. updates = (0..).into_par_iter().flat_map(|k| vec![(i0,..),(i1,..)]).collect()
 *      -- one heap-allocated 2-element Vec per pair, gathered into one Vec
 *   3. for (idx,val) in updates { new_state_vec[idx] = val } (serial scatter)
 * Phase-type operators (operator.rs:1193-1200, 1995-2011): clone, then a parallel
 * in-place multiply.  `src` is left untouched; the result goes to `dst`.
 * Supports the kinds the benchmark circuits use: H, X/CNOT/TOFFOLI, Y, RX, RY, U2,
 * SWAP (pair-type) and Z, S, SDG, T, TDG, P, RZ (phase-type).
 */
void orc_gate_faithful(const cplx* src, cplx* dst, int n, int kind,
                       const uint32_t* targets, int nt, const uint32_t* controls, int nc,
                       const double* params) {
    const int64_t dim = 1ll << n;
    memcpy(dst, src, (size_t)dim * sizeof(cplx));            /* Vec::clone */
    switch (kind) {
    case G_Z: case G_S: case G_SDG: case G_T: case G_TDG: case G_P: case G_RZ: case G_I:
        orc_gate(dst, n, kind, targets, nt, controls, nc, params);  /* par_iter_mut in place */
        return;
    default: break;
    }
    const uint64_t cmask = mask_of(controls, nc);
    const int t = (int)targets[0];
    const uint64_t tb = 1ull << t;
    const int64_t half = 1ll << (n - 1);
    int nthreads = orc_num_threads();
    upd_t** chunks = (upd_t**)calloc((size_t)nthreads, sizeof(upd_t*));
    int64_t* counts = (int64_t*)calloc((size_t)nthreads, sizeof(int64_t));
    double c = 0, s = 0; cplx m[4]; const double hs = 1.0 / sqrt(2.0);
    if (kind == G_RX || kind == G_RY) { c = cos(params[0] / 2.0); s = sin(params[0] / 2.0); }
    if (kind == G_U2) for (int j = 0; j < 4; j++) m[j] = c_mk(params[2 * j], params[2 * j + 1]);
    const int t2 = (kind == G_SWAP) ? (int)targets[1] : 0;
#pragma omp parallel
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num(), nth = omp_get_num_threads();
#else
        int tid = 0, nth = 1;
#endif
        int64_t lo = half * tid / nth, hi = half * (tid + 1) / nth;
        upd_t* buf = (upd_t*)malloc((size_t)(hi - lo) * 2 * sizeof(upd_t) + 16);
        int64_t cnt = 0;
        for (int64_t k = lo; k < hi; k++) {
            uint64_t i0 = (((uint64_t)k >> t) << (t + 1)) | ((uint64_t)k & (tb - 1));
            uint64_t i1 = i0 | tb;
            cplx v0, v1;
            if (kind == G_SWAP) {
                /* operator.rs:774-797: pairs whose two target bits differ (here t=0, t2=1) */
                if (((i0 >> t2) & 1) == 0) continue;
                i1 = i0 ^ tb ^ (1ull << t2);
                if (!controls_set(i0, cmask)) continue;
                v0 = src[i1]; v1 = src[i0];
            } else {
                if (!controls_set(i0, cmask)) continue;
                cplx a0 = src[i0], a1 = src[i1];
                switch (kind) {
                case G_H: v0 = c_scale(hs, c_add(a0, a1)); v1 = c_scale(hs, c_sub(a0, a1)); break;
                case G_X: case G_CNOT: case G_TOFFOLI: v0 = a1; v1 = a0; break;
                case G_Y: v0 = c_mul(c_mk(-0.0, -1.0), a1); v1 = c_mul(c_mk(0.0, 1.0), a0); break;
                case G_RX:
                    v0 = c_sub(c_scale(c, a0), c_mul(c_mk(0.0, s), a1));
                    v1 = c_add(c_mul(c_mk(-0.0, -s), a0), c_scale(c, a1)); break;
                case G_RY:
                    v0 = c_sub(c_scale(c, a0), c_scale(s, a1));
                    v1 = c_add(c_scale(s, a0), c_scale(c, a1)); break;
                default: /* G_U2 */
                    v0 = c_add(c_mul(m[0], a0), c_mul(m[1], a1));
                    v1 = c_add(c_mul(m[2], a0), c_mul(m[3], a1)); break;
                }
            }
            /* vec![(i0, v0), (i1, v1)]: one small heap allocation per pair */
            upd_t* pair = (upd_t*)malloc(2 * sizeof(upd_t));
            pair[0].idx = i0; pair[0].val = v0;
            pair[1].idx = i1; pair[1].val = v1;
            buf[cnt++] = pair[0]; buf[cnt++] = pair[1];
            free(pair);
        }
        chunks[tid] = buf; counts[tid] = cnt;
    }
    /* collect(): concatenate the per-thread pieces into one Vec */
    int64_t total = 0;
    for (int i = 0; i < nthreads; i++) total += counts[i];
    upd_t* updates = (upd_t*)malloc((size_t)(total > 0 ? total : 1) * sizeof(upd_t));
    int64_t off = 0;
    for (int i = 0; i < nthreads; i++) {
        if (chunks[i]) { memcpy(updates + off, chunks[i], (size_t)counts[i] * sizeof(upd_t)); free(chunks[i]); }
        off += counts[i];
    }
    /* serial scatter, operator.rs:358-360 */
    for (int64_t i = 0; i < total; i++) dst[updates[i].idx] = updates[i].val;
    free(updates); free(chunks); free(counts);
}

/* ---- state arithmetic (state.rs:2687-2862) ----------------------------- */
void orc_scale(cplx* a, int64_t len, const double* z) { /* state.rs:2692-2706: amplitude * rhs */
    const cplx w = c_mk(z[0], z[1]);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < len; i++) a[i] = c_mul(a[i], w);
}
void orc_add(cplx* a, const cplx* b, int64_t len) { /* state.rs:2779-2800 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < len; i++) a[i] = c_add(a[i], b[i]);
}
void orc_sub(cplx* a, const cplx* b, int64_t len) { /* state.rs:2841-2862 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < len; i++) a[i] = c_sub(a[i], b[i]);
}
/* state.rs:890-917: sum conj(a_i)*b_i.  rayon's sum fixes no order; the oracle sums serially. */
void orc_inner_product(const cplx* a, const cplx* b, int64_t len, double* out) {
    cplx acc = c_mk(0.0, 0.0);
    for (int64_t i = 0; i < len; i++) acc = c_add(acc, c_mul(c_conj(a[i]), b[i]));
    out[0] = acc.re; out[1] = acc.im;
}
/* state.rs:117 / 924-929 */
double orc_norm_sqr(const cplx* a, int64_t len) {
    double acc = 0.0;
    for (int64_t i = 0; i < len; i++) acc += c_norm_sqr(a[i]);
    return acc;
}
/* state.rs:924-945 normalise: returns 0 ok, 1 zero norm */
int orc_normalise(cplx* a, int64_t len) {
    double norm = sqrt(orc_norm_sqr(a, len));
    if (norm == 0.0) return 1;
    if (norm == 1.0) return 0;
    for (int64_t i = 0; i < len; i++) a[i] = c_divr(a[i], norm);
    return 0;
}

/* ---- PauliString (pauli_string.rs:139-262) ---------------------------- */
/* paulis[i] in {1:X, 2:Y, 3:Z}; applies the single Paulis one after another
 * (pauli_string.rs:172-184 apply_operators), then optionally multiplies by coeff (151). */
void orc_pauli_apply(cplx* a, int n, const uint32_t* qubits, const int32_t* paulis, int k,
                     const double* coeff /* NULL: no coefficient */) {
    for (int i = 0; i < k; i++) {
        int kind = paulis[i] == 1 ? G_X : (paulis[i] == 2 ? G_Y : G_Z);
        orc_gate(a, n, kind, &qubits[i], 1, NULL, 0, NULL);
    }
    if (coeff) orc_scale(a, 1ll << n, coeff);
}
/* pauli_string.rs:237-262 apply_exp_factor with alpha = coefficient*factor already formed by
 * the caller: out = psi*cosh(alpha) + (P psi)*sinh(alpha); empty string: psi*exp(alpha). */
void orc_pauli_exp(cplx* a, int n, const uint32_t* qubits, const int32_t* paulis, int k,
                   const double* alpha) {
    const int64_t dim = 1ll << n;
    const cplx al = c_mk(alpha[0], alpha[1]);
    if (k == 0) { cplx e = c_exp(al); double z[2] = { e.re, e.im }; orc_scale(a, dim, z); return; }
    cplx* p = (cplx*)malloc((size_t)dim * sizeof(cplx));
    memcpy(p, a, (size_t)dim * sizeof(cplx));
    orc_pauli_apply(p, n, qubits, paulis, k, NULL);
    const cplx ch = c_cosh(al), sh = c_sinh(al);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < dim; i++) a[i] = c_add(c_mul(a[i], ch), c_mul(p[i], sh));
    free(p);
}
/* pauli_string.rs:485-507: <psi| P_k |psi> for one term (phi = term.apply(psi) incl. coefficient) */
void orc_pauli_expect(const cplx* a, int n, const uint32_t* qubits, const int32_t* paulis, int k,
                      const double* coeff, double* out) {
    const int64_t dim = 1ll << n;
    cplx* p = (cplx*)malloc((size_t)dim * sizeof(cplx));
    memcpy(p, a, (size_t)dim * sizeof(cplx));
    orc_pauli_apply(p, n, qubits, paulis, k, coeff);
    orc_inner_product(a, p, dim, out);
    free(p);
}

/* ---- measurement (state.rs:525-668) ----------------------------------- */
static inline uint64_t bin_of(uint64_t idx, const uint32_t* qubits, int m) {
    uint64_t b = 0;
    for (int j = 0; j < m; j++) if ((idx >> qubits[j]) & 1) b |= 1ull << j;  /* state.rs:567-571 */
    return b;
}
/* state.rs:559-588 marginal probabilities (un-normalised); serial index order */
void orc_probabilities(const cplx* a, int n, const uint32_t* qubits, int m, double* probs) {
    const int64_t dim = 1ll << n, nb = 1ll << m;
    for (int64_t b = 0; b < nb; b++) probs[b] = 0.0;
    for (int64_t i = 0; i < dim; i++) probs[bin_of((uint64_t)i, qubits, m)] += c_norm_sqr(a[i]);
}
/* state.rs:591-619: normalise by the total, linear CDF scan, first bin with u < cumsum,
 * fallback to the last bin.  Returns -1 if the total is < f64::EPSILON (UnknownError). */
int64_t orc_sample_bin(const double* probs, int64_t nb, double u) {
    double total = 0.0;
    for (int64_t b = 0; b < nb; b++) total += probs[b];
    if (total < 2.220446049250313e-16) return -1;
    double cum = 0.0;
    for (int64_t b = 0; b < nb; b++) {
        cum += probs[b] / total;
        if (u < cum) return b;
    }
    return nb - 1;
}
/* distance of u to the nearest CDF edge (tests assert no draw sits on an edge) */
double orc_sample_margin(const double* probs, int64_t nb, double u) {
    double total = 0.0, cum = 0.0, best = 1.0;
    for (int64_t b = 0; b < nb; b++) total += probs[b];
    for (int64_t b = 0; b < nb; b++) {
        cum += probs[b] / total;
        double d = fabs(u - cum);
        if (d < best) best = d;
    }
    return best;
}
/* state.rs:622-654 collapse onto `bin` and divide by sqrt(norm^2) */
void orc_collapse(cplx* a, int n, const uint32_t* qubits, int m, uint64_t bin) {
    const int64_t dim = 1ll << n;
    double nsq = 0.0;
    for (int64_t i = 0; i < dim; i++) {
        if (bin_of((uint64_t)i, qubits, m) == bin) nsq += c_norm_sqr(a[i]);
        else a[i] = c_mk(0.0, 0.0);
    }
    if (nsq > 2.220446049250313e-16) {
        double f = sqrt(nsq);
        for (int64_t i = 0; i < dim; i++) a[i] = c_divr(a[i], f);
    }
}

/* ---- shared-seed contract (defined by this build) ---------- */
/* splitmix64; u = (x >> 11) * 2^-53 */
uint64_t orc_splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
double orc_uniform(uint64_t seed, uint64_t k) { /* k-th draw of the stream seeded with `seed` */
    uint64_t s = seed + k * 0x9E3779B97F4A7C15ull;
    return (double)(orc_splitmix64(&s) >> 11) * (1.0 / 9007199254740992.0);
}
/* sample `shots` bins from one probability table: counts are what parity compares */
void orc_sample(const double* probs, int64_t nb, uint64_t seed, int64_t shots, int64_t* bins) {
    for (int64_t k = 0; k < shots; k++) bins[k] = orc_sample_bin(probs, nb, orc_uniform(seed, (uint64_t)k));
}

/* pseudo-random normalised state: splitmix64 + Box-Muller, then normalise */
void orc_random_state(cplx* a, int n, uint64_t seed) {
    const int64_t dim = 1ll << n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < dim; i++) {
        double u1 = orc_uniform(seed, 2 * (uint64_t)i), u2 = orc_uniform(seed, 2 * (uint64_t)i + 1);
        double r = sqrt(-2.0 * log(u1 + 1.1102230246251565e-16));
        a[i] = c_mk(r * cos(6.283185307179586 * u2), r * sin(6.283185307179586 * u2));
    }
    /* norm in a FIXED association order (an OpenMP reduction combines the per-thread partials in whatever order the
     * threads finish, and depends on the thread count): blocks of 4096 summed serially, block sums added in order,
     * so the golden fixtures reproduce bit for bit on any machine */
    const int64_t blk = 4096, nblk = (dim + blk - 1) / blk;
    double* part = (double*)malloc((size_t)nblk * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nblk; b++) {
        const int64_t lo = b * blk, hi = lo + blk < dim ? lo + blk : dim;
        double acc = 0.0;
        for (int64_t i = lo; i < hi; i++) acc += c_norm_sqr(a[i]);
        part[b] = acc;
    }
    double nsq = 0.0;
    for (int64_t b = 0; b < nblk; b++) nsq += part[b];
    free(part);
    const double f = 1.0 / sqrt(nsq);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < dim; i++) a[i] = c_scale(f, a[i]);
}
