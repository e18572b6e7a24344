// quant_iron_b200.hpp -- header-only C++ host facade over the C ABI (include/qiron_b200.h).
//
// The reference is compiled Rust and its toolchain is absent from this image, so the compiled-language
// host side is C++: the same names, argument order and error behaviour as the crate's public surface
// for the state-vector path: every State constructor, metric and gate method (state.rs:99-2345), Operator and the built-in
// operators, Gate (all four kinds) / Circuit / CircuitBuilder with the full adder set (circuit.rs:288-1742), Subroutine::qft,
// PauliString / SumOp, measure / measure_n with custom bases, Trotter, heisenberg_1d/2d, ising_1d/2d, Parameter and the
// parametric gates.  Not mirrored here: the QASM emitter (Python host mirror only).  Citations are file:line under the reference root.
// `&self -> State` methods are device clone + in-place kernel; methods with a trailing underscore act in place.
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <array>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "qiron_b200.h"

namespace quant_iron {

using cplx = std::complex<double>;

// errors.rs:3-97
struct Error : std::runtime_error {
    int code;
    uint64_t payload[2];
    std::string variant;
    Error(int c, std::string v, uint64_t p0, uint64_t p1, const std::string& msg)
        : std::runtime_error(v + ": " + msg), code(c), variant(std::move(v)) { payload[0] = p0; payload[1] = p1; }
};

inline void check(int status) {
    if (status == QI_OK) return;
    static const char* names[] = {"Ok", "InvalidNumberOfMeasurements", "OverlappingControlAndTargetQubits",
        "InvalidNumberOfQubits", "InvalidQubitIndex", "StateVectorNotNormalised", "NonUnitaryMatrix", "InvalidNumberOfInputs",
        "MismatchedNumberOfParameters", "UnknownError", "CudaError", "GpuContextLockError", "CircuitMacroError",
        "InvalidInputValue", "ZeroNorm", "InvalidPauliStringCoefficient", "InvalidArgument", "PeerError"};
    uint64_t p[2] = {0, 0};
    char msg[256] = {0};
    qi_last_error(p, msg, sizeof(msg));
    throw Error(status, status >= 0 && status <= 17 ? names[status] : "UnknownError", p[0], p[1], msg);
}

class State;

// operator.rs:151-190
struct Operator {
    virtual ~Operator() = default;
    virtual int kind() const = 0;
    virtual std::vector<double> params() const { return {}; }
    virtual size_t base_qubits() const { return 1; }
    State apply(const State& state, const std::vector<size_t>& targets, const std::vector<size_t>& controls = {}) const;
};
#define QI_SIMPLE_OP(NAME, KIND, BASE) \
    struct NAME : Operator { int kind() const override { return KIND; } size_t base_qubits() const override { return BASE; } }
QI_SIMPLE_OP(Hadamard, QI_GATE_H, 1);
QI_SIMPLE_OP(PauliX, QI_GATE_X, 1);
QI_SIMPLE_OP(PauliY, QI_GATE_Y, 1);
QI_SIMPLE_OP(PauliZ, QI_GATE_Z, 1);
QI_SIMPLE_OP(Identity, QI_GATE_I, 1);
QI_SIMPLE_OP(PhaseS, QI_GATE_S, 1);
QI_SIMPLE_OP(PhaseT, QI_GATE_T, 1);
QI_SIMPLE_OP(PhaseSdag, QI_GATE_SDG, 1);
QI_SIMPLE_OP(PhaseTdag, QI_GATE_TDG, 1);
QI_SIMPLE_OP(CNOT, QI_GATE_CNOT, 2);
QI_SIMPLE_OP(SWAP, QI_GATE_SWAP, 2);
QI_SIMPLE_OP(Toffoli, QI_GATE_TOFFOLI, 3);
#undef QI_SIMPLE_OP
struct AngleOp : Operator {
    double angle;
    int k;
    AngleOp(int kind_, double a) : angle(a), k(kind_) {}
    int kind() const override { return k; }
    std::vector<double> params() const override { return {angle}; }
};
struct PhaseShift : AngleOp { explicit PhaseShift(double a) : AngleOp(QI_GATE_P, a) {} };
struct RotateX : AngleOp { explicit RotateX(double a) : AngleOp(QI_GATE_RX, a) {} };
struct RotateY : AngleOp { explicit RotateY(double a) : AngleOp(QI_GATE_RY, a) {} };
struct RotateZ : AngleOp { explicit RotateZ(double a) : AngleOp(QI_GATE_RZ, a) {} };
struct Unitary2 : Operator {   // operator.rs:2058-2275
    cplx m[2][2];
    int kind() const override { return QI_GATE_U2; }
    std::vector<double> params() const override {
        return {m[0][0].real(), m[0][0].imag(), m[0][1].real(), m[0][1].imag(), m[1][0].real(), m[1][0].imag(), m[1][1].real(), m[1][1].imag()};
    }
    static Unitary2 make(const cplx (&u)[2][2]) {   // Unitary2::new: unitarity check, operator.rs:2092-2118
        Unitary2 r;
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) r.m[i][j] = u[i][j];
        auto p = r.params();
        check(qi_unitary2_check(p.data()));
        return r;
    }
    static Unitary2 from_ry_phase(double theta, double phi) {   // operator.rs:2140-2156
        Unitary2 r;
        double c = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
        cplx e(std::cos(phi), std::sin(phi));
        r.m[0][0] = cplx(c, 0.0); r.m[0][1] = cplx(-e.real() * s, -e.imag() * s);
        r.m[1][0] = cplx(s, 0.0); r.m[1][1] = cplx(e.real() * c, e.imag() * c);
        return r;
    }
    static Unitary2 from_ry_phase_dagger(double theta, double phi) {   // operator.rs:2173-2192: [[c, s], [-e^{-i phi} s, e^{-i phi} c]]
        Unitary2 r;
        double c = std::cos(theta / 2.0), s = std::sin(theta / 2.0);
        cplx e(std::cos(phi), -std::sin(phi));
        r.m[0][0] = cplx(c, 0.0); r.m[0][1] = cplx(s, 0.0);
        r.m[1][0] = cplx(-e.real() * s, -e.imag() * s); r.m[1][1] = cplx(e.real() * c, e.imag() * c);
        return r;
    }
};
struct Matchgate : Operator {   // operator.rs:852-1019
    double theta, phi1, phi2;
    Matchgate(double t, double p1, double p2) : theta(t), phi1(p1), phi2(p2) {}
    int kind() const override { return QI_GATE_MATCHGATE; }
    size_t base_qubits() const override { return 2; }
    std::vector<double> params() const override { return {theta, phi1, phi2}; }
};

enum class MeasurementBasis { Computational = 0, X = 1, Y = 2, Custom = 3 };   // measurement.rs:76-86
// `MeasurementBasis::Custom([[Complex<f64>; 2]; 2])`: the enum above plus its payload
struct Basis {
    MeasurementBasis kind = MeasurementBasis::Computational;
    cplx u[2][2] = {{1.0, 0.0}, {0.0, 1.0}};
    Basis() = default;
    Basis(MeasurementBasis k) : kind(k) {}   // implicit: plain variants
    static Basis custom(const cplx (&m)[2][2]) { Basis b; b.kind = MeasurementBasis::Custom; for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) b.u[i][j] = m[i][j]; return b; }
};
// measurement.rs:15-60
struct MeasurementResult {
    Basis basis;
    std::vector<size_t> indices;
    std::vector<uint8_t> outcomes;
};

struct GateRecord {   // owns the control list a qi_gate points into
    qi_gate g;
    std::vector<uint32_t> controls;
};
inline GateRecord make_record(const Operator& op, const std::vector<size_t>& targets, const std::vector<size_t>& controls) {
    GateRecord r;
    r.g = qi_gate{};
    r.g.kind = op.kind();
    r.g.num_targets = (uint32_t)targets.size();
    for (size_t i = 0; i < targets.size() && i < 2; i++) r.g.targets[i] = (uint32_t)targets[i];
    for (size_t c : controls) r.controls.push_back((uint32_t)c);
    r.g.num_controls = (uint32_t)r.controls.size();
    auto p = op.params();
    for (size_t i = 0; i < p.size() && i < 8; i++) r.g.params[i] = p[i];
    return r;
}

// state.rs:74-81
class State {
    qi_state* h_ = nullptr;
    explicit State(qi_state* h) : h_(h) {}
    friend struct Operator;
    friend class Circuit;
    friend class PauliString;
    friend class SumOp;

public:
    State(const State& o) { check(qi_state_clone(o.h_, &h_)); }
    State(State&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    State& operator=(State o) { std::swap(h_, o.h_); return *this; }
    ~State() { if (h_) qi_state_free(h_); }
    qi_state* handle() const { return h_; }
    static State adopt(qi_state* h) { return State(h); }   // takes ownership of a handle the C ABI returned

    static State from_vector(const std::vector<cplx>& v) {   // State::new, state.rs:99-127
        qi_state* h = nullptr;
        check(qi_state_from_host(reinterpret_cast<const double*>(v.data()), v.size(), 0, 1, &h));
        return State(h);
    }
    static State new_zero(size_t n) { qi_state* h = nullptr; check(qi_state_new_zero((uint32_t)n, &h)); return State(h); }
    static State new_basis_n(size_t n, uint64_t k) { qi_state* h = nullptr; check(qi_state_new_basis_n((uint32_t)n, k, &h)); return State(h); }
    static State new_plus(size_t n) { qi_state* h = nullptr; check(qi_state_new_plus((uint32_t)n, &h)); return State(h); }
    static State new_minus(size_t n) { qi_state* h = nullptr; check(qi_state_new_minus((uint32_t)n, &h)); return State(h); }
    static State new_ghz(size_t n) { qi_state* h = nullptr; check(qi_state_new_ghz((uint32_t)n, &h)); return State(h); }
    static State new_hartree_fock(size_t num_electrons, size_t num_orbitals) {   // state.rs:140-151
        if (num_orbitals == 0 || num_orbitals < num_electrons) throw Error(QI_ERR_INVALID_INPUT_VALUE, "InvalidInputValue", num_orbitals, 0, "new_hartree_fock");
        return new_basis_n(num_orbitals, ((1ull << num_electrons) - 1ull) << (num_orbitals - num_electrons));
    }
    // Bell states, state.rs:28-66 / 330-373
    static State new_phi_plus() { const double a = std::sqrt(0.5); return from_vector({a, 0.0, 0.0, a}); }
    static State new_phi_minus() { const double a = std::sqrt(0.5); return from_vector({a, 0.0, 0.0, -a}); }
    static State new_psi_plus() { const double a = std::sqrt(0.5); return from_vector({0.0, a, a, 0.0}); }
    static State new_psi_minus() { const double a = std::sqrt(0.5); return from_vector({0.0, a, -a, 0.0}); }

    size_t num_qubits() const { return qi_state_num_qubits(h_); }
    State conj() const { State s(*this); check(qi_conj(s.h_)); return s; }                                  // state.rs:397-403
    bool equals_without_phase(const State& o) const {                                                        // state.rs:384-390
        return num_qubits() == o.num_qubits() && std::fabs(std::abs(inner_product(o)) - 1.0) < 1.1920928955078125e-07;
    }
    double fs_dist(const State& o) const { return std::acos(std::abs(normalise().inner_product(o.normalise()))); }          // state.rs:470-480
    double fs_fidelity(const State& o) const { double a = std::abs(normalise().inner_product(o.normalise())); return a * a; } // state.rs:492-498
    std::vector<cplx> state_vector() const {
        std::vector<cplx> v(qi_state_len(h_));
        check(qi_state_to_host(h_, reinterpret_cast<double*>(v.data()), v.size()));
        return v;
    }
    cplx amplitude(uint64_t i) const { double o[2]; check(qi_state_amplitude(h_, i, o)); return cplx(o[0], o[1]); }
    double probability(uint64_t i) const { return std::norm(amplitude(i)); }
    double norm_sqr() const { double o; check(qi_norm_sqr(h_, &o)); return o; }
    cplx inner_product(const State& o) const { double r[2]; check(qi_inner_product(h_, o.h_, r)); return cplx(r[0], r[1]); }
    State normalise() const { State s(*this); check(qi_normalise(s.h_)); return s; }
    State operator*(cplx z) const { State s(*this); double w[2] = {z.real(), z.imag()}; check(qi_scale(s.h_, w)); return s; }
    State operator+(const State& o) const { State s(*this); check(qi_add(s.h_, o.h_)); return s; }
    State operator-(const State& o) const { State s(*this); check(qi_sub(s.h_, o.h_)); return s; }
    State tensor_product(const State& o) const { qi_state* h = nullptr; check(qi_tensor_product(h_, o.h_, &h)); return State(h); }
    bool approx_eq(const State& o, double tol = 1.1920928955078125e-07) const {   // PartialEq, state.rs:2348-2372
        if (num_qubits() != o.num_qubits()) return false;
        auto a = state_vector(), b = o.state_vector();
        if (a.size() != b.size()) return false;
        for (size_t i = 0; i < a.size(); i++)
            if (std::fabs(a[i].real() - b[i].real()) > tol || std::fabs(a[i].imag() - b[i].imag()) > tol) return false;
        return true;
    }

    // Operator::apply in place
    State& apply_(const Operator& op, const std::vector<size_t>& targets, const std::vector<size_t>& controls = {}) {
        GateRecord r = make_record(op, targets, controls);
        r.g.controls = r.controls.data();
        check(qi_apply_gate(h_, &r.g));
        return *this;
    }
    // state.rs:970-1002
    State operate(const Operator& op, const std::vector<size_t>& targets, const std::vector<size_t>& controls = {}) const {
        if (op.base_qubits() != targets.size() + controls.size())
            throw Error(QI_ERR_INVALID_NUMBER_OF_QUBITS, "InvalidNumberOfQubits", op.base_qubits(), 0, "operate");
        return op.apply(*this, targets, controls);
    }
    // gate methods, state.rs:1019-2345 (same names and argument order)
    State h(size_t q) const { return Hadamard().apply(*this, {q}); }
    State x(size_t q) const { return PauliX().apply(*this, {q}); }
    State y(size_t q) const { return PauliY().apply(*this, {q}); }
    State z(size_t q) const { return PauliZ().apply(*this, {q}); }
    State i(size_t q) const { return Identity().apply(*this, {q}); }
    State s(size_t q) const { return PhaseS().apply(*this, {q}); }
    State t(size_t q) const { return PhaseT().apply(*this, {q}); }
    State s_dag(size_t q) const { return PhaseSdag().apply(*this, {q}); }
    State t_dag(size_t q) const { return PhaseTdag().apply(*this, {q}); }
    State p(size_t q, double a) const { return PhaseShift(a).apply(*this, {q}); }
    State rx(size_t q, double a) const { return RotateX(a).apply(*this, {q}); }
    State ry(size_t q, double a) const { return RotateY(a).apply(*this, {q}); }
    State rz(size_t q, double a) const { return RotateZ(a).apply(*this, {q}); }
    State cnot(size_t control, size_t target) const { return CNOT().apply(*this, {target}, {control}); }          // state.rs:2230
    State swap(size_t a, size_t b) const { return SWAP().apply(*this, {a, b}); }
    State toffoli(size_t c1, size_t c2, size_t target) const { return Toffoli().apply(*this, {target}, {c1, c2}); } // state.rs:2343
    State matchgate(size_t q, double th, double p1, double p2) const { return Matchgate(th, p1, p2).apply(*this, {q}); }
    State ry_phase(size_t q, double th, double ph) const { return Unitary2::from_ry_phase(th, ph).apply(*this, {q}); }
    State multi(const Operator& op, const std::vector<size_t>& targets, const std::vector<size_t>& controls = {}) const {
        State s(*this);
        for (size_t q : targets) s.apply_(op, {q}, controls);
        return s;
    }
    State h_multi(const std::vector<size_t>& qs) const { return multi(Hadamard(), qs); }
    State x_multi(const std::vector<size_t>& qs) const { return multi(PauliX(), qs); }
    State cx_multi(const std::vector<size_t>& t, const std::vector<size_t>& c) const { return multi(PauliX(), t, c); }
    State cp_multi(const std::vector<size_t>& t, const std::vector<size_t>& c, double a) const { return multi(PhaseShift(a), t, c); }
    // the rest of the family (state.rs:1019-2345): <g>_multi(qubits), c<g>_multi(targets, controls)
    using Q = std::vector<size_t>;
    State ch_multi(const Q& t, const Q& c) const { return multi(Hadamard(), t, c); }
    State y_multi(const Q& q) const { return multi(PauliY(), q); }
    State cy_multi(const Q& t, const Q& c) const { return multi(PauliY(), t, c); }
    State z_multi(const Q& q) const { return multi(PauliZ(), q); }
    State cz_multi(const Q& t, const Q& c) const { return multi(PauliZ(), t, c); }
    State i_multi(const Q& q) const { return multi(Identity(), q); }
    State ci_multi(const Q& t, const Q& c) const { return multi(Identity(), t, c); }
    State s_multi(const Q& q) const { return multi(PhaseS(), q); }
    State cs_multi(const Q& t, const Q& c) const { return multi(PhaseS(), t, c); }
    State t_multi(const Q& q) const { return multi(PhaseT(), q); }
    State ct_multi(const Q& t, const Q& c) const { return multi(PhaseT(), t, c); }
    State s_dag_multi(const Q& q) const { return multi(PhaseSdag(), q); }
    State cs_dag_multi(const Q& t, const Q& c) const { return multi(PhaseSdag(), t, c); }
    State t_dag_multi(const Q& q) const { return multi(PhaseTdag(), q); }
    State ct_dag_multi(const Q& t, const Q& c) const { return multi(PhaseTdag(), t, c); }
    State p_multi(const Q& q, double a) const { return multi(PhaseShift(a), q); }
    State rx_multi(const Q& q, double a) const { return multi(RotateX(a), q); }
    State crx_multi(const Q& t, const Q& c, double a) const { return multi(RotateX(a), t, c); }
    State ry_multi(const Q& q, double a) const { return multi(RotateY(a), q); }
    State cry_multi(const Q& t, const Q& c, double a) const { return multi(RotateY(a), t, c); }
    State rz_multi(const Q& q, double a) const { return multi(RotateZ(a), q); }
    State crz_multi(const Q& t, const Q& c, double a) const { return multi(RotateZ(a), t, c); }
    State unitary(size_t q, const cplx (&u)[2][2]) const { return Unitary2::make(u).apply(*this, {q}); }
    State unitary_multi(const Q& q, const cplx (&u)[2][2]) const { return multi(Unitary2::make(u), q); }
    State cunitary_multi(const Q& t, const Q& c, const cplx (&u)[2][2]) const { return multi(Unitary2::make(u), t, c); }
    State ry_phase_multi(const Q& q, double th, double ph) const { return multi(Unitary2::from_ry_phase(th, ph), q); }
    State cry_phase_gates(const Q& t, const Q& c, double th, double ph) const { return multi(Unitary2::from_ry_phase(th, ph), t, c); }
    State ry_phase_dag(size_t q, double th, double ph) const { return Unitary2::from_ry_phase_dagger(th, ph).apply(*this, {q}); }
    State ry_phase_dag_multi(const Q& q, double th, double ph) const { return multi(Unitary2::from_ry_phase_dagger(th, ph), q); }
    State cry_phase_dag_gates(const Q& t, const Q& c, double th, double ph) const { return multi(Unitary2::from_ry_phase_dagger(th, ph), t, c); }
    State cswap(size_t t1, size_t t2, const Q& controls) const { return SWAP().apply(*this, {t1, t2}, controls); }
    State cmatchgate(size_t q, double th, double p1, double p2, const Q& controls) const { return Matchgate(th, p1, p2).apply(*this, {q}, controls); }   // state.rs:2317-2324

    // State::measure (state.rs:525-730) with the crate's return type; the shared-seed contract replaces rand::rng()
    // (draw `draw` of the stream seeded `seed`).  An empty qubit list measures every qubit (531-541).
    std::pair<MeasurementResult, State> measure(const Basis& basis, const Q& qubits, uint64_t seed, uint64_t draw = 0) const {
        State out(*this);
        std::vector<uint32_t> q(qubits.begin(), qubits.end());
        MeasurementResult r;
        r.basis = basis;
        r.indices = qubits;
        if (qubits.empty()) for (size_t k = 0; k < num_qubits(); k++) r.indices.push_back(k);
        r.outcomes.assign(r.indices.size(), 0);
        uint64_t bin = 0;
        double u[8];
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) { u[4 * a + 2 * b] = basis.u[a][b].real(); u[4 * a + 2 * b + 1] = basis.u[a][b].imag(); }
        check(qi_measure(out.h_, (int)basis.kind, basis.kind == MeasurementBasis::Custom ? u : nullptr, q.data(), (uint32_t)q.size(), seed, draw,
                         r.outcomes.data(), &bin));
        return {r, out};
    }
    // State::measure_n (state.rs:750-784): n independent measurements of the same state; measurement k uses draw k
    std::vector<std::pair<MeasurementResult, State>> measure_n(const Basis& basis, const Q& qubits, size_t n, uint64_t seed) const {
        if (n == 0) throw Error(QI_ERR_INVALID_NUMBER_OF_MEASUREMENTS, "InvalidNumberOfMeasurements", 0, 0, "measure_n");
        std::vector<std::pair<MeasurementResult, State>> out;
        out.reserve(n);
        for (size_t k = 0; k < n; k++) out.push_back(measure(basis, qubits, seed, k));
        return out;
    }
    std::vector<double> probabilities(const Q& qubits) const {   // un-normalised marginal table, state.rs:559-588
        std::vector<uint32_t> q(qubits.begin(), qubits.end());
        std::vector<double> out((size_t)1 << qubits.size());
        check(qi_probabilities(h_, q.data(), (uint32_t)q.size(), out.data()));
        return out;
    }

    // State::measure (state.rs:525-730), in place; returns outcomes[j] = bit j of the sampled bin
    std::vector<uint8_t> measure_(MeasurementBasis basis, const std::vector<size_t>& qubits, uint64_t seed, uint64_t draw = 0) {
        std::vector<uint32_t> q(qubits.begin(), qubits.end());
        std::vector<uint8_t> out(qubits.empty() ? num_qubits() : qubits.size());
        uint64_t bin = 0;
        check(qi_measure(h_, (int)basis, nullptr, q.data(), (uint32_t)q.size(), seed, draw, out.data(), &bin));
        return out;
    }
    std::vector<uint64_t> sample_counts(const std::vector<size_t>& qubits, uint64_t shots, uint64_t seed) const {
        std::vector<uint32_t> q(qubits.begin(), qubits.end());
        std::vector<uint64_t> bins(shots);
        check(qi_sample(h_, q.data(), (uint32_t)q.size(), shots, seed, bins.data()));
        return bins;
    }
};

inline State Operator::apply(const State& state, const std::vector<size_t>& targets, const std::vector<size_t>& controls) const {
    State out(state);
    out.apply_(*this, targets, controls);
    return out;
}

// pauli_string.rs:13-287
enum class Pauli : uint8_t { X = 1, Y = 2, Z = 3 };
class PauliString {
    std::map<size_t, Pauli> ops_;
    cplx coefficient_;

public:
    explicit PauliString(cplx c) : coefficient_(c) {}
    PauliString& add_op(size_t q, Pauli p) {
        if (ops_.count(q)) throw std::logic_error("Duplicate Pauli string operator for qubit");   // panic, pauli_string.rs:66-70
        ops_[q] = p;
        return *this;
    }
    PauliString with_op(size_t q, Pauli p) const { PauliString r(*this); r.add_op(q, p); return r; }
    cplx coefficient() const { return coefficient_; }
    size_t len() const { return ops_.size(); }
    const std::map<size_t, Pauli>& ops() const { return ops_; }
    std::vector<struct Gate> to_gates() const;   // pauli_string.rs:118-122: one Pauli operator gate per factor
    struct Term { qi_pauli_term t; std::vector<uint32_t> q; std::vector<uint8_t> p; };
    std::unique_ptr<Term> term() const {
        auto r = std::make_unique<Term>();
        for (auto& kv : ops_) { r->q.push_back((uint32_t)kv.first); r->p.push_back((uint8_t)kv.second); }
        r->t.num_ops = (uint32_t)r->q.size();
        r->t.qubits = r->q.data();
        r->t.paulis = r->p.data();
        r->t.coefficient[0] = coefficient_.real();
        r->t.coefficient[1] = coefficient_.imag();
        return r;
    }
    State apply(const State& s) const { State o(s); auto t = term(); check(qi_apply_pauli_string(o.handle(), &t->t, 1)); return o; }                 // 139-151
    State apply_normalised(const State& s) const { State o(s); auto t = term(); check(qi_apply_pauli_string(o.handle(), &t->t, 0)); check(qi_normalise(o.handle())); return o; }
    State apply_exp_factor(const State& s, cplx f) const {                                                                                            // 237-262
        State o(s); auto t = term(); double w[2] = {f.real(), f.imag()};
        check(qi_apply_pauli_exp(o.handle(), &t->t, w));
        return o;
    }
    State apply_exp(const State& s) const { return apply_exp_factor(s, cplx(1.0, 0.0)); }
    State apply_exp_neg_i_dt(const State& s, double dt) const {                                                                                       // 281-287
        if (coefficient_.imag() != 0.0) throw Error(QI_ERR_INVALID_PAULI_STRING_COEFFICIENT, "InvalidPauliStringCoefficient", 0, 0, "imaginary coefficient");
        return apply_exp_factor(s, cplx(0.0, -dt));
    }
};

class SumOp {   // pauli_string.rs:398-507
public:
    std::vector<PauliString> terms;
    explicit SumOp(std::vector<PauliString> t = {}) : terms(std::move(t)) {}
    size_t num_terms() const { return terms.size(); }
    SumOp with_term(const PauliString& t) const { SumOp r(*this); r.terms.push_back(t); return r; }
    State apply(const State& s) const {   // pauli_string.rs:453-466: sum_k P_k psi (a new state; empty sum = 0 * psi)
        std::vector<std::unique_ptr<PauliString::Term>> keep;
        std::vector<qi_pauli_term> arr;
        for (auto& t : terms) { keep.push_back(t.term()); arr.push_back(keep.back()->t); }
        qi_state* h = nullptr;
        check(qi_apply_pauli_sum(s.handle(), arr.data(), arr.size(), &h));
        return State::adopt(h);
    }
    cplx expectation_value(const State& s) const {
        std::vector<std::unique_ptr<PauliString::Term>> keep;
        std::vector<qi_pauli_term> arr;
        for (auto& t : terms) { keep.push_back(t.term()); arr.push_back(keep.back()->t); }
        double o[2];
        check(qi_expect_pauli_sum(s.handle(), arr.data(), arr.size(), o));
        return cplx(o[0], o[1]);
    }
    // exp(f_k c_k P_k) for k = 0, 1, ... applied in order: one apply_exp_factor (pauli_string.rs:237-262) per term,
    // fused into register-window passes on the device
    State apply_exp_sequence(const State& s, const std::vector<cplx>& factors) const {
        if (factors.size() != terms.size()) throw Error(QI_ERR_MISMATCHED_NUMBER_OF_PARAMETERS, "MismatchedNumberOfParameters", terms.size(), factors.size(), "factors");
        std::vector<std::unique_ptr<PauliString::Term>> keep;
        std::vector<qi_pauli_term> arr;
        std::vector<double> f;
        for (auto& t : terms) { keep.push_back(t.term()); arr.push_back(keep.back()->t); }
        for (auto& z : factors) { f.push_back(z.real()); f.push_back(z.imag()); }
        State o(s);
        check(qi_apply_pauli_exp_sequence(o.handle(), arr.data(), arr.size(), f.data()));
        return o;
    }
    State trotter_evolve(const State& s, double dt, uint64_t steps, int order) const {   // time_evolution.rs:140-167
        std::vector<std::unique_ptr<PauliString::Term>> keep;
        std::vector<qi_pauli_term> arr;
        for (auto& t : terms) { keep.push_back(t.term()); arr.push_back(keep.back()->t); }
        State o(s);
        check(qi_trotter_evolve(o.handle(), arr.data(), arr.size(), dt, steps, order));
        return o;
    }
};

inline SumOp heisenberg_1d(size_t n, double jx, double jy, double jz, double h, double mu) {   // models/heisenberg.rs:28-102
    if (n < 2) throw Error(QI_ERR_INVALID_NUMBER_OF_INPUTS, "InvalidNumberOfInputs", n, 2, "heisenberg_1d");
    std::vector<PauliString> terms;
    if (jx == 0.0 && jy == 0.0 && jz == 0.0 && h == 0.0) return SumOp(terms);
    for (size_t i = 0; i < n; i++) {
        size_t j = (i + 1) % n;
        if (jx != 0.0) terms.push_back(PauliString(cplx(-0.5 * jx, 0.0)).with_op(i, Pauli::X).with_op(j, Pauli::X));
        if (jy != 0.0) terms.push_back(PauliString(cplx(-0.5 * jy, 0.0)).with_op(i, Pauli::Y).with_op(j, Pauli::Y));
        if (jz != 0.0) terms.push_back(PauliString(cplx(-0.5 * jz, 0.0)).with_op(i, Pauli::Z).with_op(j, Pauli::Z));
        if (h != 0.0) terms.push_back(PauliString(cplx(-mu * (-0.5 * h), 0.0)).with_op(i, Pauli::Z));   // code is ground truth, heisenberg.rs:49
    }
    return SumOp(terms);
}

// models/heisenberg.rs:122-220: site (r, c) -> qubit r*m_cols + c, periodic; per site: field Z, the vertical bond
// (XX, YY, ZZ), the horizontal bond (XX, YY, ZZ); couplings -J/2, field coefficient mu * (-h/2) (line 143)
inline SumOp heisenberg_2d(size_t n_rows, size_t m_cols, double jx, double jy, double jz, double h_field, double mu) {
    if (n_rows < 2) throw Error(QI_ERR_INVALID_NUMBER_OF_INPUTS, "InvalidNumberOfInputs", n_rows, 2, "heisenberg_2d");
    if (m_cols < 2) throw Error(QI_ERR_INVALID_NUMBER_OF_INPUTS, "InvalidNumberOfInputs", m_cols, 2, "heisenberg_2d");
    std::vector<PauliString> terms;
    if (jx == 0.0 && jy == 0.0 && jz == 0.0 && h_field == 0.0) return SumOp(terms);
    const double js[3] = {jx, jy, jz};
    const Pauli ps[3] = {Pauli::X, Pauli::Y, Pauli::Z};
    for (size_t site = 0; site < n_rows * m_cols; site++) {
        const size_t r = site / m_cols, c = site % m_cols;
        if (h_field != 0.0) terms.push_back(PauliString(cplx(mu * (-0.5 * h_field), 0.0)).with_op(site, Pauli::Z));
        const size_t nb[2] = {((r + 1) % n_rows) * m_cols + c, r * m_cols + (c + 1) % m_cols};
        for (size_t d = 0; d < 2; d++)
            for (int k = 0; k < 3; k++)
                if (js[k] != 0.0) terms.push_back(PauliString(cplx(-0.5 * js[k], 0.0)).with_op(site, ps[k]).with_op(nb[d], ps[k]));
    }
    return SumOp(terms);
}

// models/ising.rs:27-75: H = -sum_i J_i Z_i Z_{i+1} - mu sum_i h_i Z_i, periodic; per site coupling then field
inline SumOp ising_1d(const std::vector<double>& h, const std::vector<double>& j, double mu) {
    const size_t n = h.size();
    if (j.size() != n) throw Error(QI_ERR_MISMATCHED_NUMBER_OF_PARAMETERS, "MismatchedNumberOfParameters", n, j.size(), "ising_1d");
    if (n < 2) throw Error(QI_ERR_INVALID_NUMBER_OF_INPUTS, "InvalidNumberOfInputs", n, 2, "ising_1d");
    std::vector<PauliString> terms;
    bool all_zero = true;
    for (size_t i = 0; i < n; i++) all_zero = all_zero && h[i] == 0.0 && j[i] == 0.0;
    if (all_zero) return SumOp(terms);
    for (size_t i = 0; i < n; i++) {
        if (j[i] != 0.0) terms.push_back(PauliString(cplx(j[i] * -1.0, 0.0)).with_op(i, Pauli::Z).with_op((i + 1) % n, Pauli::Z));
        if (h[i] != 0.0) terms.push_back(PauliString(cplx(-1.0 * mu * h[i], 0.0)).with_op(i, Pauli::Z));
    }
    return SumOp(terms);
}
inline SumOp ising_1d_uniform(size_t n, double h, double j, double mu) {   // ising.rs:90-139
    if (n < 2) throw Error(QI_ERR_INVALID_NUMBER_OF_INPUTS, "InvalidNumberOfInputs", n, 2, "ising_1d_uniform");
    return ising_1d(std::vector<double>(n, h), std::vector<double>(n, j), mu);
}
// models/ising.rs:161-244: h[r][c] field, jv[r][c] / jh[r][c] couplings to ((r+1)%N, c) / (r, (c+1)%M); per site field, vertical, horizontal
inline SumOp ising_2d(const std::vector<std::vector<double>>& h, const std::vector<std::vector<double>>& jv,
                      const std::vector<std::vector<double>>& jh, double mu) {
    const size_t n = h.size(), m = n ? h[0].size() : 0;
    if (n < 2) throw Error(QI_ERR_INVALID_NUMBER_OF_INPUTS, "InvalidNumberOfInputs", n, 2, "ising_2d");
    if (m < 2) throw Error(QI_ERR_INVALID_NUMBER_OF_INPUTS, "InvalidNumberOfInputs", m, 2, "ising_2d");
    std::vector<PauliString> terms;
    bool all_zero = true;
    for (size_t r = 0; r < n; r++)
        for (size_t c = 0; c < m; c++) all_zero = all_zero && h[r][c] == 0.0 && jv[r][c] == 0.0 && jh[r][c] == 0.0;
    if (all_zero) return SumOp(terms);
    for (size_t site = 0; site < n * m; site++) {
        const size_t r = site / m, c = site % m;
        if (h[r][c] != 0.0) terms.push_back(PauliString(cplx(-1.0 * mu * h[r][c], 0.0)).with_op(site, Pauli::Z));
        if (jv[r][c] != 0.0) terms.push_back(PauliString(cplx(jv[r][c] * -1.0, 0.0)).with_op(site, Pauli::Z).with_op(((r + 1) % n) * m + c, Pauli::Z));
        if (jh[r][c] != 0.0) terms.push_back(PauliString(cplx(jh[r][c] * -1.0, 0.0)).with_op(site, Pauli::Z).with_op(r * m + (c + 1) % m, Pauli::Z));
    }
    return SumOp(terms);
}
inline SumOp ising_2d_uniform(size_t n, size_t m, double h, double j, double mu) {   // ising.rs:259-324
    if (n < 2) throw Error(QI_ERR_INVALID_NUMBER_OF_INPUTS, "InvalidNumberOfInputs", n, 2, "ising_2d_uniform");
    if (m < 2) throw Error(QI_ERR_INVALID_NUMBER_OF_INPUTS, "InvalidNumberOfInputs", m, 2, "ising_2d_uniform");
    std::vector<std::vector<double>> hh(n, std::vector<double>(m, h)), jj(n, std::vector<double>(m, j));
    return ising_2d(hh, jj, jj, mu);
}

// components/parametric/parameter.rs:13-74: a shared, mutable cell of N values.  Copying shares the cell (Rust's Arc
// clone); deep_clone copies it.
template <size_t N>
class Parameter {
    struct Cell { std::array<double, N> v; std::mutex mu; };
    std::shared_ptr<Cell> cell_;

public:
    explicit Parameter(const std::array<double, N>& initial) : cell_(std::make_shared<Cell>()) { cell_->v = initial; }
    Parameter clone() const { return *this; }
    Parameter deep_clone() const { return Parameter(get()); }
    std::array<double, N> get() const { std::lock_guard<std::mutex> lk(cell_->mu); return cell_->v; }
    void set(const std::array<double, N>& v) { std::lock_guard<std::mutex> lk(cell_->mu); cell_->v = v; }
};
// components/parametric/parametric_gate.rs:6-24: resolved to concrete operator gates every time it is applied
struct Gate;
struct ParametricGate {
    virtual ~ParametricGate() = default;
    virtual std::vector<Gate> to_concrete_gates(const std::vector<size_t>& targets, const std::vector<size_t>& controls) const = 0;
};

// gate.rs:13-52, circuit.rs:27-202, circuit.rs:288-1742, subroutine.rs:90-160
struct Gate {
    enum class Kind { Operator, Measurement, PauliString, PauliTimeEvolution, Parametric };
    Kind kind = Kind::Operator;
    std::shared_ptr<Operator> op;                    // Operator
    std::vector<size_t> targets, controls;           // Operator: targets / controls; Measurement: measured qubits in `targets`
    Basis basis;                                     // Measurement (gate.rs:26)
    std::shared_ptr<PauliString> pauli_string;       // PauliString / PauliTimeEvolution (gate.rs:38, 46)
    double time = 0.0;                               // PauliTimeEvolution
    std::shared_ptr<ParametricGate> p_gate;          // Parametric (gate.rs:30): targets / controls as for Operator
    Gate() = default;
    Gate(std::shared_ptr<Operator> o, std::vector<size_t> t, std::vector<size_t> c) : op(std::move(o)), targets(std::move(t)), controls(std::move(c)) {}
    static Gate measurement(const Basis& b, std::vector<size_t> qubits) { Gate g; g.kind = Kind::Measurement; g.basis = b; g.targets = std::move(qubits); return g; }
    static Gate pauli_string_gate(const PauliString& ps) { Gate g; g.kind = Kind::PauliString; g.pauli_string = std::make_shared<PauliString>(ps); for (auto& kv : ps.ops()) g.targets.push_back(kv.first); return g; }
    static Gate pauli_time_evolution(const PauliString& ps, double t) { Gate g = pauli_string_gate(ps); g.kind = Kind::PauliTimeEvolution; g.time = t; return g; }
    static Gate parametric(std::shared_ptr<ParametricGate> pg, std::vector<size_t> t, std::vector<size_t> c) {
        Gate g; g.kind = Kind::Parametric; g.p_gate = std::move(pg); g.targets = std::move(t); g.controls = std::move(c); return g;
    }
    // Circuit::to_concrete_circuit's per-gate rule (circuit.rs:205-217)
    std::vector<Gate> concrete() const {
        if (kind == Kind::Parametric) return p_gate->to_concrete_gates(targets, controls);
        if (kind == Kind::PauliString) return pauli_string->to_gates();
        return {*this};
    }
    // Gate::apply (gate.rs:99-122), in place; `seed` feeds a Measurement gate's draw (shared-seed contract)
    void apply_(State& s, uint64_t seed = 0) const {
        switch (kind) {
            case Kind::Operator: s.apply_(*op, targets, controls); break;
            case Kind::Measurement: s = s.measure(basis, targets, seed).second; break;
            case Kind::PauliString: s = pauli_string->apply_normalised(s); break;     // coefficient dropped, gate.rs:115-117
            case Kind::PauliTimeEvolution: s = pauli_string->apply_exp_neg_i_dt(s, time); break;
            case Kind::Parametric: for (const Gate& g : p_gate->to_concrete_gates(targets, controls)) g.apply_(s, seed); break;   // gate.rs:107-114
        }
    }
};
inline std::vector<Gate> PauliString::to_gates() const {
    std::vector<Gate> out;
    for (auto& kv : ops_) {
        std::shared_ptr<Operator> op;
        if (kv.second == Pauli::X) op = std::make_shared<PauliX>(); else if (kv.second == Pauli::Y) op = std::make_shared<PauliY>(); else op = std::make_shared<PauliZ>();
        out.push_back(Gate{op, {kv.first}, {}});
    }
    return out;
}
// parametric_gate.rs:35-211: one concrete gate per target (the matchgate: its single target)
template <class MakeOp, size_t N>
struct ParametricOf : ParametricGate {
    Parameter<N> parameter;
    explicit ParametricOf(Parameter<N> p) : parameter(std::move(p)) {}
    std::vector<Gate> to_concrete_gates(const std::vector<size_t>& targets, const std::vector<size_t>& controls) const override {
        auto sp = MakeOp::make(parameter.get());
        std::vector<Gate> out;
        for (size_t t : targets) out.push_back(Gate{sp, {t}, controls});
        return out;
    }
};
struct MakeRyPhase { static std::shared_ptr<Operator> make(const std::array<double, 2>& v) { return std::make_shared<Unitary2>(Unitary2::from_ry_phase(v[0], v[1])); } };
struct MakeRyPhaseDag { static std::shared_ptr<Operator> make(const std::array<double, 2>& v) { return std::make_shared<Unitary2>(Unitary2::from_ry_phase_dagger(v[0], v[1])); } };
struct MakeMatchgate { static std::shared_ptr<Operator> make(const std::array<double, 3>& v) { return std::make_shared<Matchgate>(v[0], v[1], v[2]); } };
struct MakeRx { static std::shared_ptr<Operator> make(const std::array<double, 1>& v) { return std::make_shared<RotateX>(v[0]); } };
struct MakeRy { static std::shared_ptr<Operator> make(const std::array<double, 1>& v) { return std::make_shared<RotateY>(v[0]); } };
struct MakeRz { static std::shared_ptr<Operator> make(const std::array<double, 1>& v) { return std::make_shared<RotateZ>(v[0]); } };
struct MakeP { static std::shared_ptr<Operator> make(const std::array<double, 1>& v) { return std::make_shared<PhaseShift>(v[0]); } };
using ParametricRyPhase = ParametricOf<MakeRyPhase, 2>;        // parametric_gate.rs:35-52
using ParametricRyPhaseDag = ParametricOf<MakeRyPhaseDag, 2>;  // 63-80
using ParametricMatchgate = ParametricOf<MakeMatchgate, 3>;    // 92-110
using ParametricRx = ParametricOf<MakeRx, 1>;                  // 120-136
using ParametricRy = ParametricOf<MakeRy, 1>;                  // 146-162
using ParametricRz = ParametricOf<MakeRz, 1>;                  // 172-188
using ParametricP = ParametricOf<MakeP, 1>;                    // 198-211
struct Subroutine {
    std::vector<Gate> gates;
    size_t num_qubits;
    static Subroutine qft(const std::vector<size_t>& qubits, size_t num_qubits);
    static Subroutine iqft(const std::vector<size_t>& qubits, size_t num_qubits);
};
class Circuit {
    static void validate(const Gate& g, size_t n) {   // circuit.rs:35-52
        for (size_t q : g.targets) if (q >= n) throw Error(QI_ERR_INVALID_QUBIT_INDEX, "InvalidQubitIndex", q, n, "circuit");
        for (size_t q : g.controls) if (q >= n) throw Error(QI_ERR_INVALID_QUBIT_INDEX, "InvalidQubitIndex", q, n, "circuit");
    }

public:
    std::vector<Gate> gates;
    size_t num_qubits;
    explicit Circuit(size_t n) : num_qubits(n) {}
    static Circuit with_gates(std::vector<Gate> gs, size_t n) { Circuit c(n); for (auto& g : gs) validate(g, n); c.gates = std::move(gs); return c; }   // circuit.rs:66-82
    void add_gate(const Gate& g) { validate(g, num_qubits); gates.push_back(g); }                                       // circuit.rs:97-111
    size_t get_num_qubits() const { return num_qubits; }
    const std::vector<Gate>& get_gates() const { return gates; }
    // Circuit::execute's loop (circuit.rs:160-172), in place: consecutive operator gates = ONE fused run (qi_apply_circuit),
    // consecutive PauliTimeEvolution gates = ONE fused exp sequence; measurement / PauliString gates run on their own.
    // Gate k of the circuit draws from the stream seeded `seed + k`.
    void execute_(State& s, uint64_t seed = 0) const {
        if (s.num_qubits() != num_qubits) throw Error(QI_ERR_INVALID_NUMBER_OF_QUBITS, "InvalidNumberOfQubits", s.num_qubits(), 0, "execute");
        bool has_parametric = false;
        for (auto& g : gates) has_parametric |= g.kind == Gate::Kind::Parametric;
        if (has_parametric) {      // resolved with the parameter values of THIS execution; the concrete gates join the fused runs
            Circuit flat(num_qubits);
            std::vector<size_t> origin;
            for (size_t k = 0; k < gates.size(); k++) {
                if (gates[k].kind != Gate::Kind::Parametric) { flat.gates.push_back(gates[k]); origin.push_back(k); continue; }
                for (const Gate& g : gates[k].p_gate->to_concrete_gates(gates[k].targets, gates[k].controls)) { flat.gates.push_back(g); origin.push_back(k); }
            }
            flat.execute_runs(s, seed, &origin);
            return;
        }
        execute_runs(s, seed, nullptr);
    }

private:
    void execute_runs(State& s, uint64_t seed, const std::vector<size_t>* origin) const {
        size_t i = 0;
        while (i < gates.size()) {
            const Gate::Kind k = gates[i].kind;
            size_t j = i;
            while (j < gates.size() && gates[j].kind == k && (k == Gate::Kind::Operator || k == Gate::Kind::PauliTimeEvolution)) j++;
            if (k == Gate::Kind::Operator) {
                std::vector<GateRecord> recs;
                recs.reserve(j - i);
                for (size_t g = i; g < j; g++) recs.push_back(make_record(*gates[g].op, gates[g].targets, gates[g].controls));
                std::vector<qi_gate> arr;
                for (auto& r : recs) { r.g.controls = r.controls.data(); arr.push_back(r.g); }
                check(qi_apply_circuit(s.handle(), arr.data(), arr.size()));
            } else if (k == Gate::Kind::PauliTimeEvolution) {
                std::vector<std::unique_ptr<PauliString::Term>> keep;
                std::vector<qi_pauli_term> arr;
                std::vector<double> f;
                for (size_t g = i; g < j; g++) {
                    if (gates[g].pauli_string->coefficient().imag() != 0.0)      // gate.rs:116-118 -> pauli_string.rs:281-284
                        throw Error(QI_ERR_INVALID_PAULI_STRING_COEFFICIENT, "InvalidPauliStringCoefficient", 0, 0, "imaginary coefficient");
                    keep.push_back(gates[g].pauli_string->term());
                    arr.push_back(keep.back()->t);
                    f.push_back(0.0);
                    f.push_back(-gates[g].time);
                }
                check(qi_apply_pauli_exp_sequence(s.handle(), arr.data(), arr.size(), f.data()));
            } else {
                gates[i].apply_(s, seed + (origin ? (*origin)[i] : i));
                j = i + 1;
            }
            i = j;
        }
    }

public:
    Circuit to_concrete_circuit() const {   // circuit.rs:204-221
        Circuit c(num_qubits);
        for (auto& g : gates) for (const Gate& cg : g.concrete()) c.gates.push_back(cg);
        return c;
    }
    State execute(const State& initial, uint64_t seed = 0) const { State s(initial); execute_(s, seed); return s; }   // circuit.rs:160-172
    std::vector<State> trace_execution(const State& initial, uint64_t seed = 0) const {   // circuit.rs:188-202
        if (initial.num_qubits() != num_qubits) throw Error(QI_ERR_INVALID_NUMBER_OF_QUBITS, "InvalidNumberOfQubits", initial.num_qubits(), 0, "trace_execution");
        std::vector<State> out;
        State cur(initial);
        out.push_back(cur);
        for (size_t k = 0; k < gates.size(); k++) { gates[k].apply_(cur, seed + k); out.push_back(cur); }
        return out;
    }
    // Circuit::execute for a HOST-resident state vector (state.rs:74-81): in -> circuit -> out with `work` as the device
    // buffer; the two PCIe copies overlap the circuit (qi_execute_host).  `out` may alias `in`.  Operator gates only.
    void execute_host_(State& work, const std::complex<double>* in, std::complex<double>* out, size_t len) const {
        if (work.num_qubits() != num_qubits) throw Error(QI_ERR_INVALID_NUMBER_OF_QUBITS, "InvalidNumberOfQubits", work.num_qubits(), 0, "execute_host");
        std::vector<GateRecord> recs;
        recs.reserve(gates.size());
        for (auto& g : gates) {
            if (g.kind != Gate::Kind::Operator) throw Error(QI_ERR_INVALID_ARGUMENT, "InvalidArgument", 0, 0, "execute_host_ takes operator gates only");
            recs.push_back(make_record(*g.op, g.targets, g.controls));
        }
        std::vector<qi_gate> arr;
        for (auto& r : recs) { r.g.controls = r.controls.data(); arr.push_back(r.g); }
        check(qi_execute_host(work.handle(), arr.data(), arr.size(), reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out), len));
    }
};
class CircuitBuilder {
    std::vector<Gate> gates_;
    size_t n_;
    using Q = std::vector<size_t>;
    template <class Op> CircuitBuilder& each(Op op, const Q& ts, const Q& cs = {}) {
        auto sp = std::make_shared<Op>(op);
        for (size_t q : ts) gates_.push_back(Gate{sp, {q}, cs});
        return *this;
    }

public:
    explicit CircuitBuilder(size_t n) : n_(n) {}
    CircuitBuilder& add_gate(const Gate& g) { gates_.push_back(g); return *this; }                     // circuit.rs:311-314
    CircuitBuilder& add_gates(const std::vector<Gate>& gs) { gates_.insert(gates_.end(), gs.begin(), gs.end()); return *this; }
    // <g>_gate(qubit), <g>_gates(qubits), c<g>_gates(targets, controls)  (circuit.rs:378-870)
#define QI_BUILDER_FAMILY(NAME, OP)                                                              \
    CircuitBuilder& NAME##_gate(size_t q) { return each(OP(), {q}); }                             \
    CircuitBuilder& NAME##_gates(const Q& qs) { return each(OP(), qs); }                          \
    CircuitBuilder& c##NAME##_gates(const Q& ts, const Q& cs) { return each(OP(), ts, cs); }
    QI_BUILDER_FAMILY(h, Hadamard)
    QI_BUILDER_FAMILY(x, PauliX)
    QI_BUILDER_FAMILY(y, PauliY)
    QI_BUILDER_FAMILY(z, PauliZ)
    QI_BUILDER_FAMILY(s, PhaseS)
    QI_BUILDER_FAMILY(sdag, PhaseSdag)
    QI_BUILDER_FAMILY(t, PhaseT)
    QI_BUILDER_FAMILY(tdag, PhaseTdag)
#undef QI_BUILDER_FAMILY
    CircuitBuilder& id_gate(size_t q) { return each(Identity(), {q}); }
    CircuitBuilder& id_gates(const Q& qs) { return each(Identity(), qs); }
    CircuitBuilder& ci_gates(const Q& ts, const Q& cs) { return each(Identity(), ts, cs); }
#define QI_BUILDER_ANGLE(NAME, OP)                                                                         \
    CircuitBuilder& NAME##_gate(size_t q, double a) { return each(OP(a), {q}); }                            \
    CircuitBuilder& NAME##_gates(const Q& qs, double a) { return each(OP(a), qs); }                         \
    CircuitBuilder& c##NAME##_gates(const Q& ts, const Q& cs, double a) { return each(OP(a), ts, cs); }
    QI_BUILDER_ANGLE(p, PhaseShift)
    QI_BUILDER_ANGLE(rx, RotateX)
    QI_BUILDER_ANGLE(ry, RotateY)
    QI_BUILDER_ANGLE(rz, RotateZ)
#undef QI_BUILDER_ANGLE
    // circuit.rs:875-945: the unitarity check fires here (Result in the crate, exception here)
    CircuitBuilder& unitary_gate(size_t q, const cplx (&u)[2][2]) { return each(Unitary2::make(u), {q}); }
    CircuitBuilder& unitary_gates(const Q& qs, const cplx (&u)[2][2]) { return each(Unitary2::make(u), qs); }
    CircuitBuilder& cunitary_gates(const Q& ts, const Q& cs, const cplx (&u)[2][2]) { return each(Unitary2::make(u), ts, cs); }
    CircuitBuilder& ry_phase_gate(size_t q, double th, double ph) { return each(Unitary2::from_ry_phase(th, ph), {q}); }
    CircuitBuilder& ry_phase_gates(const Q& qs, double th, double ph) { return each(Unitary2::from_ry_phase(th, ph), qs); }
    CircuitBuilder& cry_phase_gates(const Q& ts, const Q& cs, double th, double ph) { return each(Unitary2::from_ry_phase(th, ph), ts, cs); }
    CircuitBuilder& ry_phase_dag_gate(size_t q, double th, double ph) { return each(Unitary2::from_ry_phase_dagger(th, ph), {q}); }
    CircuitBuilder& ry_phase_dag_gates(const Q& qs, double th, double ph) { return each(Unitary2::from_ry_phase_dagger(th, ph), qs); }
    CircuitBuilder& cry_phase_dag_gates(const Q& ts, const Q& cs, double th, double ph) { return each(Unitary2::from_ry_phase_dagger(th, ph), ts, cs); }
    CircuitBuilder& cnot_gate(size_t target, size_t control) { gates_.push_back(Gate{std::make_shared<CNOT>(), {target}, {control}}); return *this; }   // circuit.rs:1071
    CircuitBuilder& swap_gate(size_t a, size_t b) { gates_.push_back(Gate{std::make_shared<SWAP>(), {a, b}, {}}); return *this; }
    CircuitBuilder& cswap_gate(size_t a, size_t b, const Q& cs) { gates_.push_back(Gate{std::make_shared<SWAP>(), {a, b}, cs}); return *this; }          // circuit.rs:1096
    CircuitBuilder& toffoli_gate(size_t c1, size_t c2, size_t t) { gates_.push_back(Gate{std::make_shared<Toffoli>(), {t}, {c1, c2}}); return *this; }   // circuit.rs:1118
    CircuitBuilder& matchgate(size_t q, double th, double p1, double p2) { gates_.push_back(Gate{std::make_shared<Matchgate>(th, p1, p2), {q}, {}}); return *this; }   // circuit.rs:1169
    CircuitBuilder& cmatchgate(size_t q, const Q& cs, double th, double p1, double p2) { gates_.push_back(Gate{std::make_shared<Matchgate>(th, p1, p2), {q}, cs}); return *this; }   // circuit.rs:1194: controls second
    CircuitBuilder& pauli_string_gate(const PauliString& ps) { gates_.push_back(Gate::pauli_string_gate(ps)); return *this; }                            // circuit.rs:1130
    CircuitBuilder& pauli_time_evolution_gate(const PauliString& ps, double t) { gates_.push_back(Gate::pauli_time_evolution(ps, t)); return *this; }
    template <class Op> CircuitBuilder& add_operator_gate(const Op& op, const Q& ts, const Q& cs = {}) { gates_.push_back(Gate{std::make_shared<Op>(op), ts, cs}); return *this; }   // circuit.rs:1215-1224
    // parametric adders (circuit.rs:1226-1735): one Parameter per target
    template <class PG, size_t N> CircuitBuilder& parametric_each(const Q& ts, const Q& cs, const std::vector<Parameter<N>>& ps) {
        if (ts.size() != ps.size()) throw Error(QI_ERR_MISMATCHED_NUMBER_OF_PARAMETERS, "MismatchedNumberOfParameters", ts.size(), ps.size(), "parametric gates");
        for (size_t k = 0; k < ts.size(); k++) gates_.push_back(Gate::parametric(std::make_shared<PG>(ps[k]), {ts[k]}, cs));
        return *this;
    }
#define QI_BUILDER_PARAMETRIC(NAME, PG, N)                                                                                                              \
    CircuitBuilder& parametric_##NAME##_gate(size_t t, const Parameter<N>& p) { return parametric_each<PG, N>({t}, {}, {p}); }                           \
    CircuitBuilder& parametric_##NAME##_gates(const Q& ts, const std::vector<Parameter<N>>& ps) { return parametric_each<PG, N>(ts, {}, ps); }           \
    CircuitBuilder& parametric_c##NAME##_gates(const Q& ts, const Q& cs, const std::vector<Parameter<N>>& ps) { return parametric_each<PG, N>(ts, cs, ps); }
    QI_BUILDER_PARAMETRIC(ry_phase, ParametricRyPhase, 2)
    QI_BUILDER_PARAMETRIC(ry_phase_dag, ParametricRyPhaseDag, 2)
    QI_BUILDER_PARAMETRIC(rx, ParametricRx, 1)
    QI_BUILDER_PARAMETRIC(ry, ParametricRy, 1)
    QI_BUILDER_PARAMETRIC(rz, ParametricRz, 1)
    QI_BUILDER_PARAMETRIC(p, ParametricP, 1)
#undef QI_BUILDER_PARAMETRIC
    CircuitBuilder& parametric_matchgate(size_t t, const Parameter<3>& p) { gates_.push_back(Gate::parametric(std::make_shared<ParametricMatchgate>(p), {t}, {})); return *this; }
    CircuitBuilder& parametric_cmatchgate(size_t t, const Q& cs, const Parameter<3>& p) { gates_.push_back(Gate::parametric(std::make_shared<ParametricMatchgate>(p), {t}, cs)); return *this; }
    CircuitBuilder& measure_gate(const Basis& basis, const Q& qubits) { gates_.push_back(Gate::measurement(basis, qubits)); return *this; }              // circuit.rs:1737
    CircuitBuilder& add_subroutine(const Subroutine& s) { gates_.insert(gates_.end(), s.gates.begin(), s.gates.end()); return *this; }
    Subroutine build_subroutine() { Subroutine s{gates_, n_}; gates_.clear(); return s; }
    Circuit build() const { return Circuit::with_gates(gates_, n_); }   // circuit.rs:340-343 (validation: circuit.rs:35-52)
    Circuit build_final() { Circuit c = Circuit::with_gates(gates_, n_); gates_.clear(); return c; }   // circuit.rs:352-358
};
inline Subroutine Subroutine::qft(const std::vector<size_t>& q, size_t num_qubits) {   // subroutine.rs:90-112
    CircuitBuilder b(num_qubits);
    const size_t n = q.size();
    for (size_t i = 0; i < n; i++) {
        b.h_gate(q[i]);
        double den = 2.0;
        for (size_t k = 1; k < n - i; k++) { b.cp_gates({q[i]}, {q[i + k]}, M_PI / den); den *= 2.0; }
    }
    for (size_t i = 0; i < n / 2; i++) b.swap_gate(q[i], q[n - 1 - i]);
    return b.build_subroutine();
}
inline Subroutine Subroutine::iqft(const std::vector<size_t>& q, size_t num_qubits) {   // subroutine.rs:125-160
    CircuitBuilder b(num_qubits);
    const size_t n = q.size();
    for (size_t i = 0; i < n / 2; i++) b.swap_gate(q[i], q[n - 1 - i]);
    for (size_t ii = n; ii-- > 0;) {
        if (n > ii + 1) {
            size_t k_initial = (n - 1) - ii;
            double den = std::pow(2.0, (double)k_initial);
            for (size_t it = 0; it < k_initial; it++) {
                size_t k = k_initial - it;
                b.cp_gates({q[ii]}, {q[ii + k]}, -M_PI / den);
                if (k > 1) den /= 2.0;
            }
        }
        b.h_gate(q[ii]);
    }
    return b.build_subroutine();
}

}  // namespace quant_iron
