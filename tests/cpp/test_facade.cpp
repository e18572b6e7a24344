// C++ facade test: a handful of the reference's known-answer tests (src/tests/*.rs, cited) plus the QFT
// closed form, run through include/quant_iron_b200.hpp -> C ABI -> sm_100a kernels.  Needs a GPU.
#include <cstdio>
#include <cstdlib>

#include "quant_iron_b200.hpp"

using namespace quant_iron;
static int failures = 0;
#define EXPECT(cond)                                                                 \
    do {                                                                             \
        if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

template <class F>
static bool throws(const char* variant, uint64_t p0, uint64_t p1, F f) {
    try { f(); } catch (const Error& e) { return e.variant == variant && e.payload[0] == p0 && e.payload[1] == p1; }
    return false;
}

int main(int argc, char** argv) {
    const double S2 = 1.0 / std::sqrt(2.0);
    // operator_tests.rs:22-39
    EXPECT(State::new_zero(1).h(0).approx_eq(State::new_plus(1)));
    EXPECT(State::new_basis_n(1, 1).h(0).approx_eq(State::new_minus(1)));
    EXPECT(State::new_zero(2).h_multi({0, 1}).approx_eq(State::new_plus(2)));
    // operator_tests.rs:1989-2018: cnot(control, target)
    EXPECT(State::new_basis_n(2, 1).cnot(0, 1).approx_eq(State::new_basis_n(2, 3)));
    EXPECT(State::new_basis_n(2, 3).cnot(0, 1).approx_eq(State::new_basis_n(2, 1)));
    // operator_tests.rs:2391-2430 toffoli(c1, c2, target)
    EXPECT(State::new_basis_n(3, 3).toffoli(0, 1, 2).approx_eq(State::new_basis_n(3, 7)));
    // operator_tests.rs:2073-2099 swap
    EXPECT(State::new_basis_n(2, 2).swap(0, 1).approx_eq(State::new_basis_n(2, 1)));
    // operator_tests.rs:2545-2566 error variants
    EXPECT(throws("InvalidQubitIndex", 2, 2, [] { State::new_zero(2).h(2); }));
    EXPECT(throws("OverlappingControlAndTargetQubits", 0, 0, [] { Matchgate(M_PI, 0, 0).apply(State::new_zero(3), {0}, {0}); }));
    EXPECT(throws("InvalidNumberOfQubits", 0, 0, [] { State::new_zero(0); }));
    // operate (operator_tests.rs:2493-2507)
    EXPECT(State::new_basis_n(2, 1).operate(CNOT(), {1}, {0}).approx_eq(State::new_basis_n(2, 3)));
    // pauli_string_tests.rs:408-438 golden [1/sqrt2, 0, -i/sqrt2, 0]
    PauliString ps = PauliString(cplx(0.5 * M_PI, 0.0)).with_op(0, Pauli::Z).with_op(1, Pauli::X);
    auto v = ps.apply_exp_neg_i_dt(State::new_zero(2), 0.5).state_vector();
    EXPECT(std::abs(v[0] - cplx(S2, 0)) < 1e-12 && std::abs(v[2] - cplx(0, -S2)) < 1e-12 && std::abs(v[1]) < 1e-12 && std::abs(v[3]) < 1e-12);
    // pauli_string_tests.rs:348-364: <11| 2X0 + 3Y1 + 4Z1 |11> = -4
    SumOp so({PauliString(2.0).with_op(0, Pauli::X), PauliString(3.0).with_op(1, Pauli::Y), PauliString(4.0).with_op(1, Pauli::Z)});
    EXPECT(std::abs(so.expectation_value(State::new_basis_n(2, 3)) - cplx(-4.0, 0.0)) < 1e-12);
    // circuit_tests.rs:84-95
    Circuit c = CircuitBuilder(2).h_gates({0, 1}).build();
    EXPECT(c.execute(State::new_zero(2)).approx_eq(State::new_plus(2)));
    EXPECT(throws("InvalidNumberOfQubits", 1, 0, [&] { c.execute(State::new_zero(1)); }));
    // builder argument order (circuit.rs:1071, 1118-1123)
    EXPECT(CircuitBuilder(3).x_gate(0).cnot_gate(1, 0).build().execute(State::new_zero(3)).approx_eq(State::new_basis_n(3, 3)));
    // Subroutine::qft closed forms (unpinned by the reference): QFT|+..+> = |0..0>, iqft.qft = 1
    const size_t n = 16;
    std::vector<size_t> qs(n);
    for (size_t i = 0; i < n; i++) qs[i] = i;
    State out = CircuitBuilder(n).add_subroutine(Subroutine::qft(qs, n)).build().execute(State::new_plus(n));
    EXPECT(std::abs(out.amplitude(0) - cplx(1.0, 0.0)) < 1e-12 && std::abs(out.amplitude(12345)) < 1e-12);
    State ghz = State::new_ghz(n);
    State back = CircuitBuilder(n).add_subroutine(Subroutine::qft(qs, n)).add_subroutine(Subroutine::iqft(qs, n)).build().execute(ghz);
    EXPECT(std::abs(back.inner_product(ghz) - cplx(1.0, 0.0)) < 1e-12);
    // measurement_tests.rs:4-30: outcome-conditional collapse of (|00>+|11>)/sqrt2
    for (uint64_t seed = 0; seed < 8; seed++) {
        State bell = State::from_vector({cplx(S2, 0), 0, 0, cplx(S2, 0)});
        auto o = bell.measure_(MeasurementBasis::Computational, {0}, seed);
        EXPECT(bell.approx_eq(State::new_basis_n(2, o[0] ? 3 : 0)));
    }
    // heisenberg + trotter run and keep the norm (time_evolution.rs:140-167)
    SumOp hs = heisenberg_1d(10, 1.0, 2.0, 3.0, 0.5, 0.1);
    EXPECT(hs.num_terms() == 40);
    State ev = hs.trotter_evolve(State::new_plus(10), 0.01, 5, 1);
    EXPECT(std::fabs(ev.norm_sqr() - 1.0) < 1e-12);
    EXPECT(std::fabs(hs.expectation_value(ev).imag()) < 1e-12);
    // one first-order step == the same terms as one fused apply_exp_factor sequence with factor -i*dt
    State one = hs.trotter_evolve(State::new_plus(10), 0.01, 1, 1);
    State seq = hs.apply_exp_sequence(State::new_plus(10), std::vector<cplx>(hs.num_terms(), cplx(0.0, -0.01)));
    EXPECT(seq.approx_eq(one));
    // Circuit::execute on a host-resident vector (qi_execute_host, pipelined in 8 chunks): same state as execute()
    // (run with --execute-host; kept apart from the established checks until it has run on hardware once)
    if (argc > 1 && std::string(argv[1]) == "--execute-host") {
        const size_t m = 12;
        CircuitBuilder b(m);
        for (size_t layer = 0; layer < 6; layer++) {
            for (size_t q = 0; q < m; q++) { if ((q + layer) % 3 == 0) b.h_gate(q); else if ((q + layer) % 3 == 1) b.rx_gate(q, 0.3 + 0.1 * q); else b.rz_gate(q, 0.7 - 0.05 * q); }
            for (size_t q = layer & 1; q + 1 < m; q += 2) b.cnot_gate(q + 1, q);
        }
        Circuit lc = b.build();
        State want = lc.execute(State::new_plus(m));
        std::vector<cplx> host = State::new_plus(m).state_vector();
        State work = State::new_zero(m);
        qi_set_option("host_min_qubits", 0);
        lc.execute_host_(work, host.data(), host.data(), host.size());
        qi_set_option("host_min_qubits", 26);
        EXPECT(State::from_vector(host).approx_eq(want) && work.approx_eq(want));
    }
    // the wider surface added late in round 1 (run with --surface until it has run on hardware once)
    if (argc > 1 && std::string(argv[1]) == "--surface") {
        using Q = std::vector<size_t>;
        // Bell constructors (state.rs:330-373) vs the circuit that prepares them; metrics (state.rs:384-498)
        State bell = State::new_zero(2).h(0).cnot(0, 1);
        EXPECT(bell.approx_eq(State::new_phi_plus()) && bell.equals_without_phase(State::new_phi_plus() * cplx(0.0, 1.0)));
        EXPECT(std::fabs(bell.fs_fidelity(State::new_phi_minus())) < 1e-12 && std::fabs(bell.fs_dist(State::new_phi_minus()) - M_PI / 2) < 1e-9);
        EXPECT(State::new_hartree_fock(2, 4).approx_eq(State::new_basis_n(4, 12)));
        // State families vs the builder's (state.rs:1019-2345 vs circuit.rs:378-1224)
        State p3 = State::new_plus(3);
        EXPECT(p3.cs_multi({0}, {1}).approx_eq(CircuitBuilder(3).cs_gates({0}, {1}).build().execute(p3)));
        EXPECT(p3.crx_multi({0, 2}, {1}, 0.7).approx_eq(CircuitBuilder(3).crx_gates({0, 2}, {1}, 0.7).build().execute(p3)));
        EXPECT(p3.ry_phase_dag(1, 0.7, 0.3).ry_phase(1, 0.7, 0.3).approx_eq(p3));
        EXPECT(p3.cswap(0, 1, {2}).approx_eq(CircuitBuilder(3).cswap_gate(0, 1, {2}).build().execute(p3)));
        EXPECT(p3.cmatchgate(0, 0.4, 0.2, 0.1, {2}).approx_eq(CircuitBuilder(3).cmatchgate(0, {2}, 0.4, 0.2, 0.1).build().execute(p3)));
        const cplx xm[2][2] = {{0.0, 1.0}, {1.0, 0.0}};
        EXPECT(State::new_zero(2).unitary(1, xm).approx_eq(State::new_basis_n(2, 2)));
        // SumOp::apply (pauli_string.rs:453-466): (X0 + Z0)|0> = |1> + |0>
        auto sv = SumOp({PauliString(1.0).with_op(0, Pauli::X), PauliString(1.0).with_op(0, Pauli::Z)}).apply(State::new_zero(1)).state_vector();
        EXPECT(std::abs(sv[0] - cplx(1.0, 0.0)) < 1e-12 && std::abs(sv[1] - cplx(1.0, 0.0)) < 1e-12);
        // a circuit with every gate kind (gate.rs:99-122): runs of operator gates, a PauliString gate (normalised, coefficient
        // dropped), time-evolution gates, a measurement
        PauliString zx = PauliString(cplx(0.5 * M_PI, 0.0)).with_op(0, Pauli::Z).with_op(1, Pauli::X);
        Circuit mixed = CircuitBuilder(2).h_gate(0).pauli_time_evolution_gate(zx, 0.5).pauli_string_gate(PauliString(3.0).with_op(0, Pauli::Z))
                            .h_gate(0).measure_gate(MeasurementBasis::Computational, {0}).build();
        State direct = PauliString(3.0).with_op(0, Pauli::Z).apply_normalised(zx.apply_exp_neg_i_dt(State::new_zero(2).h(0), 0.5)).h(0);
        auto trace = mixed.trace_execution(State::new_zero(2), 7);
        EXPECT(trace.size() == 6 && trace[4].approx_eq(direct));
        State fin = mixed.execute(State::new_zero(2), 7);
        EXPECT(std::fabs(fin.norm_sqr() - 1.0) < 1e-12 && fin.approx_eq(trace[5]));
        // measure / measure_n with the crate's return shape (state.rs:525-784): Bell outcomes are perfectly correlated
        auto shots = State::new_phi_plus().measure_n(MeasurementBasis::Computational, Q{}, 16, 20260003);
        EXPECT(shots.size() == 16);
        for (auto& sh : shots) EXPECT(sh.first.outcomes.size() == 2 && sh.first.outcomes[0] == sh.first.outcomes[1] &&
                                      sh.second.approx_eq(State::new_basis_n(2, sh.first.outcomes[0] ? 3 : 0)));
        EXPECT(throws("InvalidNumberOfMeasurements", 0, 0, [] { State::new_zero(1).measure_n(MeasurementBasis::X, {0}, 0, 1); }));
        // measuring |0> in the custom basis U = H is measuring |+>... the outcome state is U^dagger |b> (state.rs:708-720)
        const double r = 1.0 / std::sqrt(2.0);
        const cplx hm[2][2] = {{r, r}, {r, -r}};
        auto mr = State::new_zero(1).measure(Basis::custom(hm), {0}, 5);
        EXPECT(std::fabs(mr.second.norm_sqr() - 1.0) < 1e-12 && std::fabs(std::abs(mr.second.amplitude(0)) - r) < 1e-12);
        // a parametric circuit re-resolves at every execution (gate.rs:107-114; parametric_tests.rs:238-283)
        Parameter<1> ang({0.0});
        Circuit prx = CircuitBuilder(2).parametric_rx_gate(0, ang).cnot_gate(1, 0).build();
        EXPECT(prx.execute(State::new_zero(2)).approx_eq(State::new_zero(2)));
        ang.set({M_PI});
        EXPECT(prx.execute(State::new_zero(2)).approx_eq(State::new_zero(2).rx(0, M_PI).cnot(0, 1)));
        auto pr = State::new_phi_plus().probabilities({0, 1});
        EXPECT(pr.size() == 4 && std::fabs(pr[0] - 0.5) < 1e-12 && std::fabs(pr[3] - 0.5) < 1e-12 && pr[1] < 1e-12);
    }
    std::printf(failures ? "C++ facade: %d FAILURES\n" : "C++ facade: ALL PASS\n", failures);
    return failures ? 1 : 0;
}
